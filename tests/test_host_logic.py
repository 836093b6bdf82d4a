"""CPU tests of the host-side Python that surrounds the CUDA path (message text, records, sharding, generator)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden
from pyft8_b200 import messages, synth
from pyft8_b200.engine import bits91_to_int, int_to_bits91
from pyft8_b200.sharding import gather_records, shard_range
from pyft8_b200 import _lib as L


def test_unpack_matches_reference_on_random_payloads():
    c = load_golden("codec.npz")
    for h, acc, txt in zip(c["payload_hex"], c["accepted"], c["text"]):
        messages.call_hashes.clear()
        r = messages.unpack(int(h, 16))
        assert (r is not None) == bool(acc), h
        if acc:
            assert "|".join(r) == txt


def test_hash_history_resolves_brackets():
    messages.call_hashes.clear()
    messages.add_call_hashes("G1OJS")
    h22 = messages.hashes_for_calls["G1OJS"][2][0]
    payload = ((2063592 + h22) << 49) | (synth.pack_call28("EA6VQ")[0] << 20) | (32403 << 3) | 1
    assert messages.unpack(payload) == ("<G1OJS>", "EA6VQ", "RR73")
    messages.call_hashes.clear()
    assert messages.unpack(payload) == ("<...>", "EA6VQ", "RR73")


def test_bits91_roundtrip():
    rng = np.random.default_rng(3)
    for _ in range(50):
        v = int.from_bytes(rng.bytes(12), "big") >> 5
        assert bits91_to_int(int_to_bits91(v)) == v


def test_pack_encode_roundtrip_through_unpack():
    rng = np.random.default_rng(5)
    for _ in range(200):
        m = synth.random_message(rng)
        b = synth.pack77(*m)
        assert messages.unpack(b) == m
        cw = synth.codeword_bits(b)
        assert len(cw) == 174
        # every parity check of the decoder graph is satisfied by the encoder's codeword
        from pyft8_b200.tables import CHECK_VARS
        for row in CHECK_VARS:
            assert sum(cw[v] for v in row if v >= 0) % 2 == 0


def test_record_to_message_format():
    from pyft8_b200.receiver import record_to_message
    b77 = synth.pack77("CQ", "G1OJS", "IO90")
    r = np.zeros(1, L.RECORD_DTYPE)[0]
    r["bits91"] = int_to_bits91((b77 << 14) | synth.crc14(b77))
    r["f0_idx"], r["h0_idx"], r["snr"], r["ipass"], r["ap"], r["method"], r["ttweak"], r["ftweak"] = 400, 15, -7, 4, 1, 2, -6, -32
    m = record_to_message(r, "240101_000000")
    assert m["msg_tuple"] == ("CQ", "G1OJS", "IO90")
    assert m["decode_notes"] == "fine_CQ_LDPC20 t:-06 f:-32"
    assert m["their_snr"] == "-07"
    assert abs(m["tsec"] - (15 / 25 - 0.03)) < 1e-12 and abs(m["fHz"] - 1248.0) < 1e-12
    assert m["all_txt_format"] == "240101_000000 -07 -0.0 1248 ~ CQ G1OJS IO90"
    r["ipass"], r["method"], r["ap"] = 0, 0, 0
    assert record_to_message(r)["decode_notes"] == "grid_NoAP_GOOD91 t:+00 f:+00"


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 100000):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
from pyft8_b200 import _lib as L
from pyft8_b200.sharding import shard_range, gather_records
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
lo, hi = shard_range(5, dist.get_rank(), 2)
rec = np.zeros(2 * (hi - lo), L.RECORD_DTYPE)
rec["cycle"] = np.repeat(np.arange(hi - lo), 2)
rec["cand"] = 10 * dist.get_rank() + np.arange(len(rec))
out = gather_records(rec, lo, dist)
if dist.get_rank() == 0:
    assert list(out["cycle"]) == [0, 0, 1, 1, 2, 2, 3, 3, 4, 4], list(out["cycle"])
    assert list(out["cand"][:6]) == [0, 1, 2, 3, 4, 5] and list(out["cand"][6:]) == [10, 11, 12, 13]
    print("GATHER_OK")
else:
    assert out is None
dist.destroy_process_group()
'''


def test_gather_records_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]


_SHM_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
from pyft8_b200 import _lib as L
from pyft8_b200.sharding import ShmRecordGather
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
r = dist.get_rank()
g = ShmRecordGather(dist, capacity=16, dtype=L.RECORD_DTYPE, tag="t" + sys.argv[2])
ok = True
for step in range(3):
    rec = np.zeros(3 + r + step, L.RECORD_DTYPE)
    rec["cycle"] = np.arange(len(rec)) % 2
    rec["cand"] = 100 * r + step
    if step == 1:                                  # in-place form: records written straight into the slot, offset in the header
        g.slot_array()[:len(rec)] = rec
        g.publish_inplace(len(rec), first_cycle=2 * r)
    else:
        g.publish(rec, first_cycle=2 * r)
    parts = g.collect()
    if r == 0:
        ok &= [len(p) for p in parts] == [3 + step, 4 + step]
        ok &= set(parts[1]["cycle"].tolist()) == {2, 3} and set(parts[0]["cycle"].tolist()) == {0, 1}
        ok &= int(parts[1]["cand"][0]) == 100 + step and int(parts[0]["cand"][0]) == step
    else:
        ok &= parts is None
g.close()
print("SHM_OK" if ok else "SHM_BAD")
dist.destroy_process_group()
'''


def test_shm_record_gather_world_size_2_gloo(tmp_path):
    """The N > 1 record gather of bench.py's e2e loop (shared-memory segments + a gloo barrier), two CPU processes."""
    script = tmp_path / "w.py"
    script.write_text(_SHM_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHM_OK" in outs[0] and "SHM_OK" in outs[1], outs


def test_gather_records_single_process():
    rec = np.zeros(3, L.RECORD_DTYPE)
    out = gather_records(rec, 7)
    assert list(out["cycle"]) == [7, 7, 7] and list(rec["cycle"]) == [0, 0, 0]


def test_unpack_many_equals_sequential_unpack_with_hash_history():
    rng = np.random.default_rng(21)
    pool = [synth.pack77(*synth.random_message(rng)) for _ in range(64)]
    messages.add_call_hashes("G1OJS")
    h22 = messages.hashes_for_calls["G1OJS"][2][0]
    hashed = ((2063592 + h22) << 49) | (synth.pack_call28("EA6VQ")[0] << 20) | (32403 << 3) | 1
    g1 = synth.pack77("CQ", "G1OJS", "IO90")
    seq = [hashed, pool[3], g1, hashed, pool[3], 0, 5, pool[7], g1, hashed] + [pool[i] for i in rng.integers(0, 64, 500)]
    messages.call_hashes.clear()
    messages._CALL_CACHE.clear()
    want = [messages.unpack(b) for b in seq]
    assert want[0] == ("<...>", "EA6VQ", "RR73") and want[3] == ("<G1OJS>", "EA6VQ", "RR73")      # history matters
    messages.call_hashes.clear()
    got = messages.unpack_many(seq)
    assert got == want
    messages.call_hashes.clear()                      # warm cache, cold hash table: side effects are still replayed
    assert messages.unpack_many(seq) == want


def test_unpack_many_survives_call_cache_eviction(monkeypatch):
    """the call cache used to be cleared in the middle of a batch, then read back -> KeyError."""
    rng = np.random.default_rng(5)
    seq = [synth.pack77(*synth.random_message(rng)) for _ in range(50)]
    messages.call_hashes.clear()
    messages._CALL_CACHE.clear()
    want = [messages.unpack(b) for b in seq]
    monkeypatch.setattr(messages, "_CALL_CACHE_MAX", 8)          # every batch starts by evicting the previous one
    for _ in range(3):
        messages.call_hashes.clear()
        assert messages.unpack_many(seq) == want
        assert len(messages.call_hashes) > 0                      # registrations were replayed from the batch-local map


def test_records_bits77_vectorised():
    from pyft8_b200.receiver import records_bits77
    rng = np.random.default_rng(4)
    vals = [int.from_bytes(rng.bytes(12), "big") >> 5 for _ in range(100)]
    rec = np.zeros(100, L.RECORD_DTYPE)
    for i, v in enumerate(vals):
        rec["bits91"][i] = int_to_bits91(v)
    assert records_bits77(rec) == [v >> 14 for v in vals]
    assert records_bits77(rec[:0]) == []


def _words_from_bits77(vals):
    w = np.zeros((len(vals), 3), np.uint32)
    for i, b in enumerate(vals):
        v = int(format(int(b), "077b")[::-1], 2)
        w[i] = (v & 0xFFFFFFFF, (v >> 32) & 0xFFFFFFFF, v >> 64)
    return w


def test_unpack_words_mixed_stream_equals_sequential_unpack():
    """Vectorised batch unpack: same tuples and same hash-table end state as calling unpack() one by
    one, on a stream mixing valid standard messages, random bits, hashed-call fields, type-4 messages and repeats."""
    import copy
    from pyft8_b200 import messages, synth
    rng = np.random.default_rng(11)
    pay = []
    for _ in range(3000):
        r = rng.random()
        if r < 0.6:
            pay.append(synth.pack77(*synth.random_message(rng)))
        elif r < 0.8:
            pay.append(int.from_bytes(rng.bytes(10), "big") >> 3)
        elif r < 0.9:
            b = synth.pack77(*synth.random_message(rng))
            n28 = messages.NTOKENS + int(rng.integers(0, messages.MAX22 - 1))
            sh = 49 if rng.random() < 0.5 else 20
            pay.append((b & ~(((1 << 28) - 1) << sh)) | (n28 << sh))
        else:
            pay.append(((int.from_bytes(rng.bytes(10), "big") >> 3) & ~7) | 4)
    pay += pay[:1000] + [0]
    messages.call_hashes.clear(); messages.hashes_for_calls.clear()
    want = [messages.unpack(b) for b in pay]
    ch, hc = dict(messages.call_hashes), copy.deepcopy(messages.hashes_for_calls)
    messages.call_hashes.clear(); messages.hashes_for_calls.clear()
    got = messages.unpack_words(_words_from_bits77(pay))
    assert got == want
    assert messages.call_hashes == ch and messages.hashes_for_calls == hc
    assert sum(m is not None for m in want) > 1500
    assert messages.unpack_words(np.zeros((0, 3), np.uint32)) == []


def test_format_records_matches_record_to_message():
    from pyft8_b200 import messages, synth, _lib as L
    from pyft8_b200.receiver import format_records, record_to_message
    rng = np.random.default_rng(3)
    vals = [synth.pack77(*synth.random_message(rng)) for _ in range(40)]
    vals += vals[:5]                                       # same text twice in one cycle -> dropped; in another cycle -> kept
    rec = np.zeros(len(vals), L.RECORD_DTYPE)
    rec["bits91"] = _words_from_bits77(vals)
    rec["cycle"] = np.r_[np.repeat(np.arange(4), 10), [0, 0, 3, 3, 3]]
    rec["ipass"] = rng.integers(0, 7, len(vals))
    rec["ap"] = rng.integers(0, 5, len(vals))
    rec["method"] = rng.integers(0, 5, len(vals))
    rec["h0_idx"] = rng.integers(-37, 87, len(vals))
    rec["f0_idx"] = rng.integers(32, 960, len(vals))
    rec["ttweak"] = rng.choice(np.arange(-8, 8, 2), len(vals))
    rec["ftweak"] = rng.choice(np.arange(-32, 33, 8), len(vals))
    rec["snr"] = rng.integers(-24, 25, len(vals))
    cs = ["c%d" % i for i in range(4)]
    mb = format_records(rec, cs)
    assert mb.keep.tolist() == [True] * 40 + [False, False, True, True, True]
    kept = np.flatnonzero(mb.keep)
    dicts = [record_to_message(rec[i], cs[int(rec[i]["cycle"])]) for i in kept]
    assert mb.lines() == [d["all_txt_format"] for d in dicts]
    assert [mb.tsec[i] for i in kept] == [d["tsec"] for d in dicts]
    assert [mb.fHz[i] for i in kept] == [d["fHz"] for d in dicts]
    assert mb.notes(kept) == [d["decode_notes"] for d in dicts]
    assert mb.per_cycle_counts(4).tolist() == [10, 10, 10, 13]


def test_receiver_bank_buffers_and_hands_over_whole_cycles():
    """ReceiverBank host logic with a stub decoder: ragged per-receiver feeding, spill into the next
    cycle, double buffering, message routing per receiver."""
    from pyft8_b200 import synth, _lib as L
    from pyft8_b200.bank import ReceiverBank, CYCLE_SAMPLES
    seen = []

    def stub(audio):
        seen.append(audio.copy())
        rec = np.zeros(audio.shape[0], L.RECORD_DTYPE)     # one message per receiver, call chosen by the first sample
        for r in range(audio.shape[0]):
            rec[r]["bits91"] = _words_from_bits77([synth.pack77("CQ", "K1ABC" if audio[r, 0] % 2 else "G1OJS", "IO90")])[0]
            rec[r]["cycle"] = r
        return rec
    msgs = []
    bank = ReceiverBank(3, msgs.append, bands=["20m", "40m", None], decoder=stub, clock=lambda: 45.0)
    rng = np.random.default_rng(0)
    stream = rng.integers(-3000, 3000, (3, 2 * CYCLE_SAMPLES + 1000)).astype(np.int16)
    sent = [0, 0, 0]
    closed = 0
    while min(sent) < stream.shape[1]:
        r = int(np.argmin(sent))                           # keep the receivers within one block of each other
        k = int(rng.integers(100, 5000))
        closed += bank.feed(r, stream[r, sent[r]:sent[r] + k])
        sent[r] = min(sent[r] + k, stream.shape[1])
    assert closed == 2
    got = [bank.results(timeout=10) for _ in range(2)]
    assert [g[0] for g in got] == [0, 1]
    assert np.array_equal(seen[0], stream[:, :CYCLE_SAMPLES])
    assert np.array_equal(seen[1], stream[:, CYCLE_SAMPLES:2 * CYCLE_SAMPLES])
    assert bank._pos.tolist() == [1000, 1000, 1000]
    assert [m["receiver"] for m in got[0][1]] == [0, 1, 2] and [m["band"] for m in got[0][1]] == ["20m", "40m", None]
    want = ["CQ K1ABC IO90" if stream[r, 0] % 2 else "CQ G1OJS IO90" for r in range(3)]
    assert [" ".join(m["msg_tuple"]) for m in got[0][1]] == want
    assert len(msgs) == 6
    with pytest.raises(TypeError):
        bank.feed(0, np.zeros(10, np.float32))
    assert bank.feed_all(np.zeros((3, CYCLE_SAMPLES - 1000), np.int16)) == 1
    bank.results(timeout=10)
    with pytest.raises(BufferError):                       # one receiver running two cycles ahead of the others
        bank.feed(1, np.zeros(2 * CYCLE_SAMPLES + 1, np.int16))
    bank.close()


def test_wav_round_trip_and_cycle_split(tmp_path):
    from pyft8_b200 import wav
    g = load_golden("cycle_test_08.npz")["audio"]
    rec = np.concatenate([g, g[:50000]])                     # 1 full cycle + a partial one
    p = str(tmp_path / "two.wav")
    wav.write_wav(p, rec)
    a = wav.read_wav(p)
    assert a.shape == (2, 180000) and a.dtype == np.int16
    assert np.array_equal(a[0], g) and np.array_equal(a[1, :50000], g[:50000]) and not a[1, 50000:].any()
    assert np.array_equal(wav.read_wav(p, start_sample=1000)[0, :1000], g[1000:2000])
    import wave
    with wave.open(str(tmp_path / "bad.wav"), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(8000); w.writeframes(b"\\0\\0" * 10)
    with pytest.raises(ValueError, match="12 kHz"):
        wav.read_wav(str(tmp_path / "bad.wav"))
