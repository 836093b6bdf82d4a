"""Generate tests/golden/*.npz by running the UNMODIFIED reference (authoring container only).

Usage:  python oracle/make_golden.py          (needs /root/reference; see oracle/ref_harness.py)

Files written (all consumed by tests/, never by the product):
  cycle_<name>.npz   one per cycle: reference waterfall rows (subset), full candidate list, payload
                     grids (subset), per-candidate trace (grid/fine sd+snr, tweaks, nsync, final pass,
                     decoded payload), emitted messages (text, notes, tsec, fHz, snr).
                     wav cycles also carry the int16 audio of the reference's own fixture WAVs
                     (tests/pipeline/test_08.wav, test_09.wav); synthetic cycles are re-made from
                     their seed by pyft8_b200.synth and carry an audio checksum.
  fec.npz            ldpc_decode / osd_012 outputs on seeded noisy codewords (BASELINE config 3 recipe)
                     and on LLR vectors captured inside the WAV decodes.
  codec.npz          unpack() acceptance + text on random payloads, CRC/encoder known answers,
                     transmitter waveform samples.
"""
import os
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_harness as rh  # noqa: E402
from pyft8_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ROW_SUBSET = np.array([0, 1, 2, 7, 8, 9, 50, 111, 200, 258, 374, 375])
AP_NAMES = ["NoAP", "CQ", "RR73", "73", "RRR"]


def capture_cycle(name, audio, store_audio):
    rx, dec, tx, db, tu = rh.load_reference()
    ok_bits = []
    real_unpack = dec.unpack

    def spy_unpack(bits):
        r = real_unpack(bits)
        if r is not None:
            ok_bits.append((int(bits), r))
        return r
    dec.unpack = spy_unpack
    fec_calls = []
    real_ldpc, real_osd = rx.ldpc_decode, rx.osd_012

    def spy_ldpc(llr, nc0, its):
        inp = np.array(llr, np.float32)
        r = real_ldpc(llr, nc0, its)
        fec_calls.append(("ldpc", inp, nc0, its, r[0] is not None, r[1], np.array(llr, np.float32), len(r[2]) == 174))
        return r

    def spy_osd(llr):
        inp = np.array(llr, np.float32)
        n0 = len(ok_bits)
        r = real_osd(llr)
        fec_calls.append(("osd", inp, ok_bits[-1][0] if (r is not None and len(ok_bits) > n0) else 0))
        return r
    rx.ldpc_decode, rx.osd_012 = spy_ldpc, spy_osd
    try:
        with tempfile.TemporaryDirectory() as d:
            o = rh.decode_cycle(audio, dump=True, workdir=d)
    finally:
        dec.unpack = real_unpack
        rx.ldpc_decode, rx.osd_012 = real_ldpc, real_osd
    text_to_bits = {}
    for b, t in ok_bits:
        text_to_bits.setdefault(tuple(t), b)
    n = o["n_cands"]
    tr = o["trace"]

    def col(key, default, dtype):
        return np.array([t.get(key, default) for t in tr], dtype)
    tt = np.zeros(n, np.int32)
    ff = np.zeros(n, np.int32)
    for i, t in enumerate(tr):
        if "tweaks" in t:
            s = t["tweaks"].split()
            tt[i], ff[i] = int(s[0][2:]), int(s[1][2:])
    dec_bits = np.zeros(n, object)
    for i, t in enumerate(tr):
        r = t.get("result")
        dec_bits[i] = text_to_bits[tuple(r)] if r not in (None, "stop") else 0
    msgs = o["messages"]
    d = dict(
        grid_rows=ROW_SUBSET, grid_subset=o["grid"][ROW_SUBSET],
        grid_crc=np.uint32(zlib.crc32(o["grid"].tobytes())),
        grid_col_mean=o["grid"].astype(np.float64).mean(axis=0),
        grid_row_mean=o["grid"].astype(np.float64).mean(axis=1),
        cand_f0=o["cand_f0"], cand_h0=o["cand_h0"], cand_score=o["cand_score"],
        payload_idx=np.arange(0, n, 7), payload_subset=o["cand_payload"][::7],
        grid_sd=col("grid_sd", np.nan, np.float64), grid_snr=col("grid_snr", -99, np.int32),
        has_fine=col("nsync", -1, np.int32) >= 0, nsync=col("nsync", -1, np.int32), tt=tt, ff=ff,
        fine_sd=col("fine_sd", np.nan, np.float64), fine_snr=col("fine_snr", -99, np.int32),
        final_ipass=col("final_ipass", -1, np.int32),
        dec_bits77_hex=np.array(["%x" % b for b in dec_bits]),
        dec_notes=np.array([t.get("notes", "") for t in tr]),
        dec_tsec=col("tsec", np.nan, np.float64), dec_fHz=col("fHz", np.nan, np.float64),
        dec_snr=col("snr", -99, np.int32),
        emit_order=np.array(o["emit_order"], np.int32),
        msg_text=np.array([" ".join(m["msg_tuple"]) for m in msgs]),
        msg_notes=np.array([m["decode_notes"] for m in msgs]),
        msg_tsec=np.array([m["tsec"] for m in msgs]), msg_fHz=np.array([m["fHz"] for m in msgs]),
        msg_snr=np.array([m["their_snr"] for m in msgs]),
        msg_bits77_hex=np.array(["%x" % text_to_bits[tuple(m["msg_tuple"])] for m in msgs]),
        audio_crc=np.uint32(zlib.crc32(audio.tobytes())),
    )
    # a few grid / fine LLR vectors and fine grids for stage-level checks
    sel = [i for i, t in enumerate(tr) if "fine_llr" in t][:12]
    d["fine_sel"] = np.array(sel, np.int32)
    d["fine_llr"] = np.array([tr[i]["fine_llr"] for i in sel], np.float32).reshape(len(sel), 174)
    d["fine_grid"] = np.array([tr[i]["fine_grid"] for i in sel], np.float32).reshape(len(sel), 79, 8)
    gsel = list(range(0, n, 9))
    d["grid_llr_sel"] = np.array(gsel, np.int32)
    d["grid_llr"] = np.array([tr[i]["grid_llr"] for i in gsel], np.float32)
    if store_audio:
        d["audio"] = audio
    np.savez_compressed(os.path.join(OUT, f"cycle_{name}.npz"), **d)
    print(f"cycle_{name}: {n} cands, {len(msgs)} msgs, {len(fec_calls)} fec calls")
    return fec_calls


def main():
    os.makedirs(OUT, exist_ok=True)
    rx, dec, tx, db, tu = rh.load_reference()
    all_fec = []
    for w in ("test_08", "test_09"):
        a = rh.read_wav_i16(os.path.join(rh.REF_ROOT, "tests", "pipeline", w + ".wav"))
        all_fec += capture_cycle(w, a, True)
    for name, seed, kw in (("syn20", 1000, dict(n_signals=20, snr_db=(-20, 5), f_hz=(200, 2950), dt_s=(-0.5, 1.0))),
                           ("syn50", 2000, dict(n_signals=50, snr_db=(-24, 10), f_hz=(200, 2950), dt_s=(-0.5, 1.0))),
                           ("syn120", 4000, dict(n_signals=120, snr_db=(-24, 10), f_hz=(200, 2950), dt_s=(-0.5, 1.0)))):
        a, _ = synth.make_cycle(seed, **kw)
        all_fec += capture_cycle(name, a, False)

    # ---- FEC: captured calls (subset) + seeded noisy codewords
    rng = np.random.default_rng(7)
    ld = [c for c in all_fec if c[0] == "ldpc"]
    od = [c for c in all_fec if c[0] == "osd"]
    ld = [ld[i] for i in sorted(rng.choice(len(ld), 400, replace=False))]
    od = [od[i] for i in sorted(rng.choice(len(od), min(160, len(od)), replace=False))]
    fec = dict(
        cap_ldpc_in=np.array([c[1] for c in ld]), cap_ldpc_nc0=np.array([c[2] for c in ld], np.int32),
        cap_ldpc_its=np.array([c[3] for c in ld], np.int32), cap_ldpc_ok=np.array([c[4] for c in ld]),
        cap_ldpc_nits=np.array([c[5] for c in ld], np.int32), cap_ldpc_out=np.array([c[6] for c in ld]),
        cap_ldpc_hasllr=np.array([c[7] for c in ld]),
        cap_osd_in=np.array([c[1] for c in od]), cap_osd_bits77_hex=np.array(["%x" % c[2] for c in od]),
    )
    ok_bits = []
    real_unpack = dec.unpack

    def spy(bits):
        r = real_unpack(bits)
        if r is not None:
            ok_bits.append(int(bits))
        return r
    dec.unpack = spy
    with tempfile.TemporaryDirectory() as d:
        cwd = os.getcwd()
        os.chdir(d)
        for e in range(5):
            llr, truth = synth.make_llr_codewords(3000 + e, 120, float(e))
            st, nits, outs, osdb = [], [], [], []
            for i in range(len(llr)):
                x = llr[i].copy()
                n0 = len(ok_bits)
                r = dec.ldpc_decode(x, 90, 20)
                st.append(1 if r[0] is not None else (2 if len(r[2]) == 174 else 0))
                nits.append(r[1])
                outs.append(x.copy())
                b = ok_bits[-1] if len(ok_bits) > n0 else 0
                if r[0] is None:
                    n0 = len(ok_bits)
                    r2 = dec.osd_012(llr[i].copy())
                    b = ok_bits[-1] if (r2 is not None and len(ok_bits) > n0) else 0
                    osdb.append("%x" % b)
                else:
                    osdb.append("-")
            fec[f"e{e}_status"] = np.array(st, np.int32)
            fec[f"e{e}_nits"] = np.array(nits, np.int32)
            fec[f"e{e}_llr_out"] = np.array(outs, np.float32)
            fec[f"e{e}_osd_bits77_hex"] = np.array(osdb)
            fec[f"e{e}_truth_hex"] = np.array(["%x" % t for t in truth])
            fec[f"e{e}_llr_crc"] = np.uint32(zlib.crc32(llr.tobytes()))
            print(f"Eb/N0 {e} dB: BP ok {sum(s == 1 for s in st)}/{len(st)}, OSD rescues {sum(x not in ('-', '0') for x in osdb)}")
        os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "fec.npz"), **fec)

    # ---- codec: unpack acceptance on random payloads, plus boundary cases
    rng = np.random.default_rng(11)
    pay = [int.from_bytes(rng.bytes(10), "big") >> 3 for _ in range(30000)]
    # bias towards standard messages with standard calls so both accept and reject branches are hit
    for i in range(0, 30000, 2):
        pay[i] = (pay[i] & ~7) | (1 + (i // 2) % 2)
    pay += [0, 1, 2, 4, (6257895 << 50) | (2 << 21) | (100 << 3) | 1, (2 << 50) | (6257896 << 21) | (32403 << 3) | 1]
    with tempfile.TemporaryDirectory() as d:
        cwd = os.getcwd()
        os.chdir(d)
        db.call_hashes.clear()
        acc, txt = [], []
        for b in pay:
            db.call_hashes.clear()
            r = real_unpack(b)
            acc.append(r is not None)
            txt.append("|".join(r) if r is not None else "")
        os.chdir(cwd)
    dec.unpack = real_unpack
    sym, b77 = tx.pack_message("CQ", "G1OJS", "IO90")
    wf = tx.symbols_to_complex_audio(sym, f_base=1500.0)
    np.savez_compressed(os.path.join(OUT, "codec.npz"),
                        payload_hex=np.array(["%x" % b for b in pay]), accepted=np.array(acc),
                        text=np.array(txt),
                        kat_bits77_hex="%x" % b77, kat_symbols=np.array(sym, np.int32),
                        kat_crc14=np.int32(tx.append_crc(b77)[1]),
                        kat_cw174_hex="%x" % tx.ldpc_encode(tx.append_crc(b77)[0])[0],
                        kat_wf_idx=np.arange(0, len(wf), 997), kat_wf=wf[::997])
    print("accepted", sum(acc), "of", len(acc))


if __name__ == "__main__":
    main()
