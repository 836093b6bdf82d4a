"""Large-N parity, driver-run (`-m gpu`): the BASELINE.json configurations that round 1 only covered with builder-kept sweeps.

* cfg1 / cfg2 / cfg4 (configs[0], [1], [3]): 64 device-generated cycles each, whole CUDA path vs the CPU oracle run on
  every host core.  Asserted: identical emitted payload SETS for every cycle (bit-exact 77-bit payloads), and -- counted,
  reported and bounded -- emission order, decode notes (pass / AP / method / tweaks) and dt / df / SNR tolerances.
* cfg3 (configs[2]): 20 000 noisy codewords at Eb/N0 0..4 dB through ft8_ldpc(., 90, 20) then ft8_osd on the failures,
  every LDPC call and up to 2 000 OSD calls re-decoded by the oracle: 0 status / iteration-count / trial-word mismatches.

Each test writes what it measured to gpurun_out/parity_large_<name>.json when that directory exists (builder runs copy
it to profiles/); the assertion messages carry the same numbers for the driver's log.
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from pyft8_b200 import _lib as L
from pyft8_b200 import synth, workload
from pyft8_b200.engine import Engine, bits91_to_int
from pyft8_b200.receiver import record_to_message

pytestmark = pytest.mark.gpu
N_CYCLES = 64


def _oracle_cycle(a):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ft8_oracle as o
    recs, cl = o.decode_cycle(a)
    return [(r["bits77"], r["notes"], r["tsec"], r["fHz"], r["snr"]) for r in recs]


def _oracle_ldpc(x):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ft8_oracle as o
    z = x.copy()
    st, n, _ = o.ldpc_decode(z, 90, 20)
    return st, n, z


def _oracle_osd(x):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ft8_oracle as o
    b = o.osd(x.copy())
    return b if b else 0


@pytest.fixture(scope="module")
def pool():
    with mp.get_context("spawn").Pool(os.cpu_count() or 1) as p:
        yield p


def _report(name, d):
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"parity_large_{name}.json"), "w") as f:
            json.dump(d, f)
    print(f"[parity_large] {name}: {json.dumps(d)}")


@pytest.mark.parametrize("config,seed", [("cfg1_20sig", 11), ("cfg2_50sig", 12), ("cfg4_120sig", 14)])
def test_64_cycles_gpu_vs_oracle(pool, config, seed):
    import torch
    eng = Engine(max_cycles=N_CYCLES)
    params = workload.make_params(config, N_CYCLES, seed=seed)
    audio = torch.empty((N_CYCLES, 180000), dtype=torch.int16, device="cuda:0")
    workload.device_cycles(eng, params, audio.data_ptr())
    torch.cuda.synchronize()
    host = audio.cpu().numpy()
    rec, n = eng.decode_cycles(host)
    stats = eng.stats()
    eng.close()
    ref = pool.map(_oracle_cycle, [host[i] for i in range(N_CYCLES)], chunksize=1)
    set_diff, order_bad, notes_bad, tol_bad, sent_hit, n_ref, n_gpu = [], [], [], [], 0, 0, 0
    off = 0
    for b in range(N_CYCLES):
        r = rec[off:off + n[b]]
        off += n[b]
        em = r[r["emitted"] == 1]
        got = [bits91_to_int(x["bits91"]) >> 14 for x in em]
        want = [x[0] for x in ref[b]]
        n_ref += len(want)
        n_gpu += len(got)
        assert len(got) == len(set(got)), (config, b, "duplicate payload flagged emitted")
        d = set(got) ^ set(want)
        if d:
            set_diff.append((b, sorted("%x" % v for v in d)))
        elif got != want:
            order_bad.append(b)
        sent_hit += len(set(got) & set(params["pool_bits77"][i] for i in params["pick"][b]))
        refmap = {x[0]: x for x in ref[b]}
        for x, g in zip(em, got):
            if g in refmap:
                m = record_to_message(x)
                if m["decode_notes"] != refmap[g][1]:
                    notes_bad.append((b, "%x" % g, m["decode_notes"], refmap[g][1]))
                # tolerances of the north star: dt +-5 ms, df +-0.5 Hz, SNR +-1 dB
                if abs(m["tsec"] - refmap[g][2]) > 0.005 + 1e-9 or abs(m["fHz"] - refmap[g][3]) > 0.5 + 1e-9 \
                        or abs(int(m["their_snr"]) - refmap[g][4]) > 1:
                    tol_bad.append((b, "%x" % g))
    rep = dict(config=config, cycles=N_CYCLES, seed=seed, ref_decodes=n_ref, gpu_decodes=n_gpu,
               cycles_with_set_difference=len(set_diff), set_differences=set_diff[:8], order_differs=order_bad,
               notes_differ=notes_bad[:8], n_notes_differ=len(notes_bad), dt_df_snr_out_of_tolerance=tol_bad[:8],
               true_messages_decoded=sent_hit, candidates=stats["candidates"], fine_evals=stats["fine_evals"],
               ldpc_calls=stats["ldpc_calls"], osd_calls=stats["osd_calls"])
    _report(config, rep)
    assert n_ref > 5 * N_CYCLES, rep                       # the workload really decodes
    assert not set_diff, rep                              # payload sets bit-identical in every cycle
    assert not tol_bad, rep
    # Emission order inside a pass follows llr_sd descending and the pass name follows which AP attempt converged first;
    # both can flip on an fp32 near-tie between the CUDA FFT and numpy's (r01 sweeps: 1 + 1 in 3 008 cycles).  Counted and
    # bounded, not hidden:
    assert len(order_bad) <= 1 and len(notes_bad) <= 1, rep


def test_cfg3_fec_20k_codewords_vs_oracle(pool):
    """BASELINE configs[2] at a size the oracle finishes in seconds: decoders.py:153-171 then :223-272."""
    eng = Engine(max_cycles=1)
    rng = np.random.default_rng(33)
    msgs = [synth.pack77(*synth.random_message(rng)) for _ in range(512)]
    cws = np.array([synth.codeword_bits(b) for b in msgs], np.float32) * 2 - 1
    per = 4000
    rep = dict(points=[])
    tot_mism = tot_llr_bad = tot_osd_mism = 0
    for e in range(5):
        idx = rng.integers(0, len(msgs), per)
        sigma = np.sqrt(1.0 / (2.0 * (91.0 / 174.0) * 10 ** (e / 10)))
        y = cws[idx] + rng.normal(0, sigma, (per, 174)).astype(np.float32)
        llr = (2.83 * y / y.std(axis=1, keepdims=True)).astype(np.float32)
        x = llr.copy()
        st, ni, bits = eng.ldpc(x, 90, 20)
        ok = st == L.LDPC_OK
        fail = np.flatnonzero(~ok)
        found, ob = eng.osd(llr[fail]) if len(fail) else (np.zeros(0, np.int32), np.zeros((0, 3), np.uint32))
        assert np.all(eng.crc14(bits[ok]) == 3)
        ref = pool.map(_oracle_ldpc, [llr[i] for i in range(per)], chunksize=64)
        mism = llr_bad = 0
        for i, (s_, n_, z_) in enumerate(ref):
            # the oracle (like the reference) reports a stall as a failure that returns its llr
            mism += (s_ != (st[i] if st[i] != L.LDPC_STALL else L.LDPC_FAIL)) or (n_ != ni[i])
            llr_bad += not np.allclose(z_, x[i], rtol=2e-3, atol=2e-3, equal_nan=True)
        nosd = min(400, len(fail))
        ref_osd = pool.map(_oracle_osd, [llr[fail[j]] for j in range(nosd)], chunksize=4)
        osd_mism = sum((bits91_to_int(ob[j]) if found[j] else 0) != ref_osd[j] for j in range(nosd))
        sent = [msgs[i] for i in idx]
        wrong_bp = sum(bits91_to_int(bits[i]) >> 14 != sent[i] for i in np.flatnonzero(ok))
        rep["points"].append(dict(ebn0_db=e, n=per, bp_ok=float(ok.mean()), mean_its_ok=float(ni[ok].mean()) if ok.any() else None,
                                  osd_calls=int(len(fail)), osd_rescued=int((found > 0).sum()), wrong_bp=int(wrong_bp),
                                  oracle_ldpc_mismatch=int(mism), oracle_llr_out_of_tol=int(llr_bad),
                                  oracle_osd_checked=nosd, oracle_osd_mismatch=int(osd_mism)))
        tot_mism += mism
        tot_llr_bad += llr_bad
        tot_osd_mism += osd_mism
    eng.close()
    rep.update(codewords=5 * per, oracle_ldpc_mismatch=int(tot_mism), oracle_llr_out_of_tol=int(tot_llr_bad),
               oracle_osd_mismatch=int(tot_osd_mism))
    _report("cfg3_fec", rep)
    assert tot_mism == 0 and tot_osd_mism == 0, rep        # status, iteration count, OSD trial word: bit-exact
    assert tot_llr_bad <= 5, rep                           # post-LDPC llr within 2e-3 (fp32 tanh/atanh chains; r01: 2 in 50 000)
    bp = [p["bp_ok"] for p in rep["points"]]
    assert all(b1 >= b0 for b0, b1 in zip(bp, bp[1:])) and bp[-1] > 0.9, rep     # monotone in Eb/N0: the sweep is a sweep


def test_tensor_core_scan_equals_nine_fft_kernel():
    """fine_mode 0 (time scan + tcgen05 3xTF32 frequency scan + final transform) against fine_mode 1 (the literal nine inverse
    FFTs of receiver.py:147-159) on 256 cycles: every record field identical, incl. both tweaks of every decoded candidate."""
    import torch
    B = 256
    params = workload.make_params("cfg2_50sig", B, seed=99)
    audio, recs = None, {}
    for mode in (1, 0):
        eng = Engine(max_cycles=B, fine_mode=mode)
        if audio is None:
            audio = torch.empty((B, 180000), dtype=torch.int16, device="cuda:0")
            workload.device_cycles(eng, params, audio.data_ptr())
            torch.cuda.synchronize()
        r, n = eng.decode_cycles_dev(audio.data_ptr(), L.AUDIO_I16, B)
        recs[mode] = (np.array(r, copy=True), np.array(n, copy=True), eng.stats())
        eng.close()
    (r1, n1, s1), (r0, n0, s0) = recs[1], recs[0]
    assert np.array_equal(n0, n1) and len(r0) == len(r1) > 10 * B
    diff = {k: int(np.sum(np.any(np.atleast_2d((r0[k] != r1[k]).T), axis=0))) for k in
            ("bits91", "cycle", "cand", "ipass", "ap", "method", "ttweak", "ftweak", "nsync", "snr", "emitted", "n_its")}
    rep = dict(records=len(r0), field_mismatches=diff, fine_pass=[s0["fine_pass"], s1["fine_pass"]], fine_evals=s0["fine_evals"])
    _report("fine_tc_vs_fft", rep)
    assert not any(diff.values()), rep
    # candidates that pass the Costas count (nsync > 6): a frequency-tweak near-tie can move one across the threshold
    assert abs(s0["fine_pass"] - s1["fine_pass"]) <= 2, rep
