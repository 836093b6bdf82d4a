"""BASELINE config 3: LDPC/OSD stress on noisy 174-bit LLR codewords at Eb/N0 0..4 dB (GPU; prints one JSON line).

  python tools/fec_stress.py [--n 1000000] [--check 200]

Codewords come from a pool of valid messages (pyft8_b200.synth), BPSK + AWGN, llr = 2.83*y/std(y).
Runs ldpc_decode(.,90,20) on all vectors and osd_012 on the failures through the C ABI, reports BP success, mean
iterations, OSD fallback/rescue rates and codewords/s, and spot-checks `--check` vectors per point against the oracle.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pyft8_b200 import synth, _lib as L  # noqa: E402
from pyft8_b200.engine import Engine, bits91_to_int  # noqa: E402


def _oracle_ldpc(x):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ft8_oracle as o
    z = x.copy()
    st, n, b = o.ldpc_decode(z, 90, 20)
    return st, n, z


def _oracle_osd(x):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ft8_oracle as o
    b = o.osd(x.copy())
    return b if b else 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000000)
    ap.add_argument("--check", type=int, default=200, help="LDPC vectors per Eb/N0 point re-decoded by the CPU oracle")
    ap.add_argument("--check-osd", type=int, default=50, help="OSD vectors per point re-decoded by the CPU oracle")
    args = ap.parse_args()
    import multiprocessing as mp
    pool = mp.get_context("spawn").Pool(os.cpu_count())
    eng = Engine(max_cycles=1)
    rng = np.random.default_rng(3)
    msgs = [synth.pack77(*synth.random_message(rng)) for _ in range(1024)]
    cws = np.array([synth.codeword_bits(b) for b in msgs], np.float32) * 2 - 1
    per = args.n // 5
    points = []
    t_ldpc = t_osd = 0.0
    n_osd_total = 0
    for e in range(5):
        idx = rng.integers(0, 1024, per)
        sigma = np.sqrt(1.0 / (2.0 * (91.0 / 174.0) * 10 ** (e / 10)))
        y = cws[idx] + rng.normal(0, sigma, (per, 174)).astype(np.float32)
        llr = (2.83 * y / y.std(axis=1, keepdims=True)).astype(np.float32)
        x = llr.copy()
        st, ni, bits = eng.ldpc(x, 90, 20)
        t_ldpc += eng.last_kernel_ms(0)
        ok = st == L.LDPC_OK
        fail = np.nonzero(~ok)[0]
        found, ob = eng.osd(llr[fail])
        t_osd += eng.last_kernel_ms(0)
        n_osd_total += len(fail)
        # correctness at full size: every accepted word passes CRC+validity; wrong = accepted but not the sent message
        sent = np.array([msgs[i] for i in idx], object)
        wrong_bp = sum(bits91_to_int(bits[i]) >> 14 != sent[i] for i in np.nonzero(ok)[0][:20000])
        resc = np.nonzero(found > 0)[0]
        wrong_osd = sum(bits91_to_int(ob[j]) >> 14 != sent[fail[j]] for j in resc[:20000])
        assert np.all(eng.crc14(bits[ok]) == 3)
        if len(resc):
            assert np.all(eng.crc14(ob[resc]) == 3)
        # oracle check (decision-level identity; llr after the last update within 2e-3)
        nchk = min(args.check, per)
        ref = pool.map(_oracle_ldpc, [llr[i] for i in range(nchk)], chunksize=64)
        mism = llr_bad = 0
        for i, (s_, n_, z_) in enumerate(ref):
            mism += (s_ != (st[i] if st[i] != L.LDPC_STALL else L.LDPC_FAIL)) or (n_ != ni[i])
            llr_bad += not np.allclose(z_, x[i], rtol=2e-3, atol=2e-3, equal_nan=True)
        nosd = min(args.check_osd, len(fail))
        ref_osd = pool.map(_oracle_osd, [llr[fail[j]] for j in range(nosd)], chunksize=8)
        osd_mism = sum((bits91_to_int(ob[j]) if found[j] else 0) != ref_osd[j] for j in range(nosd))
        points.append(dict(ebn0_db=e, n=per, bp_ok=float(ok.mean()), mean_its_ok=float(ni[ok].mean()) if ok.any() else None,
                           osd_fallback=float(len(fail) / per), osd_rescued=float(len(resc) / max(len(fail), 1)),
                           wrong_bp_in_sample=int(wrong_bp), wrong_osd_in_sample=int(wrong_osd), oracle_mismatch=int(mism),
                           oracle_llr_out_of_tol=int(llr_bad), oracle_checked=nchk, oracle_osd_mismatch=int(osd_mism),
                           oracle_osd_checked=nosd))
    out = dict(metric="LDPC codewords/sec", value=args.n / (t_ldpc / 1e3), unit="codewords/s", n=args.n,
               ldpc_kernel_ms=t_ldpc, osd_calls=n_osd_total, osd_kernel_ms=t_osd,
               osd_per_sec=n_osd_total / (t_osd / 1e3), config="BASELINE configs[2]: Eb/N0 0..4 dB, ldpc_decode(.,90,20) then osd_012",
               points=points)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
