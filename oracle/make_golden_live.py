"""Golden vectors for the live (two-cycle ring) path: tests/golden/live_pairs.npz, from the UNMODIFIED reference driven by
oracle/ref_harness.decode_two_cycles (authoring container only; needs /root/reference or oracle/_ref).

Pairs: (test_08, test_09) -- the reference's own WAV fixtures back to back -- and a synthetic 30 s stream whose second cycle
contains signals that start up to 1.45 s BEFORE the cycle boundary (h0 < -32: their first payload rows lie in the previous
cycle's half of the ring) plus signals across the boundary region.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_harness as rh  # noqa: E402
from pyft8_b200 import synth  # noqa: E402


def synth_stream(seed, n_sig=24):
    """30 s of int16 audio: signals start anywhere in [0.0, 1.5] s of either cycle, and six of the second cycle's start
    1.30 .. 1.45 s before its boundary."""
    rng = np.random.default_rng(seed)
    x = rng.normal(0.0, 1000.0, 360000)
    starts = [float(rng.uniform(0.0, 1.5)) for _ in range(n_sig // 2)] + [15.0 + float(rng.uniform(0.0, 1.5)) for _ in range(n_sig // 2 - 6)] \
        + [15.0 - float(rng.uniform(1.30, 1.45)) for _ in range(6)]
    for t0 in starts:
        b77 = synth.pack77(*synth.random_message(rng))
        snr = float(rng.uniform(-14, 5))
        f = float(rng.uniform(250, 2900))
        amp = 1000.0 * np.sqrt(2.0 * (2500.0 / 6000.0) * 10.0 ** (snr / 10.0))
        wf = np.imag(synth.shift_carrier(synth.gfsk_baseband(synth.symbols_from_bits77(b77)), f)) * amp
        s0 = int(t0 * 12000)
        lo, hi = max(s0, 0), min(s0 + len(wf), len(x))
        x[lo:hi] += wf[lo - s0:hi - s0]
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def pack(out):
    d = {}
    for i, o in enumerate(out):
        d[f"text_{i}"] = np.array([" ".join(m["msg_tuple"]) for m in o["messages"]])
        d[f"notes_{i}"] = np.array([m["decode_notes"] for m in o["messages"]])
        d[f"tsec_{i}"] = np.array([m["tsec"] for m in o["messages"]], np.float64)
        d[f"fhz_{i}"] = np.array([m["fHz"] for m in o["messages"]], np.float64)
        d[f"snr_{i}"] = np.array([int(m["their_snr"]) for m in o["messages"]], np.int32)
        d[f"cand_f0_{i}"] = o["cand_f0"]
        d[f"cand_h0_{i}"] = o["cand_h0"]
    return d


def main():
    g8 = np.load(os.path.join(ROOT, "tests", "golden", "cycle_test_08.npz"))["audio"]
    g9 = np.load(os.path.join(ROOT, "tests", "golden", "cycle_test_09.npz"))["audio"]
    syn = synth_stream(777)
    out = {}
    for name, (a, b) in (("wav", (g8, g9)), ("syn", (syn[:180000], syn[180000:]))):
        with tempfile.TemporaryDirectory() as d:
            res = rh.decode_two_cycles(a, b, workdir=d)
        for k, v in pack(res).items():
            out[f"{name}_{k}"] = v
        print(name, [len(r["messages"]) for r in res], [int((r["cand_h0"] < -32).sum()) for r in res])
    out["syn_seed"] = np.int64(777)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "live_pairs.npz"), **out)


if __name__ == "__main__":
    main()
