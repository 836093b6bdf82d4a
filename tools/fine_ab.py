"""A/B of the two fine-sync implementations on the GPU: fine_mode 0 (time scan + tcgen05 frequency scan + final transform)
against fine_mode 1 (nine inverse FFTs per candidate).  Prints one JSON line: per-stage ms and every record field that differs.

  python tools/fine_ab.py [--cycles 1024] [--workload cfg2_50sig]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cycles", type=int, default=1024)
    ap.add_argument("--workload", default="cfg2_50sig")
    ap.add_argument("--seed", type=int, default=2000)
    a = ap.parse_args()
    import torch
    from pyft8_b200 import workload, _lib as L
    from pyft8_b200.engine import Engine
    B = a.cycles
    out = {"cycles": B, "workload": a.workload}
    recs = {}
    audio = None
    for mode in (1, 0):
        eng = Engine(max_cycles=B, fine_mode=mode)
        if audio is None:
            params = workload.make_params(a.workload, B, seed=a.seed)
            audio = torch.empty((B, 180000), dtype=torch.int16, device="cuda:0")
            workload.device_cycles(eng, params, audio.data_ptr())
            torch.cuda.synchronize()
        for _ in range(2):
            r, n = eng.decode_cycles_dev(audio.data_ptr(), L.AUDIO_I16, B)
        recs[mode] = (np.array(r, copy=True), np.array(n, copy=True))
        out[f"mode{mode}_ms"] = {"all": eng.last_kernel_ms(0), "fine": eng.last_kernel_ms(5)}
        out[f"mode{mode}_stats"] = {k: v for k, v in eng.stats().items() if k in ("fine_evals", "fine_pass", "decoded", "emitted")}
        eng.close()
    (r1, n1), (r0, n0) = recs[1], recs[0]
    out["same_counts"] = bool(np.array_equal(n0, n1))
    if len(r0) == len(r1):
        out["field_mismatches"] = {k: int(np.sum(r0[k] != r1[k]) if r0[k].ndim == 1 else np.sum(np.any(r0[k] != r1[k], axis=1)))
                                   for k in ("bits91", "cycle", "cand", "ipass", "ap", "method", "ttweak", "ftweak", "nsync", "snr", "emitted")}
        bad = np.flatnonzero((r0["ftweak"] != r1["ftweak"]) | (r0["ttweak"] != r1["ttweak"]))[:8]
        out["examples"] = [{"cycle": int(r0["cycle"][i]), "cand": int(r0["cand"][i]), "tc": [int(r0["ttweak"][i]), int(r0["ftweak"][i])],
                            "fft": [int(r1["ttweak"][i]), int(r1["ftweak"][i])]} for i in bad]
    else:
        out["records"] = [len(r0), len(r1)]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
