"""Four-way decode comparison on the first N cycles of the bench batch (test infrastructure; GPU box):
CUDA at the bench batch size, CUDA on the N cycles alone, the numpy port, and the unmodified reference (oracle/_ref).

  python tools/parity_probe.py [--cycles 4096] [--n 16] [--workload cfg2_50sig]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402


def cpu_both(a):
    os.environ["PYFT8_REF_ROOT"] = os.path.join(ROOT, "oracle", "_ref")
    import ft8_oracle as o
    recs, _ = o.decode_cycle(a)
    port = {r["bits77"]: r["notes"] for r in recs}
    ref = None
    if os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "PyFT8")):
        import ref_harness as rh
        with tempfile.TemporaryDirectory() as d:
            out = rh.decode_cycle(a, workdir=d)
        ref = {" ".join(m["msg_tuple"]): m["decode_notes"] for m in out["messages"]}
    return port, ref


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cycles", type=int, default=4096)
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--workload", default="cfg2_50sig")
    ap.add_argument("--seed", type=int, default=2000)
    a = ap.parse_args()
    import torch
    from pyft8_b200 import workload, messages, _lib as L
    from pyft8_b200.engine import Engine, bits91_to_int
    from pyft8_b200.receiver import record_to_message
    B = a.cycles
    eng = Engine(max_cycles=B)
    params = workload.make_params(a.workload, B, seed=a.seed)
    audio = torch.empty((B, 180000), dtype=torch.int16, device="cuda:0")
    workload.device_cycles(eng, params, audio.data_ptr())
    torch.cuda.synchronize()
    rec, cnt = eng.decode_cycles_dev(audio.data_ptr(), L.AUDIO_I16, B)
    rec2, cnt2 = eng.decode_cycles_dev(audio.data_ptr(), L.AUDIO_I16, B)
    same_runs = np.array_equal(cnt, cnt2) and np.array_equal(rec["bits91"], rec2["bits91"]) and np.array_equal(rec["emitted"], rec2["emitted"])
    host = audio[:a.n].cpu().numpy()
    small = Engine(max_cycles=a.n)
    rec_s, cnt_s = small.decode_cycles(host)
    with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
        cpu = pool.map(cpu_both, [host[i] for i in range(a.n)], chunksize=1)
    off = np.concatenate([[0], np.cumsum(cnt)])
    off_s = np.concatenate([[0], np.cumsum(cnt_s)])
    out = {"two_runs_at_B_identical": bool(same_runs), "cycles": []}

    def emitted(r):
        e = r[r["emitted"] == 1]
        return {bits91_to_int(x["bits91"]) >> 14: record_to_message(x)["decode_notes"] for x in e}
    for i in range(a.n):
        big, sm = emitted(rec[off[i]:off[i + 1]]), emitted(rec_s[off_s[i]:off_s[i + 1]])
        port, ref = cpu[i]
        ent = {"cycle": i, "n": [len(big), len(sm), len(port), len(ref) if ref is not None else None],
               "cudaB_eq_cudaN": set(big) == set(sm), "cudaN_eq_port": set(sm) == set(port)}
        messages.call_hashes.clear()
        txt = lambda d: {" ".join(messages.unpack(b) or ("?",)): n for b, n in d.items()}
        if ref is not None:
            ent["port_eq_reference"] = set(txt(port)) == set(ref)
            ent["cudaB_eq_reference"] = set(txt(big)) == set(ref)
        if not (ent["cudaB_eq_cudaN"] and ent["cudaN_eq_port"] and ent.get("port_eq_reference", True)):
            tb, ts, tp = txt(big), txt(sm), txt(port)
            allk = set(tb) | set(ts) | set(tp) | set(ref or {})
            ent["diff"] = {k: {"cudaB": tb.get(k), "cudaN": ts.get(k), "port": tp.get(k), "reference": (ref or {}).get(k)}
                           for k in allk if not (k in tb and k in ts and k in tp and (ref is None or k in ref))}
        out["cycles"].append(ent)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
