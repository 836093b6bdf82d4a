"""Decode-set parity sweep: N synthetic cycles of a BASELINE config on the GPU vs the CPU oracle (multiprocess).

  python tools/parity_sweep.py --config cfg2_50sig --n 64 [--seed 5]

Prints one JSON line: cycles with identical emitted payload sets, total symmetric difference, notes/tweak mismatches.
Test infrastructure (uses oracle/); the audio is produced by the library's generator kernel and downloaded.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _oracle(a):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ft8_oracle as o
    recs, cl = o.decode_cycle(a)
    return [(r["bits77"], r["notes"], r["tsec"], r["fHz"], r["snr"]) for r in recs], len(cl)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2_50sig")
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--dump", default="", help="directory for the audio + records of every cycle that differs (offline analysis)")
    args = ap.parse_args()
    import torch
    from pyft8_b200 import workload, _lib as L
    from pyft8_b200.engine import Engine, bits91_to_int
    from pyft8_b200.receiver import record_to_message
    eng = Engine(max_cycles=args.n)
    params = workload.make_params(args.config, args.n, seed=args.seed)
    audio = torch.empty((args.n, 180000), dtype=torch.int16, device="cuda:0")
    workload.device_cycles(eng, params, audio.data_ptr())
    torch.cuda.synchronize()
    host = audio.cpu().numpy()
    rec, n = eng.decode_cycles(host)
    # the oracle's answer depends only on (config, n, seed): cached so that several library builds can be compared in one call
    import pickle
    cache = "/tmp/parity_sweep_%s_%d_%d.pkl" % (args.config, args.n, args.seed)
    if os.path.exists(cache):
        ref = pickle.load(open(cache, "rb"))
    else:
        with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
            ref = pool.map(_oracle, [host[i] for i in range(args.n)], chunksize=1)
        pickle.dump(ref, open(cache, "wb"))
    same = symdiff = notes_bad = order_bad = dtdf_bad = cand_bad = 0
    tot_ref = tot_gpu = 0
    sent_hit = 0
    off = 0
    for b in range(args.n):
        r = rec[off:off + n[b]]
        off += n[b]
        em = r[r["emitted"] == 1]
        got = [bits91_to_int(x["bits91"]) >> 14 for x in em]
        want = [x[0] for x in ref[b][0]]
        tot_ref += len(want)
        tot_gpu += len(got)
        d = set(got) ^ set(want)
        symdiff += len(d)
        same += not d
        order_bad += (not d) and got != want
        pool_bits = set(params["pool_bits77"][i] for i in params["pick"][b])
        sent_hit += len(set(got) & pool_bits)
        refmap = {x[0]: x for x in ref[b][0]}
        cyc_bad = bool(d) or got != want
        for x, g in zip(em, got):
            if g in refmap:
                m = record_to_message(x)
                notes_bad += m["decode_notes"] != refmap[g][1]
                cyc_bad |= m["decode_notes"] != refmap[g][1]
                dtdf_bad += abs(m["tsec"] - refmap[g][2]) > 0.005 + 1e-9 or abs(m["fHz"] - refmap[g][3]) > 0.5 + 1e-9 or abs(int(m["their_snr"]) - refmap[g][4]) > 1
        if cyc_bad and args.dump:
            os.makedirs(args.dump, exist_ok=True)
            np.savez_compressed(os.path.join(args.dump, "%s_seed%d_cycle%d.npz" % (args.config, args.seed, b)), audio=host[b], rec=r,
                                ref_bits77=np.array(["%x" % x[0] for x in ref[b][0]]), ref_notes=np.array([x[1] for x in ref[b][0]]))
    print(json.dumps(dict(lib=os.path.basename(os.environ.get("PYFT8_B200_LIB", "default")), config=args.config, cycles=args.n, identical_sets=int(same), symmetric_difference=int(symdiff),
                          ref_decodes=tot_ref, gpu_decodes=tot_gpu, order_differs=int(order_bad), notes_differ=int(notes_bad),
                          dt_df_snr_out_of_tolerance=int(dtdf_bad), true_messages_decoded=int(sent_hit))))


if __name__ == "__main__":
    main()
