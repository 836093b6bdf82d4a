"""Is the unmodified reference deterministic across hosts?  (test infrastructure; needs oracle/_ref)

osd_012 orders bit reliabilities with np.argsort's default (unstable) sort (decoders.py:229) and the AP patterns force
~30 positions to |llr| = 5 (receiver.py:109-117), i.e. exact ties.  numpy picks its sort kernel by CPU features
(AVX-512 hosts use a different introsort), so the tie order -- and with it which OSD trial word is found first -- can
differ between machines.  This script prints a digest of argsort on a vector with such ties and compares the reference
with the port (stable order) on N host-generated cycles of cfg2.

  python tools/ref_tie_check.py [N]
"""
import hashlib
import json
import multiprocessing as mp
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402


def work(args):
    seed, b = args
    os.environ["PYFT8_REF_ROOT"] = os.path.join(ROOT, "oracle", "_ref")
    from pyft8_b200 import workload
    import ref_harness as rh
    import ft8_oracle as o
    p = workload.make_params("cfg2_50sig", 4, seed=seed)
    a = workload.host_cycle(p, b)
    with tempfile.TemporaryDirectory() as d:
        out = rh.decode_cycle(a, workdir=d)
    ref = {" ".join(m["msg_tuple"]): m["decode_notes"] for m in out["messages"]}
    recs, _ = o.decode_cycle(a)
    port = {" ".join(o.unpack77(r["bits77"]) or ("?",)): r["notes"] for r in recs}
    return seed, b, ref, port


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    rng = np.random.default_rng(7)
    v = np.abs(rng.normal(0, 3, 174)).astype(np.float32)
    v[rng.choice(174, 34, replace=False)] = 5.0
    d_default = hashlib.sha1(np.argsort(-v).astype(np.int64).tobytes()).hexdigest()[:12]
    d_stable = hashlib.sha1(np.argsort(-v, kind="stable").astype(np.int64).tobytes()).hexdigest()[:12]
    try:
        feats = np.lib.introspect.opt_func_info(func_name="argsort")
    except Exception:
        feats = None
    print(json.dumps({"numpy": np.__version__, "argsort_default_digest": d_default, "argsort_stable_digest": d_stable,
                      "default_equals_stable": d_default == d_stable, "argsort_dispatch": str(feats)[:300]}))
    jobs = [(100 + i // 4, i % 4) for i in range(n)]
    with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
        res = pool.map(work, jobs, chunksize=1)
    # the same unmodified reference with numpy's SIMD sort kernels switched off (scalar introsort): a different tie order
    os.environ["NPY_DISABLE_CPU_FEATURES"] = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR AVX2 FMA3"
    with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
        res_scalar = pool.map(work, jobs, chunksize=1)
    bad = bad_scalar = ref_vs_ref = n_msgs = 0
    for (seed, b, ref, port), (_, _, ref2, _) in zip(res, res_scalar):
        n_msgs += len(ref)
        for name, r in (("reference", ref), ("reference_scalar_sort", ref2)):
            if set(r) != set(port):
                print(json.dumps({"seed": seed, "cycle": b, "which": name, "only_" + name: {k: r[k] for k in set(r) - set(port)},
                                  "only_port": {k: port[k] for k in set(port) - set(r)}}))
        bad += set(ref) != set(port)
        bad_scalar += set(ref2) != set(port)
        ref_vs_ref += set(ref) != set(ref2)
    print(json.dumps({"cycles": n, "messages": n_msgs, "cycles_where_reference_differs_from_stable_order_port": bad,
                      "cycles_where_scalar_sort_reference_differs_from_port": bad_scalar,
                      "cycles_where_the_reference_differs_from_itself_across_numpy_sort_kernels": ref_vs_ref}))


if __name__ == "__main__":
    main()
