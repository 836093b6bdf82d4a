"""A/B harness for kernel variants: python tools/variant_bench.py lib1.so [lib2.so ...]

For every library (loaded in its own process through PYFT8_B200_LIB) decode the same device-generated batch a few times and
print per-stage CUDA-event times plus a digest of the record array, so that a variant can be accepted only if it is faster
AND leaves the records bit-identical (or the difference is understood)."""
import hashlib
import json
import os
import subprocess
import sys

STAGES = ["total", "spectrogram", "sync", "cycle_spectrum", "pass0", "fine", "pass234", "osd", "records"]


def child(B, reps):
    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from pyft8_b200 import workload, _lib as L
    from pyft8_b200.engine import Engine
    eng = Engine(0, max_cycles=B)
    prm = workload.make_params("cfg2_50sig", B, seed=2)
    audio = torch.empty((B, 180000), dtype=torch.int16, device="cuda:0")
    workload.device_cycles(eng, prm, audio.data_ptr())
    eng.synchronize()
    ms = np.zeros((reps, 9))
    rec = None
    for r in range(reps + 2):
        rec, n = eng.decode_cycles_dev(audio.data_ptr(), L.AUDIO_I16, B)
        if r >= 2:
            ms[r - 2] = [eng.last_kernel_ms(i) for i in range(9)]
    out = {"lib": os.environ.get("PYFT8_B200_LIB", "default"), "B": B, "records": int(len(rec)),
           "digest": hashlib.sha1(rec.tobytes()).hexdigest()[:12],
           "ms": {k: round(float(v), 3) for k, v in zip(STAGES, np.median(ms, axis=0))}}
    np.save("/tmp/vb_%s.npy" % os.path.basename(out["lib"]), rec)
    print("VARIANT " + json.dumps(out), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]))
    else:
        B = int(os.environ.get("VB_CYCLES", "4096"))
        for lib in sys.argv[1:]:
            env = dict(os.environ)
            if lib != "default":
                env["PYFT8_B200_LIB"] = os.path.abspath(lib)
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(B), "5"], env=env, capture_output=True, text=True)
            lines = [l for l in p.stdout.splitlines() if l.startswith("VARIANT ")]
            print(lines[0] if lines else "FAILED %s: %s" % (lib, (p.stderr or p.stdout)[-800:]), flush=True)
        # field-level difference of every variant's records against the first one's
        import numpy as np
        recs = [np.load("/tmp/vb_%s.npy" % os.path.basename(lib)) for lib in sys.argv[1:] if os.path.exists("/tmp/vb_%s.npy" % os.path.basename(lib))]
        for lib, r in zip(sys.argv[2:], recs[1:]):
            a = recs[0]
            if len(a) != len(r):
                print("DIFF %s: record count %d vs %d" % (lib, len(a), len(r)))
                continue
            bad = np.nonzero((a.view(np.uint8).reshape(len(a), -1) != r.view(np.uint8).reshape(len(r), -1)).any(1))[0]
            if len(bad) == 0:
                print("DIFF %s: records bit-identical" % lib)
                continue
            fields = {f: int((a[f][bad] != r[f][bad]).reshape(len(bad), -1).any(1).sum()) for f in a.dtype.names}
            print("DIFF %s: %d records differ; per field %s" % (lib, len(bad), {k: v for k, v in fields.items() if v}))
            for i in bad[:6]:
                print("   base", a[i], "\n   this", r[i])
