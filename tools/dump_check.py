"""Re-decode the cycles a parity sweep dumped (tools/parity_sweep.py --dump DIR) with both fine-sync modes and list, per cycle,
how each mode differs from the oracle's decode list stored in the dump: tells whether a deviation comes from the tensor-core
frequency scan (fine_mode 0) or is shared with the literal nine-FFT kernel (fine_mode 1), i.e. from fp32 rounding elsewhere."""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyft8_b200.engine import Engine, bits91_to_int  # noqa: E402
from pyft8_b200.receiver import record_to_message  # noqa: E402

for f in sorted(glob.glob(os.path.join(sys.argv[1], "*.npz"))):
    z = np.load(f, allow_pickle=True)
    ref = {int(h, 16): n for h, n in zip(z["ref_bits77"].tolist(), z["ref_notes"].tolist())}
    for mode in (0, 1):
        eng = Engine(max_cycles=1, fine_mode=mode)
        rec, _ = eng.decode_cycles(z["audio"][None])
        em = rec[rec["emitted"] == 1]
        got = {bits91_to_int(x["bits91"]) >> 14: record_to_message(x)["decode_notes"] for x in em}
        diffs = [("only_gpu", "%x" % k, got[k]) for k in got.keys() - ref.keys()] + [("only_ref", "%x" % k, ref[k]) for k in ref.keys() - got.keys()] \
            + [("notes", "%x" % k, got[k], ref[k]) for k in got.keys() & ref.keys() if got[k] != ref[k]]
        order = [bits91_to_int(x["bits91"]) >> 14 for x in em] != [int(h, 16) for h in z["ref_bits77"].tolist()]
        print(os.path.basename(f), "fine_mode", mode, diffs, "order differs" if (order and not diffs) else "")
        eng.close()
