"""Golden vectors for the time-gated (progressive) scheduler: tests/golden/progressive.npz, from the UNMODIFIED reference
driven by oracle/ref_harness.decode_cycle_progressive (authoring container only; needs /root/reference or oracle/_ref):
the body of Receiver.manage_cycle's loop (receiver.py:379-412) runs once after every hop, so a candidate is decoded as soon
as its payload rows are on the grid -- before the trailing Costas block and the rest of the cycle are in the audio ring --
and the messages and the hop each one comes out at are a deterministic function of the audio."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_harness as rh  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import cycle_audio  # noqa: E402


def main():
    out = {}
    for name in ("test_08", "syn20"):
        with tempfile.TemporaryDirectory() as d:
            o = rh.decode_cycle_progressive(cycle_audio(name), workdir=d)
        assert o["pending"] == 0
        out[f"{name}_text"] = np.array([" ".join(m["msg_tuple"]) for m in o["messages"]])
        out[f"{name}_notes"] = np.array([m["decode_notes"] for m in o["messages"]])
        out[f"{name}_hop"] = np.array(o["emit_hop"], np.int32)
        out[f"{name}_snr"] = np.array([int(m["their_snr"]) for m in o["messages"]], np.int32)
        print(name, o["n_cands"], "candidates,", len(o["messages"]), "messages, hops", min(o["emit_hop"]), "..", max(o["emit_hop"]))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "progressive.npz"), **out)


if __name__ == "__main__":
    main()
