import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


SYN_CYCLES = {
    "syn20": (1000, dict(n_signals=20, snr_db=(-20, 5), f_hz=(200, 2950), dt_s=(-0.5, 1.0))),
    "syn50": (2000, dict(n_signals=50, snr_db=(-24, 10), f_hz=(200, 2950), dt_s=(-0.5, 1.0))),
    "syn120": (4000, dict(n_signals=120, snr_db=(-24, 10), f_hz=(200, 2950), dt_s=(-0.5, 1.0))),
}
ALL_CYCLES = ("test_08", "test_09", "syn20", "syn50", "syn120")


def cycle_audio(name):
    """int16[180000] audio of a golden cycle (WAV fixtures are stored; synthetic ones are re-made from the seed)."""
    g = load_golden(f"cycle_{name}.npz")
    if "audio" in g.files:
        return g["audio"]
    from pyft8_b200 import synth
    seed, kw = SYN_CYCLES[name]
    a, _ = synth.make_cycle(seed, **kw)
    import zlib
    assert np.uint32(zlib.crc32(a.tobytes())) == g["audio_crc"], "synthetic generator drifted from the golden run"
    return a


@pytest.fixture(scope="session")
def golden_cycles():
    return {n: (cycle_audio(n), load_golden(f"cycle_{n}.npz")) for n in ALL_CYCLES}
