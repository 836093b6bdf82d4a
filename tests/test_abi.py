"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from pyft8_b200 import _lib as L


def _header_functions():
    src = open(os.path.join(ROOT, "include", "ft8_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ft8_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = _header_functions()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ft8_b200.h but not exported by libft8_b200.so"
    assert set(names) == set(L.SIGNATURES), "ctypes signature table out of sync with the header"


def test_struct_layouts_match_header():
    assert ctypes.sizeof(L.Record) == 64 and L.RECORD_DTYPE.itemsize == 64
    assert ctypes.sizeof(L.Cfg) == 48
    assert ctypes.sizeof(L.Stats) == 128
    for name in ("cycle", "cand", "snr", "emitted", "n_its", "score", "fine_sd"):
        assert L.RECORD_DTYPE.fields[name][1] == getattr(L.Record, name).offset


def test_default_cfg_mirrors_reference_defaults():
    cfg = L.Cfg()
    L.load().ft8_default_cfg(ctypes.byref(cfg))
    assert (cfg.max_cands, cfg.sync_score_min, cfg.llr_sd_min, cfg.osd_singleflips, cfg.osd_doubleflips) == (200, 85.0, 5.0, 30, 2)


def test_no_cpu_fallback_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pyft8_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CUDA device"):
        Engine()
    from pyft8_b200 import decoders
    import numpy as np
    with pytest.raises(RuntimeError):
        decoders.ldpc_decode(np.zeros(174, np.float32), 35, 5)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pyft8_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "ft8_oracle" not in txt and "ref_harness" not in txt, f


def test_header_compiles_and_links_from_c(tmp_path):
    """The boundary is a C ABI: a C99 translation unit including only include/ft8_b200.h links against the library."""
    import subprocess
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.join(ROOT, "pyft8_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", libdir, "-l:libft8_b200.so",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
