"""Phase timing of k_fine (debug build with -DFINE_PROFILE): clock64 sums per barrier phase.
   nvcc ... -DFINE_PROFILE -o build/variants/fineprof.so pyft8_b200/csrc/ft8_b200.cu
   PYFT8_B200_LIB=build/variants/fineprof.so python tools/fine_phase_profile.py"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyft8_b200 import workload, _lib as L
from pyft8_b200.engine import Engine
B = 1024
eng = Engine(0, max_cycles=B)
prm = workload.make_params("cfg2_50sig", B, seed=2)
audio = torch.empty((B, 180000), dtype=torch.int16, device="cuda:0")
workload.device_cycles(eng, prm, audio.data_ptr()); eng.synchronize()
lib = eng._lib
out = (C.c_ulonglong * 16)()
eng.decode_cycles_dev(audio.data_ptr(), L.AUDIO_I16, B)
lib.ft8_debug_fine_prof(out, 1)
eng.decode_cycles_dev(audio.data_ptr(), L.AUDIO_I16, B)
lib.ft8_debug_fine_prof(out, 0)
v = [int(x) for x in out]
n, npd = max(v[6], 1), max(v[7], 1)
print("transforms timed:", n, "producer transforms:", npd)
print("phase A (pass 8,25, all warps)         : %.0f clk" % (v[0] / n))
print("phase B total (barrier to barrier)     : %.0f clk" % (v[3] / n))
print("  consumer warp 0: window              : %.0f clk" % (v[4] / n))
print("  consumer warp 0: window+bar+score    : %.0f clk" % (v[1] / n))
print("  consumer warp 3: window+bar          : %.0f clk" % (v[5] / n))
print("  producer warp 4: fused passes        : %.0f clk" % (v[2] / npd))
nc = max(v[11], 1)
print("per candidate (%d timed):" % nc)
print("  time scan: pass (8,25)                 : %.0f clk" % (v[8] / nc))
print("  time scan: consumer window+8 scores    : %.0f clk   producer fused passes: %.0f clk" % (v[9] / nc, v[10] / nc))
print("  final: full last pass                  : %.0f clk" % (v[12] / nc))
print("  final: warp 0 grid+llr %.0f clk, warp 3 grid %.0f clk, producer next-item passes %.0f clk" % (v[13] / nc, v[15] / nc, v[14] / nc))
