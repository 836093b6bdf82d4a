"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): 2 cycles through every entry point."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyft8_b200 import synth
from pyft8_b200.engine import Engine

eng = Engine(max_cycles=2)
a = np.stack([synth.make_cycle(s, n_signals=8, snr_db=(-14, 4))[0] for s in (1, 2)])
g = eng.spectrogram(a)
f0, h0, sc, n, pay = eng.sync(g)
llr, sd, snr = eng.llr(pay[0, :int(n[0])])
spec = eng.cycle_spectrum(a)
k = min(int(n[0]), 12)
r = eng.fine(spec, np.zeros(k, np.int32), f0[0, :k], h0[0, :k])
x = llr[:32].copy()
st, ni, bits = eng.ldpc(x, 90, 20)
fo, ob = eng.osd(llr[:16])
fl = eng.crc14(bits)
rec, cnt = eng.decode_cycles(a)
row = eng.hop_spectrum(a[0].astype(np.float32))
print("decoded", len(rec), "emitted", int(rec["emitted"].sum()), "cands", n, "ldpc ok", int((st == 1).sum()))
eng.close()
