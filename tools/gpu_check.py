"""Stage-by-stage comparison of the CUDA path with the oracle on a GPU box (diagnostic; prints, never asserts)."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ft8_oracle as o  # noqa: E402
from conftest import cycle_audio, load_golden  # noqa: E402
from pyft8_b200 import synth  # noqa: E402
from pyft8_b200.engine import Engine, bits91_to_int, int_to_bits91  # noqa: E402


def section(name):
    print(f"\n=== {name}", flush=True)


def main():
    eng = Engine(max_cycles=4)
    names = sys.argv[1:] or ["test_08", "syn20"]
    section("fft")
    rng = np.random.default_rng(0)
    for n in (32, 256, 375, 1920, 3200):
        x = (rng.normal(size=(3, n)) + 1j * rng.normal(size=(3, n))).astype(np.complex64)
        for inv in (False, True):
            y = eng.debug_fft(x, inv)
            ref = np.fft.ifft(x.astype(np.complex128), axis=1) * n if inv else np.fft.fft(x.astype(np.complex128), axis=1)
            print(f"n={n} inv={inv} rel err {np.abs(y-ref).max()/np.abs(ref).max():.2e}")
    section("crc/valid")
    c = load_golden("codec.npz")
    pay = [int(h, 16) for h in c["payload_hex"]]
    words = np.array([int_to_bits91((b << 14) | o.crc14(b)) for b in pay], np.uint32)
    fl = eng.crc14(words)
    acc = c["accepted"]
    print("crc ok all (except payload 0):", int(((fl & 1) == 1).sum()), "of", len(pay), "; valid mismatches:",
          int((((fl & 2) != 0) != acc).sum()))
    bad = np.nonzero(((fl & 2) != 0) != acc)[0][:5]
    for i in bad:
        print("   mismatch payload", hex(pay[i]), "ref", acc[i], "text", c["text"][i])
    words2 = words.copy(); words2[:, 0] ^= 1
    print("crc rejects flipped:", int(((eng.crc14(words2) & 1) == 0).sum()), "of", len(pay))
    section("ldpc/osd on noisy codewords")
    f = load_golden("fec.npz")
    for e in range(5):
        llr, truth = synth.make_llr_codewords(3000 + e, 120, float(e))
        x = llr.copy()
        st, ni, bits = eng.ldpc(x, 90, 20)
        st_ref = f[f"e{e}_status"]
        stm = np.where(st == 3, 2, st)
        print(f"Eb/N0 {e}: status mismatch {int((stm != st_ref).sum())}, nits mismatch {int((ni != f[f'e{e}_nits']).sum())}, "
              f"llr max abs diff {np.nanmax(np.abs(x - f[f'e{e}_llr_out'])):.3e}, ok {int((st==1).sum())}")
        ok = st == 1
        wrong = sum(1 for i in np.nonzero(ok)[0] if bits91_to_int(bits[i]) >> 14 != truth[i])
        print("   wrong payloads among OK:", wrong)
        found, ob = eng.osd(llr)
        ref_osd = f[f"e{e}_osd_bits77_hex"]
        mm = 0
        for i in range(120):
            if ref_osd[i] == "-":
                continue
            got = (bits91_to_int(ob[i]) >> 14) if found[i] else 0
            mm += ("%x" % got) != ref_osd[i]
        print("   osd mismatches vs reference:", mm, "found", int((found > 0).sum()))
    xi = f["cap_ldpc_in"].copy()
    for nc0, its in ((35, 5), (90, 20)):
        sel = np.nonzero((f["cap_ldpc_nc0"] == nc0) & (f["cap_ldpc_its"] == its))[0]
        x = np.ascontiguousarray(xi[sel])
        st, ni, bits = eng.ldpc(x, nc0, its)
        ok_ref = f["cap_ldpc_ok"][sel]
        print(f"captured ldpc({nc0},{its}) n={len(sel)}: ok mismatch {int(((st==1)!=ok_ref).sum())}, nits mismatch "
              f"{int((ni!=f['cap_ldpc_nits'][sel]).sum())}, reject mismatch {int(((st==0)==f['cap_ldpc_hasllr'][sel]).sum() - int((st==1).sum()))}, "
              f"llr diff {np.nanmax(np.abs(x-f['cap_ldpc_out'][sel])):.3e}")
    found, ob = eng.osd(f["cap_osd_in"])
    mm = sum(("%x" % ((bits91_to_int(ob[i]) >> 14) if found[i] else 0)) != f["cap_osd_bits77_hex"][i] for i in range(len(found)))
    print("captured osd mismatches:", mm, "of", len(found), "found", int((found > 0).sum()))

    for name in names:
        section(f"cycle {name}")
        audio = cycle_audio(name)
        g = load_golden(f"cycle_{name}.npz")
        t = time.time(); grid_o = o.spectrogram(audio); t_o = time.time() - t
        grid = eng.spectrogram(audio)[0]
        d = np.abs(grid - grid_o)
        lin_o, lin = 10 ** (grid_o[1:] / 20), 10 ** (grid[1:] / 20)
        print(f"grid: max |dB diff| {d.max():.3e}, mean {d.mean():.3e}, row0 ones {bool(np.all(grid[0]==1))}, "
              f"rel lin err max {np.max(np.abs(lin-lin_o)/np.maximum(lin_o, 1e-3*lin_o.max())):.3e} (oracle {t_o*1e3:.0f} ms)")
        f0, h0, sc, n, pay = eng.sync(grid)
        f0o, h0o, sco, payo = o.search(grid_o)
        n = int(n[0])
        same = n == len(f0o) and np.array_equal(f0[0, :n], f0o) and np.array_equal(h0[0, :n], h0o)
        print(f"sync: n {n} vs {len(f0o)}, identical ranked list {same}")
        if not same:
            so, sg = set(zip(f0o.tolist(), h0o.tolist())), set(zip(f0[0, :n].tolist(), h0[0, :n].tolist()))
            print("   only oracle:", sorted(so - sg)[:10], " only gpu:", sorted(sg - so)[:10])
            k = min(n, len(f0o))
            print("   rank mismatches:", int(((f0[0, :k] != f0o[:k]) | (h0[0, :k] != h0o[:k])).sum()))
        k = min(n, len(f0o))
        print(f"   score diff max {np.abs(sc[0,:k]-sco[:k]).max():.3e}")
        # same-input parity: feed the ORACLE grid to the GPU sync
        f0b, h0b, scb, nb, payb = eng.sync(grid_o)
        nb = int(nb[0])
        same2 = nb == len(f0o) and np.array_equal(f0b[0, :nb], f0o) and np.array_equal(h0b[0, :nb], h0o)
        print(f"sync on oracle grid: identical {same2}; payload identical {np.array_equal(payb[0,:nb], payo) if same2 else 'n/a'}")
        llr, sd, snr = eng.llr(payo)
        lo = [o.db_to_llr(p) for p in payo]
        print(f"llr: max diff {max(np.abs(llr[i]-lo[i][0]).max() for i in range(len(lo))):.3e}, sd rel {max(abs(sd[i]-lo[i][1])/lo[i][1] for i in range(len(lo))):.2e}, snr mismatch {sum(int(snr[i])!=lo[i][2] for i in range(len(lo)))}")
        spec = eng.cycle_spectrum(audio)[0]
        spec_o = o.cycle_spectrum(audio)
        print(f"cycle spectrum: rel err {np.abs(spec-spec_o).max()/np.abs(spec_o).max():.2e}; band rel {np.abs(spec[1418:48832]-spec_o[1418:48832]).max()/np.abs(spec_o[1418:48832]).max():.2e}")
        sel = np.nonzero(g["has_fine"])[0]
        r = eng.fine(spec_o, np.zeros(len(sel), np.int32), g["cand_f0"][sel], g["cand_h0"][sel])
        print(f"fine (oracle spectrum, {len(sel)} cands): tt mismatch {int((r['tt']!=g['tt'][sel]).sum())}, ff mismatch {int((r['ff']!=g['ff'][sel]).sum())}, nsync mismatch {int((r['nsync']!=g['nsync'][sel]).sum())}")
        for k2, i in enumerate(g["fine_sel"]):
            j = int(np.nonzero(sel == i)[0][0])
            if r["tt"][j] == g["tt"][i] and r["ff"][j] == g["ff"][i]:
                gd = np.abs(r["grid"][j] - g["fine_grid"][k2]).max() / g["fine_grid"][k2].max()
                ld = np.abs(r["llr"][j] - g["fine_llr"][k2]).max()
                print(f"   cand {i}: grid rel {gd:.2e} llr diff {ld:.2e}")
        r2 = eng.fine(spec, np.zeros(len(sel), np.int32), g["cand_f0"][sel], g["cand_h0"][sel], want_grid=False)
        print(f"fine (gpu spectrum): tt mismatch {int((r2['tt']!=g['tt'][sel]).sum())}, ff mismatch {int((r2['ff']!=g['ff'][sel]).sum())}, nsync mismatch {int((r2['nsync']!=g['nsync'][sel]).sum())}")
        t = time.time(); rec, nrec = eng.decode_cycles(audio); t_g = time.time() - t
        em = rec[rec["emitted"] == 1]
        got = ["%x" % (bits91_to_int(x["bits91"]) >> 14) for x in em]
        want = list(g["msg_bits77_hex"])
        print(f"decode_cycles: {len(rec)} decoded, {len(em)} emitted vs reference {len(want)}; set equal {set(got)==set(want)}; order equal {got==want}  ({t_g*1e3:.1f} ms)")
        print("   missing:", [w for w in want if w not in got], " extra:", [x for x in got if x not in want])
        print("   stats", eng.stats())
        notes = {h: n for h, n in zip(g["msg_bits77_hex"], g["msg_notes"])}
        for x, hx in zip(em, got):
            from pyft8_b200 import _lib as L
            src = "grid" if x["ipass"] == 0 else "fine"
            nm = f"{src}_{L.AP_NAMES[x['ap']]}_{L.METHOD_NAMES[x['method']]}"
            tw = "t:+00 f:+00" if x["ipass"] == 0 else " t:%+03d f:%+03d" % (x["ttweak"], x["ftweak"])
            if hx in notes and notes[hx] != nm + tw:
                print("   notes differ:", hx, nm + tw, "| ref", notes[hx])
    eng.close()


if __name__ == "__main__":
    try:
        main()
    except Exception:
        traceback.print_exc()
        sys.exit(1)
