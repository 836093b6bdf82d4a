"""Pin the CPU oracle (oracle/ft8_oracle.py) against vectors produced by the UNMODIFIED reference.

The golden files were written by oracle/make_golden.py in the authoring container, where the reference
(G1OJS/PyFT8 v3.9.0) runs under the SURVEY 8c harness.  These tests run anywhere (no reference tree, no GPU).
"""
import zlib

import numpy as np
import pytest

import ft8_oracle as o
from conftest import ALL_CYCLES, load_golden


@pytest.fixture(scope="module")
def decoded(golden_cycles):
    out = {}
    for name in ALL_CYCLES:
        audio, g = golden_cycles[name]
        grid = o.spectrogram(audio)
        cands = o.search(grid)
        recs, cl = o.decode_cycle(audio, grid=grid, cands=cands)
        out[name] = (grid, cands, recs, cl)
    return out


@pytest.mark.parametrize("name", ALL_CYCLES)
def test_spectrogram_rows(name, golden_cycles, decoded):
    _, g = golden_cycles[name]
    grid = decoded[name][0]
    assert grid.shape == (376, 976) and grid.dtype == np.float32
    # same numpy => bit-identical; tolerance only guards against a different SIMD path for log10/abs
    np.testing.assert_allclose(grid[g["grid_rows"]], g["grid_subset"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(grid.astype(np.float64).mean(axis=0), g["grid_col_mean"], atol=1e-4)
    np.testing.assert_allclose(grid.astype(np.float64).mean(axis=1), g["grid_row_mean"], atol=1e-4)
    assert np.all(grid[0] == 1.0)


@pytest.mark.parametrize("name", ALL_CYCLES)
def test_search_candidates(name, golden_cycles, decoded):
    _, g = golden_cycles[name]
    f0, h0, sc, pay = decoded[name][1]
    assert np.array_equal(f0, g["cand_f0"])          # identical list in identical rank order
    assert np.array_equal(h0, g["cand_h0"])
    np.testing.assert_allclose(sc, g["cand_score"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(pay[g["payload_idx"]], g["payload_subset"], rtol=0, atol=2e-4)


@pytest.mark.parametrize("name", ALL_CYCLES)
def test_llr_and_fine_trace(name, golden_cycles, decoded):
    _, g = golden_cycles[name]
    cl = decoded[name][3]
    gsd = np.array([c.grid_sd for c in cl])
    np.testing.assert_allclose(gsd, g["grid_sd"], rtol=1e-5)
    for i, c in enumerate(cl):
        assert c.final_ipass == g["final_ipass"][i], (i, c.final_ipass)
        if g["has_fine"][i]:
            assert c.nsync == g["nsync"][i]
            assert c.tweaks == " t:%+03d f:%+03d" % (g["tt"][i], g["ff"][i])
            if c.nsync > 6 and np.isfinite(g["fine_sd"][i]):   # golden has no sd when the sd<=5 gate stopped it
                assert abs(c.fine_sd - g["fine_sd"][i]) <= 1e-5 * abs(g["fine_sd"][i])
    # grid-stage LLR vectors
    f0, h0, sc, pay = decoded[name][1]
    for k, i in enumerate(g["grid_llr_sel"]):
        llr, sd, snr = o.db_to_llr(pay[i])
        np.testing.assert_allclose(llr, g["grid_llr"][k], rtol=1e-5, atol=1e-6)
        assert snr == g["grid_snr"][i]


@pytest.mark.parametrize("name", ALL_CYCLES)
def test_fine_grid_vectors(name, golden_cycles):
    audio, g = golden_cycles[name]
    spec = o.cycle_spectrum(audio)
    for k, i in enumerate(g["fine_sel"]):
        fHz, tsec = 3.125 * g["cand_f0"][i], g["cand_h0"][i] / 25.0
        r = o.llr_fine(spec, fHz, tsec)
        assert (r["tt"], r["ff"], r["nsync"]) == (g["tt"][i], g["ff"][i], g["nsync"][i])
        np.testing.assert_allclose(r["grid"], g["fine_grid"][k], rtol=1e-5, atol=1e-4)
        llr, sd, snr = o.db_to_llr(20 * np.log10(r["grid"][list(o.PAYLOAD_SYMS), :]))
        np.testing.assert_allclose(llr, g["fine_llr"][k], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ALL_CYCLES)
def test_decode_set_equals_reference(name, golden_cycles, decoded):
    """The decoded message set, pass names, tweaks, dt/df/snr all equal the reference's, in emission order."""
    _, g = golden_cycles[name]
    recs = decoded[name][2]
    assert ["%x" % r["bits77"] for r in recs] == list(g["msg_bits77_hex"])
    assert [r["notes"] for r in recs] == list(g["msg_notes"])
    assert [r["cand"] for r in recs] == list(g["emit_order"])
    np.testing.assert_allclose([r["tsec"] for r in recs], g["msg_tsec"], atol=1e-9)
    np.testing.assert_allclose([r["fHz"] for r in recs], g["msg_fHz"], atol=1e-9)
    assert ["%+03d" % r["snr"] for r in recs] == list(g["msg_snr"])
    assert [" ".join(o.unpack77(r["bits77"])) for r in recs] == list(g["msg_text"])
    # per-candidate results, including duplicates that were decoded but not emitted
    cl = decoded[name][3]
    for i, c in enumerate(cl):
        want = int(g["dec_bits77_hex"][i], 16)
        got = (c.bits91 >> 14) if c.bits91 is not None else 0
        assert got == want, i
        if want:
            assert c.notes == g["dec_notes"][i]


def test_expected_message_counts(golden_cycles):
    """SURVEY 8c golden counts: 21 decodes on test_08.wav, 20 on test_09.wav, 200 candidates each."""
    assert len(golden_cycles["test_08"][1]["msg_text"]) == 21
    assert len(golden_cycles["test_09"][1]["msg_text"]) == 20
    assert "CQ SO6HQK AA37" in list(golden_cycles["test_08"][1]["msg_text"])     # the false decode is reproduced
    assert len(golden_cycles["test_08"][1]["cand_f0"]) == 200


# ------------------------------------------------------------------ FEC
def test_ldpc_captured_calls():
    f = load_golden("fec.npz")
    n = len(f["cap_ldpc_in"])
    for i in range(n):
        llr = f["cap_ldpc_in"][i].copy()
        st, nits, bits = o.ldpc_decode(llr, int(f["cap_ldpc_nc0"][i]), int(f["cap_ldpc_its"][i]))
        assert (st == o.ST_OK) == bool(f["cap_ldpc_ok"][i]), i
        assert nits == f["cap_ldpc_nits"][i]
        assert (st == o.ST_FAIL) == bool(f["cap_ldpc_hasllr"][i])
        np.testing.assert_allclose(llr, f["cap_ldpc_out"][i], rtol=1e-4, atol=1e-4, equal_nan=True)


def test_ldpc_osd_on_noisy_codewords():
    from pyft8_b200 import synth
    f = load_golden("fec.npz")
    for e in range(5):
        llr, truth = synth.make_llr_codewords(3000 + e, 120, float(e))
        assert np.uint32(zlib.crc32(llr.tobytes())) == f[f"e{e}_llr_crc"]
        for i in range(0, 120, 2):
            x = llr[i].copy()
            st, nits, bits = o.ldpc_decode(x, 90, 20)
            assert st == f[f"e{e}_status"][i] and nits == f[f"e{e}_nits"][i]
            np.testing.assert_allclose(x, f[f"e{e}_llr_out"][i], rtol=1e-4, atol=1e-4, equal_nan=True)
            if st == o.ST_OK:
                assert bits >> 14 == truth[i]
            else:
                b = o.osd(llr[i].copy())
                assert ("%x" % ((b >> 14) if b else 0)) == f[f"e{e}_osd_bits77_hex"][i]


def test_osd_captured_calls():
    f = load_golden("fec.npz")
    for i in range(len(f["cap_osd_in"])):
        b = o.osd(f["cap_osd_in"][i].copy())
        assert ("%x" % ((b >> 14) if b else 0)) == f["cap_osd_bits77_hex"][i], i


def test_osd_order0_is_systematic_reencode():
    """Property: every OSD trial word is a codeword prefix -- re-encoding its 91 bits reproduces the hard
    decisions on all 91 pivot columns for the order-0 word."""
    from pyft8_b200 import synth
    llr, _ = synth.make_llr_codewords(99, 6, 2.0)
    for x in llr:
        c0 = o.osd_candidates(x, 0, 0)[0]
        par = 0
        for m in o.GEN_MASK91:
            par = (par << 1) | (bin(c0 & m).count("1") & 1)
        cw = [(((c0 << 83) | par) >> (173 - i)) & 1 for i in range(174)]
        order = o.osd_order(x)
        agree = sum(cw[j] == (x[j] > 0) for j in order[:60])
        assert agree == 60   # the most reliable (independent) positions are reproduced


# ------------------------------------------------------------------ codec
def test_valid77_and_unpack_text():
    c = load_golden("codec.npz")
    for h, acc, txt in zip(c["payload_hex"], c["accepted"], c["text"]):
        b = int(h, 16)
        assert o.valid77(b) == bool(acc), h
        if acc:
            assert "|".join(o.unpack77(b)) == txt


def test_crc_and_encoder_kat():
    from pyft8_b200 import synth
    c = load_golden("codec.npz")
    b77 = int(str(c["kat_bits77_hex"]), 16)
    assert b77 == 0x409003831f091
    assert o.crc14(b77) == int(c["kat_crc14"]) == 0x2ca1
    assert synth.crc14(b77) == 0x2ca1
    assert synth.pack77("CQ", "G1OJS", "IO90") == b77
    assert "%x" % synth.encode174(b77) == str(c["kat_cw174_hex"])
    assert synth.symbols_from_bits77(b77) == list(c["kat_symbols"])
    assert o.crc_ok91((b77 << 14) | 0x2ca1) and not o.crc_ok91((b77 << 14) | 0x2ca0)
    assert not o.crc_ok91(0)
    wf = synth.shift_carrier(synth.gfsk_baseband(list(c["kat_symbols"])), 1500.0)
    np.testing.assert_allclose(wf[c["kat_wf_idx"]], c["kat_wf"], atol=1e-9)
