"""Synthetic FT8 cycles: message packing, LDPC/CRC encoding and the GFSK modulator.

Workload generator for tests and ``bench.py`` (BASELINE.json: "synthetic FT8 cycles produced by
PyFT8's own transmitter.py GFSK generator plus AWGN").  It restates what the reference transmitter
computes -- 77-bit packing of standard messages (transmitter.py:97-174), CRC-14 append
(transmitter.py:208-223), parity bits from the generator rows (transmitter.py:181-187), Gray map
and Costas framing (transmitter.py:189-206) and the BT=2.0 Gaussian-smoothed phase modulator
(transmitter.py:41-70) -- so that the GPU box, where the reference tree is absent, can build the
same signals.  ``tests/test_oracle_golden.py`` (encoder / CRC / modulator known answers) pins it against vectors
made with the unmodified reference.  The mixing recipe (amplitude convention, seeds) is SURVEY.md section 8d.
"""
import math

import numpy as np

from .tables import GEN_MASK91

SAMP_RATE = 12000
SPS = 1920                      # samples per symbol at 6.25 baud
COSTAS = (3, 1, 4, 0, 6, 5, 2)
GRAY = (0, 1, 3, 2, 5, 6, 4, 7)
NTOKENS, MAX22 = 2063592, 4194304
_A1 = " 0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ"
_A4 = " ABCDEFGHIJKLMNOPQRSTUVWXYZ"


# ------------------------------------------------------------------ packing
def pack_call28(call):
    """Standard callsign or token -> (n28, suffix flag).  transmitter.py:135-151."""
    if call in ("DE", "QRZ", "CQ"):
        return ("DE", "QRZ", "CQ").index(call), 0
    p = 1 if call[-2:] in ("/P", "/R") else 0
    core = call.replace("/P", "").replace("/R", "")
    if len(core) > 6:
        raise ValueError("not a standard callsign: %r" % call)
    if not core[2].isdigit():
        core = " " + core
    core = (core + "      ")[:6]
    idx = (_A1.find(core[0]), _A1[1:].find(core[1]), _A1[1:].find(core[2]),
           _A4.find(core[3]), _A4.find(core[4]), _A4.find(core[5]))
    if min(idx) < 0:
        raise ValueError("not a standard callsign: %r" % call)
    n = idx[0]
    for i, radix in zip(idx[1:], (36, 10, 27, 27, 27)):
        n = n * radix + i
    return n + NTOKENS + MAX22, p


def pack_g15(txt):
    """Grid / report / RRR / RR73 / 73 -> (g15, ir).  transmitter.py:153-174."""
    if txt[:1] in "+-" and txt[1:].isdigit():
        return 32435 + int(txt), 0
    if txt[:2] in ("R+", "R-"):
        return 32435 + int(txt[1:]), 1
    special = {"RRR": 32402, "RR73": 32403, "73": 32404}
    if txt in special:
        return special[txt], 0
    if len(txt) != 4:
        return 0, 0
    v = (ord(txt[0]) - 65) * 18 + (ord(txt[1]) - 65)
    return (v * 10 + int(txt[2])) * 10 + int(txt[3]), 0


def pack77(c1, c2, extra):
    """Three-field standard message -> 77-bit payload (i3 = 1, or 2 with a /P call).  transmitter.py:97-106."""
    n28a, pa = pack_call28(c1)
    n28b, pb = pack_call28(c2)
    g15, ir = pack_g15(extra)
    i3 = 2 if c1.endswith("/P") or c2.endswith("/P") else 1
    return (n28a << 49) | (pa << 48) | (n28b << 20) | (pb << 19) | (ir << 18) | (g15 << 3) | i3


# ------------------------------------------------------------------ channel coding
def crc14(bits77):
    """CRC-14, polynomial 0x2757, over the payload zero-extended to 82 bits.  transmitter.py:208-223."""
    r = 0
    for i in range(96):
        bit = (bits77 >> (76 - i)) & 1 if i < 77 else 0
        carry = r & 0x2000
        r = ((r << 1) & 0x3FFF) | bit
        if carry:
            r ^= 0x2757
    return r


def encode174(bits77):
    """payload -> 174-bit codeword int (message, CRC, 83 parity bits).  transmitter.py:181-201."""
    b91 = (bits77 << 14) | crc14(bits77)
    par = 0
    for m in GEN_MASK91:
        par = (par << 1) | (bin(b91 & m).count("1") & 1)
    return (b91 << 83) | par


def codeword_bits(bits77):
    cw = encode174(bits77)
    return np.array([(cw >> (173 - i)) & 1 for i in range(174)], np.uint8)


def symbols_from_bits77(bits77):
    """79 channel symbols: Costas + 29 data + Costas + 29 data + Costas.  transmitter.py:189-206."""
    cw = encode174(bits77)
    data = [GRAY[(cw >> (171 - 3 * i)) & 7] for i in range(58)]
    return list(COSTAS) + data[:29] + list(COSTAS) + data[29:] + list(COSTAS)


# ------------------------------------------------------------------ modulator
def _pulse(bt=2.0):
    from scipy.special import erf
    c = math.pi * math.sqrt(2.0 / math.log(2.0))
    t = (np.arange(3 * SPS) - 1.5 * SPS) / SPS
    return 0.5 * (erf(c * bt * (t + 0.5)) - erf(c * bt * (t - 0.5)))


_PULSE = None


def gfsk_baseband(symbols):
    """Complex GFSK waveform with carrier 0 Hz, 79*1920 samples.  transmitter.py:52-70 (f_base = 0).

    A carrier f is applied by ``shift_carrier``; the reference adds 2*pi*f*n/12000 before dropping
    its first 1920 guard samples, hence the (n + 1920) there.
    """
    global _PULSE
    if _PULSE is None:
        _PULSE = _pulse()
    n = SPS * (len(symbols) + 2)
    dphi = np.zeros(n)
    step = 2.0 * math.pi / SPS
    for i, tone in enumerate(symbols):
        dphi[i * SPS:i * SPS + 3 * SPS] += step * _PULSE * tone
    phi = np.add.accumulate(dphi)
    phi[:2 * SPS] += step * _PULSE[SPS:] * symbols[0]
    phi[-2 * SPS:] += step * _PULSE[:-SPS] * symbols[-1]
    phi = phi[SPS:-SPS]
    wf = np.exp(1j * (phi % (2 * math.pi)))
    nr = int(0.5 + SPS / 8.0)
    ramp = np.cos(np.linspace(0, math.pi, nr))
    wf[:nr] *= (1 - ramp) / 2.0
    wf[-nr:] *= (1 + ramp) / 2.0
    return wf


def shift_carrier(wf, f_hz):
    n = np.arange(len(wf)) + SPS
    return wf * np.exp(2j * math.pi * f_hz * n / SAMP_RATE)


# ------------------------------------------------------------------ random traffic
_PFX = ("G", "M", "F", "DL", "EA", "I", "OH", "SM", "PA", "ON", "OZ", "LA", "SP", "OK", "HB", "OE",
        "K", "W", "N", "VE", "JA", "VK", "ZL", "UA", "LZ", "YO", "9A", "S5", "EI", "CT", "2E", "4X")
_LET = "ABCDEFGHIJKLMNOPQRSTUVWXYZ"


def random_call(rng):
    p = _PFX[int(rng.integers(len(_PFX)))]
    d = str(int(rng.integers(10)))
    nsuf = int(rng.integers(1, 4)) if len(p) == 2 else int(rng.integers(2, 4))
    suf = "".join(_LET[int(rng.integers(26))] for _ in range(nsuf))
    return p + d + suf


def random_message(rng):
    """A random standard message of the kinds named in SURVEY 8d; every call passes the reference's validator."""
    kind = int(rng.integers(5))
    a, b = random_call(rng), random_call(rng)
    grid = _LET[int(rng.integers(18))] + _LET[int(rng.integers(18))] + "%02d" % int(rng.integers(100))
    if kind == 0:
        return ("CQ", a, grid)
    if kind == 1:
        return (a, b, grid)
    if kind == 2:
        return (a, b, "%+03d" % int(rng.integers(-24, 20)))
    if kind == 3:
        return (a, b, "R%+03d" % int(rng.integers(-24, 20)))
    return (a, b, ("RR73", "73", "RRR")[int(rng.integers(3))])


def make_cycle(seed, n_signals=20, snr_db=(-20.0, 5.0), f_hz=(200.0, 2950.0), dt_s=(-0.5, 1.0),
               noise_sigma=1000.0):
    """One 15 s cycle of int16 audio: n_signals GFSK signals + white noise.  SURVEY.md 8d.

    SNR is quoted in 2500 Hz (WSJT-X convention): amplitude = sigma*sqrt(2*(2500/6000)*10^(snr/10)).
    Signals start at sample int((0.5+dt)*12000).  Returns (audio int16[180000], truth list of dicts).
    """
    rng = np.random.default_rng(seed)
    x = rng.normal(0.0, noise_sigma, 180000)
    truth = []
    for _ in range(n_signals):
        msg = random_message(rng)
        b77 = pack77(*msg)
        snr = float(rng.uniform(*snr_db))
        f = float(rng.uniform(*f_hz))
        dt = float(rng.uniform(*dt_s))
        amp = noise_sigma * math.sqrt(2.0 * (2500.0 / 6000.0) * 10.0 ** (snr / 10.0))
        wf = np.imag(shift_carrier(gfsk_baseband(symbols_from_bits77(b77)), f)) * amp
        s0 = int((0.5 + dt) * SAMP_RATE)
        lo, hi = max(s0, 0), min(s0 + len(wf), 180000)
        x[lo:hi] += wf[lo - s0:hi - s0]
        truth.append(dict(msg=msg, bits77=b77, snr=snr, f=f, dt=dt))
    return np.clip(np.round(x), -32768, 32767).astype(np.int16), truth


def make_llr_codewords(seed, n, ebn0_db):
    """Config 3 of BASELINE.json: random valid codewords, BPSK + AWGN, llr = 2.83*y/std(y).  SURVEY 8d."""
    rng = np.random.default_rng(seed)
    sigma = math.sqrt(1.0 / (2.0 * (91.0 / 174.0) * 10.0 ** (ebn0_db / 10.0)))
    llr = np.empty((n, 174), np.float32)
    truth = []
    for i in range(n):
        b77 = pack77(*random_message(rng))
        bits = codeword_bits(b77).astype(np.float64)
        y = (2.0 * bits - 1.0) + rng.normal(0.0, sigma, 174)
        llr[i] = (2.83 * y / np.std(y)).astype(np.float32)
        truth.append(b77)
    return llr, truth
