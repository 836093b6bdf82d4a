"""Drop-in for PyFT8/decoders.py's FEC entry points, backed by the CUDA library (no CPU fallback).

Same names, arguments and return contracts as the reference:
  ldpc_decode(llr, max_ncheck0, max_iters) -> (msg_tuple|None, n_its, []|llr)   decoders.py:153-171 (llr updated in place)
  osd_012(llr, singleflips=30, doubleflips=2) -> msg_tuple|None                 decoders.py:223-272
  crc_unpack91(codeword91) -> msg_tuple|None                                     decoders.py:117-131
`unpack` (text formatting) stays in Python (messages.py).
"""
import numpy as np

from . import _lib as L
from .engine import Engine, bits91_to_int
from .messages import unpack, call_hashes, add_call_hashes  # noqa: F401  (re-exported like the reference module)

_engine = None


def get_engine():
    """Process-wide default engine on device 0 (created on first use; raises when there is no GPU / library)."""
    global _engine
    if _engine is None:
        _engine = Engine(device=0, max_cycles=1, max_cands=928)
    return _engine


def set_engine(engine):
    global _engine
    _engine = engine


def crc_unpack91(codeword91):
    bits = (np.asarray(codeword91)[:91] > 0)
    w = np.zeros(3, np.uint32)
    for j in np.nonzero(bits)[0]:
        w[j >> 5] |= np.uint32(1 << (j & 31))
    flags = get_engine().crc14(w[None, :])
    if flags[0] & 1:
        return unpack(bits91_to_int(w) >> 14)
    return None


def ldpc_decode(llr, max_ncheck0, max_iters):
    x = np.ascontiguousarray(llr, np.float32).reshape(1, 174)
    st, nits, bits = get_engine().ldpc(x, max_ncheck0, max_iters)
    if st[0] == L.LDPC_REJECT:
        return None, -1, []
    if x.base is not llr and x is not llr:
        llr[:] = x[0]                     # in-place contract of the reference (decoders.py:169)
    if st[0] == L.LDPC_OK:
        msg = unpack(bits91_to_int(bits[0]) >> 14)
        if msg:
            return msg, int(nits[0]), []
    return None, -1, llr


def osd_012(llr, singleflips=30, doubleflips=2):
    found, bits = get_engine().osd(np.asarray(llr, np.float32), singleflips, doubleflips)
    if found[0]:
        return unpack(bits91_to_int(bits[0]) >> 14)
    return None
