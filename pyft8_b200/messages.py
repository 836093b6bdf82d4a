"""77-bit payload -> message text (stays in Python, as in the reference: decoders.py:16-115, databases.py:8-26).

The CUDA path decides accept/reject on the device (csrc/codec.cuh); this module only formats the text of accepted
payloads and keeps the callsign-hash history that '<...>' resolution needs.  Deviation from the reference, stated:
rejected callsigns are not appended to ./rejected_callsigns.txt (decoders.py:114-115 side effect).
"""
from .tables import PREFIX2

call_hashes = {}          # (hash, nbits) -> callsign, like databases.call_hashes
hashes_for_calls = {}

_C38 = " 0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ/"
_A1 = " 0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ"
_A2 = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ"
_A4 = " ABCDEFGHIJKLMNOPQRSTUVWXYZ"
NTOKENS, MAX22 = 2063592, 4194304
_M64 = (1 << 64) - 1


def add_call_hashes(call):
    """10/12/22-bit hashes of a callsign (databases.py:10-26)."""
    x = 0
    for c in (call + "          ")[:11]:
        x = (38 * x + _C38.find(c)) & _M64
    x = (x * 47055833459) & _M64
    hs = []
    for m in (10, 12, 22):
        hsh = (x >> (64 - m), m)
        hs.append(hsh)
        call_hashes[hsh] = call
    hashes_for_calls[call] = hs


def _shape_ok(c):
    if " " in c or len(c) < 3:
        return False
    if c[0] in "ABCDEFGHIJKLMNOPRSTUVWXYZ" and c[1].isdigit() and not (c[0] in "BFGIKMNRW" and c[2].isdigit()):
        return True
    return c[1] in PREFIX2.get(c[0], "") and c[2].isdigit()


def _call28_text(n28):
    nn = n28 - (NTOKENS + MAX22)
    out = []
    for alphabet, div in ((_A1, 262440 * 27), (_A2, 196830), ("0123456789" + " " * 17, 19683), (_A4, 729), (_A4, 27), (_A4, 1)):
        i, nn = divmod(nn, div)
        out.append(alphabet[i])
    return "".join(out).strip()


def _call29(c29, i3):
    n28, p = c29 >> 1, c29 & 1
    if n28 < 3:
        return ("DE", "QRZ", "CQ")[n28]
    if n28 < 1004:
        return "CQ %03d" % (n28 - 3)
    if n28 < 21443:
        x, t = n28 - 1003, ""
        for _ in range(4):
            t = _A4[x % 27] + t
            x //= 27
        return "CQ " + t.strip()
    if n28 < NTOKENS + MAX22 - 1:
        return "<%s>" % call_hashes.get((n28 - NTOKENS, 22), "...")
    c = _call28_text(n28)
    if not _shape_ok(c):
        return None
    if p:
        c += "/P" if i3 == 2 else "/R"
    if c.endswith("/R") and c[0] not in "AKNW":
        return None
    add_call_hashes(c)
    return c


def unpack(bits):
    """Same contract as the reference's unpack(): (call_a, call_b, grid_or_report) or None."""
    bits = int(bits)
    if not bits:
        return None
    i3, b74 = bits & 7, bits >> 3
    if i3 in (1, 2):
        g16 = b74 & 0xFFFF
        g15 = g16 & 0x7FFF
        if g15 == 0:
            return None
        if g15 < 32400:
            a, r = divmod(g15, 1800)
            b, r = divmod(r, 100)
            extra = chr(65 + a) + chr(65 + b) + "%d%d" % divmod(r, 10)
        elif g15 <= 32404:
            extra = ("", "", "RRR", "RR73", "73")[g15 - 32400]
        else:
            extra = ("R" if g16 >> 15 else "") + "%+03d" % (g15 - 32435)
        t = (_call29((b74 >> 45) & 0x1FFFFFFF, i3), _call29((b74 >> 16) & 0x1FFFFFFF, i3), extra)
        return None if ("" in t or None in t) else t
    if i3 == 4:
        cq, rrr, swp = b74 & 1, (b74 >> 1) & 3, (b74 >> 3) & 1
        c58, h12 = (b74 >> 4) & ((1 << 58) - 1), (b74 >> 62) & 0xFFF
        if (cq and rrr) or (not cq and not rrr):
            return None
        ca = "CQ" if cq else "<%s>" % call_hashes.get((h12, 12), "...")
        cb = ""
        for _ in range(12):
            cb = _C38[c58 % 38] + cb
            c58 //= 38
        cb = cb.strip()
        add_call_hashes(cb)
        if swp:
            ca, cb = cb, ca
        return (ca, cb, ("", "RRR", "RR73", "73")[rrr])
    return None


# ---------------------------------------------------------------------------------------------- batch text formatting
_TEXT_CACHE = {}
_HASH_LO, _HASH_HI = NTOKENS, NTOKENS + MAX22 - 1      # n28 range rendered through the hash table ('<...>')


def _history_free(bits):
    """True when the text of this payload does not depend on the callsign-hash history (no '<...>' field)."""
    i3, b74 = bits & 7, bits >> 3
    if i3 not in (1, 2):
        return False
    a, b = (b74 >> 46) & 0xFFFFFFF, (b74 >> 17) & 0xFFFFFFF
    return not (_HASH_LO <= a < _HASH_HI or _HASH_LO <= b < _HASH_HI)


def unpack_many(payloads):
    """unpack() over a sequence of 77-bit payloads, in order, with the same hash-history side effects.

    Skimmer-scale output repeats payloads heavily (duplicate candidates inside a cycle, stations repeating a message in
    consecutive cycles), so history-independent payloads are formatted once and served from a cache; payloads whose text
    goes through the hash table are always re-evaluated."""
    out = []
    for b in payloads:
        b = int(b)
        t = _TEXT_CACHE.get(b)
        if t is None:
            t = unpack(b)
            if t is not None and _history_free(b):
                if len(_TEXT_CACHE) > 1 << 20:
                    _TEXT_CACHE.clear()
                _TEXT_CACHE[b] = t
        else:
            # keep the reference's side effect: every decoded standard call is (re-)entered in the hash table
            for c in t[:2]:
                if _is_plain_call(c):
                    hs = hashes_for_calls.get(c)
                    if hs is None or call_hashes.get(hs[2]) != c:
                        add_call_hashes(c)
        out.append(t)
    return out


def _is_plain_call(c):
    return not (c.startswith("CQ") or c in ("DE", "QRZ") or c.startswith("<"))
