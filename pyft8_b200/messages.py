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
# text output has to keep up with ~25 records x 30 k cycles/s per GPU.  The per-record work
# is numpy (field extraction from the packed words, table look-ups through np.unique inverses); Python-level formatting
# runs once per distinct callsign field / report field and is cached; only payloads whose text reads the hash history
# ('<...>' fields, type-4 messages) go through the scalar unpack(), at their position in the sequence.
import numpy as np

_CALL_CACHE_MAX = 1 << 21
_CALL_CACHE = {}          # c29 * 2 + (i3 == 2) -> (text | None, (h10, h12, h22) | None when nothing is registered)
_EXTRA = None             # g16 -> text ("" and None = reject; None also means: calls are not even evaluated)
_EXTRA_EVAL = _EXTRA_OK = None
_HASH_LO, _HASH_HI = NTOKENS, NTOKENS + MAX22 - 1      # n28 range rendered through the hash table ('<...>')


def _call29_pure(c29, i3_is_2):
    """(_call29 text, hashes to register) without touching the hash tables; only for n28 outside the hash range."""
    n28, p = c29 >> 1, c29 & 1
    if n28 < NTOKENS + MAX22 - 1:
        return _call29(c29, 2 if i3_is_2 else 1), None         # tokens / CQ forms: no side effect
    c = _call28_text(n28)
    if not _shape_ok(c):
        return None, None
    if p:
        c += "/P" if i3_is_2 else "/R"
    if c.endswith("/R") and c[0] not in "AKNW":
        return None, None
    x = 0
    for ch in (c + "          ")[:11]:
        x = (38 * x + _C38.find(ch)) & _M64
    x = (x * 47055833459) & _M64
    return c, tuple((x >> (64 - m), m) for m in (10, 12, 22))


def _extra_table():
    global _EXTRA, _EXTRA_EVAL, _EXTRA_OK
    if _EXTRA is None:
        t = np.empty(65536, object)
        for g16 in range(65536):
            g15 = g16 & 0x7FFF
            if g15 == 0:
                t[g16] = None
            elif g15 < 32400:
                a, r = divmod(g15, 1800)
                b, r = divmod(r, 100)
                t[g16] = chr(65 + a) + chr(65 + b) + "%d%d" % divmod(r, 10)
            elif g15 <= 32404:
                t[g16] = ("", "", "RRR", "RR73", "73")[g15 - 32400]
            else:
                t[g16] = ("R" if g16 >> 15 else "") + "%+03d" % (g15 - 32435)
        _EXTRA = t
        _EXTRA_EVAL = np.array([e is not None for e in t], bool)
        _EXTRA_OK = np.array([bool(e) for e in t], bool)
    return _EXTRA


def _rev32(v):
    v = ((v >> 1) & 0x55555555) | ((v & 0x55555555) << 1)
    v = ((v >> 2) & 0x33333333) | ((v & 0x33333333) << 2)
    v = ((v >> 4) & 0x0F0F0F0F) | ((v & 0x0F0F0F0F) << 4)
    v = ((v >> 8) & 0x00FF00FF) | ((v & 0x00FF00FF) << 8)
    return ((v >> 16) | (v << 16)) & 0xFFFFFFFF


def payload_fields(bits91_words):
    """[n,3] uint32 (ft8_record.bits91, LSB-first) -> dict of uint64 field arrays: i3, g16, c29a, c29b, hi64, lo64
    (hi64 = codeword bits 0..63, lo64 = bits 32..95, MSB first)."""
    w = np.asarray(bits91_words, np.uint32).reshape(-1, 3).astype(np.uint64)
    r0, r1, r2 = _rev32(w[:, 0]), _rev32(w[:, 1]), _rev32(w[:, 2])
    hi, lo = (r0 << np.uint64(32)) | r1, (r1 << np.uint64(32)) | r2
    u = np.uint64
    return {"hi64": hi, "lo64": lo,
            "c29a": hi >> u(35), "c29b": (hi >> u(6)) & u(0x1FFFFFFF),
            "g16": (lo >> u(22)) & u(0xFFFF), "i3": (lo >> u(19)) & u(7)}


def fields_bits77(f, idx=None):
    """77-bit payloads as Python ints (all records, or those in idx)."""
    hi, lo = (f["hi64"], f["lo64"]) if idx is None else (f["hi64"][idx], f["lo64"][idx])
    tail = (lo >> np.uint64(19)) & np.uint64(0x1FFF)
    return [(int(h) << 13) | int(t) for h, t in zip(hi.tolist(), tail.tolist())]


def _register(keys_in_order, ents):
    """Apply add_call_hashes for a sequence of call keys with the end state of doing it one by one.
    `ents` is the batch-local {key: (text, hashes)} map (never the global cache, which may be evicted between batches)."""
    if not len(keys_in_order):
        return
    rev = keys_in_order[::-1]
    uk, first_rev = np.unique(rev, return_index=True)          # first in reversed order = last occurrence
    for k in uk[np.argsort(-first_rev, kind="stable")].tolist():   # ascending last occurrence
        text, hs = ents[k]
        for h in hs:
            call_hashes[h] = text
        hashes_for_calls[text] = list(hs)


def unpack_words(bits91_words):
    """Batch unpack() over packed ft8_record.bits91 words, in order, with the reference's hash-history side effects
    (decoders.py:16-115, databases.py:8-26).  Returns a list of message tuples / None."""
    f = payload_fields(bits91_words)
    n = len(f["i3"])
    if n == 0:
        return []
    i3 = f["i3"]
    std = (i3 == 1) | (i3 == 2)
    n28a, n28b = f["c29a"] >> np.uint64(1), f["c29b"] >> np.uint64(1)
    hashed = ((n28a >= _HASH_LO) & (n28a < _HASH_HI)) | ((n28b >= _HASH_LO) & (n28b < _HASH_HI))
    nonzero = (f["hi64"] != 0) | (((f["lo64"] >> np.uint64(19)) & np.uint64(0x1FFF)) != 0)
    seq = nonzero & ((std & hashed) | (i3 == 4))               # history-dependent: scalar, in sequence
    vec = nonzero & std & ~hashed
    is2 = (i3 == 2).astype(np.uint64)
    ka, kb = f["c29a"] * np.uint64(2) + is2, f["c29b"] * np.uint64(2) + is2
    extra = _extra_table()[f["g16"].astype(np.int64)]
    # distinct call fields of the vector part -> cached scalar formatting
    vi = np.flatnonzero(vec)
    uk, inv = np.unique(np.concatenate([ka[vi], kb[vi]]), return_inverse=True)
    utext = np.empty(len(uk), object)
    ureg = np.zeros(len(uk), bool)
    uok = np.zeros(len(uk), bool)
    # the global cache is only evicted BETWEEN batches; everything this batch needs later (registration replay) is read
    # from the batch-local map `ents`, so an eviction can never pull an entry out from under a batch
    if len(_CALL_CACHE) > _CALL_CACHE_MAX:
        _CALL_CACHE.clear()
    ents = {}
    for j, k in enumerate(uk.tolist()):
        ent = _CALL_CACHE.get(k)
        if ent is None:
            ent = _CALL_CACHE[k] = _call29_pure(k >> 1, k & 1)
        ents[k] = ent
        utext[j], ureg[j], uok[j] = ent[0], ent[1] is not None, ent[0] is not None
    ia, ib = inv[:len(vi)], inv[len(vi):]
    rega, regb = ureg[ia], ureg[ib]
    g = f["g16"][vi].astype(np.int64)
    evaluated = _EXTRA_EVAL[g]
    ok = uok[ia] & uok[ib] & _EXTRA_OK[g]
    res = np.empty(n, object)
    res[vi[ok]] = np.fromiter(zip(utext[ia[ok]].tolist(), utext[ib[ok]].tolist(), extra[vi[ok]].tolist()), object, int(ok.sum()))
    out = res.tolist()
    # registrations of the vector part, as (record position, key) in evaluation order: a then b of each record
    ra, rb = rega & evaluated, regb & evaluated
    pos = np.concatenate([2 * vi[ra], 2 * vi[rb] + 1])
    keys = np.concatenate([ka[vi][ra], kb[vi][rb]])
    order = np.argsort(pos, kind="stable")
    pos, keys = pos[order], keys[order]
    si = np.flatnonzero(seq)
    if len(si) == 0:
        # no history-dependent record in the batch: the end state only depends on the LAST registration of every distinct
        # call, found in O(n) (fancy assignment keeps the last write) instead of sorting all registrations again
        uid = np.concatenate([ia[ra], ib[rb]])[order]
        if len(uid):
            last = np.full(len(uk), -1, np.int64)
            last[uid] = pos
            live = np.flatnonzero(last >= 0)
            for j in live[np.argsort(last[live], kind="stable")].tolist():
                text, hs = ents[int(uk[j])]
                for h in hs:
                    call_hashes[h] = text
                hashes_for_calls[text] = list(hs)
        return out
    cuts = np.searchsorted(pos, 2 * si)
    b77 = fields_bits77(f, si)
    lo = 0
    for c, i, b in zip(cuts.tolist(), si.tolist(), b77):
        _register(keys[lo:c], ents)
        lo = c
        out[i] = unpack(b)
    _register(keys[lo:], ents)
    return out


def unpack_many(payloads):
    """unpack() over a sequence of 77-bit payload ints, in order (thin wrapper over unpack_words)."""
    payloads = [int(b) for b in payloads]
    if not payloads:
        return []
    w = np.zeros((len(payloads), 3), np.uint32)
    for i, b in enumerate(payloads):
        v = int(format(b, "077b")[::-1], 2)                 # codeword bit j (MSB-first index) -> bit j of the LSB-first packing
        w[i] = (v & 0xFFFFFFFF, (v >> 32) & 0xFFFFFFFF, v >> 64)
    return unpack_words(w)
