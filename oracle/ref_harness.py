"""Drive the UNMODIFIED reference (G1OJS/PyFT8 under /root/reference) offline.

TEST INFRASTRUCTURE ONLY.  This module exists to (a) pin the numpy restatement in
``oracle/ft8_oracle.py`` against the real reference and (b) generate the golden
vectors committed under ``tests/golden/`` (see ``oracle/make_golden.py``).  It is
only usable in the authoring container: ``/root/reference`` does not exist on the
GPU box, and nothing in ``pyft8_b200/`` may import it.

Recipe = SURVEY.md §8c:
  * stub ``pyaudio`` and ``paho.mqtt.client`` (absent here, imported at module
    import by receiver.py:3 / pskreporter.py:1),
  * fake clock in ``time_utils`` (time_utils.py:7-14), no threads
    (receiver.py:252, receiver.py:336),
  * feed 375 blocks of 480 int16 samples through ``AudioIn._callback``
    (receiver.py:295-306) so grid row h = window ending at sample 480*h,
  * ``Receiver.search`` (receiver.py:338-367) then 8 rounds of
    ``Candidate.decode`` in ``manage_cycle`` order (receiver.py:389-398).
"""
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("PYFT8_REF_ROOT", "/root/reference")


class _FakeClock:
    def __init__(self):
        self.now = 0.0


def load_reference():
    """Import PyFT8.receiver / decoders / transmitter with the two import stubs."""
    if "PyFT8.receiver" in sys.modules:
        return (sys.modules["PyFT8.receiver"], sys.modules["PyFT8.decoders"],
                sys.modules["PyFT8.transmitter"], sys.modules["PyFT8.databases"],
                sys.modules["PyFT8.time_utils"])
    if not os.path.isdir(os.path.join(REF_ROOT, "PyFT8")):
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    pa = types.ModuleType("pyaudio")
    pa.paInt16 = 8
    pa.paContinue = 0

    class PyAudio:  # receiver.py:255-256 only needs get_device_count()
        def get_device_count(self):
            return 0
    pa.PyAudio = PyAudio
    sys.modules.setdefault("pyaudio", pa)
    for name in ("paho", "paho.mqtt", "paho.mqtt.client"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    cwd = os.getcwd()
    import PyFT8.time_utils as tu
    import PyFT8.databases as db
    import PyFT8.decoders as dec
    import PyFT8.transmitter as tx
    import PyFT8.receiver as rx
    os.chdir(cwd)

    class _NoThread:
        def __init__(self, *a, **k):
            pass

        def start(self):
            pass
    rx.threading = types.SimpleNamespace(Thread=_NoThread)
    clock = _FakeClock()
    tu.time_utils.time = lambda: clock.now
    tu.time_utils.sleep = lambda t: None
    tu.time_utils._clock = clock
    return rx, dec, tx, db, tu


def decode_cycle(audio_i16, dump=False, t0=30.0 * 1000000, clear_hashes=True, workdir=None):
    """Decode one isolated 15 s cycle with the unmodified reference.

    Returns dict(messages=[...], cands=[...]) and, when ``dump``, the per-stage
    intermediates each kernel is parity-checked against.
    """
    rx, dec, tx, db, tu = load_reference()
    clock = tu.time_utils._clock
    audio_i16 = np.asarray(audio_i16, dtype=np.int16)
    assert audio_i16.shape == (180000,)
    if clear_hashes:
        db.call_hashes.clear()
    cwd = os.getcwd()
    if workdir:  # simple_validate_call appends to ./rejected_callsigns.txt (decoders.py:114)
        os.chdir(workdir)
    try:
        clock.now = t0
        msgs = []
        r = rx.Receiver("", msgs.append)
        ai = r.audio_in
        assert ai.search_grid_ptr == 0
        for k in range(375):
            clock.now = t0 + (k + 1) * 0.04 + 1e-6
            ai._callback(audio_i16[480 * k:480 * (k + 1)].tobytes(), 480, None, None)
        clock.now = t0 + 15.0
        cs = tu.time_utils.cyclestart_string(t0)
        cands = r.search(cs, 0, range(ai.search_f0_idx_range[0], ai.search_f0_idx_range[1]))
        out = {}
        if dump:
            out["grid"] = ai.search_grid[:376].copy()
            out["grid_tail_is_one"] = bool(np.all(ai.search_grid[376:] == 1.0))
            out["cand_f0"] = np.array([c.origin["f0_idx"] for c in cands], np.int32)
            out["cand_h0"] = np.array([c.origin["h0_idx"] for c in cands], np.int32)
            out["cand_score"] = np.array([c.origin["score"] for c in cands], np.float64)
            out["cand_payload"] = np.array([c.payload_on_search_grid for c in cands], np.float32)
        trace = [dict(idx=i) for i in range(len(cands))]
        for i, c in enumerate(cands):
            c._idx = i
        dup = set()
        emit_order = []
        for rnd in range(9):
            todo = [c for c in cands if not c.decode_result]
            if not todo:
                break
            todo.sort(key=lambda c: c.llr_sd, reverse=True)
            for c in todo:
                ip = c.ipass
                c.decode(100)
                t = trace[c._idx]
                if ip == 0:
                    t["grid_sd"] = float(c.llr_sd)
                    t["grid_snr"] = int(c.snr)
                    t["grid_llr"] = np.asarray(c.llr0, np.float32).copy()
                if ip == 1:
                    t["tweaks"] = c.tweaks
                    t["nsync"] = int(c.n_sync_matches)
                    if c.decode_result != "stop":
                        t["fine_sd"] = float(c.llr_sd)
                        t["fine_snr"] = int(c.snr)
                        t["fine_llr"] = np.asarray(c.llr, np.float32).copy()
                        t["fine_grid"] = np.asarray(c.signal_grid, np.float32).copy()
                if c.decode_result is not None:
                    t["final_ipass"] = ip
                    t["result"] = c.decode_result
                    if c.decode_result != "stop":
                        t["notes"] = c.decode_notes
                        t["tsec"] = float(c.origin["tsec"])
                        t["fHz"] = float(c.origin["fHz"])
                        t["snr"] = int(c.snr)
                        n0 = len(msgs)
                        c.check_and_package(dup)
                        t["emitted"] = len(msgs) > n0
                        if t["emitted"]:
                            emit_order.append(c._idx)
        out["messages"] = msgs
        out["trace"] = trace
        out["emit_order"] = emit_order
        out["n_cands"] = len(cands)
        return out
    finally:
        os.chdir(cwd)


def decode_two_cycles(audio_a, audio_b, t0=30.0 * 1000000, workdir=None):
    """Two CONSECUTIVE cycles through one unmodified Receiver, the way a live receiver sees them (receiver.py:295-306,
    338-367): A fills ring rows 1..375 and is searched / decoded as the even cycle at its end; B's hops go on into rows
    376..749, 0 -- its first windows reach back into A's last samples and its early candidates' payload rows wrap into A's
    half -- and it is searched / decoded as the odd cycle.  Returns [messages of A, messages of B] (lists of dicts)."""
    rx, dec, tx, db, tu = load_reference()
    clock = tu.time_utils._clock
    db.call_hashes.clear()
    cwd = os.getcwd()
    if workdir:
        os.chdir(workdir)
    try:
        clock.now = t0
        out = []
        msgs = []
        r = rx.Receiver("", msgs.append)
        ai = r.audio_in
        assert ai.search_grid_ptr == 0
        for half, audio in enumerate((audio_a, audio_b)):
            audio = np.asarray(audio, dtype=np.int16)
            base = t0 + 15.0 * half
            for k in range(375):
                clock.now = base + (k + 1) * 0.04 + 1e-6
                ai._callback(audio[480 * k:480 * (k + 1)].tobytes(), 480, None, None)
            clock.now = base + 15.0
            cs = tu.time_utils.cyclestart_string(base)
            cands = r.search(cs, half, range(ai.search_f0_idx_range[0], ai.search_f0_idx_range[1]))
            n0 = len(msgs)
            dup = set()
            for rnd in range(9):
                todo = [c for c in cands if not c.decode_result]
                if not todo:
                    break
                todo.sort(key=lambda c: c.llr_sd, reverse=True)
                for c in todo:
                    c.decode(100)
                    if c.decode_result is not None and c.decode_result != "stop":
                        c.check_and_package(dup)
            out.append(dict(messages=msgs[n0:], n_cands=len(cands),
                            cand_f0=np.array([c.origin["f0_idx"] for c in cands], np.int32),
                            cand_h0=np.array([c.origin["h0_idx"] for c in cands], np.int32)))
        return out
    finally:
        os.chdir(cwd)


def decode_cycle_progressive(audio_i16, extra_hops=120, t0=30.0 * 1000000, workdir=None):
    """One cycle through the unmodified reference the way its scheduler THREAD sees it (receiver.py:376-412), made
    deterministic: after every 480-sample hop the body of manage_cycle's loop runs exactly once -- reset the searched flag
    at a cycle start, decode every candidate whose payload rows are complete (sorted by llr_sd), search once the hop
    counter passes search_start_hop (hop 260 = 10.4 s).  `extra_hops` hops of silence follow the cycle so that the late
    candidates get their turn.  Returns dict(messages, emit_hop) -- emit_hop[i] = hop after which message i came out."""
    rx, dec, tx, db, tu = load_reference()
    clock = tu.time_utils._clock
    audio_i16 = np.asarray(audio_i16, dtype=np.int16)
    db.call_hashes.clear()
    cwd = os.getcwd()
    if workdir:
        os.chdir(workdir)
    try:
        clock.now = t0
        msgs, emit_hop = [], []
        r = rx.Receiver("", msgs.append)
        ai = r.audio_in
        dup = set()
        prev = 0
        searched = False
        silence = np.zeros(480, np.int16).tobytes()
        n_cands = 0
        for k in range(375 + extra_hops):
            clock.now = t0 + (k + 1) * 0.04 + 1e-6
            ai._callback(audio_i16[480 * k:480 * (k + 1)].tobytes() if k < 375 else silence, 480, None, None)
            pos = ai.search_grid_ptr % ai.search_hops_per_cycle
            if pos < prev:
                searched = False
            prev = pos
            todo = [c for c in r.candidates if (not c.decode_result) and
                    (not (c.search_grid_bounds[0] <= ai.search_grid_ptr <= c.search_grid_bounds[1]))]
            if todo:
                todo.sort(key=lambda c: c.llr_sd, reverse=True)
                max_ipass = 10 + min(c.ipass for c in todo)
                for c in todo:
                    c.decode(max_ipass)
                    if c.decode_result is not None and c.decode_result != "stop":
                        n0 = len(msgs)
                        c.check_and_package(dup)
                        emit_hop += [k + 1] * (len(msgs) - n0)
            if not searched and pos > r.search_start_hop:
                cs = tu.time_utils.cyclestart_string(clock.now)
                r.candidates = r.search(cs, tu.time_utils.odd_even(), range(ai.search_f0_idx_range[0], ai.search_f0_idx_range[1]))
                n_cands = len(r.candidates)
                searched = True
        return dict(messages=msgs, emit_hop=emit_hop, n_cands=n_cands, pending=sum(1 for c in r.candidates if not c.decode_result))
    finally:
        os.chdir(cwd)


def read_wav_i16(path):
    import wave
    w = wave.open(path, "rb")
    assert w.getframerate() == 12000 and w.getnchannels() == 1 and w.getsampwidth() == 2
    x = np.frombuffer(w.readframes(180000), dtype=np.int16)
    w.close()
    a = np.zeros(180000, np.int16)
    a[:len(x)] = x
    return a


if __name__ == "__main__":
    import tempfile
    import time
    for name in ("test_08.wav", "test_09.wav"):
        a = read_wav_i16(os.path.join(REF_ROOT, "tests", "pipeline", name))
        t = time.time()
        with tempfile.TemporaryDirectory() as d:
            o = decode_cycle(a, workdir=d)
        print(name, o["n_cands"], "cands", len(o["messages"]), "decodes", f"{time.time()-t:.1f}s")
        for m in o["messages"]:
            print("   ", m["all_txt_format"], "|", m["decode_notes"])
