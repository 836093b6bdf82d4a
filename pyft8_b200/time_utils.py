"""Cycle arithmetic used by the receiver façade (mirrors the parts of PyFT8/time_utils.py:3-34 the path needs).

The clock is injectable so that tests and offline decoding are deterministic."""
import time as _time


class TimeUtils:
    def __init__(self, clock=None, sleep_fn=None):
        self.cycle_seconds = 15
        self._clock = clock or _time.time
        self._sleep_fn = sleep_fn

    def time(self):
        return self._clock()

    def sleep(self, t):
        """Real clock: sleep t.  Injected clock: call sleep_fn(t) if one was given, else yield the CPU for at most 10 ms
        of real time (a polling loop on a fake clock must neither spin nor stall the test that advances the clock)."""
        if t <= 0:
            return
        if self._sleep_fn is not None:
            self._sleep_fn(t)
        elif self._clock is _time.time:
            _time.sleep(t)
        else:
            _time.sleep(min(t, 0.01))

    def set_cycle_length(self, dur):
        self.cycle_seconds = dur

    def cycle_time(self):
        return self.time() % self.cycle_seconds

    def grid_time(self):
        return self.time() % (2 * self.cycle_seconds)

    def odd_even(self):
        return int(self.grid_time() / self.cycle_seconds)

    def cyclestart_string(self, t):
        cst = self.cycle_seconds * int(t / self.cycle_seconds)
        return _time.strftime("%y%m%d_%H%M%S", _time.gmtime(cst))

    def tlog(self, txt, verbose=False):
        if verbose:
            print(f"{self.cyclestart_string(self.time())} {self.cycle_time():5.2f} {txt}")
