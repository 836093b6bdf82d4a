"""Synthetic workloads of BASELINE.json: parameters on the host, waveforms on the device or host."""
import math

import numpy as np

from . import synth

CONFIGS = {
    # name: (signals per cycle, SNR range dB (2500 Hz), carrier range Hz, dt range s)
    "cfg1_20sig": (20, (-20.0, 5.0), (200.0, 2950.0), (-0.5, 1.0)),
    "cfg2_50sig": (50, (-24.0, 10.0), (200.0, 2950.0), (-0.5, 1.0)),
    "cfg4_120sig": (120, (-24.0, 10.0), (200.0, 2950.0), (-0.5, 1.0)),
}
NOISE_SIGMA = 1000.0


def make_params(config, n_cycles, seed, pool=1024):
    """Per-signal parameters for n_cycles cycles: symbols [B,S,79] u8, f_hz/dt_s/amp [B,S] f32, bits77 list of the pool."""
    n_sig, snr_r, f_r, dt_r = CONFIGS[config]
    rng = np.random.default_rng(seed)
    msgs = [synth.pack77(*synth.random_message(rng)) for _ in range(pool)]
    table = np.array([synth.symbols_from_bits77(b) for b in msgs], np.uint8)
    pick = rng.integers(0, pool, (n_cycles, n_sig))
    snr = rng.uniform(snr_r[0], snr_r[1], (n_cycles, n_sig))
    amp = NOISE_SIGMA * np.sqrt(2.0 * (2500.0 / 6000.0) * 10.0 ** (snr / 10.0))
    return dict(symbols=table[pick], f_hz=rng.uniform(f_r[0], f_r[1], (n_cycles, n_sig)).astype(np.float32),
                dt_s=rng.uniform(dt_r[0], dt_r[1], (n_cycles, n_sig)).astype(np.float32), amp=amp.astype(np.float32),
                snr=snr, pick=pick, pool_bits77=msgs, seed=seed)


def host_cycle(params, b):
    """Cycle b of the workload built on the host with the numpy modulator (used by the CPU reference arm)."""
    rng = np.random.default_rng((params["seed"] << 20) + b)
    x = rng.normal(0.0, NOISE_SIGMA, 180000)
    for s in range(params["symbols"].shape[1]):
        wf = np.imag(synth.shift_carrier(synth.gfsk_baseband(list(params["symbols"][b, s])), float(params["f_hz"][b, s])))
        s0 = int((0.5 + float(params["dt_s"][b, s])) * 12000)
        lo, hi = max(s0, 0), min(s0 + len(wf), 180000)
        x[lo:hi] += float(params["amp"][b, s]) * wf[lo - s0:hi - s0]
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def device_cycles(engine, params, out_ptr, chunk=512):
    """Fill int16 audio [B,180000] at device pointer out_ptr with the workload, using the library's generator kernel."""
    B = params["symbols"].shape[0]
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        engine.synth_cycles(params["symbols"][b0:b1], params["f_hz"][b0:b1], params["dt_s"][b0:b1], params["amp"][b0:b1],
                            NOISE_SIGMA, seed=(params["seed"] << 20) + b0, out_ptr=out_ptr + b0 * 180000 * 2)
