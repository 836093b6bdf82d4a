#!/usr/bin/env python
"""bench.py -- decoded 15-s FT8 cycles/s of the receive hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU algorithm (oracle port) on the host cores

A step = one pass of the whole path (audio -> records) over one batch of synthetic cycles (BASELINE configs[1]:
4096 cycles x 50 GFSK signals, SNR -24..+10 dB, 200-2950 Hz).  `value` is timed with the batch resident in HBM;
`e2e` is the same through Engine.decode_cycles with pinned HOST buffers (H2D of the audio and D2H of the records inside
the timed region).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "decoded 15-s FT8 cycles/sec"
UNIT = "cycles/s"
WORKLOAD = "cfg2_50sig"          # BASELINE configs[1]; --workload selects configs[0] / configs[3] shapes for extra data points
# algorithmic bytes per unit (DESIGN.md "Roofline accounting"); int16 audio in
ALG_BYTES = {
    "spectrogram": 360000 + 375 * 976 * 4,          # per cycle: audio once + grid once
    "sync": 375 * 976 * 4 + 200 * (16 + 58 * 8 * 4),  # per cycle: grid once + candidates/payloads
    "cycle_spectrum": 360000 + 47415 * 8,           # per cycle: audio once + the bins the fine stage reads
    "pass0_ldpc5": 58 * 8 * 4 + 696 + 24,           # per candidate
    "fine": 9 * 8000 + 58 * 8 * 4,                  # per fine evaluation (candidate)
    "pass234_ldpc": 696 + 696 + 24,                 # per LDPC call
    "osd": 696 + 64,                                # per OSD call
}
# measured DRAM traffic per cycle (bytes) of each stage's kernels: dram__bytes_read.sum + dram__bytes_write.sum of one
# `ncu --set full` capture at 4096 cycles/launch, divided by 4096 (profiles/r01i_summary.md; osd from the earlier r01e capture, its kernel is unchanged)
NCU_DRAM_BYTES_PER_CYCLE = {"spectrogram": (1477e6 + 5959e6) / 4096, "sync": (2328e6 + 49e6 + 46e6) / 4096, "fine": (2892e6 + 492e6) / 4096,
                            "pass0_ldpc5": (3733e6 + 616e6) / 4096, "osd": (541e6 + 25e6) / 4096,
                            "cycle_spectrum": (787e6 + 370e6 + 369e6 + 732e6) / 1024}
# issue-slot / pipe utilisation of the same captures (percent of peak): what actually bounds the non-HBM stages.
# l1_data_pipe = l1tex__data_pipe_lsu_wavefronts (shared-memory + L1 wavefronts): the binding resource of the FFT kernels.
NCU_PIPES = {"spectrogram": {"issue_active": 72.3, "fma_pipe": 31.7, "alu_pipe": 37.7, "l1_data_pipe": 86.4},
             "sync": {"issue_active": 75.2, "fma_pipe": 25.9, "alu_pipe": 43.4, "l1_data_pipe": 85.9},
             "cycle_spectrum": {"issue_active": 45.1, "fma_pipe": 16.4, "alu_pipe": 30.7, "l1_data_pipe": 88.3},
             "fine": {"issue_active": 48.0, "fma_pipe": 22.9, "alu_pipe": 19.8, "l1_data_pipe": 64.0},
             "pass0_ldpc5": {"issue_active": 56.7, "fma_pipe": 21.4, "alu_pipe": 24.0, "l1_data_pipe": 71.1},
             "osd": {"issue_active": 67.2, "fma_pipe": 6.5, "alu_pipe": 84.1}}
STAGES = ["all", "spectrogram", "sync", "cycle_spectrum", "pass0_ldpc5", "fine", "pass234_ldpc", "osd", "collect"]


def _peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def _dist_init():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # no data-path collective exists (cycles are independent); gloo carries the barrier and the max-over-ranks
        dist_mod.init_process_group("gloo", rank=rank, world_size=world)
        dist = dist_mod
    return dist, rank, local, world


def _cpu_decode_one(a):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ft8_oracle as o
    t = time.time()
    recs, cl = o.decode_cycle(a)
    return len(recs), len(cl), sum(c.n_ldpc for c in cl), sum(c.n_osd for c in cl), time.time() - t


_POOL = None


def _cpu_warm(_):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ft8_oracle  # noqa: F401
    return os.getpid()


def cpu_pool(cores):
    """Worker processes (one per host core) with the oracle imported, so that timing excludes process start-up."""
    global _POOL
    if _POOL is None:
        import multiprocessing as mp
        _POOL = mp.get_context("spawn").Pool(cores)
        _POOL.map(_cpu_warm, range(4 * cores), chunksize=1)
    return _POOL


def _host_cycle0(p):
    sys.path.insert(0, ROOT)
    from pyft8_b200 import workload
    p = dict(p)
    p["seed"] = (p["seed"] << 8) ^ int(p["pick"][0, 0]) ^ (int(p["pick"][0, 1]) << 10)     # distinct noise per cycle
    return workload.host_cycle(p, 0)


def cpu_rate(cycles, cores):
    """The oracle port on `cores` worker processes, one cycle at a time each: (cycles/s, ldpc calls/s, detail)."""
    pool = cpu_pool(cores)
    t = time.time()
    res = pool.map(_cpu_decode_one, list(cycles), chunksize=1)
    wall = time.time() - t
    return len(cycles) / wall, sum(r[2] for r in res) / wall, dict(wall_s=wall, decodes=sum(r[0] for r in res),
                                                                    per_cycle_s=float(np.mean([r[4] for r in res])))


def run_reference(args, dist, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference itself is Python
    and does not travel to the GPU box), all host cores, bounded sample of the same workload per step."""
    if rank != 0:
        return
    from pyft8_b200 import workload
    cores = os.cpu_count() or 1
    per_step = cores * args.ref_cycles_per_core
    n_cyc = per_step * (args.steps + args.warmup)
    params = workload.make_params(WORKLOAD, n_cyc, seed=args.seed)
    # the numpy modulator needs ~0.8 s per 50-signal cycle: build the inputs on the worker pool too (outside the timed region)
    small = [{k: (v[b:b + 1] if isinstance(v, np.ndarray) else v) for k, v in params.items() if k != "pool_bits77"} for b in range(n_cyc)]
    cyc = cpu_pool(cores).map(_host_cycle0, small, chunksize=1)
    for w in range(args.warmup):
        cpu_rate(cyc[w * per_step:(w + 1) * per_step], cores)
    t0 = time.time()
    ldpc = 0.0
    for s in range(args.warmup, args.warmup + args.steps):
        r, l, _ = cpu_rate(cyc[s * per_step:(s + 1) * per_step], cores)
        ldpc += l
    wall = time.time() - t0
    value = per_step * args.steps / wall
    sample = f"{per_step} cycles per step ({args.ref_cycles_per_core} per core) of {WORKLOAD}, numpy-generated from the same parameters"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "cycles_per_step": per_step, "signals_per_cycle": workload.CONFIGS[WORKLOAD][0], "host": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "ldpc_codewords_per_sec": ldpc / args.steps,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def run_gpu(args, dist, rank, local, world):
    import torch
    from pyft8_b200 import workload
    from pyft8_b200 import _lib as L
    from pyft8_b200.engine import Engine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    B = args.cycles
    eng = Engine(device=local, max_cycles=B)
    params = workload.make_params(WORKLOAD, B, seed=args.seed + 7919 * rank)
    audio = torch.empty((B, 180000), dtype=torch.int16, device=f"cuda:{local}")
    workload.device_cycles(eng, params, audio.data_ptr())
    torch.cuda.synchronize()
    host = torch.empty((B, 180000), dtype=torch.int16, pin_memory=True)
    host.copy_(audio)
    host_np = host.numpy()
    stream = torch.cuda.ExternalStream(L.load().ft8_stream(eng._h), device=f"cuda:{local}")
    rec = np.zeros(B * eng.max_cands, L.RECORD_DTYPE)
    nrec = np.zeros(B, np.int32)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def sum_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    # ---- device-resident timing
    for _ in range(args.warmup):
        eng.decode_cycles_dev(audio.data_ptr(), L.AUDIO_I16, B, rec=rec, n=nrec)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = np.zeros(9)
    launches = 0
    stats = None
    e0.record(stream)
    for _ in range(args.steps):
        r, n = eng.decode_cycles_dev(audio.data_ptr(), L.AUDIO_I16, B, rec=rec, n=nrec)
        stage_ms += [eng.last_kernel_ms(i) for i in range(9)]
        stats = eng.stats()
        launches += stats["kernel_launches"]
    e1.record(stream)
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    n_decoded, n_emitted = stats["decoded"], stats["emitted"]
    # ---- end to end: pinned host audio in, records out, every step, through Engine.decode_cycles (C ABI, host buffers).
    # Batches are streamed the way a skimmer would: each call names the next batch (ft8_decode_cycles_stream), whose PCIe
    # copy then runs on a second CUDA stream underneath this batch's kernels.  Every step's H2D copy and record D2H are
    # inside the timed region.
    # caller-owned pinned output buffers (records, per-cycle counts), reused every step
    rec_pin = torch.zeros((B * eng.max_cands, L.RECORD_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
    n_pin = torch.zeros(B, dtype=torch.int32).pin_memory()
    rec_np, n_np = rec_pin.numpy().view(L.RECORD_DTYPE).reshape(-1), n_pin.numpy()

    def e2e_steps(k):
        r_i = None
        # step 0 is not prefetched: its copy runs in chunks with the front-end kernels starting as chunks land (and is
        # inside the timed region like every other step's); from step 1 on the copy hides under the previous step
        for i in range(k):
            r_i, _ = eng.decode_cycles(host_np, next_audio=host_np if i + 1 < k else None, rec=rec_np, n=n_np)
        return r_i

    e2e_steps(min(args.warmup, 2))
    barrier()
    t0 = time.perf_counter()
    rec_last = e2e_steps(args.steps)
    barrier()
    ms_e2e = max_over_ranks(1e3 * (time.perf_counter() - t0))
    n_e2e_decoded = len(rec_last)
    # host text formatting of one step's records (outside every timed region): vectorised unpack + de-dup
    from pyft8_b200.receiver import format_records
    format_records(rec_last[:1000])
    t0 = time.perf_counter()
    mb = format_records(rec_last)
    host_text = {"records_per_sec": len(rec_last) / max(time.perf_counter() - t0, 1e-9), "records": int(len(rec_last)),
                 "messages": int(len(mb)), "note": "format_records on one step's records, one host thread, not in any timed region"}
    total_cycles = sum_over_ranks(B)
    value = total_cycles * args.steps / (ms_dev / 1e3)
    e2e = total_cycles * args.steps / (ms_e2e / 1e3)
    stage_ms /= args.steps
    ldpc_rate = sum_over_ranks(stats["ldpc_calls"]) / (ms_dev / args.steps / 1e3)
    if rank != 0:
        return
    peak, peak_src = _peaks()
    units = {"spectrogram": B, "sync": B, "cycle_spectrum": B, "pass0_ldpc5": stats["candidates"], "fine": stats["fine_evals"],
             "pass234_ldpc": max(stats["ldpc_calls"], 1), "osd": max(stats["osd_calls"], 1)}
    stages = []
    for i, name in enumerate(STAGES):
        if name in ALG_BYTES:
            ach = ALG_BYTES[name] * units[name] / (stage_ms[i] / 1e3) / 1e9 if stage_ms[i] > 0 else 0.0
            traffic = NCU_DRAM_BYTES_PER_CYCLE.get(name)
            stages.append({"kernel": name, "ms": round(float(stage_ms[i]), 4), "share": round(float(stage_ms[i] / stage_ms[0]), 4),
                           "bound": "hbm", "achieved": round(ach, 2), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 5),
                           "traffic": int(traffic * B) if traffic else None, "ncu_pct_of_peak": NCU_PIPES.get(name)})
    dom = max(stages, key=lambda s: s["ms"])
    roofline = {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "traffic": dom["traffic"], "peak_source": peak_src,
                "note": "dominant kernel by device time; LDPC/OSD/fine-sync stages are issue/ALU-bound, see DESIGN.md; "
                        "spectrogram and sync (the HBM-roofline stages of the north star) are in roofline_stages"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = [host_np[i].copy() for i in range(cores)]
        v, l, det = cpu_rate(sample, cores)
        # parity spot check on the sample is done by tests/smoke; here only the rate is reported
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {cores} cycles of this run's batch (one per core), oracle port, {det['wall_s']:.1f} s wall",
               "ldpc_codewords_per_sec": l}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "cycles_per_gpu_per_step": B, "signals_per_cycle": workload.CONFIGS[WORKLOAD][0],
                   "snr_db": list(workload.CONFIGS[WORKLOAD][1]),
                   "audio": "int16 12 kHz 15 s", "l2": "inputs_larger_than_l2 (%.2f GB per step per GPU)" % (B * 360000 / 1e9),
                   "parallelism": f"cycles sharded over {world} GPU(s), no collective"},
        "ldpc_codewords_per_sec": ldpc_rate,
        "decodes_per_cycle": n_decoded / B, "emitted_per_cycle": n_emitted / B,
        "work_per_step": {k: stats[k] for k in ("candidates", "stopped_sd", "fine_evals", "fine_pass", "ldpc_calls", "ldpc_iters", "osd_calls")},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(B * 360000), "d2h_bytes_per_step": int(n_e2e_decoded * 64 + 4 * B + 96),
                "ms_per_step": ms_e2e / args.steps, "api": "Engine.decode_cycles(pinned host int16, next_audio=...) -> ft8_decode_cycles_stream: one handle, next batch copied under the kernels"},
        "host_text": host_text, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_stages": stages, "cpu_baseline": cpu,
    }
    print(json.dumps(out), flush=True)


def main():
    global WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cycles", type=int, default=4096, help="cycles per GPU per step (BASELINE configs[1]: 4096)")
    ap.add_argument("--seed", type=int, default=2000)
    ap.add_argument("--workload", default=WORKLOAD, choices=["cfg1_20sig", "cfg2_50sig", "cfg4_120sig"],
                    help="cfg2_50sig is the headline configuration (BASELINE configs[1])")
    ap.add_argument("--ref-cycles-per-core", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    WORKLOAD = args.workload
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    dist, rank, local, world = _dist_init()
    try:
        if args.impl == "reference":
            run_reference(args, dist, rank, world)
        else:
            run_gpu(args, dist, rank, local, world)
    finally:
        if _POOL is not None:
            _POOL.terminate()
        if dist:
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
