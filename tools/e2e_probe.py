import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pyft8_b200 import workload, _lib as L
from pyft8_b200.engine import Engine
B = 4096
eng = Engine(0, max_cycles=B)
prm = workload.make_params("cfg2_50sig", B, seed=2)
audio = torch.empty((B, 180000), dtype=torch.int16, device="cuda:0")
workload.device_cycles(eng, prm, audio.data_ptr()); eng.synchronize()
host = torch.empty((B, 180000), dtype=torch.int16).pin_memory(); host.copy_(audio); torch.cuda.synchronize()
host_np = host.numpy()
rec_pin = torch.zeros((B * eng.max_cands, L.RECORD_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
n_pin = torch.zeros(B, dtype=torch.int32).pin_memory()
rec_np, n_np = rec_pin.numpy().view(L.RECORD_DTYPE).reshape(-1), n_pin.numpy()
for _ in range(2): eng.decode_cycles_dev(audio.data_ptr(), L.AUDIO_I16, B, rec=rec_np, n=n_np)
t0 = time.perf_counter(); eng.decode_cycles_dev(audio.data_ptr(), L.AUDIO_I16, B, rec=rec_np, n=n_np); t1 = time.perf_counter()
print("device-resident call: wall %.2f ms, device %.2f ms" % (1e3 * (t1 - t0), eng.last_kernel_ms(0)))
eng.prefetch(host_np)
for i in range(4):
    t0 = time.perf_counter(); eng.decode_cycles(host_np, next_audio=host_np, rec=rec_np, n=n_np); t1 = time.perf_counter()
    print("streamed call: wall %.2f ms, device %.2f ms, stages %s" % (1e3 * (t1 - t0), eng.last_kernel_ms(0), [round(eng.last_kernel_ms(k), 2) for k in range(1, 9)]))
t0 = time.perf_counter(); eng.decode_cycles(host_np, rec=rec_np, n=n_np); t1 = time.perf_counter()
print("last (prefetched, no next): wall %.2f ms, device %.2f ms" % (1e3 * (t1 - t0), eng.last_kernel_ms(0)))
# raw host->device rate of the same pinned buffer (is the streamed call copy-bound?)
dst = torch.empty_like(audio)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(host, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("H2D 1.47 GB: %.2f ms = %.1f GB/s" % (1e3 * (t1 - t0), host.numel() * 2 / (t1 - t0) / 1e9))
small = torch.empty((17_000_000,), dtype=torch.uint8).pin_memory(); dsm = torch.empty_like(small, device="cuda:0")
torch.cuda.synchronize(); t0 = time.perf_counter(); small.copy_(dsm, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print("D2H 17 MB: %.2f ms" % (1e3 * (t1 - t0)))
