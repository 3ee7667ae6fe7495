#!/bin/bash
# Runs on the GPU box: FP64 peaks (DMMA / DFMA / cuBLAS DGEMM), host<->device bandwidth, host CPU info.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/peaks_clocks.csv &
SMI=$!
./tools/fp64_peaks > gpurun_out/fp64_peaks.json 2> gpurun_out/fp64_peaks.err
python - <<'PY' > gpurun_out/dgemm_peaks.json 2> gpurun_out/dgemm_peaks.err
import torch, time, json, os
r = {}
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2): c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
r["cublas_dgemm_8192_tflops"] = 2 * n**3 / best / 1e9
# skinny shape like the Schur update: (200 x 2000) @ (2000 x 1e6)
P = torch.randn(200, 2000, dtype=torch.float64, device="cuda"); G = torch.randn(2000, 1000000, dtype=torch.float64, device="cuda")
for _ in range(2): c = P @ G
torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = P @ G; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
r["cublas_dgemm_200x2000x1e6_tflops"] = 2 * 200 * 2000 * 1e6 / best / 1e9
r["cublas_dgemm_200x2000x1e6_ms"] = best
del G, c
h = torch.empty(1 << 28, dtype=torch.uint8).pin_memory(); d = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
for name, (src, dst) in {"h2d": (h, d), "d2h": (d, h)}.items():
    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    r[name + "_pinned_gbs"] = (1 << 28) / (time.perf_counter() - t) / 1e9
x = torch.zeros(1, device="cuda"); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(200): x.add_(1); torch.cuda.synchronize()
r["launch_sync_us"] = (time.perf_counter() - t) / 200 * 1e6
r["cpu_count"] = os.cpu_count()
r["gpus"] = torch.cuda.device_count()
print(json.dumps(r))
PY
kill $SMI
lscpu | head -25 > gpurun_out/lscpu.txt
free -g > gpurun_out/free.txt
cat gpurun_out/fp64_peaks.json gpurun_out/dgemm_peaks.json
