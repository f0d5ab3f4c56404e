"""H2D bandwidth of one step's crops (1024 x 369 x 11 x 11 float32) from differently allocated page-locked host buffers.
Usage (GPU box): python tools/h2d_probe.py"""
import ctypes
import time

import numpy as np
import torch

N = 1024 * 369 * 121
dev = torch.device("cuda", 0)
dst = torch.empty(N, dtype=torch.float32, device=dev)


def bench(src, label, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        s.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        s.synchronize()
        dt = time.perf_counter() - t0
    print(f"{label:40s} {N * 4 * reps / dt / 1e9:7.2f} GB/s  pinned={src.is_pinned()}")


a = torch.rand(N).pin_memory()
bench(a, "torch pin_memory()")
try:
    from cuda.bindings import runtime as cudart
except Exception:
    from cuda import cudart
for flags, name in ((cudart.cudaHostAllocDefault, "cudaHostAlloc default"), (cudart.cudaHostAllocWriteCombined, "cudaHostAlloc write-combined"),
                    (cudart.cudaHostAllocPortable, "cudaHostAlloc portable")):
    err, ptr = cudart.cudaHostAlloc(N * 4, flags)
    assert int(err) == 0, err
    buf = (ctypes.c_float * N).from_address(int(ptr))
    arr = np.frombuffer(buf, dtype=np.float32)
    arr[:] = 0.5
    t = torch.from_numpy(arr)
    bench(t, name)
    cudart.cudaFreeHost(ptr)
# two half-size copies on two streams (does a second DMA queue help?)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h = N // 2
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1):
        dst[:h].copy_(a[:h], non_blocking=True)
    with torch.cuda.stream(s2):
        dst[h:].copy_(a[h:], non_blocking=True)
torch.cuda.synchronize()
print(f"{'two streams, half each':40s} {N * 4 * 20 / (time.perf_counter() - t0) / 1e9:7.2f} GB/s")
