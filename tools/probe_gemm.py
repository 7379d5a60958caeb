"""TF32 tcgen05 GEMM throughput sweep (CUDA events, warm): python tools/probe_gemm.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qbm_b200

dev = torch.device("cuda:0")
torch.manual_seed(0)
for M, N, K in [(256, 500, 784), (256, 784, 500), (784, 500, 256), (2048, 500, 784), (8192, 500, 784), (8192, 784, 500),
                (784, 500, 8192), (4096, 4096, 4096), (8192, 8192, 4096)]:
    A = torch.randn(M, K, device=dev)
    B = torch.randn(N, K, device=dev)
    for _ in range(3):
        C = qbm_b200.gemm_tf32(A, B)
    it = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(it):
        C = qbm_b200.gemm_tf32(A, B)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    ref = A @ B.t()
    err = float((C - ref).abs().max() / ref.abs().max())
    print(f"M={M} N={N} K={K}: {ms * 1e3:9.1f} us  {2.0 * M * N * K / ms / 1e9:8.2f} TFLOP/s  max rel err {err:.2e}", flush=True)
