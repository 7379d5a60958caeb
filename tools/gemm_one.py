"""One warm launch of the TF32 tcgen05 GEMM at a square size (for an ncu capture): python tools/gemm_one.py 8192"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, qbm_b200
dev = torch.device("cuda:0")
M = N = K = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
for _ in range(2):
    C = qbm_b200.gemm_tf32(A, B)
torch.cuda.synchronize()
