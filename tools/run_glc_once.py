"""One launch of the cluster GEMM + LayerNorm kernel per shape (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvuld_b200 import _lib
dev = 'cuda'
K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
M, N = 50176, 512
A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
b = torch.randn(N, device=dev); g = torch.ones(N, device=dev); be = torch.zeros(N, device=dev)
x32 = torch.randn(M, N, device=dev); xb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    _lib.gemm_ln_wide(A, W, g, be, 1e-5, bias=b, shortcut=x32, x32=x32, xb=xb)
torch.cuda.synchronize()
