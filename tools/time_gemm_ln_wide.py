import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, time
from mvuld_b200 import _lib
dev='cuda'
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1000
for (M,N,K) in [(50176,512,512),(50176,512,2048),(12544,1024,1024),(12544,1024,4096),(12544,1024,2048),(16896,768,768),(16896,768,3072)]:
    A=torch.randn(M,K,device=dev).bfloat16(); W=(torch.randn(N,K,device=dev)*0.05).bfloat16()
    b=torch.randn(N,device=dev); g=torch.ones(N,device=dev); be=torch.zeros(N,device=dev)
    x32=torch.randn(M,N,device=dev); xb=torch.empty(M,N,device=dev,dtype=torch.bfloat16); y=torch.empty(M,N,device=dev,dtype=torch.bfloat16)
    t_f=bench(lambda: _lib.gemm_ln_wide(A,W,g,be,1e-5,bias=b,shortcut=x32,x32=x32,xb=xb))
    def two():
        _lib.gemm(A,W,bias=b,out_bf16=y)
        _lib.call("mvuld_ln_rows", y, x32, g, be, x32, xb, M, N, 1e-5, 1)
    t_2=bench(two)
    t_g=bench(lambda: _lib.gemm(A,W,bias=b,out_bf16=y))
    extra=""
    if N==512:
        t_1=bench(lambda: _lib.gemm_ln(A,W,g,be,1e-5,bias=b,shortcut=x32,x32=x32,xb=xb)); extra=f" single-CTA gemm_ln {t_1:.1f}"
    print(f"M={M} N={N} K={K}: cluster {t_f:.1f} us | gemm+ln {t_2:.1f} us (gemm alone {t_g:.1f}){extra} | {2*M*N*K/t_f/1e6:.0f} TFLOP/s")
