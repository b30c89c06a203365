"""Times the GEMM shapes of the full step (SwinV2-B stage 2 / 3, RoBERTa-base) through the C-ABI, 20 launches each."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvuld_b200 import _lib
SHAPES = [("swin s2 fc1+gelu", 50176, 2048, 512, 1), ("swin s2 proj", 50176, 512, 512, 0), ("swin s1 fc1+gelu", 200704, 1024, 256, 1),
          ("swin s0 fc1+gelu", 802816, 512, 128, 1), ("swin s3 fc1+gelu", 12544, 4096, 1024, 1),
          ("roberta fc1+gelu", 16896, 3072, 768, 1), ("roberta fc2", 16896, 768, 3072, 0), ("roberta dense", 16896, 768, 768, 0)]
g = torch.Generator().manual_seed(0)
for name, M, N, K, act in SHAPES:
    A = (torch.randn(M, K, generator=g) * 0.5).to("cuda", torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * 0.05).to("cuda", torch.bfloat16)
    bias = torch.randn(N, generator=g).cuda()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        _lib.gemm(A, W, bias=bias, act=act, out_bf16=out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        _lib.gemm(A, W, bias=bias, act=act, out_bf16=out)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    print(f"{name:18s} M={M} N={N} K={K}: {ms * 1000:7.1f} us  {2.0 * M * N * K / ms / 1e9:6.0f} TFLOP/s", flush=True)
