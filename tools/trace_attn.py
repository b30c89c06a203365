"""Debug: timeline of one attention CTA (build with MVULD_NVCC_EXTRA='-DMV_ATT_TRACE')."""
import sys, os, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvuld_b200 import _lib
B, H, W, C, nH, ws, shift = 2, 28, 28, 512, 16, 28, 0
g = torch.Generator().manual_seed(0)
q = torch.nn.functional.normalize(torch.randn(B * nH, ws * ws, 32, generator=g), dim=-1).mul(14.0).to("cuda", torch.float16)
k = torch.nn.functional.normalize(torch.randn(B * nH, ws * ws, 32, generator=g), dim=-1).to("cuda", torch.float16)
v = torch.randn(B * nH, ws * ws, 32, generator=g).to("cuda", torch.bfloat16)
side = 2 * ws - 1
tab = (torch.rand(nH, side * side, generator=g) * 16 * 1.4427).cuda()
tmax = tab.max(1).values.contiguous()
qn = torch.full((nH,), 14.0, device="cuda") if os.environ.get("FIXED_REF", "1") == "1" else None
out = torch.empty(B * H * W, C, device="cuda", dtype=torch.bfloat16)
lib = _lib.load()
buf = (ctypes.c_longlong * (4 * 16384))()
for it in range(2):
    _lib.call(os.environ.get("ENTRY", "mvuld_swin_window_attention"), q, k, v, tab, tmax, qn, out, B, H, W, C, nH, ws, shift)
    n = lib.mvuld_debug_att_trace(buf, 16384)
recs = []
for i in range(n):
    w, clk = buf[4 * i], buf[4 * i + 1]
    recs.append((clk, (w >> 48) & 0xffff, (w >> 32) & 0xffff, (w >> 16) & 0xffff, w & 0xffff))
recs.sort()
t0 = recs[0][0]
names = {0: "sm:wait_s", 1: "sm:got_s", 2: "sm:ldtm_done", 3: "sm:max_done", 4: "sm:sweep_done", 5: "sm:got_pv", 6: "sm:sweep_begin",
         7: "sm:p_full", 10: "mma:S_begin", 11: "mma:S_end", 12: "mma:PV_begin", 13: "mma:PV_end"}
with open("gpurun_out/att_trace.txt", "w") as fh:
    for clk, tag, g_, t, j in recs:
        fh.write(f"{clk - t0:8d} g{g_} t{t} j{j} {names.get(tag, tag)}\n")
print("records", n, "span", recs[-1][0] - t0)
