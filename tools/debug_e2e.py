import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mvuld_b200.prefetch import DevicePrefetcher, ResultSink, to_device
args = bench.parse()
dev = torch.device("cuda", 0)
wl = bench.build_workload(args, 0, dev)
d = wl["to_dev"]()
for _ in range(3): wl["step"](d)
torch.cuda.synchronize()
def timed(name, fn, n=10):
    fn(2); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(n); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n * 1e3
    print(f"{name:40s} {dt:8.2f} ms/step")
def resident(n):
    for _ in range(n): wl["step"](d)
def sync_copy(n):
    for _ in range(n): wl["step"](wl["to_dev"]()).cpu()
def copy_only(n):
    for _ in range(n): wl["to_dev"]()
def copy_same_stream_nosync(n):
    outs = []
    for _ in range(n): outs.append(wl["step"](wl["to_dev"]()))
def prefetch_nosink(n):
    for b in DevicePrefetcher((wl["host"] for _ in range(n)), dev): wl["step"](b)
def prefetch_sink(n):
    s = ResultSink(n)
    for b in DevicePrefetcher((wl["host"] for _ in range(n)), dev): s.push(wl["step"](b))
    s.results()
def host_launch_only(n):
    t0 = time.perf_counter()
    for _ in range(n): wl["step"](d)
    print("   host-side launch time per step (ms):", (time.perf_counter() - t0) / n * 1e3)
for name, fn in [("resident", resident), ("copy only", copy_only), ("sync copy + .cpu()", sync_copy), ("same-stream copy, no sync", copy_same_stream_nosync),
                 ("prefetch, no sink", prefetch_nosink), ("prefetch + sink", prefetch_sink), ("host launch", host_launch_only)]:
    timed(name, fn)
