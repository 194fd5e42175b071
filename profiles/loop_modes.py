"""Debug: the bench loop's period in three forms - EnvStream.advance eager, the same captured in a CUDA graph, and bench.py's own number."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _variant  # noqa: F401
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
s = bench.EnvStream("pushing", 4096, 0, 0)
n = s.n; dev = s.dev
ids = torch.arange(n, device=dev)
for k in range(400):
    s.advance((ids % s.ep_len == k).to(torch.uint8))
torch.cuda.synchronize()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
def timed(fn, label):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for k in range(K): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1) / K:.3f} ms per step")
timed(s.advance, "eager advance")
cap = torch.cuda.Stream(device=dev)
cap.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(cap):
    for _ in range(2): s.advance()
torch.cuda.current_stream(dev).wait_stream(cap); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=cap):
    s.advance()
torch.cuda.synchronize()
timed(g.replay, "graph replay")
timed(s.advance, "eager advance again")
timed(g.replay, "graph replay again")
