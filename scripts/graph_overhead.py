"""Per-node overhead of CUDA-graph replay: N tiny dependent kernels in one graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlr_b200 import ops
e = ops.ext()
cur = torch.zeros(1, dtype=torch.int32, device="cuda:0")
for n in (50, 200, 800):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): e.advance_cursor(cur, 1)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): e.advance_cursor(cur, 1)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"graph of {n} 1-thread kernels: {e0.elapsed_time(e1) / 20 * 1e3 / n:.2f} us per node")
