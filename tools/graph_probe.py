"""Eager launches vs CUDA-graph replay of one extraction step (batch 256): how much of the step is launch gaps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from stonkgs_b200 import synthetic

dev = torch.device("cuda", 0)
model = bench.build_model(dev, 12)
b = {k: v.to(dev) for k, v in synthetic.make_batch(256, bench.N_KG, seed=3, with_labels=False).items()}
def step():
    return model.embed(b["input_ids"], b["attention_mask"], b["token_type_ids"])
for _ in range(3): out = step()
torch.cuda.synchronize()
def timeit(fn, n=10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager  ms/step", timeit(step), timeit(step))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): step()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    ref = step()
torch.cuda.synchronize()
print("graph  ms/step", timeit(g.replay), timeit(g.replay))
print("same result", torch.equal(ref, out))
