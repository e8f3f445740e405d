import os, sys, time
sys.path.insert(0, "/root/repo")
import torch, bench
from stonkgs_b200 import synthetic
from stonkgs_b200.optim import FusedAdamW
dev = torch.device("cuda", 0)
model = bench.build_model(dev, 12).train()
opt = FusedAdamW(model, lr=1e-4, weight_decay=0.0, max_grad_norm=1.0)
bs = [{k: v.to(dev) for k, v in synthetic.make_batch(64, bench.N_KG, seed=200 + i).items()} for i in range(4)]
def step(i):
    opt.zero_grad(); model(**bs[i % 4])[0].backward(); opt.step()
for i in range(5): step(i)
torch.cuda.synchronize()
host = []
t_all0 = time.perf_counter()
for i in range(20):
    t0 = time.perf_counter(); step(i); host.append(time.perf_counter() - t0)
t_enq = time.perf_counter() - t_all0
torch.cuda.synchronize()
t_all = time.perf_counter() - t_all0
print("host enqueue per step (ms): first 5", [round(h * 1e3, 1) for h in host[:5]], "mean", round(sum(host) / len(host) * 1e3, 2))
print("total enqueue %.1f ms, total with GPU %.1f ms for 20 steps" % (t_enq * 1e3, t_all * 1e3))
# pure host cost: same loop with GPU idle in between (sync after each step, time only the enqueue part)
host2 = []
for i in range(10):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(i); host2.append(time.perf_counter() - t0)
print("host enqueue per step with an empty queue (ms): mean", round(sum(host2) / len(host2) * 1e3, 2))
