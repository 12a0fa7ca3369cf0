import sys, torch
sys.path.insert(0, "/root/repo")
from qiddm_b200 import _lib as L
from qiddm_b200 import nn as qnn
dev = torch.device("cuda")
net = qnn.QDenseUndirected_old_noise(60, 28).to(dev, torch.float64)
plan = L.Plan.get(net._spec())
B = 524288
x = torch.rand(B, 784, device=dev)
w = net.weights.detach()
a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)      # 1 GiB
b = torch.empty_like(a)
s2 = torch.cuda.Stream()
def gemm():
    return plan.gemm_forward(x, w, save=True)
def copies(n=12):
    for _ in range(n):
        b.copy_(a)
def timeit(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
t_g = timeit(gemm)
t_c = timeit(copies)
def both():
    ev = torch.cuda.Event(); ev.record()
    with torch.cuda.stream(s2):
        s2.wait_event(ev)
        copies()
    gemm()
    torch.cuda.current_stream().wait_stream(s2)
t_b = timeit(both)
print(f"gemm_forward(prep+GEMM) alone {t_g:.3f} ms; 12 x 1 GiB copies alone {t_c:.3f} ms; concurrently {t_b:.3f} ms")
