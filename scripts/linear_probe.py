"""linear_down / linear_up shapes of config 1 through the skinny-linear kernels (timing + ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qiddm_b200.nn.glue import skinny_linear

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 40960
dt = torch.float64
down = torch.nn.Linear(784, 6).to("cuda", dt)
up = torch.nn.Linear(6, 784).to("cuda", dt)
x = torch.randn(rows, 784, device="cuda", dtype=dt)
g = torch.randn(rows, 784, device="cuda", dtype=dt)
for it in range(3):
    h = skinny_linear(x, down)
    y = skinny_linear(h, up)
    y.backward(g)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record(); h = skinny_linear(x, down); ev[1].record(); y = skinny_linear(h, up); ev[2].record(); y.backward(g); ev[3].record()
torch.cuda.synchronize()
print("down fwd %.1f us, up fwd %.1f us, backward (both) %.1f us" % tuple(1e3 * ev[i].elapsed_time(ev[i + 1]) for i in range(3)))
