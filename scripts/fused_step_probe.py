"""Per-kernel timers of the Diffusion(QDenseUndirected_old_noise(60,28)) training step, fused and unfused (device-resident images)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qiddm_b200 import _lib as L
from qiddm_b200 import models, nn, noise

n = int(sys.argv[1]) if len(sys.argv) > 1 else 52428
net = nn.QDenseUndirected_old_noise(60, 28)
diff = models.Diffusion(net, noise.add_normal_noise_multiple, "data", (28, 28), torch.nn.MSELoss()).to("cuda")
diff.train()
x = torch.rand(n, 784, device="cuda")
for fused in (1, 0, 1):
    os.environ["QIDDM_FUSED_STEP"] = str(fused)
    for _ in range(3):
        net.weights.grad = None
        diff(x=x, T=10)
    torch.cuda.synchronize()
    L.timing_enable(True); L.timing_collect()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        net.weights.grad = None
        with torch.no_grad():
            net.weights.add_(0.0)        # new weights version: the collapse is part of every step
        (loss,) = diff(x=x, T=10)
    b.record(); torch.cuda.synchronize()
    k = L.timing_collect(); L.timing_enable(False)
    print("fused" if fused else "unfused", "%.3f ms/step, loss %.6f" % (a.elapsed_time(b) / 5, loss.item()),
          {kk: round(v["ms"] / 5, 3) for kk, v in k.items() if v["launches"]}, flush=True)
