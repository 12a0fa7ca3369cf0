"""A few fused Diffusion(QDenseUndirected_old_noise(60,28)) training steps at the bench size (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qiddm_b200 import models, nn, noise

n, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 52428, int(sys.argv[2]) if len(sys.argv) > 2 else 4
net = nn.QDenseUndirected_old_noise(60, 28)
diff = models.Diffusion(net, noise.add_normal_noise_multiple, "data", (28, 28), torch.nn.MSELoss()).to("cuda")
diff.train()
x = torch.rand(n, 784, device="cuda")
for _ in range(steps):
    net.weights.grad = None
    with torch.no_grad():
        net.weights.add_(0.0)
    (loss,) = diff(x=x, T=10)
torch.cuda.synchronize()
print("loss", loss.item())
