import sys, torch, time
sys.path.insert(0, "/root/repo")
from qiddm_b200 import _lib as L
for m in (10, 20, 40, 80, 118):
    for kind in ("image", "random"):
        torch.manual_seed(m)
        if kind == "image":
            base = torch.rand(m, 12, dtype=torch.float64) @ torch.rand(12, 784, dtype=torch.float64)
            x = base / base.max() + 0.05 * torch.rand(m, 784, dtype=torch.float64)
        else:
            x = torch.randn(m, 784, dtype=torch.float64)
        xc = (x - x.mean(0)).cuda()
        g = xc @ xc.T
        for _ in range(3):
            L.sym_eigh(g)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            lam, v = L.sym_eigh(g)
        b.record(); torch.cuda.synchronize()
        ref = torch.linalg.eigvalsh(g.cpu()).flip(0)
        t0 = time.perf_counter(); torch.linalg.eigh(g); torch.cuda.synchronize(); t1 = time.perf_counter()
        print(m, kind, "jacobi ms", round(a.elapsed_time(b) / 10, 4), "torch eigh ms", round((t1 - t0) * 1e3, 3), "err", float((lam.cpu() - ref).abs().max() / ref.abs().max()))
