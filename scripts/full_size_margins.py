import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from conftest import rel_to_max
from oracle import qiddm_oracle as O
from qiddm_b200 import _lib as L
from qiddm_b200.functional import run_stage
from test_gpu_gemm_path import _spec
B = 262144
d = O.desc_qdense(60, 784, O.REMAP_TANH)
full = O.StageDesc(**{**d.__dict__, "read_count": 1024, "post_scale": 1.0, "clamp": False})
for seed in (99, 100, 101):
    g = torch.Generator().manual_seed(seed)
    W = (torch.randn(1, 60, 10, 3, generator=g, dtype=torch.float64) * 0.4)
    x = torch.rand(B, 784, generator=g, dtype=torch.float32)
    xd, Wd = x.cuda(), W.cuda()
    p = run_stage(_spec(full, L.PATH_GEMM), xd, Wd)
    s = p.sum(dim=1)
    idx = torch.tensor([0, 1, 777, 65535, 131072, 200001, B - 2])
    ref = O.run_stage(full, x[idx].double(), W)
    m1 = (s - 1).abs().max().item(); m2 = rel_to_max(p[idx.cuda()], ref); mn = p.min().item()
    del p
    out = run_stage(_spec(d, L.PATH_GEMM), xd, Wd)
    sl = slice(100000, 100512)
    m3 = rel_to_max(out[sl], run_stage(_spec(d, L.PATH_GATE), xd[sl], Wd))
    xs = xd[:65536].clone()
    g1 = torch.randn(65536, 784, generator=g).cuda() / 65536
    g2 = torch.randn(65536, 784, generator=g).cuda() / 65536
    def wgrad(go):
        Wp = Wd.clone().requires_grad_(True)
        o = run_stage(_spec(d, L.PATH_GEMM), xs, Wp)
        (o * go).sum().backward()
        return Wp.grad
    a, b, c = wgrad(g1), wgrad(g2), wgrad(g1 + g2)
    m4 = rel_to_max(a + b, c)
    c2 = wgrad(g1 + g2)
    print(f"seed {seed}: sum-1 {m1:.2e} (3e-5)  rows-vs-oracle {m2:.2e} (1e-5)  min p {mn:.2e}  gemm-vs-gate {m3:.2e} (2e-5)  linearity {m4:.2e} (2e-4)  repeat {rel_to_max(c2, c):.2e}")
