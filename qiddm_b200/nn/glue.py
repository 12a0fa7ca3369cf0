"""Device-side UNet glue around QConv2d: drop-in `BatchNorm2d` and `Upsample` whose CUDA path runs
the library's own HBM-bound kernels (qiddm_batchnorm_*, qiddm_upsample_bilinear_*) in the tensors' dtype (float64
like the reference `nn/unet.py:28-116`, or float32).  Same constructor arguments, parameters, buffers and
`state_dict` keys as the torch modules they subclass; CPU tensors (the classical `qdepth=0` UNet in host tests) go
through the torch implementation of the parent class."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib as L

_DT = {torch.float32: L.DTYPE_F32, torch.float64: L.DTYPE_F64}


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


RELU_NONE, RELU_POST, RELU_PRE = 0, 1, 2       # relu_mode of qiddm_batchnorm_relu_*


class _BatchNormFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, relu_mode=RELU_NONE):
        lib = L.load_library()
        x = x.contiguous()
        n, c = x.shape[0], x.shape[1]
        hw = x.numel() // (n * c)
        y = torch.empty_like(x)
        mean = torch.empty(c, dtype=torch.float64, device=x.device)
        rstd = torch.empty(c, dtype=torch.float64, device=x.device)
        ws = torch.empty(int(lib.qiddm_batchnorm_workspace_bytes(c)), dtype=torch.uint8, device=x.device)
        g = gamma.detach().to(x.dtype).contiguous() if gamma is not None else None
        b = beta.detach().to(x.dtype).contiguous() if beta is not None else None
        rm = rv = None
        if running_mean is not None:
            if running_mean.dtype != x.dtype or not running_mean.is_contiguous():
                raise L.QiddmError("BatchNorm2d: running statistics must be contiguous and have the input's dtype")
            rm, rv = running_mean, running_var
        with torch.cuda.device(x.device):
            L.check(lib.qiddm_batchnorm_relu_forward(L._ptr(x), L._ptr(y), _DT[x.dtype], n, c, hw, L._ptr(g), L._ptr(b),
                                                     L._ptr(mean), L._ptr(rstd), L._ptr(rm), L._ptr(rv), float(momentum),
                                                     float(eps), int(relu_mode), L._ptr(ws), _stream(x.device)),
                    "qiddm_batchnorm_relu_forward")
        ctx.relu_mode = int(relu_mode)
        ctx.save_for_backward(x, g, mean, rstd, b if relu_mode == RELU_POST else None)
        ctx.has_affine = gamma is not None
        ctx.param_dtype = gamma.dtype if gamma is not None else None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load_library()
        x, g, mean, rstd, b = ctx.saved_tensors
        dy = dy.contiguous()
        n, c = x.shape[0], x.shape[1]
        hw = x.numel() // (n * c)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dgamma = torch.empty(c, dtype=x.dtype, device=x.device) if ctx.has_affine else None
        dbeta = torch.empty(c, dtype=x.dtype, device=x.device) if ctx.has_affine else None
        ws = torch.empty(int(lib.qiddm_batchnorm_workspace_bytes(c)), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            L.check(lib.qiddm_batchnorm_relu_backward(L._ptr(x), L._ptr(dy), L._ptr(dx), _DT[x.dtype], n, c, hw, L._ptr(g),
                                                      L._ptr(b), L._ptr(mean), L._ptr(rstd), L._ptr(dgamma), L._ptr(dbeta),
                                                      ctx.relu_mode, L._ptr(ws), _stream(x.device)),
                    "qiddm_batchnorm_relu_backward")
        if ctx.has_affine:
            dgamma, dbeta = dgamma.to(ctx.param_dtype), dbeta.to(ctx.param_dtype)
        return dx, dgamma, dbeta, None, None, None, None, None


class BatchNorm2d(torch.nn.BatchNorm2d):
    """torch.nn.BatchNorm2d (nn/unet.py:43-45, :95, :106) with the batch-statistics path on the library's kernels.
    `fuse_relu` ("post" | "pre" | None, set by `fuse_bn_relu`): the ReLU that follows / precedes this layer in the UNet
    block is computed inside the BatchNorm kernels (the neighbouring `FusedReLU` module then passes its input through)."""

    fuse_relu = None

    def _torch_path(self, x):
        if self.fuse_relu == "pre":
            x = torch.relu(x)
        y = super().forward(x)
        return torch.relu(y) if self.fuse_relu == "post" else y

    def forward(self, x):
        if not x.is_cuda or x.dtype not in _DT or x.dim() != 4 or x.numel() == 0:
            return self._torch_path(x)
        use_batch_stats = self.training or self.running_mean is None
        if not use_batch_stats:           # eval: running statistics, plain elementwise math (differentiable)
            if self.fuse_relu == "pre":
                x = torch.relu(x)
            scale = torch.rsqrt(self.running_var.to(x.dtype) + self.eps)
            shift = -self.running_mean.to(x.dtype) * scale
            if self.affine:
                scale = scale * self.weight.to(x.dtype)
                shift = shift * self.weight.to(x.dtype) + self.bias.to(x.dtype)
            y = x * scale[None, :, None, None] + shift[None, :, None, None]
            return torch.relu(y) if self.fuse_relu == "post" else y
        factor = 0.0
        rm = rv = None
        if self.training and self.track_running_stats and self.running_mean is not None:
            # momentum None = cumulative moving average; reading the counter would sync, so only the (default)
            # exponential average takes the library path
            if self.momentum is None or self.running_mean.dtype != x.dtype:
                return self._torch_path(x)
            self.num_batches_tracked.add_(1)
            factor = self.momentum
            rm, rv = self.running_mean, self.running_var
        mode = {"post": RELU_POST, "pre": RELU_PRE}.get(self.fuse_relu, RELU_NONE)
        return _BatchNormFunction.apply(x, self.weight if self.affine else None, self.bias if self.affine else None,
                                        rm, rv, factor, self.eps, mode)


class FusedReLU(torch.nn.ReLU):
    """torch.nn.ReLU of the UNet blocks (nn/unet.py:44, :47, :96, :107).  `fused = True` (set by `fuse_bn_relu`): the
    neighbouring BatchNorm2d computes it, this module passes its input through (no parameters: state_dict unchanged)."""

    fused = False

    def forward(self, x):
        return x if self.fused else super().forward(x)


def fuse_bn_relu(seq: torch.nn.Sequential) -> torch.nn.Sequential:
    """Pair every FusedReLU of `seq` with the BatchNorm2d right behind it (Conv -> ReLU -> BN: "pre") or right in front of
    it (Conv -> BN -> ReLU: "post")."""
    mods = list(seq.children())
    for i, m in enumerate(mods):
        if not isinstance(m, FusedReLU) or m.fused:
            continue
        prev = mods[i - 1] if i > 0 else None
        nxt = mods[i + 1] if i + 1 < len(mods) else None
        if isinstance(prev, BatchNorm2d) and prev.fuse_relu is None:
            prev.fuse_relu, m.fused = "post", True
        elif isinstance(nxt, BatchNorm2d) and nxt.fuse_relu is None:
            nxt.fuse_relu, m.fused = "pre", True
    return seq


class _MaxPoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k):
        lib = L.load_library()
        x = x.contiguous()
        n, c, h, w = x.shape
        y = torch.empty((n, c, h // k, w // k), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            L.check(lib.qiddm_maxpool2d_forward(L._ptr(x), L._ptr(y), _DT[x.dtype], n * c, h, w, k, _stream(x.device)),
                    "qiddm_maxpool2d_forward")
        ctx.save_for_backward(x)
        ctx.k = k
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = L.load_library()
        (x,) = ctx.saved_tensors
        gy = gy.contiguous()
        n, c, h, w = x.shape
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            L.check(lib.qiddm_maxpool2d_backward(L._ptr(x), L._ptr(gy), L._ptr(gx), _DT[x.dtype], n * c, h, w, ctx.k,
                                                 _stream(x.device)), "qiddm_maxpool2d_backward")
        return gx, None


class MaxPool2d(torch.nn.MaxPool2d):
    """torch.nn.MaxPool2d(kernel_size=2, stride=2) of nn/unet.py:110 on the library's kernel (indices recomputed in the
    backward); any other configuration, and CPU tensors, go through torch."""

    def forward(self, x):
        k = self.kernel_size if isinstance(self.kernel_size, int) else (self.kernel_size[0] if self.kernel_size[0] == self.kernel_size[1] else None)
        st = self.stride if isinstance(self.stride, int) else (self.stride[0] if self.stride[0] == self.stride[1] else None)
        ok = (x.is_cuda and x.dtype in _DT and x.dim() == 4 and x.numel() > 0 and k is not None and st == k
              and self.padding in (0, (0, 0)) and self.dilation in (1, (1, 1)) and not self.ceil_mode and not self.return_indices
              and x.shape[2] >= k and x.shape[3] >= k)
        if not ok:
            return super().forward(x)
        return _MaxPoolFunction.apply(x, int(k))


class _UpsampleFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h_out, w_out, scale_h, scale_w):
        lib = L.load_library()
        x = x.contiguous()
        n, c, h, w = x.shape
        out = torch.empty((n, c, h_out, w_out), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            L.check(lib.qiddm_upsample_bilinear_forward(L._ptr(x), L._ptr(out), _DT[x.dtype], n * c, h, w, h_out, w_out,
                                                        scale_h, scale_w, _stream(x.device)),
                    "qiddm_upsample_bilinear_forward")
        ctx.geom = (n, c, h, w, h_out, w_out, scale_h, scale_w)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.load_library()
        n, c, h, w, h_out, w_out, scale_h, scale_w = ctx.geom
        g = g.contiguous()
        gin = torch.empty((n, c, h, w), dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            L.check(lib.qiddm_upsample_bilinear_backward(L._ptr(g), L._ptr(gin), _DT[g.dtype], n * c, h, w, h_out, w_out,
                                                         scale_h, scale_w, _stream(g.device)),
                    "qiddm_upsample_bilinear_backward")
        return gin, None, None, None, None


class Upsample(torch.nn.Upsample):
    """torch.nn.Upsample(scale_factor=2, mode="bilinear") of nn/unet.py:38 on the library's kernel
    (align_corners = False, the torch default the reference relies on)."""

    def forward(self, x):
        ok = (x.is_cuda and x.dtype in _DT and x.dim() == 4 and self.mode == "bilinear" and not self.align_corners
              and x.numel() > 0 and not getattr(self, "recompute_scale_factor", None))
        if not ok:
            return super().forward(x)
        h, w = x.shape[2], x.shape[3]
        if self.size is not None:
            h_out, w_out = (self.size, self.size) if isinstance(self.size, int) else tuple(self.size)
            sh, sw = h / h_out, w / w_out
        else:
            sf = self.scale_factor
            fh, fw = (sf, sf) if not isinstance(sf, (tuple, list)) else sf
            h_out, w_out = int(h * fh), int(w * fw)       # floor, as torch computes the output size
            sh, sw = 1.0 / fh, 1.0 / fw                   # the given scale factor drives the source coordinates
        return _UpsampleFunction.apply(x, h_out, w_out, float(sh), float(sw))


# ------------------------------------------------------------------------------------------------------------------
# Linear layers with one narrow side (linear_down: pixels -> qubits, linear_up: qubits -> pixels; nn/qdense.py:219-386, :565-670)
# ------------------------------------------------------------------------------------------------------------------
SKINNY_MAX = 16          # narrow side the streaming kernels are instantiated for
SKINNY_MIN_WIDE = 128    # below this the layer is tiny either way: torch
SKINNY_MIN_ROWS = 4096   # streaming kernels pay above a few thousand rows (2560 rows x 4096: the library GEMM was faster)


class _SkinnyLinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        lib = L.load_library()
        x = x.contiguous()
        w = weight.detach().to(x.dtype).contiguous()
        b = bias.detach().to(x.dtype).contiguous() if bias is not None else None
        rows, in_f = x.shape
        out_f = w.shape[0]
        y = torch.empty((rows, out_f), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            L.check(lib.qiddm_skinny_linear_forward(L._ptr(x), L._ptr(w), L._ptr(b), L._ptr(y), _DT[x.dtype], rows, in_f, out_f,
                                                    _stream(x.device)), "qiddm_skinny_linear_forward")
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        ctx.w_dtype = weight.dtype
        ctx.b_dtype = bias.dtype if bias is not None else None
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = L.load_library()
        x, w = ctx.saved_tensors
        gy = gy.to(x.dtype).contiguous()
        rows, in_f = x.shape
        out_f = w.shape[0]
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        gx = torch.empty_like(x) if need_x else None
        gw = torch.empty_like(w) if (need_w or need_b) else None
        gb = torch.empty(out_f, dtype=x.dtype, device=x.device) if need_b else None
        if rows == 0:
            return (gx, torch.zeros_like(w).to(ctx.w_dtype) if need_w else None,
                    torch.zeros(out_f, dtype=ctx.b_dtype, device=x.device) if need_b else None)
        with torch.cuda.device(x.device):
            ws = torch.empty(int(lib.qiddm_skinny_linear_workspace_bytes(rows, in_f, out_f)), dtype=torch.uint8, device=x.device)
            L.check(lib.qiddm_skinny_linear_backward(L._ptr(x), L._ptr(w), L._ptr(gy), L._ptr(gx), L._ptr(gw), L._ptr(gb),
                                                     _DT[x.dtype], rows, in_f, out_f, L._ptr(ws), _stream(x.device)),
                    "qiddm_skinny_linear_backward")
        return (gx, gw.to(ctx.w_dtype) if need_w else None, gb.to(ctx.b_dtype) if need_b else None)


def skinny_linear(x: torch.Tensor, layer: torch.nn.Linear) -> torch.Tensor:
    """`layer(x)` for a (rows, in) CUDA float32 / float64 input through the streaming kernels when one side of the layer is
    narrow (<= 16) and the other wide; anything else (CPU tensors, other shapes / dtypes) goes through torch."""
    in_f, out_f = layer.in_features, layer.out_features
    if (x.is_cuda and x.dim() == 2 and x.dtype in _DT and min(in_f, out_f) <= SKINNY_MAX and max(in_f, out_f) >= SKINNY_MIN_WIDE
            and layer.weight.is_cuda and x.shape[0] >= SKINNY_MIN_ROWS):
        return _SkinnyLinearFunction.apply(x, layer.weight, layer.bias)
    return layer(x)
