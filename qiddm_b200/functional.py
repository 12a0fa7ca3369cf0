"""torch.autograd.Functions over the C-ABI (forward kernel + adjoint-method backward kernel).

These replace `self.qnode(inputs[, weights])` of the reference modules (nn/qdense.py:58,:465,:1633;
nn/qconv.py H1).  Outputs are returned in the dtype of the input (weights' dtype for input-less
circuits); the simulation itself runs in fp32 on the device."""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import Plan, StageSpec, UnfoldDesc


class _StageFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: Plan, x: Optional[torch.Tensor], weights: torch.Tensor, batch: Optional[int],
                basis: Optional[torch.Tensor] = None):
        ctx.plan = plan
        ctx.basis = basis
        ctx.x_dtype = x.dtype if x is not None else None
        ctx.gemm = x is not None and plan.use_gemm(x.shape[0])
        ctx.gemm_saved = None
        ctx.gate_state = None
        if ctx.gemm:
            if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
                out, ctx.gemm_saved = plan.gemm_forward(x.detach(), weights, save=True)
            else:
                out = plan.gemm_forward(x.detach(), weights)
        elif any(ctx.needs_input_grad[1:3]):
            # training: psi_final travels to the backward (the adjoint kernel skips its forward recomputation)
            out, ctx.gate_state = plan.forward(x.detach() if x is not None else None, weights.detach(), batch=batch,
                                               basis=basis, save_state=True)
        else:
            out = plan.forward(x.detach() if x is not None else None, weights.detach(), batch=batch, basis=basis)
        ctx.save_for_backward(x if x is not None else torch.empty(0), weights)
        ctx.has_x = x is not None
        return out.to(x.dtype if x is not None else weights.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        x, weights = ctx.saved_tensors
        x = x if ctx.has_x else None
        need_x = ctx.has_x and ctx.needs_input_grad[1]
        need_w = ctx.needs_input_grad[2]
        if ctx.gemm:
            gi, gw = ctx.plan.gemm_backward(x, weights, grad_out, need_grad_in=need_x, need_grad_w=need_w,
                                            saved=ctx.gemm_saved)
            ctx.gemm_saved = None
        else:
            gi, gw = ctx.plan.backward(x, weights, grad_out, need_grad_in=need_x, need_grad_w=need_w, basis=ctx.basis,
                                       state=ctx.gate_state)
            ctx.gate_state = None
        if gi is not None:
            gi = gi.to(ctx.x_dtype)
        return None, gi, (gw.view_as(weights) if gw is not None else None), None, None


def run_stage(spec: StageSpec, x: Optional[torch.Tensor], weights: torch.Tensor,
              batch: Optional[int] = None, basis: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One QNode-equivalent evaluation: (B, n_in) -> (B, n_out), differentiable in x and weights.
    `basis` (int32, (B,)): start states of an INIT_BASIS descriptor."""
    return _StageFunction.apply(Plan.get(spec), x, weights, batch, basis)


class _QConvFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: Plan, img: torch.Tensor, weights: torch.Tensor, unfold: UnfoldDesc):
        ctx.plan, ctx.unfold = plan, unfold
        ho = img.shape[2] + 2 * unfold.pad_h - unfold.kernel_h + 1
        wo = img.shape[3] + 2 * unfold.pad_w - unfold.kernel_w + 1
        ctx.gemm = plan.use_collapse_qconv(unfold, img.shape[0] * ho * wo)       # circuit instances = patches
        ctx.gemm_saved = None
        if ctx.gemm:
            if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
                out, ctx.gemm_saved = plan.qconv_gemm_forward(img.detach(), weights, unfold, save=True)
            else:
                out = plan.qconv_gemm_forward(img.detach(), weights, unfold)
        else:
            out = plan.qconv_forward(img.detach(), weights.detach(), unfold)
        ctx.save_for_backward(img, weights)
        return out.to(img.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        img, weights = ctx.saved_tensors
        if ctx.gemm:
            gi, gw = ctx.plan.qconv_gemm_backward(img, weights, grad_out, ctx.unfold,
                                                  need_grad_in=ctx.needs_input_grad[1],
                                                  need_grad_w=ctx.needs_input_grad[2], saved=ctx.gemm_saved)
            ctx.gemm_saved = None
        else:
            gi, gw = ctx.plan.qconv_backward(img, weights, grad_out, ctx.unfold,
                                             need_grad_in=ctx.needs_input_grad[1],
                                             need_grad_w=ctx.needs_input_grad[2])
        if gi is not None:
            gi = gi.to(img.dtype)
        return None, gi, (gw.view_as(weights) if gw is not None else None), None


def run_qconv(spec: StageSpec, img: torch.Tensor, weights: torch.Tensor, kernel_size, padding) -> torch.Tensor:
    """Fused unfold + amplitude-embed + SEL + probs readout: (N,C,H,W) -> (N,out,H_out,W_out)."""
    n, c, h, w = img.shape
    unfold = UnfoldDesc(c, h, w, kernel_size[0], kernel_size[1], padding[0], padding[1])
    return _QConvFunction.apply(Plan.get(spec), img, weights, unfold)


class _QConvUpFunction(torch.autograd.Function):
    """Bilinear Upsample -> 1 x 1 QConv2d (nn/unet.py:36-41) with the interpolation inside the convolution's staging."""

    @staticmethod
    def forward(ctx, plan: Plan, src: torch.Tensor, weights: torch.Tensor, unfold: UnfoldDesc, scale_h: float, scale_w: float):
        ctx.plan, ctx.unfold, ctx.scales = plan, unfold, (scale_h, scale_w)
        ctx.saved_y = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            out, ctx.saved_y = plan.qconv_up_forward(src.detach(), weights, unfold, scale_h, scale_w, save=True)
        else:
            out = plan.qconv_up_forward(src.detach(), weights, unfold, scale_h, scale_w)
        ctx.save_for_backward(src, weights)
        return out.to(src.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        from . import _lib as L
        src, weights = ctx.saved_tensors
        sh, sw = ctx.scales
        g_up, gw = ctx.plan.qconv_up_backward(src, weights, grad_out, ctx.unfold, sh, sw, ctx.saved_y,
                                              need_grad_in=ctx.needs_input_grad[1], need_grad_w=ctx.needs_input_grad[2])
        ctx.saved_y = None
        gi = None
        if g_up is not None:      # transpose of the interpolation: gradient of the upsampled image -> gradient of the source
            n, c, h, w = src.shape
            gi = torch.empty((n, c, h, w), dtype=g_up.dtype, device=g_up.device)
            with torch.cuda.device(g_up.device):
                L.check(L.load_library().qiddm_upsample_bilinear_backward(
                    L._ptr(g_up), L._ptr(gi), L.DTYPE_F64 if g_up.dtype == torch.float64 else L.DTYPE_F32, n * c, h, w,
                    ctx.unfold.height, ctx.unfold.width, float(sh), float(sw),
                    L.C.c_void_p(torch.cuda.current_stream(g_up.device).cuda_stream)), "qiddm_upsample_bilinear_backward")
            gi = gi.to(src.dtype)
        return None, gi, (gw.view_as(weights) if gw is not None else None), None, None, None


def run_qconv_up(spec: StageSpec, src: torch.Tensor, weights: torch.Tensor, h_out: int, w_out: int, scale_h: float,
                 scale_w: float):
    """Upsample(bilinear, align_corners=False) to (h_out, w_out) followed by a 1 x 1 QConv, fused; None when the layer has no
    direct-convolution form (the caller then runs the two modules)."""
    if not src.is_cuda or src.dim() != 4 or src.dtype not in (torch.float32, torch.float64) or src.numel() == 0:
        return None
    unfold = UnfoldDesc(src.shape[1], h_out, w_out, 1, 1, 0, 0)
    plan = Plan.get(spec)
    if not (plan.qconv_direct(unfold) and plan.use_collapse_qconv(unfold, src.shape[0] * h_out * w_out)):
        return None
    return _QConvUpFunction.apply(plan, src, weights, unfold, float(scale_h), float(scale_w))


class _QConvReferenceMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img: torch.Tensor, unfold: UnfoldDesc, out_channels: int):
        from ._lib import qconv_reference_map
        ctx.unfold, ctx.out_channels = unfold, out_channels
        ctx.save_for_backward(img)
        return qconv_reference_map(img.detach(), unfold, out_channels)

    @staticmethod
    def backward(ctx, grad_out):
        from ._lib import qconv_reference_map
        (img,) = ctx.saved_tensors
        return qconv_reference_map(img.detach(), ctx.unfold, ctx.out_channels, grad_out=grad_out), None, None


def run_qconv_reference_map(img: torch.Tensor, kernel_size, padding, out_channels: int) -> torch.Tensor:
    """What `_QConv2d_FAST.forward` literally computes in the reference (nn/qconv.py:71-90; the circuit is never called)."""
    n, c, h, w = img.shape
    unfold = UnfoldDesc(c, h, w, kernel_size[0], kernel_size[1], padding[0], padding[1])
    n_ch = min(out_channels, (c * kernel_size[0] * kernel_size[1] + 1) // 2)
    return _QConvReferenceMap.apply(img, unfold, n_ch)


def build_unitary(spec: StageSpec, weights: torch.Tensor) -> torch.Tensor:
    return Plan.get(spec).build_unitary(weights.detach())
