"""Noise ladder used by the diffusion training step (reference `src/noise.py:105-126`; the other
schedules there are unused by every driver)."""
import torch


def add_normal_noise_multiple(data: torch.Tensor, tau: int, decay_mod: float = 1.0, eps: torch.Tensor = None):
    """x_t = clamp(x (1 - w_t) + eps w_t, 0, 1), w_t = (t/(tau-1))**decay_mod, eps ~ N(0.5, 0.2) drawn once
    per sample (float32 draw as in the reference).  Returns ((batch tau), pixels), batch-major.
    `eps` may be passed in for reproducible tests; everything stays on `data.device`."""
    if data.dim() == 1:
        data = data.unsqueeze(0)
    batch, pixels = data.shape
    if eps is None:
        eps = torch.normal(mean=0.5, std=0.2, size=(batch, pixels), device=data.device)
    w = torch.linspace(0, 1, tau, device=data.device) ** decay_mod
    w = (w / w.max()).to(data.dtype)[None, :, None]                      # (1, tau, 1)
    noisy = data[:, None, :] * (1 - w) + eps.to(data.device)[:, None, :] * w
    return noisy.clamp(0, 1).reshape(batch * tau, pixels)


def _level_weights(tau: int, decay_mod: float, device, dtype) -> torch.Tensor:
    w = torch.linspace(0, 1, tau, device=device) ** decay_mod
    return (w / w.max()).to(dtype)


def ladder_pair(data: torch.Tensor, T: int, decay_mod: float = 3.0, eps: torch.Tensor = None, want_clean: bool = True,
                return_draw: bool = False):
    """(noisy, clean) of the training step in one kernel launch (qiddm_noise_ladder): the tau = T + 1 level ladder of
    `add_normal_noise_multiple` (src/noise.py:105-126) written directly as noisy = levels 1..T and clean = levels 0..T-1,
    each ((batch T), pixels) batch-major -- what src/models.py:46-63 slices out of the full ladder.  CUDA tensors only.
    `want_clean=False`: only `noisy` is materialised (clean = None); `return_draw`: also (images, eps, level weights), from
    which mse_ladder_loss_and_grad recomputes the target."""
    import ctypes as C
    from . import _lib as L
    if not data.is_cuda:
        raise L.QiddmError("ladder_pair needs CUDA tensors; use add_normal_noise_multiple on the host")
    if data.dim() == 1:
        data = data.unsqueeze(0)
    data = data.contiguous()
    batch, pixels = data.shape
    if eps is None:
        eps = torch.normal(mean=0.5, std=0.2, size=(batch, pixels), device=data.device)      # float32 draw, as the reference
    eps = eps.to(device=data.device, dtype=torch.float32).contiguous()
    dt = {torch.float32: L.DTYPE_F32, torch.float64: L.DTYPE_F64}[data.dtype]
    w = _level_weights(T + 1, decay_mod, data.device, data.dtype)
    noisy = torch.empty((batch * T, pixels), dtype=data.dtype, device=data.device)
    clean = torch.empty_like(noisy) if want_clean else None
    with torch.cuda.device(data.device):
        L.check(L.load_library().qiddm_noise_ladder(L._ptr(data), L._ptr(eps), L._ptr(w), dt, batch, pixels, T + 1,
                                                    L._ptr(noisy), L._ptr(clean),
                                                    C.c_void_p(torch.cuda.current_stream(data.device).cuda_stream)),
                "qiddm_noise_ladder")
    if return_draw:
        return noisy, clean, (data, eps, w)
    return noisy, clean


def mse_ladder_loss_and_grad(pred: torch.Tensor, draw, T: int, scale: float = 1.0, shift: float = 0.0, c0: float = 1.0,
                             c1: float = 0.0):
    """loss = mean((scale * pred + shift - (c0 level_t + c1 level_{t+1}))^2) and d loss / d pred with the ladder levels
    recomputed from `draw` = (images, eps, level weights) of ladder_pair(..., return_draw=True) (qiddm_mse_ladder_loss_grad):
    goal "data" c0 = 1, c1 = 0 (src/models.py:65-67); goal "noise" scale 0.1, shift -0.05, c0 = -1, c1 = 1 (:95-99)."""
    import ctypes as C
    from . import _lib as L
    lib = L.load_library()
    data, eps, w = draw
    pred_c = pred.detach().contiguous()
    batch, pixels = data.shape
    if pred_c.numel() != batch * T * pixels or pred_c.dtype != data.dtype:
        raise L.QiddmError(f"mse_ladder_loss_and_grad: pred {tuple(pred_c.shape)} {pred_c.dtype} vs {batch} x {T} x {pixels} {data.dtype}")
    dt = {torch.float32: L.DTYPE_F32, torch.float64: L.DTYPE_F64}[pred_c.dtype]
    grad = torch.empty_like(pred_c)
    loss = torch.empty((), dtype=pred_c.dtype, device=pred_c.device)
    ws = torch.empty(int(lib.qiddm_mse_workspace_bytes()), dtype=torch.uint8, device=pred_c.device)
    with torch.cuda.device(pred_c.device):
        L.check(lib.qiddm_mse_ladder_loss_grad(L._ptr(pred_c), L._ptr(data), L._ptr(eps), L._ptr(w), dt, batch, pixels, T + 1,
                                               float(scale), float(shift), float(c0), float(c1), L._ptr(grad), L._ptr(loss),
                                               L._ptr(ws), C.c_void_p(torch.cuda.current_stream(pred_c.device).cuda_stream)),
                "qiddm_mse_ladder_loss_grad")
    return loss, grad


def mse_loss_and_grad(pred: torch.Tensor, target: torch.Tensor, target_add: torch.Tensor = None, scale: float = 1.0,
                      shift: float = 0.0):
    """loss = mean((scale * pred + shift - target + target_add)^2) and d loss / d pred in one pass (qiddm_mse_loss_grad):
    the MSELoss + `.mean().backward()` seed of src/models.py:65-67 (goal "data") and :95-99 (goal "noise": scale 0.1,
    shift -0.05, target = noisy, target_add = clean).  Returns (loss 0-d tensor, grad like pred)."""
    import ctypes as C
    from . import _lib as L
    lib = L.load_library()
    pred_c = pred.detach().contiguous()
    dt = {torch.float32: L.DTYPE_F32, torch.float64: L.DTYPE_F64}[pred_c.dtype]
    t1 = target.detach().to(pred_c.dtype).contiguous()
    t2 = target_add.detach().to(pred_c.dtype).contiguous() if target_add is not None else None
    grad = torch.empty_like(pred_c)
    loss = torch.empty((), dtype=pred_c.dtype, device=pred_c.device)
    ws = torch.empty(int(lib.qiddm_mse_workspace_bytes()), dtype=torch.uint8, device=pred_c.device)
    with torch.cuda.device(pred_c.device):
        L.check(lib.qiddm_mse_loss_grad(L._ptr(pred_c), L._ptr(t1), L._ptr(t2), dt, float(scale), float(shift), pred_c.numel(),
                                        L._ptr(grad), L._ptr(loss), L._ptr(ws),
                                        C.c_void_p(torch.cuda.current_stream(pred_c.device).cuda_stream)),
                "qiddm_mse_loss_grad")
    return loss, grad


class _LinearUpMSE(torch.autograd.Function):
    """loss = mean((scale * linear_up(h) + shift - target)^2) with the target recomputed from the noise draw; the loss and all
    three gradients come out of the forward call (qiddm_linear_up_mse_step), the backward hands them to autograd."""

    @staticmethod
    def forward(ctx, h, weight, bias, data, eps, w, T, scale, shift, c0, c1):
        import ctypes as C
        from . import _lib as L
        lib = L.load_library()
        dt = {torch.float32: L.DTYPE_F32, torch.float64: L.DTYPE_F64}[h.dtype]
        hc = h.detach().contiguous()
        wc = weight.detach().to(h.dtype).contiguous()
        bc = bias.detach().to(h.dtype).contiguous() if bias is not None else None
        batch, pixels = data.shape
        hidden = wc.shape[1]
        need_h = ctx.needs_input_grad[0]
        loss = torch.empty((), dtype=h.dtype, device=h.device)
        gw = torch.empty_like(wc)
        gb = torch.empty(pixels, dtype=h.dtype, device=h.device) if bias is not None else None
        gh = torch.empty_like(hc)                      # always written (its V term also feeds the loss)
        ws = torch.empty(int(lib.qiddm_linear_up_mse_workspace_bytes(pixels, hidden)), dtype=torch.uint8, device=h.device)
        with torch.cuda.device(h.device):
            L.check(lib.qiddm_linear_up_mse_step(L._ptr(hc), L._ptr(wc), L._ptr(bc), L._ptr(data), L._ptr(eps), L._ptr(w), dt, batch,
                                                 pixels, T + 1, hidden, float(scale), float(shift), float(c0), float(c1), L._ptr(loss),
                                                 L._ptr(gw), L._ptr(gb), L._ptr(gh), L._ptr(ws),
                                                 C.c_void_p(torch.cuda.current_stream(h.device).cuda_stream)),
                    "qiddm_linear_up_mse_step")
        ctx.grads = (gh if need_h else None, gw.to(weight.dtype), gb.to(bias.dtype) if gb is not None else None)
        return loss

    @staticmethod
    def backward(ctx, g):
        gh, gw, gb = ctx.grads
        ctx.grads = None
        return (gh * g if gh is not None else None, gw * g, gb * g if gb is not None else None,
                None, None, None, None, None, None, None, None)


def linear_up_mse_ok(h: torch.Tensor, layer: torch.nn.Linear, draw, T: int) -> bool:
    data = draw[0]
    return (h.is_cuda and h.dim() == 2 and h.dtype in (torch.float32, torch.float64) and h.dtype == data.dtype
            and isinstance(layer, torch.nn.Linear) and layer.in_features <= 16 and layer.in_features == h.shape[1]
            and layer.out_features == data.shape[1] and h.shape[0] == data.shape[0] * T and layer.weight.is_cuda
            and (data.shape[1] + 8 * 33) * layer.in_features * 8 <= 200 * 1024 and T <= 32)


def linear_up_mse_loss(h: torch.Tensor, layer: torch.nn.Linear, draw, T: int, scale: float = 1.0, shift: float = 0.0,
                       c0: float = 1.0, c1: float = 0.0) -> torch.Tensor:
    """Differentiable scalar: the tail `linear_up -> MSELoss(..., target).mean()` of a re-upload network's training step with the
    ladder target recomputed from `draw` (ladder_pair(..., return_draw=True)); see qiddm_linear_up_mse_step."""
    data, eps, w = draw
    return _LinearUpMSE.apply(h, layer.weight, layer.bias, data, eps, w, T, scale, shift, c0, c1)
