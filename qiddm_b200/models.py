"""Diffusion wrapper (reference `src/models.py:8-150`): training step (noise ladder -> net -> MSE ->
`.backward()` INSIDE forward, as the reference does) and the fixed-point sampler.  Same constructor,
`forward(x, T=..., verbose=...)`, `sample(...)` and `save_name()`; tensors stay on the device."""
import os
import typing

import torch

from . import noise as _noise


class Diffusion(torch.nn.Module):
    def __init__(self, net: torch.nn.Module, noise_f, prediction_goal: str, shape: typing.Tuple[int, int],
                 loss: torch.nn.Module = torch.nn.MSELoss(reduction="none")) -> None:
        super().__init__()
        self.net = net
        self.prediction_goal = prediction_goal
        self.add_noise = noise_f
        self.width, self.height = shape
        self.loss = loss

    def forward(self, x: typing.Optional[torch.Tensor], **kwargs):
        """Training mode: one training step (incl. backward).  Eval mode: sample.  src/models.py:29-42."""
        if self.training:
            if self.prediction_goal == "data":
                return self.run_training_step_data(x, **kwargs)
            return self.run_training_step_noise(x, **kwargs)
        return self.sample(first_x=x, **kwargs)

    def _ladder(self, x: torch.Tensor, T: int, want_clean: bool = True):
        """src/models.py:46-63: (batch, T+1, pixels) ladder -> noisy = steps 1..T, clean = steps 0..T-1.
        On CUDA with the reference schedule the pair comes out of one kernel (noise.ladder_pair); with `want_clean=False`
        only `noisy` is written and the third return value is the draw (images, eps, level weights) from which the loss
        kernel recomputes its target.  Returns (noisy, clean or None, draw or None)."""
        shape = (-1, 1, self.width, self.height)
        if (x.is_cuda and self.add_noise is _noise.add_normal_noise_multiple
                and x.dtype in (torch.float32, torch.float64) and x.dim() == 2):
            noisy, clean, draw = _noise.ladder_pair(x, T, decay_mod=3.0, want_clean=want_clean, return_draw=True)
            return noisy.reshape(shape), (clean.reshape(shape) if clean is not None else None), draw
        whole = self.add_noise(x, tau=T + 1, decay_mod=3.0).reshape(-1, T + 1, x.shape[-1])
        return whole[:, 1:, :].reshape(shape), whole[:, :-1, :].reshape(shape), None

    def _recompute_target(self, x: torch.Tensor, verbose: bool) -> bool:
        """The loss kernel can recompute the clean ladder levels from the images and the noise draw (no `clean` tensor: a third
        to a half of the ladder's and the loss pass's bytes).  QIDDM_MSE_LADDER=0 switches it off."""
        return (not verbose and os.environ.get("QIDDM_MSE_LADDER", "1") != "0" and x.is_cuda and x.dim() == 2
                and self.add_noise is _noise.add_normal_noise_multiple and x.dtype in (torch.float32, torch.float64)
                and type(self.loss) is torch.nn.MSELoss and self.loss.reduction in ("mean", "none"))

    def _fused_tail(self, noisy, draw, T, verbose, **coef):
        """Re-upload networks that end in `linear_up` (nn/qdense.py:642, :676): hidden features -> loss in ONE pass
        (noise.linear_up_mse_loss: no (rows x pixels) output / gradient tensors).  Returns the loss after `.backward()`, or None
        when not applicable.  The plain FP64 FMA rate bounds the B200 here (~5 TFLOP/s), so the kernels work on expanded
        products (second moments of the weight and of h): 222 us at 40 960 x 784 x 6 against 326 us for the four streaming
        kernels they replace; from 16 384 rows on (below, its five launches cost what it saves).  QIDDM_FUSED_TAIL=0: off."""
        fh = getattr(self.net, "forward_hidden", None)
        if (fh is None or draw is None or verbose or os.environ.get("QIDDM_FUSED_TAIL", "1") == "0"
                or type(self.loss) is not torch.nn.MSELoss or self.loss.reduction not in ("mean", "none")):
            return None
        layer = getattr(self.net, "linear_up", None)
        if not isinstance(layer, torch.nn.Linear) or getattr(self.net, "_restore", "linear") != "linear":
            return None
        if not (layer.in_features <= 16 and noisy.shape[0] >= 16384):         # below: its five launches cost what it saves
            return None
        h = fh(noisy)
        if h is None or not _noise.linear_up_mse_ok(h, layer, draw, T) or not (h.requires_grad or layer.weight.requires_grad):
            return None
        loss = _noise.linear_up_mse_loss(h, layer, draw, T, **coef)
        loss.backward()
        return loss.detach()

    def _clean_of(self, x, T, draw):
        """The clean levels after all (the net's output cannot take the fused loss): same draw, same kernel."""
        _, clean = _noise.ladder_pair(x, T, decay_mod=3.0, eps=draw[1])
        return clean.reshape(-1, 1, self.width, self.height)

    def _fused_mse(self, recon: torch.Tensor, verbose: bool) -> bool:
        """MSELoss followed by `.mean().backward()` == one kernel giving the loss and d loss / d recon."""
        return (not verbose and recon.is_cuda and type(self.loss) is torch.nn.MSELoss and self.loss.reduction in ("mean", "none")
                and recon.dtype in (torch.float32, torch.float64) and recon.requires_grad)

    def _fused_step(self, x: torch.Tensor, kwargs):
        """Single-layer nets that offer `fused_mse_step` (the amplitude-embedding QDense layers on the unitary-collapse path)
        run ladder -> layer -> MSE -> backward as ONE library call; None = not applicable, take the unfused sequence.
        QIDDM_FUSED_STEP=0 switches it off."""
        fn = getattr(self.net, "fused_mse_step", None)
        if (fn is None or kwargs.get("verbose", False) or os.environ.get("QIDDM_FUSED_STEP", "1") == "0"
                or self.add_noise is not _noise.add_normal_noise_multiple or type(self.loss) is not torch.nn.MSELoss
                or self.loss.reduction not in ("mean", "none") or (self.width * self.height) != x.shape[-1]):
            return None
        return fn(x, kwargs["T"], self.prediction_goal, decay_mod=3.0)

    def run_training_step_data(self, x: torch.Tensor, **kwargs):
        loss = self._fused_step(x, kwargs)
        if loss is not None:
            return (loss.abs(),)
        T, verbose = kwargs["T"], kwargs.get("verbose", False)
        noisy, clean, draw = self._ladder(x, T, want_clean=not self._recompute_target(x, verbose))
        if clean is None:
            loss = self._fused_tail(noisy, draw, T, verbose)                                  # target = level_t
            if loss is not None:
                return (loss.abs(),)
        recon = self.net.forward(x=noisy)
        if self._fused_mse(recon, verbose):
            if clean is None and recon.dtype == x.dtype and recon.numel() == noisy.numel():
                loss, grad = _noise.mse_ladder_loss_and_grad(recon, draw, T)                  # target = level_t
            else:
                clean = clean if clean is not None else self._clean_of(x, T, draw)
                loss, grad = _noise.mse_loss_and_grad(recon, clean)
            recon.backward(grad)
            return (loss.abs(),)
        if clean is None:
            clean = self._clean_of(x, T, draw)
        batch_loss = self.loss(recon, clean)
        batch_loss_mean = batch_loss.mean()
        batch_loss_mean.backward()
        if kwargs.get("verbose", False):
            return batch_loss.abs(), recon.abs()
        return (batch_loss_mean.abs(),)

    def run_training_step_noise(self, x: torch.Tensor, **kwargs):
        loss = self._fused_step(x, kwargs)
        if loss is not None:
            return (loss,)
        T, verbose = kwargs["T"], kwargs.get("verbose", False)
        noisy, clean, draw = self._ladder(x, T, want_clean=not self._recompute_target(x, verbose))
        if clean is None:
            loss = self._fused_tail(noisy, draw, T, verbose, scale=0.1, shift=-0.05, c0=-1.0, c1=1.0)
            if loss is not None:
                return (loss,)
        out = self.net.forward(x=noisy)
        if self._fused_mse(out, verbose):
            # predicted_noise = (out - 0.5) * 0.1, target = noisy - clean = level_{t+1} - level_t
            if clean is None and out.dtype == x.dtype and out.numel() == noisy.numel():
                loss, grad = _noise.mse_ladder_loss_and_grad(out, draw, T, scale=0.1, shift=-0.05, c0=-1.0, c1=1.0)
            else:
                clean = clean if clean is not None else self._clean_of(x, T, draw)
                loss, grad = _noise.mse_loss_and_grad(out, noisy.reshape(out.shape), clean.reshape(out.shape), scale=0.1, shift=-0.05)
            out.backward(grad)
            return (loss,)
        if clean is None:
            clean = self._clean_of(x, T, draw)
        predicted_noise = (out - 0.5) * 0.1
        batch_loss = self.loss(predicted_noise, noisy - clean)
        batch_loss_mean = batch_loss.mean()
        batch_loss_mean.backward()
        if kwargs.get("verbose", False):
            return batch_loss, torch.clamp(noisy - predicted_noise, 0, 1)
        return (batch_loss_mean,)

    def sample(self, n_iters, first_x: typing.Optional[torch.Tensor] = None, labels=None, show_progress: bool = False,
               only_last=False, step=1, noise_factor=1.0) -> torch.Tensor:
        """x <- net(x) (goal "data") or x <- clamp(x - 0.1 nf (net(x) - 0.5), 0, 1).  src/models.py:106-147."""
        if first_x is None:
            p = next(self.net.parameters())
            first_x = torch.rand((10, 1, self.width, self.height), device=p.device, dtype=p.dtype)
        outp = [first_x]
        with torch.no_grad():
            x = first_x
            for i in range(n_iters):
                predicted = self.net(x)
                if self.prediction_goal == "data":
                    x = predicted
                else:
                    x = torch.clamp(x - (predicted - 0.5) * 0.1 * noise_factor, 0, 1)
                if i % step == 0:
                    outp.append(x)
        if only_last:
            return outp[-1]
        outp = torch.stack(outp)                                   # iters batch 1 height width
        it, b, _, h, w = outp.shape
        return outp[:, :, 0].permute(0, 2, 1, 3).reshape(it * h, b * w)

    def save_name(self):
        return f"{self.net.save_name()}{'_noise' if self.prediction_goal == 'noise' else ''}"
