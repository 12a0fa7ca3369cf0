"""UNet callers of QConv2d (reference `nn/unet.py:9-190`).  Float64 glue like the reference; the Conv2d factory's
quantum branch is the B200 QConv2d, and BatchNorm2d / the bilinear Upsample run the library's own kernels on CUDA
(qiddm_b200.nn.glue: same parameters and state_dict keys as the torch modules)."""
import os

import torch

from ..functional import run_qconv_up
from .glue import BatchNorm2d, FusedReLU, MaxPool2d, Upsample, fuse_bn_relu
from .qconv import QConv2d
from .utils import autopad, get_label_embedding


# Upsample -> 1 x 1 QConv in one pass (functional.run_qconv_up).  Measured on UNetUndirected(3, 8, 3), 64 images: 3.94 ms per step
# fused against 3.91 ms with the two modules -- the interpolation is redone in the staging of three kernels (forward, weight
# gradient, normalisation term of the image gradient), which costs what the upsample kernel and its 64 MB tensor saved; the
# fused form saves that tensor's memory.  Off by default; QIDDM_UPCONV_FUSION=1 turns it on.
UPCONV_FUSION = os.environ.get("QIDDM_UPCONV_FUSION", "0").lower() in ("1", "true", "yes")


def Conv2d(**kwargs):
    """QConv2d when qdepth > 0, else torch.nn.Conv2d in double.  nn/unet.py:9-24."""
    qdepth = kwargs.pop("qdepth", 3)
    if qdepth > 0:
        return QConv2d(qdepth=qdepth, **kwargs)
    return torch.nn.Conv2d(**kwargs).double()


class UpBlock(torch.nn.Module):
    """nn/unet.py:28-75."""

    def __init__(self, in_channels, out_channels, kernel_size=3, qdepth=3):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.up_conv = torch.nn.Sequential(
            Upsample(scale_factor=2, mode="bilinear"),
            Conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=1, padding=0, qdepth=qdepth),
        ).double()
        self.net = fuse_bn_relu(torch.nn.Sequential(
            Conv2d(in_channels=2 * out_channels, out_channels=out_channels, kernel_size=kernel_size, padding=1,
                   qdepth=qdepth),
            FusedReLU(),
            BatchNorm2d(out_channels, dtype=torch.double),
            Conv2d(in_channels=out_channels, out_channels=out_channels, kernel_size=kernel_size, padding=1,
                   qdepth=qdepth),
            BatchNorm2d(out_channels, dtype=torch.double),
            FusedReLU(),
        )).double()

    def _up_conv(self, x):
        """`Upsample -> 1 x 1 Conv2d` (nn/unet.py:36-41); with a quantum convolution that has the direct form the interpolation
        runs inside the convolution's staging and the upsampled tensor never exists (opt-in: QIDDM_UPCONV_FUSION=1)."""
        up, conv = self.up_conv[0], self.up_conv[1]
        if (UPCONV_FUSION and isinstance(up, Upsample) and isinstance(conv, QConv2d) and not conv.reference_forward
                and conv.kernel_size == (1, 1) and conv.padding == (0, 0) and up.mode == "bilinear" and not up.align_corners
                and up.size is None and up.scale_factor is not None and not getattr(up, "recompute_scale_factor", None)
                and x.dim() == 4 and x.shape[1] == conv.in_channels):
            sf = up.scale_factor
            fh, fw = (sf, sf) if not isinstance(sf, (tuple, list)) else sf
            out = run_qconv_up(conv._spec(), x, conv.weights, int(x.shape[2] * fh), int(x.shape[3] * fw), 1.0 / fh, 1.0 / fw)
            if out is not None:
                return out
        return self.up_conv(x)

    def forward(self, from_down, from_up):
        from_up = self._up_conv(from_up)
        from_down, from_up = autopad(from_down.double(), from_up.double())
        return self.net(torch.cat([from_up, from_down], dim=1).double())


class DownBlock(torch.nn.Module):
    """nn/unet.py:78-116."""

    def __init__(self, in_channels, out_channels, pooling, kernel_size=3, qdepth=3):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.pooling = kernel_size, pooling
        self.net = fuse_bn_relu(torch.nn.Sequential(
            Conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size, qdepth=qdepth,
                   padding=1),
            BatchNorm2d(out_channels, dtype=torch.double),
            FusedReLU(),
            Conv2d(in_channels=out_channels, out_channels=out_channels, kernel_size=kernel_size, qdepth=qdepth,
                   padding=1),
            BatchNorm2d(out_channels, dtype=torch.double),
            FusedReLU(),
        )).double()
        if self.pooling:
            self.pooling_layer = MaxPool2d(kernel_size=2, stride=2)

    def forward(self, x):
        x = self.net(x.double())
        before_pool = x
        if self.pooling:
            x = self.pooling_layer(x)
        return x, before_pool


class UNetUndirected(torch.nn.Module):
    """U-shaped network, undirected (no labels).  nn/unet.py:119-180."""

    def __init__(self, depth=3, start_channels=8, qdepth=3):
        super().__init__()
        self.depth, self.start_channels, self.qdepth = depth, start_channels, qdepth
        assert self.depth > 0, "Depth must be greater than 0"
        out_channel = -1
        down_blocks = []
        for i in range(depth):
            in_channel = 1 if i == 0 else out_channel
            out_channel = start_channels * 2 ** i
            down_blocks.append(DownBlock(in_channel, out_channel, pooling=i < depth - 1, qdepth=qdepth))
        up_blocks = []
        for i in range(depth - 1):
            in_channel = out_channel
            out_channel = out_channel // 2
            up_blocks.append(UpBlock(in_channel, out_channel, qdepth=qdepth))
        self.down_blocks = torch.nn.ModuleList(down_blocks).double()
        self.up_blocks = torch.nn.ModuleList(up_blocks).double()
        self.final_conv = Conv2d(in_channels=out_channel, out_channels=1, kernel_size=1, padding=0,
                                 qdepth=qdepth).double()

    def forward(self, x):
        skips = []
        x = x.double()
        for block in self.down_blocks:
            x, before_pool = block(x)
            skips.append(before_pool)
        for i, block in enumerate(self.up_blocks):
            x = block(skips[-(i + 2)].double(), x.double())
        return self.final_conv(x)

    def extra_repr(self) -> str:
        return f"depth={self.depth}"

    def save_name(self) -> str:
        return f"unet_undirected_d{self.depth}_s{self.start_channels}_d{self.qdepth}"


class UnetDirected(UNetUndirected):
    """nn/unet.py:183-190."""

    def forward(self, x, y):
        mask = get_label_embedding(y.double(), x.shape[2], x.shape[3])
        return super().forward(x.double() + mask)

    def save_name(self) -> str:
        return f"unet_directed_d{self.depth}_s{self.start_channels}_d{self.qdepth}"
