"""One-QConv-per-block UNet variant (reference `nn/unet_simple.py:6-94`)."""
import torch

from .glue import BatchNorm2d, Upsample
from .qconv import QConv2d
from .unet import DownBlock, UNetUndirected, UpBlock, get_label_embedding


class DownBlockS(DownBlock):
    def __init__(self, in_channels, out_channels, pooling, kernel_size=3, qdepth=3):
        super().__init__(in_channels, out_channels, pooling, kernel_size, qdepth)
        self.net = torch.nn.Sequential(
            QConv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size, qdepth=qdepth,
                    padding=1),
            BatchNorm2d(out_channels),
        )


class UpBlockS(UpBlock):
    def __init__(self, in_channels, out_channels, kernel_size=3, qdepth=3):
        super().__init__(in_channels, out_channels, kernel_size, qdepth=0)
        self.net = torch.nn.Sequential(
            QConv2d(in_channels=2 * out_channels, out_channels=out_channels, kernel_size=kernel_size, padding=1,
                    qdepth=qdepth),
            BatchNorm2d(out_channels),
        )
        self.up_conv = torch.nn.Sequential(
            Upsample(scale_factor=2, mode="bilinear"),
            QConv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=1, padding=0, qdepth=qdepth),
        )


class UNetUndirectedS(UNetUndirected):
    def __init__(self, depth=3, start_channels=8, qdepth=3):
        super().__init__(depth, start_channels, qdepth=0)
        self.qdepth = qdepth
        self.down_blocks = torch.nn.ModuleList([
            DownBlockS(in_channels=db.in_channels, out_channels=db.out_channels, pooling=db.pooling,
                       kernel_size=db.kernel_size, qdepth=qdepth) for db in self.down_blocks])
        self.up_blocks = torch.nn.ModuleList([
            UpBlockS(in_channels=ub.in_channels, out_channels=ub.out_channels, kernel_size=ub.kernel_size,
                     qdepth=qdepth) for ub in self.up_blocks])

    def save_name(self) -> str:
        return f"unet_s_undirected_d{self.depth}_s{self.start_channels}_d{self.qdepth}"


class UnetDirectedS(UNetUndirectedS):
    def forward(self, x, y):
        mask = get_label_embedding(y, x.shape[2], x.shape[3])
        return super().forward(x + mask)

    def save_name(self) -> str:
        return f"unet_s_directed_d{self.depth}_s{self.start_channels}_d{self.qdepth}"
