"""`nn` namespace the reference drivers import from (`from nn import ...`, `eval(f"nn.{name}")`,
src/mnist_exm.py:24-25, :419-424).  The reference ships a 1-byte placeholder here."""
from .qconv import QConv2d, _QConv2d_FAST
from .qdense import *  # noqa: F401,F403
from .qdense import __all__ as _qdense_all
from .unet import Conv2d, DownBlock, UNetUndirected, UnetDirected, UpBlock
from .unet_simple import DownBlockS, UNetUndirectedS, UnetDirectedS, UpBlockS
from .utils import autocrop, autopad, get_label_embedding

__all__ = list(_qdense_all) + [
    "QConv2d", "_QConv2d_FAST", "Conv2d", "DownBlock", "UpBlock", "UNetUndirected", "UnetDirected",
    "DownBlockS", "UpBlockS", "UNetUndirectedS", "UnetDirectedS", "autocrop", "autopad", "get_label_embedding",
]
