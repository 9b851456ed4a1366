"""Optional drop-ins for the two padding consumers of the reference's `model/blocks.py` that are pure
memory movement plus a tiny stencil: `SepConv` (blocks.py:92-116) and `PhysicalDownsample` (blocks.py:57-71).
Both use `paradis::geocyclic_dwconv`, i.e. the GeoCyclic padding is applied inside the convolution kernel and
the padded tensor is never materialised.  Parameter names match the reference (state_dict compatible)."""
import torch
from torch import nn

from .ops import geocyclic_avgpool5, geocyclic_dwconv
from .padding import GeoCyclicPadding


class SepConv(nn.Module):
    """Separable convolution: GeoCyclic pad + depthwise k x k (fused, k in 3/5/7) + pointwise 1x1."""

    def __init__(self, input_dim: int, output_dim: int, mesh_size: tuple, kernel_size: int = 3, bias: bool = True):
        super().__init__()
        self.padding = (kernel_size - 1) // 2
        self.geo_padding = GeoCyclicPadding(self.padding)
        self.depthwise = nn.Conv2d(input_dim, input_dim, kernel_size, groups=input_dim, bias=False)
        self.pointwise = nn.Conv2d(input_dim, output_dim, kernel_size=1, bias=bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.depthwise.kernel_size[0] in (3, 5, 7):
            x = geocyclic_dwconv(x, self.depthwise.weight)
        else:
            x = self.depthwise(self.geo_padding(x))
        return self.pointwise(x)


class CLinear(nn.Module):
    """1x1 convolution (model/blocks.py:74-89); stays cuDNN."""

    def __init__(self, input_dim, output_dim, mesh_size, kernel_size=1, bias=True):
        super().__init__()
        self.conv = nn.Conv2d(input_dim, output_dim, kernel_size=1, bias=bias)

    def forward(self, x):
        return self.conv(x)


class PhysicalDownsample(nn.Module):
    """GeoCyclic pad 2 + 5x5 average pooling with `stride` (blocks.py:57-71) as one kernel that computes only the
    strided outputs (`paradis::geocyclic_avgpool5`)."""

    def __init__(self, stride=4):
        super().__init__()
        # both kept for structural parity with the reference (the fused kernel replaces them in forward)
        self.pool = nn.AvgPool2d(kernel_size=5, stride=stride, count_include_pad=False)
        self.padding = GeoCyclicPadding(2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        s = self.pool.stride if isinstance(self.pool.stride, int) else self.pool.stride[0]
        return geocyclic_avgpool5(x, s)
