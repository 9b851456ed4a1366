"""Optional drop-ins for the two padding consumers of the reference's `model/blocks.py` that are pure
memory movement plus a tiny stencil: `SepConv` (blocks.py:92-116) and `PhysicalDownsample` (blocks.py:57-71).
Both use `paradis::geocyclic_dwconv`, i.e. the GeoCyclic padding is applied inside the convolution kernel and
the padded tensor is never materialised.  Parameter names match the reference (state_dict compatible)."""
import torch
from torch import nn

from .ops import geocyclic_dwconv
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


class PhysicalDownsample(nn.Module):
    """GeoCyclic pad 2 + 5x5 average pooling with `stride` (blocks.py:57-71).  The pooling has no padding of
    its own, so it is the 5x5 box filter of the GeoCyclic-padded field sampled every `stride` points: one
    fused depthwise convolution with constant taps 1/25."""

    def __init__(self, stride=4):
        super().__init__()
        # both kept for structural parity with the reference (the fused kernel replaces them in forward)
        self.pool = nn.AvgPool2d(kernel_size=5, stride=stride, count_include_pad=False)
        self.padding = GeoCyclicPadding(2)
        self.register_buffer("_box", torch.full((1, 1, 5, 5), 1.0 / 25.0), persistent=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        C = x.shape[1]
        y = geocyclic_dwconv(x, self._box.expand(C, 1, 5, 5).contiguous())
        s = self.pool.stride if isinstance(self.pool.stride, int) else self.pool.stride[0]
        return y if s == 1 else y[:, :, ::s, ::s].contiguous()
