"""Drop-in for the reference's ``model/padding.py::GeoCyclicPadding`` (model/padding.py:4-39).

Same constructor, same ``forward(x[B, C, H, W]) -> [B, C, H+2p, W+2p]``, same asserts; the
roll/flip/cat chain is replaced by one CUDA gather kernel (and one deterministic fold-add
kernel in backward) behind ``torch.ops.paradis.geocyclic_pad``.
"""
import torch

from .ops import geocyclic_pad


class GeoCyclicPadding(torch.nn.Module):
    """Cyclic padding layer for equiangular grids with poles."""

    def __init__(self, pad_width):
        super().__init__()
        self.pad_width = pad_width

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.pad_width == 0:
            return x
        assert len(x.shape) == 4, "Input must be 4-dimensional [batch, channels, lat, lon]"
        assert x.shape[3] % 2 == 0, "Number of longitude points must be even"
        return geocyclic_pad(x, self.pad_width)

    def extra_repr(self) -> str:
        return f"pad_width={self.pad_width}"
