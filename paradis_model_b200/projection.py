"""Channel projections either side of the fused operator (model/advection.py:38-52, 127, 171-173).

The projections are dense convolutions and stay PyTorch/cuDNN (out of scope for
the hand-written path).  When the package is used inside the reference tree the reference's own
``model.blocks.GMBlock`` is used unchanged; otherwise this minimal builder provides the two layer
kinds the shipped configuration uses (config/paradis_settings.yaml:43-51) with identical submodule
names, so ``state_dict`` keys match (``down_projection.0-SepConv.depthwise.weight`` ...).
"""
from collections import OrderedDict

import torch
from torch import nn

from .blocks import CLinear, SepConv        # one definition of each layer (blocks.py)


_LAYERS = {"SepConv": SepConv, "CLinear": CLinear}


class ProjectionBlock(nn.Sequential):
    """Activation-free stack of SepConv / CLinear layers named ``{idx}-{Type}``
    (the subset of model/blocks.py:210-304 the advection projections use)."""

    def __init__(self, layers, input_dim, output_dim, mesh_size, hidden_dim=0, kernel_size=5, **_ignored):
        if len(layers) == 0:
            raise ValueError("ProjectionBlock: must specify at least one layer")
        if hidden_dim <= 0:
            hidden_dim = max(input_dim, output_dim)
        blocks, cin = [], input_dim
        for idx, name in enumerate(layers):
            if name not in _LAYERS:
                raise ValueError(f"Unknown layer type: {name}. Available standalone: {list(_LAYERS)}; "
                                 "run inside the reference tree for the full registry")
            cout = output_dim if idx == len(layers) - 1 else hidden_dim
            blocks.append((f"{idx}-{name}", _LAYERS[name](input_dim=cin, output_dim=cout, mesh_size=mesh_size,
                                                         kernel_size=kernel_size)))
            if idx < len(layers) - 1:
                blocks.append((f"{idx}-SiLU", nn.SiLU()))
            cin = cout
        super().__init__(OrderedDict(blocks))
        convs = [m for m in self.modules() if isinstance(m, nn.Conv2d)]
        for i, conv in enumerate(convs):  # model/blocks.py:33-54
            nn.init.kaiming_normal_(conv.weight, mode="fan_in", nonlinearity="relu")
            if i == len(convs) - 1:
                with torch.no_grad():
                    conv.weight.mul_(0.1)
            if conv.bias is not None:
                nn.init.constant_(conv.bias, 0.0)


def resolve_block_factory():
    """The reference's GMBlock when importable (drop-in deployment), else ProjectionBlock."""
    try:
        from model.blocks import GMBlock  # type: ignore
        return GMBlock
    except Exception:
        return ProjectionBlock
