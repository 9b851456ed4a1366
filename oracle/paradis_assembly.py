"""TEST INFRASTRUCTURE -- restatement of the reference's model assembly around the hot path, so that BASELINE
configs 1 and 4 ("PARADIS forward / training step with the fused op") can be checked on the GPU box, where
/root/reference does not exist.  Never imported by the product path (only tests/ and tools/).

What is restated, each pinned against the real classes on the CPU by tests/test_assembly.py (same state_dict keys
and shapes, `load_state_dict(strict=True)` from the reference model, bit-identical forward):

    model/blocks.py:57-71    PhysicalDownsample        -> `Downsample`
    model/blocks.py:74-116   CLinear, SepConv          -> `_clinear`, `SepConvLayer`
    model/blocks.py:118-134  ChannelNorm               -> `ChannelNorm`
    model/blocks.py:138-197  GlobalBias (low rank)     -> `GlobalBias`
    model/blocks.py:210-304  GMBlock builder + init    -> `make_block`
    model/advection.py:10-175 NeuralSemiLagrangian     -> `OracleAdvection` (core = oracle.sl_oracle.sl_advect)
    model/paradis.py:31-269  Paradis                   -> `Assembly`

`Assembly(..., dropin=False)` is the reference model in torch ops (roll/flip/cat padding, grid_sample advection);
`Assembly(..., dropin=True)` is what the maintainer gets after INTEGRATION.md section 1: the same assembly with
`paradis_model_b200.GeoCyclicPadding` and `paradis_model_b200.NeuralSemiLagrangian` in place of the two modules.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F
from torch import nn
from torch.utils.checkpoint import checkpoint

from . import sl_oracle as O

OMEGA = 7.29212e-5          # model/paradis.py:13-14


class AttrDict(dict):
    """cfg shim: attribute + .get access over nested dicts (what the reference needs of OmegaConf)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as exc:
            raise AttributeError(k) from exc
        return AttrDict(v) if isinstance(v, dict) else v

    def get(self, k, d=None):
        v = dict.get(self, k, d)
        return AttrDict(v) if isinstance(v, dict) else v


def default_cfg(latent=16, vels=8, layers=2, interp="bicubic", bias_channels=8, checkpointing=False):
    """The shipped configuration (config/paradis_settings.yaml:1-52, 63, 80) with the three size knobs shrunk."""
    blk = lambda names, hid: {"layers": list(names), "hidden_dim": hid}
    return AttrDict({
        "model": {"latent_size": latent, "base_dt": 21600, "num_layers": layers, "bias_channels": bias_channels,
                  "velocity_vectors": vels, "adv_interpolation": interp, "activation": "SiLU", "coarsening_factor": 1,
                  "physblock": {"input_proj": blk(["CLinear"], 0), "velocity_net": blk(["CLinear", "SepConv"], 12),
                                "diffusion": blk(["SepConv"], 0), "reaction": blk(["CLinear"] * 4, 24),
                                "output_proj": blk(["CLinear"] * 3, 20),
                                "advection": {"down_projection": blk(["SepConv"], 0), "up_projection": blk(["CLinear"], 0)}}},
        "dataset": {"n_time_inputs": 2},
        "compute": {"gradient_checkpointing": checkpointing},
        "features": {"input": {"constants": [f"c{i}" for i in range(10)]}},
    })


class FakeDataModule:
    """The four counts Paradis.__init__ reads (model/paradis.py:55-59, 172)."""

    class dataset:
        num_in_dyn_features = 12
        num_in_static_features = 10
    num_common_features = 5
    num_out_features = 7


# ---------------------------------------------------------------------------------------------------------------
# padding / layers
# ---------------------------------------------------------------------------------------------------------------
class OraclePadding(nn.Module):
    """model/padding.py:4-39 through the oracle's index-map restatement."""

    def __init__(self, pad_width):
        super().__init__()
        self.pad_width = pad_width

    def forward(self, x):
        if self.pad_width == 0:
            return x
        assert x.dim() == 4 and x.shape[3] % 2 == 0
        return O.geocyclic_pad(x, self.pad_width)


class _CLinear(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=1, bias=True)

    def forward(self, x):
        return self.conv(x)


class SepConvLayer(nn.Module):
    def __init__(self, cin, cout, k, pad_cls):
        super().__init__()
        self.padding = (k - 1) // 2
        self.geo_padding = pad_cls(self.padding)
        self.depthwise = nn.Conv2d(cin, cin, k, groups=cin, bias=False)
        self.pointwise = nn.Conv2d(cin, cout, kernel_size=1, bias=True)

    def forward(self, x):
        return self.pointwise(self.depthwise(self.geo_padding(x)))


class ChannelNorm(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.eps = 1e-5
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))

    def forward(self, x):
        var, mean = torch.var_mean(x, dim=-3, keepdim=False)
        inv_std = (self.eps + var) ** -0.5
        y = torch.einsum("...cij,...ij,c->...cij", x - mean[..., None, :, :], inv_std, self.weight)
        return y + self.bias[..., :, None, None]


class GlobalBias(nn.Module):
    def __init__(self, cin, cout, mesh, rank=128):
        super().__init__()
        H, W = mesh
        self.A = nn.Parameter(torch.zeros(cin, rank))
        self.U = nn.Parameter(torch.zeros(rank, H))
        self.V = nn.Parameter(torch.zeros(rank, W))
        with torch.no_grad():
            for p in (self.A, self.U, self.V):
                nn.init.normal_(p, mean=0.0, std=1e-3)
        self.projection = nn.Linear(cin, cout, bias=False) if cin != cout else None

    def forward(self, x):
        maps = torch.einsum("ck,kh,kw->chw", self.A, self.U, self.V)
        if self.projection is not None:
            maps = torch.einsum("oc,chw->ohw", self.projection.weight, maps)
        return x + maps.unsqueeze(0)


def make_block(layers, cin, cout, mesh, pad_cls, k=5, hidden=0, act=nn.SiLU, bias_channels=0, activation=False,
               pre_normalize=False):
    """nn.Sequential with the reference GMBlock's layer names, order and initialisation (blocks.py:210-304, 33-54)."""
    n = len(layers)
    acts = (True,) * (n - 1) + (activation,)
    if hidden <= 0:
        hidden = max(cin, cout)
    seq = []
    if pre_normalize:
        seq.append(("0-ChannelNorm", ChannelNorm(cin)))
    c = cin
    for i, name in enumerate(layers):
        co = cout if i == n - 1 else hidden
        seq.append((f"{i}-{name}", _CLinear(c, co) if name == "CLinear" else SepConvLayer(c, co, k, pad_cls)))
        if i == 0 and bias_channels > 0:
            seq.append(("0-GlobalBias", GlobalBias(bias_channels, co, mesh)))
        if acts[i]:
            seq.append((f"{i}-{act.__name__}", act()))
        c = co
    block = nn.Sequential(OrderedDict(seq))
    convs = []
    for m in block.modules():
        if isinstance(m, nn.Conv2d):
            convs.append(m)
    for i, conv in enumerate(convs):
        nn.init.kaiming_normal_(conv.weight, mode="fan_in", nonlinearity="relu")
        if i == len(convs) - 1:
            with torch.no_grad():
                conv.weight.mul_(0.1)
        if conv.bias is not None:
            nn.init.constant_(conv.bias, 0.0)
    return block


class Downsample(nn.Module):
    def __init__(self, stride, pad_cls):
        super().__init__()
        self.pool = nn.AvgPool2d(kernel_size=5, stride=stride, count_include_pad=False)
        self.padding = pad_cls(2)

    def forward(self, x):
        return self.pool(self.padding(x))


class OracleAdvection(nn.Module):
    """model/advection.py: projections + the oracle's op replay of lines 129-169."""

    def __init__(self, cfg, hidden, mesh, num_vels, lat_grid, lon_grid, interpolation, pad_cls):
        super().__init__()
        a = cfg.model.physblock.advection
        self.padding_interp = pad_cls(2 if interpolation == "bicubic" else 1)
        self.interpolation = interpolation
        self.down_projection = make_block(a.down_projection.layers, hidden, num_vels, mesh, pad_cls,
                                          hidden=a.down_projection.hidden_dim)
        self.up_projection = make_block(a.up_projection.layers, num_vels, hidden, mesh, pad_cls,
                                        hidden=a.up_projection.hidden_dim)
        self.register_buffer("lat2d", lat_grid.clone(), persistent=False)
        self.register_buffer("lon2d", lon_grid.clone(), persistent=False)

    def forward(self, hidden, u, v, dt):
        core = O.sl_advect(self.down_projection(hidden), u, v, self.lat2d, self.lon2d, dt, self.interpolation)
        return self.up_projection(core)


# ---------------------------------------------------------------------------------------------------------------
# the model
# ---------------------------------------------------------------------------------------------------------------
class Assembly(nn.Module):
    def __init__(self, datamodule, cfg, lat_grid, lon_grid, dropin: bool, math: str = "fast"):
        super().__init__()
        if dropin:
            import paradis_model_b200 as pkg
            pad_cls = pkg.GeoCyclicPadding
        else:
            pad_cls = OraclePadding
        m = cfg.model
        self.nlat, self.nlon = lat_grid.shape
        mesh = (self.nlat, self.nlon)
        hidden, self.num_vels = m.get("latent_size"), m.get("velocity_vectors")
        nb = m.get("bias_channels", 4)
        self.num_layers = max(1, m.num_layers)
        self.dt = m.get("base_dt") * OMEGA / self.num_layers
        act = {"SiLU": nn.SiLU, "GELU": nn.GELU}[m.activation]
        cin = datamodule.dataset.num_in_dyn_features + datamodule.dataset.num_in_static_features
        self.gradient_checkpoint = cfg.compute.get("gradient_checkpointing", False)
        pb = m.physblock
        stride = m.get("coarsening_factor", 1)
        self.nlat_coarse, self.nlon_coarse = (self.nlat - 1) // stride + 1, self.nlon // stride
        cmesh = (self.nlat_coarse, self.nlon_coarse)
        static_dim = 128
        self.input_proj = make_block(pb.input_proj.layers, cin, hidden, mesh, pad_cls, hidden=pb.input_proj.hidden_dim,
                                     activation=True, act=act)
        L = range(self.num_layers)
        self.velocity_nets = nn.ModuleList([
            make_block(pb.velocity_net.layers, hidden, 2 * self.num_vels, cmesh, pad_cls, hidden=pb.velocity_net.hidden_dim,
                       bias_channels=nb, act=act, pre_normalize=True) for _ in L])
        if dropin:
            self.advection = nn.ModuleList([
                pkg.NeuralSemiLagrangian(cfg, hidden, cmesh, num_vels=self.num_vels, lat_grid=lat_grid[::stride, ::stride],
                                         lon_grid=lon_grid[::stride, ::stride], interpolation=m.get("adv_interpolation"),
                                         math=math) for _ in L])
        else:
            self.advection = nn.ModuleList([
                OracleAdvection(cfg, hidden, cmesh, self.num_vels, lat_grid[::stride, ::stride],
                                lon_grid[::stride, ::stride], m.get("adv_interpolation"), pad_cls) for _ in L])
        self.diffusion = nn.ModuleList([
            make_block(pb.diffusion.layers, hidden, hidden, cmesh, pad_cls, hidden=pb.diffusion.hidden_dim,
                       pre_normalize=True, act=act, bias_channels=nb) for _ in L])
        self.reaction = nn.ModuleList([
            make_block(pb.reaction.layers, hidden + static_dim, hidden, cmesh, pad_cls, hidden=pb.reaction.hidden_dim,
                       pre_normalize=True, act=act, bias_channels=nb) for _ in L])
        self.output_proj = make_block(pb.output_proj.layers, hidden, datamodule.num_out_features, mesh, pad_cls,
                                      hidden=pb.output_proj.hidden_dim, pre_normalize=True, activation=False, act=act,
                                      bias_channels=nb)
        self.alpha_adv = nn.Parameter(torch.full((self.num_layers, hidden), -1.0))
        self.downsample = Downsample(stride, pad_cls)
        self.n_static = ns = len(cfg.features.input.constants)
        self.static_encoder = nn.Sequential(
            SepConvLayer(ns, 64, 7, pad_cls), nn.SiLU(), pad_cls(3), nn.Conv2d(64, 64, groups=64, kernel_size=7), nn.SiLU(),
            SepConvLayer(64, static_dim, 5, pad_cls))

    def _ckpt(self, fn, *args):
        return checkpoint(fn, *args, use_reentrant=False) if self.gradient_checkpoint else fn(*args)

    def upsample(self, x):
        ext = torch.cat([x, x[..., :1]], dim=-1)
        y = F.interpolate(ext, size=(self.nlat, self.nlon + 1), mode="bilinear", align_corners=True)
        return y[..., :-1]

    def layer_step(self, i, hidden, hidden_static):
        B = hidden.shape[0]
        vel = self.velocity_nets[i](hidden).view(B, 2, self.num_vels, self.nlat_coarse, self.nlon_coarse)
        u, v = vel[:, 0], vel[:, 1]
        gate = torch.sigmoid(self.alpha_adv[i]).to(hidden.dtype).view(1, -1, 1, 1)
        hidden = hidden + gate * (self.advection[i](hidden, u, v, self.dt) - hidden)
        hidden = hidden + self.diffusion[i](hidden)
        return hidden + self.reaction[i](torch.cat([hidden, hidden_static], dim=1))

    def forward(self, fields):
        hidden = self._ckpt(self.input_proj, fields)
        hstat = self._ckpt(self.static_encoder, fields[:, -self.n_static:])
        skip = hidden
        hidden, hstat = self.downsample(hidden), self.downsample(hstat)
        for i in range(self.num_layers):
            hidden = self._ckpt(self.layer_step, i, hidden, hstat)
        return self._ckpt(self.output_proj, self.upsample(hidden) + skip)
