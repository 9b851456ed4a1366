"""Synthetic meshes and inputs for benchmarks and examples (SURVEY 8d shapes)."""
import math

import numpy as np
import torch

DT_DEFAULT = 21600 * 7.29212e-5 / 8  # model/paradis.py:13-14,50 with base_dt 21600 s, 8 layers


def make_grids(H: int, W: int, poles: bool):
    """[H, W] lat/lon grids in radians, fp64 deg2rad then fp32 (data/era5_dataset.py:178-182).
    poles=True: lat = linspace(-90, 90, H); poles=False: cell-centred WB2-style latitudes."""
    lat = np.linspace(-90.0, 90.0, H) if poles else -90.0 + 180.0 / H * (np.arange(H) + 0.5)
    lon = 360.0 / W * np.arange(W)
    lat_g, lon_g = np.meshgrid(np.deg2rad(lat), np.deg2rad(lon), indexing="ij")
    return torch.from_numpy(lat_g).float(), torch.from_numpy(lon_g).float()


def white_noise_inputs(H, W, B, V, dt=DT_DEFAULT, seed=0, cells_sigma=2.0, cells_clip=4.0, pin=False):
    """field ~ N(0,1); u, v ~ N(0, sigma^2), sigma = cells_sigma * dphi / dt, clipped at
    +-cells_clip latitude cells (the CFL bound of the benchmark); grad_out ~ N(0,1).  CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    dphi = math.pi / H
    sigma, clip = cells_sigma * dphi / dt, cells_clip * dphi / dt
    mk = lambda: torch.empty(B, V, H, W, pin_memory=pin)
    field = mk().normal_(generator=g)
    u = mk().normal_(generator=g).mul_(sigma).clamp_(-clip, clip)
    v = mk().normal_(generator=g).mul_(sigma).clamp_(-clip, clip)
    grad_out = mk().normal_(generator=g)
    return field, u, v, grad_out
