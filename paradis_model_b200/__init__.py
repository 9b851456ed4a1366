"""paradis_model_b200 -- B200-native semi-Lagrangian advection operator for PARADIS.

Scope: the hot path model/advection.py:129-169 + model/padding.py:11-39 of
Wx-Alliance-Alliance-Meteo/paradis_model, as hand-written sm_100a kernels behind a C ABI
(include/paradis_sl.h), exposed as torch.library custom ops and as drop-in modules.
"""
from .ops import (SLGeometry, check_status, geocyclic_avgpool5, geocyclic_dwconv, geocyclic_pad, host_fwd_bwd,  # noqa: F401
                  poll_status, sl_advect)
from .padding import GeoCyclicPadding  # noqa: F401
from .advection import NeuralSemiLagrangian  # noqa: F401

__version__ = "0.1.0"
