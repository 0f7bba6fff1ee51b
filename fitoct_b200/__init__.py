"""fitoct_b200 — B200-native batched NUTS engine for the FitOCT decay models (fitMonoExp / fitExpGP).

Host-side mirror of the reference's R interface (FitOCTLib::fitExpGP / fitMonoExp as called from
FitOCT.R:95,110-124) over the C ABI in include/fitoct_b200.h.  The CUDA library is loaded lazily on first
use; there is no CPU fallback.
"""
from ._abi import (FOCT_EXPGP, FOCT_MONOEXP, default_cfg, default_spec, make_problems, make_problems_dense,  # noqa: F401
                   param_names)

__all__ = ["fitExpGP", "fitMonoExp", "fitExpGP_batch", "StanFit", "load_ctrl_params", "monitor"]


def __getattr__(name):
    if name in __all__:
        from . import api

        return getattr(api, name)
    raise AttributeError(name)
