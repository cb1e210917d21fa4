"""focusflow_official_b200 -- B200-native correlation hot path of FocusFlow.

Public surface (same names as the reference modules they replace):
    CorrBlock, coords_grid          <- FF_RAFT_Core/corr.py, utils/utils.py
    FunctionCorrelation, ModuleCorrelation  <- PWCNet_Core/correlation.py
"""
from .corr import (AlternateCorrBlock, CorrBlock, coords_grid, get_sampler_semantics, set_sampler_semantics, correlation_pyramid, correlation_volume, lookup, lookup_tiled,  # noqa: F401
                   tile_levels, tiled_pyramid, untile_levels)
from .correlation import FunctionCorrelation, ModuleCorrelation, correlation_leaky  # noqa: F401

__version__ = "0.1.0"
