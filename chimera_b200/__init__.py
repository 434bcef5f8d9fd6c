"""chimera_b200 -- B200-native (sm_100a CUDA) implementation of CHIMERA's hierarchical-likelihood
hot path behind the reference's object surface.

    from chimera_b200 import hyperlikelihood, selection_function, population, cosmo, mass, rate
"""
__version__ = "0.1.0"

from . import data
from .data import theta_pe_det, theta_inj_det, theta_src
from .population import (population, theta_det2src, get_theta_src_and_weights, p_cbc, pop_rate_det,
                         compute_z_grids, cosmo, mass, rate)
from .catalog import completeness, empty_catalog, pixelated_catalog, dVdz_completeness
from .likelihood import hyperlikelihood
from .selection_function import selection_function
from . import parallel
from . import sky
from . import sampling
from .sky import pixelize_gw_catalog
