from .pop_wrapper import (population, theta_det2src, get_theta_src_and_weights, p_cbc, pop_rate_det,
                          compute_z_grids)
from . import cosmo
from . import mass
from . import rate
from ..catalog import *  # noqa: F401,F403  (the reference re-exports the catalogue objects here)
