"""scgbm_b200 -- B200-native (sm_100a) implementation of scGBM's Poisson bilinear
model fit (`gbm.sc`, `get.se`) behind the reference's interface.  The numerics
live in libscgbm.so (hand-written CUDA, C ABI in include/scgbm.h)."""
from .api import gbm_sc, get_se, RError  # noqa: F401
from .projection import gbm_proj_parallel  # noqa: F401
from .simulate import sim_data, sim_null, sim_null_nb, cci_perturb  # noqa: F401
from ._lib import ScgbmError, load  # noqa: F401
