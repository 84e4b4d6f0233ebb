"""vela.jl_b200 — B200-native batched posterior hot path behind Vela.jl's API names.

The directory name contains a dot, so it is loaded under the module name `vela_jl_b200`
(see `__graft_entry__.load_package()`).
"""

from . import model, priors, readwrite, descriptor, jlso, lib, api, synth, parallel  # noqa: F401
from .model import *  # noqa: F401,F403
from .readwrite import load_pulsar_data  # noqa: F401
from .descriptor import ModelDescriptor  # noqa: F401
from .api import (Pulsar, calc_lnlike, calc_lnlike_serial, calc_lnlike_vectorized, calc_lnprior,  # noqa: F401
                  prior_transform, calc_lnpost, calc_lnpost_serial, calc_lnpost_vectorized, calc_chi2,
                  calc_chi2_serial, calc_chi2_reduced, degrees_of_freedom, form_residuals,
                  _calc_resids_and_Ninvdiag, calc_noise_weights_inv, get_lnlike_serial_func,
                  get_lnlike_parallel_func, get_lnlike_func, get_chi2_serial_func,
                  get_chi2_parallel_func, get_chi2_func, get_lnprior_func, get_prior_transform_func,
                  get_lnpost_func)
from .lib import VelaB200Error  # noqa: F401
