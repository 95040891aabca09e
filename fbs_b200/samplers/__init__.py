"""API of ``fbs/samplers/__init__.py:1-3``."""
from .smc import bootstrap_filter, pmcmc_kernel, pmcmc_filter_step, pcn_proposal, bootstrap_backward_smoother, twisted_smc
from .resampling import multinomial, systematic, stratified, killing
from .gibbs import gibbs_init, gibbs_kernel, force_move
from .common import MCMCState
from . import csmc
