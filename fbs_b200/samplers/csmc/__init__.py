from .csmc import csmc_kernel, forward_pass, backward_scanning_pass, backward_sampling_pass, normalise, barker_move
from . import resamplings
