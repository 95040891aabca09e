from typing import NamedTuple, Any


class MCMCState(NamedTuple):
    """``fbs/samplers/common.py:5-9``."""
    acceptance_prob: Any
    is_accepted: Any
    prop_log_ell: Any
    log_ell: Any
