from .linear import (LinearSDE, StationaryConstLinearSDE, StationaryLinLinearSDE, StationaryExpLinearSDE,
                     make_linear_sde, make_ou_sde)
from .simulators import euler_maruyama, AffineDrift
from .bridges import make_gaussian_bw_sb
