"""Score network of zgbkdlm/fbs (fbs/nn/unet.py) on sm_100a tensor cores, and the NN-score closures."""
from .unet import ScoreUNet, ScoreNetModel, random_unet_params, unet_param_shapes  # noqa: F401
