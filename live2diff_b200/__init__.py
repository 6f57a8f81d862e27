"""live2diff_b200 -- B200-native per-frame streaming UNet step for Live2Diff."""
from .weights import UNetDims, random_state_dict, unet_param_spec  # noqa: F401
