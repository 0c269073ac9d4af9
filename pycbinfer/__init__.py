"""Drop-in alias: ``import pycbinfer`` resolves to the B200-native implementation."""
from cbinfer_b200 import *          # noqa: F401,F403
from cbinfer_b200 import conv2d, conv2d_cg, conv2d_fg, verbose  # noqa: F401
