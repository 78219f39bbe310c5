"""gan_lib_tensorflow_b200: B200-native (sm_100a) hot path of watsonyanghx/GAN_Lib_Tensorflow.

Layer functions keep the reference's names and arguments (gan_lib_tensorflow_b200.common.ops.*,
gan_lib_tensorflow_b200.common.resnet_block); arithmetic runs in hand-written CUDA kernels behind the C ABI
declared in include/ganb200.h.  There is no CPU fallback.
"""
from . import cabi  # noqa: F401

__version__ = "0.1.0"
