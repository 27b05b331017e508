"""mmdgan_b200 -- B200-native engine for the MMD-GAN (SNGan + repulsive MMD) training hot path.

Host side: Python mirror of the reference's layer DSL / losses / training loop (GeneralTools/layer_func.py,
GeneralTools/math_func.py, DeepLearning/my_sngan.py) over a C-ABI library of hand-written sm_100a CUDA kernels
(csrc/, include/mmdgan_b200.h).  PyTorch supplies device memory, streams and torch.distributed only.
"""
__version__ = '0.1.0'
