"""CPU oracle for the MMD-GAN SNGan hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (PyTorch-CPU / numpy) of the algorithm the
reference implements in TensorFlow-1.8 for the path named in BASELINE.json:
the SNGan training step (DeepLearning/my_sngan.py) with the DCGAN-style G/D
(GeneralTools/layer_func.py), PICO spectral normalisation and the rep / rmb
MMD losses (GeneralTools/math_func.py).

PARITY PINNED BY EXECUTING THE REFERENCE'S OWN PYTHON.  The reference ships no tests, golden vectors or known-answer
fixtures for this path and TensorFlow 1.x cannot be installed in this image (no cp312
wheel, no network).  Instead the UNMODIFIED reference modules (GeneralTools/math_func.py, layer_func.py, graph_func.py,
DeepLearning/my_sngan.py) are imported on top of `oracle/tfshim` -- an eager PyTorch-CPU float64 stand-in for the
tf.* calls they make -- and the reference's own GANLoss / get_squared_dist / mmd_g / mmd_g_bounded / mixture_mmd_g,
SpectralNorm, Net / Routine / ParametricOperation and SNGan.__gpu_task__ + multi_opt_config + apply_gradients +
UPDATE_OPS are executed (tests/golden/make_reference_fixtures.py -> tests/golden/ref_*.npz, committed).  This oracle
reproduces those outputs to <= 4e-15 relative (losses bit-exact): tests/test_oracle_golden.py.  Caveat, stated
plainly: the TensorFlow LIBRARY primitives underneath the reference's Python (conv2d SAME padding, conv2d_transpose
as conv2d-backprop-input, Maximum/Minimum tie gradients, fused batch norm, AdamOptimizer) are third-party code absent
from /root/reference (TensorFlow 1.8.0, pinned only as a string in misc_fun.py:33) and are restated in the shim from
their published behaviour.  Independent checks remain: closed forms, a double-loop restatement and float64 finite
differences (tests/test_oracle_selfcheck.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this package.  The product (mmd-gan_b200/) never does.
"""
