"""CPU oracle for the MMD-GAN SNGan hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (PyTorch-CPU / numpy) of the algorithm the
reference implements in TensorFlow-1.8 for the path named in BASELINE.json:
the SNGan training step (DeepLearning/my_sngan.py) with the DCGAN-style G/D
(GeneralTools/layer_func.py), PICO spectral normalisation and the rep / rmb
MMD losses (GeneralTools/math_func.py).

PARITY UNPINNED: the reference ships no tests, golden vectors or known-answer
fixtures for this path and TensorFlow 1.x cannot be
imported or installed in this image (no cp312 wheel, no network), so the
oracle cannot be checked against outputs of the reference itself.  It is pinned
instead against closed forms and float64 finite differences
(tests/test_oracle_*.py) and every function cites the reference file:line it
follows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this package.  The product (mmd-gan_b200/) never does.
"""
