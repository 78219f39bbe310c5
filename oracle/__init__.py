"""CPU oracle: a restatement of the reference's hot path (watsonyanghx/GAN_Lib_Tensorflow) in torch CPU ops.

TEST INFRASTRUCTURE ONLY.  Nothing under gan_lib_tensorflow_b200/ may import this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and only as the checker
or the reported CPU baseline -- never as the product path.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and TensorFlow 1.5 (where the
arithmetic lives) is not installable here, so this oracle cannot be checked against outputs of the
reference itself.  It is pinned instead by its own self-checks (tests/test_oracle_*.py): finite-difference
gradients in float64, sigma against numpy SVD, an independent NumPy loop convolution with TF SAME padding,
closed-form identities (UpsampleConv == conv of np.repeat, ConvMeanPool == avg-pool of conv, CBN with equal
labels == BN) and a hand-computed Adam vector.  TF-1.x semantics that are not visible in the repository are
listed in SURVEY.md 8(c) and encoded in oracle/ops.py with the rule each one follows.

Every function cites the reference file:line it restates (paths relative to the reference root).
"""
