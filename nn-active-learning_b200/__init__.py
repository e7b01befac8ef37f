"""nn-active-learning_b200 (import name ``nnal_b200``): B200-native drop-in for the
query-scoring step of jsourati/nn-active-learning.

Sub-modules mirror the reference's module and function names for the hot path
(``patch_utils.get_patches``, ``PW_NN.batch_eval``, ``NNAL_tools.compute_entropy``,
``PW_NNAL.CNN_query`` / ``query_multimg``, ``NNAL.CNN_query``); all arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``include/nnal_b200.h``.  There is
no CPU fallback: importing works anywhere (so host logic can be tested), but the first
compute call raises if ``libnnal_b200.so`` is missing or no B200 is present.
"""
from . import _lib            # noqa: F401
from .engine import Engine, get_engine, reset_engine      # noqa: F401
from . import NN, patch_utils, PW_NN, NNAL_tools, PW_NNAL, NNAL, PW_AL, dist, fi, rep   # noqa: F401

__version__ = '0.1.0'
