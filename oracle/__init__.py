"""CPU oracle for the nn-active-learning query-scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``nn-active-learning_b200`` / ``nnal_b200``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker / CPU baseline.

Parity status
-------------
* Pure-NumPy pieces (patch gather, multi-image gather + normalise,
  global2local_inds, compute_entropy, uncertainty_filtering, shrink_gradient,
  binary_uncertainty_filter, sample_query_dstr, append_zero) are PINNED against
  the reference's own functions executed unchanged under import stubs:
  ``oracle/check_against_reference.py`` (run in the dev container, where
  /root/reference exists) compares them on seeded inputs and writes the golden
  vectors in ``tests/golden/``.
* TensorFlow-side arithmetic (conv/pool/fc/softmax forward, tf.gradients) is
  "parity unpinned": TF 1.x is not installable here and the reference ships no
  golden vectors for it.  The float64 restatement below encodes
  the documented TF semantics and is cross-checked against torch-CPU float64
  (forward and autograd) in ``tests/test_oracle.py``.
"""
from .nnal_oracle import *   # noqa: F401,F403
from .fi_oracle import *     # noqa: F401,F403
from .rep_oracle import *    # noqa: F401,F403
from .mc_oracle import *     # noqa: F401,F403
