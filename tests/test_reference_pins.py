"""Runs oracle/check_against_reference.py where the reference tree exists (the dev container): every pin of the oracle
against the UNMODIFIED reference -- NumPy helpers, the query dispatch over a fake TF session / fake cvxopt, the SDP
programme -- is re-executed, and the goldens it would write must equal the committed ones.  Skipped on boxes without
/root/reference (the GPU box): there the committed goldens stand in."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('NNAL_REFERENCE', '/root/reference')


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')
def test_pins_hold_and_goldens_are_current(tmp_path, golden):
    env = dict(os.environ, NNAL_GOLD_OUT=str(tmp_path), OMP_NUM_THREADS='2')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'oracle', 'check_against_reference.py')], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    out = r.stdout.decode()
    assert r.returncode == 0, out[-3000:]
    for needle in ('get_patches: oracle == reference', "PW_NNAL.CNN_query 'fi'", "query_multimg 'MC-entropy' / 'BALD'",
                   "NNAL.CNN_query 'entropy' / 'rep-entropy' / 'fi'", 'SDP_query_distribution / inequality_cvx_matrix',
                   'LLFC_grads / LLFC_hess'):
        assert needle in out, needle
    fresh = np.load(os.path.join(str(tmp_path), 'reference_numpy_helpers.npz'))
    assert sorted(fresh.files) == sorted(golden.files)
    for k in fresh.files:
        assert np.array_equal(fresh[k], golden[k]), k
