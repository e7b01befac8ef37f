"""CPU checks of arithmetic the device kernels rely on (no GPU): NumPy emulations of the exact float32 sequences."""
import numpy as np
import pytest


@pytest.mark.parametrize('mu,sg', [(100.123456789, 30.0123456), (0.0371, 1.7e-3), (5000.7, 0.931), (1e-3, 250.0)])
def test_fused_conv1_two_term_normalisation_error(mu, sg):
    """conv_tc.cu's fused gather normalises in float32 two-term arithmetic, d = (x - mu_hi) - mu_lo,
    v = fma(d, r_lo, d * r_hi) with r = 1 / sigma, instead of batch_eval's float64 (x - mu) / sigma
    (PW_NN.py:357-539): the result is within 2^-22 (relative) of the float64 quotient -- below what the fp16 hi/lo operand
    pair (22 significant bits) carries into the tensor core anyway."""
    rs = np.random.RandomState(0)
    x = np.maximum(rs.standard_normal(400000) * sg * 1.3 + mu, 0).astype(np.float32)
    x[:100] = 0
    mh = np.float32(mu)
    ml = np.float32(mu - np.float64(mh))
    r = 1.0 / sg
    rh = np.float32(r)
    rl = np.float32(r - np.float64(rh))
    d = ((x - mh).astype(np.float32) - ml).astype(np.float32)
    p = (d * rh).astype(np.float32)
    v = (d.astype(np.float64) * np.float64(rl) + p.astype(np.float64)).astype(np.float32)     # fma: one rounding
    exact = (x.astype(np.float64) - mu) / sg
    nz = np.abs(exact) > 0
    rel = np.abs(v.astype(np.float64) - exact)[nz] / np.abs(exact)[nz]
    assert rel.max() < 2.0 ** -22
