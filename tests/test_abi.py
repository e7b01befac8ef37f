"""C-ABI checks that need no GPU: the library builds/loads, exports every symbol include/nnal_b200.h declares
(and the ctypes table covers exactly those), refuses to create a context without a B200, and the product
package never imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'nnal_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(nnal_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported_and_bound():
    from nnal_b200 import _lib
    names = _declared()
    assert len(names) > 30
    lib = _lib.load()                      # raises if the .so is missing or a bound symbol is not exported
    for n in names:
        assert hasattr(lib, n), 'libnnal_b200.so does not export %s' % n
    assert sorted(_lib.SIGNATURES) == names, 'ctypes table and header disagree: %s' % (
        set(_lib.SIGNATURES) ^ set(names))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from nnal_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.nnal_ctx_create(0, ctypes.byref(h)) == _lib.ERR_NO_DEVICE and not h.value
    import nnal_b200
    nnal_b200.reset_engine()
    with pytest.raises(_lib.NnalError):
        nnal_b200.get_engine()
    import numpy as np
    with pytest.raises(_lib.NnalError):
        nnal_b200.NNAL_tools.compute_entropy(np.array([[.5], [.5]]))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'nn-active-learning_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(import|from)\s+oracle\b', txt, flags=re.M), f
                assert '/root/reference' not in txt, f


def test_host_index_bookkeeping(golden):
    import numpy as np
    import nnal_b200
    got = nnal_b200.patch_utils.global2local_inds(golden['g2l_inds'], golden['g2l_sizes'])
    for s in range(len(got)):
        assert np.array_equal(got[s], golden['g2l_out%d' % s])
