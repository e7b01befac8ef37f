"""The built-in HDF5 reader/writer (nnal_b200.hdf5) on the reference's weight-file layout (NN.save_weights,
NN.py:379-396): ``<layer>/Weight``, ``<layer>/Bias``.  CPU only."""
import os
import struct

import numpy as np
import pytest

import nnal_b200
from nnal_b200 import hdf5


def _weights(seed=0):
    m = nnal_b200.NN.create_PW1(2)
    rs = np.random.RandomState(seed)
    return m, {name: (rs.randn(*ws).astype(np.float32), rs.randn(*bs).astype(np.float32))
               for name, (ws, bs) in m.weight_shapes().items() if name in ('conv1', 'conv3', 'fc3')}


def test_round_trip_and_model_io(tmp_path):
    small = nnal_b200.NN.CNN((9, 7, 2), [('conv1', [6, 'conv', [3, 3]]), ('max1', [[2, 2], 'pool']), ('fc1', [10, 'fc']), ('fc2', [3, 'fc'])])
    small.initialize(3, bias_scale=0.2)
    p = str(tmp_path / 'w.h5')
    small.save_weights(p)
    other = nnal_b200.NN.CNN((9, 7, 2), small.layer_dict)
    other.perform_assign_ops(p)                          # NN.perform_assign_ops (NN.py:397-419)
    for name in small.var_dict:
        assert np.array_equal(small.var_dict[name][0], other.var_dict[name][0])
        assert np.array_equal(small.var_dict[name][1], other.var_dict[name][1])
    with hdf5.File(p) as f:
        assert sorted(f.keys()) == ['conv1', 'fc1', 'fc2']
        assert sorted(f['conv1'].keys()) == ['Bias', 'Weight']
        assert f['conv1']['Weight'].shape == (3, 3, 2, 6) and f['conv1/Weight'].dtype == np.float32
        assert np.array_equal(f['fc2']['Bias'][...], small.var_dict['fc2'][1])
        with pytest.raises(KeyError):
            f['nope']
    # NPZ twin keeps working
    q = str(tmp_path / 'w.npz')
    small.save_weights(q)
    other.initialize(9)
    other.load_weights(q)
    assert np.array_equal(small.var_dict['fc1'][0], other.var_dict['fc1'][0])


def test_dtypes_many_members_and_structure(tmp_path):
    rs = np.random.RandomState(1)
    w = {'layer%02d' % i: (rs.randn(3, i + 1).astype([np.float32, np.float64][i % 2]), np.arange(i + 2, dtype=[np.int32, np.int64][i % 2]))
         for i in range(40)}                             # > 32 members: several symbol-table nodes under one B-tree node
    p = str(tmp_path / 'many.h5')
    hdf5.write_weights(p, w)
    r = hdf5.read_weights(p)
    assert sorted(r) == sorted(w)
    for k in w:
        assert r[k][0].dtype == w[k][0].dtype and np.array_equal(r[k][0], w[k][0])
        assert r[k][1].dtype == w[k][1].dtype and np.array_equal(r[k][1], w[k][1])
    raw = open(p, 'rb').read()
    # fixed points of the on-disk format (HDF5 File Format Specification): signature, superblock version 0, 8-byte offsets
    # and lengths, end-of-file address = file size, root symbol-table entry caching the B-tree / heap addresses
    assert raw[:8] == b'\x89HDF\r\n\x1a\n' and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    base, free, eof, drv = struct.unpack('<QQQQ', raw[24:56])
    assert base == 0 and eof == len(raw) and free == drv == 0xFFFFFFFFFFFFFFFF
    name_off, ohdr, cache, _, btree, heap = struct.unpack('<QQIIQQ', raw[56:96])
    assert cache == 1 and raw[btree:btree + 4] == b'TREE' and raw[heap:heap + 4] == b'HEAP' and raw[ohdr] == 1
    assert ohdr % 8 == 0 and btree % 8 == 0


def test_big_endian_and_compact_and_errors(tmp_path):
    """Reader paths the writer does not produce: big-endian data, compact layout, version-2 dataspace; and loud failures."""
    w = hdf5._Writer()
    arr = np.arange(6, dtype='>f4').reshape(2, 3)
    raw = w.alloc(arr.tobytes())
    dt = bytearray(hdf5._dtype_msg(np.float32))
    dt[1] |= 1                                           # byte order bit: big-endian
    space = struct.pack('<BBBB', 2, 2, 0, 1) + struct.pack('<QQ', 2, 3)
    d1 = w.alloc(hdf5._object_header([hdf5._msg(1, space), hdf5._msg(3, bytes(dt)), hdf5._msg(8, struct.pack('<BBQQ', 3, 1, raw, 24))]))
    comp = np.array([7, 8, 9], dtype='<i4')
    d2 = w.alloc(hdf5._object_header([hdf5._msg(1, struct.pack('<BBB5xQ', 1, 1, 0, 3)), hdf5._msg(3, hdf5._dtype_msg(np.int32)),
                                      hdf5._msg(8, struct.pack('<BBH', 3, 0, 12) + comp.tobytes())]))
    d3 = w.alloc(hdf5._object_header([hdf5._msg(1, space), hdf5._msg(3, hdf5._dtype_msg(np.float32)),
                                      hdf5._msg(8, struct.pack('<BB', 3, 2) + b'\x00' * 22)]))
    root, bt, hp = w.group([('be', d1), ('compact', d2), ('chunked', d3)])
    p = str(tmp_path / 'x.h5')
    open(p, 'wb').write(w.finish(root, bt, hp))
    with hdf5.File(p) as f:
        assert np.array_equal(f['be'][...], np.arange(6, dtype=np.float32).reshape(2, 3)) and f['be'].dtype == np.float32
        assert np.array_equal(np.array(f['compact']), comp)
        with pytest.raises(NotImplementedError):
            f['chunked']
    open(p, 'wb').write(b'not hdf5 at all' * 10)
    with pytest.raises(hdf5.Hdf5FormatError):
        hdf5.File(p)
    m = nnal_b200.NN.create_PW1(2)
    q = str(tmp_path / 'partial.h5')
    hdf5.write_weights(q, {'conv1': (np.zeros((5, 5, 3, 24), np.float32), np.zeros(24, np.float32))})
    with pytest.raises(KeyError):
        m.load_weights(q)
