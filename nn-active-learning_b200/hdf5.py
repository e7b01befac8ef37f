"""Minimal HDF5 reader/writer for the reference's weight files (NN.save_weights / NN.perform_assign_ops,
NN.py:379-419): a file of groups ``<layer>`` holding the datasets ``Weight`` and ``Bias``.

h5py (libhdf5) is not part of this image, so the subset of the HDF5 file format that such files use is implemented
here from the format specification ("HDF5 File Format Specification Version 3.0"):

* reading: superblock versions 0-3; object headers version 1 and 2 (with continuation blocks); old-style groups
  (symbol-table message -> version-1 B-tree -> symbol-table nodes -> local heap) and compact new-style groups (link
  messages); datasets with contiguous or compact layout (layout message versions 3 and 4, and the version-1/2 contiguous
  form), simple dataspaces (versions 1 and 2), fixed-point and IEEE floating-point datatypes of either byte order.
  Chunked / filtered datasets, dense (fractal-heap) groups and variable-length types are outside what
  ``f.create_group(layer).create_dataset('Weight', data=array)`` produces and raise ``NotImplementedError``.
* writing: what h5py writes with its default ``libver='earliest'``: superblock version 0, version-1 object headers,
  symbol-table groups, contiguous little-endian datasets (float32/float64/int32/int64).

Only NumPy and the standard library are used; nothing here touches the GPU (file formats are caller-side plumbing,
SURVEY.md 8f rank 3)."""
import struct

import numpy as np

SIG = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5FormatError(ValueError):
    pass


# ======================================================================================================================
# reader
# ======================================================================================================================
class _Reader(object):
    def __init__(self, buf):
        self.b = buf
        self.so = 8          # size of offsets
        self.sl = 8          # size of lengths
        self.base = 0

    def u(self, off, n):
        return int.from_bytes(self.b[off:off + n], 'little')

    def addr(self, off):
        v = self.u(off, self.so)
        return None if v == (1 << (8 * self.so)) - 1 else v + self.base


class Dataset(object):
    def __init__(self, name, array):
        self.name = name
        self._a = array
        self.shape = array.shape
        self.dtype = array.dtype

    def __getitem__(self, key):
        return self._a[key]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._a, dtype=dtype)

    @property
    def value(self):
        return self._a


class Group(object):
    def __init__(self, file, name, links):
        self._f = file
        self.name = name
        self._links = links              # name -> object header address

    def keys(self):
        return list(self._links.keys())

    def __iter__(self):
        return iter(self._links)

    def __len__(self):
        return len(self._links)

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in str(path).split('/') if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError("Unable to open object (object '%s' doesn't exist)" % part)
            node = node._f._object(node._links[part], (node.name.rstrip('/') + '/' + part))
        return node

    def items(self):
        return [(k, self[k]) for k in self._links]


class File(Group):
    """Read-only view of an HDF5 file: ``File(path)['conv1']['Weight'][...]`` / ``np.array(f['conv1/Bias'])``."""

    def __init__(self, path, mode='r'):
        if mode != 'r':
            raise ValueError('File is read-only; use write_weights() to create files')
        with open(path, 'rb') as fh:
            data = fh.read()
        self._r = _Reader(data)
        self._cache = {}
        root = self._superblock()
        Group.__init__(self, self, '/', self._group_links(root))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass

    # -- superblock -------------------------------------------------------------------------------------------------
    def _superblock(self):
        r = self._r
        off = 0
        while r.b[off:off + 8] != SIG:                      # the superblock may sit at 0, 512, 1024, ...
            off = 512 if off == 0 else off * 2
            if off >= len(r.b):
                raise Hdf5FormatError('not an HDF5 file (signature not found)')
        ver = r.b[off + 8]
        if ver in (0, 1):
            r.so, r.sl = r.b[off + 13], r.b[off + 14]
            p = off + 24 + (4 if ver == 1 else 0)
            r.base = r.u(p, r.so)
            p += 4 * r.so                                   # base, free-space, end-of-file, driver-info addresses
            # root group symbol-table entry: link name offset, object header address, cache type, reserved, scratch
            return r.addr(p + r.so)
        if ver in (2, 3):
            r.so, r.sl = r.b[off + 9], r.b[off + 10]
            p = off + 12
            r.base = r.u(p, r.so)
            return r.addr(p + 3 * r.so)                     # base, superblock extension, end of file, ROOT object header
        raise Hdf5FormatError('unsupported superblock version %d' % ver)

    # -- object headers ---------------------------------------------------------------------------------------------
    def _messages(self, addr):
        """[(type, flags, data offset, size)] of the object header at ``addr`` (continuations followed)."""
        r = self._r
        msgs = []
        if r.b[addr:addr + 4] == b'OHDR':
            if r.b[addr + 4] != 2:
                raise Hdf5FormatError('unsupported object header version')
            flags = r.b[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16                                      # access, modification, change, birth times
            if flags & 0x10:
                p += 4                                       # max compact / min dense
            csz = 1 << (flags & 3)
            chunk0 = r.u(p, csz)
            p += csz
            blocks = [(p, p + chunk0)]
            track = bool(flags & 0x04)
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 + (2 if track else 0) <= end:
                    mtype, size, mflags = r.b[p], r.u(p + 1, 2), r.b[p + 3]
                    p += 4 + (2 if track else 0)
                    if mtype == 0x10:                        # continuation: offset, length; block = 'OCHK' ... checksum
                        co, cl = r.addr(p), r.u(p + r.so, r.sl)
                        blocks.append((co + 4, co + cl - 4))
                    elif mtype != 0:
                        msgs.append((mtype, mflags, p, size))
                    p += size
            return msgs
        ver = r.b[addr]
        if ver != 1:
            raise Hdf5FormatError('unsupported object header version %d at %d' % (ver, addr))
        nmsg = r.u(addr + 2, 2)
        hsize = r.u(addr + 8, 4)
        blocks = [(addr + 16, addr + 16 + hsize)]
        while blocks and len(msgs) < nmsg + 64:
            p, end = blocks.pop(0)
            while p + 8 <= end:
                mtype, size, mflags = r.u(p, 2), r.u(p + 2, 2), r.b[p + 4]
                p += 8
                if mtype == 0x10:
                    blocks.append((r.addr(p), r.addr(p) + r.u(p + r.so, r.sl)))
                elif mtype != 0:
                    msgs.append((mtype, mflags, p, size))
                p += size
        return msgs

    def _object(self, addr, name):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr)
        types = set(m[0] for m in msgs)
        if 0x08 in types or 0x01 in types and 0x03 in types:
            obj = self._dataset(msgs, name)
        else:
            obj = Group(self, name, self._group_links(addr, msgs))
        self._cache[addr] = obj
        return obj

    # -- groups -----------------------------------------------------------------------------------------------------
    def _group_links(self, addr, msgs=None):
        r = self._r
        msgs = self._messages(addr) if msgs is None else msgs
        links = {}
        for mtype, mflags, p, size in msgs:
            if mtype == 0x11:                                # symbol table: B-tree address, local heap address
                btree, heap = r.addr(p), r.addr(p + r.so)
                if r.b[heap:heap + 4] != b'HEAP':
                    raise Hdf5FormatError('bad local heap signature')
                heap_data = r.addr(heap + 8 + 2 * r.sl)
                self._walk_btree(btree, heap_data, links)
            elif mtype == 0x06:                              # link message (compact new-style group)
                ver, lf = r.b[p], r.b[p + 1]
                q = p + 2
                ltype = 0
                if lf & 0x08:
                    ltype = r.b[q]
                    q += 1
                if lf & 0x04:
                    q += 8
                if lf & 0x10:
                    q += 1
                nsz = 1 << (lf & 3)
                nlen = r.u(q, nsz)
                q += nsz
                lname = bytes(r.b[q:q + nlen]).decode('utf-8')
                q += nlen
                if ltype == 0:
                    links[lname] = r.addr(q)
            elif mtype == 0x02 and ver_dense(r, p):
                raise NotImplementedError('dense (fractal-heap) groups are not supported')
        return links

    def _walk_btree(self, addr, heap_data, links):
        r = self._r
        if addr is None:
            return
        sig = bytes(r.b[addr:addr + 4])
        if sig == b'TREE':
            if r.b[addr + 4] != 0:
                raise Hdf5FormatError('not a group B-tree')
            n = r.u(addr + 6, 2)
            p = addr + 8 + 2 * r.so
            for i in range(n):                               # key_i (length), child_i (offset), ..., key_n
                child = r.addr(p + r.sl + i * (r.sl + r.so))
                self._walk_btree(child, heap_data, links)
        elif sig == b'SNOD':
            n = r.u(addr + 6, 2)
            p = addr + 8
            esz = 2 * r.so + 8 + 16
            for i in range(n):
                e = p + i * esz
                noff, oaddr = r.u(e, r.so), r.addr(e + r.so)
                s = heap_data + noff
                t = r.b.find(b'\x00', s)
                links[bytes(r.b[s:t]).decode('utf-8')] = oaddr
        else:
            raise Hdf5FormatError('unknown B-tree node signature %r' % sig)

    # -- datasets ---------------------------------------------------------------------------------------------------
    def _dataset(self, msgs, name):
        r = self._r
        shape = dtype = None
        data = None
        for mtype, mflags, p, size in msgs:
            if mtype == 0x01:                                # dataspace
                ver, rank, fl = r.b[p], r.b[p + 1], r.b[p + 2]
                q = p + (8 if ver == 1 else 4)
                shape = tuple(r.u(q + i * r.sl, r.sl) for i in range(rank))
            elif mtype == 0x03:                              # datatype
                cv = r.b[p]
                cls, ver = cv & 0x0f, cv >> 4
                bits0 = r.b[p + 1]
                nbytes = r.u(p + 4, 4)
                order = '>' if bits0 & 1 else '<'
                if cls == 0:
                    dtype = np.dtype('%s%s%d' % (order, 'i' if bits0 & 0x08 else 'u', nbytes))
                elif cls == 1:
                    if nbytes not in (2, 4, 8):
                        raise NotImplementedError('floating-point size %d' % nbytes)
                    dtype = np.dtype('%sf%d' % (order, nbytes))
                else:
                    raise NotImplementedError('HDF5 datatype class %d' % cls)
            elif mtype == 0x08:                              # data layout
                ver = r.b[p]
                if ver in (3, 4):
                    lclass = r.b[p + 1]
                    if lclass == 1:
                        data = ('contig', r.addr(p + 2), r.u(p + 2 + r.so, r.sl))
                    elif lclass == 0:
                        n = r.u(p + 2, 2)
                        data = ('compact', p + 4, n)
                    else:
                        raise NotImplementedError('chunked / virtual dataset layouts are not supported')
                elif ver in (1, 2):
                    rank, lclass = r.b[p + 1], r.b[p + 2]
                    if lclass != 1:
                        raise NotImplementedError('only contiguous version-1/2 layouts are supported')
                    data = ('contig', r.addr(p + 8), None)
                else:
                    raise Hdf5FormatError('unsupported layout message version %d' % ver)
            elif mtype == 0x0B:
                raise NotImplementedError('filtered datasets are not supported')
        if shape is None or dtype is None or data is None:
            raise Hdf5FormatError('dataset %s lacks a dataspace, datatype or layout message' % name)
        count = int(np.prod(shape)) if shape else 1
        nb = count * dtype.itemsize
        if data[0] == 'contig':
            if data[1] is None:
                arr = np.zeros(shape, dtype=dtype.newbyteorder('='))        # never written: fill value 0
            else:
                arr = np.frombuffer(r.b, dtype=dtype, count=count, offset=data[1]).reshape(shape)
        else:
            arr = np.frombuffer(r.b, dtype=dtype, count=count, offset=data[1]).reshape(shape)
        if len(r.b) < (data[1] or 0) + (nb if data[1] is not None else 0):
            raise Hdf5FormatError('dataset %s extends past the end of the file' % name)
        return Dataset(name, arr.astype(dtype.newbyteorder('='), copy=True))


def ver_dense(r, p):
    """Link-info message: is the group stored densely (fractal heap address defined)?"""
    flags = r.b[p + 1]
    q = p + 2 + (8 if flags & 1 else 0)
    return r.addr(q) is not None


def read_weights(path):
    """{layer: (Weight, Bias)} of a reference weight file (NN.save_weights layout)."""
    out = {}
    with File(path) as f:
        for layer in f.keys():
            g = f[layer]
            if isinstance(g, Group) and 'Weight' in g.keys() and 'Bias' in g.keys():
                out[layer] = (np.array(g['Weight']), np.array(g['Bias']))
    return out


# ======================================================================================================================
# writer (superblock v0, object headers v1, symbol-table groups, contiguous datasets -- h5py's 'earliest' format)
# ======================================================================================================================
def _pad8(b):
    return b + b'\x00' * (-len(b) % 8)


def _msg(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack('<HHB3x', mtype, len(data), flags) + data


def _object_header(msgs):
    body = b''.join(msgs)
    return struct.pack('<BBHII4x', 1, 0, len(msgs), 1, len(body)) + body


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == 'f' and dt.itemsize in (4, 8):
        # class 1 (floating point), version 1; bit field: little-endian, IEEE mantissa normalisation (implied msb), sign position
        exp_bits, man_bits = (8, 23) if dt.itemsize == 4 else (11, 52)
        bits = bytes([0x20, dt.itemsize * 8 - 1, 0])
        props = struct.pack('<HHBBBBI', 0, dt.itemsize * 8, man_bits, exp_bits, 0, man_bits, (1 << (exp_bits - 1)) - 1)
        return struct.pack('<B3sI', 0x11, bits, dt.itemsize) + props
    if dt.kind in 'iu' and dt.itemsize in (1, 2, 4, 8):
        bits = bytes([0x08 if dt.kind == 'i' else 0x00, 0, 0])
        return struct.pack('<B3sI', 0x10, bits, dt.itemsize) + struct.pack('<HH', 0, dt.itemsize * 8)
    raise NotImplementedError('dtype %s' % dt)


class _Writer(object):
    def __init__(self):
        self.buf = bytearray(96)                      # superblock goes here at the end

    def alloc(self, data):
        while len(self.buf) % 8:
            self.buf += b'\x00'
        off = len(self.buf)
        self.buf += data
        return off

    def dataset(self, arr):
        arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == '>':
            arr = arr.astype(arr.dtype.newbyteorder('<'))
        raw = self.alloc(arr.tobytes())
        rank = arr.ndim
        space = struct.pack('<BBB5x', 1, rank, 1) + b''.join(struct.pack('<Q', d) for d in arr.shape) * 2
        layout = struct.pack('<BBQQ', 3, 1, raw, arr.nbytes)
        fill = struct.pack('<BBBB', 2, 2, 0, 0)       # version 2, late allocation... fill value undefined-size 0: default
        msgs = [_msg(0x01, space), _msg(0x03, _dtype_msg(arr.dtype), flags=1), _msg(0x05, fill), _msg(0x08, layout)]
        return self.alloc(_object_header(msgs))

    def group(self, entries):
        """entries: [(name, object header address)] -> address of the group's object header."""
        entries = sorted(entries, key=lambda e: e[0].encode())
        # local heap: offset 0 = empty name, then the names; remaining space is one free block
        names = bytearray(8)
        offs = []
        for name, _ in entries:
            offs.append(len(names))
            names += _pad8(name.encode('utf-8') + b'\x00')
        free_off = len(names)
        names += struct.pack('<QQ', 1, 16)             # free block: next = H5HL_FREE_NULL (1), size of this block
        heap_data = self.alloc(bytes(names))
        heap = self.alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(names), free_off, heap_data))
        # one symbol-table node (leaf K = 16 -> up to 32 entries per node; larger groups get several nodes)
        K = 16
        nodes = []
        for s in range(0, max(len(entries), 1), 2 * K):
            chunk = list(zip(offs[s:s + 2 * K], entries[s:s + 2 * K]))
            body = b'SNOD' + struct.pack('<BBH', 1, 0, len(chunk))
            for noff, (name, oaddr) in chunk:
                body += struct.pack('<QQII16x', noff, oaddr, 0, 0)
            body += b'\x00' * ((2 * K - len(chunk)) * 40)
            nodes.append((self.alloc(body), chunk[-1][0] if chunk else 0))
        if len(nodes) > 2 * K:
            raise NotImplementedError('groups with more than %d members' % (4 * K * K))
        tree = b'TREE' + struct.pack('<BBHQQ', 0, 0, len(nodes), UNDEF, UNDEF) + struct.pack('<Q', 0)
        for naddr, last_name_off in nodes:
            tree += struct.pack('<QQ', naddr, last_name_off)
        tree += b'\x00' * ((2 * K - len(nodes)) * 16)
        btree = self.alloc(tree)
        hdr = self.alloc(_object_header([_msg(0x11, struct.pack('<QQ', btree, heap))]))
        return hdr, btree, heap

    def finish(self, root_hdr, btree, heap):
        eof = len(self.buf)
        sb = SIG + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, 16, 16, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
        sb += struct.pack('<QQII', 0, root_hdr, 1, 0) + struct.pack('<QQ', btree, heap)
        assert len(sb) == 96
        self.buf[0:96] = sb
        return bytes(self.buf)


def write_weights(path, weights):
    """Writes ``{layer: (Weight, Bias)}`` in the layout of NN.save_weights (NN.py:379-396): one group per layer with the
    datasets ``Weight`` and ``Bias``."""
    w = _Writer()
    top = []
    for layer, (W, b) in weights.items():
        members = [('Weight', w.dataset(np.asarray(W))), ('Bias', w.dataset(np.asarray(b)))]
        top.append((layer, w.group(members)[0]))
    root, btree, heap = w.group(top)
    data = w.finish(root, btree, heap)
    with open(path, 'wb') as fh:
        fh.write(data)
