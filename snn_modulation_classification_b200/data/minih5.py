"""Install-free reader (and a small writer) for the HDF5 subset the RadioML per-(class, SNR) files use.

``data/load_radio_ml.py`` reads ``class<c>_snr<z>.hdf5`` files, each holding one dataset ``X`` written by
``h5py.File(...).create_dataset('X', data=array)`` (reference data/load_radio_ml.py:46-53): HDF5's *earliest* file format --
superblock version 0, version-1 object headers, the root group as a symbol table (v1 B-tree + local heap + ``SNOD`` nodes),
dataspace v1/v2, IEEE / integer little-endian datatypes, data layout v3 (contiguous or compact).  That subset of the HDF5 File
Format Specification (version 1.1 / 2.0, sections III.A-III.D, IV.A.1-IV.A.2) is what this module implements, so that the loader
works where h5py is not installed; h5py is preferred when importable (``load_radio_ml._h5_module``).

API = the part of h5py the loader touches: ``File(path, 'r'|'w')``, ``f['X'][:]`` / ``f['X'][a:b]``, ``.shape``, ``.dtype``,
``f.create_dataset(name, data=)``, ``f.close()``, ``name in f``, ``f.keys()``.  Chunked / compressed datasets, new-style groups
(object header v2, superblock >= 2) raise ``NotImplementedError`` naming what was found -- nothing is guessed.

Not validated against libhdf5 in this repository's build container (neither h5py nor the HDF5 tools are installed there): the
writer and the reader were written independently from the specification and are tested against each other and against
hand-checked byte offsets (tests/test_host_cpu_loader.py).
"""
import struct

import numpy as np

SIG = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(n):
    return (n + 7) // 8 * 8


# ---------------------------------------------------------------------------------------------------------------------
# datatype message (IV.A.2.d)
# ---------------------------------------------------------------------------------------------------------------------
def _encode_dtype(dt):
    dt = np.dtype(dt)
    if dt.byteorder == '>':
        raise NotImplementedError('big-endian data')
    if dt.kind == 'f' and dt.itemsize in (4, 8):
        bits = dt.itemsize * 8
        exp_size, mant = (8, 23) if bits == 32 else (11, 52)
        head = struct.pack('<BBBBI', 0x11, 0x20, bits - 1, 0, dt.itemsize)        # version 1 | class 1; msb-implied mantissa; sign bit
        return head + struct.pack('<HHBBBBI', 0, bits, mant, exp_size, 0, mant, (1 << (exp_size - 1)) - 1)
    if dt.kind in 'iu':
        head = struct.pack('<BBBBI', 0x10, 0x08 if dt.kind == 'i' else 0x00, 0, 0, dt.itemsize)
        return head + struct.pack('<HH', 0, dt.itemsize * 8)
    raise NotImplementedError('dtype %r' % dt)


def _decode_dtype(b):
    cls_ver, f0, f1, _f2, size = struct.unpack_from('<BBBBI', b, 0)
    cls = cls_ver & 0x0F
    if f0 & 1:
        raise NotImplementedError('big-endian data')
    if cls == 1:
        if size not in (4, 8):
            raise NotImplementedError('%d-byte floating point' % size)
        return np.dtype('<f%d' % size)
    if cls == 0:
        return np.dtype('<%s%d' % ('i' if f0 & 0x08 else 'u', size))
    raise NotImplementedError('HDF5 datatype class %d' % cls)


# ---------------------------------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------------------------------
class Dataset:
    def __init__(self, f, shape, dtype, layout):
        self._f, self.shape, self.dtype, self._layout = f, tuple(shape), dtype, layout

    def _read_all(self):
        n = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        kind, a, b = self._layout
        if kind == 'compact':
            raw = a
        else:
            if a == UNDEF:
                return np.zeros(self.shape, self.dtype)             # never written: fill value
            self._f._fh.seek(a)
            raw = self._f._fh.read(n * self.dtype.itemsize)
        return np.frombuffer(raw, dtype=self.dtype, count=n).reshape(self.shape).copy()

    def __getitem__(self, key):
        kind, addr, _ = self._layout
        if kind == 'contiguous' and addr != UNDEF and isinstance(key, slice) and self.shape and (key.step in (None, 1)):
            start, stop, _ = key.indices(self.shape[0])              # leading-axis slice: read only those rows
            row = int(np.prod(self.shape[1:], dtype=np.int64)) * self.dtype.itemsize
            self._f._fh.seek(addr + start * row)
            raw = self._f._fh.read(max(0, stop - start) * row)
            return np.frombuffer(raw, dtype=self.dtype).reshape((max(0, stop - start),) + self.shape[1:]).copy()
        return self._read_all()[key]

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        a = self._read_all()
        return a.astype(dtype) if dtype is not None else a


class File:
    def __init__(self, path, mode='r'):
        if mode not in ('r', 'w'):
            raise ValueError("minih5.File supports modes 'r' and 'w'")
        self.mode, self._path = mode, path
        self._pending = {}
        if mode == 'r':
            self._fh = open(path, 'rb')
            self._links = self._read_root()
        else:
            self._fh = None

    # -- context / dict protocol
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def keys(self):
        return list(self._links if self.mode == 'r' else self._pending)

    def __contains__(self, name):
        return name in (self._links if self.mode == 'r' else self._pending)

    def __getitem__(self, name):
        if self.mode != 'r':
            raise IOError('file opened for writing')
        if name not in self._links:
            raise KeyError(name)
        return self._read_dataset(self._links[name])

    # -- low-level reads
    def _at(self, off, n):
        self._fh.seek(off)
        b = self._fh.read(n)
        if len(b) != n:
            raise IOError('truncated HDF5 file %s' % self._path)
        return b

    def _read_root(self):
        if self._at(0, 8) != SIG:
            raise IOError('%s is not an HDF5 file (no signature at offset 0)' % self._path)
        ver = self._at(8, 1)[0]
        if ver not in (0, 1):
            raise NotImplementedError('HDF5 superblock version %d (written with a newer libver): only the earliest format '
                                      '(versions 0 / 1, h5py default) is implemented' % ver)
        so, sl = self._at(13, 1)[0], self._at(14, 1)[0]
        if (so, sl) != (8, 8):
            raise NotImplementedError('%d-byte offsets / %d-byte lengths' % (so, sl))
        p = 24 + (4 if ver == 1 else 0)                              # v1 adds indexed-storage K + reserved
        base = struct.unpack('<Q', self._at(p, 8))[0]
        if base != 0:
            raise NotImplementedError('non-zero base address')
        root = p + 32                                                # root group symbol-table entry
        _name_off, ohdr, cache, _res = struct.unpack('<QQII', self._at(root, 24))
        btree = heap = None
        if cache == 1:
            btree, heap = struct.unpack('<QQ', self._at(root + 24, 16))
        else:
            for mtype, data in self._messages(ohdr):
                if mtype == 0x0011:
                    btree, heap = struct.unpack_from('<QQ', data, 0)
        if btree is None:
            raise NotImplementedError('root group without a symbol table (new-style group)')
        return self._walk_group(btree, heap)

    def _heap_name(self, heap_data, off):
        end = heap_data.index(b'\0', off)
        return heap_data[off:end].decode('utf-8')

    def _walk_group(self, btree, heap):
        h = self._at(heap, 32)
        if h[:4] != b'HEAP':
            raise IOError('bad local heap signature')
        seg_size, _free, seg_addr = struct.unpack_from('<QQQ', h, 8)
        heap_data = self._at(seg_addr, seg_size)
        links = {}

        def node(addr):
            hd = self._at(addr, 24)
            if hd[:4] == b'SNOD':
                nsym = struct.unpack_from('<H', hd, 6)[0]
                ent = self._at(addr + 8, 40 * nsym)
                for i in range(nsym):
                    name_off, oh = struct.unpack_from('<QQ', ent, 40 * i)
                    links[self._heap_name(heap_data, name_off)] = oh
                return
            if hd[:4] != b'TREE' or hd[4] != 0:
                raise IOError('bad group B-tree node')
            used = struct.unpack_from('<H', hd, 6)[0]
            body = self._at(addr + 24, (2 * used + 1) * 8)
            for i in range(used):
                node(struct.unpack_from('<Q', body, 8 + 16 * i)[0])  # key0 child0 key1 child1 ...

        node(btree)
        return links

    def _messages(self, addr):
        """(type, data) of every message of a version-1 object header, continuation blocks included (IV.A.1.a)."""
        hd = self._at(addr, 16)
        if hd[:4] == b'OHDR':
            raise NotImplementedError('version-2 object header (file written with libver >= 1.8 "latest")')
        if hd[0] != 1:
            raise NotImplementedError('object header version %d' % hd[0])
        nmsg, _refs, size = struct.unpack_from('<HII', hd, 2)
        blocks, out = [(addr + 16, size)], []
        while blocks and len(out) < nmsg:
            off, left = blocks.pop(0)
            while left >= 8 and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack('<HHB', self._at(off, 5))
                data = self._at(off + 8, msize)
                if mtype == 0x0010:
                    blocks.append(struct.unpack('<QQ', data[:16]))
                out.append((mtype, data))
                off += 8 + msize
                left -= 8 + msize
        return out

    def _read_dataset(self, ohdr):
        shape = dtype = layout = None
        for mtype, data in self._messages(ohdr):
            if mtype == 0x0001:                                      # dataspace
                ver, rank, flags = data[0], data[1], data[2]
                p = 8 if ver == 1 else 4
                shape = struct.unpack_from('<%dQ' % rank, data, p) if rank else ()
            elif mtype == 0x0003:
                dtype = _decode_dtype(data)
            elif mtype == 0x0008:                                    # data layout
                if data[0] != 3:
                    raise NotImplementedError('data layout message version %d' % data[0])
                cls = data[1]
                if cls == 1:
                    layout = ('contiguous',) + struct.unpack_from('<QQ', data, 2)
                elif cls == 0:
                    n = struct.unpack_from('<H', data, 2)[0]
                    layout = ('compact', bytes(data[4:4 + n]), n)
                else:
                    raise NotImplementedError('chunked dataset (layout class %d): re-write it contiguous, or install h5py' % cls)
            elif mtype == 0x000B:
                raise NotImplementedError('filtered (compressed) dataset: install h5py')
        if shape is None or dtype is None or layout is None:
            raise IOError('object at %d is not a simple dataset' % ohdr)
        return Dataset(self, shape, dtype, layout)

    # -- writer
    def create_dataset(self, name, data=None, **_ignored):
        if self.mode != 'w':
            raise IOError('file opened read-only')
        if len(self._pending) >= 8:
            raise NotImplementedError('more than 8 objects in the root group (one symbol-table node)')
        self._pending[name] = np.ascontiguousarray(data)

    def close(self):
        if self.mode == 'r':
            if self._fh:
                self._fh.close()
                self._fh = None
            return
        if self._pending is None:
            return
        names = sorted(self._pending)                                # symbol-table entries are sorted by name
        # layout of the file: superblock | root object header | B-tree node | local heap (+ data) | SNOD | per dataset: header, data
        heap_data = bytearray(8)                                     # offset 0: the empty name
        name_off = {}
        for n in names:
            name_off[n] = len(heap_data)
            heap_data += n.encode('utf-8') + b'\0'
            heap_data += b'\0' * (_pad8(len(heap_data)) - len(heap_data))
        heap_seg = max(_pad8(len(heap_data)) + 8, 88)
        heap_data += b'\0' * (heap_seg - len(heap_data))
        k_leaf, k_int = 4, 16
        off_root = 96
        root_msgs = struct.pack('<HHBBBB', 0x0011, 16, 0, 0, 0, 0) + b'\0' * 16     # patched below
        off_btree = off_root + 16 + len(root_msgs)
        btree_size = 24 + (2 * k_int + 1) * 8 + 2 * k_int * 8
        off_heap = off_btree + btree_size
        off_heap_data = off_heap + 32
        off_snod = off_heap_data + heap_seg
        snod_size = 8 + 2 * k_leaf * 40
        pos = off_snod + snod_size
        hdrs = {}
        for n in names:
            a = self._pending[n]
            space = struct.pack('<BBBB4x', 1, a.ndim, 0, 0) + struct.pack('<%dQ' % a.ndim, *a.shape)
            dtype = _encode_dtype(a.dtype)
            msgs = b''
            layout = struct.pack('<BBQQ', 3, 1, 0, a.nbytes)     # version 3, contiguous; the address is patched below
            for mtype, body in ((0x0001, space), (0x0003, dtype), (0x0008, layout)):
                body += b'\0' * (_pad8(len(body)) - len(body))
                msgs += struct.pack('<HHBBBB', mtype, len(body), 0, 0, 0, 0) + body
            off_hdr = pos
            off_data = _pad8(off_hdr + 16 + len(msgs))
            # patch the contiguous-layout address (last message: 8-byte header, then version, class, address)
            lay_at = len(msgs) - _pad8(18) + 2
            msgs = msgs[:lay_at] + struct.pack('<Q', off_data) + msgs[lay_at + 8:]
            hdrs[n] = (off_hdr, msgs, off_data)
            pos = _pad8(off_data + a.nbytes)
        eof = pos
        out = bytearray(eof)
        # superblock, version 0 (III.A)
        out[0:8] = SIG
        out[8:16] = struct.pack('<BBBBBBBB', 0, 0, 0, 0, 0, 8, 8, 0)
        out[16:24] = struct.pack('<HHI', k_leaf, k_int, 0)
        out[24:56] = struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
        out[56:96] = struct.pack('<QQII', 0, off_root, 1, 0) + struct.pack('<QQ', off_btree, off_heap)
        # root group object header (version 1) with its symbol-table message
        root_msgs = struct.pack('<HHBBBB', 0x0011, 16, 0, 0, 0, 0) + struct.pack('<QQ', off_btree, off_heap)
        out[off_root:off_root + 16] = struct.pack('<BBHII4x', 1, 0, 1, 1, len(root_msgs))
        out[off_root + 16:off_root + 16 + len(root_msgs)] = root_msgs
        # B-tree node (type 0 = group, level 0): one child, keys = heap offsets of the smallest / largest name below it
        out[off_btree:off_btree + 24] = b'TREE' + struct.pack('<BBHQQ', 0, 0, 1, UNDEF, UNDEF)
        out[off_btree + 24:off_btree + 48] = struct.pack('<QQQ', 0, off_snod, name_off[names[-1]] if names else 0)
        # local heap
        out[off_heap:off_heap + 32] = b'HEAP' + struct.pack('<B3xQQQ', 0, heap_seg, UNDEF, off_heap_data)
        out[off_heap_data:off_heap_data + heap_seg] = heap_data
        # symbol-table node
        out[off_snod:off_snod + 8] = b'SNOD' + struct.pack('<BBH', 1, 0, len(names))
        for i, n in enumerate(names):
            out[off_snod + 8 + 40 * i:off_snod + 48 + 40 * i] = struct.pack('<QQII16x', name_off[n], hdrs[n][0], 0, 0)
        # datasets
        for n in names:
            off_hdr, msgs, off_data = hdrs[n]
            out[off_hdr:off_hdr + 16] = struct.pack('<BBHII4x', 1, 0, 3, 1, len(msgs))
            out[off_hdr + 16:off_hdr + 16 + len(msgs)] = msgs
            a = self._pending[n]
            out[off_data:off_data + a.nbytes] = a.tobytes()
        with open(self._path, 'wb') as fh:
            fh.write(out)
        self._pending = None
