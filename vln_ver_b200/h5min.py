"""Minimal HDF5 container support for the two files either side of the path, for
environments without h5py (this image): READ the view-feature file `VoxelFormer.get_image_feature` opens
(voxelformer.py:317-325: root-group datasets `<scan>_<vp>_i<e>_<deg>` of shape (1, 197, C), fp16 / fp32) and
WRITE the `getbev` export (HEAD:627-638: root-group datasets (C, Z, H, W), float64, gzip).

Scope -- the subset of the HDF5 File Format Specification (version 1.1 / 2.0 structures) that h5py / libhdf5
write with default settings (`libver='earliest'`):
  superblock v0 / v1 (at offset 0, 512, 1024, ...), base address;
  old-style groups: symbol-table message -> v1 B-tree (node type 0) -> SNOD symbol nodes + local heap;
  object header v1 with continuation blocks; dataspace v1 / v2; fixed-point and IEEE floating-point datatypes;
  data layout v1-v3: compact, contiguous, chunked (v1 B-tree, node type 1); filters: deflate, shuffle, fletcher32.
Not supported (raises H5Error): superblock v2+ / object header v2 / link messages (libver='latest'),
compound / string / variable-length types, external storage, datasets inside sub-groups on write.

The reader was checked against a file written by a real HDF5 library (scipy's MATLAB 7.3 sample,
tests/test_h5min.py); the writer is checked by round trips through that reader and structurally (node
sizes, name order, addresses) -- libhdf5 itself is not available here to cross-read its output.
`File` offers the slice of h5py.File's interface the reference uses: context manager, `f[key]` -> array-like
with numpy indexing, `create_dataset(key, shape, dtype=..., compression='gzip')`, `f[key][...] = data`, `close()`.
"""
import os
import struct
import zlib

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


# =========================================================================== reading
class _Reader:
    def __init__(self, data):
        self.d = data
        self.base = self._find_superblock()
        self._parse_superblock()

    def _find_superblock(self):
        off = 0
        while off + 8 <= len(self.d):
            if self.d[off:off + 8] == SIGNATURE:
                return off
            off = 512 if off == 0 else off * 2
        raise H5Error('not an HDF5 file (no superblock signature)')

    def u(self, off, n):
        return int.from_bytes(self.d[off:off + n], 'little')

    def _parse_superblock(self):
        p = self.base + 8
        version = self.d[p]
        if version not in (0, 1):
            raise H5Error(f'superblock version {version}: only the v0/v1 layout h5py writes by default is supported')
        self.O, self.L = self.d[p + 5], self.d[p + 6]
        if self.O != 8 or self.L != 8:
            raise H5Error('only 8-byte offsets / lengths are supported')
        self.leaf_k, self.internal_k = self.u(p + 8, 2), self.u(p + 10, 2)
        p += 16
        self.chunk_k = 32
        if version == 1:
            self.chunk_k = self.u(p, 2)
            p += 4
        base_addr = self.u(p, 8)
        if base_addr not in (0, self.base):
            raise H5Error('unexpected base address')
        self.eof = self.u(p + 16, 8)
        p += 32                                            # base, free-space, end-of-file, driver-info addresses
        # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
        self.root_header = self.u(p + 8, 8)

    def a(self, addr):
        """file address (relative to the base address) -> byte offset"""
        return self.base + addr

    # ------------------------------------------------------------------ object headers
    def messages(self, header_addr):
        p = self.a(header_addr)
        if self.d[p] != 1:
            raise H5Error(f'object header version {self.d[p]} (only v1; the file was written with libver="latest"?)')
        nmsg, size = self.u(p + 2, 2), self.u(p + 8, 4)
        out, blocks = [], [(p + 16, size)]
        while blocks and len(out) < nmsg:
            q, n = blocks.pop(0)
            end = q + n
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self.u(q, 2), self.u(q + 2, 2), self.d[q + 4]
                body = q + 8
                if mtype == 0x0010:                        # continuation: offset, length
                    blocks.append((self.a(self.u(body, 8)), self.u(body + 8, 8)))
                out.append((mtype, body, msize, flags))
                q = body + msize
        return out

    # ------------------------------------------------------------------ groups
    def _heap_string(self, heap_addr, off):
        p = self.a(heap_addr)
        if self.d[p:p + 4] != b'HEAP':
            raise H5Error('bad local heap signature')
        seg = self.a(self.u(p + 24, 8))
        end = self.d.index(b'\0', seg + off)
        return self.d[seg + off:end].decode()

    def _walk_group_btree(self, node_addr, heap_addr, out):
        p = self.a(node_addr)
        sig = self.d[p:p + 4]
        if sig == b'SNOD':
            n = self.u(p + 6, 2)
            q = p + 8
            for _ in range(n):
                out[self._heap_string(heap_addr, self.u(q, 8))] = self.u(q + 8, 8)
                q += 40
            return
        if sig != b'TREE' or self.d[p + 4] != 0:
            raise H5Error('bad group B-tree node')
        used = self.u(p + 6, 2)
        q = p + 24 + 8                                     # skip header and key 0
        for _ in range(used):
            self._walk_group_btree(self.u(q, 8), heap_addr, out)
            q += 16                                        # child pointer + next key

    def links(self, header_addr):
        for mtype, body, _, _ in self.messages(header_addr):
            if mtype == 0x0011:                            # symbol table message: B-tree address, heap address
                out = {}
                self._walk_group_btree(self.u(body, 8), self.u(body + 8, 8), out)
                return out
            if mtype in (0x0002, 0x0006):
                raise H5Error('new-style group (link messages): not supported')
        return None                                        # not a group

    # ------------------------------------------------------------------ datasets
    def _dtype(self, body):
        cls, version = self.d[body] & 15, self.d[body] >> 4
        bits0 = self.d[body + 1]
        size = self.u(body + 4, 4)
        order = '>' if bits0 & 1 else '<'
        if cls == 0:
            kind = 'i' if bits0 & 8 else 'u'
        elif cls == 1:
            kind = 'f'
            if size not in (2, 4, 8):
                raise H5Error(f'{size}-byte floating point type')
        else:
            raise H5Error(f'datatype class {cls} (only integers and IEEE floats are supported)')
        return np.dtype(f'{order}{kind}{size}')

    def _dataspace(self, body):
        version, rank, flags = self.d[body], self.d[body + 1], self.d[body + 2]
        p = body + (8 if version == 1 else 4)
        return tuple(self.u(p + 8 * i, 8) for i in range(rank))

    def _filters(self, body):
        version, n = self.d[body], self.d[body + 1]
        p = body + (8 if version == 1 else 2)
        out = []
        for _ in range(n):
            fid = self.u(p, 2)
            if version == 1 or fid >= 256:
                name_len = self.u(p + 2, 2)
                p += 2
            else:
                name_len = 0
            ncd = self.u(p + 4, 2)
            p += 6
            p += (name_len + 7) // 8 * 8 if version == 1 else name_len
            cd = [self.u(p + 4 * i, 4) for i in range(ncd)]
            p += 4 * ncd
            if version == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _chunks(self, node_addr, rank, out):
        p = self.a(node_addr)
        if self.d[p:p + 4] != b'TREE' or self.d[p + 4] != 1:
            raise H5Error('bad chunk B-tree node')
        level, used = self.d[p + 5], self.u(p + 6, 2)
        key = 8 + 8 * (rank + 1)
        q = p + 24
        for _ in range(used):
            size, mask = self.u(q, 4), self.u(q + 4, 4)
            offs = tuple(self.u(q + 8 + 8 * i, 8) for i in range(rank))
            child = self.u(q + key, 8)
            if level:
                self._chunks(child, rank, out)
            else:
                out.append((offs, size, mask, child))
            q += key + 8

    def dataset(self, header_addr):
        shape = dtype = layout = None
        filters = []
        for mtype, body, msize, _ in self.messages(header_addr):
            if mtype == 0x0001:
                shape = self._dataspace(body)
            elif mtype == 0x0003:
                dtype = self._dtype(body)
            elif mtype == 0x000B:
                filters = self._filters(body)
            elif mtype == 0x0008:
                layout = body
        if shape is None or dtype is None or layout is None:
            raise H5Error('not a dataset')
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if self.d[layout] in (1, 2):
            # layout v1 / v2 (HDF5 <= 1.6): dimensionality, class, 5 reserved, [address], 4-byte dims, [element size]
            ndim, cls = self.d[layout + 1], self.d[layout + 2]
            q = layout + 8
            addr = UNDEF
            if cls != 0:
                addr = self.u(q, 8)
                q += 8
            dims = tuple(self.u(q + 4 * i, 4) for i in range(ndim))
            q += 4 * ndim
            if cls == 1:
                if addr == UNDEF:
                    return np.zeros(shape, dtype)
                p = self.a(addr)
                return np.frombuffer(self.d[p:p + count * dtype.itemsize], dtype, count).reshape(shape).copy()
            if cls == 0:
                n = self.u(q, 4)
                return np.frombuffer(self.d[q + 4:q + 4 + n], dtype, count).reshape(shape).copy()
            return self._read_chunked(addr, ndim - 1, dims[:ndim - 1], shape, dtype, filters)
        if self.d[layout] != 3:
            raise H5Error(f'data layout message version {self.d[layout]} (only v1-v3)')
        cls = self.d[layout + 1]
        if cls == 0:                                       # compact
            n = self.u(layout + 2, 2)
            raw = self.d[layout + 4:layout + 4 + n]
            return np.frombuffer(raw, dtype, count).reshape(shape).copy()
        if cls == 1:                                       # contiguous
            addr, n = self.u(layout + 2, 8), self.u(layout + 10, 8)
            if addr == UNDEF:
                return np.zeros(shape, dtype)
            p = self.a(addr)
            return np.frombuffer(self.d[p:p + n], dtype, count).reshape(shape).copy()
        if cls != 2:
            raise H5Error(f'layout class {cls}')
        rank = self.d[layout + 2] - 1
        btree = self.u(layout + 3, 8)
        cdims = tuple(self.u(layout + 11 + 4 * i, 4) for i in range(rank))
        return self._read_chunked(btree, rank, cdims, shape, dtype, filters)

    def _read_chunked(self, btree, rank, cdims, shape, dtype, filters):
        out = np.zeros(shape, dtype)
        if btree == UNDEF:
            return out
        chunks = []
        self._chunks(btree, rank, chunks)
        for offs, size, mask, addr in chunks:
            p = self.a(addr)
            raw = bytes(self.d[p:p + size])
            for i, (fid, cd) in reversed(list(enumerate(filters))):
                if mask >> i & 1:
                    continue
                if fid == 3:
                    raw = raw[:-4]
                elif fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else dtype.itemsize
                    raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                else:
                    raise H5Error(f'filter {fid} is not supported')
            block = np.frombuffer(raw, dtype, int(np.prod(cdims))).reshape(cdims)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
            out[sel] = block[tuple(slice(0, s.stop - s.start) for s in sel)]
        return out


# =========================================================================== writing
def _pad8(b):
    return b + b'\0' * (-len(b) % 8)


def _msg(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack('<HHB3x', mtype, len(body), flags) + body


def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt.byteorder == '>':
        raise H5Error('big-endian data is not written')
    size = dt.itemsize
    if dt.kind == 'f':
        # class 1 v1; bits: little endian, pad zeros, mantissa normalisation 2 (msb implied); sign position
        spec = {2: (15, 10, 5, 0, 10, 15), 4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}.get(size)
        if spec is None:
            raise H5Error(f'float{8 * size}')
        sign, epos, esize, mpos, msize, bias = spec
        head = struct.pack('<BBBBI', 0x11, 0x20, sign, 0, size)
        return head + struct.pack('<HHBBBBI', 0, 8 * size, epos, esize, mpos, msize, bias)
    if dt.kind in 'iu':
        head = struct.pack('<BBBBI', 0x10, 0x08 if dt.kind == 'i' else 0x00, 0, 0, size)
        return head + struct.pack('<HH', 0, 8 * size)
    raise H5Error(f'dtype {dt} is not supported')


def _dataspace_message(shape):
    return struct.pack('<BBB5x', 1, len(shape), 0) + b''.join(struct.pack('<Q', int(s)) for s in shape)


class _Writer:
    """Appends objects to a byte buffer; every address is relative to base address 0."""
    MAX_LEAF_K, INTERNAL_K, CHUNK_K = 1024, 16, 32        # a SNOD holds up to 2048 links, the root B-tree 32 SNODs

    def __init__(self):
        self.buf = bytearray(b'\0' * 96)                  # superblock v0 is 96 bytes with 8-byte offsets
        self.LEAF_K = 4                                    # libhdf5's default; grown in finish() for large groups

    def alloc(self, data, align=8):
        self.buf += b'\0' * (-len(self.buf) % align)
        addr = len(self.buf)
        self.buf += data
        return addr

    def object_header(self, messages):
        body = b''.join(messages)
        head = struct.pack('<BBHII', 1, 0, len(messages), 1, len(body)) + b'\0' * 4
        return self.alloc(head + body)

    def dataset(self, arr, gzip=None):
        arr = np.ascontiguousarray(arr)
        shape = arr.shape if arr.ndim else (1,)
        msgs = [_msg(0x0001, _dataspace_message(shape)), _msg(0x0003, _dtype_message(arr.dtype), flags=1)]
        raw = arr.tobytes()
        if gzip is None:
            msgs.append(_msg(0x0005, struct.pack('<BBBB', 2, 2, 0, 0)))                  # fill value v2: late, undefined
            addr = self.alloc(raw) if raw else UNDEF
            msgs.append(_msg(0x0008, struct.pack('<BBQQ', 3, 1, addr, len(raw))))
        else:
            rank = len(shape)
            msgs.append(_msg(0x0005, struct.pack('<BBBB', 2, 3, 0, 0)))                  # incremental allocation
            msgs.append(_msg(0x000B, struct.pack('<BB6x', 1, 1) + struct.pack('<HHHH', 1, 8, 1, 1) + b'deflate\0'
                             + struct.pack('<II', int(gzip), 0)))
            comp = zlib.compress(raw, int(gzip))
            caddr = self.alloc(comp)
            key = 8 + 8 * (rank + 1)
            node = bytearray(b'TREE' + struct.pack('<BBHQQ', 1, 0, 1, UNDEF, UNDEF))
            node += struct.pack('<II', len(comp), 0) + b'\0' * (8 * (rank + 1)) + struct.pack('<Q', caddr)
            # final key: one chunk past the end along the slowest dimension
            node += struct.pack('<II', 0, 0) + struct.pack('<Q', shape[0]) + b'\0' * (8 * rank)
            node += b'\0' * (24 + 2 * self.CHUNK_K * (key + 8) + key - len(node))         # full-size node
            baddr = self.alloc(bytes(node))
            dims = b''.join(struct.pack('<I', int(s)) for s in shape) + struct.pack('<I', arr.dtype.itemsize)
            msgs.append(_msg(0x0008, struct.pack('<BBBQ', 3, 2, rank + 1, baddr) + dims))
        return self.object_header(msgs)

    def finish(self, links):
        """links: name -> object header address.  Writes heap, symbol nodes, B-tree, root group, superblock."""
        names = sorted(links, key=lambda s: s.encode())
        while 2 * self.LEAF_K < len(names) and self.LEAF_K < self.MAX_LEAF_K:
            self.LEAF_K *= 2                               # the K values travel in the superblock
        per = 2 * self.LEAF_K
        if len(names) > per * 2 * self.INTERNAL_K:
            raise H5Error(f'{len(names)} datasets: more than this writer\'s single-level root B-tree holds')
        heap = bytearray(b'\0' * 8)                        # offset 0: the empty string (B-tree key 0)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode() + b'\0')
        free_off = len(heap)
        heap += struct.pack('<QQ', 1, 32) + b'\0' * 16     # one free block: next = 1 (none), size 32
        seg_addr = self.alloc(bytes(heap))
        heap_addr = self.alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap), free_off, seg_addr))
        groups = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        snods, keys = [], [0]
        for g in groups:
            node = bytearray(b'SNOD' + struct.pack('<BBH', 1, 0, len(g)))
            for n in g:
                node += struct.pack('<QQII16x', offs[n], links[n], 0, 0)
            node += b'\0' * (8 + per * 40 - len(node))     # full-size node
            snods.append(self.alloc(bytes(node)))
            keys.append(offs[g[-1]] if g else 0)
        tree = bytearray(b'TREE' + struct.pack('<BBHQQ', 0, 0, len(snods), UNDEF, UNDEF) + struct.pack('<Q', keys[0]))
        for addr, k in zip(snods, keys[1:]):
            tree += struct.pack('<QQ', addr, k)
        tree += b'\0' * (24 + 8 + 2 * self.INTERNAL_K * 16 - len(tree))
        tree_addr = self.alloc(bytes(tree))
        root = self.object_header([_msg(0x0011, struct.pack('<QQ', tree_addr, heap_addr))])
        eof = len(self.buf)
        sb = SIGNATURE + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
        sb += struct.pack('<QQII', 0, root, 1, 0) + struct.pack('<QQ', tree_addr, heap_addr)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


# =========================================================================== the h5py-like face
class _PendingDataset:
    """What create_dataset returns / f[key] yields before the file is written: numpy-indexable and assignable."""

    def __init__(self, shape, dtype, gzip):
        self.array = np.zeros(shape, dtype)
        self.gzip = gzip
        self.shape, self.dtype = self.array.shape, self.array.dtype

    def __getitem__(self, idx):
        return self.array[idx]

    def __setitem__(self, idx, value):
        self.array[idx] = value

    def __array__(self, dtype=None, copy=None):
        return self.array if dtype is None else self.array.astype(dtype)


class File:
    """`h5py.File(path, mode)` stand-in for modes 'r', 'w', 'a' over root-group datasets (see module docstring).
    'a' on an existing file reads every dataset and rewrites the file on close (the files of this path hold one
    dataset per panorama; appends are rare and small next to the run that produces them)."""

    def __init__(self, path, mode='r'):
        if mode not in ('r', 'w', 'a'):
            raise ValueError(f'mode {mode!r}')
        self.path, self.mode = path, mode
        self._reader, self._links, self._new = None, {}, {}
        if mode == 'r' or (mode == 'a' and os.path.exists(path)):
            with open(path, 'rb') as f:
                self._reader = _Reader(f.read())
            self._links = self._reader.links(self._reader.root_header) or {}
        self._closed = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def keys(self):
        return sorted(set(self._links) | set(self._new))

    def __contains__(self, key):
        return key in self._links or key in self._new

    def __len__(self):
        return len(self.keys())

    def __getitem__(self, key):
        if key in self._new:
            return self._new[key]
        if key not in self._links:
            raise KeyError(key)
        addr = self._links[key]
        if self._reader.links(addr) is not None:
            raise H5Error(f'{key!r} is a group: only root-group datasets are supported')
        return self._reader.dataset(addr)

    def create_dataset(self, key, shape=None, dtype=None, data=None, compression=None, compression_opts=None):
        if self.mode == 'r':
            raise H5Error('file is open read-only')
        if key in self:
            raise ValueError(f'unable to create dataset (name already exists): {key}')
        if compression not in (None, 'gzip'):
            raise H5Error(f'compression {compression!r}: only gzip')
        if data is not None:
            data = np.asarray(data)
            shape, dtype = data.shape if shape is None else shape, data.dtype if dtype is None else dtype
        if dtype in ('float', float):
            dtype = np.float64                             # h5py: dtype='float' is float64 (HEAD:635)
        gzip = (4 if compression_opts is None else compression_opts) if compression == 'gzip' else None
        ds = _PendingDataset(shape, np.dtype(dtype or np.float32), gzip)
        if data is not None:
            ds[...] = data.reshape(ds.shape)
        self._new[key] = ds
        return ds

    def close(self):
        if self._closed:
            return
        self._closed = True
        if self.mode == 'r' or (not self._new and self._reader is not None):
            return
        w = _Writer()
        links = {}
        for key, addr in self._links.items():             # carry the existing datasets over (mode 'a')
            links[key] = w.dataset(self._reader.dataset(addr))
        for key, ds in self._new.items():
            links[key] = w.dataset(ds.array, ds.gzip)
        data = w.finish(links)
        tmp = self.path + '.tmp'
        with open(tmp, 'wb') as f:
            f.write(data)
        os.replace(tmp, self.path)
