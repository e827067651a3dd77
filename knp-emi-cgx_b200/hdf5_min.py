"""Minimal HDF5 reader (and fixture writer) for mesh ingest -- host side, setup time, numpy only.

Replaces the HDF5 half of ``dfx.io.XDMFFile(...).read_mesh / read_meshtags``
(src/CGx/utils/mixed_dim_problem.py:634-681): neither libhdf5 nor h5py exists in this image, so the subset of the
published file format (HDF5 File Format Specification 3.0) that DOLFINx / meshio / emimesh files use is parsed directly:

  superblock        versions 0-3 (a user block of 512 * 2^k bytes is skipped)
  groups            symbol tables (B-tree v1 + local heap + SNOD) and compact link messages (object header v2)
  object headers    versions 1 and 2, continuation blocks
  datasets          fixed-point and IEEE floating-point types, either byte order; compact, contiguous and chunked
                    (layout versions 1-3, B-tree v1 chunk index) storage; deflate / shuffle / fletcher32 filters

Dense link storage (fractal heaps: groups with more than 8 links written with libver=latest) and layout version 4 chunk
indices raise ``Hdf5FormatError`` naming the feature.  Contiguous datasets are returned as ``numpy.memmap`` views, so a
100 M-cell topology costs no copy.  If h5py is importable it is used instead (``open_file``).

``write_file`` writes the earliest-format layout (superblock 0, symbol-table groups, contiguous datasets) -- what
libhdf5 writes by default.  It exists for the test fixtures and for exporting generated meshes; the reader is pinned
against a file written by libhdf5 itself (scipy's MATLAB 7.3 test file, ``tests/test_mesh_ingest.py``).
"""
import struct
import zlib
import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5FormatError(RuntimeError):
    pass


class _Reader:
    def __init__(self, path):
        self.path = path
        self.buf = np.memmap(path, dtype=np.uint8, mode="r")
        self.base = 0
        self._superblock()

    # ---------------------------------------------------------------- primitives
    def bytes(self, off, n):
        if off + n > self.buf.size:
            raise Hdf5FormatError(f"{self.path}: read past the end of the file (offset {off}, {n} bytes)")
        return self.buf[off:off + n].tobytes()

    def uint(self, off, n):
        return int.from_bytes(self.bytes(off, n), "little")

    def addr(self, off):
        v = self.uint(off, self.so)
        return None if v == (1 << (8 * self.so)) - 1 else v + self.base

    # ---------------------------------------------------------------- superblock
    def _superblock(self):
        off = 0
        while True:
            if off + 8 > self.buf.size:
                raise Hdf5FormatError(f"{self.path}: not an HDF5 file (no superblock signature)")
            if self.bytes(off, 8) == SIG:
                break
            off = 512 if off == 0 else off * 2
        ver = self.uint(off + 8, 1)
        if ver in (0, 1):
            self.so, self.sl = self.uint(off + 13, 1), self.uint(off + 14, 1)
            p = off + 24 + (4 if ver == 1 else 0)
            base = self.uint(p, self.so)
            # root group symbol table entry after base / free-space / end-of-file / driver addresses
            ent = p + 4 * self.so
            for cand in (base, off):          # files with a user block store addresses relative to the superblock
                self.base = cand
                root = self.addr(ent + self.so)
                if root is not None and root + 16 <= self.buf.size and self.uint(root, 1) == 1:
                    break
            else:
                raise Hdf5FormatError(f"{self.path}: root group object header not found")
            self.root = root
        elif ver in (2, 3):
            self.so, self.sl = self.uint(off + 9, 1), self.uint(off + 10, 1)
            base = self.uint(off + 12, self.so)
            for cand in (base, off):
                self.base = cand
                root = self.addr(off + 12 + 3 * self.so)
                if root is not None and root + 4 <= self.buf.size and self.bytes(root, 4) == b"OHDR":
                    break
            else:
                raise Hdf5FormatError(f"{self.path}: root group object header not found")
            self.root = root
        else:
            raise Hdf5FormatError(f"{self.path}: superblock version {ver} is not supported")

    # ---------------------------------------------------------------- object headers
    def messages(self, oh):
        """[(type, flags, offset of the message body, size)] of the object header at `oh`, continuation blocks followed."""
        out = []
        if self.bytes(oh, 4) == b"OHDR":
            flags = self.uint(oh + 5, 1)
            p = oh + 6 + (16 if flags & 0x20 else 0) + (4 if flags & 0x10 else 0)
            w = 1 << (flags & 3)
            size0 = self.uint(p, w)
            p += w
            blocks = [(p, p + size0)]
            track = bool(flags & 0x04)
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 <= end:
                    mtype, msize, mflags = self.uint(p, 1), self.uint(p + 1, 2), self.uint(p + 3, 1)
                    p += 4 + (2 if track else 0)
                    if mtype == 0x10:
                        a, ln = self.addr(p), self.uint(p + self.so, self.sl)
                        if self.bytes(a, 4) != b"OCHK":
                            raise Hdf5FormatError(f"{self.path}: bad object header continuation block")
                        blocks.append((a + 4, a + ln - 4))
                    elif mtype != 0:
                        out.append((mtype, mflags, p, msize))
                    p += msize
        else:
            ver = self.uint(oh, 1)
            if ver != 1:
                raise Hdf5FormatError(f"{self.path}: object header version {ver} at {oh} is not supported")
            nmsg, hsize = self.uint(oh + 2, 2), self.uint(oh + 8, 4)
            blocks = [(oh + 16, oh + 16 + hsize)]
            while blocks and len(out) < nmsg + 64:
                p, end = blocks.pop(0)
                while p + 8 <= end:
                    mtype, msize, mflags = self.uint(p, 2), self.uint(p + 2, 2), self.uint(p + 4, 1)
                    p += 8
                    if mtype == 0x10:
                        a, ln = self.addr(p), self.uint(p + self.so, self.sl)
                        blocks.append((a, a + ln))
                    elif mtype != 0:
                        out.append((mtype, mflags, p, msize))
                    p += msize                                  # v1 sizes include the alignment padding
        return out

    # ---------------------------------------------------------------- groups
    def links(self, oh):
        """{name: object header address} of the group at `oh`."""
        out = {}
        for mtype, _f, p, _n in self.messages(oh):
            if mtype == 0x11:                                   # symbol table: B-tree v1 + local heap
                btree, heap = self.addr(p), self.addr(p + self.so)
                if self.bytes(heap, 4) != b"HEAP":
                    raise Hdf5FormatError(f"{self.path}: bad local heap")
                hdata = self.addr(heap + 8 + 2 * self.sl)
                self._group_node(btree, hdata, out)
            elif mtype == 0x06:                                 # link message
                flags = self.uint(p + 1, 1)
                q = p + 2
                ltype = 0
                if flags & 0x08:
                    ltype = self.uint(q, 1)
                    q += 1
                if flags & 0x04:
                    q += 8
                if flags & 0x10:
                    q += 1
                w = 1 << (flags & 3)
                ln = self.uint(q, w)
                q += w
                name = self.bytes(q, ln).decode("utf-8")
                q += ln
                if ltype == 0:
                    out[name] = self.addr(q)
            elif mtype == 0x02:                                 # link info: dense storage?
                flags = self.uint(p + 1, 1)
                q = p + 2 + (8 if flags & 1 else 0)
                if self.addr(q) is not None:
                    raise Hdf5FormatError(f"{self.path}: dense link storage (fractal heap) is not supported by the built-in "
                                          "reader; rewrite the file with the default (earliest) libver or h5repack")
        return out

    def _group_node(self, node, hdata, out):
        sig = self.bytes(node, 4)
        if sig == b"TREE":
            level, used = self.uint(node + 5, 1), self.uint(node + 6, 2)
            p = node + 8 + 2 * self.so
            for i in range(used):
                child = self.addr(p + self.sl + i * (self.sl + self.so))
                self._group_node(child, hdata, out)
        elif sig == b"SNOD":
            n = self.uint(node + 6, 2)
            p = node + 8
            for i in range(n):
                e = p + i * (2 * self.so + 24)
                noff, oh = self.uint(e, self.so), self.addr(e + self.so)
                end = noff
                while self.buf[hdata + end] != 0:
                    end += 1
                out[self.bytes(hdata + noff, end - noff).decode("utf-8")] = oh
        else:
            raise Hdf5FormatError(f"{self.path}: bad group node signature {sig!r}")

    def resolve(self, path):
        oh = self.root
        for part in [s for s in path.split("/") if s]:
            ln = self.links(oh)
            if part not in ln:
                raise KeyError(f"{self.path}: no object {path!r} (missing {part!r}; have {sorted(ln)})")
            oh = ln[part]
        return oh

    # ---------------------------------------------------------------- datasets
    def dataset(self, path):
        oh = self.resolve(path)
        shape = dtype = layout = None
        filters = []
        for mtype, _f, p, n in self.messages(oh):
            if mtype == 0x01:
                ver, rank = self.uint(p, 1), self.uint(p + 1, 1)
                q = p + (8 if ver == 1 else 4)
                shape = tuple(self.uint(q + i * self.sl, self.sl) for i in range(rank))
            elif mtype == 0x03:
                cv, b0, size = self.uint(p, 1), self.uint(p + 1, 1), self.uint(p + 4, 4)
                cls = cv & 15
                order = ">" if b0 & 1 else "<"
                if cls == 0:
                    dtype = np.dtype(f"{order}{'i' if b0 & 8 else 'u'}{size}")
                elif cls == 1:
                    dtype = np.dtype(f"{order}f{size}")
                else:
                    raise Hdf5FormatError(f"{self.path}:{path}: datatype class {cls} is not supported (integers / floats only)")
            elif mtype == 0x08:
                layout = self._layout(p)
            elif mtype == 0x0B:
                filters = self._filters(p)
        if shape is None or dtype is None or layout is None:
            raise Hdf5FormatError(f"{self.path}:{path} is not a dataset")
        count = int(np.prod(shape)) if shape else 1
        kind = layout[0]
        if kind == "compact":
            return np.frombuffer(self.bytes(layout[1], count * dtype.itemsize), dtype).reshape(shape)
        if kind == "contiguous":
            if layout[1] is None or count == 0:
                return np.zeros(shape, dtype)
            return np.ndarray(shape, dtype, buffer=self.buf, offset=layout[1])
        return self._chunked(layout, shape, dtype, filters)

    def _layout(self, p):
        ver = self.uint(p, 1)
        if ver == 3:
            cls = self.uint(p + 1, 1)
            if cls == 0:
                return ("compact", p + 4)
            if cls == 1:
                return ("contiguous", self.addr(p + 2))
            if cls == 2:
                nd = self.uint(p + 2, 1)
                bt = self.addr(p + 3)
                dims = [self.uint(p + 3 + self.so + 4 * i, 4) for i in range(nd)]
                return ("chunked", bt, dims)
        elif ver in (1, 2):
            nd, cls = self.uint(p + 1, 1), self.uint(p + 2, 1)
            q = p + 8
            a = None
            if cls != 0:
                a = self.addr(q)
                q += self.so
            dims = [self.uint(q + 4 * i, 4) for i in range(nd)]
            q += 4 * nd
            if cls == 1:
                return ("contiguous", a)
            if cls == 2:
                return ("chunked", a, dims + [self.uint(q, 4)])
            return ("compact", q + 4)
        raise Hdf5FormatError(f"{self.path}: data layout version {ver} is not supported (libver=latest chunk indices); "
                              "rewrite the dataset contiguous (h5repack -l CONTI)")

    def _filters(self, p):
        ver, nf = self.uint(p, 1), self.uint(p + 1, 1)
        q = p + (8 if ver == 1 else 2)
        out = []
        for _ in range(nf):
            fid = self.uint(q, 2)
            q += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = self.uint(q, 2)
                q += 2
            ncd = self.uint(q + 2, 2)
            q += 4
            q += (nlen + 7) & ~7 if ver == 1 else nlen
            cd = [self.uint(q + 4 * i, 4) for i in range(ncd)]
            q += 4 * ncd + (4 if ver == 1 and ncd & 1 else 0)
            out.append((fid, cd))
        return out

    def _chunked(self, layout, shape, dtype, filters):
        _k, bt, dims = layout
        cshape = tuple(dims[:-1])
        out = np.zeros(shape, dtype)
        if bt is None:
            return out
        nd = len(cshape)

        def walk(node):
            if self.bytes(node, 4) != b"TREE":
                raise Hdf5FormatError(f"{self.path}: bad chunk index node")
            level, used = self.uint(node + 5, 1), self.uint(node + 6, 2)
            ksz = 8 + 8 * (nd + 1)
            p = node + 8 + 2 * self.so
            for i in range(used):
                k = p + i * (ksz + self.so)
                child = self.addr(k + ksz)
                if level > 0:
                    walk(child)
                    continue
                nbytes, mask = self.uint(k, 4), self.uint(k + 4, 4)
                offs = [self.uint(k + 8 + 8 * j, 8) for j in range(nd)]
                raw = self.bytes(child, nbytes)
                for fi, (fid, cd) in reversed(list(enumerate(filters))):
                    if mask & (1 << fi):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cd[0] if cd else dtype.itemsize
                        a = np.frombuffer(raw, np.uint8)
                        n = a.size // es
                        raw = a[:n * es].reshape(es, n).T.tobytes() + a[n * es:].tobytes()
                    elif fid == 3:
                        raw = raw[:-4]
                    else:
                        raise Hdf5FormatError(f"{self.path}: filter {fid} is not supported (deflate / shuffle / fletcher32 only)")
                chunk = np.frombuffer(raw, dtype, count=int(np.prod(cshape))).reshape(cshape)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]

        walk(bt)
        return out


class File:
    """Read access by path: ``File(p)["/Mesh/mesh/topology"]`` -> ndarray; ``.keys(group)`` lists a group."""

    def __init__(self, path):
        self._r = _Reader(path)

    def __getitem__(self, path):
        return self._r.dataset(path)

    def keys(self, group="/"):
        return sorted(self._r.links(self._r.resolve(group)))

    def __contains__(self, path):
        try:
            self._r.resolve(path)
            return True
        except KeyError:
            return False

    def close(self):
        self._r = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class _H5pyFile:
    def __init__(self, path):
        import h5py
        self._f = h5py.File(path, "r")

    def __getitem__(self, path):
        return self._f[path][()]

    def keys(self, group="/"):
        return sorted(self._f[group].keys())

    def __contains__(self, path):
        return path in self._f

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def open_file(path):
    """h5py when it is installed, the built-in reader otherwise."""
    try:
        import h5py  # noqa: F401
    except ImportError:
        return File(path)
    return _H5pyFile(path)


# ------------------------------------------------------------------------------------------------ writer
def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits = (0x08 if dt.kind == "i" else 0)
        return struct.pack("<BBBBIHH", 0x10 | 0, bits, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        # IEEE: sign position, exponent location / size, mantissa location / size, exponent bias
        if dt.itemsize == 8:
            prop = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            sign = 63
        else:
            prop = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            sign = 31
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, sign, 0, dt.itemsize) + prop
    raise TypeError(f"unsupported dtype {dt}")


def _message(mtype, body):
    body = body + b"\0" * (-len(body) % 8)
    return struct.pack("<HHBBBB", mtype, len(body), 0, 0, 0, 0) + body


def _object_header(msgs):
    data = b"".join(msgs)
    return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(data)) + b"\0" * 4 + data


def write_file(path, datasets, chunk_rows=None, compress=False):
    """datasets: {"/Group/sub/name": ndarray}.  Earliest-format HDF5: superblock 0, one symbol-table node per group
    (at most 64 entries), little-endian datasets, 8-byte offsets and lengths.  Datasets are contiguous, or -- with
    chunk_rows -- chunked along the first axis (one B-tree v1 leaf, at most 64 chunks) and, with compress, passed
    through the shuffle + deflate pipeline h5py / meshio use for ``compression="gzip"``."""
    tree = {}
    for full, arr in datasets.items():
        parts = [s for s in full.split("/") if s]
        node = tree
        for s in parts[:-1]:
            node = node.setdefault(s, {})
            if not isinstance(node, dict):
                raise ValueError(f"{full}: a dataset is used as a group")
        node[parts[-1]] = np.ascontiguousarray(arr)
    out = bytearray(b"\0" * 96)               # superblock (56 bytes + 40-byte root symbol table entry)

    def alloc(b, align=8):
        out.extend(b"\0" * (-len(out) % align))
        a = len(out)
        out.extend(b)
        return a

    def put_dataset(arr):
        dt = arr.dtype.newbyteorder("<")
        space = struct.pack("<BBBBI", 1, arr.ndim, 0, 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape)
        msgs = [_message(0x01, space), _message(0x03, _dtype_message(dt))]
        if chunk_rows and arr.ndim >= 1 and arr.size:
            cshape = (min(chunk_rows, arr.shape[0]),) + arr.shape[1:]
            starts = list(range(0, arr.shape[0], cshape[0]))
            if len(starts) > 64:
                raise ValueError("write_file: more than 64 chunks in one dataset")
            keys = bytearray()
            for st in starts:
                blk = np.zeros(cshape, dt)
                part = arr[st:st + cshape[0]]
                blk[:part.shape[0]] = part
                raw = blk.tobytes()
                if compress:
                    raw = np.frombuffer(raw, np.uint8).reshape(-1, dt.itemsize).T.tobytes()      # shuffle
                    raw = zlib.compress(raw, 4)
                a = alloc(raw)
                keys += struct.pack("<II", len(raw), 0) + struct.pack("<Q", st) + b"\0" * (8 * arr.ndim) + struct.pack("<Q", a)
            keys += struct.pack("<II", 0, 0) + struct.pack("<Q", arr.shape[0]) + b"\0" * (8 * arr.ndim)
            bt = alloc(b"TREE" + struct.pack("<BBHQQ", 1, 0, len(starts), UNDEF, UNDEF) + bytes(keys))
            layout = struct.pack("<BBBQ", 3, 2, arr.ndim + 1, bt) + b"".join(struct.pack("<I", c) for c in cshape + (dt.itemsize,))
            if compress:
                msgs.append(_message(0x0B, struct.pack("<BB6x", 1, 2) + struct.pack("<HHHHII", 2, 0, 0, 1, dt.itemsize, 0)
                                     + struct.pack("<HHHHII", 1, 0, 0, 1, 4, 0)))
        else:
            a = alloc(arr.astype(dt, copy=False).tobytes()) if arr.size else UNDEF
            layout = struct.pack("<BBQQ", 3, 1, a, arr.size * dt.itemsize)
        return alloc(_object_header(msgs + [_message(0x08, layout)]))

    def put_group(node):
        if len(node) > 64:
            raise ValueError("write_file: more than 64 entries in one group")
        entries = []
        for name in sorted(node):
            child = node[name]
            entries.append((name, put_group(child) if isinstance(child, dict) else put_dataset(child)))
        heap = bytearray(b"\0" * 8)
        offs = []
        for name, _a in entries:
            offs.append(len(heap))
            nb = name.encode("utf-8") + b"\0"
            heap.extend(nb + b"\0" * (-len(nb) % 8))
        heap.extend(b"\0" * 16)
        hdata = alloc(bytes(heap))
        hdr = alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap), len(heap) - 16, hdata))
        # free block at the end of the heap: next = 1 (none), size 16
        out[hdata + len(heap) - 16:hdata + len(heap)] = struct.pack("<QQ", 1, 16)
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(entries)))
        for (name, a), o in zip(entries, offs):
            snod.extend(struct.pack("<QQII", o, a, 0, 0) + b"\0" * 16)
        snod.extend(b"\0" * ((64 - len(entries)) * 40))
        sn = alloc(bytes(snod))
        last = offs[-1] if offs else 0
        bt = alloc(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if entries else 0, UNDEF, UNDEF)
                   + struct.pack("<QQQ", 0, sn, last) + b"\0" * (2 * 32 * 16))
        return alloc(_object_header([_message(0x11, struct.pack("<QQ", bt, hdr))]))

    root = put_group(tree)
    sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 32, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(out), UNDEF)
    sb += struct.pack("<QQII", 0, root, 0, 0) + b"\0" * 16
    out[:len(sb)] = sb
    with open(path, "wb") as f:
        f.write(bytes(out))
