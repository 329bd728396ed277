"""Minimal HDF5 writer/reader (no h5py / libhdf5 in the target image).

Writes the subset of the HDF5 1.8 "earliest" file format the reference's output schema needs
(base_solver.py:105-133): superblock v0, one root group (v1 object header,
symbol table = v1 B-tree + local heap + one symbol-table node), float64 datasets with
contiguous layout (written at once) or chunked layout with one chunk per time frame
(v3 layout message + v1 chunk B-tree, built when the file is closed), and root attributes
(float64 arrays of any rank, int64/float64 scalars, fixed-length strings).

Frames are appended sequentially, so a frame write is one contiguous file write; all metadata
goes to the end of the file at close and the superblock at offset 0 is rewritten last.

LIMITATION (stated in DESIGN.md): no HDF5 library is available in the build image, so files are
validated by the independent reader below and by structure tests, not by libhdf5 itself.
"""
from __future__ import annotations

import mmap
import os
import struct
import threading

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"
GROUP_LEAF_K, GROUP_INT_K, CHUNK_K = 4, 16, 32     # library defaults for a v0 superblock


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


# ------------------------------------------------------------------------------------------
# message encoders
# ------------------------------------------------------------------------------------------
def _dt_f64():
    # class 1 (float) v1; LE, mantissa normalisation "implied msb" (2<<4), sign bit 63
    return struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)


def _dt_i64():
    return struct.pack("<BBBBI", 0x10, 0x08, 0, 0, 8) + struct.pack("<HH", 0, 64)


def _dt_str(n):
    # class 3 (string) v1; null-padded (1), ASCII/UTF-8 compatible charset 0
    return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, n)


def _dataspace(shape):
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _msg(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data


def _attr_msg(name, value):
    if isinstance(value, str):
        raw = value.encode("utf-8")
        raw = raw if raw else b"\0"
        dt, ds, data = _dt_str(len(raw)), _dataspace(()), raw
    elif isinstance(value, (int, np.integer)) and not isinstance(value, bool):
        dt, ds, data = _dt_i64(), _dataspace(()), struct.pack("<q", int(value))
    else:
        a = np.asarray(value, dtype="<f8")          # 0-d stays 0-d (scalar dataspace)
        dt, ds, data = _dt_f64(), _dataspace(a.shape), a.tobytes()
    nm = name.encode("ascii") + b"\0"
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + data
    if len(body) + 8 > 0xFFF8:
        raise ValueError("attribute %r too large for a v1 object header message" % name)
    return _msg(0x000C, body)


def _object_header(msgs):
    body = b"".join(msgs)
    return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\0" * 4 + body


class _Chunked:
    def __init__(self, name, shape):
        self.name, self.shape = name, tuple(int(s) for s in shape)
        self.frame_bytes = int(np.prod(self.shape[:-1])) * 8
        self.addr = {}          # time index -> file address


class H5Writer:
    """hdf = H5Writer(path); hdf.attrs[...] = ...; hdf.create_dataset(name, array);
    d = hdf.create_chunked(name, (a, b, c, frames)); hdf.write_frame(d, t, array); hdf.close()"""

    def __init__(self, path):
        # raw descriptor + positional writes: frames of different datasets can be written by several threads at once
        self.fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)      # read-write: the native writer may map the frame extents
        self._end = 96                      # superblock placeholder
        self._lock = threading.Lock()
        self.attrs = {}
        self._contig = []                   # (name, shape, addr, nbytes)
        self._chunked = []
        self.closed = False

    # -- data ------------------------------------------------------------------------------
    def _pwrite(self, raw, pos):
        mv = memoryview(raw).cast("B")
        while len(mv):
            n = os.pwrite(self.fd, mv, pos)
            mv, pos = mv[n:], pos + n

    def _append(self, raw):
        """Reserve an 8-byte aligned extent at the end of the file and write `raw` there (thread-safe; the
        write itself runs outside the lock)."""
        nbytes = memoryview(raw).nbytes
        with self._lock:
            pos = self._end + (-self._end % 8)
            self._end = pos + nbytes
        self._pwrite(raw, pos)
        return pos

    def create_dataset(self, name, data):
        a = np.ascontiguousarray(np.asarray(data, dtype="<f8"))
        addr = self._append(a.tobytes()) if a.size else UNDEF
        self._contig.append((name, a.shape, addr, a.nbytes))

    def create_dataset_from(self, name, shape, pieces):
        """Contiguous float64 dataset of `shape` written piece by piece (an iterable of arrays that concatenate, in C
        order, to the whole): the 1 GB `density` of a 512^3 grid never exists as one host array nor as a bytes copy."""
        shape = tuple(int(v) for v in shape)
        nbytes = int(np.prod(shape)) * 8
        with self._lock:
            addr = self._end + (-self._end % 8)
            self._end = addr + nbytes
        pos = addr
        for a in pieces:
            a = np.ascontiguousarray(np.asarray(a, dtype="<f8"))
            self._pwrite(a.reshape(-1).data, pos)
            pos += a.nbytes
        if pos != addr + nbytes:
            raise ValueError("dataset %s: pieces hold %d bytes, shape %s needs %d" % (name, pos - addr, shape, nbytes))
        self._contig.append((name, shape, addr if nbytes else UNDEF, nbytes))

    def settle(self):
        """Flush what has been written so far to the device (used after the large static datasets)."""
        try:
            os.fsync(self.fd)
        except OSError:
            pass

    def preallocate(self, nbytes, start=None):
        """Reserve file space for `nbytes` of data now (posix_fallocate), from `start` (default: the current end): on
        tmpfs / page cache the pages are then allocated outside the stepping loop and the frame writes are plain
        copies.  Best effort."""
        try:
            os.posix_fallocate(self.fd, self._end if start is None else int(start), int(nbytes))
            return True
        except (OSError, AttributeError):
            return False

    def create_chunked(self, name, shape):
        d = _Chunked(name, shape)
        self._chunked.append(d)
        return d

    def write_frame(self, d, t, frame):
        a = np.ascontiguousarray(np.asarray(frame, dtype="<f8"))
        if a.nbytes != d.frame_bytes:
            raise ValueError("frame of %d bytes for dataset %s, expected %d" % (a.nbytes, d.name, d.frame_bytes))
        with self._lock:
            if not (0 <= t < d.shape[-1]) or t in d.addr:
                raise ValueError("bad or repeated frame index %d for %s" % (t, d.name))
            d.addr[t] = None
        d.addr[t] = self._append(a.data)

    def reserve_frame(self, d, t):
        """Reserve the extent of frame t of chunked dataset d and return its file offset; the caller fills it
        with pwrite() calls (any number of pieces, from any thread) before close()."""
        with self._lock:
            if not (0 <= t < d.shape[-1]) or t in d.addr:
                raise ValueError("bad or repeated frame index %d for %s" % (t, d.name))
            pos = self._end + (-self._end % 8)
            self._end = pos + d.frame_bytes
            d.addr[t] = pos
        return pos

    def pwrite(self, raw, pos):
        self._pwrite(raw, pos)

    def reserve_frames(self, datasets, frames):
        """Reserve the extents of frames 0 .. frames-1 of several chunked datasets at once, interleaved frame by
        frame (frame t of every dataset, then frame t+1, ...) so that a run appends to the file sequentially.
        Returns (base, stride): frame t of datasets[c] lives at base[c] + t * stride.  The native writer threads
        (phb_writer_start) fill them; call keep_frames(n) before close() if fewer were written."""
        with self._lock:
            pos = self._end + (-self._end % 8)
            stride = sum(d.frame_bytes for d in datasets)
            base, off = [], 0
            for d in datasets:
                if d.addr:
                    raise ValueError("dataset %s already has frames" % d.name)
                if frames > d.shape[-1]:
                    raise ValueError("%d frames for dataset %s of %d" % (frames, d.name, d.shape[-1]))
                base.append(pos + off)
                for t in range(frames):
                    d.addr[t] = pos + off + t * stride
                off += d.frame_bytes
            self._end = pos + stride * frames
        return base, stride

    def keep_frames(self, n):
        """Forget the reserved frames with index >= n (a cancelled run wrote fewer than it reserved): readers
        then see them as never written (fill value) instead of whatever the extent holds."""
        for d in self._chunked:
            for t in [t for t in d.addr if t >= n]:
                del d.addr[t]

    # -- chunk B-tree (v1, node type 1) --------------------------------------------------------
    def _chunk_btree(self, d):
        rank = len(d.shape)
        items = sorted(d.addr.items())
        if not items:
            return UNDEF
        keysize = 8 + 8 * (rank + 1)
        node_bytes = 24 + 2 * CHUNK_K * 8 + (2 * CHUNK_K + 1) * keysize

        def key(t, nbytes):
            return struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", v) for v in (0,) * (rank - 1) + (t, 0))

        # level 0 entries: (first_t, last_t_plus_1, child_addr)
        level, entries = 0, [(t, t + 1, addr) for t, addr in items]
        while True:
            groups = [entries[i:i + 2 * CHUNK_K] for i in range(0, len(entries), 2 * CHUNK_K)]
            base = self._end + (-self._end % 8)
            addrs = [base + n * node_bytes for n in range(len(groups))]
            out = []
            for n, grp in enumerate(groups):
                left = addrs[n - 1] if n > 0 else UNDEF
                right = addrs[n + 1] if n + 1 < len(groups) else UNDEF
                b = b"TREE" + struct.pack("<BBHQQ", 1, level, len(grp), left, right)
                for (t0, _t1, child) in grp:
                    b += key(t0, d.frame_bytes) + struct.pack("<Q", child)
                b += key(grp[-1][1], 0)                      # final key: one past the last chunk
                b += b"\0" * (node_bytes - len(b))
                out.append(b)
            pos = self._append(b"".join(out))
            assert pos == base
            entries = [(grp[0][0], grp[-1][1], addrs[n]) for n, grp in enumerate(groups)]
            if len(entries) == 1:
                return entries[0][2]
            level += 1

    # -- close: all metadata -----------------------------------------------------------------
    def close(self):
        if self.closed:
            return
        try:
            self._write_metadata()
        finally:            # the descriptor is released whatever the metadata writes hit (ENOSPC, EBADF, ...)
            self.closed = True
            try:
                os.close(self.fd)
            except OSError:
                pass

    def _write_metadata(self):
        fill_early = struct.pack("<BBBB", 2, 1, 2, 0)           # fill value v2: alloc early, write "if set", undefined
        fill_incr = struct.pack("<BBBB", 2, 3, 2, 0)            # alloc incremental (chunked)
        objects = {}                                            # name -> object header address
        for name, shape, addr, nbytes in self._contig:
            msgs = [_msg(0x0001, _dataspace(shape)), _msg(0x0003, _dt_f64(), flags=1), _msg(0x0005, fill_early),
                    _msg(0x0008, struct.pack("<BBQQ", 3, 1, addr, nbytes))]
            objects[name] = self._append(_object_header(msgs))
        for d in self._chunked:
            bt = self._chunk_btree(d)
            rank = len(d.shape)
            chunk_dims = d.shape[:-1] + (1, 8)
            layout = struct.pack("<BBBQ", 3, 2, rank + 1, bt) + b"".join(struct.pack("<I", v) for v in chunk_dims)
            msgs = [_msg(0x0001, _dataspace(d.shape)), _msg(0x0003, _dt_f64(), flags=1), _msg(0x0005, fill_incr),
                    _msg(0x0008, layout)]
            objects[d.name] = self._append(_object_header(msgs))
        names = sorted(objects)
        if len(names) > 2 * GROUP_LEAF_K:
            raise ValueError("h5lite supports at most %d objects in the root group" % (2 * GROUP_LEAF_K))
        # local heap: offset 0 = empty string
        heap, offs = bytearray(b"\0" * 8), {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode("ascii") + b"\0")
        heap_data = self._append(bytes(heap))
        heap_addr = self._append(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap), 1, heap_data))
        # symbol table node
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for n in names:
            snod += struct.pack("<QQII", offs[n], objects[n], 0, 0) + b"\0" * 16
        snod += b"\0" * (8 + 2 * GROUP_LEAF_K * 40 - len(snod))
        snod_addr = self._append(snod)
        # group B-tree (node type 0, level 0, one child)
        bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF)
        bt += struct.pack("<QQQ", 0, snod_addr, offs[names[-1]] if names else 0)
        bt += b"\0" * (24 + 2 * GROUP_INT_K * 8 + (2 * GROUP_INT_K + 1) * 8 - len(bt))
        bt_addr = self._append(bt)
        # root group object header: symbol table message + attributes
        msgs = [_msg(0x0011, struct.pack("<QQ", bt_addr, heap_addr))]
        msgs += [_attr_msg(k, v) for k, v in self.attrs.items()]
        root_addr = self._append(_object_header(msgs))
        eof = self._end
        sb = SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", GROUP_LEAF_K, GROUP_INT_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", bt_addr, heap_addr)
        assert len(sb) == 96
        self._pwrite(sb, 0)
        try:
            os.ftruncate(self.fd, eof)          # drop preallocated space that was not used
        except OSError:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ------------------------------------------------------------------------------------------
# independent reader (tests, and bench/e2e verification)
# ------------------------------------------------------------------------------------------
class H5Reader:
    """Parses exactly what H5Writer emits, following the addresses stored in the file."""

    def __init__(self, path):
        self._fh = open(path, "rb")
        self.b = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)     # frames are read where they lie
        self._chunks = {}
        b = self.b
        if b[:8] != SIG:
            raise ValueError("not an HDF5 file")
        if b[8] != 0 or b[13] != 8 or b[14] != 8:
            raise ValueError("unsupported superblock")
        self.eof = struct.unpack_from("<Q", b, 40)[0]
        root_addr, cache = struct.unpack_from("<QI", b, 64)
        assert cache == 1
        self.attrs, self.datasets = {}, {}
        bt_addr = heap_addr = None
        for mtype, data in self._messages(root_addr):
            if mtype == 0x0011:
                bt_addr, heap_addr = struct.unpack_from("<QQ", data, 0)
            elif mtype == 0x000C:
                k, v = self._attr(data)
                self.attrs[k] = v
        assert b[heap_addr:heap_addr + 4] == b"HEAP"
        heap_data = struct.unpack_from("<Q", b, heap_addr + 24)[0]
        assert b[bt_addr:bt_addr + 4] == b"TREE" and b[bt_addr + 4] == 0
        nent = struct.unpack_from("<H", b, bt_addr + 6)[0]
        for e in range(nent):
            snod = struct.unpack_from("<Q", b, bt_addr + 24 + 8 + e * 16)[0]
            assert b[snod:snod + 4] == b"SNOD"
            nsym = struct.unpack_from("<H", b, snod + 6)[0]
            for s in range(nsym):
                noff, oaddr = struct.unpack_from("<QQ", b, snod + 8 + s * 40)
                end = b.find(b"\0", heap_data + noff)
                self.datasets[b[heap_data + noff:end].decode()] = oaddr

    def _messages(self, addr):
        b = self.b
        ver, _, nmsg, _ref, size = struct.unpack_from("<BBHII", b, addr)
        assert ver == 1
        pos, end = addr + 16, addr + 16 + size
        for _ in range(nmsg):
            mtype, msize = struct.unpack_from("<HH", b, pos)
            yield mtype, b[pos + 8:pos + 8 + msize]
            pos += 8 + msize
        assert pos == end

    @staticmethod
    def _dtype(dt):
        cls = dt[0] & 15
        size = struct.unpack_from("<I", dt, 4)[0]
        return {1: "<f8", 0: "<i8"}.get(cls, "S%d" % size)

    def _attr(self, data):
        _v, _r, nsz, tsz, ssz = struct.unpack_from("<BBHHH", data, 0)
        p = 8
        name = data[p:p + nsz - 1].decode()
        p += (nsz + 7) // 8 * 8
        dt = data[p:p + tsz]
        p += (tsz + 7) // 8 * 8
        rank = data[p + 1]
        shape = struct.unpack_from("<%dQ" % rank, data, p + 8) if rank else ()
        p += (ssz + 7) // 8 * 8
        kind = self._dtype(dt)
        n = int(np.prod(shape)) if rank else 1
        if kind.startswith("S"):
            return name, data[p:p + int(kind[1:])].rstrip(b"\0").decode("utf-8")
        a = np.frombuffer(data, kind, n, p).reshape(shape)
        return name, (a.copy() if rank else a.reshape(()).item())

    def shape(self, name):
        for mtype, data in self._messages(self.datasets[name]):
            if mtype == 0x0001:
                return struct.unpack_from("<%dQ" % data[1], data, 8)

    def close(self):
        try:
            self.b.close()
        except (BufferError, ValueError):      # views into the map are still alive: the map goes with them
            pass
        self._fh.close()

    def _layout(self, name):
        for mtype, data in self._messages(self.datasets[name]):
            if mtype == 0x0008:
                return data
        raise KeyError(name)

    def is_chunked(self, name):
        return self._layout(name)[1] == 2

    def frame_view(self, name, t):
        """Frame t of a chunked dataset as a read-only array over the file mapping (no copy); a frame that was
        never written reads as zeros (the HDF5 fill value)."""
        shape = self.shape(name)
        layout = self._layout(name)
        assert layout[0] == 3 and layout[1] == 2
        if name not in self._chunks:
            self._chunks[name] = {}
            self._walk(struct.unpack_from("<Q", layout, 3)[0], layout[2], self._chunks[name])
        fshape = tuple(shape[:-1])
        addr = self._chunks[name].get(int(t))
        if addr is None:
            return np.zeros(fshape)
        return np.frombuffer(self.b, "<f8", int(np.prod(fshape)), addr).reshape(fshape)

    def read(self, name, frame=None):
        """Whole dataset (contiguous, or chunked with all frames present) or one time frame."""
        b = self.b
        shape = layout = None
        for mtype, data in self._messages(self.datasets[name]):
            if mtype == 0x0001:
                shape = struct.unpack_from("<%dQ" % data[1], data, 8)
            elif mtype == 0x0008:
                layout = data
        assert layout[0] == 3
        if layout[1] == 1:
            addr, nbytes = struct.unpack_from("<QQ", layout, 2)
            return np.frombuffer(b, "<f8", nbytes // 8, addr).reshape(shape).copy()
        assert layout[1] == 2
        nd = layout[2]
        bt = struct.unpack_from("<Q", layout, 3)[0]
        cdims = struct.unpack_from("<%dI" % nd, layout, 11)
        assert tuple(cdims[:-1]) == tuple(shape[:-1]) + (1,) and cdims[-1] == 8
        if name not in self._chunks:
            self._chunks[name] = {}
            self._walk(bt, nd, self._chunks[name])
        chunks = self._chunks[name]
        fshape = tuple(shape[:-1])
        n = int(np.prod(fshape))
        if frame is not None:
            return np.frombuffer(b, "<f8", n, chunks[frame]).reshape(fshape).copy()
        out = np.zeros(shape)
        for t, addr in chunks.items():
            out[..., t] = np.frombuffer(b, "<f8", n, addr).reshape(fshape)
        return out

    def _walk(self, addr, nd, out):
        b = self.b
        if addr == UNDEF:
            return
        assert b[addr:addr + 4] == b"TREE" and b[addr + 4] == 1
        level, nent = b[addr + 5], struct.unpack_from("<H", b, addr + 6)[0]
        keysize = 8 + 8 * nd
        pos = addr + 24
        for _ in range(nent):
            offs = struct.unpack_from("<%dQ" % nd, b, pos + 8)
            child = struct.unpack_from("<Q", b, pos + keysize)[0]
            if level == 0:
                out[offs[nd - 2]] = child
            else:
                self._walk(child, nd, out)
            pos += keysize + 8


def merge_slabs(paths, out_path):
    """Concatenate the per-slab output files of a multi-GPU run (attrs x0 / nxl, one file per rank,
    solver_b200.Solver) along x into one file with the single-GPU schema.  Frame by frame, so the
    working set is one frame."""
    rs = sorted((H5Reader(p) for p in paths), key=lambda r: int(r.attrs["x0"]))
    x0 = 0
    for r in rs:
        if int(r.attrs["x0"]) != x0:
            raise ValueError("slab files do not tile x: expected a slab starting at %d, found %d" % (x0, int(r.attrs["x0"])))
        x0 += int(r.attrs["nxl"])
    nx = len(np.atleast_1d(rs[0].attrs.get("x_full", rs[0].attrs["x"])))      # decimated volumes keep the full mesh lines as x_full
    if x0 != nx:
        raise ValueError("slab files cover %d of %d planes" % (x0, nx))
    names = [k for k in ("ux", "uy", "uz") if all(k in r.datasets for r in rs)]       # cfg["record_fields"] may drop some
    frames = min(int(r.attrs.get("frames_written", r.shape(names[0])[3])) for r in rs) if names else 0
    with H5Writer(out_path) as w:
        w.attrs.update(rs[0].attrs)
        w.attrs.update({"x0": 0, "nxl": nx, "frames_written": frames})
        for name in ("density", "elasticity"):
            if all(name in r.datasets for r in rs):
                w.create_dataset(name, np.concatenate([r.read(name) for r in rs], axis=0))
        for name in names:
            shapes = [r.shape(name) for r in rs]
            full = (sum(sh[0] for sh in shapes),) + tuple(shapes[0][1:3]) + (shapes[0][3],)
            d = w.create_chunked(name, full)
            for t in range(frames):
                w.write_frame(d, t, np.concatenate([r.read(name, frame=t) for r in rs], axis=0))
    return out_path
