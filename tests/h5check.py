"""Strict structural validator for the HDF5 files the plugin writes -- TEST INFRASTRUCTURE.

Written from the HDF5 File Format Specification (version 2.0, the format libhdf5 1.8 reads with
"earliest" bounds), NOT from phonomena_b200/h5lite.py: it does not import it and shares no code with
its writer or reader.  It walks the file the way libhdf5 does (superblock -> root symbol-table entry
-> root object header -> group B-tree / local heap / symbol-table node -> dataset object headers ->
chunk B-trees) and checks every field the specification constrains, so that a file that passes is a
file libhdf5 / h5py can open (neither exists in this image; tests/test_h5_real.py runs when one does).

validate(path) returns a dict describing what was found (datasets with shape / layout / chunk map,
attribute names and values); it raises H5FormatError at the first violation.
"""
import os
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(AssertionError):
    pass


def _req(cond, msg, *a):
    if not cond:
        raise H5FormatError(msg % a if a else msg)


class _File:
    def __init__(self, path):
        self.size = os.path.getsize(path)
        with open(path, "rb") as fh:
            self.b = fh.read()

    def u(self, fmt, off):
        _req(0 <= off and off + struct.calcsize(fmt) <= self.size, "read of %s at %d runs past the end of the file (%d)", fmt, off, self.size)
        return struct.unpack_from("<" + fmt, self.b, off)

    def addr_ok(self, a, n=1, what="address"):
        _req(a != UNDEF and 0 <= a and a + n <= self.size, "%s %d (+%d) outside the file (%d bytes)", what, a, n, self.size)
        _req(a % 8 == 0, "%s %d is not 8-byte aligned", what, a)


def _datatype(f, raw, where):
    """Datatype message (IV.A.2.d).  Returns a numpy dtype string or ('S', n)."""
    _req(len(raw) >= 8, "%s: datatype message too short", where)
    cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", raw, 0)
    version, cls = cv >> 4, cv & 15
    _req(version == 1, "%s: datatype version %d (expected 1)", where, version)
    if cls == 1:        # floating point
        _req(size == 8, "%s: float of %d bytes", where, size)
        _req(b0 & 1 == 0, "%s: float is not little-endian", where)
        _req((b0 >> 1) & 7 == 0, "%s: float padding bits set", where)
        _req((b0 >> 4) & 3 == 2, "%s: mantissa normalisation %d (IEEE needs 2 = implied msb)", where, (b0 >> 4) & 3)
        _req(b0 >> 6 == 0 and b2 == 0, "%s: reserved float bits set", where)
        _req(b1 == 63, "%s: sign bit at %d (IEEE double: 63)", where, b1)
        _req(len(raw) >= 20, "%s: float properties missing", where)
        off, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", raw, 8)
        _req((off, prec, eloc, esize, mloc, msize, bias) == (0, 64, 52, 11, 0, 52, 1023),
             "%s: float properties %r are not IEEE-754 binary64", where, (off, prec, eloc, esize, mloc, msize, bias))
        return "<f8"
    if cls == 0:        # fixed point
        _req(size == 8, "%s: integer of %d bytes", where, size)
        _req(b0 & 1 == 0, "%s: integer is not little-endian", where)
        _req((b0 >> 3) & 1 == 1, "%s: integer is not signed two's complement", where)
        _req(b1 == 0 and b2 == 0 and b0 >> 4 == 0, "%s: reserved integer bits set", where)
        off, prec = struct.unpack_from("<HH", raw, 8)
        _req((off, prec) == (0, 64), "%s: integer bit offset / precision %r", where, (off, prec))
        return "<i8"
    if cls == 3:        # string
        _req(b0 & 15 in (0, 1, 2), "%s: string padding type %d", where, b0 & 15)
        _req(b0 >> 4 in (0, 1), "%s: string character set %d", where, b0 >> 4)
        _req(b1 == 0 and b2 == 0, "%s: reserved string bits set", where)
        _req(size >= 1, "%s: empty fixed-length string type", where)
        return ("S", size)
    raise H5FormatError("%s: unsupported datatype class %d" % (where, cls))


def _dataspace(raw, where):
    """Dataspace message version 1 (IV.A.2.b)."""
    _req(len(raw) >= 8, "%s: dataspace message too short", where)
    version, rank, flags, r0, r1 = struct.unpack_from("<BBBBI", raw, 0)
    _req(version == 1, "%s: dataspace version %d", where, version)
    _req(flags & ~1 == 0 and r0 == 0 and r1 == 0, "%s: dataspace flags / reserved bytes", where)
    need = 8 + 8 * rank * (2 if flags & 1 else 1)
    _req(len(raw) >= need, "%s: dataspace of rank %d needs %d bytes, message has %d", where, rank, need, len(raw))
    return tuple(struct.unpack_from("<%dQ" % rank, raw, 8)) if rank else ()


def _object_header(f, addr, where):
    """Version 1 object header (IV.A.1.a): list of (type, flags, data)."""
    f.addr_ok(addr, 16, where + " object header")
    version, r0, nmsg, refc, size = f.u("BBHII", addr)
    _req(version == 1 and r0 == 0, "%s: object header version %d / reserved %d", where, version, r0)
    _req(refc >= 1, "%s: object reference count %d", where, refc)
    _req(size % 8 == 0, "%s: header size %d not a multiple of 8", where, size)
    pos, end = addr + 16, addr + 16 + size        # prefix padded to 8 bytes
    _req(end <= f.size, "%s: object header runs past the end of the file", where)
    msgs = []
    for k in range(nmsg):
        _req(pos + 8 <= end, "%s: message %d starts beyond the header (%d > %d)", where, k, pos + 8, end)
        mtype, msize, flags, a, b, c = f.u("HHBBBB", pos)
        _req(msize % 8 == 0, "%s: message %d (type 0x%04x) size %d is not a multiple of 8", where, k, mtype, msize)
        _req((a, b, c) == (0, 0, 0), "%s: message %d reserved bytes", where, k)
        _req(pos + 8 + msize <= end, "%s: message %d overruns the header", where, k)
        _req(mtype != 0x0010, "%s: continuation messages are not expected", where)
        msgs.append((mtype, flags, f.b[pos + 8:pos + 8 + msize]))
        pos += 8 + msize
    _req(pos == end, "%s: messages fill %d of %d header bytes (libhdf5 requires null messages for gaps)", where, pos - addr - 16, size)
    return msgs


def _attribute(f, raw, where):
    """Attribute message version 1 (IV.A.2.m): name, datatype and dataspace each padded to 8 bytes."""
    version, r0, nsz, tsz, ssz = struct.unpack_from("<BBHHH", raw, 0)
    _req(version == 1 and r0 == 0, "%s: attribute version %d", where, version)
    p = 8
    name = raw[p:p + nsz]
    _req(nsz >= 2 and name[-1:] == b"\0" and b"\0" not in name[:-1], "%s: attribute name not a NUL-terminated string", where)
    name = name[:-1].decode("ascii")
    p += (nsz + 7) // 8 * 8
    kind = _datatype(f, raw[p:p + tsz], "%s attr %s" % (where, name))
    p += (tsz + 7) // 8 * 8
    shape = _dataspace(raw[p:p + ssz], "%s attr %s" % (where, name))
    p += (ssz + 7) // 8 * 8
    n = int(np.prod(shape)) if shape else 1
    if isinstance(kind, tuple):
        _req(shape == (), "%s attr %s: string arrays not expected", where, name)
        _req(p + kind[1] <= len(raw), "%s attr %s: data truncated", where, name)
        return name, raw[p:p + kind[1]].rstrip(b"\0").decode("utf-8")
    _req(p + 8 * n <= len(raw), "%s attr %s: %d elements do not fit the message", where, name, n)
    a = np.frombuffer(raw, kind, n, p)
    return name, (a.reshape(shape).copy() if shape else a[0].item())


def _chunk_btree(f, addr, ndims, frame_bytes, where, level_expected=None, out=None, bounds=None):
    """Version 1 B-tree, node type 1 (III.A.1): keys = chunk size, filter mask, ndims offsets; 2K+1 keys / 2K children
    of room per node (K = 32, the default for chunked raw data when the superblock is version 0)."""
    K = 32
    keysize = 8 + 8 * ndims
    f.addr_ok(addr, 24 + 2 * K * 8 + (2 * K + 1) * keysize, where + " chunk B-tree node")
    _req(f.b[addr:addr + 4] == b"TREE", "%s: chunk B-tree signature", where)
    ntype, level, used, left, right = f.u("BBHQQ", addr + 4)
    _req(ntype == 1, "%s: B-tree node type %d (expected 1 = raw data chunks)", where, ntype)
    _req(1 <= used <= 2 * K, "%s: B-tree node with %d entries (1..%d)", where, used, 2 * K)
    if level_expected is not None:
        _req(level == level_expected, "%s: child node level %d, parent says %d", where, level, level_expected)
    pos = addr + 24
    keys, children = [], []
    for k in range(used + 1):
        csize, mask = f.u("II", pos)
        offs = f.u("%dQ" % ndims, pos + 8)
        keys.append((csize, mask, offs))
        pos += keysize
        if k < used:
            children.append(f.u("Q", pos)[0])
            pos += 8
    for k in range(used):
        csize, mask, offs = keys[k]
        _req(mask == 0, "%s: chunk filter mask %d", where, mask)
        _req(offs[-1] == 0 and all(v == 0 for v in offs[:-2]), "%s: chunk offset %r (only the time axis may be non-zero)", where, offs)
        _req(keys[k + 1][2][-2] > offs[-2], "%s: B-tree keys not strictly increasing (%d then %d)", where, offs[-2], keys[k + 1][2][-2])
        if bounds is not None:
            _req(bounds[0] <= offs[-2] < bounds[1], "%s: key %d outside the parent's key range %r", where, offs[-2], bounds)
        if level == 0:
            _req(csize == frame_bytes, "%s: chunk of %d bytes, a frame has %d", where, csize, frame_bytes)
            f.addr_ok(children[k], frame_bytes, where + " chunk")
            _req(offs[-2] not in out, "%s: frame %d stored twice", where, offs[-2])
            out[offs[-2]] = children[k]
        else:
            _chunk_btree(f, children[k], ndims, frame_bytes, where, level - 1, out, (offs[-2], keys[k + 1][2][-2]))
    return level, left, right


def _dataset(f, addr, name):
    msgs = _object_header(f, addr, "dataset " + name)
    types = [m[0] for m in msgs]
    for need in (0x0001, 0x0003, 0x0008):
        _req(types.count(need) == 1, "dataset %s: message 0x%04x appears %d times", name, need, types.count(need))
    info = {"name": name}
    for mtype, flags, raw in msgs:
        if mtype == 0x0001:
            info["shape"] = _dataspace(raw, "dataset " + name)
        elif mtype == 0x0003:
            info["dtype"] = _datatype(f, raw, "dataset " + name)
            _req(flags & 1, "dataset %s: datatype message must be flagged constant", name)
        elif mtype == 0x0005:
            version, alloc, wtime, defined = struct.unpack_from("<BBBB", raw, 0)
            _req(version == 2 and alloc in (1, 2, 3) and wtime in (0, 1, 2) and defined in (0, 1), "dataset %s: fill value message %r", name, (version, alloc, wtime, defined))
            info["fill"] = (alloc, wtime, defined)
    _req(info["dtype"] == "<f8", "dataset %s is not float64", name)
    raw = next(m[2] for m in msgs if m[0] == 0x0008)
    version, lclass = raw[0], raw[1]
    _req(version == 3, "dataset %s: layout version %d", name, version)
    shape = info["shape"]
    if lclass == 1:
        a, n = struct.unpack_from("<QQ", raw, 2)
        _req(n == 8 * int(np.prod(shape)), "dataset %s: contiguous size %d for shape %r", name, n, shape)
        if n:
            f.addr_ok(a, n, "dataset %s data" % name)
        info.update(layout="contiguous", addr=a)
    elif lclass == 2:
        nd = raw[2]
        _req(nd == len(shape) + 1, "dataset %s: chunk dimensionality %d for rank %d", name, nd, len(shape))
        bt = struct.unpack_from("<Q", raw, 3)[0]
        cdims = struct.unpack_from("<%dI" % nd, raw, 11)
        _req(cdims[-1] == 8, "dataset %s: chunk element size %d", name, cdims[-1])
        _req(tuple(cdims[:-2]) == tuple(shape[:-1]) and cdims[-2] == 1, "dataset %s: chunk dims %r for shape %r (one frame per chunk)", name, cdims, shape)
        _req(info.get("fill", (3,))[0] == 3, "dataset %s: chunked data must use incremental allocation", name)
        chunks = {}
        if bt != UNDEF:
            frame_bytes = 8 * int(np.prod(shape[:-1]))
            _chunk_btree(f, bt, nd, frame_bytes, "dataset " + name, None, chunks)
            _req(all(0 <= t < shape[-1] for t in chunks), "dataset %s: frame index outside [0, %d)", name, shape[-1])
        info.update(layout="chunked", chunks=chunks)
    else:
        raise H5FormatError("dataset %s: layout class %d" % (name, lclass))
    return info


def validate(path):
    f = _File(path)
    b = f.b
    # ---- superblock version 0 (II.A) ----
    _req(b[:8] == b"\x89HDF\r\n\x1a\n", "format signature")
    sb_ver, fs_ver, root_ver, r0, shm_ver, so, sl, r1 = f.u("BBBBBBBB", 8)
    _req((sb_ver, fs_ver, root_ver, r0, shm_ver, r1) == (0, 0, 0, 0, 0, 0), "superblock version bytes %r", (sb_ver, fs_ver, root_ver, r0, shm_ver, r1))
    _req(so == 8 and sl == 8, "size of offsets / lengths %d / %d", so, sl)
    leaf_k, int_k, flags = f.u("HHI", 16)
    _req(leaf_k >= 1 and int_k >= 1, "group K values %d / %d", leaf_k, int_k)
    _req(flags == 0, "file consistency flags %d (file not closed cleanly?)", flags)
    base, freesp, eof, drv = f.u("QQQQ", 24)
    _req(base == 0 and freesp == UNDEF and drv == UNDEF, "base / free-space / driver addresses %r", (base, freesp, drv))
    _req(eof == f.size, "end-of-file address %d, file has %d bytes", eof, f.size)
    name_off, root_addr, cache, r2 = f.u("QQII", 56)
    _req(name_off == 0 and r2 == 0, "root symbol table entry name offset / reserved")
    _req(cache == 1, "root entry cache type %d (expected 1: B-tree + heap cached)", cache)
    sc_bt, sc_heap = f.u("QQ", 80)
    # ---- root group ----
    msgs = _object_header(f, root_addr, "root group")
    stab = [m for m in msgs if m[0] == 0x0011]
    _req(len(stab) == 1, "root group has %d symbol table messages", len(stab))
    bt_addr, heap_addr = struct.unpack_from("<QQ", stab[0][2], 0)
    _req((bt_addr, heap_addr) == (sc_bt, sc_heap), "scratch-pad B-tree / heap addresses differ from the symbol table message")
    attrs = {}
    for mtype, flags, raw in msgs:
        if mtype == 0x000C:
            k, v = _attribute(f, raw, "root")
            _req(k not in attrs, "attribute %s stored twice", k)
            attrs[k] = v
        else:
            _req(mtype in (0x0011, 0x0000), "unexpected message 0x%04x in the root group header", mtype)
    # local heap (III.D)
    f.addr_ok(heap_addr, 32, "local heap")
    _req(b[heap_addr:heap_addr + 4] == b"HEAP", "local heap signature")
    hver, h0, h1, h2 = f.u("BBBB", heap_addr + 4)
    _req((hver, h0, h1, h2) == (0, 0, 0, 0), "local heap version / reserved")
    hsize, hfree, hdata = f.u("QQQ", heap_addr + 8)
    f.addr_ok(hdata, hsize, "local heap data segment")
    _req(hsize % 8 == 0 and hsize >= 8, "local heap data size %d", hsize)
    _req(hfree == 1 or hfree == UNDEF or (hfree % 8 == 0 and hfree + 16 <= hsize), "local heap free-list head %d", hfree)
    _req(b[hdata] == 0, "heap offset 0 must hold the empty string")

    def heap_name(off):
        _req(0 <= off < hsize, "name offset %d outside the heap (%d)", off, hsize)
        end = b.find(b"\0", hdata + off, hdata + hsize)
        _req(end >= 0, "unterminated name at heap offset %d", off)
        return b[hdata + off:end].decode("ascii")

    # group B-tree (III.A.1, node type 0) -- a single leaf-level node is all the writer needs
    f.addr_ok(bt_addr, 24 + 2 * int_k * 8 + (2 * int_k + 1) * 8, "group B-tree node")
    _req(b[bt_addr:bt_addr + 4] == b"TREE", "group B-tree signature")
    ntype, level, used, left, right = f.u("BBHQQ", bt_addr + 4)
    _req(ntype == 0 and level == 0, "group B-tree node type %d level %d", ntype, level)
    _req(left == UNDEF and right == UNDEF, "root B-tree node has siblings")
    _req(used <= 2 * int_k, "group B-tree entries %d", used)
    datasets, names_seen = {}, []
    pos = bt_addr + 24
    prev_key = f.u("Q", pos)[0]
    _req(used == 0 or heap_name(prev_key) == "", "first group B-tree key must be the empty string")
    for e in range(used):
        child, key = f.u("QQ", pos + 8)
        pos += 16
        f.addr_ok(child, 8 + 2 * leaf_k * 40, "symbol table node")
        _req(b[child:child + 4] == b"SNOD", "symbol table node signature")
        sver, s0, nsym = f.u("BBH", child + 4)
        _req(sver == 1 and s0 == 0, "symbol table node version %d", sver)
        _req(1 <= nsym <= 2 * leaf_k, "symbol table node holds %d symbols (1..%d)", nsym, 2 * leaf_k)
        node_names = []
        for s in range(nsym):
            noff, oaddr, ctype, r3 = f.u("QQII", child + 8 + 40 * s)
            _req(ctype == 0 and r3 == 0, "symbol %d: cache type %d", s, ctype)
            nm = heap_name(noff)
            _req(nm != "", "empty link name")
            node_names.append(nm)
            datasets[nm] = _dataset(f, oaddr, nm)
        _req(node_names == sorted(node_names) and len(set(node_names)) == len(node_names), "symbols not in strictly increasing name order: %r", node_names)
        _req(heap_name(key) == node_names[-1], "B-tree key %r is not the last name of its node (%r)", heap_name(key), node_names[-1])
        _req(not names_seen or names_seen[-1] < node_names[0], "symbol table nodes out of order")
        names_seen += node_names
    # no two extents may overlap: chunks, contiguous data
    ext = []
    for d in datasets.values():
        if d["layout"] == "contiguous" and d["shape"] and int(np.prod(d["shape"])):
            ext.append((d["addr"], d["addr"] + 8 * int(np.prod(d["shape"])), d["name"]))
        elif d["layout"] == "chunked":
            n = 8 * int(np.prod(d["shape"][:-1]))
            ext += [(a, a + n, "%s[%d]" % (d["name"], t)) for t, a in d["chunks"].items()]
    ext.sort()
    for (a0, a1, n0), (b0, b1, n1) in zip(ext, ext[1:]):
        _req(a1 <= b0, "data extents overlap: %s [%d, %d) and %s [%d, %d)", n0, a0, a1, n1, b0, b1)
    _req(not ext or ext[0][0] >= 96, "data overlaps the superblock")
    return {"attrs": attrs, "datasets": datasets, "eof": eof}


def read_frame(path, info, name, t):
    """Frame t of a chunked dataset through the validator's own chunk map."""
    d = info["datasets"][name]
    fshape = d["shape"][:-1]
    with open(path, "rb") as fh:
        fh.seek(d["chunks"][t])
        return np.frombuffer(fh.read(8 * int(np.prod(fshape))), "<f8").reshape(fshape)
