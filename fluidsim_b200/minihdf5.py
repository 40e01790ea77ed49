"""Minimal HDF5 writer / reader (pure Python + numpy) for fluidsim's ``state_phys`` files.

fluidsim stores checkpoints with h5py / h5netcdf (``/root/reference/fluidsim/util/output.py:47-160``,
``base/init_fields.py:140-298``).  Neither library nor libhdf5 exists in this image, so the subset of the
HDF5 file format those files need is written here directly, following the published HDF5 File Format
Specification (version 0 superblock, version 1 object headers, "old style" groups = v1 B-tree + local
heap + symbol-table nodes, contiguous datasets, version 1 attribute messages) -- the layout libhdf5
itself produces with its default ``libver='earliest'`` settings.

**Format parity is unpinned**: no HDF5 implementation is available offline to open the files this
module writes; the tests check writer -> reader round trips and the structural rules of the
specification only.

Data model: a *group* is a ``dict`` with optional key ``"@attrs"`` (dict of attributes); every other key
maps to a child group (``dict``) or a dataset (``numpy.ndarray``).  Attribute values: ``int``, ``float``,
``bool`` (stored as int, like h5py does for Python bools -> 8-bit enum is NOT used), ``str`` / ``bytes``
(fixed-length strings), ``None`` (stored as the string "None", fluiddyn's convention), numpy arrays of
int64 / float64 / ``|S``.
"""

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K = 16  # symbol-table node holds up to 2 * LEAF_K entries (stored in the superblock)
INTERNAL_K = 16
H5HL_FREE_NULL = 1  # end-of-free-list marker of a local heap


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


# ------------------------------------------------------------------------------------------- datatypes
def _datatype_message(dtype):
    """Datatype message body (version 1) for a numpy dtype."""
    dtype = np.dtype(dtype)
    if dtype.kind in "iu":
        bits0 = 0x08 if dtype.kind == "i" else 0x00  # little endian, zero padding, signedness
        return struct.pack("<BBBBI", 0x10 | 0, bits0, 0, 0, dtype.itemsize) + struct.pack(
            "<HH", 0, 8 * dtype.itemsize
        )
    if dtype.kind == "f":
        if dtype.itemsize == 8:
            sign, eloc, esize, msize, bias = 63, 52, 11, 52, 1023
        elif dtype.itemsize == 4:
            sign, eloc, esize, msize, bias = 31, 23, 8, 23, 127
        else:
            raise TypeError(dtype)
        # byte 0: little endian, mantissa normalisation 2 (msb implied); byte 1: sign bit location
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, sign, 0, dtype.itemsize) + struct.pack(
            "<HHBBBBI", 0, 8 * dtype.itemsize, eloc, esize, 0, msize, bias
        )
    if dtype.kind == "S":
        return struct.pack("<BBBBI", 0x10 | 3, 0x01, 0, 0, max(dtype.itemsize, 1))  # null-padded ASCII
    raise TypeError(f"unsupported dtype {dtype}")


def _dataspace_message(shape):
    """Dataspace message body (version 1); shape () = scalar."""
    body = struct.pack("<BBBBI", 1, len(shape), 0, 0, 0)
    for n in shape:
        body += struct.pack("<Q", int(n))
    return body


def _normalise_attr(value):
    """Python attribute value -> numpy array (0-d for scalars)."""
    if value is None:
        value = "None"
    if isinstance(value, (bool, np.bool_)):
        return np.array(int(value), dtype="<i8")
    if isinstance(value, (int, np.integer)):
        return np.array(int(value), dtype="<i8")
    if isinstance(value, (float, np.floating)):
        return np.array(float(value), dtype="<f8")
    if isinstance(value, str):
        value = value.encode("utf-8")
    if isinstance(value, (bytes, np.bytes_)):
        return np.array(bytes(value), dtype=f"|S{max(len(value), 1)}")
    a = np.asarray(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf-8")
    if a.dtype.kind == "b":
        a = a.astype("<i8")
    if a.dtype.kind == "O":
        raise TypeError(f"cannot store attribute value {value!r}")
    return np.ascontiguousarray(a)


def _message(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHBBBB", mtype, len(body), flags, 0, 0, 0) + body


def _attribute_message(name, value):
    a = _normalise_attr(value)
    nameb = name.encode("utf-8") + b"\0"
    dt = _datatype_message(a.dtype)
    ds = _dataspace_message(a.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nameb), len(dt), len(ds))
    body += _pad8(nameb) + _pad8(dt) + _pad8(ds) + a.tobytes()
    if len(body) > 0xFFF0:
        raise ValueError(f"attribute {name!r} too large for a version 1 object header message")
    return _message(0x000C, body)


def _object_header(messages):
    data = b"".join(messages)
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(data)) + b"\0" * 4 + data


# ----------------------------------------------------------------------------------------------- writer
class _Writer:
    def __init__(self):
        self.buf = bytearray(96)  # superblock written last

    def alloc(self, data):
        """Append an 8-byte aligned block, return its address."""
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def write_dataset(self, arr):
        arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        raw = arr.tobytes()
        data_addr = self.alloc(raw) if raw else UNDEF
        msgs = [
            _message(0x0001, _dataspace_message(arr.shape)),
            _message(0x0003, _datatype_message(arr.dtype), flags=1),
            # fill value v2: late allocation, written "if set", default fill value (defined, size 0)
            _message(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0)),
            _message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, len(raw))),  # contiguous layout v3
        ]
        return self.alloc(_object_header(msgs))

    def write_group(self, group):
        """Returns (object header address, b-tree address, heap address)."""
        children = {}
        for name, child in group.items():
            if name == "@attrs":
                continue
            if isinstance(child, dict):
                children[name] = ("group",) + self.write_group(child)
            else:
                children[name] = ("dataset", self.write_dataset(child))
        names = sorted(children, key=lambda s: s.encode("utf-8"))
        if len(names) > 2 * LEAF_K:
            raise ValueError(f"a group may hold at most {2 * LEAF_K} links in this writer")
        # local heap: "" at offset 0, then the link names (null terminated, 8-byte aligned), one free block
        heap_data = bytearray(8)
        offsets = {}
        for name in names:
            offsets[name] = len(heap_data)
            heap_data += _pad8(name.encode("utf-8") + b"\0")
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", H5HL_FREE_NULL, 16)  # free block: next = none, size = 16
        heap_data_addr = self.alloc(bytes(heap_data))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap_data), free_off, heap_data_addr))
        # symbol-table node (full size: 2 * LEAF_K entries)
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(names)))
        for name in names:
            kind = children[name]
            if kind[0] == "group":
                snod += struct.pack("<QQII", offsets[name], kind[1], 1, 0) + struct.pack("<QQ", kind[2], kind[3])
            else:
                snod += struct.pack("<QQII", offsets[name], kind[1], 0, 0) + b"\0" * 16
        snod += b"\0" * (8 + 2 * LEAF_K * 40 - len(snod))
        # v1 B-tree, one leaf node (full size: 2 * INTERNAL_K entries)
        if names:
            snod_addr = self.alloc(bytes(snod))
            tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF))
            tree += struct.pack("<QQQ", 0, snod_addr, offsets[names[-1]])
        else:
            tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 0, UNDEF, UNDEF))
        tree += b"\0" * (24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8 - len(tree))
        btree_addr = self.alloc(bytes(tree))
        msgs = [_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]
        for key, value in (group.get("@attrs") or {}).items():
            msgs.append(_attribute_message(key, value))
        return self.alloc(_object_header(msgs)), btree_addr, heap_addr

    def finish(self, root):
        oh, bt, hp = self.write_group(root)
        self.buf += b"\0" * (-len(self.buf) % 8)
        sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0)
        sb += struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", bt, hp)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_hdf5(path, root):
    """Write the nested-dict tree ``root`` (see the module docstring) as an HDF5 file."""
    data = _Writer().finish(root)
    with open(path, "wb") as f:
        f.write(data)


# ----------------------------------------------------------------------------------------------- reader
class _Reader:
    def __init__(self, data):
        self.d = data
        if data[:8] != SIGNATURE:
            raise ValueError("not an HDF5 file (signature at offset 0 expected)")
        if data[8] != 0 or data[13] != 8 or data[14] != 8:
            raise NotImplementedError("only version 0 superblocks with 8-byte offsets / lengths are read")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", data, 16)
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", data, 24)
        _name_off, self.root_oh, _cache, _res = struct.unpack_from("<QQII", data, 56)

    # -- object headers
    def messages(self, addr):
        d = self.d
        version, _r, nmsg, _ref, size = struct.unpack_from("<BBHII", d, addr)
        if version != 1:
            raise NotImplementedError("only version 1 object headers are read")
        out = []
        blocks = [(addr + 16, size)]
        while blocks and len(out) < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", d, pos)
                body = d[pos + 8 : pos + 8 + msize]
                if mtype == 0x0010:  # continuation
                    blocks.append(struct.unpack_from("<QQ", body, 0))
                out.append((mtype, body))
                pos += 8 + msize
        return out

    def _datatype(self, body):
        cls = body[0] & 0x0F
        bits = body[1:4]
        (size,) = struct.unpack_from("<I", body, 4)
        if cls == 0:
            return np.dtype(("<i" if bits[0] & 0x08 else "<u") + str(size)), None
        if cls == 1:
            return np.dtype("<f" + str(size)), None
        if cls == 3:
            return np.dtype(f"|S{size}"), None
        if cls == 9 and (bits[0] & 0x0F) == 1:
            return None, "vlen_str"
        raise NotImplementedError(f"datatype class {cls}")

    @staticmethod
    def _dataspace(body):
        version, rank, flags = body[0], body[1], body[2]
        off = 8 if version == 1 else 4
        return tuple(struct.unpack_from("<Q", body, off + 8 * i)[0] for i in range(rank))

    def _vlen_strings(self, raw, count):
        out = []
        for i in range(count):
            length, gaddr, index = struct.unpack_from("<IQI", raw, 16 * i)
            out.append(self._global_heap_object(gaddr, index)[:length].decode("utf-8"))
        return out

    def _global_heap_object(self, addr, index):
        d = self.d
        if d[addr : addr + 4] != b"GCOL":
            raise ValueError("bad global heap collection")
        (size,) = struct.unpack_from("<Q", d, addr + 8)
        pos, end = addr + 16, addr + size
        while pos + 16 <= end:
            idx, _ref, _res, osize = struct.unpack_from("<HHIQ", d, pos)
            if idx == index:
                return d[pos + 16 : pos + 16 + osize]
            if idx == 0:
                break
            pos += 16 + osize + (-osize % 8)
        raise KeyError(index)

    def _attribute(self, body):
        version, _r, nsize, dsize, ssize = struct.unpack_from("<BBHHH", body, 0)
        if version not in (1, 2, 3):
            raise NotImplementedError(f"attribute message version {version}")
        pos = 8 + (1 if version == 3 else 0)
        pad = (lambda n: n + (-n % 8)) if version == 1 else (lambda n: n)
        name = body[pos : pos + nsize].split(b"\0")[0].decode("utf-8")
        pos += pad(nsize)
        dtype, special = self._datatype(body[pos : pos + dsize])
        pos += pad(dsize)
        shape = self._dataspace(body[pos : pos + ssize])
        pos += pad(ssize)
        count = int(np.prod(shape)) if shape else 1
        if special == "vlen_str":
            vals = self._vlen_strings(body[pos:], count)
            value = vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)
        else:
            a = np.frombuffer(body, dtype=dtype, count=count, offset=pos).reshape(shape)
            value = a[()] if not shape else a.copy()
            if isinstance(value, np.bytes_):
                value = bytes(value)
        return name, value

    # -- groups
    def _heap_name(self, heap_addr, offset):
        d = self.d
        if d[heap_addr : heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap")
        (data_addr,) = struct.unpack_from("<Q", d, heap_addr + 24)
        start = data_addr + offset
        return d[start : d.index(b"\0", start)].decode("utf-8")

    def _btree_entries(self, addr, heap_addr):
        d = self.d
        if d[addr : addr + 4] != b"TREE":
            raise ValueError("bad B-tree node")
        _ntype, level, used = struct.unpack_from("<BBH", d, addr + 4)
        out = []
        for i in range(used):
            (child,) = struct.unpack_from("<Q", d, addr + 24 + 8 + 16 * i)
            if level > 0:
                out += self._btree_entries(child, heap_addr)
                continue
            if d[child : child + 4] != b"SNOD":
                raise ValueError("bad symbol table node")
            (nsym,) = struct.unpack_from("<H", d, child + 6)
            for j in range(nsym):
                name_off, oh = struct.unpack_from("<QQ", d, child + 8 + 40 * j)
                out.append((self._heap_name(heap_addr, name_off), oh))
        return out

    def read_object(self, addr):
        msgs = self.messages(addr)
        attrs = dict(self._attribute(b) for t, b in msgs if t == 0x000C)
        symtab = [b for t, b in msgs if t == 0x0011]
        if symtab:
            btree, heap = struct.unpack_from("<QQ", symtab[0], 0)
            node = {"@attrs": attrs} if attrs else {}
            for name, oh in self._btree_entries(btree, heap):
                node[name] = self.read_object(oh)
            return node
        shape = dtype = layout = None
        for t, b in msgs:
            if t == 0x0001:
                shape = self._dataspace(b)
            elif t == 0x0003:
                dtype, special = self._datatype(b)
                if special:
                    raise NotImplementedError("variable-length datasets")
            elif t == 0x0008:
                layout = b
        if shape is None or dtype is None or layout is None:
            raise ValueError("object is neither an old-style group nor a dataset")
        if layout[0] != 3 or layout[1] != 1:
            raise NotImplementedError("only contiguous (layout version 3) datasets are read")
        data_addr, size = struct.unpack_from("<QQ", layout, 2)
        count = int(np.prod(shape)) if shape else 1
        if data_addr == UNDEF:
            return np.zeros(shape, dtype=dtype)
        return np.frombuffer(self.d, dtype=dtype, count=count, offset=data_addr).reshape(shape).copy()


def read_hdf5(path):
    """Read an HDF5 file of the subset described in the module docstring into the nested-dict model."""
    with open(path, "rb") as f:
        data = f.read()
    r = _Reader(data)
    return r.read_object(r.root_oh)
