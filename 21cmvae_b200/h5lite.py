"""Minimal HDF5 reader/writer for Keras-2.x model files (no h5py needed).

The reference loads its weights with ``tf.keras.models.load_model``
(/root/reference/VeryAccurateEmulator/emulator.py:333-337) from classic
HDF5 files written by Keras 2.7 (superblock v0, symbol-table groups, v1
object headers, contiguous little-endian datasets, variable-length string
attributes in global heaps).  ``h5py`` is the primary loader when it is
importable (see ``keras_h5.py``); this module is the dependency-free
fallback and also provides the writer used by ``DirectEmulator.save`` and
the test fixtures.

Only the subset of the HDF5 File Format Specification (v1 structures) that
those files use is implemented:

* superblock v0, 8-byte offsets/lengths
* object header v1 with continuation blocks
* messages: dataspace (v1/v2), datatype (fixed, float, string, vlen-string),
  layout v3 (compact / contiguous), attribute v1, symbol table, continuation
* group B-tree v1 + symbol-table nodes + local heaps
* global heap collections (vlen strings)

Anything else raises ``H5LiteError`` rather than guessing.
"""

from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5LiteError(IOError):
    """Raised for unsupported or corrupt files (IOError like the reference's load_model)."""


# --------------------------------------------------------------------------
# Reader
# --------------------------------------------------------------------------


class _Datatype:
    __slots__ = ("cls", "size", "np_dtype", "is_vlen_str", "str_pad")

    def __init__(self, cls, size, np_dtype=None, is_vlen_str=False, str_pad=0):
        self.cls = cls
        self.size = size
        self.np_dtype = np_dtype
        self.is_vlen_str = is_vlen_str
        self.str_pad = str_pad


class Dataset:
    """A contiguous (or compact) dataset; ``read()`` returns a numpy array."""

    def __init__(self, f: "File", name: str, shape, dtype: _Datatype, data_addr, data_size, compact=None, attrs=None):
        self._f = f
        self.name = name
        self.shape = tuple(shape)
        self._dtype = dtype
        self._addr = data_addr
        self._size = data_size
        self._compact = compact
        self.attrs = attrs or {}

    @property
    def dtype(self):
        return self._dtype.np_dtype

    def read(self) -> np.ndarray:
        if self._dtype.np_dtype is None:
            raise H5LiteError(f"dataset {self.name}: unsupported datatype class {self._dtype.cls}")
        n = int(np.prod(self.shape)) if self.shape else 1
        nbytes = n * self._dtype.size
        if self._compact is not None:
            raw = self._compact[:nbytes]
        else:
            if self._addr == _UNDEF:
                return np.zeros(self.shape, dtype=self._dtype.np_dtype)
            raw = self._f._buf[self._addr : self._addr + nbytes]
        if len(raw) < nbytes:
            raise H5LiteError(f"dataset {self.name}: truncated data")
        return np.frombuffer(raw, dtype=self._dtype.np_dtype, count=n).reshape(self.shape).copy()

    def __getitem__(self, key):
        return self.read()[key]


class Group:
    def __init__(self, f: "File", name: str, links: Dict[str, int], attrs):
        self._f = f
        self.name = name
        self._links = links
        self.attrs = attrs

    def keys(self):
        return list(self._links.keys())

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __getitem__(self, path: str):
        node: Union[Group, Dataset] = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(path)
            node = node._f._load_object(node._links[part], (node.name.rstrip("/") + "/" + part))
        return node

    def visit_datasets(self, prefix=""):
        out = []
        for k in self.keys():
            obj = self[k]
            p = prefix + "/" + k
            if isinstance(obj, Group):
                out.extend(obj.visit_datasets(p))
            else:
                out.append((p, obj))
        return out


class File(Group):
    """Read-only view of a classic HDF5 file held fully in memory."""

    def __init__(self, path: str):
        try:
            with open(path, "rb") as fh:
                self._buf = fh.read()
        except OSError as e:  # keep the reference's IOError contract
            raise H5LiteError(f"cannot open {path}: {e}") from e
        self.path = path
        self._cache: Dict[int, object] = {}
        if self._buf[:8] != _SIG:
            raise H5LiteError(f"{path}: not an HDF5 file (bad signature)")
        ver = self._buf[8]
        if ver != 0:
            raise H5LiteError(f"{path}: superblock version {ver} unsupported (need 0)")
        so, sl = self._buf[13], self._buf[14]
        if (so, sl) != (8, 8):
            raise H5LiteError(f"{path}: offsets/lengths of {so}/{sl} bytes unsupported")
        self._base = struct.unpack_from("<Q", self._buf, 24)[0]
        if self._base != 0:
            raise H5LiteError("non-zero base address unsupported")
        # root symbol table entry at byte 56: name off, obj hdr addr, cache type, reserved, scratch
        root_hdr = struct.unpack_from("<Q", self._buf, 56 + 8)[0]
        root = self._load_object(root_hdr, "/")
        if not isinstance(root, Group):
            raise H5LiteError("root object is not a group")
        super().__init__(self, "/", root._links, root.attrs)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    # ---- low-level helpers -------------------------------------------------
    def _u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self._buf, off)

    def _messages(self, addr) -> List[Tuple[int, int, bytes]]:
        ver = self._buf[addr]
        if ver != 1:
            raise H5LiteError(f"object header version {ver} at {addr} unsupported")
        nmsgs = self._u("H", addr + 2)[0]
        hdr_size = self._u("I", addr + 8)[0]
        blocks = [(addr + 16, hdr_size)]
        msgs = []
        bi = 0
        while bi < len(blocks) and len(msgs) < nmsgs:
            off, size = blocks[bi]
            end = off + size
            while off + 8 <= end and len(msgs) < nmsgs:
                mtype, msize, mflags = self._u("HHB", off)
                body = self._buf[off + 8 : off + 8 + msize]
                if mtype == 0x0010:
                    c_off, c_len = struct.unpack_from("<QQ", body, 0)
                    blocks.append((c_off, c_len))
                msgs.append((mtype, mflags, body))
                off += 8 + msize
            bi += 1
        return msgs

    def _parse_datatype(self, b: bytes, off=0) -> Tuple[_Datatype, int]:
        cv = b[off]
        cls = cv & 0x0F
        bf0, bf1, bf2 = b[off + 1], b[off + 2], b[off + 3]
        size = struct.unpack_from("<I", b, off + 4)[0]
        if cls == 0:  # fixed point
            signed = bool(bf0 & 0x08)
            order = ">" if (bf0 & 1) else "<"
            dt = np.dtype(f"{order}{'i' if signed else 'u'}{size}")
            return _Datatype(cls, size, dt), 8 + 4
        if cls == 1:  # float
            order = ">" if (bf0 & 1) else "<"
            dt = np.dtype(f"{order}f{size}")
            return _Datatype(cls, size, dt), 8 + 12
        if cls == 3:  # fixed string
            return _Datatype(cls, size, np.dtype(f"S{size}"), str_pad=bf0 & 0x0F), 8
        if cls == 9:  # variable length
            vtype = bf0 & 0x0F
            base, blen = self._parse_datatype(b, off + 8)
            return _Datatype(cls, size, None, is_vlen_str=(vtype == 1)), 8 + blen
        return _Datatype(cls, size, None), 8

    @staticmethod
    def _parse_dataspace(b: bytes):
        ver = b[0]
        rank = b[1]
        flags = b[2]
        if ver == 1:
            off = 8
        elif ver == 2:
            off = 4
        else:
            raise H5LiteError(f"dataspace version {ver} unsupported")
        dims = struct.unpack_from("<" + "Q" * rank, b, off) if rank else ()
        return tuple(dims)

    def _global_heap_object(self, gaddr, index) -> bytes:
        if self._buf[gaddr : gaddr + 4] != b"GCOL":
            raise H5LiteError("bad global heap signature")
        csize = self._u("Q", gaddr + 8)[0]
        off = gaddr + 16
        end = gaddr + csize
        while off + 16 <= end:
            idx, _ref, _res, osize = struct.unpack_from("<HHIQ", self._buf, off)
            if idx == 0:
                break
            if idx == index:
                return self._buf[off + 16 : off + 16 + osize]
            off += 16 + ((osize + 7) & ~7)
        raise H5LiteError(f"global heap object {index} not found")

    def _decode_attr_value(self, dt: _Datatype, shape, raw: bytes):
        n = int(np.prod(shape)) if shape else 1
        if dt.is_vlen_str:
            vals = []
            for i in range(n):
                ln, gaddr, gidx = struct.unpack_from("<IQI", raw, i * 16)
                s = self._global_heap_object(gaddr, gidx)[:ln] if ln else b""
                vals.append(s.decode("utf-8", "replace"))
            if not shape:
                return vals[0]
            return np.array(vals, dtype=object).reshape(shape)
        if dt.np_dtype is None:
            return None
        arr = np.frombuffer(raw[: n * dt.size], dtype=dt.np_dtype, count=n)
        if dt.cls == 3:
            arr = np.array([x.rstrip(b"\x00 ") for x in arr], dtype=object)
            if not shape:
                return arr[0]
            return arr.reshape(shape)
        if not shape:
            return arr[0]
        return arr.reshape(shape).copy()

    def _parse_attribute(self, b: bytes):
        ver = b[0]
        if ver not in (1, 2, 3):
            raise H5LiteError(f"attribute version {ver} unsupported")
        name_sz, dt_sz, ds_sz = struct.unpack_from("<HHH", b, 2)
        off = 8
        if ver == 3:
            off = 9  # + name charset byte
        pad = (lambda x: (x + 7) & ~7) if ver == 1 else (lambda x: x)
        name = b[off : off + name_sz].split(b"\x00")[0].decode("utf-8")
        off += pad(name_sz)
        dt, _ = self._parse_datatype(b, off)
        off += pad(dt_sz)
        shape = self._parse_dataspace(b[off : off + ds_sz]) if ds_sz >= 4 else ()
        off += pad(ds_sz)
        return name, self._decode_attr_value(dt, shape, b[off:])

    def _group_links(self, btree_addr, heap_addr) -> Dict[str, int]:
        if self._buf[heap_addr : heap_addr + 4] != b"HEAP":
            raise H5LiteError("bad local heap signature")
        heap_data = self._u("Q", heap_addr + 24)[0]
        links: Dict[str, int] = {}

        def name_at(off):
            s = heap_data + off
            e = self._buf.index(b"\x00", s)
            return self._buf[s:e].decode("utf-8")

        def walk(addr):
            sig = self._buf[addr : addr + 4]
            if sig == b"TREE":
                ntype, level, nent = self._u("BBH", addr + 4)
                if ntype != 0:
                    raise H5LiteError("non-group B-tree in group")
                off = addr + 24  # after sig(4) type(1) level(1) entries(2) left(8) right(8)
                off += 8  # key 0
                for _ in range(nent):
                    child = self._u("Q", off)[0]
                    off += 16  # child + next key
                    walk(child)
            elif sig == b"SNOD":
                nsym = self._u("H", addr + 6)[0]
                off = addr + 8
                for _ in range(nsym):
                    name_off, hdr = self._u("QQ", off)
                    links[name_at(name_off)] = hdr
                    off += 40
            else:
                raise H5LiteError(f"unexpected node signature {sig!r}")

        walk(btree_addr)
        return links

    def _load_object(self, addr: int, name: str):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr)
        attrs = {}
        stab = None
        shape = None
        dtype = None
        layout = None
        for mtype, _flags, body in msgs:
            if mtype == 0x000C:
                k, v = self._parse_attribute(body)
                attrs[k] = v
            elif mtype == 0x0011:
                stab = struct.unpack_from("<QQ", body, 0)
            elif mtype == 0x0001:
                shape = self._parse_dataspace(body)
            elif mtype == 0x0003:
                dtype, _ = self._parse_datatype(body)
            elif mtype == 0x0008:
                layout = body
            elif mtype == 0x000B:
                raise H5LiteError(f"{name}: filtered datasets unsupported")
        if stab is not None:
            obj = Group(self, name, self._group_links(*stab), attrs)
        elif layout is not None and dtype is not None:
            ver = layout[0]
            if ver != 3:
                raise H5LiteError(f"{name}: layout version {ver} unsupported")
            lclass = layout[1]
            if lclass == 1:
                daddr, dsize = struct.unpack_from("<QQ", layout, 2)
                obj = Dataset(self, name, shape or (), dtype, daddr, dsize, attrs=attrs)
            elif lclass == 0:
                csz = struct.unpack_from("<H", layout, 2)[0]
                obj = Dataset(self, name, shape or (), dtype, 0, csz, compact=bytes(layout[4 : 4 + csz]), attrs=attrs)
            else:
                raise H5LiteError(f"{name}: chunked datasets unsupported")
        else:
            raise H5LiteError(f"{name}: object is neither group nor dataset")
        self._cache[addr] = obj
        return obj


# --------------------------------------------------------------------------
# Writer (same classic layout; enough for Keras-2.x style model files)
# --------------------------------------------------------------------------


def _pad8(n):
    return (n + 7) & ~7


class _WNode:
    def __init__(self):
        self.children: Dict[str, "_WNode"] = {}
        self.attrs: Dict[str, object] = {}
        self.data: Optional[np.ndarray] = None


class Writer:
    """Build a classic HDF5 file in memory: groups, contiguous datasets, attributes.

    String attributes are written as variable-length UTF-8 strings (global
    heap), string arrays likewise -- the same encoding Keras/h5py 3 produce
    for ``model_config``, ``layer_names`` and ``weight_names``.
    """

    LEAF_K = 4  # symbol-table node holds up to 2*K entries

    def __init__(self):
        self.root = _WNode()

    def _node(self, path, create=True) -> _WNode:
        n = self.root
        for part in [p for p in path.split("/") if p]:
            if part not in n.children:
                if not create:
                    raise KeyError(path)
                n.children[part] = _WNode()
            n = n.children[part]
        return n

    def create_group(self, path):
        self._node(path)

    def create_dataset(self, path, data):
        node = self._node(path)
        arr = np.asarray(data)
        if arr.ndim:  # (np.ascontiguousarray would turn a 0-d array -- Keras' `Adam/iter:0` -- into shape (1,))
            arr = np.ascontiguousarray(arr)
        if arr.dtype.kind not in "fiu":
            raise H5LiteError("only numeric datasets supported")
        node.data = arr.astype(arr.dtype.newbyteorder("<"))

    def set_attr(self, path, name, value):
        self._node(path).attrs[name] = value

    # ---- serialisation -----------------------------------------------------
    def tobytes(self) -> bytes:
        buf = bytearray(b"\x00" * 96)  # superblock v0 (56) + root symbol table entry (40)
        gheap_items: List[bytes] = []

        def alloc(b: bytes, align=8) -> int:
            while len(buf) % align:
                buf.append(0)
            a = len(buf)
            buf.extend(b)
            return a

        def _is_str_list(v):
            return (isinstance(v, (list, tuple)) and len(v) > 0 and isinstance(v[0], (str, bytes))) or (
                isinstance(v, np.ndarray) and v.dtype == object
            )

        # one global heap collection for all vlen strings, in the order write_node consumes them
        # (children before the node's own attributes)
        def collect_in_write_order(node: _WNode):
            if node.data is None:
                for n in sorted(node.children.keys()):
                    collect_in_write_order(node.children[n])
            for v in node.attrs.values():
                if isinstance(v, str):
                    gheap_items.append(v.encode("utf-8"))
                elif _is_str_list(v):
                    for s in np.asarray(v, dtype=object).ravel():
                        gheap_items.append(s.encode("utf-8") if isinstance(s, str) else s)

        collect_in_write_order(self.root)
        gheap_addr = 0
        if gheap_items:
            body = bytearray()
            for i, s in enumerate(gheap_items, start=1):
                body += struct.pack("<HHIQ", i, 1, 0, len(s)) + s + b"\x00" * (_pad8(len(s)) - len(s))
            total = max(_pad8(16 + len(body) + 16), 4096)
            free = total - 16 - len(body) - 16
            body += struct.pack("<HHIQ", 0, 0, 0, free) + b"\x00" * free
            hdr = b"GCOL" + bytes([1, 0, 0, 0]) + struct.pack("<Q", total)
            gheap_addr = alloc(hdr + bytes(body))
        str_counter = [0]

        def vlen_ref(s: Union[str, bytes]) -> bytes:
            b = s.encode("utf-8") if isinstance(s, str) else s
            i = str_counter[0]
            assert gheap_items[i] == b, "string order mismatch"
            str_counter[0] += 1
            return struct.pack("<IQI", len(b), gaddr_final, i + 1)

        gaddr_final = gheap_addr

        def dt_msg_for(arr_dtype: np.dtype) -> bytes:
            size = arr_dtype.itemsize
            if arr_dtype.kind == "f":
                if size == 4:
                    props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
                    bits = (0x20, 31, 0)
                elif size == 8:
                    props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
                    bits = (0x20, 63, 0)
                else:
                    raise H5LiteError("float size unsupported")
                return bytes([0x11, bits[0], bits[1], bits[2]]) + struct.pack("<I", size) + props
            signed = 0x08 if arr_dtype.kind == "i" else 0
            return bytes([0x10, signed, 0, 0]) + struct.pack("<I", size) + struct.pack("<HH", 0, size * 8)

        vlen_str_dt = (
            bytes([0x19, 0x01, 0x01, 0x00]) + struct.pack("<I", 16) + bytes([0x13, 0x11, 0, 0]) + struct.pack("<I", 1)
        )  # vlen string, utf-8, base = 1-byte string

        def ds_msg(shape) -> bytes:
            rank = len(shape)
            return bytes([1, rank, 0, 0, 0, 0, 0, 0]) + b"".join(struct.pack("<Q", d) for d in shape)

        def attr_msg(name: str, value) -> bytes:
            nm = name.encode("utf-8") + b"\x00"
            if isinstance(value, str):
                dt, shape, data = vlen_str_dt, (), vlen_ref(value)
            elif _is_str_list(value):
                vals = list(np.asarray(value, dtype=object).ravel())
                dt, shape, data = vlen_str_dt, (len(vals),), b"".join(vlen_ref(s) for s in vals)
            else:
                arr = np.ascontiguousarray(value)
                if arr.dtype.kind not in "fiu":
                    raise H5LiteError(f"attribute {name}: unsupported type {arr.dtype}")
                arr = arr.astype(arr.dtype.newbyteorder("<"))
                dt, shape, data = dt_msg_for(arr.dtype), arr.shape, arr.tobytes()
            ds = ds_msg(shape)
            body = bytes([1, 0]) + struct.pack("<HHH", len(nm), len(dt), len(ds))
            body += nm + b"\x00" * (_pad8(len(nm)) - len(nm))
            body += dt + b"\x00" * (_pad8(len(dt)) - len(dt))
            body += ds + b"\x00" * (_pad8(len(ds)) - len(ds))
            body += data
            return body

        def obj_header(msgs: List[Tuple[int, bytes]]) -> bytes:
            out = bytearray()
            for mtype, body in msgs:
                body = body + b"\x00" * (_pad8(len(body)) - len(body))
                out += struct.pack("<HHB3x", mtype, len(body), 0) + body
            return bytes([1, 0]) + struct.pack("<HII", len(msgs), 1, len(out)) + b"\x00" * 4 + bytes(out)

        def write_node(node: _WNode) -> int:
            """Returns object header address."""
            if node.data is not None:
                arr = node.data
                daddr = alloc(arr.tobytes()) if arr.size else _UNDEF
                msgs = [
                    (0x0001, ds_msg(arr.shape)),
                    (0x0003, dt_msg_for(arr.dtype)),
                    (0x0008, bytes([3, 1]) + struct.pack("<QQ", daddr, arr.nbytes)),
                ]
                msgs += [(0x000C, attr_msg(k, v)) for k, v in node.attrs.items()]
                return alloc(obj_header(msgs))
            # group: children first
            names = sorted(node.children.keys())
            child_addrs = {n: write_node(node.children[n]) for n in names}
            # local heap: offset 0 is the empty string
            heap = bytearray(b"\x00" * 8)
            name_off = {}
            for n in names:
                name_off[n] = len(heap)
                nb = n.encode("utf-8") + b"\x00"
                heap += nb + b"\x00" * (_pad8(len(nb)) - len(nb))
            heap_size = max(_pad8(len(heap)) + 16, 64)
            free_off = len(heap)
            heap += b"\x00" * (heap_size - len(heap))
            # free block: next-free offset (1 = none), size
            struct.pack_into("<QQ", heap, free_off, 1, heap_size - free_off)
            heap_data_addr = alloc(bytes(heap))
            heap_addr = alloc(b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", heap_size, free_off, heap_data_addr))
            # symbol table nodes (leaves), 2K entries each
            cap = 2 * self.LEAF_K
            leaves = []
            for i in range(0, max(len(names), 1), cap):
                part = names[i : i + cap]
                body = bytearray(b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(part)))
                for n in part:
                    body += struct.pack("<QQII16x", name_off[n], child_addrs[n], 0, 0)
                body += b"\x00" * (40 * (cap - len(part)))
                leaves.append((alloc(bytes(body)), part))
            if len(leaves) > 2 * 16:
                raise H5LiteError("too many links in one group for single-level B-tree")
            bt = bytearray(b"TREE" + bytes([0, 0]) + struct.pack("<H", len(leaves)) + struct.pack("<QQ", _UNDEF, _UNDEF))
            bt += struct.pack("<Q", 0)  # key 0: empty string
            for addr, part in leaves:
                last_key = name_off[part[-1]] if part else 0
                bt += struct.pack("<QQ", addr, last_key)
            bt += b"\x00" * (24 + 8 + 16 * 32 - len(bt))
            bt_addr = alloc(bytes(bt))
            msgs = [(0x0011, struct.pack("<QQ", bt_addr, heap_addr))]
            msgs += [(0x000C, attr_msg(k, v)) for k, v in node.attrs.items()]
            node._stab = (bt_addr, heap_addr)  # type: ignore[attr-defined]
            return alloc(obj_header(msgs))

        root_addr = write_node(self.root)
        eof = len(buf)
        sb = bytearray()
        sb += _SIG
        sb += bytes([0, 0, 0, 0, 0, 8, 8, 0])
        sb += struct.pack("<HH", self.LEAF_K, 16)
        sb += struct.pack("<I", 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
        bt_addr, heap_addr = self.root._stab  # type: ignore[attr-defined]
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", bt_addr, heap_addr)
        assert len(sb) == 96, len(sb)
        buf[0:96] = sb
        return bytes(buf)

    def save(self, path: str):
        data = self.tobytes()
        with open(path, "wb") as fh:
            fh.write(data)
