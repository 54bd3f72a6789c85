"""Minimal pure-Python HDF5 codec for the Keras ``.h5`` weight contract.

h5py / libhdf5 are not available in this environment, so the subset Keras 2.4 ``save_weights`` produces
through h5py's defaults is implemented by hand from the HDF5 file-format specification:

  * superblock version 0, 8-byte offsets / lengths;
  * "old style" groups: symbol-table message -> v1 B-tree (one leaf node) + local heap + SNOD symbol nodes;
  * version-1 object headers (continuation blocks are followed when reading);
  * contiguous little-endian float32/float64 datasets, simple dataspaces of rank 0..4;
  * attributes: fixed-length byte strings (scalar and 1-D arrays); the reader also accepts variable-length
    strings stored in global heap collections (files re-saved by newer h5py).

File layout mirrored from the reference loader (common/utils/weight_io.py:76-263): root attributes
``layer_names`` / ``backend`` / ``keras_version``; one group per top-level layer; group attribute
``weight_names`` (ordered); one dataset per weight at ``<group>/<weight_name>`` (names contain "/", hence
nested groups).  Matching is by group name, then by POSITION — inner names are written as Keras would name
them but never compared on load.  An optional ``model_weights`` wrapper group is descended into
(weight_io.py:119-120).

VALIDATION LIMIT: there is no foreign HDF5 implementation in the image to cross-check against; the codec is
pinned by self round-trips and by structural checks against the published format only.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import weights as W

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K = 4, 16


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


# =================================================================================================
# Writer
# =================================================================================================
class _Node:
    """In-memory tree: a group (children dict, attrs) or a dataset (array)."""

    def __init__(self, array: Optional[np.ndarray] = None):
        self.children: Dict[str, "_Node"] = {}
        self.attrs: List[Tuple[str, object]] = []
        self.array = array

    def ensure_group(self, path: List[str]) -> "_Node":
        node = self
        for p in path:
            node = node.children.setdefault(p, _Node())
        return node


def _dt_float32() -> bytes:
    return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)


def _dt_string(n: int) -> bytes:
    return struct.pack("<BBBBI", 0x13, 0x01, 0x00, 0x00, n)          # null-padded, ASCII, fixed length n


def _dataspace(shape: Tuple[int, ...]) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _attr_message(name: str, value) -> bytes:
    nm = name.encode() + b"\x00"
    if isinstance(value, (bytes, str)):
        raw = value.encode() if isinstance(value, str) else value
        dt, ds, data = _dt_string(max(1, len(raw))), _dataspace(()), raw.ljust(max(1, len(raw)), b"\x00")
    else:                                            # 1-D array of byte strings
        items = [v.encode() if isinstance(v, str) else bytes(v) for v in value]
        width = max([1] + [len(v) for v in items])
        dt, ds = _dt_string(width), _dataspace((len(items),))
        data = b"".join(v.ljust(width, b"\x00") for v in items)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + data
    if len(body) >= 65536 - 8:
        raise ValueError(f"attribute {name!r} exceeds the 64 KB object-header message limit")
    return _message(0x000C, body)


def _object_header(messages: List[bytes]) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class _Writer:
    def __init__(self):
        self.buf = bytearray(b"\x00" * 96)           # superblock placeholder

    def alloc(self, data: bytes) -> int:
        while len(self.buf) % 8:
            self.buf += b"\x00"
        addr = len(self.buf)
        self.buf += data
        return addr

    def write_dataset(self, arr: np.ndarray) -> int:
        a = np.ascontiguousarray(arr, dtype="<f4")
        data_addr = self.alloc(a.tobytes()) if a.size else UNDEF
        msgs = [
            _message(0x0001, _dataspace(a.shape)),
            _message(0x0003, _dt_float32(), flags=1),
            _message(0x0005, struct.pack("<BBBB", 2, 2, 2, 0)),                       # fill value v2, undefined
            _message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, a.nbytes)),        # layout v3, contiguous
        ]
        return self.alloc(_object_header(msgs))

    def write_group(self, node: _Node) -> Tuple[int, int, int]:
        """Returns (object header address, B-tree address, heap address)."""
        names = sorted(node.children)                # symbol nodes are ordered by strcmp (bytewise)
        names.sort(key=lambda s: s.encode())
        child_addr = {}
        for n in names:
            ch = node.children[n]
            child_addr[n] = self.write_dataset(ch.array) if ch.array is not None else self.write_group(ch)
        # local heap: offset 0 holds the empty string, then the link names
        heap = bytearray(b"\x00" * 8)
        name_off = {}
        for n in names:
            name_off[n] = len(heap)
            heap += _pad8(n.encode() + b"\x00")
        heap_data_addr = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, heap_data_addr))
        # symbol nodes of at most 2*LEAF_K entries
        snods, keys = [], [0]
        for i in range(0, len(names), 2 * LEAF_K):
            chunk = names[i:i + 2 * LEAF_K]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
            for n in chunk:
                ca = child_addr[n]
                if isinstance(ca, tuple):            # group: cache the B-tree / heap addresses (cache type 1)
                    body += struct.pack("<QQII", name_off[n], ca[0], 1, 0) + struct.pack("<QQ", ca[1], ca[2])
                else:
                    body += struct.pack("<QQII16x", name_off[n], ca, 0, 0)
            body += b"\x00" * (40 * (2 * LEAF_K - len(chunk)))
            snods.append(self.alloc(body))
            keys.append(name_off[chunk[-1]])
        if len(snods) > 2 * INTERNAL_K:
            raise ValueError("group has too many links for a single B-tree leaf node")
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        for i in range(2 * INTERNAL_K):
            tree += struct.pack("<Q", keys[i] if i < len(keys) else 0)
            tree += struct.pack("<Q", snods[i] if i < len(snods) else UNDEF)
        tree += struct.pack("<Q", keys[2 * INTERNAL_K] if len(keys) > 2 * INTERNAL_K else 0)
        btree_addr = self.alloc(tree)
        msgs = [_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]
        msgs += [_attr_message(k, v) for k, v in node.attrs]
        ohdr = self.alloc(_object_header(msgs))
        return (ohdr, btree_addr, heap_addr)

    def finish(self, root: Tuple[int, int, int]) -> bytes:
        eof = len(self.buf)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root[0], 1, 0) + struct.pack("<QQ", root[1], root[2])
        assert len(sb) == 96
        self.buf[0:96] = sb
        return bytes(self.buf)


def write_h5(path: str, root: _Node) -> None:
    w = _Writer()
    data = w.finish(w.write_group(root))
    with open(path, "wb") as f:
        f.write(data)


def save_keras_weights(path: str, spec, w: Dict[W.WeightKey, np.ndarray], keras_version: str = "2.4.0") -> None:
    """model.save_weights("*.h5") layout (train.py:706, :719)."""
    inv = W.inventory(spec)
    root = _Node()
    root.attrs = [("layer_names", list(inv)), ("backend", b"tensorflow"), ("keras_version", keras_version.encode())]
    for gname, tensors in inv.items():
        grp = root.ensure_group([gname])
        grp.attrs = [("weight_names", [n for n, _, _ in tensors])]
        for i, (wname, shape, _) in enumerate(tensors):
            a = np.asarray(w[(gname, i)], dtype=np.float32)
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"{gname}[{i}]: array shape {a.shape}, inventory {shape}")
            parts = wname.split("/")
            grp.ensure_group(parts[:-1]).children[parts[-1]] = _Node(a)
    write_h5(path, root)


# =================================================================================================
# Reader
# =================================================================================================
class H5Object:
    def __init__(self, f: "H5File", addr: int):
        self.f, self.addr = f, addr
        self.messages = f._read_object_header(addr)
        self.attrs = {}
        for t, d in self.messages:
            if t == 0x000C:
                k, v = f._parse_attribute(d)
                self.attrs[k] = v

    @property
    def is_group(self) -> bool:
        return any(t == 0x0011 for t, _ in self.messages)

    def links(self) -> Dict[str, int]:
        for t, d in self.messages:
            if t == 0x0011:
                btree, heap = struct.unpack_from("<QQ", d)
                return self.f._read_group(btree, heap)
        if any(t in (0x0002, 0x0006) for t, _ in self.messages):
            raise NotImplementedError("new-style (link message) groups are not supported; re-save with libver='earliest'")
        raise ValueError("not a group")

    def __getitem__(self, path: str) -> "H5Object":
        obj = self
        for p in [q for q in path.split("/") if q]:
            ln = obj.links()
            if p not in ln:
                raise KeyError(p)
            obj = H5Object(self.f, ln[p])
        return obj

    def __contains__(self, name: str) -> bool:
        return self.is_group and name in self.links()

    def read(self) -> np.ndarray:
        shape = dtype = None
        layout = None
        for t, d in self.messages:
            if t == 0x0001:
                shape = self.f._parse_dataspace(d)
            elif t == 0x0003:
                dtype = self.f._parse_datatype(d)[0]
            elif t == 0x0008:
                layout = d
        if shape is None or dtype is None or layout is None:
            raise ValueError("not a dataset")
        if not isinstance(dtype, np.dtype) or dtype.kind != "f":
            raise NotImplementedError(f"dataset type {dtype} is not a float type")
        n = int(np.prod(shape)) if shape else 1
        ver, cls = layout[0], layout[1]
        if ver == 3 and cls == 1:
            addr, size = struct.unpack_from("<QQ", layout, 2)
            raw = b"" if addr == UNDEF else self.f.data[addr:addr + size]
        elif ver == 3 and cls == 0:
            (size,) = struct.unpack_from("<H", layout, 2)
            raw = layout[4:4 + size]
        elif ver in (1, 2) and layout[2] == 1:                 # old contiguous layout: dims follow
            rank = layout[1]
            (addr,) = struct.unpack_from("<Q", layout, 8)
            raw = self.f.data[addr:addr + n * dtype.itemsize]
            del rank
        else:
            raise NotImplementedError("only contiguous / compact dataset layouts are supported (Keras writes contiguous)")
        if n == 0:
            return np.zeros(shape, dtype=np.float32)
        return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).astype(np.float32)


class H5File(H5Object):
    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.data = fh.read()
        d = self.data
        if d[:8] != SIGNATURE:
            raise ValueError("not an HDF5 file (bad signature)")
        ver = d[8]
        if ver not in (0, 1):
            raise NotImplementedError(f"superblock version {ver} (only 0/1 — h5py default — is supported)")
        if d[13] != 8 or d[14] != 8:
            raise NotImplementedError("only 8-byte offsets / lengths are supported")
        off = 24 + (4 if ver == 1 else 0)
        base, _fs, _eof, _drv = struct.unpack_from("<QQQQ", d, off)
        if base != 0:
            raise NotImplementedError("non-zero base address")
        _name_off, root_addr = struct.unpack_from("<QQ", d, off + 32)
        super().__init__(self, root_addr)

    # ---- object headers ---------------------------------------------------------------------------
    def _read_object_header(self, addr: int) -> List[Tuple[int, bytes]]:
        d = self.data
        if d[addr:addr + 4] == b"OHDR":
            raise NotImplementedError("version-2 object headers are not supported; re-save with libver='earliest'")
        ver, _r, nmsg, _ref, size = struct.unpack_from("<BBHII", d, addr)
        if ver != 1:
            raise ValueError(f"unsupported object header version {ver}")
        msgs: List[Tuple[int, bytes]] = []
        chunks = [(addr + 16, size)]
        while chunks and len(msgs) < nmsg:
            pos, remaining = chunks.pop(0)
            end = pos + remaining
            while pos + 8 <= end and len(msgs) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", d, pos)
                body = d[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:                      # continuation: (address, length)
                    chunks.append(struct.unpack_from("<QQ", body))
                msgs.append((mtype, body))
        return msgs

    # ---- groups -----------------------------------------------------------------------------------
    def _heap_name(self, heap_addr: int, offset: int) -> str:
        d = self.data
        if d[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        (seg,) = struct.unpack_from("<Q", d, heap_addr + 24)
        end = d.index(b"\x00", seg + offset)
        return d[seg + offset:end].decode("utf8")

    def _read_group(self, btree: int, heap: int) -> Dict[str, int]:
        out: Dict[str, int] = {}
        self._walk_btree(btree, heap, out)
        return out

    def _walk_btree(self, addr: int, heap: int, out: Dict[str, int]) -> None:
        d = self.data
        if d[addr:addr + 4] != b"TREE":
            raise ValueError("bad B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", d, addr + 4)
        if ntype != 0:
            raise ValueError("not a group B-tree")
        pos = addr + 24
        for i in range(used):
            (child,) = struct.unpack_from("<Q", d, pos + 8 + 16 * i)
            if level > 0:
                self._walk_btree(child, heap, out)
                continue
            if d[child:child + 4] != b"SNOD":
                raise ValueError("bad symbol node signature")
            (nsym,) = struct.unpack_from("<H", d, child + 6)
            for k in range(nsym):
                name_off, ohdr = struct.unpack_from("<QQ", d, child + 8 + 40 * k)
                out[self._heap_name(heap, name_off)] = ohdr

    # ---- messages ---------------------------------------------------------------------------------
    @staticmethod
    def _parse_dataspace(b: bytes) -> Tuple[int, ...]:
        ver, rank, flags = b[0], b[1], b[2]
        off = 8 if ver == 1 else 4
        return tuple(struct.unpack_from("<Q", b, off + 8 * i)[0] for i in range(rank))

    def _parse_datatype(self, b: bytes):
        """-> (descriptor, bytes consumed); descriptor is a numpy dtype, ('S', n) or ('vlen_str',)."""
        cls = b[0] & 0x0F
        bits0 = b[1]
        (size,) = struct.unpack_from("<I", b, 4)
        if cls == 1:
            order = ">" if bits0 & 1 else "<"
            return np.dtype(f"{order}f{size}"), 20
        if cls == 0:
            order = ">" if bits0 & 1 else "<"
            signed = "i" if bits0 & 0x08 else "u"
            return np.dtype(f"{order}{signed}{size}"), 12
        if cls == 3:
            return ("S", size), 8
        if cls == 9:
            if (bits0 & 0x0F) == 1:
                return ("vlen_str",), 8
            raise NotImplementedError("variable-length sequences")
        raise NotImplementedError(f"datatype class {cls}")

    def _parse_attribute(self, b: bytes):
        ver = b[0]
        if ver == 1:
            _v, _r, nsz, dsz, ssz = struct.unpack_from("<BBHHH", b)
            pos = 8
            name = b[pos:pos + nsz].split(b"\x00")[0].decode()
            pos += (nsz + 7) & ~7
            dt = b[pos:pos + dsz]; pos += (dsz + 7) & ~7
            sp = b[pos:pos + ssz]; pos += (ssz + 7) & ~7
        elif ver in (2, 3):
            _v, _f, nsz, dsz, ssz = struct.unpack_from("<BBHHH", b)
            pos = 8 + (1 if ver == 3 else 0)
            name = b[pos:pos + nsz].split(b"\x00")[0].decode(); pos += nsz
            dt = b[pos:pos + dsz]; pos += dsz
            sp = b[pos:pos + ssz]; pos += ssz
        else:
            raise NotImplementedError(f"attribute message version {ver}")
        shape = self._parse_dataspace(sp)
        desc, _ = self._parse_datatype(dt)
        n = int(np.prod(shape)) if shape else 1
        raw = b[pos:]
        if isinstance(desc, np.dtype):
            val = np.frombuffer(raw, dtype=desc, count=n).reshape(shape)
            return name, (val if shape else val.reshape(()).item())
        if desc[0] == "S":
            w = desc[1]
            items = [raw[i * w:(i + 1) * w].split(b"\x00")[0] for i in range(n)]
        else:                                            # variable-length strings -> global heap
            items = []
            for i in range(n):
                ln, gaddr, gidx = struct.unpack_from("<IQI", raw, 16 * i)
                items.append(self._global_heap_object(gaddr, gidx)[:ln])
        return name, (np.array(items, dtype=object) if shape else items[0])

    def _global_heap_object(self, addr: int, index: int) -> bytes:
        d = self.data
        if d[addr:addr + 4] != b"GCOL":
            raise ValueError("bad global heap signature")
        (csize,) = struct.unpack_from("<Q", d, addr + 8)
        pos, end = addr + 16, addr + csize
        while pos + 16 <= end:
            idx, _ref, _r, osize = struct.unpack_from("<HHIQ", d, pos)
            if idx == index:
                return d[pos + 16:pos + 16 + osize]
            if idx == 0:
                break
            pos += 16 + ((osize + 7) & ~7)
        raise KeyError(f"global heap object {index} not found")


def _decode(v) -> str:
    return v.decode("utf8") if isinstance(v, (bytes, np.bytes_)) else str(v)


def load_keras_weights(path: str, spec, skip_mismatch: bool = False, verbose: bool = True,
                       report: Optional[dict] = None) -> Dict[W.WeightKey, np.ndarray]:
    """Name-based group lookup, position-based tensor matching, shape check — weight_io.py:125-263, with its rules:
    a model layer the file does not hold is NOT an error (it keeps its current values and is listed under "not assigned any
    weights", :247-251; this is how a pre-trained file without some layers is loaded); a layer whose tensor count or a tensor
    whose shape disagrees raises ValueError unless ``skip_mismatch`` (:185-195, :219-232: then it is skipped with a warning);
    file layers the model does not have are listed as "not consumed" (:241-245).  The returned dict may therefore be partial;
    ``report`` (optional dict) receives the three lists."""
    f: H5Object = H5File(path)
    if "layer_names" not in f.attrs and "model_weights" in f:                 # weight_io.py:119-120
        f = f["model_weights"]
    if "layer_names" not in f.attrs:
        raise ValueError("no 'layer_names' attribute: not a Keras weights file")
    file_layers = [_decode(n) for n in np.atleast_1d(f.attrs["layer_names"])]
    inv = W.inventory(spec)
    out: Dict[W.WeightKey, np.ndarray] = {}
    unassigned, skipped = [], []
    for gname, tensors in inv.items():
        if not tensors:
            continue
        if gname not in file_layers:
            unassigned.append(gname)
            continue
        g = f[gname]
        names = [_decode(n) for n in np.atleast_1d(g.attrs.get("weight_names", []))]
        if len(names) != len(tensors):                                        # weight_io.py:185-195
            msg = f"layer {gname!r}: file has {len(names)} weights, model expects {len(tensors)}"
            if not skip_mismatch:
                raise ValueError(msg)
            skipped.append(msg)
            continue
        for i, (wn, (_, shape, _)) in enumerate(zip(names, tensors)):
            a = g[wn].read()
            if tuple(a.shape) != tuple(shape):                                # weight_io.py:219-232
                msg = f"layer {gname!r} weight {i}: file shape {a.shape}, model expects {tuple(shape)}"
                if not skip_mismatch:
                    raise ValueError(msg)
                skipped.append(msg)
                continue
            out[(gname, i)] = a
    model_layers = {gname for gname, tensors in inv.items() if tensors}
    unconsumed = [n for n in file_layers if n not in model_layers and len(np.atleast_1d(f[n].attrs.get("weight_names", [])))]
    if report is not None:
        report.update({"unassigned_layers": unassigned, "unconsumed_layers": unconsumed, "skipped": skipped})
    if verbose:
        if unconsumed:
            print("The following layers were not consumed from .h5 file:")
            for n in unconsumed:
                print("- " + n)
        if unassigned:
            print("The following layers were not assigned any weights:")
            for n in unassigned:
                print("- " + n)
        for msg in skipped:
            print("Skipping loading of weights: " + msg)
    return out
