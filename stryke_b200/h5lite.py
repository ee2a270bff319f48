"""A small HDF5 *writer* (pure Python + NumPy) for the one layout the reference's result file uses: the pandas / PyTables
"table" format — ``/<key>/table``, a chunked, extendible 1-D compound dataset with an ``index`` member and one
``values_block_N`` array member per dtype block (fixed-width strings for labels), under groups that carry the PyTables and
pandas attributes (``tools/analyze_sim_table_h5py.py:4-9`` reads exactly that with h5py; ``Stryke/stryke.py:3344, :3379,
:3457, :3497, :4093-4099`` write it through ``pandas.HDFStore``).

This image has neither libhdf5 nor PyTables, so the file is emitted byte by byte from the HDF5 file-format
specification, restricted to the oldest, most widely readable structures: superblock version 0, version-1 object headers,
symbol-table groups (version-1 B-tree + local heap + symbol nodes), layout message version 3 (chunked, version-1 chunk
B-tree), datatype classes 0 / 1 / 3 / 6 / 10 (fixed point, IEEE float, string, compound, array), version-1 attributes.
``tests/h5parse.py`` is an independent reader of the same subset used by the tests (structure + values); reading the files
back with ``pandas.HDFStore`` is tested wherever PyTables exists and skipped elsewhere — on this image the file has NOT
been opened by a real HDF5 library.
"""
from __future__ import annotations

import pickle
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
GROUP_LEAF_K = 32          # symbol nodes hold 2 K = 64 entries
GROUP_INTERNAL_K = 16      # group B-tree nodes hold 2 K = 32 children
CHUNK_K = 32               # chunk B-tree nodes hold 2 K = 64 children (library default; superblock v0 cannot say otherwise)


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ---- datatype messages ---------------------------------------------------------------------------------------------------
def datatype_message(dt: np.dtype) -> bytes:
    """The body of a Datatype message (type 0x0003) for a NumPy dtype: integers, IEEE floats, fixed-width byte strings,
    sub-array members and structured (compound) types."""
    dt = np.dtype(dt)
    if dt.subdtype is not None:                                   # array member: class 10, version 2
        base, shape = dt.subdtype
        body = struct.pack("<B3x", len(shape)) + b"".join(struct.pack("<I", int(s)) for s in shape)
        body += b"".join(struct.pack("<I", i) for i in range(len(shape)))              # permutation indices (unused by the library)
        return struct.pack("<B3BI", (2 << 4) | 10, 0, 0, 0, dt.itemsize) + body + datatype_message(base)
    if dt.names is not None:                                      # compound: class 6, version 2 (array members need >= 2)
        n = len(dt.names)
        out = struct.pack("<B3BI", (2 << 4) | 6, n & 0xFF, (n >> 8) & 0xFF, 0, dt.itemsize)
        for name in dt.names:
            sub, off = dt.fields[name][0], dt.fields[name][1]
            out += _pad8(name.encode("utf-8") + b"\0") + struct.pack("<I", off) + datatype_message(sub)
        return out
    if dt.kind in "iu":                                           # fixed point: class 0, version 1
        flags = 0x08 if dt.kind == "i" else 0x00                  # bit 0 = 0: little endian; bit 3: signed
        return struct.pack("<B3BI", (1 << 4) | 0, flags, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "f" and dt.itemsize in (4, 8):                  # IEEE float: class 1, version 1
        sign, eloc, esize, msize, bias = (31, 23, 8, 23, 127) if dt.itemsize == 4 else (63, 52, 11, 52, 1023)
        return (struct.pack("<B3BI", (1 << 4) | 1, 0x20, sign, 0, dt.itemsize)              # 0x20: mantissa normalisation 2 (implied msb)
                + struct.pack("<HHBBBBI", 0, dt.itemsize * 8, eloc, esize, 0, msize, bias))
    if dt.kind == "S":                                            # string: class 3, null-terminated, ASCII
        return struct.pack("<B3BI", (1 << 4) | 3, 0x00, 0, 0, max(dt.itemsize, 1))
    raise TypeError(f"h5lite: unsupported dtype {dt!r}")


def _string_type(size: int, utf8: bool) -> bytes:
    return struct.pack("<B3BI", (1 << 4) | 3, 0x10 if utf8 else 0x00, 0, 0, max(size, 1))


def dataspace_message(shape, maxshape=None) -> bytes:
    """Dataspace message (type 0x0001), version 1.  ``shape=()`` is a scalar."""
    rank = len(shape)
    out = struct.pack("<BBB5x", 1, rank, 1 if maxshape is not None else 0)
    out += b"".join(struct.pack("<Q", int(s)) for s in shape)
    if maxshape is not None:
        out += b"".join(struct.pack("<Q", UNDEF if m is None else int(m)) for m in maxshape)
    return out


def attribute_message(name: str, value) -> bytes:
    """Attribute message (type 0x000C), version 1.  ``str`` -> UTF-8 fixed string, ``bytes`` -> ASCII fixed string, NumPy /
    Python scalars -> scalar of that type, anything else -> its protocol-0 pickle as an ASCII string (PyTables' convention)."""
    if isinstance(value, str):
        data = value.encode("utf-8")
        dt = _string_type(len(data), True)
        data = data or b"\0"
    elif isinstance(value, (bytes, np.bytes_)):
        data = bytes(value)
        dt = _string_type(len(data), False)
        data = data or b"\0"
    elif isinstance(value, (bool, np.bool_)):
        data, dt = np.uint8(1 if value else 0).tobytes(), datatype_message(np.uint8)
    elif isinstance(value, (int, np.integer)):
        data, dt = np.int64(value).tobytes(), datatype_message(np.int64)
    elif isinstance(value, (float, np.floating)):
        data, dt = np.float64(value).tobytes(), datatype_message(np.float64)
    else:
        data = pickle.dumps(value, protocol=0)
        dt = _string_type(len(data), False)
    ds = dataspace_message(())
    nm = name.encode("utf-8") + b"\0"
    return struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + data


def object_header(messages) -> bytes:
    """Version-1 object header holding ``messages`` = [(type, body bytes)], all in the first (only) block."""
    body = b""
    for mtype, data in messages:
        data = _pad8(data)
        if len(data) > 0xFFFF:
            raise ValueError("h5lite: header message larger than 64 KB")
        body += struct.pack("<HHB3x", mtype, len(data), 0) + data
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body


# ---- the tree the caller builds -------------------------------------------------------------------------------------
class Group:
    def __init__(self):
        self.attrs = {}
        self.children = {}      # name -> Group | Table


class Table:
    """A 1-D compound dataset: ``data`` is a structured array; chunked, extendible along its only axis."""

    def __init__(self, data: np.ndarray, attrs=None, chunk_rows=None):
        if data.dtype.names is None or data.ndim != 1:
            raise TypeError("h5lite.Table wants a 1-D structured array")
        self.data = data
        self.attrs = dict(attrs or {})
        rows = max(1, min(len(data), max(1, (1 << 20) // max(1, data.dtype.itemsize))))      # ~1 MB chunks, like PyTables
        self.chunk_rows = int(chunk_rows or rows)


class Writer:
    """Build a tree with ``group`` / ``table``, then ``write`` it."""

    def __init__(self):
        self.root = Group()

    def group(self, path: str) -> Group:
        g = self.root
        for part in [p for p in path.split("/") if p]:
            nxt = g.children.get(part)
            if nxt is None:
                nxt = g.children[part] = Group()
            if not isinstance(nxt, Group):
                raise ValueError(f"h5lite: {part!r} in {path!r} is a dataset")
            g = nxt
        return g

    def table(self, path: str, data, attrs=None, chunk_rows=None) -> Table:
        parts = [p for p in path.split("/") if p]
        g = self.group("/".join(parts[:-1]))
        t = g.children[parts[-1]] = Table(data, attrs, chunk_rows)
        return t

    # -- serialisation ---------------------------------------------------------------------------------------------------
    def write(self, filename: str):
        with open(filename, "wb") as fh:
            self._fh = fh
            fh.write(b"\0" * 96)                                  # superblock, filled in at the end
            addr, btree, heap = self._write_group(self.root)
            eof = self._tell()
            sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
            sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
            sb += struct.pack("<QQII", 0, addr, 1, 0) + struct.pack("<QQ", btree, heap)        # root symbol table entry (cached: group)
            assert len(sb) == 96
            fh.seek(0)
            fh.write(sb)
        self._fh = None

    def _tell(self):
        return self._fh.tell()

    def _put(self, data: bytes) -> int:
        """Append ``data`` at the next 8-byte boundary; returns its address."""
        pos = self._fh.tell()
        if pos % 8:
            self._fh.write(b"\0" * (8 - pos % 8))
            pos += 8 - pos % 8
        self._fh.write(data)
        return pos

    def _write_group(self, g: Group):
        entries = []                                              # (name bytes, object header address, cache type, scratch)
        for name in g.children:
            child = g.children[name]
            if isinstance(child, Group):
                a, bt, hp = self._write_group(child)
                entries.append((name.encode("utf-8"), a, 1, struct.pack("<QQ", bt, hp)))
            else:
                entries.append((name.encode("utf-8"), self._write_table(child), 0, b"\0" * 16))
        entries.sort(key=lambda e: e[0])                          # symbol nodes hold their entries in strcmp order
        # local heap: the empty string at offset 0, then the names (each padded to 8 bytes)
        seg, offs = bytearray(b"\0" * 8), []
        for nm, *_ in entries:
            offs.append(len(seg))
            seg += _pad8(nm + b"\0")
        seg_addr = self._put(bytes(seg))
        heap_addr = self._put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), 1, seg_addr))      # free list head 1 = H5HL_FREE_NULL
        # symbol nodes, then one B-tree node over them (2 K x 2 K entries: 2048 links per group)
        per = 2 * GROUP_LEAF_K
        nodes = [list(range(i, min(len(entries), i + per))) for i in range(0, max(len(entries), 1), per)]
        if len(nodes) > 2 * GROUP_INTERNAL_K:
            raise ValueError("h5lite: more than 2048 links in one group")
        snods = []
        for idx in nodes:
            body = b"SNOD" + struct.pack("<BxH", 1, len(idx))
            for i in idx:
                nm, a, ctype, scratch = entries[i]
                body += struct.pack("<QQII", offs[i], a, ctype, 0) + scratch
            body += b"\0" * (40 * (per - len(idx)))
            snods.append(self._put(body))
        keys = [0] + [offs[idx[-1]] if idx else 0 for idx in nodes]
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(nodes) if entries else 0, UNDEF, UNDEF)
        slots = 2 * GROUP_INTERNAL_K
        for i in range(slots):
            node += struct.pack("<Q", keys[i] if i < len(keys) else 0)
            node += struct.pack("<Q", snods[i] if (i < len(snods) and entries) else 0)
        node += struct.pack("<Q", keys[slots] if slots < len(keys) else 0)
        btree_addr = self._put(node)
        msgs = [(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]
        msgs += [(0x000C, attribute_message(k, v)) for k, v in g.attrs.items()]
        return self._put(object_header(msgs)), btree_addr, heap_addr

    def _write_table(self, t: Table) -> int:
        data, rowsize, cr = t.data, t.data.dtype.itemsize, t.chunk_rows
        n = len(data)
        chunk_bytes = cr * rowsize
        leaves = []                                               # (row offset, address)
        raw = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        for lo in range(0, n, cr):
            piece = raw[lo * rowsize:min(n, lo + cr) * rowsize].tobytes()
            leaves.append((lo, self._put(piece + b"\0" * (chunk_bytes - len(piece)))))      # edge chunks are allocated whole
        end = ((n + cr - 1) // cr) * cr
        btree_addr = self._write_chunk_btree(leaves, chunk_bytes, end)
        msgs = [
            (0x0001, dataspace_message((n,), (None,))),
            (0x0003, datatype_message(data.dtype)),
            (0x0005, struct.pack("<BBBBI", 2, 3, 2, 1, 0)),      # fill value v2: incremental allocation, written if set, default value
            (0x0008, struct.pack("<BBBQII", 3, 2, 2, btree_addr, cr, rowsize)),      # layout v3, chunked, rank + 1 dimensions
        ]
        msgs += [(0x000C, attribute_message(k, v)) for k, v in t.attrs.items()]
        return self._put(object_header(msgs))

    def _write_chunk_btree(self, leaves, chunk_bytes, end_offset) -> int:
        """Version-1 B-tree (node type 1) over the chunks; ``leaves`` = [(row offset, address)] in order."""
        def key(nbytes, off):
            return struct.pack("<IIQQ", nbytes, 0, off, 0)
        slots = 2 * CHUNK_K
        size = 24 + (slots + 1) * 24 + slots * 8
        level = 0
        items = [(off, addr, chunk_bytes) for off, addr in leaves]         # (first row offset, child address, key size field)
        if not items:
            node = b"TREE" + struct.pack("<BBHQQ", 1, 0, 0, UNDEF, UNDEF)
            return self._put(node + b"\0" * (size - len(node)))
        while True:
            groups = [items[i:i + slots] for i in range(0, len(items), slots)]
            base = self._tell()
            base += (-base) % 8
            addrs = [base + i * size for i in range(len(groups))]           # nodes of one level are written back to back
            nxt = []
            for gi, grp in enumerate(groups):
                left = addrs[gi - 1] if gi > 0 else UNDEF
                right = addrs[gi + 1] if gi + 1 < len(groups) else UNDEF
                node = b"TREE" + struct.pack("<BBHQQ", 1, level, len(grp), left, right)
                for off, addr, nb in grp:
                    node += key(nb, off) + struct.pack("<Q", addr)
                last_end = groups[gi + 1][0][0] if gi + 1 < len(groups) else end_offset
                node += key(0, last_end)
                node += b"\0" * (size - len(node))
                a = self._put(node)
                assert a == addrs[gi]
                nxt.append((grp[0][0], a, grp[0][2]))
            if len(nxt) == 1:
                return nxt[0][1]
            items, level = nxt, level + 1


# ---- pandas "table" layout ----------------------------------------------------------------------------------------
def _block_order(df):
    """Columns grouped the way pandas' consolidated BlockManager orders its blocks for ``format='table'`` — by dtype name:
    bool, datetime64[ns], float64, int64, object — each block's columns in frame order.  (A reader goes by the ``*_kind``
    attributes, not by the order.)"""
    blocks = {}
    for c in df.columns:
        col = df[c]
        if col.dtype.kind == "f":
            k = ("float64", "f8")
        elif col.dtype.kind in "iu":
            k = ("int64", "i8")
        elif col.dtype.kind == "b":
            k = ("bool", "u1")
        elif col.dtype.kind == "M":
            k = ("datetime64[ns]", "i8")
        else:
            k = ("string", None)
        blocks.setdefault(k, []).append(c)
    order = [("bool", "u1"), ("datetime64[ns]", "i8"), ("float64", "f8"), ("int64", "i8"), ("string", None)]
    return [(k, blocks[k]) for k in order if k in blocks]


def frame_table(df, nan_rep="nan", min_itemsize=None, data_columns=False):
    """(structured array, group attributes, table attributes) of a DataFrame in pandas' ``format='table'`` layout:
    ``index`` + one ``values_block_N`` per dtype block (or one member per column with ``data_columns=True``)."""
    import pandas as pd
    n = len(df)
    index = np.asarray(df.index)
    if index.dtype.kind not in "iu":
        index = np.arange(n, dtype=np.int64)
    fields, arrays, tattrs = [("index", "<i8")], {"index": index.astype(np.int64)}, {}
    tattrs["index_kind"] = "integer"
    values_cols, data_cols = [], []

    def as_strings(cols):
        """[n, len(cols)] fixed-width byte strings; every column is encoded through its distinct values (label columns of
        long frames hold a handful of them)."""
        enc, width = [], 1
        for c in cols:
            col = df[c]
            if isinstance(col.dtype, pd.CategoricalDtype):
                codes, uniq = np.asarray(col.cat.codes), list(col.cat.categories)
            else:
                try:
                    codes, uniq = pd.factorize(col, use_na_sentinel=True)
                    uniq = list(uniq)
                except TypeError:                                 # unhashable cells (lists): stringify first
                    codes, uniq = pd.factorize(col.map(str), use_na_sentinel=True)
                    uniq = list(uniq)
            labels = [(x if isinstance(x, str) else str(x)).encode("utf-8") for x in uniq] + [nan_rep.encode("utf-8")]      # [-1] = missing
            width = max([width] + [len(x) for x in labels[:-1]] + ([len(labels[-1])] if (np.asarray(codes) < 0).any() else []))
            enc.append((np.asarray(codes), labels))
        if isinstance(min_itemsize, dict):
            width = max([width] + [int(min_itemsize[c]) for c in cols if c in min_itemsize])
        elif min_itemsize:
            width = max(width, int(min_itemsize))
        out = np.zeros((n, len(cols)), dtype=f"S{width}")
        for j, (codes, labels) in enumerate(enc):
            out[:, j] = np.array(labels, dtype=f"S{width}")[codes]
        return out, width

    blocks = _block_order(df)
    if data_columns:
        blocks = [((kind, code), [c]) for (kind, code), cols in blocks for c in cols]
        blocks.sort(key=lambda b: list(df.columns).index(b[1][0]))
    for bi, ((kind, code), cols) in enumerate(blocks):
        name = str(cols[0]) if data_columns else f"values_block_{bi}"
        if kind == "string":
            arr, width = as_strings(cols)
            base, dtype_name = f"S{width}", f"bytes{8 * width}"
        else:
            arr = np.stack([df[c].to_numpy().astype(code if kind != "datetime64[ns]" else "datetime64[ns]").view("i8")
                            if kind == "datetime64[ns]" else df[c].to_numpy().astype(code) for c in cols], axis=1) if n else \
                np.zeros((0, len(cols)), dtype=code)
            base, dtype_name = "<" + code, kind
        if data_columns:
            fields.append((name, base))
            arrays[name] = arr[:, 0]
            data_cols.append(name)
        else:
            fields.append((name, base, (len(cols),)))
            arrays[name] = arr
        values_cols.append(name)
        tattrs[f"{name}_kind"] = [str(c) for c in cols]
        tattrs[f"{name}_meta"] = None
        tattrs[f"{name}_dtype"] = dtype_name
    rec = np.zeros(n, dtype=np.dtype(fields))
    for k, v in arrays.items():
        rec[k] = v
    table_attrs = {"CLASS": "TABLE", "VERSION": "2.7", "TITLE": ""}
    for i, nm in enumerate(rec.dtype.names):
        sub = rec.dtype.fields[nm][0]
        base = sub.subdtype[0] if sub.subdtype else sub
        table_attrs[f"FIELD_{i}_NAME"] = nm
        table_attrs[f"FIELD_{i}_FILL"] = (b"" if base.kind == "S" else (np.float64(0.0) if base.kind == "f" else np.int64(0)))
    table_attrs["NROWS"] = np.int64(n)
    table_attrs.update(tattrs)
    group_attrs = {
        "CLASS": "GROUP", "TITLE": "", "VERSION": "1.0",
        "pandas_type": "frame_table", "pandas_version": "0.15.2", "table_type": "appendable_frame",
        "index_cols": [(0, "index")], "values_cols": values_cols,
        "non_index_axes": [(1, [str(c) for c in df.columns])], "data_columns": data_cols,
        "nan_rep": nan_rep, "encoding": "UTF-8", "errors": "strict", "levels": 1, "metadata": [],
        "info": {1: {"type": "Index", "names": [None]}, "index": {}},
    }
    return rec, group_attrs, table_attrs


def write_frames(filename, frames, min_itemsize=None, data_columns=()):
    """Write ``{key: DataFrame}`` as one HDF5 file in the reference's layout (``/<key>/table`` per frame; keys may be nested,
    e.g. ``simulations/<scenario>/<species>``)."""
    w = Writer()
    w.root.attrs.update({"CLASS": "GROUP", "PYTABLES_FORMAT_VERSION": "2.1", "TITLE": "", "VERSION": "1.0"})
    for key, df in frames.items():
        path = "/".join(p for p in str(key).split("/") if p)
        rec, gattrs, tattrs = frame_table(df, min_itemsize=(min_itemsize or {}).get(key) if isinstance(min_itemsize, dict) else min_itemsize,
                                          data_columns=key in data_columns)
        parts = path.split("/")
        g = w.root
        for p in parts[:-1]:                                      # intermediate groups carry the plain PyTables group attributes
            g = w.group("/".join(parts[:parts.index(p) + 1]))
            g.attrs.setdefault("CLASS", "GROUP"); g.attrs.setdefault("TITLE", ""); g.attrs.setdefault("VERSION", "1.0")
        w.group(path).attrs.update(gattrs)
        w.table(path + "/table", rec, tattrs)
    w.write(filename)
    return filename
