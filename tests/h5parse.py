"""An independent reader of the HDF5 subset ``stryke_b200/h5lite.py`` writes — TEST INFRASTRUCTURE.

Written from the HDF5 file-format specification on its own (it shares no code with the writer): superblock version 0,
symbol-table groups (version-1 B-tree, local heap, symbol nodes), version-1 object headers, datatype classes 0 / 1 / 3 / 6 /
10, dataspace version 1, layout version 3 (contiguous or chunked with a version-1 chunk B-tree), version-1 attributes.  It
checks what a strict reader checks on the way (signatures, alignment, sorted links, B-tree key order, sizes) and raises
``H5Error`` when something does not hold.
"""
import pickle
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


def need(cond, msg):
    if not cond:
        raise H5Error(msg)


class File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            raw = fh.read()
        user = 0                                                   # a user block of 512, 1024, ... bytes may precede the superblock
        while raw[user:user + 8] != b"\x89HDF\r\n\x1a\n":
            user = 512 if user == 0 else 2 * user
            need(user < len(raw), "signature")
        self.b = b = raw[user:]                                    # addresses in the file are relative to the base address
        ver, fsver, rootver, _, shver, so, sl, _ = struct.unpack_from("<8B", b, 8)
        need((ver, fsver, rootver, shver) == (0, 0, 0, 0), "superblock version 0 expected")
        need((so, sl) == (8, 8), "8-byte offsets and lengths expected")
        self.leaf_k, self.internal_k, flags = struct.unpack_from("<HHI", b, 16)
        base, free, eof, drv = struct.unpack_from("<4Q", b, 24)
        need(base == user and free == UNDEF and drv == UNDEF, "base / free-space / driver addresses")
        need(eof == len(raw), f"end-of-file address {eof} != file size {len(raw)}")
        name_off, hdr, ctype, _ = struct.unpack_from("<QQII", b, 56)
        need(ctype == 1, "root entry caches its B-tree and heap")
        self.root_btree, self.root_heap = struct.unpack_from("<QQ", b, 80)
        self.root = self.object(hdr)
        need(self.root["kind"] == "group", "root is a group")
        need((self.root["btree"], self.root["heap"]) == (self.root_btree, self.root_heap), "root entry scratch == symbol table message")

    # ---- object headers ---------------------------------------------------------------------------------------------------
    def object(self, addr):
        b = self.b
        need(addr % 8 == 0, "object header alignment")
        ver, _, nmsg, refc, size = struct.unpack_from("<BBHII", b, addr)
        need(ver == 1 and refc >= 1, "object header version 1")
        p, end = addr + 16, addr + 16 + size
        need(end <= len(b), "object header inside the file")
        out = {"attrs": {}, "kind": None, "addr": addr}
        seen = 0
        while p < end:
            mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
            need(msize % 8 == 0, "message size multiple of 8")
            body = b[p + 8:p + 8 + msize]
            seen += 1
            if mtype == 0x0011:
                out["kind"] = "group"
                out["btree"], out["heap"] = struct.unpack_from("<QQ", body, 0)
            elif mtype == 0x0001:
                out["shape"], out["maxshape"] = self.dataspace(body)
            elif mtype == 0x0003:
                out["dtype"], _ = self.datatype(body, 0)
            elif mtype == 0x0005:
                need(body[0] in (1, 2), "fill value message version 1 or 2")
                out["fill"] = tuple(body[1:4])
                if body[3] or body[0] == 1:
                    need(struct.unpack_from("<i", body, 4)[0] >= 0, "fill value size")
            elif mtype in (0x0012, 0x000E):                       # modification time (new / old): informational
                pass
            elif mtype == 0x0008:
                out["kind"] = "dataset"
                out["layout"] = self.layout(body)
            elif mtype == 0x000C:
                k, v = self.attribute(body)
                need(k not in out["attrs"], f"duplicate attribute {k}")
                out["attrs"][k] = v
            elif mtype == 0x0000:
                pass
            else:
                raise H5Error(f"unexpected message type {mtype:#x}")
            p += 8 + msize
        need(p == end and seen == nmsg, "message count / header size")
        return out

    def dataspace(self, body):
        ver, rank, flags = struct.unpack_from("<BBB", body, 0)
        need(ver == 1, "dataspace version 1")
        dims = struct.unpack_from(f"<{rank}Q", body, 8)
        mx = struct.unpack_from(f"<{rank}Q", body, 8 + 8 * rank) if flags & 1 else None
        return tuple(dims), (tuple(None if m == UNDEF else m for m in mx) if mx else None)

    def datatype(self, body, p):
        cv, f0, f1, f2, size = struct.unpack_from("<BBBBI", body, p)
        cls, ver = cv & 15, cv >> 4
        p += 8
        if cls == 0:
            need(ver == 1 and (f0 & 1) == 0, "little-endian fixed point")
            off, prec = struct.unpack_from("<HH", body, p)
            need(off == 0 and prec == 8 * size, "full-precision integer")
            return np.dtype(("<i" if f0 & 8 else "<u") + str(size)), p + 4
        if cls == 1:
            need(ver == 1 and (f0 & 1) == 0 and (f0 >> 4) & 3 == 2, "little-endian IEEE float, implied mantissa bit")
            off, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", body, p)
            want = {4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[size]
            need((f1, eloc, esize, mloc, msize, bias) == want and off == 0 and prec == 8 * size, "IEEE layout")
            return np.dtype("<f" + str(size)), p + 12
        if cls == 3:
            need(ver == 1 and (f0 & 15) in (0, 1) and (f0 >> 4) in (0, 1), "null-terminated / null-padded ASCII or UTF-8 string")
            return np.dtype(f"S{size}"), p
        if cls == 10:
            need(ver == 2, "array datatype version 2")
            nd = body[p]
            dims = struct.unpack_from(f"<{nd}I", body, p + 4)
            p += 4 + 8 * nd
            base, p = self.datatype(body, p)
            dt = np.dtype((base, tuple(dims)))
            need(dt.itemsize == size, "array size")
            return dt, p
        if cls == 6:
            need(ver == 2, "compound datatype version 2")
            n = f0 | (f1 << 8)
            names, formats, offsets = [], [], []
            for _ in range(n):
                e = body.index(b"\0", p)
                names.append(body[p:e].decode("utf-8"))
                p += ((e - p + 1) + 7) // 8 * 8
                (off,) = struct.unpack_from("<I", body, p)
                sub, p = self.datatype(body, p + 4)
                formats.append(sub)
                offsets.append(off)
            return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), p
        raise H5Error(f"datatype class {cls}")

    def layout(self, body):
        ver, cls = body[0], body[1]
        if ver in (1, 2):                                          # files written by HDF5 <= 1.6: only contiguous is read here
            nd, cls = body[1], body[2]
            need(cls == 1, "old-style layout: contiguous only")
            (addr,) = struct.unpack_from("<Q", body, 8)
            dims = struct.unpack_from(f"<{nd}I", body, 16)
            return {"class": "contiguous", "addr": addr, "size": int(np.prod(dims))}
        need(ver == 3, "layout version 3")
        if cls == 1:
            addr, size = struct.unpack_from("<QQ", body, 2)
            return {"class": "contiguous", "addr": addr, "size": size}
        need(cls == 2, "contiguous or chunked layout")
        nd = body[2]
        (bt,) = struct.unpack_from("<Q", body, 3)
        dims = struct.unpack_from(f"<{nd}I", body, 11)
        return {"class": "chunked", "btree": bt, "chunk": tuple(dims[:-1]), "elem": dims[-1]}

    def attribute(self, body):
        ver, _, nsz, dsz, ssz = struct.unpack_from("<BBHHH", body, 0)
        need(ver == 1, "attribute version 1")
        p = 8
        name = body[p:p + nsz]
        need(name.endswith(b"\0"), "attribute name terminator")
        p += (nsz + 7) // 8 * 8
        dt, _ = self.datatype(body[p:p + dsz], 0)
        p += (dsz + 7) // 8 * 8
        shape, _ = self.dataspace(body[p:p + ssz])
        p += (ssz + 7) // 8 * 8
        need(shape == (), "scalar attributes only")
        raw = body[p:p + dt.itemsize]
        need(len(raw) == dt.itemsize, "attribute data present")
        if dt.kind == "S":
            v = raw
            if v[:1] in (b"(", b"]", b"}", b"N", b"c", b"\x80") and v.rstrip(b"\0").endswith(b"."):      # PyTables: a pickle
                try:
                    return name[:-1].decode(), pickle.loads(v)
                except Exception:
                    pass
            return name[:-1].decode(), v.rstrip(b"\0").decode("utf-8")
        return name[:-1].decode(), np.frombuffer(raw, dtype=dt)[0]

    # ---- groups -------------------------------------------------------------------------------------------------------------
    def heap_string(self, heap_addr, off):
        b = self.b
        need(b[heap_addr:heap_addr + 4] == b"HEAP" and b[heap_addr + 4] == 0, "local heap")
        size, free, seg = struct.unpack_from("<QQQ", b, heap_addr + 8)
        need(free == 1 or free < size, "heap free-list head")
        need(off < size, "name inside the heap")
        e = b.index(b"\0", seg + off)
        need(e < seg + size, "name terminated inside the heap")
        return b[seg + off:e]

    def links(self, obj):
        """{name: object header address} of a group, walking its B-tree; checks ordering and keys."""
        need(obj["kind"] == "group", "not a group")
        out = {}
        names = []

        def node(addr):
            b = self.b
            need(b[addr:addr + 4] == b"TREE", "B-tree signature")
            ntype, level, used, left, right = struct.unpack_from("<BBHQQ", b, addr + 4)
            need(ntype == 0, "group B-tree node")
            need(used <= 2 * self.internal_k, "entries used <= 2 K")
            p = addr + 24
            keys, kids = [], []
            for i in range(used):
                keys.append(struct.unpack_from("<Q", b, p)[0])
                kids.append(struct.unpack_from("<Q", b, p + 8)[0])
                p += 16
            keys.append(struct.unpack_from("<Q", b, p)[0])
            for i, kid in enumerate(kids):
                if level > 0:
                    node(kid)
                    continue
                need(b[kid:kid + 4] == b"SNOD" and b[kid + 4] == 1, "symbol node")
                (n,) = struct.unpack_from("<H", b, kid + 6)
                need(1 <= n <= 2 * self.leaf_k, "symbols per node")
                last = None
                for j in range(n):
                    off, hdr, ctype, _ = struct.unpack_from("<QQII", b, kid + 8 + 40 * j)
                    nm = self.heap_string(obj["heap"], off)
                    names.append(nm)
                    out[nm.decode("utf-8")] = hdr
                    last = nm
                    if ctype == 1:
                        bt, hp = struct.unpack_from("<QQ", b, kid + 8 + 40 * j + 24)
                        child = self.object(hdr)
                        need((child["btree"], child["heap"]) == (bt, hp), "cached B-tree / heap of a subgroup")
                lo = self.heap_string(obj["heap"], keys[i])
                hi = self.heap_string(obj["heap"], keys[i + 1])
                need(hi == last, "right key = largest name of the node")
                need(lo < names[len(names) - n], "left key < smallest name of the node")
        node(obj["btree"])
        need(names == sorted(names), "links in strcmp order")
        return out

    def get(self, path):
        obj = self.root
        for part in [p for p in path.split("/") if p]:
            ln = self.links(obj)
            need(part in ln, f"no link {part!r}")
            obj = self.object(ln[part])
        return obj

    def walk(self, obj=None, prefix=""):
        obj = obj or self.root
        yield prefix or "/", obj
        if obj["kind"] == "group":
            for name, addr in self.links(obj).items():
                yield from self.walk(self.object(addr), f"{prefix}/{name}")

    # ---- datasets -----------------------------------------------------------------------------------------------------------
    def read(self, obj):
        need(obj["kind"] == "dataset", "not a dataset")
        dt, lay = obj["dtype"], obj["layout"]
        if lay["class"] == "contiguous":
            n = int(np.prod(obj["shape"]))
            return np.frombuffer(self.b, dtype=dt, count=n, offset=lay["addr"]).copy().reshape(obj["shape"])
        (n,) = obj["shape"]
        need(lay["elem"] == dt.itemsize and len(lay["chunk"]) == 1, "chunk geometry")
        need(obj["maxshape"] == (None,), "extendible along the row axis")
        cr = lay["chunk"][0]
        out = np.zeros(n, dtype=dt)
        seen = []

        def node(addr, lo_key=None, hi_key=None):
            b = self.b
            need(b[addr:addr + 4] == b"TREE", "chunk B-tree signature")
            ntype, level, used, left, right = struct.unpack_from("<BBHQQ", b, addr + 4)
            need(ntype == 1 and used <= 64, "chunk B-tree node")
            p = addr + 24
            prev = None
            for i in range(used):
                nbytes, mask, off, zero = struct.unpack_from("<IIQQ", b, p)
                (child,) = struct.unpack_from("<Q", b, p + 24)
                nxt_off = struct.unpack_from("<Q", b, p + 32 + 8)[0]
                need(mask == 0 and zero == 0, "no filters; element offset 0")
                need(prev is None or off > prev, "keys increase")
                need(nxt_off > off, "right key beyond the chunk")
                prev = off
                if level > 0:
                    node(child)
                else:
                    need(nbytes == cr * dt.itemsize and off % cr == 0, "chunk size / offset")
                    rows = min(cr, n - off)
                    out[off:off + rows] = np.frombuffer(b, dtype=dt, count=rows, offset=child)
                    seen.append(off)
                p += 32
            return left, right
        if n:
            node(lay["btree"])
        need(seen == list(range(0, n, cr)), "every chunk present once, in order")
        return out
