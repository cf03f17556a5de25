"""Minimal read-only HDF5 parser (superblock v0, v1 object headers, contiguous layout).

Just enough to read GauXC's golden test fixtures (tests/ref_data/*.hdf5 in the
reference tree) without libhdf5/h5py.  Only used offline by
tools/make_golden.py to convert fixtures to flat .npz files under tests/golden/.
"""
import struct
import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        assert self.b[:8] == b"\x89HDF\r\n\x1a\n", "not HDF5"
        assert self.b[8] == 0, "superblock version != 0"
        assert self.b[13] == 8 and self.b[14] == 8
        # root symbol table entry at byte 56
        self.root = self._read_ste(56)

    def _read_ste(self, off):
        name_off, ohdr, cache, _ = struct.unpack_from("<QQII", self.b, off)
        scratch = self.b[off + 24: off + 40]
        return dict(name_off=name_off, ohdr=ohdr, cache=cache, scratch=scratch)

    # ---- object headers -------------------------------------------------
    def _messages(self, addr):
        ver, _, nmsgs, _ref, hsize = struct.unpack_from("<BBHII", self.b, addr)
        assert ver == 1
        msgs = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsgs:
            p, ln = blocks.pop(0)
            end = p + ln
            while p + 8 <= end and len(msgs) < nmsgs:
                t, sz, fl = struct.unpack_from("<HHB", self.b, p)
                body = p + 8
                if t == 0x10:
                    ca, cl = struct.unpack_from("<QQ", self.b, body)
                    blocks.append((ca, cl))
                msgs.append((t, body, sz))
                p = body + sz
        return msgs

    # ---- groups ----------------------------------------------------------
    def _group_entries(self, btree, heap):
        assert self.b[heap:heap + 4] == b"HEAP"
        data_addr = struct.unpack_from("<Q", self.b, heap + 24)[0]
        out = {}

        def walk(node):
            assert self.b[node:node + 4] == b"TREE"
            ntype, level, nent = struct.unpack_from("<BBH", self.b, node + 4)
            p = node + 24
            # key0, child0, key1, child1, ... key_n
            for i in range(nent):
                child = struct.unpack_from("<Q", self.b, p + 8 + 16 * i)[0]
                if level > 0:
                    walk(child)
                else:
                    assert self.b[child:child + 4] == b"SNOD"
                    nsym = struct.unpack_from("<H", self.b, child + 6)[0]
                    for k in range(nsym):
                        ste = self._read_ste(child + 8 + 40 * k)
                        s = data_addr + ste["name_off"]
                        e = self.b.index(b"\0", s)
                        out[self.b[s:e].decode()] = ste
        walk(btree)
        return out

    def _children(self, ohdr):
        for t, body, sz in self._messages(ohdr):
            if t == 0x11:
                bt, hp = struct.unpack_from("<QQ", self.b, body)
                return self._group_entries(bt, hp)
        return None

    def _resolve(self, path):
        ohdr = self.root["ohdr"]
        for part in [p for p in path.split("/") if p]:
            ch = self._children(ohdr)
            if ch is None or part not in ch:
                raise KeyError(path)
            ohdr = ch[part]["ohdr"]
        return ohdr

    def keys(self, path="/"):
        ch = self._children(self._resolve(path))
        return sorted(ch.keys()) if ch else []

    def is_group(self, path):
        return self._children(self._resolve(path)) is not None

    # ---- datasets --------------------------------------------------------
    def raw(self, path):
        """Returns (dims, elem_size, type_class, bytes)."""
        ohdr = self._resolve(path)
        dims, esize, tclass, addr, size = (), None, None, None, None
        for t, body, sz in self._messages(ohdr):
            if t == 0x01:
                ver, rank, flags = struct.unpack_from("<BBB", self.b, body)
                off = body + (8 if ver == 1 else 4)
                dims = struct.unpack_from("<%dQ" % rank, self.b, off) if rank else ()
            elif t == 0x03:
                tclass = self.b[body] & 0x0F
                esize = struct.unpack_from("<I", self.b, body + 4)[0]
            elif t == 0x08:
                ver, cls = struct.unpack_from("<BB", self.b, body)
                assert ver == 3, "layout version %d" % ver
                if cls == 1:
                    addr, size = struct.unpack_from("<QQ", self.b, body + 2)
                elif cls == 0:  # compact
                    size = struct.unpack_from("<H", self.b, body + 2)[0]
                    addr = body + 4
                else:
                    raise NotImplementedError("chunked layout")
        n = int(np.prod(dims)) if dims else 1
        if addr == UNDEF or addr is None:
            data = b""
        else:
            data = self.b[addr: addr + n * esize]
        return dims, esize, tclass, data

    def array(self, path):
        dims, esize, tclass, data = self.raw(path)
        if tclass == 1:
            dt = {8: "<f8", 4: "<f4"}[esize]
        elif tclass == 0:
            dt = {8: "<i8", 4: "<i4", 2: "<i2", 1: "<i1"}[esize]
        else:
            raise TypeError("class %d: use raw()" % tclass)
        a = np.frombuffer(data, dtype=dt)
        return a.reshape(dims) if dims else a

    def walk(self, path="/", depth=0, out=None):
        out = [] if out is None else out
        for k in self.keys(path):
            p = path.rstrip("/") + "/" + k
            if self.is_group(p):
                out.append((p, "group"))
                self.walk(p, depth + 1, out)
            else:
                dims, es, tc, data = self.raw(p)
                out.append((p, (dims, es, tc, len(data))))
        return out


if __name__ == "__main__":
    import sys
    f = H5File(sys.argv[1])
    for p, info in f.walk():
        print(p, info)
