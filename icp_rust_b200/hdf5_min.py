"""Dependency-free reader for the one HDF5 layout tests/golden needs: superblock version 0, a root
group stored as v1 B-tree + symbol-table nodes + local heap, datasets with version-1 object
headers holding a simple dataspace, an IEEE f64 little-endian datatype and a CONTIGUOUS layout
(message version 3).  That is how /root/reference/scans/3d/scans.hdf5 is written
(examples/scan3d.rs:34-61 reads it through the hdf5 crate).  Not a general HDF5 implementation."""
import struct

import numpy as np


class MiniH5:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        assert b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        assert b[8] == 0, "superblock version 0 expected"
        assert b[13] == 8 and b[14] == 8, "8-byte offsets/lengths expected"
        # superblock v0: ... base addr @24, free-space @32, eof @40, driver @48, root symbol-table entry @56
        self.root = self._symtab_entry(56)

    def _u(self, off, n):
        return int.from_bytes(self.b[off:off + n], "little")

    def _symtab_entry(self, off):
        name_off, obj_hdr, cache = self._u(off, 8), self._u(off + 8, 8), self._u(off + 16, 4)
        scratch = self.b[off + 24:off + 40]
        return dict(name_off=name_off, obj=obj_hdr, cache=cache, scratch=scratch)

    def _heap_data(self, heap_addr):
        assert self.b[heap_addr:heap_addr + 4] == b"HEAP"
        return self._u(heap_addr + 24, 8)

    def _walk_btree(self, addr, heap_data, out):
        b = self.b
        assert b[addr:addr + 4] == b"TREE", "v1 B-tree expected"
        node_type, level, used = b[addr + 4], b[addr + 5], self._u(addr + 6, 2)
        assert node_type == 0
        p = addr + 24  # after signature, type, level, entries, left, right siblings
        children = []
        for i in range(used):
            p += 8                      # key i
            children.append(self._u(p, 8))
            p += 8
        for c in children:
            if level > 0:
                self._walk_btree(c, heap_data, out)
            else:
                assert b[c:c + 4] == b"SNOD"
                n = self._u(c + 6, 2)
                for e in range(n):
                    ent = self._symtab_entry(c + 8 + 40 * e)
                    s = heap_data + ent["name_off"]
                    name = b[s:b.index(b"\0", s)].decode()
                    out[name] = ent["obj"]

    def datasets(self):
        """name -> object header address, for the root group."""
        assert self.root["cache"] == 1, "root group with cached B-tree / heap addresses expected"
        btree, heap = struct.unpack("<QQ", self.root["scratch"])
        out = {}
        self._walk_btree(btree, self._heap_data(heap), out)
        return dict(sorted(out.items()))

    def read(self, obj_addr):
        b = self.b
        assert b[obj_addr] == 1, "version-1 object header expected"
        nmsg = self._u(obj_addr + 2, 2)
        hdr_size = self._u(obj_addr + 8, 4)
        p = obj_addr + 16
        end = p + hdr_size
        dims = None
        data_addr = data_size = None
        blocks = [(p, end)]
        seen = 0
        while blocks and seen < nmsg:
            p, end = blocks.pop(0)
            while p + 8 <= end and seen < nmsg:
                mtype, msize, _flags = self._u(p, 2), self._u(p + 2, 2), b[p + 4]
                body = p + 8
                if mtype == 0x0001:      # dataspace
                    ver, rank = b[body], b[body + 1]
                    assert ver == 1
                    dims = [self._u(body + 8 + 8 * i, 8) for i in range(rank)]
                elif mtype == 0x0003:    # datatype
                    cls = b[body] & 0x0f
                    size = self._u(body + 4, 4)
                    assert cls == 1 and size == 8 and (b[body + 1] & 1) == 0, "IEEE f64 little-endian expected"
                elif mtype == 0x0008:    # layout
                    ver, lclass = b[body], b[body + 1]
                    assert ver == 3 and lclass == 1, "contiguous layout (v3) expected"
                    data_addr, data_size = self._u(body + 2, 8), self._u(body + 10, 8)
                elif mtype == 0x0010:    # continuation
                    blocks.append((self._u(body, 8), self._u(body, 8) + self._u(body + 8, 8)))
                p = body + msize
                seen += 1
        assert dims is not None and data_addr is not None
        n = int(np.prod(dims))
        assert data_size == 8 * n
        return np.frombuffer(b, dtype="<f8", count=n, offset=data_addr).reshape(dims).copy()
