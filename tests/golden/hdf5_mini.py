"""Minimal read-only HDF5 reader (superblock v0, v1 object headers, symbol-table groups,
contiguous / compact / chunked-deflate datasets of little-endian ints and floats).

h5py is not installed in this image; this is just enough to pull the integrals out of the
reference's `profiling/Hring_12.hdf5` (written by openfermion's MolecularData.save) so that
BASELINE.json's config 0 can be turned into a committed fixture.  Used only by
tests/golden/make_golden.py."""
import struct
import zlib

import numpy as np


class MiniHDF5:
    def __init__(self, path):
        self.buf = open(path, "rb").read()
        b = self.buf
        assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0, "only superblock version 0"
        assert b[13] == 8 and b[14] == 8, "only 8-byte offsets/lengths"
        # root symbol table entry at byte 56: name off, header addr, cache type, rsv, scratch
        _, self.root_header, cache, _ = struct.unpack_from("<QQII", b, 56)
        assert cache == 1, "root group must cache its B-tree/heap"
        btree, heap = struct.unpack_from("<QQ", b, 80)
        self.datasets = {}
        self._walk_group(btree, heap)

    # ---- groups ---------------------------------------------------------------------------
    def _heap_data(self, heap):
        assert self.buf[heap:heap + 4] == b"HEAP"
        _size, _free, addr = struct.unpack_from("<QQQ", self.buf, heap + 8)
        return addr

    def _walk_group(self, btree, heap):
        data = self._heap_data(heap)
        for snod in self._group_leaves(btree):
            assert self.buf[snod:snod + 4] == b"SNOD"
            nsym = struct.unpack_from("<H", self.buf, snod + 6)[0]
            for k in range(nsym):
                off = snod + 8 + 40 * k
                name_off, header, _cache = struct.unpack_from("<QQI", self.buf, off)
                end = self.buf.index(b"\x00", data + name_off)
                name = self.buf[data + name_off:end].decode()
                self.datasets[name] = header

    def _group_leaves(self, node):
        b = self.buf
        assert b[node:node + 4] == b"TREE" and b[node + 4] == 0
        level, used = b[node + 5], struct.unpack_from("<H", b, node + 6)[0]
        pos = node + 24
        out = []
        for _ in range(used):
            pos += 8  # key
            child = struct.unpack_from("<Q", b, pos)[0]
            pos += 8
            out += [child] if level == 0 else self._group_leaves(child)
        return out

    # ---- object headers -------------------------------------------------------------------
    def _messages(self, header):
        b = self.buf
        assert b[header] == 1, "only version-1 object headers"
        nmsg = struct.unpack_from("<H", b, header + 2)[0]
        size = struct.unpack_from("<I", b, header + 8)[0]
        blocks = [(header + 16, size)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(msgs) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = pos + 8
                if mtype == 0x10:  # continuation
                    blocks.append(struct.unpack_from("<QQ", b, body))
                msgs.append((mtype, body, msize))
                pos = body + msize
        return msgs

    def read(self, name):
        b = self.buf
        shape, dtype, layout, filters = (), None, None, []
        for mtype, body, msize in self._messages(self.datasets[name]):
            if mtype == 0x01:  # dataspace
                ver, rank = b[body], b[body + 1]
                off = body + (8 if ver == 1 else 4)
                shape = struct.unpack_from("<%dQ" % rank, b, off) if rank else ()
            elif mtype == 0x03:  # datatype
                cls = b[body] & 0x0F
                size = struct.unpack_from("<I", b, body + 4)[0]
                if cls == 1:
                    dtype = np.dtype("<f%d" % size)
                elif cls == 0:
                    signed = (b[body + 1] >> 3) & 1
                    dtype = np.dtype("<%s%d" % ("i" if signed else "u", size))
                else:
                    dtype = np.dtype("S%d" % size)
            elif mtype == 0x08:  # layout
                assert b[body] == 3, "only layout version 3"
                cls = b[body + 1]
                if cls == 0:
                    n = struct.unpack_from("<H", b, body + 2)[0]
                    layout = ("compact", body + 4, n)
                elif cls == 1:
                    layout = ("contiguous",) + struct.unpack_from("<QQ", b, body + 2)
                else:
                    nd = b[body + 2]
                    tree = struct.unpack_from("<Q", b, body + 3)[0]
                    dims = struct.unpack_from("<%dI" % nd, b, body + 11)
                    layout = ("chunked", tree, dims)
            elif mtype == 0x0B:  # filter pipeline (version 1)
                nfilt = b[body + 1]
                pos = body + 8
                for _ in range(nfilt):
                    fid, nlen, _fl, ncl = struct.unpack_from("<HHHH", b, pos)
                    pos += 8 + ((nlen + 7) // 8) * 8 + 4 * ncl + (4 if ncl % 2 else 0)
                    filters.append(fid)
        count = int(np.prod(shape)) if shape else 1
        if layout[0] == "compact":
            raw = b[layout[1]:layout[1] + layout[2]]
        elif layout[0] == "contiguous":
            raw = b[layout[1]:layout[1] + layout[2]]
        else:
            return self._read_chunked(layout[1], layout[2], shape, dtype, filters)
        arr = np.frombuffer(raw, dtype=dtype, count=count)
        return arr.reshape(shape) if shape else arr[0]

    def _read_chunked(self, tree, dims, shape, dtype, filters):
        chunk_shape = dims[:-1]
        out = np.zeros(shape, dtype=dtype)
        for offsets, addr, nbytes in self._chunk_leaves(tree, len(dims)):
            raw = self.buf[addr:addr + nbytes]
            for fid in reversed(filters):
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:  # shuffle
                    arr = np.frombuffer(raw, dtype=np.uint8).reshape(dtype.itemsize, -1)
                    raw = arr.T.tobytes()
                else:
                    raise NotImplementedError("HDF5 filter %d" % fid)
            chunk = np.frombuffer(raw, dtype=dtype).reshape(chunk_shape)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offsets, chunk_shape, shape))
            out[sel] = chunk[tuple(slice(0, s.stop - s.start) for s in sel)]
        return out

    def _chunk_leaves(self, node, nd):
        b = self.buf
        assert b[node:node + 4] == b"TREE" and b[node + 4] == 1
        level, used = b[node + 5], struct.unpack_from("<H", b, node + 6)[0]
        pos = node + 24
        key_size = 8 + 8 * nd
        out = []
        for _ in range(used):
            nbytes, _mask = struct.unpack_from("<II", b, pos)
            offsets = struct.unpack_from("<%dQ" % nd, b, pos + 8)[:-1]
            child = struct.unpack_from("<Q", b, pos + key_size)[0]
            pos += key_size + 8
            out += [(offsets, child, nbytes)] if level == 0 else self._chunk_leaves(child, nd)
        return out
