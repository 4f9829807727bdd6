"""Minimal read-only HDF5 reader for the chain files the reference writes.

The reference stores its chains through emcee's ``HDFBackend`` (``linna/sampler.py:322-368``: group ``mcmc`` with the
resizable datasets ``chain`` / ``chain_transformed`` / ``log_prob`` / ``accepted`` and the attribute ``iteration``) and
reads them back in ``read_chain_and_cut`` (``linna/util.py:68-94``).  h5py / emcee are not dependencies of this
package, so the part of the published HDF5 file format those files use is restated here: version-0 superblock,
symbol-table groups, version-1 object headers, contiguous and chunked float / integer datasets (optionally deflate /
shuffle filtered) and scalar or 1-D attributes.  Host-side file reading only -- nothing of the likelihood path.
``ChainStore`` (``sampler.py``) uses it to open a reference-written ``chemcee_256.h5`` / ``zeus_256.h5``; the reader is
pinned by the numbers the reference's own ``test_reading`` asserts on its shipped chain (``tests/test_main.py:51-52``).
"""
import struct
import zlib

import numpy as np


class H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        assert self.b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        ver = self.b[8]
        assert ver == 0, "only version-0 superblocks are supported (got %d)" % ver
        assert self.b[13] == 8 and self.b[14] == 8, "8-byte offsets and lengths expected"
        # signature 8, versions / sizes 8, leaf k 2, internal k 2, flags 4, base 8, free 8, eof 8, driver 8, root entry
        self.base = struct.unpack_from("<Q", self.b, 24)[0]
        self.root = self._symbol_entry(56)

    # ---- low level
    def _u(self, off, n):
        return int.from_bytes(self.b[off:off + n], "little")

    def _symbol_entry(self, off):
        name_off, header, cache = struct.unpack_from("<QQI", self.b, off)
        scratch = self.b[off + 24:off + 40]
        return dict(name_off=name_off, header=header, cache=cache, scratch=scratch)

    def _messages(self, addr):
        """(type, flags, payload bytes) of every message of the version-1 object header at addr."""
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", self.b, addr)
        assert ver == 1, "object header version %d" % ver
        out, blocks = [], [(addr + 16, size)]
        while blocks and len(out) < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", self.b, pos)
                body = self.b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:   # continuation
                    o, l = struct.unpack_from("<QQ", body, 0)
                    blocks.append((o + self.base, l))
                out.append((mtype, flags, body))
        return out

    def _heap_name(self, heap_addr, off):
        assert self.b[heap_addr:heap_addr + 4] == b"HEAP"
        data = struct.unpack_from("<Q", self.b, heap_addr + 24)[0] + self.base
        end = self.b.index(b"\x00", data + off)
        return self.b[data + off:end].decode()

    def _group_entries(self, btree, heap):
        """name -> symbol entry of a symbol-table group."""
        out = {}

        def walk(addr):
            assert self.b[addr:addr + 4] == b"TREE", "group B-tree expected"
            ntype, level, used = struct.unpack_from("<BBH", self.b, addr + 4)
            assert ntype == 0
            pos = addr + 24
            for i in range(used):
                child = struct.unpack_from("<Q", self.b, pos + 8)[0] + self.base
                pos += 16
                if level > 0:
                    walk(child)
                else:
                    assert self.b[child:child + 4] == b"SNOD"
                    n = struct.unpack_from("<H", self.b, child + 6)[0]
                    for j in range(n):
                        e = self._symbol_entry(child + 8 + 40 * j)
                        out[self._heap_name(heap, e["name_off"])] = e
        walk(btree)
        return out

    def _group_of(self, header):
        for mtype, _, body in self._messages(header):
            if mtype == 0x11:
                bt, hp = struct.unpack_from("<QQ", body, 0)
                return self._group_entries(bt + self.base, hp + self.base)
        raise KeyError("not a group")

    def _resolve(self, path):
        header = self.root["header"] + self.base
        for part in [p for p in path.split("/") if p]:
            header = self._group_of(header)[part]["header"] + self.base
        return header

    # ---- public
    def keys(self, path="/"):
        return sorted(self._group_of(self._resolve(path)))

    @staticmethod
    def _dtype(body):
        cls, bits0, size = body[0] & 0x0F, body[1], struct.unpack_from("<I", body, 4)[0]
        order = ">" if (bits0 & 1) else "<"
        if cls == 1:
            return np.dtype("%sf%d" % (order, size))
        if cls == 0:
            signed = (bits0 >> 3) & 1
            return np.dtype("%s%s%d" % (order, "i" if signed else "u", size))
        raise NotImplementedError("datatype class %d" % cls)

    @staticmethod
    def _dims(body):
        ver, rank, flags = body[0], body[1], body[2]
        off = 8 if ver == 1 else 4
        return tuple(struct.unpack_from("<Q", body, off + 8 * i)[0] for i in range(rank))

    def attrs(self, path):
        out = {}
        for mtype, _, body in self._messages(self._resolve(path)):
            if mtype != 0x0C:
                continue
            ver, _, nsz, tsz, ssz = struct.unpack_from("<BBHHH", body, 0)
            assert ver == 1, "attribute message version %d" % ver
            pad = lambda n: (n + 7) & ~7
            pos = 8
            name = body[pos:pos + nsz].split(b"\x00")[0].decode()
            pos += pad(nsz)
            tbody = body[pos:pos + tsz]
            pos += pad(tsz)
            sbody = body[pos:pos + ssz]
            pos += pad(ssz)
            try:
                dt = self._dtype(tbody)
            except NotImplementedError:
                continue
            dims = self._dims(sbody)
            n = int(np.prod(dims)) if dims else 1
            val = np.frombuffer(body[pos:pos + n * dt.itemsize], dt).reshape(dims)
            out[name] = val[()] if not dims else val
        return out

    def dataset(self, path):
        dt = dims = layout = None
        filters = []
        for mtype, _, body in self._messages(self._resolve(path)):
            if mtype == 0x01:
                dims = self._dims(body)
            elif mtype == 0x03:
                dt = self._dtype(body)
            elif mtype == 0x08:
                layout = body
            elif mtype == 0x0B:
                ver, nf = body[0], body[1]
                pos = 8 if ver == 1 else 2
                for _ in range(nf):
                    fid, nlen, _, ncd = struct.unpack_from("<HHHH", body, pos)
                    pos += 8 + ((nlen + 7) & ~7 if ver == 1 else nlen) + 4 * ncd
                    if ver == 1 and ncd % 2:
                        pos += 4
                    filters.append(fid)
        assert dt is not None and dims is not None and layout is not None, "not a simple dataset"
        assert layout[0] == 3, "layout message version %d" % layout[0]
        n = int(np.prod(dims)) if dims else 1
        if layout[1] == 1:   # contiguous
            addr, size = struct.unpack_from("<QQ", layout, 2)
            if addr == 0xFFFFFFFFFFFFFFFF:
                return np.zeros(dims, dt)
            return np.frombuffer(self.b[addr + self.base:addr + self.base + n * dt.itemsize], dt).reshape(dims).copy()
        assert layout[1] == 2, "layout class %d" % layout[1]
        rank1 = layout[2]
        btree = struct.unpack_from("<Q", layout, 3)[0]
        cdims = struct.unpack_from("<%dI" % rank1, layout, 11)
        chunk = cdims[:-1]
        assert cdims[-1] == dt.itemsize and len(chunk) == len(dims)
        out = np.zeros(dims, dt)
        if btree == 0xFFFFFFFFFFFFFFFF:
            return out

        def walk(addr):
            assert self.b[addr:addr + 4] == b"TREE", "chunk B-tree expected"
            ntype, level, used = struct.unpack_from("<BBH", self.b, addr + 4)
            assert ntype == 1
            pos = addr + 24
            keysz = 8 + 8 * rank1
            for i in range(used):
                csize, fmask = struct.unpack_from("<II", self.b, pos)
                offs = struct.unpack_from("<%dQ" % rank1, self.b, pos + 8)[:-1]
                child = struct.unpack_from("<Q", self.b, pos + keysz)[0] + self.base
                pos += keysz + 8
                if level > 0:
                    walk(child)
                    continue
                raw = self.b[child:child + csize]
                for fid in reversed(filters):
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:   # shuffle
                        a = np.frombuffer(raw, np.uint8).reshape(dt.itemsize, -1)
                        raw = a.T.tobytes()
                    else:
                        raise NotImplementedError("filter %d" % fid)
                blk = np.frombuffer(raw, dt, count=int(np.prod(chunk))).reshape(chunk)
                sl = tuple(slice(o, min(o + c, d)) for o, c, d in zip(offs, chunk, dims))
                out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
        walk(btree + self.base)
        return out
