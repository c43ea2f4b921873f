"""A minimal HDF5 reader (test infrastructure): superblock version 0, old-style groups (version-1 B-tree + local heap +
symbol-table nodes), version-1 object headers, contiguous and chunked (unfiltered) datasets of fixed-point / IEEE float /
fixed-length string types, version-1 attributes.  That is the subset the reference's PSPWriter produces with the HDF5
1.10 library's default settings (cpp/lib/PSPHDF5.ipp) -- checked here against the reference's own fixtures
cpp/test/inputs/unstruct_nodal_pencil*.h5 -- and therefore the subset host/psp_hdf5.hpp writes.  No HDF5 library, h5py or
pytables exists in this image; the file format is followed from the HDF5 File Format Specification version 2.0."""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


class Dataset:
    def __init__(self, name, shape, dtype, strlen, layout, attrs):
        self.name, self.shape, self.dtype, self.strlen, self.layout, self.attrs = name, shape, dtype, strlen, layout, attrs
        self.filters = []       # filter ids of the pipeline message (0x000B), in order
        self.data = None


class Group:
    def __init__(self, name):
        self.name, self.attrs, self.children = name, {}, {}

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            node = node.children[part]
        return node


class File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("not an HDF5 file")
        if b[8] != 0 or b[13] != 8 or b[14] != 8:
            raise H5Error("only superblock version 0 with 8-byte offsets / lengths")
        self.leaf_k, self.int_k = struct.unpack_from("<HH", b, 16)
        self.base, _, self.eof, _ = struct.unpack_from("<4Q", b, 24)
        if self.eof > len(b):
            raise H5Error("end-of-file address beyond the file")
        _, ohdr, cache, _ = struct.unpack_from("<QQII", b, 56)
        self.root = self._read_object("/", ohdr)

    # ---- object headers
    def _messages(self, addr):
        b = self.b
        ver, _, nmsg, refcnt, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise H5Error(f"object header version {ver} at {addr}")
        out = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, pos)
                data = b[pos + 8: pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:      # continuation
                    caddr, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((caddr, clen))
                out.append((mtype, data))
        return out

    @staticmethod
    def _parse_dataspace(d):
        ver, rank, flags = struct.unpack_from("<BBB", d, 0)
        if ver == 1:
            off = 8
        elif ver == 2:
            off = 4
        else:
            raise H5Error(f"dataspace version {ver}")
        return tuple(struct.unpack_from(f"<{rank}Q", d, off)) if rank else ()

    @staticmethod
    def _parse_datatype(d):
        cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", d, 0)
        cls, ver = cv & 15, cv >> 4
        if cls == 0:
            signed = bool(b0 & 8)
            if b0 & 1:
                raise H5Error("big-endian fixed point")
            return np.dtype(("<i" if signed else "<u") + str(size)), 0, 8 + 4
        if cls == 1:
            if b0 & 1:
                raise H5Error("big-endian float")
            return np.dtype("<f" + str(size)), 0, 8 + 12
        if cls == 3:
            return np.dtype(f"S{size}"), size, 8
        raise H5Error(f"datatype class {cls}")

    @staticmethod
    def _parse_pipeline(d):
        ver, nf = d[0], d[1]
        if ver != 1:
            raise H5Error(f"filter pipeline version {ver}")
        pos, ids = 8, []
        for _ in range(nf):
            fid, nlen, flags, ncd = struct.unpack_from("<HHHH", d, pos)
            pos += 8 + ((nlen + 7) & ~7) + 4 * ncd + (4 if ncd % 2 else 0)
            ids.append(fid)
        return ids

    def _parse_attribute(self, d):
        ver, _, nlen, tlen, slen = struct.unpack_from("<BBHHH", d, 0)
        if ver != 1:
            raise H5Error(f"attribute version {ver}")
        pad = lambda n: (n + 7) & ~7
        pos = 8
        name = d[pos: pos + nlen].split(b"\0")[0].decode()
        pos += pad(nlen)
        dt, strlen, _ = self._parse_datatype(d[pos: pos + tlen])
        pos += pad(tlen)
        shape = self._parse_dataspace(d[pos: pos + slen])
        pos += pad(slen)
        n = int(np.prod(shape)) if shape else 1
        val = np.frombuffer(d, dt, n, pos).copy()
        if strlen:
            val = [v.split(b"\0")[0].decode() for v in val]
        return name, val

    def _read_object(self, name, addr):
        msgs = self._messages(addr)
        attrs = {}
        shape = dtype = layout = None
        strlen = 0
        symtab = None
        filters = []
        for mtype, d in msgs:
            if mtype == 0x000C:
                k, v = self._parse_attribute(d)
                attrs[k] = v
            elif mtype == 0x0001:
                shape = self._parse_dataspace(d)
            elif mtype == 0x0003:
                dtype, strlen, _ = self._parse_datatype(d)
            elif mtype == 0x0008:
                layout = d
            elif mtype == 0x0011:
                symtab = struct.unpack_from("<QQ", d, 0)
            elif mtype == 0x000B:
                filters = self._parse_pipeline(d)
        if symtab is not None:
            g = Group(name)
            g.attrs = attrs
            for cname, caddr in self._group_entries(*symtab):
                g.children[cname] = self._read_object(cname, caddr)
            return g
        if shape is None or dtype is None or layout is None:
            raise H5Error(f"object {name}: not a group and not a complete dataset")
        ds = Dataset(name, shape, dtype, strlen, layout, attrs)
        ds.filters = filters
        ds.data = self._read_data(ds)
        return ds

    def addr_of(self, path):
        """Object-header address of the object at `path` ("/Grid/x")."""
        _, addr, _, _ = struct.unpack_from("<QQII", self.b, 56)
        for part in [p for p in path.split("/") if p]:
            st = [d for t, d in self._messages(addr) if t == 0x0011]
            if not st:
                raise H5Error(f"{part}: parent is not a group")
            ents = dict(self._group_entries(*struct.unpack_from("<QQ", st[0], 0)))
            addr = ents[part]
        return addr

    # ---- groups
    def _heap_name(self, heap, off):
        b = self.b
        if b[heap: heap + 4] != b"HEAP":
            raise H5Error("local heap signature")
        dseg = struct.unpack_from("<Q", b, heap + 24)[0]
        end = b.index(b"\0", dseg + off)
        return b[dseg + off: end].decode()

    def _group_entries(self, btree, heap):
        b = self.b
        out = []

        def node(addr):
            if b[addr: addr + 4] == b"SNOD":
                n = struct.unpack_from("<H", b, addr + 6)[0]
                for i in range(n):
                    noff, ohdr = struct.unpack_from("<QQ", b, addr + 8 + 40 * i)
                    out.append((self._heap_name(heap, noff), ohdr))
                return
            if b[addr: addr + 4] != b"TREE":
                raise H5Error(f"B-tree signature at {addr}")
            ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
            if ntype != 0:
                raise H5Error("group B-tree expected")
            pos = addr + 24
            for i in range(used):
                child = struct.unpack_from("<Q", b, pos + 8)[0]      # key, child, key, child ...
                node(child)
                pos += 16

        node(btree)
        return out

    # ---- raw data
    def _read_data(self, ds):
        b, d = self.b, ds.layout
        ver, cls = d[0], d[1]
        if ver != 3:
            raise H5Error(f"layout version {ver}")
        n = int(np.prod(ds.shape)) if ds.shape else 1
        if cls == 1:            # contiguous
            addr, size = struct.unpack_from("<QQ", d, 2)
            if addr == UNDEF:
                return np.zeros(ds.shape, ds.dtype)
            return np.frombuffer(b, ds.dtype, n, addr).reshape(ds.shape).copy()
        if cls == 0:            # compact
            size = struct.unpack_from("<H", d, 2)[0]
            return np.frombuffer(d, ds.dtype, n, 4).reshape(ds.shape).copy()
        if cls == 2:            # chunked, version-1 B-tree of raw-data chunks
            rank1 = d[2]
            addr = struct.unpack_from("<Q", d, 3)[0]
            cdims = struct.unpack_from(f"<{rank1}I", d, 11)
            chunk = cdims[:-1]
            out = np.zeros(ds.shape, ds.dtype)
            if addr == UNDEF:
                return out

            def node(a):
                if b[a: a + 4] != b"TREE":
                    raise H5Error("chunk B-tree signature")
                ntype, level, used = struct.unpack_from("<BBH", b, a + 4)
                ksz = 8 + 8 * rank1
                pos = a + 24
                for i in range(used):
                    csize, fmask = struct.unpack_from("<II", b, pos)
                    offs = struct.unpack_from(f"<{rank1}Q", b, pos + 8)[:-1]
                    child = struct.unpack_from("<Q", b, pos + ksz)[0]
                    if level > 0:
                        node(child)
                    else:
                        raw = b[child: child + csize]
                        nbytes = int(np.prod(chunk)) * ds.dtype.itemsize
                        if ds.filters:      # the reference's fixtures carry a deflate pipeline (filter id 1) on "frames"
                            if ds.filters != [1] or fmask:
                                raise H5Error(f"filter pipeline {ds.filters} (mask {fmask}) is not supported")
                            raw = zlib.decompress(raw)
                        if len(raw) != nbytes:
                            raise H5Error("chunk size mismatch")
                        blk = np.frombuffer(raw, ds.dtype).reshape(chunk)
                        sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, ds.shape))
                        out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
                    pos += ksz + 8

            node(addr)
            return out
        raise H5Error(f"layout class {cls}")


def walk(g, prefix=""):
    """[(path, object)] of everything below a group, depth first, names sorted."""
    out = []
    for k in sorted(g.children):
        c = g.children[k]
        out.append((prefix + "/" + k, c))
        if isinstance(c, Group):
            out += walk(c, prefix + "/" + k)
    return out
