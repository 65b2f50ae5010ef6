"""Minimal read-only HDF5 reader for the recipe's statistics files (``stats.h5``), no h5py needed.

The reference reads two small float datasets per file: ``mean`` / ``scale`` of the vocoder
(jatts/vocoder/vocoder.py:47-54 via ``read_hdf5``, jatts/utils/utils.py:103-127) and ``mel_mean`` / ``mel_scale`` of
the text2mel model (jatts/bin/tts_decode.py:160-164); they are written by ``write_hdf5`` (utils.py:130-175:
``h5py.File(..., "w")`` / ``create_dataset(path, data=array)``), i.e. with h5py's default ``libver="earliest"``:

  * superblock version 0 (or 1), 8-byte offsets and lengths,
  * "old-style" groups: a symbol-table message -> version-1 B-tree ("TREE") -> symbol-table nodes ("SNOD") with the
    link names in a local heap ("HEAP"),
  * version-1 object headers (with continuation blocks),
  * datasets with a simple dataspace, a fixed-point or IEEE floating-point datatype and a CONTIGUOUS (or compact) data
    layout -- ``create_dataset(data=...)`` without chunks / compression.

That subset of the HDF5 File Format Specification (version 1.x structures) is what this module implements; anything
else (chunked or compressed datasets, new-style "OHDR" object headers, variable-length types) raises
``NotImplementedError`` naming the construct, so a file it cannot read fails loudly instead of yielding wrong
statistics.  h5py is absent from this image, so the reader is checked against files produced by an independent
writer of the same specification (tests/h5_writer.py) -- not against h5py output [format spec, unverified against h5py].
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5LiteFile:
    """``f = H5LiteFile(path); f["mean"]`` -> numpy array; ``f.keys()`` lists dataset paths (nested groups as "a/b")."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        self.base = self._find_superblock()
        self._parse_superblock()
        self._datasets: Dict[str, int] = {}
        self._walk_group(self.root_header, "")

    # ---- low level -------------------------------------------------------------------------------------------
    def _u(self, off: int, n: int) -> int:
        return int.from_bytes(self.buf[off:off + n], "little")

    def _find_superblock(self) -> int:
        off = 0
        while off < len(self.buf):
            if self.buf[off:off + 8] == _SIG:
                return off
            off = 512 if off == 0 else off * 2
        raise ValueError("not an HDF5 file (signature not found)")

    def _parse_superblock(self):
        b = self.base
        ver = self.buf[b + 8]
        if ver not in (0, 1):
            raise NotImplementedError(f"HDF5 superblock version {ver} (only the 'earliest' format, versions 0 and 1, is read)")
        self.so, self.sl = self.buf[b + 13], self.buf[b + 14]
        if self.so != 8 or self.sl != 8:
            raise NotImplementedError("HDF5 files with offsets / lengths other than 8 bytes")
        p = b + 24 + (4 if ver == 1 else 0)          # past group K values, consistency flags (+ indexed storage K in v1)
        self.base_addr = self._u(p, 8)
        p += 32                                        # base, free-space info, end of file, driver info addresses
        # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch pad
        self.root_header = self._u(p + 8, 8) + self.base_addr

    # ---- object headers --------------------------------------------------------------------------------------
    def _messages(self, addr: int) -> List[Tuple[int, int, int]]:
        """version-1 object header at ``addr`` -> [(message type, data offset, data size)] incl. continuation blocks"""
        if self.buf[addr:addr + 4] == b"OHDR":
            raise NotImplementedError("version-2 object headers (file written with libver='latest')")
        if self.buf[addr] != 1:
            raise ValueError(f"bad object header version {self.buf[addr]} at {addr}")
        n_msgs = self._u(addr + 2, 2)
        size = self._u(addr + 8, 4)
        out, blocks = [], [(addr + 16, size)]
        while blocks and len(out) < n_msgs:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < n_msgs:
                mtype, msize = self._u(p, 2), self._u(p + 2, 2)
                data = p + 8
                if mtype == 0x0010:                    # continuation: offset, length
                    blocks.append((self._u(data, 8) + self.base_addr, self._u(data + 8, 8)))
                out.append((mtype, data, msize))
                p = data + msize
        return out

    # ---- groups ----------------------------------------------------------------------------------------------
    def _heap_string(self, heap_addr: int, off: int) -> str:
        if self.buf[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        seg = self._u(heap_addr + 24, 8) + self.base_addr
        end = self.buf.index(b"\x00", seg + off)
        return self.buf[seg + off:end].decode("utf-8")

    def _btree_leaves(self, addr: int) -> List[int]:
        if self.buf[addr:addr + 4] != b"TREE":
            raise ValueError("bad B-tree signature")
        if self.buf[addr + 4] != 0:
            raise NotImplementedError("chunked datasets (raw-data B-tree)")
        level, used = self.buf[addr + 5], self._u(addr + 6, 2)
        p = addr + 8 + 16                              # past left / right sibling
        kids = [self._u(p + 8 + i * 16, 8) + self.base_addr for i in range(used)]   # key0 child0 key1 child1 ...
        if level == 0:
            return kids
        out = []
        for k in kids:
            out += self._btree_leaves(k)
        return out

    def _walk_group(self, header: int, prefix: str):
        msgs = self._messages(header)
        sym = [m for m in msgs if m[0] == 0x0011]
        if not sym:
            if any(m[0] in (0x0002, 0x0006) for m in msgs):
                raise NotImplementedError("new-style groups (link messages)")
            return
        btree, heap = self._u(sym[0][1], 8) + self.base_addr, self._u(sym[0][1] + 8, 8) + self.base_addr
        for snod in self._btree_leaves(btree):
            if self.buf[snod:snod + 4] != b"SNOD":
                raise ValueError("bad symbol table node signature")
            n = self._u(snod + 6, 2)
            for i in range(n):
                e = snod + 8 + i * 40
                name = self._heap_string(heap, self._u(e, 8))
                obj = self._u(e + 8, 8) + self.base_addr
                path = prefix + name
                kinds = {m[0] for m in self._messages(obj)}
                if 0x0011 in kinds:
                    self._walk_group(obj, path + "/")
                elif 0x0008 in kinds:
                    self._datasets[path] = obj

    # ---- datasets --------------------------------------------------------------------------------------------
    def keys(self):
        return list(self._datasets)

    def __contains__(self, name: str) -> bool:
        return name.strip("/") in self._datasets

    def __getitem__(self, name: str) -> np.ndarray:
        obj = self._datasets.get(name.strip("/"))
        if obj is None:
            raise KeyError(name)
        shape, dtype, raw = None, None, None
        for mtype, data, size in self._messages(obj):
            if mtype == 0x0001:                        # dataspace
                ver, rank, flags = self.buf[data], self.buf[data + 1], self.buf[data + 2]
                p = data + (8 if ver == 1 else 4)
                shape = tuple(self._u(p + 8 * i, 8) for i in range(rank))
            elif mtype == 0x0003:                      # datatype
                cls, bits0 = self.buf[data] & 0x0F, self.buf[data + 1]
                nbytes = self._u(data + 4, 4)
                order = ">" if bits0 & 1 else "<"
                if cls == 1:
                    if nbytes not in (2, 4, 8):
                        raise NotImplementedError(f"{nbytes}-byte floating point")
                    dtype = np.dtype(f"{order}f{nbytes}")
                elif cls == 0:
                    signed = bool(bits0 & 0x08)
                    dtype = np.dtype(f"{order}{'i' if signed else 'u'}{nbytes}")
                else:
                    raise NotImplementedError(f"HDF5 datatype class {cls}")
            elif mtype == 0x0008:                      # data layout
                ver = self.buf[data]
                if ver == 3:
                    lclass = self.buf[data + 1]
                    if lclass == 1:
                        raw = (self._u(data + 2, 8), self._u(data + 10, 8))
                    elif lclass == 0:
                        n = self._u(data + 2, 2)
                        raw = (data + 4 - self.base_addr, n)
                    else:
                        raise NotImplementedError("chunked dataset layout (write the statistics without chunks / compression)")
                elif ver in (1, 2):
                    rank, lclass = self.buf[data + 1], self.buf[data + 2]
                    if lclass != 1:
                        raise NotImplementedError("non-contiguous dataset layout (version 1/2 message)")
                    raw = (self._u(data + 8, 8), None)
                else:
                    raise NotImplementedError(f"data layout message version {ver}")
            elif mtype == 0x000B:
                raise NotImplementedError("filtered (compressed) dataset")
        if shape is None or dtype is None or raw is None:
            raise ValueError(f"dataset {name!r}: incomplete object header")
        count = int(np.prod(shape)) if shape else 1
        addr, nbytes = raw
        if addr == _UNDEF:
            return np.zeros(shape, dtype=dtype.newbyteorder("="))
        start = addr + self.base_addr
        arr = np.frombuffer(self.buf, dtype=dtype, count=count, offset=start).reshape(shape)
        return arr.astype(dtype.newbyteorder("="))


def read_hdf5(path, name: str) -> np.ndarray:
    """jatts/utils/utils.py:103-127 ``read_hdf5(hdf5_name, hdf5_path)`` without h5py."""
    f = H5LiteFile(path)
    if name not in f:
        raise KeyError(f"There is no such a data in hdf5 file. ({name})")
    return f[name]
