"""Minimal, dependency-free HDF5 reader for BubbleML-style files (h5py is not part of this image).

Supports what the reference's data files use (upstream bubbleformer/data/dataset.py:46-52 opens them with h5py and
reads whole datasets by name): superblock version 0, version-1 object headers, symbol-table groups (B-tree v1 +
local heap), and little-endian float32/float64/int datasets with CONTIGUOUS layout in the root group.  Chunked or
filtered (compressed) datasets raise NotImplementedError rather than being read wrongly.
"""
from __future__ import annotations

import struct
from typing import Dict, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class _File:
    def __init__(self, buf: bytes):
        if bytes(buf[:8]) != _SIG:
            raise ValueError("not an HDF5 file")
        if buf[8] != 0:
            raise NotImplementedError(f"HDF5 superblock version {buf[8]} (only version 0 is supported)")
        self.buf = buf
        self.O, self.L = buf[13], buf[14]                  # size of offsets / lengths
        if self.O != 8 or self.L != 8:
            raise NotImplementedError("HDF5 files with 4-byte offsets")
        p = 24                                             # after versions, sizes, K values and consistency flags
        self.base = self.u(p, 8)
        p += 4 * 8                                         # base, free-space, end-of-file, driver-info addresses
        # root group symbol table entry
        self.root_header = self.u(p + 8, 8)
        cache_type = self.u(p + 16, 4)
        self.root_btree = self.u(p + 24, 8) if cache_type == 1 else None
        self.root_heap = self.u(p + 32, 8) if cache_type == 1 else None

    def u(self, off: int, n: int) -> int:
        return int.from_bytes(bytes(self.buf[off:off + n]), "little")

    # ---- object headers (version 1) ----------------------------------------------------------
    def messages(self, addr: int):
        b = self.buf
        if b[addr] != 1:
            raise NotImplementedError(f"object header version {b[addr]}")
        nmsg = self.u(addr + 2, 2)
        size = self.u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize = self.u(p, 2), self.u(p + 2, 2)
                data = p + 8
                if mtype == 0x10:                          # continuation
                    blocks.append((self.u(data, 8), self.u(data + 8, 8)))
                out.append((mtype, data, msize))
                p = data + msize
        return out

    def dataset(self, addr: int) -> np.ndarray:
        shape = dtype = None
        layout = None
        for mtype, d, msize in self.messages(addr):
            b = self.buf
            if mtype == 0x1:                               # dataspace
                ver, rank = b[d], b[d + 1]
                p = d + (8 if ver == 1 else 4)
                shape = tuple(self.u(p + 8 * i, 8) for i in range(rank))
            elif mtype == 0x3:                             # datatype
                cls = b[d] & 0x0F
                bits0 = b[d + 1]
                size = self.u(d + 4, 4)
                if bits0 & 1:
                    raise NotImplementedError("big-endian HDF5 data")
                if cls == 1:
                    dtype = {4: "<f4", 8: "<f8"}[size]
                elif cls == 0:
                    dtype = ("<i" if (bits0 & 0x08) else "<u") + str(size)
                else:
                    raise NotImplementedError(f"HDF5 datatype class {cls}")
            elif mtype == 0x8:                             # data layout
                ver = b[d]
                if ver != 3:
                    raise NotImplementedError(f"data layout message version {ver}")
                if b[d + 1] != 1:
                    raise NotImplementedError("only contiguous HDF5 datasets are supported (chunked / compact found)")
                layout = (self.u(d + 2, 8), self.u(d + 10, 8))
            elif mtype == 0xB:
                raise NotImplementedError("filtered (compressed) HDF5 datasets are not supported")
        if shape is None or dtype is None or layout is None:
            raise ValueError("object is not a simple dataset")
        off, nbytes = layout
        n = int(np.prod(shape)) if shape else 1
        if n * np.dtype(dtype).itemsize != nbytes:
            raise ValueError("dataset size does not match its dataspace")
        start = self.base + off
        return np.asarray(self.buf[start:start + nbytes]).view(dtype).reshape(shape)

    # ---- root group ----------------------------------------------------------------------------
    def links(self) -> Dict[str, int]:
        if self.root_btree is None:
            raise NotImplementedError("root group without a cached symbol table")
        heap = self.root_heap
        if bytes(self.buf[heap:heap + 4]) != b"HEAP":
            raise ValueError("bad local heap")
        heap_data = self.u(heap + 8 + 2 * 8, 8)
        out: Dict[str, int] = {}

        def name_at(o: int) -> str:
            s = self.base + heap_data + o
            e = s
            while self.buf[e] != 0:
                e += 1
            return bytes(self.buf[s:e]).decode("utf-8")

        def walk(node: int) -> None:
            b = self.buf
            if bytes(b[node:node + 4]) == b"TREE":
                level, used = b[node + 5], self.u(node + 6, 2)
                p = node + 8 + 2 * 8                       # skip sibling addresses
                for i in range(used):
                    child = self.u(p + 8 + i * 16, 8)      # key (8) then child pointer (8)
                    walk(self.base + child)
                return
            if bytes(b[node:node + 4]) != b"SNOD":
                raise ValueError("bad group node")
            n = self.u(node + 6, 2)
            p = node + 8
            for i in range(n):
                e = p + i * 40
                out[name_at(self.u(e, 8))] = self.base + self.u(e + 8, 8)

        walk(self.base + self.root_btree)
        return out


class LazyFile:
    """Mapping name -> array over the root group that decodes an object only when it is asked for (like h5py reads by
    name): sub-groups, compound / string tables or chunked datasets elsewhere in the file do not get in the way of
    reading the plain fields.  The file is memory-mapped, so only the datasets actually read are paged in."""

    def __init__(self, path: str):
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")
        self._hf = _File(self._mm)
        self._links = self._hf.links()

    def keys(self):
        return self._links.keys()

    def __contains__(self, name: str) -> bool:
        return name in self._links

    def __iter__(self):
        return iter(self._links)

    def __len__(self) -> int:
        return len(self._links)

    def __getitem__(self, name: str) -> np.ndarray:
        if name not in self._links:
            raise KeyError(f"no object named {name!r} in the root group (have: {sorted(self._links)})")
        return self._hf.dataset(self._links[name])

    def items(self):
        return ((k, self[k]) for k in self._links)


def open_hdf5(path: str) -> LazyFile:
    return LazyFile(path)


def read_hdf5(path: str, fields=None) -> Dict[str, np.ndarray]:
    """Datasets of the root group as name -> array (views of the memory-mapped file).  `fields`: only these names are
    decoded; by default every object of the root group that is a simple contiguous dataset (others are skipped)."""
    f = LazyFile(path)
    if fields is not None:
        return {k: f[k] for k in fields}
    out = {}
    for k in f.keys():
        try:
            out[k] = f[k]
        except (NotImplementedError, ValueError):
            continue                      # not a plain dataset (group, compound table, chunked ...): read by name if needed
    return out


def dataset_shapes(path: str) -> Dict[str, Tuple[int, ...]]:
    return {k: v.shape for k, v in read_hdf5(path).items()}
