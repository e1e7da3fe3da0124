"""Self-contained reader / writer for the classic HDF5 on-disk format, covering exactly what the tomogram files of the
reference need (seam B5: ``h5py.File(path, "w")`` + ``create_dataset`` with and without ``compression="gzip"`` in
run/dino_features.py:109-153, ``h5py.File(path)`` + ``fh[key][()]`` in datasets/vit_dataset.py:83-84 and
datasets/tomo_dataset.py:110-123).

``h5py`` with its default ``libver="earliest"`` writes -- and every libhdf5 since 1.6 reads -- this subset:

    superblock version 0 (1 accepted on read), 8-byte offsets and lengths
    groups as symbol tables: version-1 object header -> symbol-table message -> B-tree v1 (node type 0) + local heap
        + symbol-table nodes (``SNOD``)
    datasets as version-1 object headers: dataspace v1, datatype v1 (fixed-point / IEEE float), fill value v2,
        layout v3 (contiguous | chunked | compact on read), filter pipeline v1 (deflate; shuffle on read)
    chunk index: B-tree v1 (node type 1), keys = (bytes, filter mask, element offsets)

The structures follow the published "HDF5 File Format Specification Version 2.0" (sections II.A superblock, III.A
B-trees, III.B symbol-table nodes, III.D local heaps, IV.A object headers and messages 0x01 0x03 0x05 0x08 0x0B
0x10 0x11). Neither h5py nor libhdf5 is in the build image; the READER is pinned against the one libhdf5-written
file the image holds (a MATLAB 7.3 file in scipy's test data, tests/test_hdf5_classic.py) and the writer against the
reader plus a byte-level walk of everything libhdf5 validates on open (node sizes, end-of-file address, key order).
When ``h5py`` is importable, host/hdf.py uses it instead and this module is only the cross-check.
"""
from __future__ import annotations

import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from itertools import product

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
GROUP_LEAF_K = 4        # libhdf5 defaults: a symbol-table node holds <= 2 * 4 entries,
GROUP_INTERNAL_K = 16   # a group B-tree node <= 2 * 16 children,
CHUNK_K = 32            # a chunk B-tree node <= 2 * 32 children (implied by superblock version 0)
FREE_NULL = 1           # local heap: "no free block"

MSG_DATASPACE, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LAYOUT, MSG_FILTERS = 0x01, 0x03, 0x04, 0x05, 0x08, 0x0B
MSG_CONTINUATION, MSG_SYMBOL_TABLE = 0x10, 0x11
FILTER_DEFLATE, FILTER_SHUFFLE, FILTER_FLETCHER32 = 1, 2, 3


class Hdf5FormatError(ValueError):
    """The file uses a part of HDF5 outside the classic subset (or is damaged)."""


def is_hdf5(path) -> bool:
    """True when an HDF5 signature sits at one of the offsets the format allows (0, 512, 1024, ...: user blocks)."""
    try:
        with open(path, "rb") as fh:
            return _find_superblock(fh) is not None
    except OSError:
        return False


def _find_superblock(fh) -> int | None:
    off = 0
    size = os.fstat(fh.fileno()).st_size
    while off + 8 <= size:
        fh.seek(off)
        if fh.read(8) == SIGNATURE:
            return off
        off = 512 if off == 0 else off * 2
    return None


def _pad8(n: int) -> int:
    return (n + 7) & ~7


# ------------------------------------------------------------------------------------------------------------------
# datatypes
# ------------------------------------------------------------------------------------------------------------------

_FLOAT_LAYOUT = {2: (15, 10, 5, 10, 15), 4: (31, 23, 8, 23, 127), 8: (63, 52, 11, 52, 1023)}  # sign, eloc, esz, msz, bias


def _encode_datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.byteorder == ">":
        raise Hdf5FormatError("big-endian arrays are not written")
    if dt.kind in "iu" or dt.kind == "b":
        size = dt.itemsize
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10 | 0, bits0, 0, 0, size) + struct.pack("<HH", 0, 8 * size)
    if dt.kind == "f" and dt.itemsize in _FLOAT_LAYOUT:
        sign, eloc, esz, msz, bias = _FLOAT_LAYOUT[dt.itemsize]
        head = struct.pack("<BBBBI", 0x10 | 1, 0x20, sign, 0, dt.itemsize)  # mantissa normalisation 2: implied msb
        return head + struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, eloc, esz, 0, msz, bias)
    raise Hdf5FormatError(f"dtype {dt} is outside the tomogram layout (integers and IEEE floats)")


def _decode_datatype(buf: bytes) -> np.dtype:
    cls, version = buf[0] & 0x0F, buf[0] >> 4
    bits0, bits1 = buf[1], buf[2]
    size = struct.unpack_from("<I", buf, 4)[0]
    order = ">" if bits0 & 1 else "<"
    if version not in (1, 2, 3):
        raise Hdf5FormatError(f"datatype message version {version}")
    if cls == 0:
        return np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}")
    if cls == 1:
        eloc, esz, _, msz, bias = struct.unpack_from("<BBBBI", buf, 12)
        if size in _FLOAT_LAYOUT and (bits1, eloc, esz, msz, bias) == _FLOAT_LAYOUT[size]:
            return np.dtype(f"{order}f{size}")
        raise Hdf5FormatError("non-IEEE floating-point datatype")
    raise Hdf5FormatError(f"datatype class {cls} (only fixed-point and floating-point datasets are read)")


def undo_filters(blob: bytes, filters: list[tuple[int, tuple[int, ...]]], mask: int, itemsize: int) -> bytes:
    """A stored chunk back to raw element bytes: the pipeline's filters in reverse order, skipping those the chunk's
    filter mask switched off. deflate, shuffle and fletcher32 (the checksum is dropped, not verified) are what h5py
    offers without plug-ins; anything else needs h5py itself."""
    for idx in range(len(filters) - 1, -1, -1):
        if mask >> idx & 1:
            continue
        fid, vals = filters[idx]
        if fid == FILTER_DEFLATE:
            blob = zlib.decompress(blob)
        elif fid == FILTER_SHUFFLE:
            width = vals[0] if vals else itemsize
            blob = np.frombuffer(blob, np.uint8).reshape(width, -1).T.tobytes()
        elif fid == FILTER_FLETCHER32:
            blob = blob[:-4]
        else:
            raise Hdf5FormatError(f"filter {fid} needs h5py")
    return blob


# ------------------------------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------------------------------

def default_chunks(shape: tuple[int, ...], itemsize: int, target_bytes: int = 1 << 20) -> tuple[int, ...]:
    """Chunk shape for a compressed dataset: whole trailing planes, as many leading rows as fit ~1 MiB (any chunking is
    valid HDF5; h5py's own guess differs and does not have to be mirrored: readers follow the B-tree)."""
    if not shape:
        raise Hdf5FormatError("scalar datasets cannot be chunked")
    chunk = [max(1, int(s)) for s in shape]
    for axis in range(len(shape)):
        rest = int(np.prod(chunk[axis + 1:], dtype=np.int64)) * itemsize
        if rest <= target_bytes:
            chunk[axis] = max(1, min(chunk[axis], target_bytes // max(rest, 1)))
            break
        chunk[axis] = 1
    return tuple(chunk)


class _Writer:
    def __init__(self, fh):
        self.fh = fh
        self.end = 96  # superblock (56) + root symbol-table entry (40)

    def alloc(self, size: int) -> int:
        addr = _pad8(self.end)
        self.end = addr + size
        return addr

    def put(self, addr: int, data) -> None:
        self.fh.seek(addr)
        self.fh.write(data)

    def add(self, data) -> int:
        addr = self.alloc(len(data) if not isinstance(data, memoryview) else data.nbytes)
        self.put(addr, data)
        return addr

    # -- object headers ------------------------------------------------------------------------------------------
    def object_header(self, messages: list[tuple[int, int, bytes]]) -> int:
        body = b""
        for mtype, flags, data in messages:
            data = data + b"\x00" * (_pad8(len(data)) - len(data))
            body += struct.pack("<HHB3x", mtype, len(data), flags) + data
        head = struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body))  # version 1, ref count 1, 16-byte prefix
        return self.add(head + body)

    # -- datasets ------------------------------------------------------------------------------------------------
    def dataset(self, arr: np.ndarray, gzip_level: int | None, chunks: tuple[int, ...] | None, pool) -> int:
        arr = np.asarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        if not arr.flags.c_contiguous:  # (np.ascontiguousarray would turn a 0-d array into a 1-d one)
            arr = np.ascontiguousarray(arr)
        rank = arr.ndim
        space = struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape)
        msgs = [(MSG_DATASPACE, 0, space), (MSG_DATATYPE, 1, _encode_datatype(arr.dtype))]
        if (gzip_level is None and chunks is None) or rank == 0 or arr.size == 0:
            addr = self.add(memoryview(arr).cast("B")) if arr.size else UNDEF
            msgs.append((MSG_FILL, 1, struct.pack("<BBBBI", 2, 2, 2, 1, 0)))  # late allocation, default fill value
            msgs.append((MSG_LAYOUT, 0, struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)))
            return self.object_header(msgs)
        chunks = tuple(int(c) for c in (chunks or default_chunks(arr.shape, arr.itemsize)))
        if len(chunks) == rank:  # libhdf5 refuses chunks larger than a fixed-size dataset
            chunks = tuple(min(c, s) for c, s in zip(chunks, arr.shape))
        if len(chunks) != rank or any(c < 1 for c in chunks):
            raise Hdf5FormatError(f"chunk shape {chunks} does not fit a rank-{rank} dataset")
        grid = [range(0, s, c) for s, c in zip(arr.shape, chunks)]

        def pack(origin):
            block = arr[tuple(slice(o, o + c) for o, c in zip(origin, chunks))]
            if block.shape != chunks:  # edge chunks are stored whole, the overhang is fill value
                full = np.zeros(chunks, arr.dtype)
                full[tuple(slice(0, s) for s in block.shape)] = block
                block = full
            block = np.ascontiguousarray(block)
            return zlib.compress(block, gzip_level) if gzip_level is not None else block.tobytes()

        origins = list(product(*grid))  # C order = the lexicographic key order of the chunk B-tree
        entries = []
        for origin, blob in zip(origins, pool.map(pack, origins)):
            entries.append((origin, len(blob), self.add(blob)))
        btree = self.chunk_btree(entries, chunks, arr.itemsize)
        filters = struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", FILTER_DEFLATE, 8, 1, 1) + b"deflate\x00" \
            + struct.pack("<I4x", gzip_level or 0)
        msgs.append((MSG_FILL, 1, struct.pack("<BBBBI", 2, 3, 2, 1, 0)))  # incremental allocation
        if gzip_level is not None:
            msgs.append((MSG_FILTERS, 1, filters))
        msgs.append((MSG_LAYOUT, 0, struct.pack("<BBBQ", 3, 2, rank + 1, btree)
                     + b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", arr.itemsize)))
        return self.object_header(msgs)

    def chunk_btree(self, entries, chunks, itemsize) -> int:
        """B-tree v1 over (element offsets..., 0) keys; <= 2 * CHUNK_K children per node, full-size nodes."""
        rank1 = len(chunks) + 1
        key_size = 8 + 8 * rank1
        node_size = 24 + (2 * CHUNK_K + 1) * key_size + 2 * CHUNK_K * 8

        def key(origin, nbytes):
            return struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in origin) + struct.pack("<Q", 0)

        last = entries[-1][0]
        max_key = key(tuple(o + c for o, c in zip(last, chunks)), 0)
        level_items = [(key(origin, nbytes), addr) for origin, nbytes, addr in entries]  # (left key, child address)
        level = 0
        while True:
            groups = [level_items[i:i + 2 * CHUNK_K] for i in range(0, len(level_items), 2 * CHUNK_K)]
            addrs = [self.alloc(node_size) for _ in groups]
            for gi, group in enumerate(groups):
                right = groups[gi + 1][0][0] if gi + 1 < len(groups) else max_key
                body = struct.pack("<4sBBHQQ", b"TREE", 1, level, len(group),
                                   addrs[gi - 1] if gi else UNDEF, addrs[gi + 1] if gi + 1 < len(groups) else UNDEF)
                for k, child in group:
                    body += k + struct.pack("<Q", child)
                body += right
                self.put(addrs[gi], body + b"\x00" * (node_size - len(body)))
            if len(groups) == 1:
                return addrs[0]
            level_items = [(group[0][0], addr) for group, addr in zip(groups, addrs)]
            level += 1

    # -- groups --------------------------------------------------------------------------------------------------
    def group(self, children: dict[str, tuple[int, tuple[int, int] | None]]) -> tuple[int, int, int]:
        """children: name -> (object header address, (btree, heap) for groups | None). Returns (header, btree, heap)."""
        names = sorted(children, key=lambda s: s.encode())
        if len(names) > 2 * GROUP_LEAF_K * 2 * GROUP_INTERNAL_K:
            raise Hdf5FormatError("more links in one group than a single-level group B-tree holds")
        heap = bytearray(8)  # offset 0: the empty string every group B-tree's first key points at
        offsets = {}
        for n in names:
            raw = n.encode() + b"\x00"
            offsets[n] = len(heap)
            heap += raw + b"\x00" * (_pad8(len(raw)) - len(raw))
        heap_data = self.add(bytes(heap))
        heap_addr = self.add(struct.pack("<4sB3xQQQ", b"HEAP", 0, len(heap), FREE_NULL, heap_data))
        node_cap = 2 * GROUP_LEAF_K
        keys, nodes = [0], []
        for i in range(0, max(len(names), 1), node_cap):
            part = names[i:i + node_cap]
            body = struct.pack("<4sBBH", b"SNOD", 1, 0, len(part))
            for n in part:
                addr, sub = children[n]
                if sub is None:
                    body += struct.pack("<QQII16x", offsets[n], addr, 0, 0)
                else:
                    body += struct.pack("<QQIIQQ", offsets[n], addr, 1, 0, sub[0], sub[1])
            nodes.append(self.add(body + b"\x00" * (8 + node_cap * 40 - len(body))))
            keys.append(offsets[part[-1]] if part else 0)
        tree_size = 24 + (2 * GROUP_INTERNAL_K + 1) * 8 + 2 * GROUP_INTERNAL_K * 8
        body = struct.pack("<4sBBHQQ", b"TREE", 0, 0, len(nodes) if names else 0, UNDEF, UNDEF)
        if names:
            for k, child in zip(keys, nodes):
                body += struct.pack("<QQ", k, child)
            body += struct.pack("<Q", keys[-1])
        btree = self.add(body + b"\x00" * (tree_size - len(body)))
        header = self.object_header([(MSG_SYMBOL_TABLE, 0, struct.pack("<QQ", btree, heap_addr))])
        return header, btree, heap_addr


def write_file(path, datasets: dict[str, np.ndarray], gzip: dict[str, int | None] | None = None,
               chunks: dict[str, tuple[int, ...]] | None = None, threads: int = 8) -> None:
    """One HDF5 file holding ``datasets`` (keys may contain ``/``: intermediate groups are created, as
    ``h5py.File.create_dataset`` does). ``gzip[key]`` = deflate level for a chunked, compressed dataset (h5py's
    ``compression="gzip"`` is level 4); keys absent from it are stored contiguous and raw, unless ``chunks[key]`` asks
    for a chunked (uncompressed) layout, e.g. ``dino_features`` in depth slabs so that a training crop reads only its
    slabs. Chunks are compressed on ``threads`` threads (zlib releases the GIL)."""
    gzip = gzip or {}
    chunks = chunks or {}
    tree: dict = {}
    for key, arr in datasets.items():
        parts = [p for p in key.split("/") if p]
        if not parts:
            raise Hdf5FormatError("empty dataset name")
        node = tree
        for p in parts[:-1]:
            node = node.setdefault(p, {})
            if not isinstance(node, dict):
                raise Hdf5FormatError(f"{key}: {p} is a dataset, not a group")
        if isinstance(node.get(parts[-1]), dict):
            raise Hdf5FormatError(f"{key} is already a group")
        node[parts[-1]] = key
    with open(path, "wb") as fh, ThreadPoolExecutor(max(1, threads)) as pool:
        w = _Writer(fh)

        def emit(node) -> tuple[int, int, int]:
            children = {}
            for name, item in node.items():
                if isinstance(item, dict):
                    header, btree, heap = emit(item)
                    children[name] = (header, (btree, heap))
                else:
                    children[name] = (w.dataset(datasets[item], gzip.get(item), chunks.get(item), pool), None)
            return w.group(children)

        header, btree, heap = emit(tree)
        eof = _pad8(w.end)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQIIQQ", 0, header, 1, 0, btree, heap)  # root symbol-table entry, cached B-tree / heap
        w.put(0, sb)
        fh.truncate(eof)


# ------------------------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------------------------

class DatasetInfo:
    __slots__ = ("shape", "dtype", "layout", "address", "size", "chunks", "btree", "filters", "compact", "error")

    def __init__(self):
        self.shape = self.dtype = self.layout = self.address = self.size = self.chunks = self.btree = self.compact = None
        self.error: str | None = None  # why this dataset cannot be read (the others of the file still can)
        self.filters: list[tuple[int, tuple[int, ...]]] = []


class File:
    """Read side: ``File(path).keys()`` lists every dataset as ``group/member`` paths, ``read(key)`` returns it."""

    def __init__(self, path):
        self.fh = open(path, "rb")
        try:
            base = _find_superblock(self.fh)
            if base is None:
                raise Hdf5FormatError(f"{path}: no HDF5 signature")
            self.fh.seek(base + 8)
            version = self.fh.read(1)[0]
            if version not in (0, 1):
                raise Hdf5FormatError(f"{path}: superblock version {version} (written with libver='latest'?) is outside "
                                      "the classic subset this reader covers; install h5py for such files")
            head = self._at(base + 8, 16 if version == 0 else 20, absolute=True)
            if head[5] != 8 or head[6] != 8:
                raise Hdf5FormatError("only 8-byte offsets / lengths are read")
            self.leaf_k, self.internal_k = struct.unpack_from("<HH", head, 8)
            self.chunk_k = struct.unpack_from("<H", head, 16)[0] if version == 1 else CHUNK_K
            pos = base + 8 + (16 if version == 0 else 20)
            self.base, _, self.eof, _ = struct.unpack("<QQQQ", self._at(pos, 32, absolute=True))
            if self.base == 0 and base:  # user block in front of a file whose addresses are relative to the superblock
                self.base = base
            root = self._at(pos + 32, 40, absolute=True)
            _, self.root_header, cache, _, bt, hp = struct.unpack("<QQIIQQ", root)
            self._datasets: dict[str, DatasetInfo] = {}
            self._walk_group(self.root_header, "", (bt, hp) if cache == 1 else None, 0)
        except Exception:
            self.fh.close()
            raise

    def close(self):
        self.fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def keys(self) -> list[str]:
        return list(self._datasets)

    def info(self, key: str) -> DatasetInfo:
        return self._datasets[key.strip("/")]

    # -- low level -----------------------------------------------------------------------------------------------
    def _at(self, addr: int, size: int, absolute: bool = False) -> bytes:
        self.fh.seek(addr if absolute else addr + self.base)
        data = self.fh.read(size)
        if len(data) != size:
            raise Hdf5FormatError(f"truncated file: {size} bytes wanted at {addr}, {len(data)} there")
        return data

    def _messages(self, addr: int) -> list[tuple[int, bytes]]:
        version, _, count, _, size = struct.unpack("<BBHII", self._at(addr, 12))
        if version != 1:
            raise Hdf5FormatError(f"object header version {version} at {addr}: only version 1 (classic) is read")
        blocks = [(addr + 16, size)]
        out: list[tuple[int, bytes]] = []
        while blocks and len(out) < count:
            pos, left = blocks.pop(0)
            while left >= 8 and len(out) < count:
                mtype, msize, _ = struct.unpack("<HHB", self._at(pos, 5))
                data = self._at(pos + 8, msize)
                if mtype == MSG_CONTINUATION:
                    blocks.append(struct.unpack("<QQ", data[:16]))
                out.append((mtype, data))
                pos += 8 + msize
                left -= 8 + msize
        return out

    def _heap_name(self, heap_addr: int, offset: int) -> str:
        sig, _, size, _, data = struct.unpack("<4sB3xQQQ", self._at(heap_addr, 32))
        if sig != b"HEAP":
            raise Hdf5FormatError("local heap signature")
        raw = self._at(data + offset, min(size - offset, 1024))
        return raw.split(b"\x00", 1)[0].decode()

    def _walk_group(self, header: int, prefix: str, cached, depth: int) -> None:
        if depth > 32:
            raise Hdf5FormatError("group nesting too deep (cycle?)")
        msgs = self._messages(header)
        table = next((d for t, d in msgs if t == MSG_SYMBOL_TABLE), None)
        if table is None:
            if any(t == MSG_LAYOUT for t, _ in msgs):
                try:
                    self._datasets[prefix.strip("/")] = self._dataset(msgs)
                except Hdf5FormatError as e:  # e.g. a string / compound dataset next to the tomogram's own ones
                    bad = DatasetInfo()
                    bad.error = str(e)
                    self._datasets[prefix.strip("/")] = bad
                return
            if any(t in (0x02, 0x06) for t, _ in msgs):  # link info / link messages
                raise Hdf5FormatError(f"group '{prefix or '/'}' is stored in the new link-message format (libver='latest' or "
                                      "track_order): outside the classic subset, needs h5py")
            if cached is None:
                return  # committed datatype: nothing the tomogram layout holds
            btree, heap = cached
        else:
            btree, heap = struct.unpack("<QQ", table[:16])
        for snod in self._group_leaves(btree):
            sig, _, _, used = struct.unpack("<4sBBH", self._at(snod, 8))
            if sig != b"SNOD":
                raise Hdf5FormatError("symbol-table node signature")
            for i in range(used):
                name_off, obj, cache, _, bt, hp = struct.unpack("<QQIIQQ", self._at(snod + 8 + 40 * i, 40))
                name = self._heap_name(heap, name_off)
                self._walk_group(obj, f"{prefix}/{name}", (bt, hp) if cache == 1 else None, depth + 1)

    def _group_leaves(self, node: int) -> list[int]:
        sig, ntype, level, used, _, _ = struct.unpack("<4sBBHQQ", self._at(node, 24))
        if sig != b"TREE" or ntype != 0:
            raise Hdf5FormatError("group B-tree node")
        body = self._at(node + 24, used * 16 + 8)
        children = [struct.unpack_from("<Q", body, 8 + 16 * i)[0] for i in range(used)]
        if level == 0:
            return children
        return [leaf for c in children for leaf in self._group_leaves(c)]

    # -- datasets ------------------------------------------------------------------------------------------------
    def _dataset(self, msgs) -> DatasetInfo:
        d = DatasetInfo()
        for mtype, data in msgs:
            if mtype == MSG_DATASPACE:
                version, rank, flags = data[0], data[1], data[2]
                if version == 1:
                    off = 8
                elif version == 2:
                    off = 4
                    if data[3] == 2:  # null dataspace
                        rank = 0
                else:
                    raise Hdf5FormatError(f"dataspace version {version}")
                d.shape = tuple(struct.unpack_from("<Q", data, off + 8 * i)[0] for i in range(rank))
            elif mtype == MSG_DATATYPE:
                d.dtype = _decode_datatype(data)
            elif mtype == MSG_FILTERS:
                d.filters = self._filters(data)
            elif mtype == MSG_LAYOUT:
                version = data[0]
                if version == 3:
                    cls = data[1]
                    if cls == 0:
                        n = struct.unpack_from("<H", data, 2)[0]
                        d.layout, d.compact = "compact", data[4:4 + n]
                    elif cls == 1:
                        d.layout = "contiguous"
                        d.address, d.size = struct.unpack_from("<QQ", data, 2)
                    elif cls == 2:
                        d.layout = "chunked"
                        nd = data[2]
                        d.btree = struct.unpack_from("<Q", data, 3)[0]
                        d.chunks = tuple(struct.unpack_from("<I", data, 11 + 4 * i)[0] for i in range(nd - 1))
                    else:
                        raise Hdf5FormatError(f"layout class {cls}")
                elif version in (1, 2):
                    nd, cls = data[1], data[2]
                    off = 8
                    if cls != 0:
                        d.address = struct.unpack_from("<Q", data, off)[0]
                        off += 8
                    dims = [struct.unpack_from("<I", data, off + 4 * i)[0] for i in range(nd)]
                    off += 4 * nd
                    if cls == 2:
                        d.layout, d.btree, d.chunks = "chunked", d.address, tuple(dims[:-1])
                    elif cls == 1:
                        d.layout = "contiguous"
                    else:
                        n = struct.unpack_from("<I", data, off)[0]
                        d.layout, d.compact = "compact", data[off + 4:off + 4 + n]
                else:
                    raise Hdf5FormatError(f"layout message version {version} (virtual / v4 indexes need h5py)")
        if d.shape is None or d.dtype is None or d.layout is None:
            raise Hdf5FormatError("dataset header without dataspace / datatype / layout")
        return d

    @staticmethod
    def _filters(data: bytes) -> list[tuple[int, tuple[int, ...]]]:
        version, count = data[0], data[1]
        pos = 8 if version == 1 else 2
        out = []
        for _ in range(count):
            fid = struct.unpack_from("<H", data, pos)[0]
            if version == 1 or fid >= 256:
                name_len = struct.unpack_from("<H", data, pos + 2)[0]
                pos += 4
            else:
                name_len = 0
                pos += 2
            _, nvals = struct.unpack_from("<HH", data, pos)
            pos += 4 + (_pad8(name_len) if version == 1 else name_len)
            vals = struct.unpack_from(f"<{nvals}I", data, pos)
            pos += 4 * nvals + (4 if version == 1 and nvals % 2 else 0)
            out.append((fid, tuple(vals)))
        return out

    def _chunk_entries(self, node: int, rank1: int):
        sig, ntype, level, used, _, _ = struct.unpack("<4sBBHQQ", self._at(node, 24))
        if sig != b"TREE" or ntype != 1:
            raise Hdf5FormatError("chunk B-tree node")
        key_size = 8 + 8 * rank1
        body = self._at(node + 24, used * (key_size + 8) + key_size)
        for i in range(used):
            pos = i * (key_size + 8)
            nbytes, mask = struct.unpack_from("<II", body, pos)
            origin = struct.unpack_from(f"<{rank1}Q", body, pos + 8)
            child = struct.unpack_from("<Q", body, pos + key_size)[0]
            if level == 0:
                yield origin[:-1], nbytes, mask, child
            else:
                yield from self._chunk_entries(child, rank1)

    def read(self, key: str, threads: int = 8) -> np.ndarray:
        d = self.info(key)
        if d.error:
            raise Hdf5FormatError(f"{key}: {d.error}")
        count = int(np.prod(d.shape, dtype=np.int64)) if d.shape else 1
        native = d.dtype.newbyteorder("=") if d.dtype.byteorder == ">" else d.dtype
        if d.layout == "compact":
            return np.frombuffer(d.compact, d.dtype, count).reshape(d.shape).astype(native)
        if d.layout == "contiguous":
            if d.address == UNDEF or count == 0:
                return np.zeros(d.shape, native)
            out = np.empty(d.shape, d.dtype)
            self.fh.seek(d.address + self.base)
            if self.fh.readinto(memoryview(out.reshape(-1)).cast("B")) != out.nbytes:
                raise Hdf5FormatError(f"{key}: truncated contiguous dataset")
            return out.astype(native, copy=False)
        out = np.zeros(d.shape, native)  # unallocated chunks read as the (default) fill value
        if d.btree == UNDEF or count == 0:
            return out
        entries = list(self._chunk_entries(d.btree, len(d.chunks) + 1))
        blobs = [(origin, mask, self._at(addr, nbytes)) for origin, nbytes, mask, addr in entries]

        def unpack(item):
            origin, mask, blob = item
            blob = undo_filters(blob, d.filters, mask, d.dtype.itemsize)
            block = np.frombuffer(blob, d.dtype, int(np.prod(d.chunks))).reshape(d.chunks)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(origin, d.chunks, d.shape))
            out[sel] = block[tuple(slice(0, s.stop - s.start) for s in sel)]

        if threads > 1 and len(blobs) > 1:
            with ThreadPoolExecutor(threads) as pool:
                list(pool.map(unpack, blobs))
        else:
            for item in blobs:
                unpack(item)
        return out


def read_file(path, keys: list[str] | None = None) -> dict[str, np.ndarray]:
    with File(path) as fh:
        return {k: fh.read(k) for k in fh.keys() if keys is None or k in keys}


def list_keys(path) -> list[str]:
    with File(path) as fh:
        return fh.keys()


if __name__ == "__main__":  # python -m cryovit_b200.host.hdf5_classic FILE...: what `h5ls -r -v` would say, without libhdf5
    import sys

    for arg in sys.argv[1:]:
        with File(arg) as fh_:
            print(f"{arg}: superblock at {fh_.base}, end of file address {fh_.eof}")
            for key_ in fh_.keys():
                d_ = fh_.info(key_)
                if d_.error:
                    print(f"  /{key_}  (not readable here: {d_.error})")
                    continue
                how = d_.layout + (f" {d_.chunks}" if d_.chunks else "") + ("".join(
                    f" + {'deflate' if f == FILTER_DEFLATE else 'shuffle' if f == FILTER_SHUFFLE else f'filter {f}'}{list(v)}"
                    for f, v in d_.filters))
                print(f"  /{key_}  {d_.dtype}  {d_.shape}  {how}")
