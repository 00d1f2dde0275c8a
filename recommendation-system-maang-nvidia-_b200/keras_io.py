"""Reading the encoder weights the reference saves: `encoder.keras` (src/trainer.py:230), a Keras-v3 zip archive
whose `model.weights.h5` member is an HDF5 file, without h5py (not available in this image).

Two layers:

* `read_hdf5(data) -> {path: ndarray}` -- a minimal HDF5 reader covering what h5py writes with default settings
  (superblock v0/v1, version-1 object headers, symbol-table groups = B-tree v1 + local heap, contiguous / compact
  datasets, fixed-point and IEEE float types) plus the cheap neighbours: superblock v2/v3, version-2 object headers
  with compact link messages, chunked datasets through a B-tree v1 index with no filter or deflate(+shuffle).
  Dense (fractal-heap) groups and version-4 chunk indexes raise NotImplementedError.  `write_hdf5` produces the
  h5py-default subset and exists so that the reader has something to be tested against here.
  PARITY: both were written from the HDF5 file-format specification.  The reader's generic layer (user block, superblock
  v0, version-1 object headers, symbol-table groups, contiguous IEEE datasets) is checked against ONE file the genuine
  HDF5 library wrote (a MATLAB "-v7.3" file from scipy's test data, tests/test_keras_io.py); no file produced by h5py /
  Keras was available, so the Keras weight-path layout below stays UNPINNED.

* `load_keras_encoder(path) -> weights` -- maps the Keras-v3 weight paths of the reference's MultiTowerModel
  (src/models.py:59-77: attributes user_embedding / item_embedding = Embedding, user_tower / item_tower =
  Sequential of Dense) to the weight dict `models.load_weights` takes:
      <side>_embedding/vars/0                      -> "<side>_embedding"   [V + 1, E]
      <side>_tower/layers/dense[_i]/vars/0, /1     -> "<side>_tower"[i] = (kernel (in, out), bias)
  Components in front of these (e.g. a `layers/` or `encoder/` prefix) are ignored.
"""

import io
import struct
import zipfile
import zlib
from typing import Dict, List, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


# ---------------------------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, data: bytes):
        self.d = memoryview(data)
        base = 0
        while bytes(self.d[base:base + 8]) != _SIG:          # the superblock may sit at 0, 512, 1024, ...
            base = 512 if base == 0 else base * 2
            if base + 8 > len(self.d):
                raise ValueError("not an HDF5 file (no superblock signature)")
        self.base = base
        ver = self.d[base + 8]
        if ver in (0, 1):
            self.O, self.L = self.d[base + 13], self.d[base + 14]
            p = base + 24 + (4 if ver == 1 else 0)
            p += 4 * self.O                                   # base, free-space, end-of-file, driver-info addresses
            self.root_header = self._uint(p + self.O, self.O)  # root symbol-table entry: name offset, header address
        elif ver in (2, 3):
            self.O, self.L = self.d[base + 9], self.d[base + 10]
            self.root_header = self._uint(base + 12 + 3 * self.O, self.O)
        else:
            raise NotImplementedError(f"HDF5 superblock version {ver}")
        if self.O != 8 or self.L != 8:
            raise NotImplementedError("only 8-byte offsets / lengths are supported")

    def _uint(self, p: int, n: int) -> int:
        return int.from_bytes(self.d[p:p + n], "little")

    def _addr(self, a: int) -> int:
        return self.base + a

    # ---- object headers -> list of (type, payload offset, size) ----
    def messages(self, addr: int) -> List[Tuple[int, int, int]]:
        p = self._addr(addr)
        if bytes(self.d[p:p + 4]) == b"OHDR":
            return self._messages_v2(p)
        if self.d[p] != 1:
            raise NotImplementedError(f"object header version {self.d[p]}")
        n_msgs = self._uint(p + 2, 2)
        size = self._uint(p + 8, 4)
        out, blocks = [], [(p + 16, size)]
        while blocks and len(out) < n_msgs:
            q, left = blocks.pop(0)
            end = q + left
            while q + 8 <= end and len(out) < n_msgs:
                mtype, msize = self._uint(q, 2), self._uint(q + 2, 2)
                body = q + 8
                if mtype == 0x10:                                 # continuation: offset, length
                    blocks.append((self._addr(self._uint(body, 8)), self._uint(body + 8, 8)))
                out.append((mtype, body, msize))
                q = body + msize
        return out

    def _messages_v2(self, p: int) -> List[Tuple[int, int, int]]:
        flags = self.d[p + 5]
        q = p + 6
        if flags & 0x20:
            q += 16
        if flags & 0x10:
            q += 4
        nsz = 1 << (flags & 3)
        chunk0 = self._uint(q, nsz)
        q += nsz
        out, blocks = [], [(q, chunk0)]
        while blocks:
            q, left = blocks.pop(0)
            end = q + left
            while q + 4 <= end:
                mtype, msize, _mflags = self.d[q], self._uint(q + 1, 2), self.d[q + 3]
                body = q + 4 + (2 if flags & 0x04 else 0)
                if mtype == 0 and msize == 0:
                    break                                         # gap / padding at the end of the chunk
                if mtype == 0x10:
                    a, ln = self._addr(self._uint(body, 8)), self._uint(body + 8, 8)
                    if bytes(self.d[a:a + 4]) != b"OCHK":
                        raise ValueError("bad object header continuation block")
                    blocks.append((a + 4, ln - 8))               # minus signature and checksum
                out.append((mtype, body, msize))
                q = body + msize
        return out

    # ---- groups ----
    def children(self, msgs) -> Dict[str, int]:
        out: Dict[str, int] = {}
        for mtype, body, _ in msgs:
            if mtype == 0x11:                                     # symbol table: B-tree v1 + local heap
                heap = self._addr(self._uint(body + 8, 8))
                if bytes(self.d[heap:heap + 4]) != b"HEAP":
                    raise ValueError("bad local heap")
                heap_data = self._addr(self._uint(heap + 8 + 16, 8))
                self._walk_group_btree(self._addr(self._uint(body, 8)), heap_data, out)
            elif mtype == 0x06:                                   # link message (compact new-style group)
                lflags = self.d[body + 1]
                q = body + 2
                ltype = 0
                if lflags & 0x08:
                    ltype = self.d[q]
                    q += 1
                if lflags & 0x04:
                    q += 8
                if lflags & 0x10:
                    q += 1
                nsz = 1 << (lflags & 3)
                nlen = self._uint(q, nsz)
                q += nsz
                name = bytes(self.d[q:q + nlen]).decode("utf-8")
                q += nlen
                if ltype == 0:
                    out[name] = self._uint(q, 8)
            elif mtype == 0x02:                                   # link info: a fractal heap means dense storage
                lflags = self.d[body + 1]
                q = body + 2 + (8 if lflags & 1 else 0)
                if self._uint(q, 8) != _UNDEF:
                    raise NotImplementedError("dense (fractal-heap) group storage is not supported")
        return out

    def _walk_group_btree(self, p: int, heap_data: int, out: Dict[str, int]) -> None:
        if bytes(self.d[p:p + 4]) != b"TREE" or self.d[p + 4] != 0:
            raise ValueError("bad group B-tree node")
        level, used = self.d[p + 5], self._uint(p + 6, 2)
        q = p + 8 + 16                                            # past the sibling addresses
        for i in range(used):
            child = self._addr(self._uint(q + 8 + i * 16, 8))     # key_i (8) child_i (8) key_{i+1} ...
            if level > 0:
                self._walk_group_btree(child, heap_data, out)
                continue
            if bytes(self.d[child:child + 4]) != b"SNOD":
                raise ValueError("bad symbol table node")
            for s in range(self._uint(child + 6, 2)):
                e = child + 8 + s * 40
                name_off, header = self._uint(e, 8), self._uint(e + 8, 8)
                end = heap_data + name_off
                while self.d[end] != 0:
                    end += 1
                out[bytes(self.d[heap_data + name_off:end]).decode("utf-8")] = header

    # ---- datasets ----
    def dataset(self, msgs) -> np.ndarray:
        shape, dtype, layout, filters = None, None, None, []
        for mtype, body, size in msgs:
            if mtype == 0x01:
                ver, rank = self.d[body], self.d[body + 1]
                q = body + (8 if ver == 1 else 4)
                shape = tuple(self._uint(q + 8 * i, 8) for i in range(rank))
            elif mtype == 0x03:
                dtype = self._dtype(body)
            elif mtype == 0x08:
                layout = (body, size)
            elif mtype == 0x0B:
                filters = self._filters(body)
        if shape is None or dtype is None or layout is None:
            raise ValueError("object is not a dataset")
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        body, _ = layout
        ver = self.d[body]
        if ver == 3:
            cls = self.d[body + 1]
            if cls == 0:
                sz = self._uint(body + 2, 2)
                raw = bytes(self.d[body + 4:body + 4 + sz])
            elif cls == 1:
                a, sz = self._uint(body + 2, 8), self._uint(body + 10, 8)
                raw = b"" if a == _UNDEF else bytes(self.d[self._addr(a):self._addr(a) + sz])
            elif cls == 2:
                ndim = self.d[body + 2]
                btree = self._uint(body + 3, 8)
                cdims = tuple(self._uint(body + 11 + 4 * i, 4) for i in range(ndim - 1))
                return self._chunked(btree, shape, cdims, dtype, filters)
            else:
                raise NotImplementedError(f"data layout class {cls}")
        elif ver in (1, 2):
            ndim, cls = self.d[body + 1], self.d[body + 2]
            q = body + 8
            a = None
            if cls != 0:
                a = self._uint(q, 8)
                q += 8
            dims = [self._uint(q + 4 * i, 4) for i in range(ndim)]
            q += 4 * ndim
            if cls == 1:
                raw = bytes(self.d[self._addr(a):self._addr(a) + n * dtype.itemsize])
            elif cls == 0:
                sz = self._uint(q, 4)
                raw = bytes(self.d[q + 4:q + 4 + sz])
            else:
                return self._chunked(a, shape, tuple(dims[:-1]), dtype, filters)
        else:
            raise NotImplementedError(f"data layout message version {ver}")
        if len(raw) < n * dtype.itemsize:
            return np.zeros(shape, dtype.newbyteorder("="))        # never written: the fill value (zeros)
        return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).astype(dtype.newbyteorder("="))

    def _dtype(self, body: int) -> np.dtype:
        cls, bits0, size = self.d[body] & 0x0F, self.d[body + 1], self._uint(body + 4, 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 1:
            if size not in (2, 4, 8):
                raise NotImplementedError(f"{size}-byte floating-point type")
            return np.dtype(f"{order}f{size}")
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}")
        raise NotImplementedError(f"HDF5 datatype class {cls}")

    def _filters(self, body: int) -> List[int]:
        ver, n = self.d[body], self.d[body + 1]
        q = body + (8 if ver == 1 else 2)
        ids = []
        for _ in range(n):
            fid = self._uint(q, 2)
            if ver == 1 or fid >= 256:
                nlen = self._uint(q + 2, 2)
                q += 4
            else:
                nlen = 0
                q += 2
            ncd = self._uint(q + 2, 2)
            q += 4 + nlen + 4 * ncd
            if ver == 1 and ncd % 2:
                q += 4
            ids.append(fid)
        return ids

    def _chunked(self, btree: int, shape, cdims, dtype: np.dtype, filters: List[int]) -> np.ndarray:
        for f in filters:
            if f not in (1, 2):
                raise NotImplementedError(f"HDF5 filter {f} (only deflate and shuffle are supported)")
        out = np.zeros(shape, dtype.newbyteorder("="))
        if btree == _UNDEF:
            return out
        nd = len(shape)
        csize = int(np.prod(cdims)) * dtype.itemsize

        def walk(p):
            if bytes(self.d[p:p + 4]) != b"TREE" or self.d[p + 4] != 1:
                raise ValueError("bad chunk B-tree node")
            level, used = self.d[p + 5], self._uint(p + 6, 2)
            q = p + 8 + 16
            ksz = 8 + 8 * (nd + 1)
            for i in range(used):
                k = q + i * (ksz + 8)
                nbytes, mask = self._uint(k, 4), self._uint(k + 4, 4)
                offs = [self._uint(k + 8 + 8 * j, 8) for j in range(nd)]
                child = self._addr(self._uint(k + ksz, 8))
                if level > 0:
                    walk(child)
                    continue
                raw = bytes(self.d[child:child + nbytes])
                if 1 in filters and not (mask >> filters.index(1)) & 1:
                    raw = zlib.decompress(raw)
                if 2 in filters and not (mask >> filters.index(2)) & 1:
                    raw = np.frombuffer(raw, np.uint8).reshape(dtype.itemsize, -1).T.tobytes()
                chunk = np.frombuffer(raw[:csize], dtype=dtype).reshape(cdims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]

        walk(self._addr(btree))
        return out


def read_hdf5(data) -> Dict[str, np.ndarray]:
    """All datasets of an HDF5 file (bytes, or a path) as {"group/sub/name": array}."""
    if not isinstance(data, (bytes, bytearray, memoryview)):
        with open(data, "rb") as f:
            data = f.read()
    r = _Reader(bytes(data))
    out: Dict[str, np.ndarray] = {}
    seen = set()

    def visit(addr: int, path: str) -> None:
        if addr in seen:
            return
        seen.add(addr)
        msgs = r.messages(addr)
        types = {m[0] for m in msgs}
        if 0x08 in types:
            out[path] = r.dataset(msgs)
            return
        for name, child in sorted(r.children(msgs).items()):
            visit(child, f"{path}/{name}" if path else name)

    visit(r.root_header, "")
    return out


# ---------------------------------------------------------------------------------------------------------------
# writer (the h5py-default subset: superblock v0, v1 object headers, symbol-table groups, contiguous datasets)
# ---------------------------------------------------------------------------------------------------------------
def write_hdf5(path, datasets: Dict[str, np.ndarray], chunks: Dict[str, Tuple[int, ...]] = None, deflate: bool = False) -> None:
    """Write {"group/sub/name": array} as an HDF5 file laid out the way h5py lays out a default file.  Datasets named in
    `chunks` ({path: chunk shape}) are stored chunked behind a one-node B-tree (<= 64 chunks), shuffled + deflated when
    `deflate` is set -- what h5py produces for create_dataset(..., chunks=..., compression="gzip", shuffle=True)."""
    tree: Dict = {}
    chunk_of = {}
    for full, arr in datasets.items():
        node = tree
        parts = [p for p in full.split("/") if p]
        for p in parts[:-1]:
            node = node.setdefault(p, {})
            if not isinstance(node, dict):
                raise ValueError(f"{full}: a dataset is in the way")
        node[parts[-1]] = np.require(arr, requirements="C")
        if chunks and full in chunks:
            chunk_of[id(node[parts[-1]])] = tuple(int(c) for c in chunks[full])
    buf = bytearray(96)                                           # superblock v0 with 8-byte offsets: 24 + 32 + 40

    def alloc(n: int) -> int:
        while len(buf) % 8:
            buf.append(0)
        a = len(buf)
        buf.extend(b"\0" * n)
        return a

    def message(mtype: int, body: bytes) -> bytes:
        body += b"\0" * (-len(body) % 8)
        return struct.pack("<HHB3x", mtype, len(body), 0) + body

    def header(msgs: List[bytes]) -> int:
        blob = b"".join(msgs)
        a = alloc(16 + len(blob))
        buf[a:a + 16] = struct.pack("<BxHII4x", 1, len(msgs), 1, len(blob))
        buf[a + 16:a + 16 + len(blob)] = blob
        return a

    def write_dataset(arr: np.ndarray) -> int:
        raw = arr.astype(arr.dtype.newbyteorder("<")).tobytes()
        data = alloc(max(len(raw), 1))
        buf[data:data + len(raw)] = raw
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape)
        if arr.dtype.kind == "f":
            size = arr.dtype.itemsize
            exp, man = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[size]
            dt = struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size)
            dt += struct.pack("<HHBBBBI", 0, size * 8, man, exp, 0, man, (1 << (exp - 1)) - 1)
        elif arr.dtype.kind in "iu":
            size = arr.dtype.itemsize
            dt = struct.pack("<BBBBI", 0x10, 0x08 if arr.dtype.kind == "i" else 0, 0, 0, size) + struct.pack("<HH", 0, size * 8)
        else:
            raise NotImplementedError(f"dtype {arr.dtype}")
        cdims = chunk_of.get(id(arr))
        if cdims is None or arr.ndim == 0 or arr.size == 0:
            layout = struct.pack("<BBQQ", 3, 1, data, len(raw))
            return header([message(0x01, space), message(0x03, dt), message(0x08, layout)])
        # chunked: every chunk is a full cdims block (edge chunks zero-padded), indexed by one leaf B-tree node
        little = arr.astype(arr.dtype.newbyteorder("<"))
        grid = [range(0, s, c) for s, c in zip(arr.shape, cdims)]
        entries = []
        for offs in np.ndindex(*[len(g) for g in grid]):
            start = [g[i] for g, i in zip(grid, offs)]
            block = np.zeros(cdims, little.dtype)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(start, cdims, arr.shape))
            block[tuple(slice(0, x.stop - x.start) for x in sl)] = little[sl]
            payload = block.tobytes()
            if deflate:
                payload = np.frombuffer(payload, np.uint8).reshape(-1, arr.dtype.itemsize).T.tobytes()   # shuffle
                payload = zlib.compress(payload, 4)
            a = alloc(len(payload))
            buf[a:a + len(payload)] = payload
            entries.append((start, a, len(payload)))
        if len(entries) > 64:
            raise NotImplementedError("more than 64 chunks in one dataset")
        nd = arr.ndim
        ksz = 8 + 8 * (nd + 1)
        bt = alloc(24 + (ksz + 8) * 64 + ksz)
        buf[bt:bt + 24] = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), _UNDEF, _UNDEF)
        q = bt + 24
        for start, a, nbytes in entries:
            buf[q:q + ksz] = struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in start) + struct.pack("<Q", 0)
            buf[q + ksz:q + ksz + 8] = struct.pack("<Q", a)
            q += ksz + 8
        buf[q:q + ksz] = struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape) + struct.pack("<Q", 0)
        layout = struct.pack("<BBBQ", 3, 2, nd + 1, bt) + b"".join(struct.pack("<I", c) for c in cdims) + struct.pack("<I", arr.dtype.itemsize)
        msgs = [message(0x01, space), message(0x03, dt), message(0x08, layout)]
        if deflate:   # filter pipeline v1: shuffle (id 2, element size), then deflate (id 1, level)
            pipe = struct.pack("<BB6x", 1, 2)
            pipe += struct.pack("<HHHHI4x", 2, 0, 0, 1, arr.dtype.itemsize)
            pipe += struct.pack("<HHHHI4x", 1, 0, 0, 1, 4)
            msgs.append(message(0x0B, pipe))
        return header(msgs)

    def write_group(node: Dict) -> int:
        entries = []
        for name in sorted(node):
            child = node[name]
            entries.append((name, write_group(child) if isinstance(child, dict) else write_dataset(child)))
        heap_bytes = bytearray(b"\0" * 8)                          # offset 0: the empty name
        offsets = []
        for name, _ in entries:
            offsets.append(len(heap_bytes))
            heap_bytes += name.encode("utf-8") + b"\0"
            heap_bytes += b"\0" * (-len(heap_bytes) % 8)
        heap_data = alloc(len(heap_bytes))
        buf[heap_data:heap_data + len(heap_bytes)] = heap_bytes
        heap = alloc(32)
        buf[heap:heap + 32] = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_bytes), _UNDEF, heap_data)
        snods = []                                                # at most 8 symbols per node (2 * leaf K, K = 4)
        for i in range(0, max(len(entries), 1), 8):
            part = list(zip(entries[i:i + 8], offsets[i:i + 8]))
            a = alloc(8 + 40 * 8)
            buf[a:a + 8] = b"SNOD" + struct.pack("<BxH", 1, len(part))
            for s, ((_, hdr), off) in enumerate(part):
                buf[a + 8 + 40 * s:a + 48 + 40 * s] = struct.pack("<QQII16x", off, hdr, 0, 0)
            snods.append((a, part[-1][1] if part else 0))
        if len(snods) > 32:
            raise NotImplementedError("more than 256 links in one group")
        bt = alloc(24 + 8 + 16 * 32)
        buf[bt:bt + 24] = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), _UNDEF, _UNDEF)
        q = bt + 24
        buf[q:q + 8] = struct.pack("<Q", 0)
        for i, (a, last_off) in enumerate(snods):
            buf[q + 8 + 16 * i:q + 24 + 16 * i] = struct.pack("<QQ", a, last_off)
        return header([message(0x11, struct.pack("<QQ", bt, heap))])

    root = write_group(tree)
    eof = len(buf)
    sb = _SIG + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
    sb += struct.pack("<QQII16x", 0, root, 0, 0)
    buf[0:len(sb)] = sb
    with open(path, "wb") as f:
        f.write(bytes(buf))


# ---------------------------------------------------------------------------------------------------------------
# Keras-v3 archive -> encoder weight dict
# ---------------------------------------------------------------------------------------------------------------
def _dense_index(name: str) -> int:
    tail = name.rsplit("_", 1)
    return int(tail[1]) if len(tail) == 2 and tail[1].isdigit() else 0


def encoder_weights_from_datasets(ds: Dict[str, np.ndarray]) -> Dict:
    """{h5 path: array} of a saved MultiTowerModel -> the dict models.load_weights / service.load take."""
    w: Dict = {}
    for side in ("user", "item"):
        emb = [v for k, v in ds.items() if f"{side}_embedding" in k.split("/") and k.endswith("vars/0")]
        if len(emb) != 1:
            raise KeyError(f"expected one '{side}_embedding/vars/0' dataset, found {len(emb)}")
        w[f"{side}_embedding"] = np.ascontiguousarray(emb[0], dtype=np.float32)
        layers: Dict[str, Dict[str, np.ndarray]] = {}
        for k, v in ds.items():
            parts = k.split("/")
            if f"{side}_tower" not in parts:
                continue
            rest = parts[parts.index(f"{side}_tower") + 1:]
            if len(rest) == 4 and rest[0] == "layers" and rest[2] == "vars":
                layers.setdefault(rest[1], {})[rest[3]] = v
        names = sorted(layers, key=lambda n: (_dense_index(n), n))
        if not names:
            raise KeyError(f"no '{side}_tower/layers/<dense>/vars/<i>' datasets found")
        tower = []
        for n in names:
            if set(layers[n]) != {"0", "1"}:
                raise KeyError(f"{side}_tower/layers/{n}: expected vars/0 (kernel) and vars/1 (bias)")
            kernel, bias = layers[n]["0"], layers[n]["1"]
            if kernel.ndim != 2 or bias.shape != (kernel.shape[1],):
                raise ValueError(f"{side}_tower/layers/{n}: kernel {kernel.shape} / bias {bias.shape} do not form a Dense layer")
            tower.append((np.ascontiguousarray(kernel, dtype=np.float32), np.ascontiguousarray(bias, dtype=np.float32)))
        w[f"{side}_tower"] = tower
    return w


def load_keras_encoder(path) -> Dict:
    """encoder.keras (zip with model.weights.h5), or a bare .h5 weights file -> encoder weight dict."""
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:2] == b"PK":
        with zipfile.ZipFile(io.BytesIO(blob)) as z:
            names = [n for n in z.namelist() if n.endswith(".weights.h5")]
            if not names:
                raise KeyError(f"{path}: no *.weights.h5 member in the archive")
            blob = z.read(names[0])
    return encoder_weights_from_datasets(read_hdf5(blob))


def save_keras_encoder(path, weights: Dict) -> None:
    """The inverse, for tests and for exporting synthetic models in the reference's artifact layout."""
    import os
    import tempfile
    ds = {}
    for side in ("user", "item"):
        ds[f"{side}_embedding/vars/0"] = np.asarray(weights[f"{side}_embedding"], dtype=np.float32)
        for i, (kernel, bias) in enumerate(weights[f"{side}_tower"]):
            name = "dense" if i == 0 else f"dense_{i}"
            ds[f"{side}_tower/layers/{name}/vars/0"] = np.asarray(kernel, dtype=np.float32)
            ds[f"{side}_tower/layers/{name}/vars/1"] = np.asarray(bias, dtype=np.float32)
    with tempfile.TemporaryDirectory() as tmp:
        h5 = os.path.join(tmp, "model.weights.h5")
        write_hdf5(h5, ds)
        with zipfile.ZipFile(path, "w") as z:
            z.write(h5, "model.weights.h5")
            z.writestr("metadata.json", '{"keras_version": "3", "writer": "recsys_b200.keras_io"}')
            z.writestr("config.json", '{"class_name": "MultiTowerModel"}')
