#!/usr/bin/env python
"""Generate tests/golden/symmetricblockexamples_{cuboid,sphere}.npz from the one fixture the
reference ships: /root/reference/test/assets/symmetricblockexamples.jld2
(loaded by /root/reference/test/test_symmetricblockmatrix.jl:9-16 as
 blockdict[example] = (diagonalblocks, selfindices, offblocks, testindices, trialindices)).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

No Julia, no h5py: a minimal reader for the HDF5 subset JLD2 0.2.0 writes (superblock v2 at byte
512, version-2 object headers, contiguous/compact layouts, object-reference arrays).
The .npz stores, per example, the matrices flattened column-major (exactly the bytes Julia holds)
plus shapes and the 1-based Int64 index vectors, unmodified.
"""
import struct
import sys
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/test/assets/symmetricblockexamples.jld2")
OUT = Path(__file__).resolve().parent
BASE = 512  # superblock base address; every in-file address is relative to it

# absolute offsets of the five container object headers per example, in tuple order
ANCHORS = {
    "cuboid": (7449, 358835, 379704, 1891616, 1892488),
    "sphere": (1955387, 2231408, 2252192, 3667352, 3668312),
}


def parse_ohdr(buf, pos):
    """Return the list of (type, payload bytes) messages of the v2 object header at `pos`."""
    assert buf[pos:pos + 4] == b"OHDR", (pos, buf[pos:pos + 4])
    flags = buf[pos + 5]
    p = pos + 6
    if flags & 0x20:
        p += 16
    if flags & 0x10:
        p += 4
    szw = 1 << (flags & 3)
    chunk0 = int.from_bytes(buf[p:p + szw], "little")
    p += szw
    msgs = []
    todo = [(p, p + chunk0)]
    while todo:
        q, end = todo.pop(0)
        while q + 4 <= end:
            mtype = buf[q]
            msize = struct.unpack_from("<H", buf, q + 1)[0]
            q += 4
            if flags & 0x04:
                q += 2
            payload = buf[q:q + msize]
            q += msize
            if mtype == 0x10:  # continuation → OCHK block, 4-byte signature, 4-byte checksum tail
                addr, length = struct.unpack_from("<QQ", payload, 0)
                a = BASE + addr
                assert buf[a:a + 4] == b"OCHK"
                todo.append((a + 4, a + length - 4))
            elif mtype != 0:
                msgs.append((mtype, payload))
    return msgs


def read_dataset(buf, pos):
    """Read the dataset whose object header is at absolute offset `pos`.
    Returns ('ref', [abs offsets]) | ('i64', array) | ('c128', 2-D column-major array)."""
    msgs = parse_ohdr(buf, pos)
    dims = None
    dclass = None
    dsize = None
    data = None
    for mtype, pl in msgs:
        if mtype == 0x01:  # dataspace v2
            rank = pl[1]
            dims = struct.unpack_from("<%dQ" % rank, pl, 4)
        elif mtype == 0x03:  # datatype (possibly shared → treat by rank below)
            dclass = pl[0] & 0x0F
            dsize = struct.unpack_from("<I", pl, 4)[0] if len(pl) >= 8 else None
        elif mtype == 0x08:  # layout v3/4
            lclass = pl[1]
            if lclass == 1:
                addr, size = struct.unpack_from("<QQ", pl, 2)
                data = buf[BASE + addr:BASE + addr + size]
            elif lclass == 0:
                size = struct.unpack_from("<H", pl, 2)[0]
                data = pl[4:4 + size]
            else:
                raise ValueError("unsupported layout class %d" % lclass)
    assert dims is not None and data is not None, pos
    if len(dims) == 2:
        # HDF5 dims are reversed w.r.t. Julia; data is Julia column-major
        ncols, nrows = dims
        a = np.frombuffer(data, dtype=np.complex128, count=nrows * ncols)
        return "c128", a.reshape((ncols, nrows)).T  # → (nrows, ncols) view, Fortran order
    n = dims[0]
    if dclass == 7:
        refs = np.frombuffer(data, dtype="<u8", count=n)
        return "ref", [BASE + int(r) for r in refs]
    if dclass == 0 and dsize == 8:
        return "i64", np.frombuffer(data, dtype="<i8", count=n).copy()
    raise ValueError("unexpected dataset at %d: dims=%s class=%s size=%s" % (pos, dims, dclass, dsize))


def load_container(buf, pos, want):
    kind, refs = read_dataset(buf, pos)
    assert kind == "ref", (pos, kind)
    out = []
    for r in refs:
        k, a = read_dataset(buf, r)
        assert k == want, (r, k, want)
        out.append(a)
    return out


def pack_matrices(mats):
    shapes = np.array([m.shape for m in mats], dtype=np.int64).reshape(-1, 2)
    flat = np.concatenate([np.asarray(m).ravel(order="F") for m in mats]) if mats else np.zeros(0, np.complex128)
    return shapes, flat


def pack_indices(vecs):
    ptr = np.zeros(len(vecs) + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(v) for v in vecs])
    pool = np.concatenate(vecs).astype(np.int64) if vecs else np.zeros(0, np.int64)
    return pool, ptr


def main():
    buf = SRC.read_bytes()
    assert buf[BASE:BASE + 8] == b"\x89HDF\r\n\x1a\n"
    expect = {"cuboid": (1344, 96, 92, 21264, 93842), "sphere": (1203, 106, 103, 16501, 87718)}
    for name, anchors in ANCHORS.items():
        diag = load_container(buf, anchors[0], "c128")
        selfidx = load_container(buf, anchors[1], "i64")
        off = load_container(buf, anchors[2], "c128")
        testidx = load_container(buf, anchors[3], "i64")
        trialidx = load_container(buf, anchors[4], "i64")
        N = max(int(v.max()) for v in selfidx)
        nd = sum(m.size for m in diag)
        no = sum(m.size for m in off)
        got = (N, len(diag), len(off), nd, no)
        assert got == expect[name], (name, got)
        for d, ix in zip(diag, selfidx):
            assert d.shape == (len(ix), len(ix))
            assert np.array_equal(d, d.T)
        for o, r, c in zip(off, testidx, trialidx):
            assert o.shape == (len(r), len(c))
        dshape, dflat = pack_matrices(diag)
        oshape, oflat = pack_matrices(off)
        dpool, dptr = pack_indices(selfidx)
        rpool, rptr = pack_indices(testidx)
        cpool, cptr = pack_indices(trialidx)
        np.savez(OUT / ("symmetricblockexamples_%s.npz" % name), n=np.int64(N),
                 diag_shapes=dshape, diag_values=dflat, diag_idx=dpool, diag_ptr=dptr,
                 off_shapes=oshape, off_values=oflat, row_idx=rpool, row_ptr=rptr,
                 col_idx=cpool, col_ptr=cptr)
        print(name, got, "first diag row:", diag[0][0, :2])
    return 0


if __name__ == "__main__":
    sys.exit(main())
