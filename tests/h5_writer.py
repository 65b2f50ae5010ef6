"""Test helper: writes a small HDF5 file in the 'earliest' on-disk format (superblock version 0, symbol-table root
group, version-1 object headers, contiguous datasets) straight from the HDF5 File Format Specification -- the
structures h5py produces for ``h5py.File(path, "w").create_dataset(name, data=array)``.  h5py itself is not in this
image; this writer and jatts_b200/_h5lite.py are two independent statements of the same specification, and the
tests make them meet in the middle.  Never imported by the product."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _msg(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _object_header(msgs) -> bytes:
    body = b"".join(msgs)
    return struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body


def _datatype(dt: np.dtype) -> bytes:
    if dt.kind == "f":
        nb = dt.itemsize
        exp_bits, man_bits, bias = {4: (8, 23, 127), 8: (11, 52, 1023)}[nb]
        head = bytes([0x11, 0x20, nb * 8 - 1, 0]) + struct.pack("<I", nb)
        return head + struct.pack("<HHBBBBI", 0, nb * 8, man_bits, exp_bits, 0, man_bits, bias)
    if dt.kind in "iu":
        head = bytes([0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0]) + struct.pack("<I", dt.itemsize)
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    raise ValueError(dt)


def write_h5(path, datasets: dict):
    """datasets: name -> numpy array (little-endian float32/float64/int)."""
    names = sorted(datasets)
    assert 0 < len(names) <= 8, "one symbol-table node (2K = 8 entries) is enough for a statistics file"
    # ---- layout plan: superblock | root header | B-tree | SNOD | heap header | heap data | per dataset: header, raw
    sb_size = 96
    root_hdr = _object_header([_msg(0x0011, struct.pack("<QQ", 0, 0))])   # patched below
    heap_data = _pad8(b"\x00")
    name_off = {}
    for n in names:
        name_off[n] = len(heap_data)
        heap_data += _pad8(n.encode() + b"\x00")
    off_root = sb_size
    off_btree = off_root + len(root_hdr)
    btree_size = 8 + 16 + 8 + 8 + 8
    off_snod = off_btree + btree_size
    snod_size = 8 + 40 * len(names)
    off_heap = off_snod + snod_size
    off_heap_data = off_heap + 32
    p = off_heap_data + len(heap_data)
    ds_hdr, ds_raw, blobs = {}, {}, {}
    for n in names:
        a = np.ascontiguousarray(datasets[n])
        a = a.astype(a.dtype.newbyteorder("<"))
        space = struct.pack("<BBB5x", 1, a.ndim, 0) + b"".join(struct.pack("<Q", d) for d in a.shape)
        hdr_len = len(_object_header([_msg(1, space), _msg(3, _datatype(a.dtype)), _msg(8, bytes(18))]))
        ds_hdr[n] = p
        p += hdr_len
        p += -p % 8
        ds_raw[n] = p
        blobs[n] = (a, space)
        p += a.nbytes
        p += -p % 8
    eof = p
    out = bytearray(eof)
    out[0:8] = b"\x89HDF\r\n\x1a\n"
    out[8:24] = struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, 4, 16, 0)
    out[24:56] = struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    out[56:96] = struct.pack("<QQII", 0, off_root, 1, 0) + struct.pack("<QQ", off_btree, off_heap)
    root_hdr = _object_header([_msg(0x0011, struct.pack("<QQ", off_btree, off_heap))])
    out[off_root:off_root + len(root_hdr)] = root_hdr
    out[off_btree:off_btree + btree_size] = (b"TREE" + struct.pack("<BBH", 0, 0, 1) + struct.pack("<QQ", UNDEF, UNDEF) +
                                             struct.pack("<QQQ", 0, off_snod, name_off[names[-1]]))
    snod = b"SNOD" + struct.pack("<BxH", 1, len(names))
    for n in names:
        snod += struct.pack("<QQII16x", name_off[n], ds_hdr[n], 0, 0)
    out[off_snod:off_snod + snod_size] = snod
    out[off_heap:off_heap + 32] = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, off_heap_data)
    out[off_heap_data:off_heap_data + len(heap_data)] = heap_data
    for n in names:
        a, space = blobs[n]
        layout = struct.pack("<BBQQ", 3, 1, ds_raw[n], a.nbytes)
        hdr = _object_header([_msg(1, space), _msg(3, _datatype(a.dtype)), _msg(8, layout)])
        out[ds_hdr[n]:ds_hdr[n] + len(hdr)] = hdr
        out[ds_raw[n]:ds_raw[n] + a.nbytes] = a.tobytes()
    with open(path, "wb") as f:
        f.write(bytes(out))
