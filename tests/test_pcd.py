"""PCD v0.7 reader (elm_pcd_read_xyz / elm_map_add_points_pcd; the node's pcl::io::loadPCDFile, pcm_matching.cpp:69-79):
ascii, binary and binary_compressed (LZF) files written by this test — PointXYZINormal-like field lists, extra fields, a
COUNT > 1 field, double-precision coordinates, NaN points — read back bit for bit; malformed files are refused."""
import struct

import numpy as np
import pytest

import elimaloc_b200 as E
from elimaloc_b200 import _capi, synth


def lzf_compress(data: bytes) -> bytes:
    """A small LZF encoder (liblzf stream format): greedy matching through a dict of 3-byte prefixes; emits literal runs
    (ctrl = n - 1 < 32) and back references (len - 2 in the top 3 bits, 7 = extended; 13-bit distance - 1)."""
    out, lit, i, n, table = bytearray(), bytearray(), 0, len(data), {}

    def flush():
        for k in range(0, len(lit), 32):
            chunk = lit[k:k + 32]
            out.append(len(chunk) - 1)
            out.extend(chunk)
        lit.clear()

    while i < n:
        key = data[i:i + 3]
        ref = table.get(key) if len(key) == 3 else None
        table[key] = i
        if ref is not None and 0 < i - ref <= 8192:
            length = 3
            while i + length < n and length < 264 and data[ref + length] == data[i + length]:
                length += 1
            flush()
            dist, l2 = i - ref - 1, length - 2
            if l2 < 7:
                out.append((l2 << 5) | (dist >> 8))
            else:
                out.append((7 << 5) | (dist >> 8))
                out.append(l2 - 7)
            out.append(dist & 0xFF)
            i += length
        else:
            lit.append(data[i])
            i += 1
    flush()
    return bytes(out)


FIELDS = [("x", "F", 4, 1), ("y", "F", 4, 1), ("z", "F", 4, 1), ("intensity", "F", 4, 1), ("normal", "F", 4, 3), ("ring", "U", 2, 1)]


def write_pcd(path, xyz, mode, fields=FIELDS, coord_type=("F", 4)):
    n = len(xyz)
    fields = [(nm, coord_type[0], coord_type[1], c) if nm in "xyz" else (nm, t, s, c) for nm, t, s, c in fields]
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n"
           f"FIELDS {' '.join(f[0] for f in fields)}\nSIZE {' '.join(str(f[2]) for f in fields)}\n"
           f"TYPE {' '.join(f[1] for f in fields)}\nCOUNT {' '.join(str(f[3]) for f in fields)}\n"
           f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {mode}\n").encode()
    fmt = {("F", 4): "f", ("F", 8): "d", ("U", 2): "H", ("I", 4): "i"}
    cols = []  # per field: list of per-point byte strings
    for nm, t, s, c in fields:
        if nm in ("x", "y", "z"):
            v = xyz[:, "xyz".index(nm)]
            cols.append([struct.pack("<" + fmt[(t, s)], float(a)) for a in v])
        else:
            cols.append([struct.pack("<" + fmt[(t, s)] * c, *([7] * c if t != "F" else [0.5] * c)) for _ in range(n)])
    if mode == "ascii":
        lines = []
        for i in range(n):
            vals = []
            for (nm, t, s, c), col in zip(fields, cols):
                vals += [repr(float(xyz[i, "xyz".index(nm)])) if nm in "xyz" else ("7" if t != "F" else "0.5")] * (1 if nm in "xyz" else c)
            lines.append(" ".join(vals))
        body = ("\n".join(lines) + "\n").encode()
    elif mode == "binary":
        body = b"".join(b"".join(col[i] for col in cols) for i in range(n))
    else:
        raw = b"".join(b"".join(col) for col in cols)  # field-major
        comp = lzf_compress(raw)
        body = struct.pack("<II", len(comp), len(raw)) + comp
    with open(path, "wb") as f:
        f.write(hdr + body)


@pytest.mark.parametrize("mode", ["ascii", "binary", "binary_compressed"])
@pytest.mark.parametrize("coord_type", [("F", 4), ("F", 8)])
def test_pcd_round_trip(tmp_path, mode, coord_type):
    rng = np.random.default_rng(3)
    xyz = ((rng.random((1500, 3)) * 2 - 1) * 300.0).astype(np.float32)
    xyz[::50] = np.round(xyz[::50])          # long runs of repeated bytes: real LZF back references
    xyz[7] = [np.nan, 1.0, 2.0]              # an invalid return: dropped
    xyz[900, 2] = np.inf
    path = tmp_path / "cloud.pcd"
    write_pcd(path, xyz, mode, coord_type=coord_type)
    got, dropped = E.read_pcd_xyz(path)
    keep = np.isfinite(xyz).all(axis=1)
    assert dropped == 2 and np.array_equal(got, xyz[keep])
    pm = E.VoxelHashMap(1.0, 30, device=-1)
    assert pm.AddPointsFromPcd(path) == keep.sum()
    qm = E.VoxelHashMap(1.0, 30, device=-1)
    qm.AddPoints(xyz[keep])
    assert np.array_equal(pm.export()["pxyz"], qm.export()["pxyz"])


def test_lzf_back_references_and_overlap(tmp_path):
    """a cloud whose field-major bytes are extremely repetitive (constant z, arithmetic x): long and overlapping matches"""
    n = 5000
    xyz = np.zeros((n, 3), np.float32)
    xyz[:, 0] = np.arange(n) % 4
    xyz[:, 2] = 1.25
    path = tmp_path / "rep.pcd"
    write_pcd(path, xyz, "binary_compressed", fields=FIELDS[:4])
    assert path.stat().st_size < n * 16 // 4     # it really compressed
    got, dropped = E.read_pcd_xyz(path)
    assert dropped == 0 and np.array_equal(got, xyz)


def test_pcd_errors(tmp_path):
    good = tmp_path / "g.pcd"
    write_pcd(good, synth.map_u(100, 5.0), "binary")
    raw = good.read_bytes()
    cases = {
        "trunc.pcd": raw[:-40],
        "nofields.pcd": raw.replace(b"FIELDS x y z", b"FIELDS a b c"),
        "nodata.pcd": raw[: raw.index(b"DATA")],
        "mode.pcd": raw.replace(b"DATA binary", b"DATA zipped"),
    }
    for name, content in cases.items():
        p = tmp_path / name
        p.write_bytes(content)
        with pytest.raises(E.ElmError) as ei:
            E.read_pcd_xyz(p)
        assert ei.value.status == _capi.ELM_ERR_IO, name
    comp = tmp_path / "c.pcd"
    write_pcd(comp, synth.map_u(300, 5.0), "binary_compressed")
    b = bytearray(comp.read_bytes())
    b[-5] ^= 0xFF
    b[-9] ^= 0x5A
    (tmp_path / "c_bad.pcd").write_bytes(bytes(b[:-3]))
    with pytest.raises(E.ElmError):
        E.read_pcd_xyz(tmp_path / "c_bad.pcd")
    with pytest.raises(E.ElmError):
        E.read_pcd_xyz(tmp_path / "missing.pcd")


def test_lzf_known_answer_streams(tmp_path):
    """Hand-assembled LZF streams straight from the liblzf format definition (independent of the encoder above):
      02 'a' 'b' 'c' | 20 02          literal "abc", then ctrl 0x20: length (1) + 2 = 3 bytes from distance (0 << 8 | 2) + 1 = 3
      00 'x' | E0 01 00               literal "x", then ctrl 0xE0: length 7 + next byte (1) + 2 = 10 bytes from distance 1 (overlap)
    wrapped as binary_compressed PCDs whose x, y, z are 1-byte unsigned fields (field-major: all x, all y, all z)."""
    def pcd(stream, n):
        hdr = (f"VERSION 0.7\nFIELDS x y z\nSIZE 1 1 1\nTYPE U U U\nCOUNT 1 1 1\nWIDTH {n}\nHEIGHT 1\nPOINTS {n}\nDATA binary_compressed\n").encode()
        return hdr + struct.pack("<II", len(stream), 3 * n) + stream
    p = tmp_path / "kat1.pcd"
    p.write_bytes(pcd(bytes([0x02]) + b"abc" + bytes([0x20, 0x02]), 2))          # "abcabc" -> x = a b, y = c a, z = b c
    got, _ = E.read_pcd_xyz(p)
    assert np.array_equal(got, np.array([[97, 99, 98], [98, 97, 99]], np.float32))
    p = tmp_path / "kat2.pcd"
    p.write_bytes(pcd(bytes([0x00]) + b"x" + bytes([0xE0, 0x01, 0x00]) + bytes([0x00]) + b"y", 4))   # 11 x's then "y"
    got, _ = E.read_pcd_xyz(p)
    assert np.array_equal(got, np.array([[120, 120, 120]] * 3 + [[120, 120, 121]], np.float32))
