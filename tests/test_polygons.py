"""Host half of updateRigidBodyHitbox / updateChunkMesh (world.cpp:497-563): csrc/polygons.hpp behind fse_hitbox_triangles against
THE REFERENCE'S OWN TPPL (oracle/_ref: physics_math.cpp compiled where it lies) — same triangles, same order, same grouping — plus
known answers that hold without the reference.  CPU only: the function is pure host code of the product library."""
import ctypes as C
import os

import numpy as np
import pytest

from falling_sand_engine_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libfse_ref_outline.so")


def _flat(contours):
    pts = np.concatenate([np.asarray(c, dtype=np.float32).reshape(-1, 2) for c in contours]) if contours else np.zeros((0, 2), np.float32)
    off = np.zeros(len(contours) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(c) for c in contours])
    return np.ascontiguousarray(pts), off


def ours(contours, cap_tris=4096, cap_groups=256):
    L = api.load_library()
    pts, off = _flat(contours)
    tris = np.zeros((cap_tris, 3, 2), dtype=np.float64)
    goff = np.zeros(cap_groups + 1, dtype=np.int32)
    ng = C.c_int32(0)
    L.fse_hitbox_triangles.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    rc = L.fse_hitbox_triangles(pts.ctypes.data, off.ctypes.data, len(contours), tris.ctypes.data, cap_tris, goff.ctypes.data, cap_groups, C.byref(ng))
    assert rc == 0, L.fse_last_error()
    return [tris[goff[g]:goff[g + 1]].copy() for g in range(ng.value)]


def theirs(contours, cap_tris=4096, cap_groups=256):
    L = C.CDLL(REF_LIB)
    pts, off = _flat(contours)
    tris = np.zeros((cap_tris, 3, 2), dtype=np.float64)
    goff = np.zeros(cap_groups + 1, dtype=np.int32)
    L.ref_hitbox_triangles.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    ng = L.ref_hitbox_triangles(pts.ctypes.data, off.ctypes.data, len(contours), tris.ctypes.data, cap_tris, goff.ctypes.data, cap_groups)
    assert ng >= 0
    return [tris[goff[g]:goff[g + 1]].copy() for g in range(ng)]


def _area(tris):
    a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
    return 0.5 * np.abs((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])).sum()


def test_square_and_square_with_hole_known_answers(oracle):
    """A filled block gives one polygon of two triangles covering its outline; a block with a hole gives one polygon whose triangles
    cover outline minus hole (the hole was bridged, not dropped); two separate blocks give two groups."""
    m = np.zeros((16, 20), dtype=np.uint8)
    m[3:11, 4:16] = 1
    cs = oracle.outlines(m)
    g = ours(cs)
    assert len(g) == 1 and len(g[0]) == len(cs[0]) - 2
    outline_area = 0.5 * abs(float(np.sum(cs[0][:, 0] * np.roll(cs[0][:, 1], -1) - np.roll(cs[0][:, 0], -1) * cs[0][:, 1])))
    assert abs(_area(g[0]) - outline_area) < 1e-9
    # a hole: an outline traced the other way round (the reference's scan finds holes only from candidates on diagonal steps, so
    # this one is written by hand)
    hole = np.array([[7, 5], [12, 5], [12, 8], [7, 8]], dtype=np.float32)
    cs = [cs[0], hole]
    g = ours(cs)
    assert len(g) == 1
    areas = [0.5 * abs(float(np.sum(c[:, 0] * np.roll(c[:, 1], -1) - np.roll(c[:, 0], -1) * c[:, 1]))) for c in cs]
    assert abs(_area(g[0]) - (areas[0] - areas[1])) < 1e-9 and len(g[0]) == (4 + 4 + 2) - 2
    if os.path.exists(REF_LIB):
        t = theirs(cs)
        assert len(t) == 1 and np.array_equal(t[0], g[0])
    m2 = np.zeros((16, 40), dtype=np.uint8)
    m2[3:11, 4:16] = 1
    m2[2:9, 22:35] = 1
    assert len(ours(oracle.outlines(m2))) == 2


def test_degenerate_inputs():
    assert ours([]) == []
    assert ours([np.array([[0, 0], [1, 0]], dtype=np.float32)]) == []                       # fewer than 3 points: no polygon
    assert ours([np.array([[0, 0], [1, 0], [2, 0]], dtype=np.float32)]) == []               # one collinear triangle: dropped (world.cpp:558)
    tri = ours([np.array([[0, 0], [0, 4], [3, 0]], dtype=np.float32)])
    assert len(tri) == 1 and len(tri[0]) == 1


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built (make -C oracle -f Makefile.ref needs /root/reference)")
@pytest.mark.parametrize("seed,fill", [(1, 0.5), (2, 0.7), (3, 0.85), (4, 0.95)])
def test_triangles_equal_the_reference_tppl(oracle, seed, fill):
    """Masks of rectangles, rectangular holes and hashed holes (rigid bodies of BASELINE configs[3] are 70 % fill): the triangle groups
    equal the reference's RemoveHoles + Triangulate_EC output bit for bit — same count, same order, same vertices."""
    rng = np.random.default_rng(seed)
    checked = tri_total = holes = 0
    for k in range(40):
        h, w = int(rng.integers(10, 36)), int(rng.integers(10, 44))
        m = np.zeros((h, w), dtype=np.uint8)
        for _ in range(int(rng.integers(1, 5))):
            x0, y0 = int(rng.integers(1, w - 4)), int(rng.integers(1, h - 4))
            m[y0:min(h - 1, y0 + int(rng.integers(3, 16))), x0:min(w - 1, x0 + int(rng.integers(3, 20)))] = 1
        if k % 2:
            m &= (rng.random((h, w)) < fill).astype(np.uint8)
        else:
            for _ in range(int(rng.integers(0, 4))):
                x0, y0 = int(rng.integers(2, w - 3)), int(rng.integers(2, h - 3))
                m[y0:y0 + int(rng.integers(1, 4)), x0:x0 + int(rng.integers(1, 5))] = 0
        cs = oracle.outlines(m)
        if not cs:
            continue
        a, b = theirs(cs), ours(cs)
        assert len(a) == len(b), (k, len(a), len(b))
        for ga, gb in zip(a, b):
            assert ga.shape == gb.shape and np.array_equal(ga, gb), k
            tri_total += len(ga)
        holes += len(cs) - len(a)
        checked += 1
    assert checked >= 30 and tri_total > 200 and holes > 0
