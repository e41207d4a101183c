"""Ground truth for the restated third-party arithmetic that does NOT depend on the restatement itself (the packages --
commonroad-drivability-checker's pycrcc / pycrccosy -- are not installable here):

* the separating-axis tests (box / box, box / triangle; closed sets) against an EXACT decision procedure in rational
  arithmetic (vertex containment + segment intersection with orientation predicates on ``fractions.Fraction``);
* ``convert_to_curvilinear_coords(convert_to_cartesian_coords(s, d)) == (s, d)`` inside the projection domain of curved
  paths;
* the OBB-sum hull of two boxes contains both and is tight along the first box's axes.

The oracle (oracle/third_party.py) is tested on the CPU; the CUDA restatements (rp_collide_poses, rp_initial_states) on
the GPU against the same exact procedure."""
import math
from fractions import Fraction as F

import numpy as np
import pytest

from oracle import third_party as tp


# ---- exact polygon intersection (closed sets) in rational arithmetic ------------------------------------------------
def _orient(a, b, c):
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])


def _on_segment(a, b, p):
    return _orient(a, b, p) == 0 and min(a[0], b[0]) <= p[0] <= max(a[0], b[0]) and min(a[1], b[1]) <= p[1] <= max(a[1], b[1])


def _segments_intersect(a, b, c, d):
    o1, o2, o3, o4 = _orient(a, b, c), _orient(a, b, d), _orient(c, d, a), _orient(c, d, b)
    if ((o1 > 0) != (o2 > 0)) and o1 != 0 and o2 != 0 and ((o3 > 0) != (o4 > 0)) and o3 != 0 and o4 != 0:
        return True
    return _on_segment(a, b, c) or _on_segment(a, b, d) or _on_segment(c, d, a) or _on_segment(c, d, b)


def _inside_convex(poly, p):
    """p inside or on the boundary of the convex polygon (either orientation)"""
    signs = [_orient(poly[k], poly[(k + 1) % len(poly)], p) for k in range(len(poly))]
    return all(s >= 0 for s in signs) or all(s <= 0 for s in signs)


def exact_intersect(pa, pb):
    if any(_inside_convex(pb, p) for p in pa) or any(_inside_convex(pa, p) for p in pb):
        return True
    for k in range(len(pa)):
        for m in range(len(pb)):
            if _segments_intersect(pa[k], pa[(k + 1) % len(pa)], pb[m], pb[(m + 1) % len(pb)]):
                return True
    return False


def exact_gap(pa, pb):
    """squared distance between the two convex polygons (0 if they intersect), as float -- to skip near-degenerate draws"""
    if exact_intersect(pa, pb):
        return 0.0
    best = None
    for P, Q in ((pa, pb), (pb, pa)):
        for p in P:
            for m in range(len(Q)):
                a, b = Q[m], Q[(m + 1) % len(Q)]
                ab = (b[0] - a[0], b[1] - a[1])
                t = ((p[0] - a[0]) * ab[0] + (p[1] - a[1]) * ab[1]) / (ab[0] * ab[0] + ab[1] * ab[1])
                t = max(F(0), min(F(1), t))
                q = (a[0] + t * ab[0], a[1] + t * ab[1])
                d2 = (p[0] - q[0]) ** 2 + (p[1] - q[1]) ** 2
                best = d2 if best is None or d2 < best else best
    return float(best)


def box_corners_exact(cx, cy, theta, hl, hw):
    """the box the SAT sees: centre and half extents as given, axes (cos, sin) as the floats libm returns, in rationals"""
    c, s = F(math.cos(theta)), F(math.sin(theta))
    cx, cy, hl, hw = F(cx), F(cy), F(hl), F(hw)
    return [(cx + sx * hl * c - sy * hw * s, cy + sx * hl * s + sy * hw * c) for sx, sy in ((1, 1), (-1, 1), (-1, -1), (1, -1))]


def _random_pairs(rng, n):
    for _ in range(n):
        a = (rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-math.pi, math.pi), rng.uniform(0.3, 3.0), rng.uniform(0.2, 1.5))
        b = (rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-math.pi, math.pi), rng.uniform(0.3, 3.0), rng.uniform(0.2, 1.5))
        yield a, b


TOUCHING = [   # axis-aligned, integer data: exact in floating point.  (a, b, expected)
    ((0, 0, 0.0, 2, 1), (4, 0, 0.0, 2, 1), True),        # faces touch along x = 2
    ((0, 0, 0.0, 2, 1), (4, 2, 0.0, 2, 1), True),        # corners touch at (2, 1)
    ((0, 0, 0.0, 2, 1), (4.000001, 0, 0.0, 2, 1), False),
    ((0, 0, 0.0, 2, 1), (0, 0, 0.0, 1, 0.5), True),      # containment
    ((0, 0, 0.0, 2, 1), (0, 2.5, 0.0, 5, 1.5), True),    # faces touch along y = 1
    ((0, 0, 0.0, 2, 1), (0, 2.5000001, 0.0, 5, 1.5), False),
]


def test_box_box_sat_equals_exact_polygon_intersection():
    rng = np.random.default_rng(7)
    n_hit = n_checked = 0
    for a, b in _random_pairs(rng, 1500):
        pa, pb = box_corners_exact(*a), box_corners_exact(*b)
        want = exact_intersect(pa, pb)
        if not want and exact_gap(pa, pb) < 1e-18:
            continue                                       # separated by less than 1e-9 m: beyond what doubles decide
        if want and not exact_intersect(pa, box_corners_exact(b[0], b[1], b[2], b[3] - 1e-9, b[4] - 1e-9)):
            continue                                       # touching by less than 1e-9 m
        got = tp.obb_obb_overlap(tp.RectOBB(a[3], a[4], a[2], a[0], a[1]), tp.RectOBB(b[3], b[4], b[2], b[0], b[1]))
        assert got == want, (a, b)
        n_hit += want
        n_checked += 1
    assert n_checked > 1400 and 300 < n_hit < 1300
    for a, b, want in TOUCHING:                            # closed sets: touching counts as collision
        assert exact_intersect(box_corners_exact(*a), box_corners_exact(*b)) == want
        assert tp.obb_obb_overlap(tp.RectOBB(a[3], a[4], a[2], a[0], a[1]), tp.RectOBB(b[3], b[4], b[2], b[0], b[1])) == want, (a, b)


def _random_triangles(rng, n):
    for _ in range(n):
        box = (rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(-math.pi, math.pi), rng.uniform(0.5, 3.0), rng.uniform(0.3, 1.5))
        c = rng.uniform(-4, 4, 2)
        tri = (c + rng.uniform(-2.5, 2.5, (3, 2))).ravel()
        yield box, tri


def test_box_triangle_sat_equals_exact_polygon_intersection():
    rng = np.random.default_rng(8)
    n_hit = n_checked = 0
    for box, tri in _random_triangles(rng, 1500):
        pt = [(F(float(tri[0])), F(float(tri[1]))), (F(float(tri[2])), F(float(tri[3]))), (F(float(tri[4])), F(float(tri[5])))]
        if abs(float(_orient(*pt))) < 1e-6:
            continue                                       # degenerate triangle
        pa = box_corners_exact(*box)
        want = exact_intersect(pa, pt)
        if not want and exact_gap(pa, pt) < 1e-18:
            continue
        got = tp.obb_triangle_overlap(tp.RectOBB(box[3], box[4], box[2], box[0], box[1]), tp.Triangle(*tri))
        assert got == want, (box, tri)
        n_hit += want
        n_checked += 1
    assert n_checked > 1400 and 200 < n_hit < 1300
    # touching: a triangle vertex on the box's face, an edge along the face
    b = tp.RectOBB(2, 1, 0.0, 0, 0)
    assert tp.obb_triangle_overlap(b, tp.Triangle(2, 0, 4, 1, 4, -1))
    assert tp.obb_triangle_overlap(b, tp.Triangle(2, -3, 2, 3, 5, 0))
    assert not tp.obb_triangle_overlap(b, tp.Triangle(2.000001, 0, 4, 1, 4, -1))


def _curved_paths():
    u = np.linspace(0.0, 1.0, 240)
    return {"sine": np.stack([np.arange(0, 300, 1.0), 20.0 * np.sin(np.arange(0, 300, 1.0) / 40.0)], axis=1),
            "tight": np.stack([np.arange(0, 200, 1.0), 20.0 * np.sin(np.arange(0, 200, 1.0) / 25.0)], axis=1),
            "arc": np.stack([60.0 * np.cos(1.5 * np.pi * u), 60.0 * np.sin(1.5 * np.pi * u)], axis=1)}


@pytest.mark.parametrize("name", ["sine", "tight", "arc"])
def test_curvilinear_round_trip_is_the_identity_inside_the_domain(name):
    """(s, d) -> (x, y) -> (s, d): the pseudo-normal projection and its inverse (a quadratic per segment) are consistent
    wherever the lateral offset stays below the path's radius of curvature"""
    from oracle import rp_oracle as O
    path = _curved_paths()[name]
    ref, ccosy, _ = O.reference_tables(path)
    cc = tp.CurvilinearCoordinateSystem.__new__(tp.CurvilinearCoordinateSystem)
    frame = O._ArrayCCosy(ccosy)
    rng = np.random.default_rng(3)
    S = np.asarray(ccosy["S"])
    kmax = float(np.max(np.abs(ref["ref_curv"])))
    dmax = min(6.0, 0.5 / max(kmax, 1e-9))
    n = 0
    for _ in range(400):
        s = rng.uniform(S[2], S[-3])
        d = rng.uniform(-dmax, dmax)
        xy = frame.convert_to_cartesian_coords(s, d)
        s2, d2 = frame.convert_to_curvilinear_coords(xy[0], xy[1])
        assert abs(s2 - s) < 1e-8 and abs(d2 - d) < 1e-8, (name, s, d, s2, d2)
        n += 1
    assert n == 400
    # geometric meaning: (x, y) lies at distance ~|d| from the polyline (pseudo-normals are unit at the vertices)
    xy = frame.convert_to_cartesian_coords(float(S[10]), 2.0)
    P = np.asarray(ccosy["path"])
    assert abs(np.min(np.hypot(P[:, 0] - xy[0], P[:, 1] - xy[1])) - 2.0) < 0.05


def _corners(b):
    c, s = math.cos(b.orientation), math.sin(b.orientation)
    return [(b.cx + sx * b.r_x * c - sy * b.r_y * s, b.cy + sx * b.r_x * s + sy * b.r_y * c) for sx, sy in ((1, 1), (-1, 1), (-1, -1), (1, -1))]


@pytest.mark.parametrize("which", ["oracle", "product"])
def test_obb_sum_hull_contains_both_boxes_and_is_tight(which):
    if which == "oracle":
        mk, hull = tp.RectOBB, tp.obb_sum_hull
    else:
        from commonroad_rp_b200 import collision
        mk, hull = collision.RectOBB, collision.obb_sum_hull
    rng = np.random.default_rng(11)
    for _ in range(300):
        a = mk(rng.uniform(0.5, 3), rng.uniform(0.3, 1.5), rng.uniform(-math.pi, math.pi), rng.uniform(-5, 5), rng.uniform(-5, 5))
        b = mk(rng.uniform(0.5, 3), rng.uniform(0.3, 1.5), a.orientation + rng.uniform(-0.6, 0.6), a.cx + rng.uniform(-3, 3),
               a.cy + rng.uniform(-3, 3))
        h = hull(a, b)
        assert h.orientation == a.orientation
        c, s = math.cos(h.orientation), math.sin(h.orientation)
        us, vs = [], []
        for px, py in _corners(a) + _corners(b):
            u, v = (px - h.cx) * c + (py - h.cy) * s, -(px - h.cx) * s + (py - h.cy) * c
            assert abs(u) <= h.r_x + 1e-12 and abs(v) <= h.r_y + 1e-12          # contains both boxes
            us.append(u)
            vs.append(v)
        # tight along the first box's axes: a corner touches each of the four faces
        assert max(us) == pytest.approx(h.r_x, abs=1e-12) and min(us) == pytest.approx(-h.r_x, abs=1e-12)
        assert max(vs) == pytest.approx(h.r_y, abs=1e-12) and min(vs) == pytest.approx(-h.r_y, abs=1e-12)


# ---- the CUDA restatements against the same exact procedure --------------------------------------------------------
@pytest.mark.gpu
def test_device_sat_equals_exact_polygon_intersection():
    from commonroad_rp_b200._lib import Engine
    rng = np.random.default_rng(21)
    hl, hw = 2.254, 0.805
    eng = Engine(0)
    eng.set_vehicle(2 * hl, 2 * hw, 0.0, 2.5, 10.0, 7.0, 1.0, 0.4)
    n_checked = n_hit = 0
    for rep in range(6):
        boxes = np.array([[rng.uniform(-8, 8), rng.uniform(-8, 8), rng.uniform(-math.pi, math.pi), rng.uniform(0.3, 3.0),
                           rng.uniform(0.2, 1.5)] for _ in range(3)])
        tris = np.array([(rng.uniform(-8, 8, 2)[None, :] + rng.uniform(-2.5, 2.5, (3, 2))).ravel() for _ in range(3)])
        eng.set_obstacles(boxes, None, None, tris)
        poses = np.stack([rng.uniform(-10, 10, 250), rng.uniform(-10, 10, 250), rng.uniform(-math.pi, math.pi, 250)], axis=1)
        got = eng.collide_poses(poses, np.zeros(len(poses), dtype=np.int32), hl, hw)
        polys = [box_corners_exact(*b) for b in boxes] + \
                [[(F(float(t[0])), F(float(t[1]))), (F(float(t[2])), F(float(t[3]))), (F(float(t[4])), F(float(t[5])))] for t in tris]
        for q, (x, y, th) in enumerate(poses):
            ego = box_corners_exact(x, y, th, hl, hw)
            gaps = [exact_gap(ego, pb) for pb in polys]
            want = min(gaps) == 0.0
            if not want and min(gaps) < 1e-16:
                continue
            assert bool(got[q]) == want, (rep, q)
            n_checked += 1
            n_hit += want
    assert n_checked > 1400 and 100 < n_hit < 1300
    # closed sets on exact data: touching faces / corners collide
    eng.set_obstacles(np.array([[4.0, 0.0, 0.0, 2.0, 1.0], [4.0, 10.0, 0.0, 2.0, 1.0]]), None, None, None)
    eng.set_vehicle(4.0, 2.0, 0.0, 2.5, 10.0, 7.0, 1.0, 0.4)
    got = eng.collide_poses(np.array([[0.0, 0.0, 0.0], [0.0, 2.0, 0.0], [-0.000001, 0.0, 0.0], [0.0, 8.0, 0.0], [0.0, 7.999999, 0.0]]),
                            np.zeros(5, dtype=np.int32), 2.0, 1.0)
    assert list(got) == [True, True, False, True, False]
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sine", "tight", "arc"])
def test_device_curvilinear_round_trip(name):
    """rp_initial_states' Cartesian -> curvilinear projection inverts the frame's (s, d) -> (x, y) map on the device tables"""
    from commonroad_rp_b200._lib import Engine
    from commonroad_rp_b200.utility.utils_coordinate_system import CoordinateSystem
    co = CoordinateSystem(_curved_paths()[name])
    eng = Engine(0)
    eng.set_vehicle(4.5, 1.6, 1.4, 2.58, 11.5, 7.3, 1.066, 0.4)
    tb = co.device_tables()
    eng.set_reference(tb["ref_pos"], tb["ref_theta"], tb["ref_curv"], tb["ref_curv_d"], tb["path_xy"], tb["path_s"],
                      tb["path_normals"], tb["proj_limit"])
    rng = np.random.default_rng(5)
    S = tb["path_s"]
    dmax = min(6.0, 0.5 / max(float(np.max(np.abs(tb["ref_curv"]))), 1e-9))
    sd = np.stack([rng.uniform(S[2], S[-3], 300), rng.uniform(-dmax, dmax, 300)], axis=1)
    xy = np.array([co.convert_to_cartesian_coords(s, d) for s, d in sd])
    j = np.searchsorted(tb["ref_pos"], sd[:, 0]) - 1
    x0 = np.stack([xy[:, 0], xy[:, 1], tb["ref_theta"][j], np.full(300, 5.0), np.zeros(300), np.zeros(300)], axis=1)
    lon, lat, status = eng.initial_states(x0, 0)
    assert not status.any()
    assert np.max(np.abs(lon[:, 0] - sd[:, 0])) < 1e-8 and np.max(np.abs(lat[:, 0] - sd[:, 1])) < 1e-8
    eng.close()
