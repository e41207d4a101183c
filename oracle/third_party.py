"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

CPU restatement of the THIRD-PARTY arithmetic the reference's hot path calls but does
not ship (the packages are un-vendored pip dependencies, absent from /root/reference):

  * commonroad-drivability-checker 2024.1 (poetry.lock:163-164)
      - pycrccosy.CurvilinearCoordinateSystem  (utility/utils_coordinate_system.py:128-129, :170, :178)
      - pycrcc.{RectOBB, RectAABB, Triangle, ShapeGroup, TimeVariantCollisionObject, CollisionChecker}
        (reactive_planner.py:234-251, :1040-1042)
      - commonroad_dc.geometry.util.{compute_pathlength,curvature,orientation}_from_polyline,
        resample_polyline  (utility/utils_coordinate_system.py:114-117, :57, :68, :82)
  * commonroad-io 2024.1 (poetry.lock:220-221)
      - commonroad.common.util.make_valid_orientation  (utility/utils_coordinate_system.py:43)
  * commonroad-vehicle-models 3.0.2 (poetry.lock:265-266): parameters of vehicle type 2
    (utility/config.py:198-219)

PARITY UNPINNED: the reference ships no tests / golden vectors for any of these and the
real packages cannot be installed here (no network).  The semantics below are the
published algorithms restated from memory; the GPU product is checked
against THIS restatement.  Everything in commonroad_rp itself is pinned by running the
reference's own code (oracle/ref_shims.py).
"""
import math

import numpy as np

TWO_PI = 2.0 * np.pi


# ----------------------------------------------------------------------------------------------
# commonroad.common.util.make_valid_orientation  ([-2pi, 2pi] form)
# ----------------------------------------------------------------------------------------------
def make_valid_orientation(angle):
    while angle > TWO_PI:
        angle = angle - TWO_PI
    while angle < -TWO_PI:
        angle = angle + TWO_PI
    return angle


# ----------------------------------------------------------------------------------------------
# commonroad_dc.geometry.util  (setup-time only)
# ----------------------------------------------------------------------------------------------
def compute_pathlength_from_polyline(polyline):
    polyline = np.asarray(polyline, dtype=np.float64)
    dist = np.zeros(len(polyline))
    for i in range(1, len(polyline)):
        dist[i] = dist[i - 1] + np.linalg.norm(polyline[i] - polyline[i - 1])
    return dist


def compute_orientation_from_polyline(polyline):
    polyline = np.asarray(polyline, dtype=np.float64)
    n = len(polyline)
    out = np.zeros(n)
    for i in range(n - 1):
        seg = polyline[i + 1] - polyline[i]
        out[i] = np.arctan2(seg[1], seg[0])
    out[n - 1] = out[n - 2]
    return out


def compute_curvature_from_polyline(polyline):
    polyline = np.asarray(polyline, dtype=np.float64)
    pl = compute_pathlength_from_polyline(polyline)
    x_d = np.gradient(polyline[:, 0], pl)
    x_dd = np.gradient(x_d, pl)
    y_d = np.gradient(polyline[:, 1], pl)
    y_dd = np.gradient(y_d, pl)
    return (x_d * y_dd - x_dd * y_d) / ((x_d ** 2 + y_d ** 2) ** (3.0 / 2.0))


def resample_polyline(polyline, step=2.0):
    polyline = np.asarray(polyline, dtype=np.float64)
    if len(polyline) < 2:
        return polyline.copy()
    new = [polyline[0]]
    pos = step
    seg_len = np.linalg.norm(polyline[0] - polyline[1])
    idx = 0
    while idx < len(polyline) - 1:
        if pos >= seg_len:
            pos -= seg_len
            idx += 1
            if idx > len(polyline) - 2:
                break
            seg_len = np.linalg.norm(polyline[idx + 1] - polyline[idx])
        else:
            rel = pos / seg_len
            new.append((1.0 - rel) * polyline[idx] + rel * polyline[idx + 1])
            pos += step
    if np.linalg.norm(new[-1] - polyline[-1]) >= 1e-6:
        new.append(polyline[-1])
    return np.array(new)


def chaikins_corner_cutting(polyline, refinements=1):
    pts = np.asarray(polyline, dtype=np.float64)
    for _ in range(refinements):
        L = pts.repeat(2, axis=0)
        R = np.empty_like(L)
        R[0] = L[0]
        R[2::2] = L[1:-1:2]
        R[1:-1:2] = L[2::2]
        R[-1] = L[-1]
        pts = L * 0.75 + R * 0.25
    return pts


# ----------------------------------------------------------------------------------------------
# pycrccosy.CurvilinearCoordinateSystem  
# ----------------------------------------------------------------------------------------------
class CurvilinearCoordinateSystem:
    """Polyline curvilinear frame with pseudo-normal projection.

    ctor: copy of the polyline extended by one vertex at each end, offset ``eps2`` along
    the first / last segment.  Vertex pseudo-tangents: end vertices use the adjacent
    segment direction, interior vertices the normalised chord p[i+1]-p[i-1].
    (s,d)->(x,y): segment j with S[j] <= s <= S[j+1]; lam=(s-S[j])/len_j;
    base = p_j + lam*(p_{j+1}-p_j); pseudo-normal = n_j + lam*(n_{j+1}-n_j) (NOT
    re-normalised) with n = 90deg rotation of the tangent; result = base + d*pn.
    Raises ValueError outside the longitudinal range or beyond the lateral projection
    limit (the wrapper turns any exception into None, utils_coordinate_system.py:169-172).
    """

    def __init__(self, reference_path, default_projection_domain_limit=20.0, eps=0.1, eps2=1e-4):
        ref = np.asarray(reference_path, dtype=np.float64)
        if ref.ndim != 2 or ref.shape[0] < 3 or ref.shape[1] != 2:
            raise ValueError("reference path must be an (n>=3, 2) polyline")
        d0 = ref[1] - ref[0]
        d1 = ref[-1] - ref[-2]
        first = ref[0] - eps2 * d0 / np.linalg.norm(d0)
        last = ref[-1] + eps2 * d1 / np.linalg.norm(d1)
        self._path = np.vstack([first[None, :], ref, last[None, :]])
        self._limit = float(default_projection_domain_limit)
        self._S = compute_pathlength_from_polyline(self._path)
        n = len(self._path)
        tang = np.zeros((n, 2))
        tang[0] = self._path[1] - self._path[0]
        tang[-1] = self._path[-1] - self._path[-2]
        tang[1:-1] = self._path[2:] - self._path[:-2]
        tang = tang / np.linalg.norm(tang, axis=1)[:, None]
        self._tangent = tang
        self._normal = np.stack([-tang[:, 1], tang[:, 0]], axis=1)

    # -- accessors used by the reference wrapper --
    def reference_path(self):
        return [p.copy() for p in self._path]

    # -- arrays for packing into device tables / the numpy port --
    @property
    def path(self):
        return self._path

    @property
    def pathlength(self):
        return self._S

    @property
    def normals(self):
        return self._normal

    @property
    def projection_domain_limit(self):
        return self._limit

    def convert_to_cartesian_coords(self, s, d):
        S = self._S
        if not (s >= S[0] and s <= S[-1]):  # also rejects NaN
            raise ValueError("longitudinal coordinate outside of projection domain")
        if not (abs(d) <= self._limit):
            raise ValueError("lateral coordinate outside of projection domain")
        # first index with S[j] > s, minus one; s == S[-1] belongs to the last segment
        j = int(np.searchsorted(S, s, side="right")) - 1
        if j > len(S) - 2:
            j = len(S) - 2
        lam = (s - S[j]) / (S[j + 1] - S[j])
        p0 = self._path[j]
        p1 = self._path[j + 1]
        n0 = self._normal[j]
        n1 = self._normal[j + 1]
        bx = p0[0] + lam * (p1[0] - p0[0])
        by = p0[1] + lam * (p1[1] - p0[1])
        nx = n0[0] + lam * (n1[0] - n0[0])
        ny = n0[1] + lam * (n1[1] - n0[1])
        return np.array([bx + d * nx, by + d * ny])

    def convert_to_curvilinear_coords(self, x, y):
        """Inverse of the pseudo-normal map (per-segment quadratic in lam); smallest |d| wins."""
        best = None
        p = np.array([x, y], dtype=np.float64)
        for j in range(len(self._path) - 1):
            p0 = self._path[j]
            e = self._path[j + 1] - p0
            m = self._normal[j]
            dn = self._normal[j + 1] - m
            a = p - p0
            cr = lambda u, v: u[0] * v[1] - u[1] * v[0]
            qa = -cr(e, dn)
            qb = cr(a, dn) - cr(e, m)
            qc = cr(a, m)
            roots = []
            if abs(qa) < 1e-14:
                if abs(qb) > 0.0:
                    roots.append(-qc / qb)
            else:
                disc = qb * qb - 4.0 * qa * qc
                if disc >= 0.0:
                    # cancellation-free form of (-qb +- sqrt(disc)) / (2 qa): on gently curved paths qa is tiny and the
                    # textbook formula loses half the digits (1e-7 m on s; tests/test_third_party_properties.py)
                    sq = math.sqrt(disc)
                    qq = -0.5 * (qb + math.copysign(sq, qb))
                    if qq == 0.0:
                        roots.extend([0.0, 0.0])
                    elif qb >= 0.0:
                        roots.extend([qc / qq, qq / qa])           # (-qb + sq) / 2qa, (-qb - sq) / 2qa
                    else:
                        roots.extend([qq / qa, qc / qq])
            for lam in roots:
                if -1e-12 <= lam <= 1.0 + 1e-12:
                    lam = min(max(lam, 0.0), 1.0)
                    base = p0 + lam * e
                    pn = m + lam * dn
                    dd = float(np.dot(p - base, pn) / np.dot(pn, pn))
                    if abs(dd) <= self._limit and (best is None or abs(dd) < abs(best[1])):
                        best = (float(self._S[j] + lam * (self._S[j + 1] - self._S[j])), dd)
        if best is None:
            raise ValueError("point outside of projection domain")
        return np.array(best)


# ----------------------------------------------------------------------------------------------
# pycrcc shapes + checker  (SURVEY App. D#2).  Closed sets: touching counts as collision,
# i.e. "separated iff projected gap > 0".
# ----------------------------------------------------------------------------------------------
class RectOBB:
    def __init__(self, r_x, r_y, orientation, cx, cy):
        self.r_x = float(r_x)
        self.r_y = float(r_y)
        self.orientation = float(orientation)
        self.cx = float(cx)
        self.cy = float(cy)

    def center(self):
        return np.array([self.cx, self.cy])


class RectAABB(RectOBB):
    def __init__(self, r_x, r_y, cx, cy):
        super().__init__(r_x, r_y, 0.0, cx, cy)


class Triangle:
    def __init__(self, x1, y1, x2, y2, x3, y3):
        self.v = np.array([[x1, y1], [x2, y2], [x3, y3]], dtype=np.float64)


class ShapeGroup:
    def __init__(self):
        self.shapes = []
        self._bound = None

    def add_shape(self, shape):
        self.shapes.append(shape)
        self._bound = None

    def near(self, box):
        """Members whose bounding circle reaches the bounding circle of ``box`` (exact pre-selection:
        shapes that fail this cannot overlap; keeps the pure-Python narrow phase affordable)."""
        if self._bound is None:
            c = np.zeros((len(self.shapes), 2))
            r = np.zeros(len(self.shapes))
            for k, sh in enumerate(self.shapes):
                if isinstance(sh, Triangle):
                    c[k] = sh.v.mean(axis=0)
                    r[k] = np.max(np.linalg.norm(sh.v - c[k], axis=1))
                else:
                    c[k] = (sh.cx, sh.cy)
                    r[k] = math.hypot(sh.r_x, sh.r_y)
            self._bound = (c, r)
        c, r = self._bound
        if len(r) == 0:
            return []
        rb = math.hypot(box.r_x, box.r_y)
        reach = (r + rb) * (1.0 + 1e-9) + 1e-9
        hit = (c[:, 0] - box.cx) ** 2 + (c[:, 1] - box.cy) ** 2 <= reach * reach
        return [self.shapes[k] for k in np.nonzero(hit)[0]]


class TimeVariantCollisionObject:
    def __init__(self, time_start_idx):
        self._t0 = int(time_start_idx)
        self._obstacles = []

    def append_obstacle(self, obstacle):
        self._obstacles.append(obstacle)

    def time_start_idx(self):
        return self._t0

    def time_end_idx(self):
        return self._t0 + len(self._obstacles) - 1

    def obstacle_at_time(self, t):
        k = t - self._t0
        if 0 <= k < len(self._obstacles):
            return self._obstacles[k]
        return None


def obb_sum_hull(a, b):
    """commonroad_dc's "OBB sum hull" of two consecutive boxes (trajectory_preprocess_obb_sum; continuous collision
    check, reactive_planner.py:240-241, :1049-1058).  PARITY UNPINNED: the package is not installed and the reference
    ships no vectors; restated as the tight box ALONG THE FIRST BOX'S AXES enclosing both boxes."""
    ca, sa = math.cos(a.orientation), math.sin(a.orientation)
    cb, sb = math.cos(b.orientation), math.sin(b.orientation)
    dx = b.cx - a.cx
    dy = b.cy - a.cy
    u0 = dx * ca + dy * sa
    v0 = -dx * sa + dy * ca
    c = ca * cb + sa * sb
    s = ca * sb - sa * cb
    eu = b.r_x * abs(c) + b.r_y * abs(s)
    ev = b.r_x * abs(s) + b.r_y * abs(c)
    umin, umax = min(-a.r_x, u0 - eu), max(a.r_x, u0 + eu)
    vmin, vmax = min(-a.r_y, v0 - ev), max(a.r_y, v0 + ev)
    um, vm = 0.5 * (umin + umax), 0.5 * (vmin + vmax)
    return RectOBB(0.5 * (umax - umin), 0.5 * (vmax - vmin), a.orientation, a.cx + um * ca - vm * sa, a.cy + um * sa + vm * ca)


def trajectory_preprocess_obb_sum(tvo):
    """commonroad_dc.collision.trajectory_queries.trajectory_preprocess_obb_sum: a time-variant object of n boxes ->
    one of n - 1 hull boxes (steps k, k + 1) starting at the same time index; (object, error code)."""
    out = TimeVariantCollisionObject(tvo.time_start_idx())
    obs = tvo._obstacles
    for k in range(len(obs) - 1):
        if not isinstance(obs[k], RectOBB) or not isinstance(obs[k + 1], RectOBB):
            return None, -1
        out.append_obstacle(obb_sum_hull(obs[k], obs[k + 1]))
    return out, 0


def obb_obb_overlap(a, b):
    """4-axis SAT on two oriented boxes; separated iff gap > 0 on some axis."""
    ca, sa = math.cos(a.orientation), math.sin(a.orientation)
    cb, sb = math.cos(b.orientation), math.sin(b.orientation)
    dx = b.cx - a.cx
    dy = b.cy - a.cy
    # axes of a
    c = ca * cb + sa * sb        # ua . ub
    s = ca * sb - sa * cb        # ua x ub
    ac = abs(c)
    as_ = abs(s)
    if abs(dx * ca + dy * sa) > a.r_x + (b.r_x * ac + b.r_y * as_):
        return False
    if abs(-dx * sa + dy * ca) > a.r_y + (b.r_x * as_ + b.r_y * ac):
        return False
    if abs(dx * cb + dy * sb) > b.r_x + (a.r_x * ac + a.r_y * as_):
        return False
    if abs(-dx * sb + dy * cb) > b.r_y + (a.r_x * as_ + a.r_y * ac):
        return False
    return True


def obb_triangle_overlap(a, tri):
    """SAT with the 2 box axes + 3 triangle edge normals."""
    ca, sa = math.cos(a.orientation), math.sin(a.orientation)
    rel = tri.v - np.array([a.cx, a.cy])
    # box frame coordinates of the triangle vertices
    px = [r[0] * ca + r[1] * sa for r in rel]
    py = [-r[0] * sa + r[1] * ca for r in rel]
    if min(px) > a.r_x or max(px) < -a.r_x:
        return False
    if min(py) > a.r_y or max(py) < -a.r_y:
        return False
    for k in range(3):
        x0, y0 = px[k], py[k]
        x1, y1 = px[(k + 1) % 3], py[(k + 1) % 3]
        nx, ny = -(y1 - y0), (x1 - x0)
        tp = [px[i] * nx + py[i] * ny for i in range(3)]
        rb = a.r_x * abs(nx) + a.r_y * abs(ny)
        if min(tp) > rb or max(tp) < -rb:
            return False
    return True


def shapes_overlap(a, b):
    if isinstance(a, ShapeGroup):
        members = a.near(b) if isinstance(b, RectOBB) else a.shapes
        return any(shapes_overlap(s, b) for s in members)
    if isinstance(b, ShapeGroup):
        members = b.near(a) if isinstance(a, RectOBB) else b.shapes
        return any(shapes_overlap(a, s) for s in members)
    if isinstance(a, RectOBB) and isinstance(b, RectOBB):
        return obb_obb_overlap(a, b)
    if isinstance(a, RectOBB) and isinstance(b, Triangle):
        return obb_triangle_overlap(a, b)
    if isinstance(a, Triangle) and isinstance(b, RectOBB):
        return obb_triangle_overlap(b, a)
    raise TypeError("unsupported shape pair %r / %r" % (type(a), type(b)))


class CollisionChecker:
    def __init__(self):
        self._objects = []

    def add_collision_object(self, obj):
        self._objects.append(obj)

    def obstacles(self):
        return list(self._objects)

    def collide(self, obj):
        """any object of the checker intersects obj; time-variant vs time-variant only at
        common time indices, static objects at every index of obj."""
        if isinstance(obj, TimeVariantCollisionObject):
            for t in range(obj.time_start_idx(), obj.time_end_idx() + 1):
                ego = obj.obstacle_at_time(t)
                for other in self._objects:
                    if isinstance(other, TimeVariantCollisionObject):
                        o = other.obstacle_at_time(t)
                        if o is None:
                            continue
                    else:
                        o = other
                    if shapes_overlap(ego, o):
                        return True
            return False
        for other in self._objects:
            if isinstance(other, TimeVariantCollisionObject):
                raise TypeError("static query against a time-variant obstacle")
            if shapes_overlap(obj, other):
                return True
        return False


# ----------------------------------------------------------------------------------------------
# commonroad-vehicle-models 3.0.2, vehicle type 2 (BMW 320i)  
# ----------------------------------------------------------------------------------------------
class _Longitudinal:
    a_max = 11.5
    v_switch = 7.319
    v_min = -13.9
    v_max = 45.8


class _Steering:
    min = -1.066
    max = 1.066
    v_min = -0.4
    v_max = 0.4


class VehicleParametersBMW320i:
    l = 4.508
    w = 1.610
    a = 1.1561957064
    b = 1.4227170936
    longitudinal = _Longitudinal()
    steering = _Steering()


def vehicle_parameters(vehicle_type_id=2):
    if int(vehicle_type_id) != 2:
        raise NotImplementedError("only vehicle type 2 (BMW 320i) is restated")
    return VehicleParametersBMW320i()
