"""Obstacle containers with the surface of ``commonroad_dc.pycrcc`` that the reference planner uses
(``reactive_planner.py:28-31, :234-251, :1040-1042``): ``RectOBB, RectAABB, Triangle, ShapeGroup,
TimeVariantCollisionObject, CollisionChecker``.  They are plain host-side descriptions; the narrow
phase runs on the device (separating-axis kernels in csrc/rp_device.cuh).  ``CollisionChecker`` packs
itself into the arrays of ``rp_ctx_set_obstacles``; ``collide()`` answers single queries through
``rp_collide_poses`` -- there is no host narrow phase.

commonroad_dc is not installable here, so semantics follow SURVEY.md App. D#2 (parity unpinned):
closed shapes (touching counts as collision), time-variant objects collide only at common time
indices, static objects at every index.
"""
import math
from typing import List, Optional

import numpy as np


class RectOBB:
    """Oriented box: half extents r_x (along the heading) and r_y, heading, centre."""

    def __init__(self, r_x: float, r_y: float, orientation: float, cx: float, cy: float):
        self.r_x, self.r_y, self.orientation, self.cx, self.cy = float(r_x), float(r_y), float(orientation), float(cx), float(cy)

    def center(self):
        return np.array([self.cx, self.cy])

    def row(self):
        return (self.cx, self.cy, self.orientation, self.r_x, self.r_y)


class RectAABB(RectOBB):
    def __init__(self, r_x: float, r_y: float, cx: float, cy: float):
        super().__init__(r_x, r_y, 0.0, cx, cy)


class Triangle:
    def __init__(self, x1, y1, x2, y2, x3, y3):
        self.vertices = (float(x1), float(y1), float(x2), float(y2), float(x3), float(y3))

    def row(self):
        return self.vertices


class ShapeGroup:
    def __init__(self):
        self._shapes = []
        self._version = 0

    def add_shape(self, shape):
        self._shapes.append(shape)
        self._version += 1

    def unpack(self):
        return list(self._shapes)


class TimeVariantCollisionObject:
    """One shape per consecutive time index starting at ``time_start_idx``."""

    def __init__(self, time_start_idx: int):
        self._t0 = int(time_start_idx)
        self._shapes = []
        self._version = 0

    def append_obstacle(self, shape):
        self._shapes.append(shape)
        self._version += 1

    def time_start_idx(self) -> int:
        return self._t0

    def time_end_idx(self) -> int:
        return self._t0 + len(self._shapes) - 1

    def obstacle_at_time(self, time_idx: int):
        k = time_idx - self._t0
        return self._shapes[k] if 0 <= k < len(self._shapes) else None


def obb_sum_hull(a: RectOBB, b: RectOBB) -> RectOBB:
    """The OBB-sum hull of two consecutive boxes (continuous collision check, reference :240-241, :1049-1058): the
    tight box along ``a``'s axes that encloses both.  commonroad_dc is not installable here, so this follows
    SURVEY.md App. D#2 / oracle/third_party.py (parity unpinned); the device builds the ego's hulls with the same
    expressions (continuous_check_kernel)."""
    ca, sa = math.cos(a.orientation), math.sin(a.orientation)
    cb, sb = math.cos(b.orientation), math.sin(b.orientation)
    dx, dy = b.cx - a.cx, b.cy - a.cy
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


def trajectory_preprocess_obb_sum(tvo: "TimeVariantCollisionObject"):
    """``commonroad_dc.collision.trajectory_queries.trajectory_preprocess_obb_sum``: n boxes -> n - 1 hull boxes
    (steps k, k + 1) from the same start index; returns (object, error code) like the original."""
    out = TimeVariantCollisionObject(tvo.time_start_idx())
    shapes = [tvo.obstacle_at_time(k) for k in range(tvo.time_start_idx(), tvo.time_end_idx() + 1)]
    for k in range(len(shapes) - 1):
        if not isinstance(shapes[k], RectOBB) or not isinstance(shapes[k + 1], RectOBB):
            return None, -1
        out.append_obstacle(obb_sum_hull(shapes[k], shapes[k + 1]))
    return out, 0


class CollisionChecker:
    """Set of static shapes, shape groups and time-variant objects."""

    def __init__(self):
        self._objects = []
        self._engine = None
        self._engine_dims = None
        self._version = 0

    def add_collision_object(self, obj):
        self._objects.append(obj)
        self._version += 1

    def obstacles(self):
        return list(self._objects)

    @property
    def version(self) -> int:
        """Changes whenever the checker's content changes: objects added here, or shapes appended to a contained
        shape group / time-variant object after it was added (device tables are re-uploaded when it moves)."""
        return self._version + sum(getattr(o, "_version", 0) for o in self._objects)

    # ---- packing for rp_ctx_set_obstacles ----
    def device_arrays(self) -> dict:
        static, tris, dyn_t0, dyn_boxes = [], [], [], []

        def add_static(shape):
            if isinstance(shape, ShapeGroup):
                for s in shape.unpack():
                    add_static(s)
            elif isinstance(shape, Triangle):
                tris.append(shape.row())
            elif isinstance(shape, RectOBB):
                static.append(shape.row())
            else:
                raise TypeError("<CollisionChecker>: unsupported shape %r (only boxes and triangles)" % type(shape))

        for obj in self._objects:
            if isinstance(obj, TimeVariantCollisionObject):
                rows = []
                for k in range(obj.time_start_idx(), obj.time_end_idx() + 1):
                    sh = obj.obstacle_at_time(k)
                    if not isinstance(sh, RectOBB):
                        raise TypeError("<CollisionChecker>: time-variant obstacles must consist of boxes")
                    rows.append(sh.row())
                dyn_t0.append(obj.time_start_idx())
                dyn_boxes.append(np.asarray(rows, dtype=np.float64).reshape(-1, 5))
            else:
                add_static(obj)
        return {"static_obb": np.asarray(static, dtype=np.float64).reshape(-1, 5),
                "dyn_t0": np.asarray(dyn_t0, dtype=np.int32), "dyn_boxes": dyn_boxes,
                "tris": np.asarray(tris, dtype=np.float64).reshape(-1, 6)}

    def upload(self, engine, cell_size: float = 0.0):
        arr = self.device_arrays()
        engine.set_obstacles(arr["static_obb"], arr["dyn_t0"], arr["dyn_boxes"], arr["tris"], cell_size)

    # ---- pycrcc-style query ----
    def collide(self, obj) -> bool:
        """True if any object of the checker intersects ``obj`` (a box or a time-variant object of boxes).
        Runs on the device."""
        if isinstance(obj, TimeVariantCollisionObject):
            times = list(range(obj.time_start_idx(), obj.time_end_idx() + 1))
            boxes = [obj.obstacle_at_time(k) for k in times]
            static_query = False
        elif isinstance(obj, RectOBB):
            times, boxes, static_query = [0], [obj], True
        else:
            raise TypeError("<CollisionChecker.collide>: unsupported query object %r" % type(obj))
        if not boxes:
            return False
        if static_query and any(isinstance(o, TimeVariantCollisionObject) for o in self._objects):
            raise TypeError("<CollisionChecker.collide>: static query against time-variant obstacles")
        hl = max(b.r_x for b in boxes)
        hw = max(b.r_y for b in boxes)
        if any(b.r_x != hl or b.r_y != hw for b in boxes):
            return any(self.collide(_single(b, t)) for b, t in zip(boxes, times))
        eng = self._query_engine(hl, hw)
        pose = np.array([[b.cx, b.cy, b.orientation] for b in boxes], dtype=np.float64)
        return bool(eng.collide_poses(pose, np.asarray(times, dtype=np.int32), hl, hw).any())

    def _query_engine(self, hl, hw):
        from commonroad_rp_b200._device import current_device_and_stream
        from commonroad_rp_b200._lib import Engine
        state = (hl, hw, self.version)
        if self._engine is None:
            dev, stream = current_device_and_stream()
            self._engine = Engine(dev, stream)
        if self._engine_dims != state:
            # the broad-phase grid is inflated by the query box's circumradius
            self._engine.set_vehicle(2 * hl, 2 * hw, 0.0, 1.0, 1.0, 1.0, 0.5, 0.5)
            self.upload(self._engine)
            self._engine_dims = state
        return self._engine


def _single(box, t):
    tvo = TimeVariantCollisionObject(t)
    tvo.append_obstacle(box)
    return tvo


# ---- construction from CommonRoad scenario objects (duck-typed commonroad-io) -----------------------
def _shape_box(shape, position, orientation):
    """commonroad.geometry.shape.Rectangle (length, width[, center, orientation]) placed at a state."""
    length = getattr(shape, "length", None)
    width = getattr(shape, "width", None)
    if length is None or width is None:
        raise TypeError("<create_collision_object>: only rectangular obstacle shapes are supported")
    off = np.asarray(getattr(shape, "center", (0.0, 0.0)), dtype=np.float64)
    local_th = float(getattr(shape, "orientation", 0.0))
    c, s = math.cos(orientation), math.sin(orientation)
    cx = position[0] + c * off[0] - s * off[1]
    cy = position[1] + s * off[0] + c * off[1]
    return RectOBB(0.5 * length, 0.5 * width, orientation + local_th, cx, cy)


def create_collision_object(obstacle):
    """commonroad_dc ... pycrcc_collision_dispatch.create_collision_object for static / dynamic obstacles
    with rectangular shapes (reference call sites reactive_planner.py:236, :239)."""
    init = obstacle.initial_state
    prediction = getattr(obstacle, "prediction", None)
    if prediction is None:
        return _shape_box(obstacle.obstacle_shape, init.position, init.orientation)
    tvo = TimeVariantCollisionObject(int(init.time_step))
    tvo.append_obstacle(_shape_box(obstacle.obstacle_shape, init.position, init.orientation))
    states = prediction.trajectory.state_list
    expected = int(init.time_step) + 1
    for st in states:
        if int(st.time_step) != expected:
            raise ValueError("<create_collision_object>: prediction must cover consecutive time steps")
        tvo.append_obstacle(_shape_box(obstacle.obstacle_shape, st.position, st.orientation))
        expected += 1
    return tvo


def _points_in_polygon(pts, poly):
    """Even-odd rule for points pts[n][2] against one polygon poly[m][2] (vectorised ray casting)."""
    x, y = pts[:, 0][:, None], pts[:, 1][:, None]
    x1, y1 = poly[:, 0][None, :], poly[:, 1][None, :]
    x2, y2 = np.roll(poly[:, 0], -1)[None, :], np.roll(poly[:, 1], -1)[None, :]
    straddle = (y1 > y) != (y2 > y)
    with np.errstate(divide="ignore", invalid="ignore"):
        xi = (x2 - x1) * (y - y1) / (y2 - y1) + x1
    return (np.count_nonzero(straddle & (x < xi), axis=1) & 1).astype(bool)


def road_boundary_segments(lanelets, probe: float = 0.1):
    """The border segments of a lanelet network that separate drivable area from off-road area.

    ``lanelets``: objects with ``left_vertices`` / ``right_vertices`` ((n, 2) arrays, driving direction) and
    ``adj_left`` / ``adj_right`` (None = no laterally adjacent lanelet).  A border without a lateral neighbour is not
    yet a road boundary: at intersections, merges and forks such borders run through OTHER lanelets' drivable area.  A
    segment is kept iff the point ``probe`` metres beyond it -- on the side away from its own lanelet -- lies in no
    lanelet's polygon, i.e. iff the road really ends there (the criterion commonroad_dc's boundary construction
    implements by triangulating the complement of the union of all lanelet polygons; reference call site
    reactive_planner.py:247).  Returns an (m, 2, 2) array of segments with the off-road side to the LEFT of p -> q.
    """
    polys = [np.vstack([np.asarray(ll.left_vertices, dtype=np.float64),
                        np.asarray(ll.right_vertices, dtype=np.float64)[::-1]]) for ll in lanelets]
    segs = []
    for ll in lanelets:
        for side, adj, sign in (("left_vertices", "adj_left", 1.0), ("right_vertices", "adj_right", -1.0)):
            if getattr(ll, adj, None) is not None:
                continue
            pts = np.asarray(getattr(ll, side), dtype=np.float64)
            p, q = pts[:-1], pts[1:]
            if sign < 0:
                p, q = q, p                     # off-road side to the left of p -> q
            keep = np.hypot(*(q - p).T) > 0.0
            segs.append(np.stack([p[keep], q[keep]], axis=1))
    if not segs:
        return np.zeros((0, 2, 2))
    segs = np.concatenate(segs, axis=0)
    d = segs[:, 1] - segs[:, 0]
    normal = np.stack([-d[:, 1], d[:, 0]], axis=1) / np.hypot(d[:, 0], d[:, 1])[:, None]
    probes = 0.5 * (segs[:, 0] + segs[:, 1]) + probe * normal
    inside = np.zeros(len(segs), dtype=bool)
    for poly in polys:
        inside |= _points_in_polygon(probes, poly)
    return segs[~inside]


def create_road_boundary_obstacle(scenario, method: str = "obb_rectangles", width: float = 0.1, band: float = 1.0):
    """Road boundary of a CommonRoad scenario as a static shape group (reference call site reactive_planner.py:247,
    ``commonroad_dc.boundary.boundary.create_road_boundary_obstacle``).  Returns (None, ShapeGroup).

    commonroad_dc's default method triangulates the off-road area inside an enlarged bounding box with the ``triangle``
    package, which is not installable here (parity unpinned); both methods below are built from the
    true road boundary (``road_boundary_segments``: borders whose far side is off-road) and give the same collision
    verdict for every pose that touches the road's edge:

    ``obb_rectangles``   one thin box (``width``) per boundary segment -- commonroad_dc's method of the same name;
    ``triangulation``    two triangles per boundary segment covering a band of ``band`` metres on the off-road side
                         (triangles reaching into another lanelet's drivable area are dropped).
    """
    lanelets = list(scenario.lanelet_network.lanelets)
    segs = road_boundary_segments(lanelets)
    sg = ShapeGroup()
    if method == "obb_rectangles":
        for p, q in segs:
            seg = q - p
            mid = 0.5 * (p + q)
            sg.add_shape(RectOBB(0.5 * float(np.hypot(seg[0], seg[1])), 0.5 * width, math.atan2(seg[1], seg[0]), mid[0], mid[1]))
    elif method == "triangulation":
        for tri in boundary_band_triangles(lanelets, segs, band):
            sg.add_shape(Triangle(*tri))
    else:
        raise ValueError("<create_road_boundary_obstacle>: unknown method %r" % (method,))
    return None, sg


def boundary_band_triangles(lanelets, segs, band: float = 1.0):
    """(m, 6) triangles: per boundary segment p -> q (off-road side to the left) the quad p, q, q + band n, p + band n
    split along its diagonal; a triangle with a vertex or its centroid inside a lanelet polygon (the band of one road
    reaching into another road across a narrow gore) is dropped."""
    if len(segs) == 0:
        return np.zeros((0, 6))
    p, q = segs[:, 0], segs[:, 1]
    d = q - p
    n = np.stack([-d[:, 1], d[:, 0]], axis=1) / np.hypot(d[:, 0], d[:, 1])[:, None]
    a, b = p + band * n, q + band * n
    tris = np.concatenate([np.stack([p, q, b], axis=1), np.stack([p, b, a], axis=1)], axis=0)      # (2m, 3, 2)
    # probe points: the centroid and the two vertices off the segment, pulled slightly towards the centroid
    cen = tris.mean(axis=1)
    probes = np.concatenate([cen[:, None, :], cen[:, None, :] + 0.98 * (tris - cen[:, None, :])], axis=1)   # (2m, 4, 2)
    on_edge = np.zeros(probes.shape[:2], dtype=bool)
    on_edge[:, 1] = True                                  # vertex 0 is p (on the border itself)
    on_edge[:len(p), 2] = True                            # first family: vertex 1 is q
    bad = np.zeros(len(tris), dtype=bool)
    polys = [np.vstack([np.asarray(ll.left_vertices, dtype=np.float64),
                        np.asarray(ll.right_vertices, dtype=np.float64)[::-1]]) for ll in lanelets]
    flat = probes.reshape(-1, 2)
    for poly in polys:
        inside = _points_in_polygon(flat, poly).reshape(probes.shape[:2])
        bad |= (inside & ~on_edge).any(axis=1)
    return tris[~bad].reshape(-1, 6)


def checker_from_arrays(static_boxes=(), dyn_t0=(), dyn_states=(), dyn_lw=(), boundary_boxes=(), boundary_tris=()):
    """CollisionChecker from a scenario dict of ``utility.synthetic`` (static_boxes rows are
    cx, cy, theta, LENGTH, WIDTH; boundary_boxes rows carry half extents)."""
    cc = CollisionChecker()
    for cx, cy, th, l, w in np.asarray(static_boxes, dtype=np.float64).reshape(-1, 5):
        cc.add_collision_object(RectOBB(0.5 * l, 0.5 * w, th, cx, cy))
    for t0, st, lw in zip(dyn_t0, dyn_states, dyn_lw):
        tvo = TimeVariantCollisionObject(int(t0))
        for cx, cy, th in np.asarray(st, dtype=np.float64).reshape(-1, 3):
            tvo.append_obstacle(RectOBB(0.5 * lw[0], 0.5 * lw[1], th, cx, cy))
        cc.add_collision_object(tvo)
    sg = ShapeGroup()
    for cx, cy, th, hl, hw in np.asarray(boundary_boxes, dtype=np.float64).reshape(-1, 5):
        sg.add_shape(RectOBB(hl, hw, th, cx, cy))
    for tri in np.asarray(boundary_tris, dtype=np.float64).reshape(-1, 6):
        sg.add_shape(Triangle(*tri))
    cc.add_collision_object(sg)
    return cc
