"""TEST INFRASTRUCTURE ONLY -- minimal CommonRoad XML (2018b / 2020a) reader for the reference's four
bundled scenarios (/root/reference/example_scenarios/*.xml).  commonroad-io is not
installed here, and the XML files themselves must not be copied into this repo: this module turns a
scenario into the plain-array "scenario dict" of ``commonroad_rp_b200.utility.synthetic`` (plus the
planning problem), which oracle/make_golden.py stores as a compact fixture.

Road boundary: ``commonroad_rp_b200.collision.road_boundary_segments`` (lanelet borders whose far side is off-road)
as thin boxes (width 0.1 m) or as a band of triangles (parity unpinned: the reference's default is commonroad_dc's
triangulation of the whole off-road area).
"""
import math
import xml.etree.ElementTree as ET

import numpy as np


def _points(node):
    return np.array([[float(p.find("x").text), float(p.find("y").text)] for p in node.findall("point")])


def _exact(node, tag, default=None):
    el = node.find(tag)
    if el is None:
        return default
    ex = el.find("exact")
    if ex is not None:
        return float(ex.text)
    lo, hi = el.find("intervalStart"), el.find("intervalEnd")
    if lo is not None and hi is not None:
        return 0.5 * (float(lo.text) + float(hi.text))
    return default


def _state(node):
    pos = node.find("position").find("point")
    return {"x": float(pos.find("x").text), "y": float(pos.find("y").text),
            "orientation": _exact(node, "orientation", 0.0), "time": int(_exact(node, "time", 0)),
            "velocity": _exact(node, "velocity", 0.0), "yaw_rate": _exact(node, "yawRate", 0.0),
            "acceleration": _exact(node, "acceleration", 0.0)}


def _rect(node):
    r = node.find("shape").find("rectangle")
    return float(r.find("length").text), float(r.find("width").text)


def load(path, boundary_method="obb_rectangles"):
    root = ET.parse(path).getroot()
    lanelets = {}
    for ll in root.findall("lanelet"):
        lid = int(ll.get("id"))
        lanelets[lid] = {
            "left": _points(ll.find("leftBound")), "right": _points(ll.find("rightBound")),
            "adj_left": None if ll.find("adjacentLeft") is None else int(ll.find("adjacentLeft").get("ref")),
            "adj_right": None if ll.find("adjacentRight") is None else int(ll.find("adjacentRight").get("ref")),
            "successor": [int(s.get("ref")) for s in ll.findall("successor")],
        }
    statics, dyn_t0, dyn_states, dyn_lw = [], [], [], []
    for ob in list(root.findall("obstacle")) + list(root.findall("staticObstacle")) + list(root.findall("dynamicObstacle")):
        role = ob.find("role").text if ob.find("role") is not None else (
            "static" if ob.tag == "staticObstacle" else "dynamic")
        length, width = _rect(ob)
        init = _state(ob.find("initialState"))
        traj = ob.find("trajectory")
        if role == "static" or traj is None:
            statics.append((init["x"], init["y"], init["orientation"], length, width))
        else:
            rows = [(init["x"], init["y"], init["orientation"])]
            expect = init["time"] + 1
            for st in traj.findall("state"):
                s = _state(st)
                assert s["time"] == expect, "non-consecutive obstacle trajectory"
                rows.append((s["x"], s["y"], s["orientation"]))
                expect += 1
            dyn_t0.append(init["time"])
            dyn_states.append(np.array(rows))
            dyn_lw.append((length, width))
    # road boundary: the product's own construction from the lanelet borders (input data for BOTH sides of the
    # parity tests, not arithmetic under test): thin boxes along the true boundary segments, or a triangle band
    from types import SimpleNamespace
    from commonroad_rp_b200 import collision as rpc
    lls = [SimpleNamespace(left_vertices=ll["left"], right_vertices=ll["right"], adj_left=ll["adj_left"],
                           adj_right=ll["adj_right"]) for ll in lanelets.values()]
    segs = rpc.road_boundary_segments(lls)
    boundary, tris = [], np.zeros((0, 6))
    if boundary_method == "triangulation":
        tris = rpc.boundary_band_triangles(lls, segs, 1.0)
    else:
        for p, q in segs:
            seg = q - p
            mid = 0.5 * (p + q)
            boundary.append((mid[0], mid[1], math.atan2(seg[1], seg[0]), 0.5 * float(np.hypot(seg[0], seg[1])), 0.05))
    pp = root.find("planningProblem")
    init = _state(pp.find("initialState"))
    goal = pp.find("goalState")
    gpos = goal.find("position")
    goal_info = {"time": (int(float(goal.find("time").find("intervalStart").text)),
                          int(float(goal.find("time").find("intervalEnd").text)))}
    if gpos is not None and gpos.find("rectangle") is not None:
        r = gpos.find("rectangle")
        c = r.find("center")
        goal_info["rectangle"] = (float(c.find("x").text), float(c.find("y").text),
                                  float(r.find("orientation").text) if r.find("orientation") is not None else 0.0,
                                  float(r.find("length").text), float(r.find("width").text))
    elif gpos is not None and gpos.find("lanelet") is not None:
        goal_info["lanelet"] = int(gpos.find("lanelet").get("ref"))
    gv = goal.find("velocity")
    if gv is not None:
        goal_info["velocity"] = (float(gv.find("intervalStart").text), float(gv.find("intervalEnd").text))
    return {
        "lanelets": lanelets,
        "scn": {"static_boxes": np.array(statics, dtype=np.float64).reshape(-1, 5),
                "dyn_t0": np.array(dyn_t0, dtype=np.int64), "dyn_states": dyn_states,
                "dyn_lw": np.array(dyn_lw, dtype=np.float64).reshape(-1, 2),
                "boundary_boxes": np.array(boundary, dtype=np.float64).reshape(-1, 5),
                "boundary_tris": np.asarray(tris, dtype=np.float64).reshape(-1, 6)},
        "initial_state": init, "goal": goal_info,
    }


def route_centerline(lanelets, ids):
    """Concatenated centre lines (mean of left / right bound) of the route's lanelets."""
    pts = []
    for lid in ids:
        ll = lanelets[lid]
        c = 0.5 * (ll["left"] + ll["right"])
        if pts and np.allclose(pts[-1], c[0]):
            c = c[1:]
        pts.extend(list(c))
    return np.array(pts)


def desired_velocity(sc):
    """retrieve_desired_velocity_from_pp (utility/general.py:32-46)."""
    g = sc["goal"]
    if "velocity" in g:
        lo, hi = g["velocity"]
        return (lo + hi) / 2 if lo > 0 else hi / 2
    return sc["initial_state"]["velocity"]


class GoalStub:
    """goal.is_reached(state): inside the goal shape during the goal time interval."""

    def __init__(self, sc):
        self.g = sc["goal"]
        self.lanelets = sc["lanelets"]

    def is_reached(self, state):
        t0, t1 = self.g["time"]
        if not (t0 <= state.time_step <= t1):
            return False
        x, y = state.position
        if "rectangle" in self.g:
            cx, cy, th, l, w = self.g["rectangle"]
            dx, dy = x - cx, y - cy
            lx = dx * math.cos(th) + dy * math.sin(th)
            ly = -dx * math.sin(th) + dy * math.cos(th)
            return abs(lx) <= 0.5 * l and abs(ly) <= 0.5 * w
        if "lanelet" in self.g:
            ll = self.lanelets[self.g["lanelet"]]
            poly = np.vstack([ll["left"], ll["right"][::-1]])
            inside = False
            n = len(poly)
            for i in range(n):
                x1, y1 = poly[i]
                x2, y2 = poly[(i + 1) % n]
                if (y1 > y) != (y2 > y) and x < (x2 - x1) * (y - y1) / (y2 - y1) + x1:
                    inside = not inside
            return inside
        return True
