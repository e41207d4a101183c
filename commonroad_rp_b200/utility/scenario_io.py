"""Scenario files -> the objects the planner turns into device tables ("on-disk format step").

The reference reads CommonRoad XML through commonroad-io (``utility/general.py:11-29``) and builds its collision checker
from the scenario's obstacles and lanelet network (``reactive_planner.py:218-256``).  commonroad-io is used here when it is
installed; otherwise this module reads the XML (format versions 2018b and 2020a: the four scenarios the reference bundles)
into light objects with the attribute names the planner consumes:

    scenario.static_obstacles / .dynamic_obstacles   obstacle_id, obstacle_shape (length, width), initial_state, prediction
    scenario.lanelet_network.lanelets                lanelet_id, left_vertices, right_vertices, center_vertices, adj_left,
                                                     adj_right, successor, predecessor
    planning_problem.initial_state                   position (vehicle centre), orientation, velocity, yaw_rate, slip_angle,
                                                     time_step, acceleration
    planning_problem.goal                            is_reached(state), state_list[0].velocity / .time_step intervals

``ReactivePlanner.reset(config)`` / ``set_collision_checker(scenario=...)`` turn them into the checker whose content
``rp_ctx_set_obstacles`` uploads (static boxes, per-time-index dynamic boxes, road-boundary primitives), and
``parallel.ScenarioBatch.add_scenario_from`` does the same for a batch of scenarios.
"""
import math
import xml.etree.ElementTree as ET
from types import SimpleNamespace
from typing import List, Optional

import numpy as np

from commonroad_rp_b200._compat import InitialState


class Interval:
    def __init__(self, start, end):
        self.start, self.end = start, end

    def contains(self, value) -> bool:
        return self.start <= value <= self.end

    def __repr__(self):
        return "Interval(%r, %r)" % (self.start, self.end)


class Rectangle:
    """commonroad.geometry.shape.Rectangle stand-in: length along the heading, width across, optional local offset"""

    def __init__(self, length, width, center=(0.0, 0.0), orientation=0.0):
        self.length, self.width = float(length), float(width)
        self.center = np.asarray(center, dtype=np.float64)
        self.orientation = float(orientation)

    def contains_point(self, point, position=(0.0, 0.0), heading=0.0) -> bool:
        th = heading + self.orientation
        c, s = math.cos(heading), math.sin(heading)
        cx = position[0] + c * self.center[0] - s * self.center[1]
        cy = position[1] + s * self.center[0] + c * self.center[1]
        dx, dy = point[0] - cx, point[1] - cy
        lx = dx * math.cos(th) + dy * math.sin(th)
        ly = -dx * math.sin(th) + dy * math.cos(th)
        return abs(lx) <= 0.5 * self.length and abs(ly) <= 0.5 * self.width


class Lanelet:
    def __init__(self, lanelet_id, left, right, adj_left, adj_right, successor, predecessor):
        self.lanelet_id = lanelet_id
        self.left_vertices = left
        self.right_vertices = right
        self.center_vertices = 0.5 * (left + right)
        self.adj_left, self.adj_right = adj_left, adj_right
        self.successor, self.predecessor = successor, predecessor

    @property
    def polygon(self) -> np.ndarray:
        return np.vstack([self.left_vertices, self.right_vertices[::-1]])

    def contains_point(self, point) -> bool:
        from commonroad_rp_b200.collision import _points_in_polygon
        return bool(_points_in_polygon(np.asarray(point, dtype=np.float64).reshape(1, 2), self.polygon)[0])


class LaneletNetwork:
    def __init__(self, lanelets: List[Lanelet]):
        self.lanelets = lanelets
        self._by_id = {ll.lanelet_id: ll for ll in lanelets}

    def find_lanelet_by_id(self, lanelet_id) -> Lanelet:
        return self._by_id[lanelet_id]

    def find_lanelet_by_position(self, point) -> List[int]:
        return [ll.lanelet_id for ll in self.lanelets if ll.contains_point(point)]


class Goal:
    """goal region: a position (rectangle or lanelet) during a time-step interval, optional velocity interval"""

    def __init__(self, time_step: Interval, shape=None, lanelet: Optional[Lanelet] = None, velocity: Optional[Interval] = None):
        st = SimpleNamespace(time_step=time_step)
        if velocity is not None:
            st.velocity = velocity
        self.state_list = [st]
        self._shape, self._lanelet = shape, lanelet
        self.lanelets_of_goal_position = {0: [lanelet.lanelet_id]} if lanelet is not None else None

    def is_reached(self, state) -> bool:
        if not self.state_list[0].time_step.contains(state.time_step):
            return False
        if self._shape is not None:
            return self._shape.contains_point(state.position)
        if self._lanelet is not None:
            return self._lanelet.contains_point(state.position)
        return True


def _points(node) -> np.ndarray:
    return np.array([[float(p.find("x").text), float(p.find("y").text)] for p in node.findall("point")], dtype=np.float64)


def _scalar(node, tag, default=None):
    """<tag><exact>v</exact></tag> or the mid-point of an interval"""
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


def _interval(node, tag, cast=float) -> Optional[Interval]:
    el = node.find(tag)
    if el is None:
        return None
    ex = el.find("exact")
    if ex is not None:
        return Interval(cast(float(ex.text)), cast(float(ex.text)))
    return Interval(cast(float(el.find("intervalStart").text)), cast(float(el.find("intervalEnd").text)))


def _state(node):
    pos = node.find("position").find("point")
    return SimpleNamespace(position=np.array([float(pos.find("x").text), float(pos.find("y").text)]),
                           orientation=_scalar(node, "orientation", 0.0), time_step=int(_scalar(node, "time", 0)),
                           velocity=_scalar(node, "velocity", 0.0))


def _rectangle(node) -> Rectangle:
    r = node.find("shape").find("rectangle")
    if r is None:
        raise ValueError("<read_commonroad_xml>: only rectangular obstacle shapes are supported")
    center = r.find("center")
    off = (float(center.find("x").text), float(center.find("y").text)) if center is not None else (0.0, 0.0)
    ori = float(r.find("orientation").text) if r.find("orientation") is not None else 0.0
    return Rectangle(float(r.find("length").text), float(r.find("width").text), off, ori)


def read_commonroad_xml(path, idx_planning_problem: Optional[int] = None):
    """(scenario, planning_problem) of a CommonRoad 2018b / 2020a file."""
    root = ET.parse(path).getroot()
    dt = float(root.get("timeStepSize", 0.1))
    ref = lambda ll, tag: None if ll.find(tag) is None else int(ll.find(tag).get("ref"))
    lanelets = [Lanelet(int(ll.get("id")), _points(ll.find("leftBound")), _points(ll.find("rightBound")),
                        ref(ll, "adjacentLeft"), ref(ll, "adjacentRight"),
                        [int(s.get("ref")) for s in ll.findall("successor")],
                        [int(s.get("ref")) for s in ll.findall("predecessor")]) for ll in root.findall("lanelet")]
    network = LaneletNetwork(lanelets)
    statics, dynamics = [], []
    for ob in list(root.findall("obstacle")) + list(root.findall("staticObstacle")) + list(root.findall("dynamicObstacle")):
        role = ob.find("role").text if ob.find("role") is not None else ("static" if ob.tag == "staticObstacle" else "dynamic")
        init = _state(ob.find("initialState"))
        traj = ob.find("trajectory")
        obstacle = SimpleNamespace(obstacle_id=int(ob.get("id")), obstacle_shape=_rectangle(ob), initial_state=init, prediction=None)
        if role == "static" or traj is None:
            statics.append(obstacle)
        else:
            states = [_state(st) for st in traj.findall("state")]
            obstacle.prediction = SimpleNamespace(trajectory=SimpleNamespace(initial_time_step=init.time_step + 1, state_list=states))
            dynamics.append(obstacle)
    scenario = SimpleNamespace(scenario_id=root.get("benchmarkID", ""), dt=dt, lanelet_network=network,
                               static_obstacles=statics, dynamic_obstacles=dynamics)
    problems = root.findall("planningProblem")
    if not problems:
        return scenario, None
    pp = problems[0] if idx_planning_problem is None else next(
        (q for q in problems if int(q.get("id")) == idx_planning_problem), None)
    if pp is None:
        raise KeyError(f"<ReactivePlannerConfiguration.update()>:Planning Problem with ID: {idx_planning_problem} does not exist!")
    ini = pp.find("initialState")
    st = _state(ini)
    acc = ini.find("acceleration")
    initial = InitialState(time_step=st.time_step, position=st.position, orientation=st.orientation, velocity=st.velocity,
                           steering_angle=None, yaw_rate=_scalar(ini, "yawRate", 0.0), slip_angle=_scalar(ini, "slipAngle", 0.0),
                           acceleration=_scalar(ini, "acceleration", 0.0) if acc is not None else None)
    g = pp.find("goalState")
    shape = lanelet = None
    gpos = g.find("position")
    if gpos is not None and gpos.find("rectangle") is not None:
        r = gpos.find("rectangle")
        c = r.find("center")
        shape = Rectangle(float(r.find("length").text), float(r.find("width").text),
                          (float(c.find("x").text), float(c.find("y").text)),
                          float(r.find("orientation").text) if r.find("orientation") is not None else 0.0)
    elif gpos is not None and gpos.find("lanelet") is not None:
        lanelet = network.find_lanelet_by_id(int(gpos.find("lanelet").get("ref")))
    goal = Goal(_interval(g, "time", int), shape, lanelet, _interval(g, "velocity"))
    problem = SimpleNamespace(planning_problem_id=int(pp.get("id")), initial_state=initial, goal=goal)
    return scenario, problem


def find_route(scenario, planning_problem) -> List[int]:
    """Lanelet ids from the initial position to the goal by breadth-first search over the successor graph (a stand-in for
    commonroad_route_planner, run_planner.py:43 -- enough for scenarios without lane changes)."""
    net = scenario.lanelet_network
    starts = net.find_lanelet_by_position(planning_problem.initial_state.position)
    goal = planning_problem.goal
    if goal.lanelets_of_goal_position:
        targets = set(goal.lanelets_of_goal_position[0])
    elif goal._shape is not None:
        centre = np.array([goal._shape.center[0], goal._shape.center[1]])
        targets = set(net.find_lanelet_by_position(centre))
    else:
        targets = set()
    best = None
    for s in starts:
        prev, queue = {s: None}, [s]
        while queue:
            cur = queue.pop(0)
            if cur in targets or (not targets and not net.find_lanelet_by_id(cur).successor):
                path = []
                while cur is not None:
                    path.append(cur)
                    cur = prev[cur]
                path.reverse()
                if best is None or len(path) < len(best):
                    best = path
                break
            for nxt in net.find_lanelet_by_id(cur).successor:
                if nxt not in prev:
                    prev[nxt] = cur
                    queue.append(nxt)
    if best is None:
        raise ValueError("<find_route>: no lanelet route from the initial state to the goal")
    return best


def route_reference_path(scenario, lanelet_ids, extend_back: float = 0.0) -> np.ndarray:
    """Concatenated centre lines of the route's lanelets, optionally extended backwards along the first segment (the
    planner's state refers to the REAR axle, which may lie before the first centre point)."""
    pts = []
    for lid in lanelet_ids:
        c = scenario.lanelet_network.find_lanelet_by_id(lid).center_vertices
        if pts and np.allclose(pts[-1], c[0]):
            c = c[1:]
        pts.extend(list(c))
    path = np.array(pts)
    if extend_back > 0.0:
        u = (path[0] - path[1]) / np.hypot(*(path[0] - path[1]))
        path = np.vstack([np.array([path[0] + u * k for k in range(int(extend_back), 0, -1)]), path])
    return path
