"""TEST INFRASTRUCTURE ONLY -- works only where /root/reference exists (this container).

Makes the reference's OWN, UNMODIFIED ``commonroad_rp`` package importable although none of
its pip dependencies are installed: a ``sys.meta_path`` finder that
auto-stubs the absent package roots, plus behavioural shims for the handful of names whose
behaviour matters on the hot path.  Third-party arithmetic comes from
``oracle/third_party.py`` (parity unpinned, see there).  Nothing from /root/reference is
copied: the modules are executed from where they lie.

Used by ``oracle/make_golden.py`` (fixture generation) and by the ``not gpu`` tests that
validate the portable restatement ``oracle/rp_oracle.py`` against the real reference code.
"""
import dataclasses
import functools
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np

from . import third_party as tp

REFERENCE_ROOT = os.environ.get("RP_REFERENCE_ROOT", "/root/reference")

_STUB_ROOTS = ("commonroad", "commonroad_dc", "commonroad_route_planner", "vehiclemodels",
               "omegaconf", "methodtools", "matplotlib", "imageio", "commonroad_reach")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "commonroad_rp"))


class _Dummy:
    """Accept-anything placeholder for classes only used in type hints / cold code."""

    def __init__(self, *a, **k):
        self.__dict__.update(k)

    def __call__(self, *a, **k):
        return _Dummy()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Dummy,), {})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            if fullname == "commonroad_reach" or fullname.startswith("commonroad_reach."):
                return None  # sampling.py probes it inside try/except ImportError
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        _apply_behaviour(module)


# ---- behavioural shims -------------------------------------------------------------------------
def _lru_cache(maxsize=128):
    """methodtools.lru_cache stacked on @classmethod (polynomial_trajectory.py:292-293, :341-342)."""
    def deco(fn):
        raw = fn.__func__ if isinstance(fn, (classmethod, staticmethod)) else fn
        cached = functools.lru_cache(maxsize)(lambda *a: raw(None, *a))
        return staticmethod(cached)
    return deco


@dataclasses.dataclass(eq=False)
class KSState:
    time_step: object = None
    position: object = None
    steering_angle: object = None
    velocity: object = None
    orientation: object = None

    def translate_rotate(self, translation, angle):
        new = dataclasses.replace(self)
        c, s = np.cos(angle), np.sin(angle)
        p = np.asarray(self.position, dtype=np.float64) + np.asarray(translation, dtype=np.float64)
        new.position = np.array([c * p[0] - s * p[1], s * p[0] + c * p[1]])
        new.orientation = self.orientation + angle
        return new

    def convert_state_to_state(self, other):
        for f in dataclasses.fields(other):
            if hasattr(self, f.name):
                setattr(other, f.name, getattr(self, f.name))
        return other


class InitialState(_Dummy):
    pass


class CustomState:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Trajectory:
    def __init__(self, initial_time_step, state_list):
        self.initial_time_step = initial_time_step
        self.state_list = state_list


class _VehicleParameterMapping:
    @staticmethod
    def from_vehicle_type(vehicle_type):
        vid = getattr(vehicle_type, "value", vehicle_type)
        return tp.vehicle_parameters(vid)


class _VehicleType:
    def __init__(self, value):
        self.value = int(value)


def _create_collision_object(obj):
    """commonroad_dc ... pycrcc_collision_dispatch.create_collision_object (reactive_planner.py:236,239):
    our scenario stubs already carry third_party shapes."""
    return obj.to_collision_object()


def _create_road_boundary_obstacle(scenario, *a, **k):
    """commonroad_dc.boundary.boundary.create_road_boundary_obstacle (reactive_planner.py:247)."""
    return None, scenario.road_boundary_shape_group()


def _trajectory_preprocess_obb_sum(tvo):
    """commonroad_dc ... trajectory_preprocess_obb_sum (reactive_planner.py:241, :1053) -> oracle/third_party.py"""
    return tp.trajectory_preprocess_obb_sum(tvo)


_BEHAVIOUR = {
    "methodtools": {"lru_cache": _lru_cache},
    "commonroad.common.validity": {
        "is_real_number": lambda x: isinstance(x, (int, float, np.integer, np.floating)),
        "is_natural_number": lambda x: isinstance(x, (int, np.integer)) and x >= 0,
        "is_positive": lambda x: x > 0,
        "is_real_number_vector": lambda x, length=None: True,
    },
    "commonroad.common.util": {"make_valid_orientation": tp.make_valid_orientation},
    "commonroad.scenario.state": {"KSState": KSState, "InitialState": InitialState,
                                  "CustomState": CustomState, "FloatExactOrInterval": object},
    "commonroad.scenario.trajectory": {"Trajectory": Trajectory},
    "commonroad_dc.feasibility.vehicle_dynamics": {"VehicleParameterMapping": _VehicleParameterMapping},
    "commonroad.common.solution": {"VehicleType": _VehicleType},
    "commonroad_dc.pycrcc": {"RectOBB": tp.RectOBB, "RectAABB": tp.RectAABB, "Triangle": tp.Triangle,
                             "ShapeGroup": tp.ShapeGroup,
                             "TimeVariantCollisionObject": tp.TimeVariantCollisionObject,
                             "CollisionChecker": tp.CollisionChecker},
    "commonroad_dc.pycrccosy": {"CurvilinearCoordinateSystem": tp.CurvilinearCoordinateSystem},
    "commonroad_dc.geometry.util": {
        "compute_pathlength_from_polyline": tp.compute_pathlength_from_polyline,
        "compute_curvature_from_polyline": tp.compute_curvature_from_polyline,
        "compute_orientation_from_polyline": tp.compute_orientation_from_polyline,
        "resample_polyline": tp.resample_polyline,
        "chaikins_corner_cutting": tp.chaikins_corner_cutting,
    },
    "commonroad_dc.boundary.boundary": {"create_road_boundary_obstacle": _create_road_boundary_obstacle},
    "commonroad_dc.collision.collision_detection.pycrcc_collision_dispatch":
        {"create_collision_object": _create_collision_object},
    "commonroad_dc.collision.trajectory_queries.trajectory_queries":
        {"trajectory_preprocess_obb_sum": _trajectory_preprocess_obb_sum},
}


def _apply_behaviour(module):
    for k, v in _BEHAVIOUR.get(module.__name__, {}).items():
        setattr(module, k, v)


_installed = False


def install():
    """Install the stub finder and put the reference on sys.path.  Idempotent."""
    global _installed
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if not _installed:
        sys.meta_path.append(_StubFinder())
        if REFERENCE_ROOT not in sys.path:
            sys.path.append(REFERENCE_ROOT)
        _installed = True
    return importlib.import_module("commonroad_rp.reactive_planner")


# ---- scenario stubs the reference planner consumes ----------------------------------------------
class StaticObstacleStub:
    def __init__(self, cx, cy, theta, length, width):
        self.box = (float(cx), float(cy), float(theta), float(length), float(width))

    def to_collision_object(self):
        cx, cy, th, l, w = self.box
        return tp.RectOBB(0.5 * l, 0.5 * w, th, cx, cy)


class DynamicObstacleStub:
    """states[k] = (cx, cy, theta) at time index t0 + k (k = 0 is the initial state)."""

    def __init__(self, t0, states, length, width):
        self.t0 = int(t0)
        self.states = np.asarray(states, dtype=np.float64).reshape(-1, 3)
        self.length = float(length)
        self.width = float(width)

    def to_collision_object(self):
        tvo = tp.TimeVariantCollisionObject(self.t0)
        for cx, cy, th in self.states:
            tvo.append_obstacle(tp.RectOBB(0.5 * self.length, 0.5 * self.width, th, cx, cy))
        return tvo


class ScenarioStub:
    """config.scenario stand-in (reactive_planner.py:235-248)."""

    def __init__(self, static_obstacles=(), dynamic_obstacles=(), boundary_boxes=(), boundary_triangles=()):
        self.static_obstacles = list(static_obstacles)
        self.dynamic_obstacles = list(dynamic_obstacles)
        self.boundary_boxes = np.asarray(boundary_boxes, dtype=np.float64).reshape(-1, 5)
        self.boundary_triangles = np.asarray(boundary_triangles, dtype=np.float64).reshape(-1, 6)

    def road_boundary_shape_group(self):
        sg = tp.ShapeGroup()
        for cx, cy, th, hl, hw in self.boundary_boxes:   # (cx, cy, theta, half_len, half_wid)
            sg.add_shape(tp.RectOBB(hl, hw, th, cx, cy))
        for t in self.boundary_triangles:
            sg.add_shape(tp.Triangle(*t))
        return sg
