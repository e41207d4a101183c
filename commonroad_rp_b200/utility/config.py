"""Configuration tree of the planner.  Class names, field names, defaults and the YAML layout are those
of the reference's ``commonroad_rp/utility/config.py`` (:107-290) so that existing configuration files
and user code keep working; the classes themselves are generated from the specification tables below.

Not identical on purpose (none of it changes the API):
* YAML is parsed with OmegaConf when importable, else PyYAML (OmegaConf is not installed in this image);
* vehicle parameters come from ``vehiclemodels`` / ``commonroad_dc`` when importable, else from the
  built-in commonroad-vehicle-models 3.0.2 table (type 2 = BMW 320i is the reference default, :198);
* ``debug.multiproc`` and ``debug.num_workers`` are accepted and ignored: candidates are GPU threads,
  there is nothing to fork.
"""
import dataclasses
import inspect
import os.path
import pathlib
import warnings
from typing import Any, List, Optional

import numpy as np


class _Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


# l, w, a, b, a_max, v_switch, steering min / max, steering-rate min / max
_VEHICLES = {
    1: (4.298, 1.674, 0.88392, 1.50876, 11.5, 4.755, -0.910, 0.910, -0.4, 0.4),          # Ford Escort
    2: (4.508, 1.610, 1.1561957064, 1.4227170936, 11.5, 7.319, -1.066, 1.066, -0.4, 0.4),  # BMW 320i
    3: (4.569, 1.844, 1.08966, 1.35634, 11.5, 7.824, -1.023, 1.023, -0.4, 0.4),           # VW Vanagon
}


def vehicle_parameters_from_type(vehicle_type_id: int):
    """What ``VehicleParameterMapping.from_vehicle_type(VehicleType(id))`` returns (reference :200)."""
    try:
        from commonroad.common.solution import VehicleType
        from commonroad_dc.feasibility.vehicle_dynamics import VehicleParameterMapping
        return VehicleParameterMapping.from_vehicle_type(VehicleType(vehicle_type_id))
    except Exception:
        l, w, a, b, a_max, v_switch, s_lo, s_hi, sv_lo, sv_hi = _VEHICLES[int(vehicle_type_id)]
        return _Bag(l=l, w=w, a=a, b=b, longitudinal=_Bag(a_max=a_max, v_switch=v_switch),
                    steering=_Bag(min=s_lo, max=s_hi, v_min=sv_lo, v_max=sv_hi))


def _read_yaml(path) -> dict:
    try:
        from omegaconf import OmegaConf
        tree = OmegaConf.to_object(OmegaConf.load(path))
        if isinstance(tree, dict):
            return tree
    except Exception:       # absent (or a test stub of it): plain YAML is equivalent for these files
        pass
    import yaml
    with open(path) as fh:
        return yaml.safe_load(fh) or {}


def _from_dict(tree: dict, cls):
    """Nested dictionaries -> nested configuration objects; unknown keys are ignored like in the reference."""
    picked = {}
    for f in dataclasses.fields(cls):
        if f.init and f.name in tree:
            nested = inspect.isclass(f.type) and issubclass(f.type, BaseConfiguration)
            picked[f.name] = _from_dict(tree[f.name], f.type) if nested else tree[f.name]
    return cls(**picked)


class BaseConfiguration:
    """Item access (``cfg["planning"]["dt"]``) and the YAML loader shared by all configuration classes."""

    def __getitem__(self, item: str) -> Any:
        if not hasattr(self, item):
            raise KeyError(f"{item} is not a parameter of {self.__class__.__name__}")
        return getattr(self, item)

    def __setitem__(self, key: str, value: Any):
        try:
            setattr(self, key, value)
        except AttributeError as e:
            raise KeyError(f"{key} is not a parameter of {self.__class__.__name__}") from e

    @classmethod
    def load(cls, file_path, scenario_name: Optional[str] = None, validate_types: bool = True):
        """Configuration from a .yaml file; ``scenario_name`` fills ``general.path_scenario`` (reference :84-104)."""
        file_path = pathlib.Path(file_path)
        assert file_path.suffix == ".yaml", f"File type {file_path.suffix} is unsupported! Please use .yaml!"
        cfg = _from_dict(_read_yaml(file_path), cls)
        if scenario_name:
            cfg.general.set_path_scenario(scenario_name)
        return cfg


def _make(name: str, doc: str, spec, namespace=None):
    fields = []
    for fname, ftype, default in spec:
        if isinstance(default, (list, dict)):
            fields.append((fname, ftype, dataclasses.field(default_factory=lambda d=default: type(d)(d))))
        else:
            fields.append((fname, ftype, dataclasses.field(default=default)))
    cls = dataclasses.make_dataclass(name, fields, bases=(BaseConfiguration,), namespace=namespace or {})
    cls.__doc__ = doc
    cls.__module__ = __name__
    return cls


_DT, _STEPS = 0.1, 60

PlanningConfiguration = _make("PlanningConfiguration", "Planning parameters for reactive planner.", [
    ("dt", float, _DT),                                   # planner time step [s]
    ("time_steps_computation", int, _STEPS),              # horizon = dt * time_steps_computation
    ("planning_horizon", float, _DT * _STEPS),
    ("replanning_frequency", int, 3),                     # re-plan every n time steps
    ("continuous_collision_check", bool, False),
    ("factor", int, 1),                                   # planner step / scenario step for collision time indices
    ("low_vel_mode_threshold", float, 4.0),               # [m/s] below: lateral motion sampled over arc length
    ("constraints_to_check", List[str], ["velocity", "acceleration", "kappa", "kappa_dot", "yaw_rate"]),
    ("standstill_lookahead", int, 10),
])

SamplingConfiguration = _make("SamplingConfiguration", "Sampling parameters for reactive planner.", [
    ("sampling_method", int, 1),                          # 1 fixed intervals, 2 corridor sampling (CommonRoad-Reach)
    ("longitudinal_mode", str, "velocity_keeping"),       # or "stopping"
    ("num_sampling_levels", int, 4),
    ("t_min", float, 0.4),
    ("v_min", float, 0), ("v_max", float, 0),
    ("s_min", float, -1), ("s_max", float, 1),
    ("d_min", float, -3), ("d_max", float, 3),
])

DebugConfiguration = _make("DebugConfiguration", "Parameters specifying debug-related information.", [
    ("save_plots", bool, False), ("save_config", bool, False), ("show_plots", bool, False),
    ("draw_ref_path", bool, True), ("draw_planning_problem", bool, True), ("draw_icons", bool, False),
    ("draw_traj_set", bool, False),
    ("logging_level", str, "INFO"),
    ("multiproc", bool, True), ("num_workers", int, 6),   # accepted, ignored (see module docstring)
])


def _vehicle_post_init(self):
    vp = self.vehicle_parameters or vehicle_parameters_from_type(self.id_type_vehicle)
    self.vehicle_parameters = vp
    derived = {"length": vp.l, "width": vp.w, "wb_front_axle": vp.a, "wb_rear_axle": vp.b,
               "a_max": vp.longitudinal.a_max, "v_switch": vp.longitudinal.v_switch,
               "delta_min": vp.steering.min, "delta_max": vp.steering.max,
               "v_delta_min": vp.steering.v_min, "v_delta_max": vp.steering.v_max, "wheelbase": vp.a + vp.b}
    for key, value in derived.items():
        if getattr(self, key) is None:        # explicit values (e.g. from YAML) win
            setattr(self, key, value)
    self.kappa_max = np.tan(self.delta_max) / self.wheelbase


VehicleConfiguration = _make("VehicleConfiguration", "Class to store vehicle configurations", [
    ("id_type_vehicle", int, 2), ("vehicle_parameters", Any, None),
    ("length", float, None), ("width", float, None),
    ("wb_front_axle", float, None), ("wb_rear_axle", float, None),
    ("a_max", float, None), ("v_switch", float, None),
    ("delta_min", float, None), ("delta_max", float, None),
    ("v_delta_min", float, None), ("v_delta_max", float, None),
    ("wheelbase", float, None),
], namespace={"__post_init__": _vehicle_post_init})


def _set_path_scenario(self, scenario_name: str):
    self.path_scenario = os.path.join(self.path_scenarios, scenario_name)


GeneralConfiguration = _make("GeneralConfiguration", "General parameters for evaluations.", [
    ("path_scenarios", str, "example_scenarios/"), ("path_output", str, "output/"),
    ("path_logs", str, "output/logs/"), ("path_pickles", str, "output/pickles/"),
    ("path_scenario", Optional[str], None), ("name_scenario", Optional[str], None),
], namespace={"set_path_scenario": _set_path_scenario})


def _root_post_init(self):
    self.scenario = None
    self.planning_problem = None
    self.planning_problem_set = None


def _root_update(self, scenario=None, planning_problem=None, idx_planning_problem: Optional[int] = None):
    """Attach scenario / planning problem; with neither given they are read from ``general.path_scenario``
    (needs commonroad-io) -- reference :265-290."""
    self.scenario, self.planning_problem = scenario, planning_problem
    if scenario is None and planning_problem is None:
        try:
            from commonroad_rp_b200.utility.general import load_scenario_and_planning_problem
            self.scenario, self.planning_problem, self.planning_problem_set = \
                load_scenario_and_planning_problem(self.general.path_scenario, idx_planning_problem)
        except FileNotFoundError:
            warnings.warn(f"<ReactivePlannerConfiguration.update()>: No scenario .xml file found at "
                          f"path_scenario = {self.general.path_scenario}")
    assert self.scenario is not None, "<Configuration.update()>: no scenario has been specified"


def _sub(cls):
    return dataclasses.field(default_factory=cls)


ReactivePlannerConfiguration = dataclasses.make_dataclass(
    "ReactivePlannerConfiguration",
    [("vehicle", VehicleConfiguration, _sub(VehicleConfiguration)),
     ("planning", PlanningConfiguration, _sub(PlanningConfiguration)),
     ("sampling", SamplingConfiguration, _sub(SamplingConfiguration)),
     ("debug", DebugConfiguration, _sub(DebugConfiguration)),
     ("general", GeneralConfiguration, _sub(GeneralConfiguration))],
    bases=(BaseConfiguration,),
    namespace={"__post_init__": _root_post_init, "update": _root_update,
               "name_scenario": property(lambda self: self.general.name_scenario)})
ReactivePlannerConfiguration.__doc__ = "Configuration parameters for reactive planner."
ReactivePlannerConfiguration.__module__ = __name__
