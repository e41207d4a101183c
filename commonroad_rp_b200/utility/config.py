"""Configuration tree of the planner -- same classes, field names and defaults as the reference's
``commonroad_rp/utility/config.py`` (dataclass tree :107-290), so existing YAML files and user code
keep working.  Differences that do not touch the API:

* YAML is read with OmegaConf when it is installed, otherwise with PyYAML (not installed here).
* Vehicle parameters come from ``vehiclemodels`` when importable, otherwise from the built-in table of
  commonroad-vehicle-models 3.0.2 values below (type 2 = BMW 320i is the reference default, :198).
* ``debug.multiproc`` / ``debug.num_workers`` are accepted and ignored: the candidate loop runs on the
  GPU, there is nothing to fork.
"""
import dataclasses
import inspect
import os.path
import pathlib
import warnings
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Union

import numpy as np


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


# commonroad-vehicle-models 3.0.2 (type 2 values per SURVEY.md App. D#5; 1 and 3 for completeness)
_VEHICLE_TABLE = {
    1: dict(l=4.298, w=1.674, a=0.88392, b=1.50876, a_max=11.5, v_switch=4.755, s_min=-0.910, s_max=0.910,
            sv_min=-0.4, sv_max=0.4),
    2: dict(l=4.508, w=1.610, a=1.1561957064, b=1.4227170936, a_max=11.5, v_switch=7.319, s_min=-1.066, s_max=1.066,
            sv_min=-0.4, sv_max=0.4),
    3: dict(l=4.569, w=1.844, a=1.08966, b=1.35634, a_max=11.5, v_switch=7.824, s_min=-1.023, s_max=1.023,
            sv_min=-0.4, sv_max=0.4),
}


def vehicle_parameters_from_type(vehicle_type_id: int):
    """VehicleParameterMapping.from_vehicle_type(VehicleType(id)) stand-in (reference :200)."""
    try:
        from commonroad.common.solution import VehicleType
        from commonroad_dc.feasibility.vehicle_dynamics import VehicleParameterMapping
        return VehicleParameterMapping.from_vehicle_type(VehicleType(vehicle_type_id))
    except Exception:
        pass
    row = _VEHICLE_TABLE[int(vehicle_type_id)]
    return _NS(l=row["l"], w=row["w"], a=row["a"], b=row["b"],
               longitudinal=_NS(a_max=row["a_max"], v_switch=row["v_switch"]),
               steering=_NS(min=row["s_min"], max=row["s_max"], v_min=row["sv_min"], v_max=row["sv_max"]))


def _dict_to_params(dict_params: Dict[str, Any], cls: Any) -> Any:
    kwargs = {}
    for f in dataclasses.fields(cls):
        if not f.init or f.name not in dict_params:
            continue
        if inspect.isclass(f.type) and issubclass(f.type, BaseConfiguration):
            kwargs[f.name] = _dict_to_params(dict_params[f.name], f.type)
        else:
            kwargs[f.name] = dict_params[f.name]
    return cls(**kwargs)


def _load_yaml(file_path) -> dict:
    try:
        from omegaconf import OmegaConf
        loaded = OmegaConf.to_object(OmegaConf.load(file_path))
        if isinstance(loaded, dict):
            return loaded
    except Exception:       # not installed (or a test stub of it): plain YAML is equivalent for these files
        pass
    import yaml
    with open(file_path) as f:
        return yaml.safe_load(f) or {}


@dataclass
class BaseConfiguration:
    """Reactive planner base parameters."""

    def __getitem__(self, item: str) -> Any:
        try:
            return self.__getattribute__(item)
        except AttributeError as e:
            raise KeyError(f"{item} is not a parameter of {self.__class__.__name__}") from e

    def __setitem__(self, key: str, value: Any):
        try:
            self.__setattr__(key, value)
        except AttributeError as e:
            raise KeyError(f"{key} is not a parameter of {self.__class__.__name__}") from e

    @classmethod
    def load(cls, file_path: Union[pathlib.Path, str], scenario_name: Optional[str] = None,
             validate_types: bool = True) -> 'ReactivePlannerConfiguration':
        """Loads parameters from a config yaml file (reference :84-104)."""
        file_path = pathlib.Path(file_path)
        assert file_path.suffix == ".yaml", f"File type {file_path.suffix} is unsupported! Please use .yaml!"
        params = _dict_to_params(_load_yaml(file_path), cls)
        if scenario_name:
            params.general.set_path_scenario(scenario_name)
        return params


@dataclass
class PlanningConfiguration(BaseConfiguration):
    """Planning parameters for reactive planner."""
    dt: float = 0.1
    time_steps_computation: int = 60
    planning_horizon: float = dt * time_steps_computation
    replanning_frequency: int = 3
    continuous_collision_check: bool = False
    factor: int = 1
    low_vel_mode_threshold: float = 4.0
    constraints_to_check: List[str] = \
        field(default_factory=lambda: ["velocity", "acceleration", "kappa", "kappa_dot", "yaw_rate"])
    standstill_lookahead: int = 10


@dataclass
class SamplingConfiguration(BaseConfiguration):
    """Sampling parameters for reactive planner."""
    sampling_method: int = 1
    longitudinal_mode: str = "velocity_keeping"
    num_sampling_levels: int = 4
    t_min: float = 0.4
    v_min: float = 0
    v_max: float = 0
    s_min: float = -1
    s_max: float = 1
    d_min: float = -3
    d_max: float = 3


@dataclass
class DebugConfiguration(BaseConfiguration):
    """Parameters specifying debug-related information."""
    save_plots: bool = False
    save_config: bool = False
    show_plots: bool = False
    draw_ref_path: bool = True
    draw_planning_problem: bool = True
    draw_icons: bool = False
    draw_traj_set: bool = False
    logging_level: str = "INFO"
    multiproc: bool = True
    num_workers: int = 6


@dataclass
class VehicleConfiguration(BaseConfiguration):
    """Class to store vehicle configurations"""
    id_type_vehicle: int = 2
    vehicle_parameters: Any = None
    length: float = None
    width: float = None
    wb_front_axle: float = None
    wb_rear_axle: float = None
    a_max: float = None
    v_switch: float = None
    delta_min: float = None
    delta_max: float = None
    v_delta_min: float = None
    v_delta_max: float = None
    wheelbase: float = None

    def __post_init__(self):
        vp = self.vehicle_parameters or vehicle_parameters_from_type(self.id_type_vehicle)
        self.vehicle_parameters = vp
        derived = dict(length=vp.l, width=vp.w, wb_front_axle=vp.a, wb_rear_axle=vp.b,
                       a_max=vp.longitudinal.a_max, v_switch=vp.longitudinal.v_switch, delta_min=vp.steering.min,
                       delta_max=vp.steering.max, v_delta_min=vp.steering.v_min, v_delta_max=vp.steering.v_max,
                       wheelbase=vp.a + vp.b)
        for name, value in derived.items():
            if getattr(self, name) is None:
                setattr(self, name, value)
        self.kappa_max = np.tan(self.delta_max) / self.wheelbase


@dataclass
class GeneralConfiguration(BaseConfiguration):
    """General parameters for evaluations."""
    path_scenarios: str = "example_scenarios/"
    path_output: str = "output/"
    path_logs: str = "output/logs/"
    path_pickles: str = "output/pickles/"
    path_scenario: Optional[str] = None
    name_scenario: Optional[str] = None

    def set_path_scenario(self, scenario_name: str):
        self.path_scenario = os.path.join(self.path_scenarios, scenario_name)


@dataclass
class ReactivePlannerConfiguration(BaseConfiguration):
    """Configuration parameters for reactive planner."""
    vehicle: VehicleConfiguration = field(default_factory=VehicleConfiguration)
    planning: PlanningConfiguration = field(default_factory=PlanningConfiguration)
    sampling: SamplingConfiguration = field(default_factory=SamplingConfiguration)
    debug: DebugConfiguration = field(default_factory=DebugConfiguration)
    general: GeneralConfiguration = field(default_factory=GeneralConfiguration)

    def __post_init__(self):
        self.scenario = None
        self.planning_problem = None
        self.planning_problem_set = None

    @property
    def name_scenario(self) -> str:
        return self.general.name_scenario

    def update(self, scenario=None, planning_problem=None, idx_planning_problem: Optional[int] = None):
        """Updates configuration based on the given attributes (reference :265-290)."""
        self.scenario = scenario
        self.planning_problem = planning_problem
        if scenario is None and planning_problem is None:
            try:
                from commonroad_rp_b200.utility.general import load_scenario_and_planning_problem
                self.scenario, self.planning_problem, self.planning_problem_set = \
                    load_scenario_and_planning_problem(self.general.path_scenario, idx_planning_problem)
            except FileNotFoundError:
                warnings.warn(f"<ReactivePlannerConfiguration.update()>: No scenario .xml file found at "
                              f"path_scenario = {self.general.path_scenario}")
        assert self.scenario is not None, "<Configuration.update()>: no scenario has been specified"
