"""``ReactivePlanner`` -- drop-in for the reference's ``commonroad_rp/reactive_planner.py``: same
constructor, public methods, properties and return values; the candidate hot path underneath
(``_create_trajectory_bundle`` + ``_get_optimal_trajectory``, reference :421-444 and :1065-1136 with
``_check_kinematics`` :715-969, ``_check_constraints`` :971-1017 and ``_check_collisions`` :1019-1063
beneath them) runs as hand-written CUDA through the C-ABI of include/rp_b200.h.

Per cycle the host iterates the sampling sets (their order IS the enumeration order), sends three small
arrays + one struct, and reads back one result record and the winner's 14 x (N+1) state block.  The
reference path and the obstacle tables are uploaded once and stay device resident across ``reset()``
calls.  There is no CPU fallback: without the CUDA library or a device, ``plan()`` raises.

``config.debug.multiproc`` / ``num_workers`` are accepted and ignored (the reference forks workers over
candidate chunks, :1084-1111; here every (candidate, time step) pair is a GPU thread).
"""
import collections.abc
import copy
import logging
import math
import time
from typing import Dict, List, Optional, Tuple, Type, Union

import numpy as np

from commonroad_rp_b200 import _lib
from commonroad_rp_b200 import collision as rpc
from commonroad_rp_b200._compat import HAVE_COMMONROAD_IO, CustomState, InputState, Trajectory
from commonroad_rp_b200.cost_function import CostFunction, DefaultCostFunction
from commonroad_rp_b200.polynomial_trajectory import QuarticTrajectory, QuinticTrajectory
from commonroad_rp_b200.sampling import (PositionSampling, SamplingSpace, TimeSampling, VelocitySampling,
                                          sampling_space_factory)
from commonroad_rp_b200.state import ReactivePlannerState
from commonroad_rp_b200.trajectories import (CartesianSample, CurviLinearSample, DeviceTrajectorySample, FeasibilityStatus,
                                              TrajectoryBundle, TrajectorySample, _DeviceBacking)
from commonroad_rp_b200.utility.config import ReactivePlannerConfiguration, VehicleConfiguration
from commonroad_rp_b200.utility.general import retrieve_desired_velocity_from_pp, shift_orientation
from commonroad_rp_b200.utility.utils_coordinate_system import CoordinateSystem, interpolate_angle

class _LazyViews(collections.abc.Sequence):
    """``stored_trajectories`` in draw mode: a read-only sequence of TrajectorySample views on the device results of
    the cycle, each created (and cached) on first access.  ``copy.deepcopy`` gives a plain list of materialised copies."""

    def __init__(self, make, order):
        """order: the candidate indices in list order, or a callable producing them on first use (the per-candidate
        verdicts are then fetched from the device only if somebody looks at the stored trajectories)"""
        self._make, self._order_src, self._cache = make, order, {}

    @property
    def _order(self):
        if callable(self._order_src):
            self._order_src = self._order_src()
        return self._order_src

    def __len__(self):
        return len(self._order)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError("trajectory index out of range")
        view = self._cache.get(i)
        if view is None:
            view = self._cache[i] = self._make(int(self._order[i]))
        return view

    def __deepcopy__(self, memo):
        return [copy.deepcopy(v, memo) for v in self]


logger = logging.getLogger("RP_LOGGER")

try:        # CPython helper for the output packing (csrc/rp_pack.c, built by build.py); the Python loop below is the fallback
    from commonroad_rp_b200 import _rp_pack
except ImportError:
    _rp_pack = None
_PACK_MAKES_POSITIONS = bool(getattr(_rp_pack, "MAKES_POSITIONS", 0))

_EPS = 1e-5
_TWO_PI = 2.0 * math.pi


def _curvilinear_states(t0, factor, pos_curv, v_l, a_l, th_l, kap_l):
    """the curvilinear CustomState list of plan()'s result (reference :543-556: velocity, acceleration and orientation
    are the CARTESIAN ones, yaw_rate carries the curvature)"""
    out = []
    new_cs = CustomState.__new__
    ts = t0
    for pos, vel, acc, th, kap in zip(pos_curv, v_l, a_l, th_l, kap_l):
        c = new_cs(CustomState)
        c.__dict__ = {"time_step": ts, "position": pos, "velocity": vel, "acceleration": acc, "orientation": th, "yaw_rate": kap}
        out.append(c)
        ts += factor
    return out


_REASON_INDEX = {name: i for i, name in enumerate(_lib.REASON_NAMES)}


class _LazyTrajectory(Trajectory):
    """Trajectory whose ``state_list`` is built on first access (only used with the package's own stand-in classes)"""

    def __init__(self, initial_time_step, make_states):
        self.initial_time_step = initial_time_step
        self._make_states = make_states
        self._states = None

    @property
    def state_list(self):
        if self._states is None:
            self._states = self._make_states()
            self._make_states = None
        return self._states

    @state_list.setter
    def state_list(self, value):
        self._states = value
        self._make_states = None


class _BundleArrays(dict):
    """Per-candidate verdict arrays of the device bundle, fetched on first use and shared by all the
    lazy TrajectorySample views of that bundle."""

    def __init__(self, engine, generation, all_states, level=None):
        super().__init__()
        self._engine = engine
        self._generation = generation
        self._level = level              # cycle launch (several levels evaluated together): this bundle's level index
        dict.__setitem__(self, "all_states", all_states)

    def select(self):
        """make the engine's fetch_* calls address this bundle's level (cycle launches hold several)"""
        if self._level is not None:
            self._engine.select_level(self._level)

    def __missing__(self, key):
        if self._engine.plan_generation != self._generation:
            raise RuntimeError("device bundle replaced by a newer plan() call")
        self.select()
        cost, status, reason, step = self._engine.fetch_candidates()
        self.update(cost=cost, status=status, reason=reason, step=step)
        return dict.__getitem__(self, key)


class ReactivePlanner(object):
    """Reactive planner class that plans trajectories in a sampling-based fashion (GPU candidate path)."""

    def __init__(self, config: ReactivePlannerConfiguration, device: Optional[int] = None, stream=None):
        """
        :param config: configuration object holding all planner-relevant configurations
        :param device: CUDA device index (default: torch's current device)
        :param stream: cudaStream_t handle to run on (default: torch's current stream)
        """
        self.dt: float = config.planning.dt
        self.N: int = config.planning.time_steps_computation
        self.horizon: float = config.planning.dt * config.planning.time_steps_computation
        self.vehicle_params: VehicleConfiguration = config.vehicle

        self.x_0: Optional[ReactivePlannerState] = None
        self.x_0_cl: Optional[Tuple[List, List]] = None
        self._co: Optional[CoordinateSystem] = None
        self._cc: Optional[rpc.CollisionChecker] = None

        self._infeasible_count_collision: int = 0
        self._infeasible_count_kinematics: int = 0
        self._infeasible_reason_dict: Dict = dict()
        self._optimal_cost: float = 0.0
        self._planning_times_list: List = list()
        self._record_state_list: List[ReactivePlannerState] = list()
        self._record_input_list: List[InputState] = list()
        self.stored_trajectories: Optional[List[TrajectorySample]] = None

        self._desired_speed: Optional[float] = None
        self._desired_lon_position: Optional[float] = None
        self._low_vel_mode = False
        self._draw_traj_set = config.debug.draw_traj_set and (config.debug.show_plots or config.debug.save_plots)

        # device side (created lazily: constructing a planner needs no GPU, planning does)
        self._device_index = device
        self._stream = stream
        self._engine: Optional[_lib.Engine] = None
        self._uploaded_co = None
        self._uploaded_cc = None
        self._uploaded_vehicle = None
        self.last_result = None        # rp_plan_result of the last evaluated level
        # level escalation (:616-636) in one submission: plan() lets the device evaluate the next sampling levels together
        # with the current one -- as long as their candidates x time steps stay below this budget (about one wave of the
        # step-parallel kernel) -- and picks the lowest level with a winner; 0 = one level per submission
        self.speculation_budget = 200000
        # when plan() lets the device evaluate the following sampling levels in the same launch: "after_failure" -- the
        # first level of a cycle goes alone (it nearly always has a winner), and when it has none ALL the remaining
        # levels go in the second launch (an escalating cycle costs two submissions, never three); "always" -- from the
        # first launch on (one submission per cycle whatever happens, ~15 us more host work in the cycles that do not
        # escalate); "never"
        self.speculation = "after_failure"
        self._escalating = False       # inside plan()'s escalation loop
        self._level_failed = False     # a level of the current plan() call ended without a winner
        self._spec = None              # records of the levels evaluated ahead in the current cycle
        self._plan_serial = 0
        self._inputs = None            # the rp_plan_inputs struct, reused across cycles
        self._constraint_key, self._constraint_mask = None, 0
        self._grid_cache = {}          # per sampling level: ordered t / lon arrays (+ traj_len) while the sample sets live

        self.config: Optional[ReactivePlannerConfiguration] = None
        self.reset(config)

        self.sampling_space: Optional[Type[SamplingSpace]] = None
        self.set_sampling_space()
        self.sampling_level = config.sampling.num_sampling_levels

        self.cost_function: Optional[Type[CostFunction]] = None
        self.set_cost_function()
        self._standstill_lookahead = config.planning.standstill_lookahead

    # ------------------------------------------------------------------ properties (reference :115-160)
    @property
    def collision_checker(self):
        return self._cc

    @property
    def coordinate_system(self) -> CoordinateSystem:
        return self._co

    @property
    def reference_path(self) -> np.ndarray:
        return self._co.reference

    @property
    def infeasible_count_collision(self) -> float:
        return self._infeasible_count_collision

    @property
    def infeasible_count_kinematics(self) -> float:
        return self._infeasible_count_kinematics

    @property
    def infeasible_reason_dict(self) -> dict:
        return self._infeasible_reason_dict

    @property
    def optimal_cost(self) -> float:
        return self._optimal_cost

    @property
    def planning_times(self) -> List:
        return self._planning_times_list

    @property
    def record_state_list(self) -> List:
        return self._record_state_list

    @property
    def record_input_list(self) -> List:
        return self._record_input_list

    @property
    def engine(self) -> _lib.Engine:
        """The device context of this planner (created on first use)."""
        if self._engine is None:
            if self._device_index is None or self._stream is None:
                from commonroad_rp_b200._device import current_device_and_stream
                dev, stream = current_device_and_stream()
                self._device_index = dev if self._device_index is None else self._device_index
                self._stream = stream if self._stream is None else self._stream
            self._engine = _lib.Engine(self._device_index, self._stream)
        return self._engine

    # ------------------------------------------------------------------ configuration (reference :162-419)
    def goal_reached(self) -> bool:
        x_0_shifted = self.x_0.shift_positions_to_center(self.vehicle_params.wb_rear_axle)
        if self.config.planning_problem.goal.is_reached(x_0_shifted):
            logger.info("Goal of planning problem reached")
            return True
        return False

    def reset(self, config: ReactivePlannerConfiguration = None, initial_state_cart: ReactivePlannerState = None,
              initial_state_curv: Tuple[List, List] = None, collision_checker=None,
              coordinate_system: CoordinateSystem = None):
        """Initializes/resets configuration of the planner for re-planning purposes (reference :172-216)."""
        if config is not None:
            self.config = config
        else:
            assert self.config is not None, "<ReactivePlanner.reset(). No Configuration object provided>"
        self._spec = None                  # levels evaluated ahead belong to the previous initial state
        self._plan_serial += 1
        self._reset_statistics()
        if collision_checker is None:
            self.set_collision_checker(scenario=self.config.scenario)
        else:
            self.set_collision_checker(collision_checker=collision_checker)
        if coordinate_system is not None:
            self.set_reference_path(coordinate_system=coordinate_system)
        if self.x_0 is None and initial_state_cart is None:
            if self.config.planning_problem:
                self.x_0 = ReactivePlannerState.create_from_initial_state(self.config.planning_problem.initial_state,
                                                                          self.vehicle_params.wheelbase,
                                                                          self.vehicle_params.wb_rear_axle)
            else:
                self.x_0 = None
        else:
            self.x_0 = initial_state_cart if initial_state_cart is not None else self.x_0
        self.x_0_cl = initial_state_curv if initial_state_curv is not None else self._compute_initial_states(self.x_0)

    def set_collision_checker(self, scenario=None, collision_checker=None, road_boundary_obstacle=None):
        """Either adopt a ``collision.CollisionChecker`` or build one from a CommonRoad scenario: static
        obstacles, dynamic obstacles as time-variant objects, and the road boundary (reference :218-256).
        The checker's content becomes the device-resident obstacle table on the next plan()."""
        if collision_checker is None:
            assert scenario is not None, '<ReactivePlanner.set collision checker>: Please provide a CommonRoad ' \
                                         'scenario OR a CollisionChecker object to the planner.'
            cc_scenario = rpc.CollisionChecker()
            for co in scenario.static_obstacles:
                cc_scenario.add_collision_object(rpc.create_collision_object(co))
            for co in scenario.dynamic_obstacles:
                tvo = rpc.create_collision_object(co)
                if self.config.planning.continuous_collision_check:
                    tvo, err = rpc.trajectory_preprocess_obb_sum(tvo)
                    if err == -1:
                        raise Exception("Invalid input for trajectory_preprocess_obb_sum: dynamic "
                                        "obstacle elements overlap")
                cc_scenario.add_collision_object(tvo)
            if road_boundary_obstacle is None:
                _, road_boundary_sg = rpc.create_road_boundary_obstacle(scenario)
                cc_scenario.add_collision_object(road_boundary_sg)
            else:
                cc_scenario.add_collision_object(road_boundary_obstacle)
            self._cc = cc_scenario
        else:
            assert scenario is None, '<ReactivePlanner.set collision checker>: Please provide a CommonRoad scenario ' \
                                     'OR a CollisionChecker object to the planner.'
            if not isinstance(collision_checker, rpc.CollisionChecker):
                raise TypeError("<ReactivePlanner.set_collision_checker>: expected a commonroad_rp_b200.collision."
                                "CollisionChecker (a pycrcc checker cannot be unpacked into device tables)")
            self._cc = collision_checker

    def set_reference_path(self, reference_path: np.ndarray = None, coordinate_system: CoordinateSystem = None):
        """Create the curvilinear coordinate system from a polyline or adopt a given one (reference :258-272)."""
        if coordinate_system is None:
            assert reference_path is not None, '<set reference path>: Please provide a reference path OR a ' \
                                               'CoordinateSystem object to the planner.'
            self._co = CoordinateSystem(reference_path)
        else:
            assert reference_path is None, '<set reference path>: Please provide a reference path OR a ' \
                                           'CoordinateSystem object to the planner.'
            self._co = coordinate_system

    def set_t_sampling_parameters(self, t_min):
        self.sampling_space.samples_t = TimeSampling(t_min, self.horizon, self.sampling_level, self.dt)
        logger.debug("Sampled interval of time: {} s - {} s".format(t_min, self.horizon))

    def set_d_sampling_parameters(self, delta_d_min, delta_d_max):
        self.sampling_space.samples_d = PositionSampling(delta_d_min, delta_d_max, self.sampling_level)
        logger.debug("Sampled interval of lateral position: {} m - {} m".format(delta_d_min, delta_d_max))

    def set_v_sampling_parameters(self, v_min, v_max):
        self.sampling_space.samples_v = VelocitySampling(v_min, v_max, self.sampling_level)
        logger.info("Sampled interval of velocity: {} m/s - {} m/s".format(v_min, v_max))

    def set_s_sampling_parameters(self, s_min, s_max):
        self.sampling_space.samples_s = PositionSampling(s_min, s_max, self.sampling_level)
        logger.info("Sampled interval of longitudinal position: {} m - {} m".format(s_min, s_max))

    def set_desired_velocity(self, desired_velocity: float = None, current_speed: float = None, stopping: bool = False):
        """Sets desired velocity and re-calculates velocity samples (reference :309-347)."""
        self._desired_lon_position = None
        if desired_velocity is None and self._desired_speed is None:
            self._desired_speed = retrieve_desired_velocity_from_pp(self.config.planning_problem)
        else:
            self._desired_speed = desired_velocity if desired_velocity is not None else self._desired_speed
        assert self._desired_speed >= 0.0, f"<ReactivePlanner.set_desired_velocity(): desired speed has to be " \
                                           f"positive. Provided speed{self._desired_speed}>"
        if not stopping:
            reference_speed = current_speed if current_speed is not None else self._desired_speed
            min_v = max(0, reference_speed - (0.125 * self.horizon * self.vehicle_params.a_max))
            max_v = max(min_v + 5.0, reference_speed + 2)
            self.set_v_sampling_parameters(min_v, max_v)
        else:
            self.set_v_sampling_parameters(v_min=self._desired_speed, v_max=self._desired_speed)
        if hasattr(self.cost_function, "desired_speed"):
            self.cost_function.desired_speed = self._desired_speed
        if hasattr(self.cost_function, "w_a"):
            self.cost_function.w_a = 5
        if hasattr(self.cost_function, "desired_s"):
            self.cost_function.desired_s = self._desired_lon_position

    def set_desired_lon_position(self, lon_position: float, delta_s_min: Optional[float] = None,
                                 delta_s_max: Optional[float] = None):
        """Sets a desired longitudinal position for stopping and re-calculates s samples (reference :349-376)."""
        self._desired_lon_position = lon_position
        self._desired_speed = 0.0
        if delta_s_min is None and delta_s_max is None:
            delta_s_min = self.config.sampling.s_min
            delta_s_max = self.config.sampling.s_max
        self.set_s_sampling_parameters(s_min=lon_position + delta_s_min, s_max=lon_position + delta_s_max)
        if hasattr(self.cost_function, "desired_s"):
            self.cost_function.desired_s = self._desired_lon_position
        if hasattr(self.cost_function, "desired_speed"):
            self.cost_function.desired_speed = self._desired_speed
        if hasattr(self.cost_function, "w_a"):
            self.cost_function.w_a = 1

    def set_cost_function(self, cost_function: Type[CostFunction] = None):
        if cost_function:
            self.cost_function = cost_function
        else:
            self.cost_function = DefaultCostFunction(self._desired_speed, desired_d=0.0,
                                                     desired_s=self._desired_lon_position)

    def set_sampling_space(self, sampling_space: Type[SamplingSpace] = None):
        if sampling_space:
            self.sampling_space = sampling_space
        else:
            self.sampling_space = sampling_space_factory(self.config)

    def record_state_and_input(self, state: ReactivePlannerState):
        """Adds state and the derived control input to the recorded lists (reference :391-408)."""
        self.record_state_list.append(state)
        if len(self.record_state_list) > 1:
            steering_angle_speed = (state.steering_angle - self.record_state_list[-2].steering_angle) / self.dt
        else:
            steering_angle_speed = 0.0
        self.record_input_list.append(InputState(time_step=state.time_step, acceleration=state.acceleration,
                                                 steering_angle_speed=steering_angle_speed))

    def _reset_statistics(self):
        self._optimal_cost = 0
        self._infeasible_count_kinematics = 0
        self._infeasible_count_collision = 0
        for constraint in self.config.planning.constraints_to_check:
            self._infeasible_reason_dict[constraint] = 0

    # ------------------------------------------------------------------ device tables
    def _sync_frame_tables(self):
        """Upload vehicle / reference tables when they changed since the last cycle."""
        eng = self.engine
        vp = self.vehicle_params
        vkey = (vp.length, vp.width, vp.wb_rear_axle, vp.wheelbase, vp.a_max, vp.v_switch, vp.delta_max, vp.v_delta_max)
        if self._uploaded_vehicle != vkey:
            # kappa_max exactly as the reference computes it per check (:985)
            eng.set_vehicle(*vkey, kappa_max=np.tan(vp.delta_max) / vp.wheelbase)
            self._uploaded_vehicle = vkey
        if self._uploaded_co is not self._co:
            tb = self._co.device_tables()
            eng.set_reference(tb["ref_pos"], tb["ref_theta"], tb["ref_curv"], tb["ref_curv_d"], tb["path_xy"],
                              tb["path_s"], tb["path_normals"], tb["proj_limit"])
            self._uploaded_co = self._co

    def _sync_device_tables(self):
        """Upload vehicle / reference / obstacle tables when they changed since the last cycle."""
        self._sync_frame_tables()
        eng = self.engine
        # the uploaded checker is kept alive and compared by identity (id() values are reused once an object is freed),
        # plus its content version (objects added / shapes appended since the upload)
        up = self._uploaded_cc
        if up is None or up[0] is not self._cc or up[1] != self._cc.version:
            self._cc.upload(eng)
            self._uploaded_cc = (self._cc, self._cc.version)

    def _plan_inputs(self, x_0_lon, x_0_lat, cost_spec, want_all_states, check_collision=None):
        """The cycle's rp_plan_inputs (one struct per planner, refilled every cycle)."""
        p = self.config.planning
        pi = self._inputs
        if pi is None:
            pi = self._inputs = _lib.PlanInputs()
        key = tuple(p.constraints_to_check)
        if key != self._constraint_key:
            mask = 0
            for name in key:
                mask |= _lib.CONSTRAINT_BITS[name]
            self._constraint_key, self._constraint_mask = key, mask
        kind = cost_spec["cost_kind"]
        ds, dp = cost_spec["desired_speed"], cost_spec["desired_s"]
        if check_collision is None:
            # the reference's collision pass is lazy (:1031-1063); full flags only when every trajectory is kept
            check_collision = _lib.COLLISION_ALL if (want_all_states or kind == _lib.COST_NONE) else _lib.COLLISION_LAZY
        # one pack_into for the whole struct (field order of _lib.PlanInputs / rp_plan_inputs)
        _lib.PLAN_INPUTS_STRUCT.pack_into(
            pi, 0, x_0_lon[0], x_0_lon[1], x_0_lon[2], x_0_lat[0], x_0_lat[1], x_0_lat[2], self.x_0.orientation,
            int(self.x_0.time_step), 1 if self._low_vel_mode else 0,
            _lib.STOPPING if self.config.sampling.longitudinal_mode == "stopping" else _lib.VELOCITY_KEEPING, self.N,
            self.dt, int(p.factor), 1 if self._draw_traj_set else 0, self._constraint_mask,
            kind, 0 if ds is None else 1, 0 if dp is None else 1,
            0.0 if ds is None else ds, 0.0 if dp is None else dp, cost_spec["desired_d"], cost_spec["w_a"],
            1 if want_all_states else 0, check_collision,
            # reference :1049-1058 (the hull check of the first discretely collision-free candidate, on the device)
            1 if (p.continuous_collision_check and kind != _lib.COST_NONE) else 0, 0)
        return pi

    def _device_cost_spec(self):
        """Fused device cost for the built-in cost functions; None for user subclasses that bring their own
        ``evaluate`` (generic plug-in path)."""
        cf = self.cost_function
        spec = getattr(cf, "device_spec", None)
        if spec is None:
            return None
        for base in type(cf).__mro__:
            if "evaluate" in base.__dict__:
                # the class that defines evaluate must be the one that defines device_spec
                return spec() if "device_spec" in base.__dict__ else None
        return None

    # ------------------------------------------------------------------ hot path
    def _create_trajectory_bundle(self, x_0_lon: np.array, x_0_lat: np.array, samp_level: int) -> TrajectoryBundle:
        """Candidate bundle of a sampling level (reference :421-444).  For grid sampling spaces only the three
        ordered sample arrays are produced here; TrajectorySample objects are views created on access."""
        log = logger.isEnabledFor(logging.INFO)        # (the f-strings below are formatted before logger.info sees them)
        if log:
            logger.info("===== Sampling trajectories ... =====")
            logger.info(f"Sampling density {samp_level + 1} of {self.sampling_level}")
        mode = self.config.sampling.longitudinal_mode
        if hasattr(self.sampling_space, "sample_grid"):
            t, lon, d = self.sampling_space.sample_grid(samp_level, x_0_lat, mode)
            bundle = TrajectoryBundle(lambda: self._materialise_views(bundle), cost_function=self.cost_function)
            # (the initial states as plain tuples: read three times per cycle, turned into arrays only by the lazy
            # polynomial objects)
            bundle.device = {"kind": "grid", "t": t, "lon": lon, "d": d, "x_0_lon": tuple(x_0_lon),
                             "x_0_lat": tuple(x_0_lat), "mode": mode,
                             "low_vel": bool(self._low_vel_mode), "n": len(t) * len(lon) * len(d), "level": samp_level,
                             "serial": self._plan_serial}
        else:
            trajectories = self.sampling_space.generate_trajectories_at_level(samp_level, x_0_lon, x_0_lat, mode,
                                                                              self._low_vel_mode)
            bundle = TrajectoryBundle(trajectories, cost_function=self.cost_function)
            bundle.device = {"kind": "list", "n": len(trajectories), "x_0_lon": np.asarray(x_0_lon, dtype=np.float64),
                             "x_0_lat": np.asarray(x_0_lat, dtype=np.float64)}
        if log:
            logger.info(f"Number of trajectory samples: {bundle.device['n']}")
        return bundle

    def _grid_polynomials(self, dev, k):
        """Polynomial objects of grid candidate k (coefficients are solved lazily on the device if read)."""
        n_lon, n_d = len(dev["lon"]), len(dev["d"])
        it, rem = divmod(k, n_lon * n_d)
        il, idd = divmod(rem, n_d)
        t, lon, d = float(dev["t"][it]), float(dev["lon"][il]), float(dev["d"][idd])
        if dev["mode"] == "velocity_keeping":
            tl = QuarticTrajectory(tau_0=0, delta_tau=t, x_0=np.array(dev["x_0_lon"], dtype=np.float64), x_d=np.array([lon, 0.0]))
        else:
            tl = QuinticTrajectory(tau_0=0, delta_tau=t, x_0=np.array(dev["x_0_lon"], dtype=np.float64), x_d=np.array([lon, 0.0, 0.0]))
        tau_lat = t
        if dev["low_vel"]:
            s_goal = tl.evaluate_state_at_tau(t)[0] - dev["x_0_lon"][0]
            tau_lat = t if s_goal <= 0 else s_goal
        lat = QuinticTrajectory(tau_0=0, delta_tau=tau_lat, x_0=np.array(dev["x_0_lat"], dtype=np.float64), x_d=np.array([d, 0.0, 0.0]))
        return tl, lat

    def _view(self, bundle, k, arrays):
        dev = bundle.device
        if dev["kind"] == "grid":
            # the polynomial objects are built when somebody reads them (most callers only read the sampled states)
            sample = DeviceTrajectorySample(self.horizon, self.dt, lambda: self._grid_polynomials(dev, k))
        else:
            sample = dev["candidates"][k]
        return sample._attach(_DeviceBacking(self.engine, k, arrays, dev["generation"]))

    def _materialise_views(self, bundle):
        dev = bundle.device
        if "generation" not in dev:
            # bundle not evaluated yet: plain candidates with lazily solved polynomials
            out = []
            for k in range(dev["n"]):
                tl, lat = self._grid_polynomials(dev, k)
                out.append(TrajectorySample(self.horizon, self.dt, tl, lat))
            return out
        return [self._view(bundle, k, dev["arrays"]) for k in range(dev["n"])]

    def _cycle_levels(self, dev, cost_generic):
        """The sampling levels one cycle launch (rp_plan_levels) evaluates for this bundle: the bundle's own level and --
        inside plan()'s escalation loop -- the following ones while the launch stays within ``speculation_budget``
        candidate-timesteps and the library's limits.  None: not eligible (one rp_plan_grid per level instead)."""
        p = self.config.planning
        Np1 = self.N + 1
        if dev["kind"] != "grid" or Np1 > 256 or p.continuous_collision_check:
            return None
        max_levels, max_samples, max_segments, max_work = self.engine.cycle_limits()
        dt = self.dt
        levels = [(dev["level"], dev["t"], dev["lon"], dev["d"], self._traj_lens(dev["t"], dt))]
        n_samp = len(dev["t"]) + len(dev["lon"]) + len(dev["d"])
        n_seg = len(dev["t"])
        work = dev["n"] * Np1
        if n_samp > max_samples or n_seg > max_segments or work > max_work:
            return None
        ahead = self._escalating and not cost_generic and self.speculation_budget > 0 and (
            self.speculation == "always" or (self.speculation == "after_failure" and self._level_failed))
        if ahead:
            for lvl in range(dev["level"] + 1, self.sampling_level):
                if len(levels) >= max_levels:
                    break
                t, lon, d = self.sampling_space.sample_grid(lvl, dev["x_0_lat"], dev["mode"])
                n = len(t) * len(lon) * len(d)
                n_samp += len(t) + len(lon) + len(d)
                n_seg += len(t)
                work += n * Np1
                if work > min(self.speculation_budget, max_work) or n_samp > max_samples or n_seg > max_segments:
                    break
                levels.append((lvl, t, lon, d, self._traj_lens(t, dt)))
        return levels

    _traj_len_cache: Dict = {}

    @classmethod
    def _traj_lens(cls, t, dt):
        """traj_len of every sampled horizon (reactive_planner.py:733); the sampled horizons of a level repeat every cycle"""
        hit = cls._traj_len_cache.get(id(t))           # (the ordered t arrays are cached objects, sampling._ordered)
        if hit is None or hit[0] is not t or hit[1] != dt:
            if len(cls._traj_len_cache) > 256:
                cls._traj_len_cache.clear()
            tl = np.array([_lib.traj_len_of(x, dt) for x in t], dtype=np.int32)
            tl.flags.writeable = False
            hit = cls._traj_len_cache[id(t)] = (t, dt, tl)
        return hit[2]

    def _get_optimal_trajectory(self, trajectory_bundle: TrajectoryBundle) -> Union[TrajectorySample, None]:
        """Kinematic check, cost, collision check and selection of the optimal candidate in one device pass
        (reference :1065-1136).  Returns the optimal TrajectorySample or None."""
        log = logger.isEnabledFor(logging.INFO)
        if log:
            logger.info("===== Checking trajectories ... =====")
        self._reset_statistics()
        self._sync_device_tables()
        eng = self.engine
        dev = trajectory_bundle.device
        if dev is None:         # bundle built by user code from a plain list
            dev = trajectory_bundle.device = {"kind": "list", "n": len(trajectory_bundle.trajectories),
                                              "x_0_lon": np.asarray(self.x_0_cl[0], dtype=np.float64),
                                              "x_0_lat": np.asarray(self.x_0_cl[1], dtype=np.float64)}
        cost_spec = self._device_cost_spec()
        generic_cost = cost_spec is None
        if generic_cost and self.config.planning.continuous_collision_check:
            raise NotImplementedError("continuous collision check with a user-defined cost function: the hull check runs "
                                      "on the device right after the device-side selection (built-in cost functions)")
        if generic_cost:
            cost_spec = {"cost_kind": _lib.COST_NONE, "desired_speed": None, "desired_s": None, "desired_d": 0.0, "w_a": 1.0}
        want_all = bool(self._draw_traj_set or generic_cost)

        t0 = time.time()
        level_index = None
        spec = self._spec
        if (spec is not None and dev["kind"] == "grid" and spec["serial"] == dev.get("serial") and dev["level"] in spec["index"]
                and spec["generation"] == eng.plan_generation and tuple(spec["x_0_lon"]) == tuple(dev["x_0_lon"])
                and tuple(spec["x_0_lat"]) == tuple(dev["x_0_lat"])):
            # this level was evaluated ahead, together with the previous one: no new submission
            level_index = spec["index"][dev["level"]]
            res = spec["records"][level_index]
        else:
            self._spec = None
            levels = self._cycle_levels(dev, generic_cost) if dev["kind"] == "grid" else None
            # replanning-size launches check every feasible candidate: the lazy gate's extra barriers cost a small
            # bundle more than the skipped checks save (winner and counters are the same in both modes; the views
            # restore the reference's labels, trajectories._STATUS_TO_LABEL / TrajectorySample.feasibility_label)
            inputs = self._plan_inputs(dev["x_0_lon"], dev["x_0_lat"], cost_spec, want_all,
                                       check_collision=_lib.COLLISION_ALL if levels is not None else None)
            if levels is not None:
                # the whole cycle in one launch: inputs inside the launch, result block written to mapped host memory
                records, chosen = eng.plan_levels(inputs, [lv[1:] for lv in levels])
                records = [_lib.PlanResult.from_buffer_copy(records[j]) for j in range(chosen + 1)]
                level_index = 0
                res = records[0]
                if len(levels) > 1:
                    self._spec = {"serial": dev.get("serial"), "generation": eng.plan_generation, "records": records,
                                  "index": {lv[0]: j for j, lv in enumerate(levels) if j <= chosen},
                                  "x_0_lon": dev["x_0_lon"], "x_0_lat": dev["x_0_lat"]}
            elif dev["kind"] == "grid":
                res = eng.plan_grid(inputs, dev["t"], dev["lon"], dev["d"])
            else:
                cands = trajectory_bundle.trajectories
                skip = None
                if self.config.sampling.longitudinal_mode == 'stopping':
                    # filter_goals_behind (trajectories.py:545-550) as skip flags so that indices stay stable
                    skip = np.array([not (c.trajectory_long.x_0[0] < c.trajectory_long.x_d[0]) for c in cands], dtype=np.uint8)
                cl = np.array([c.trajectory_long.coeffs for c in cands], dtype=np.float64).reshape(-1, 6)
                ct = np.array([c.trajectory_lat.coeffs for c in cands], dtype=np.float64).reshape(-1, 6)
                tl = np.array([_lib.traj_len_of(c.trajectory_long.delta_tau, self.dt) for c in cands], dtype=np.int32)
                dev["candidates"] = cands
                res = eng.plan_list(inputs, cl, ct, tl, skip)
        dev["generation"] = eng.plan_generation
        arrays = dev["arrays"] = _BundleArrays(eng, eng.plan_generation, want_all, level=level_index)
        from_block = False
        if level_index is not None:
            eng.select_level(level_index)           # (no library call when the engine already addresses that level)
            # the chosen level's winner states lie in the mapped result block the kernel wrote
            from_block = getattr(eng, "_cyc_chosen", -1) == level_index
        self.last_result = res
        if log:
            logger.info(f"Kinematic + cost + collision checks took:  \t{time.time() - t0:.7f}s")

        winner = res.winner
        n_collision = res.n_infeasible_collision
        dict.__setitem__(arrays, "winner_key", (float(res.winner_cost), int(winner)) if winner >= 0 else None)
        if generic_cost:
            winner, n_collision = self._select_with_user_cost(trajectory_bundle, arrays)

        self._infeasible_count_kinematics = int(res.n_infeasible_kinematics)
        self._infeasible_count_collision = int(n_collision)
        for constraint in self.config.planning.constraints_to_check:
            self._infeasible_reason_dict[constraint] = int(res.reason_counts[_REASON_INDEX[constraint]])

        if self._draw_traj_set:
            # feasible candidates first, then the kinematically infeasible ones (reference :1121-1128); the views are
            # created when they are looked at (plotting), not 3 000 Python objects per cycle up front
            def order():
                status = np.asarray(arrays["status"])
                return np.concatenate([np.nonzero(np.isin(status, (0, 2, 4)))[0], np.nonzero(status == 1)[0]])

            self.stored_trajectories = _LazyViews(lambda k: self._view(trajectory_bundle, k, arrays), order)

        if winner < 0:
            return None
        best = self._view(trajectory_bundle, winner, arrays)
        best._set_states(eng.cycle_winner_states() if (from_block and not generic_cost) else eng.fetch_states(winner))
        best._label = FeasibilityStatus.FEASIBLE
        if not generic_cost:
            best._cost = float(res.winner_cost)
            best._cost_function = self.cost_function
        return best

    def _select_with_user_cost(self, bundle, arrays):
        """Generic cost plug-in: the device did kinematics, projection and the collision flags of every
        candidate; the user's Python ``evaluate`` ranks the feasible ones (reference :1131-1135 semantics:
        stable sort by cost, first collision-free wins, colliders ranked before it are counted)."""
        status = arrays["status"]
        feasible = [k for k in range(len(status)) if status[k] in (0, 2)]
        views = {k: self._view(bundle, k, arrays) for k in feasible}
        costs = {}
        for k in feasible:
            views[k]._materialise()
            views[k].cost = self.cost_function
            costs[k] = views[k].cost
        arrays["cost"] = np.array([costs.get(k, np.nan) for k in range(len(status))])
        n_collision = 0
        for k in sorted(feasible, key=lambda q: costs[q]):
            if status[k] == 2:
                n_collision += 1
            else:
                return k, n_collision
        return -1, n_collision

    # ------------------------------------------------------------------ host glue around the hot path
    def _compute_initial_states(self, x_0: ReactivePlannerState) -> (np.ndarray, np.ndarray):
        """Cartesian initial state -> curvilinear (lon, lat) initial states (reference :446-512) on the device:
        the pseudo-normal projection over all path segments and the Frenet derivative formulas are one kernel
        (rp_initial_states); errors are raised exactly as the reference raises them."""
        if not self._co:
            return None
        self._sync_frame_tables()
        lon, lat, status = self.engine.initial_states(
            [x_0.position[0], x_0.position[1], x_0.orientation, x_0.velocity, x_0.acceleration, x_0.steering_angle],
            self._low_vel_mode)
        if status[0] == 1:
            logger.critical("Initial state could not be transformed.")
            raise ValueError("Initial state could not be transformed.")
        if status[0] == 2:
            raise Exception("Initial state or reference incorrect! Curvilinear velocity is negative which indicates"
                            "that the ego vehicle is not driving in the same direction as specified by the reference")
        return [float(v) for v in lon[0]], [float(v) for v in lat[0]]

    def _compute_trajectory_pair(self, trajectory: TrajectorySample) -> Tuple[Trajectory, Trajectory, List, List]:
        """Optimal sample -> (Cartesian Trajectory, curvilinear Trajectory, lon list, lat list)
        (reference :514-568)."""
        factor = self.config.planning.factor
        x_0 = self.x_0
        block = trajectory._device_block() if hasattr(trajectory, "_device_block") else None
        device_block = block is not None
        if device_block and _rp_pack is not None and not HAVE_COMMONROAD_IO:
            # the whole packing in one C call (csrc/rp_pack.c): state objects, steering angles, yaw rates, orientation
            # shift, the two curvilinear lists -- the same per-state arithmetic as the reference's loop (:520-556)
            t0 = x_0.time_step
            positions = None if _PACK_MAKES_POSITIONS else list(np.ascontiguousarray(block[0:2].T))
            cart_list, lon_list, lat_list = _rp_pack.pack(ReactivePlannerState, block, positions,
                                                          int(t0), int(factor), self.dt, self.vehicle_params.wheelbase,
                                                          x_0.yaw_rate, x_0.orientation - math.pi, x_0.orientation + math.pi)
            pos_curv = block[7:9].T
            rows = block[2:6]
            curv_traj = _LazyTrajectory(t0, lambda: _curvilinear_states(t0, factor, np.ascontiguousarray(pos_curv), rows[1].tolist(),
                                                                        rows[2].tolist(), rows[0].tolist(), rows[3].tolist()))
            return Trajectory(t0, cart_list), curv_traj, lon_list, lat_list
        ca, cu = trajectory.cartesian, trajectory.curvilinear
        n = len(ca.x)
        theta = ca.theta
        steering = np.arctan2(self.vehicle_params.wheelbase * ca.kappa, 1.0)
        # element-wise the same arithmetic as the reference's per-state loop (:520-556), built once as arrays
        yaw_rate = np.empty(n)
        yaw_rate[0] = x_0.yaw_rate
        yaw_rate[1:] = (theta[1:] - theta[:-1]) / self.dt
        t0 = x_0.time_step
        if device_block:
            # device result: the 14 rows are views of one block -- one transpose / tolist per group instead of one per row
            rows = block.tolist()
            th_l, v_l, a_l, kap_l = rows[2], rows[3], rows[4], rows[5]
            pos_cart = np.ascontiguousarray(block[0:2].T)
            pos_curv = np.ascontiguousarray(block[7:9].T)
            lon_list = [list(r) for r in zip(rows[7], rows[10], rows[11])]
            lat_list = [list(r) for r in zip(rows[8], rows[12], rows[13])]
        else:
            pos_cart = np.stack([ca.x, ca.y], axis=1)
            pos_curv = np.stack([cu.s, cu.d], axis=1)
            th_l, v_l, a_l, kap_l = theta.tolist(), ca.v.tolist(), ca.a.tolist(), ca.kappa.tolist()
            lon_list = np.stack([cu.s, cu.s_dot, cu.s_ddot], axis=1).tolist()
            lat_list = np.stack([cu.d, cu.d_dot, cu.d_ddot], axis=1).tolist()
        if HAVE_COMMONROAD_IO:
            cart_list = [ReactivePlannerState(time_step=t0 + factor * i, position=pos_cart[i], orientation=th, velocity=vel,
                                              acceleration=acc, yaw_rate=yr, steering_angle=st)
                         for i, (th, vel, acc, yr, st) in enumerate(zip(th_l, v_l, a_l, yaw_rate.tolist(), steering.tolist()))]
            cl_list = [CustomState(time_step=t0 + factor * i, position=pos_curv[i], velocity=vel, acceleration=acc,
                                   orientation=th, yaw_rate=kap)
                       for i, (vel, acc, th, kap) in enumerate(zip(v_l, a_l, th_l, kap_l))]
            cart_traj = shift_orientation(Trajectory(t0, cart_list), interval_start=x_0.orientation - np.pi,
                                          interval_end=x_0.orientation + np.pi)
            return cart_traj, Trajectory(t0, cl_list), lon_list, lat_list
        # ---- the package's own plain state classes (no commonroad-io) ----------------------------------------------
        # orientation shift (:565, utility/general.py:49-55) on the list, before the objects exist: whole turns into
        # [x_0.orientation - pi, x_0.orientation + pi]
        lo, hi = x_0.orientation - math.pi, x_0.orientation + math.pi
        th_shift = th_l
        if min(th_l) < lo or max(th_l) > hi:
            th_shift = []
            for th in th_l:
                while th < lo:
                    th += _TWO_PI
                while th > hi:
                    th -= _TWO_PI
                th_shift.append(th)
        new_rs = ReactivePlannerState.__new__
        cart_list = []
        append = cart_list.append
        ts = t0
        for pos, th, vel, acc, yr, st in zip(pos_cart, th_shift, v_l, a_l, yaw_rate.tolist(), steering.tolist()):
            o = new_rs(ReactivePlannerState)
            o.__dict__ = {"time_step": ts, "position": pos, "steering_angle": st, "velocity": vel, "orientation": th,
                          "acceleration": acc, "yaw_rate": yr}
            append(o)
            ts += factor
        # the curvilinear state list (second element of the result; run_planner.py never reads it) is built when read
        curv_traj = _LazyTrajectory(t0, lambda: _curvilinear_states(t0, factor, pos_curv, v_l, a_l, th_l, kap_l))
        return Trajectory(t0, cart_list), curv_traj, lon_list, lat_list

    def plan(self, current_sampling_level: int = None) -> tuple:
        """Plans an optimal trajectory (reference :570-665): sampling levels are escalated until one yields a
        feasible, collision-free candidate."""
        planning_start_time = time.time()
        assert self.x_0 is not None, "<ReactivePlanner.plan(): Planner Cartesian initial state is empty!>"
        assert self._co is not None, "<ReactivePlanner.plan(): No coordinate system given. Call set_reference_path()>"
        if not self.x_0_cl:
            self.x_0_cl = self._compute_initial_states(self.x_0)
        assert self.x_0_cl is not None, "<ReactivePlanner.plan(): Planner curvilinear initial state is empty!>"
        x_0_lon, x_0_lat = self.x_0_cl
        self._low_vel_mode = True if self.x_0.velocity < self.config.planning.low_vel_mode_threshold else False

        log = logger.isEnabledFor(logging.INFO)
        if log:                                    # the f-strings format numpy arrays: skip when nobody listens
            logger.info("=================== Starting Planning Cycle ===================")
            logger.info(f"time_step={self.x_0.time_step} position={self.x_0.position} velocity={self.x_0.velocity} "
                        f"orientation={self.x_0.orientation}")
            logger.info(f"longitudinal state = {x_0_lon}  lateral state = {x_0_lat}")
            logger.info(f"mode: {self.config.sampling.longitudinal_mode}  desired velocity: {self._desired_speed} m/s  "
                        f"desired longitudinal position: {self._desired_lon_position} m")

        optimal_trajectory = None
        bundle = None
        i = 1 if current_sampling_level is None else current_sampling_level
        self._plan_serial += 1
        self._spec = None
        self._escalating = current_sampling_level is None        # the device may evaluate the following levels ahead
        self._level_failed = False
        try:
            while optimal_trajectory is None and i < self.sampling_level:
                bundle = self._create_trajectory_bundle(x_0_lon, x_0_lat, samp_level=i)
                t0 = time.time()
                optimal_trajectory = self._get_optimal_trajectory(bundle)
                if log:
                    logger.info(f"Total checking time: {time.time() - t0:.7f}")
                    logger.info(f"Rejected {self.infeasible_count_kinematics} infeasible trajectories due to kinematics")
                    logger.info(f"Rejected {self.infeasible_count_collision} infeasible trajectories due to collisions")
                if current_sampling_level is not None:
                    break
                self._level_failed = optimal_trajectory is None
                i += 1
        finally:
            self._escalating = False

        if self.x_0.velocity <= 0.05 and \
                (optimal_trajectory is None or optimal_trajectory.cartesian.v[self._standstill_lookahead] <= 0.05):
            logger.info("Planning standstill for the current scenario")
            optimal_trajectory = self._compute_standstill_trajectory()
            if optimal_trajectory is None and current_sampling_level == self.sampling_level:
                logger.warning("Could not find a valid trajectory")
            else:
                self._optimal_cost = optimal_trajectory.cost

        planning_result = self._compute_trajectory_pair(optimal_trajectory) if optimal_trajectory is not None else None
        self._planning_times_list.append(time.time() - planning_start_time)
        if log:
            logger.info(f"Total planning time: {self.planning_times[-1]:.7f}")
        if planning_result is None:
            logger.warning("Planner failed to find an optimal trajectory with given sampling configuration!")
        return planning_result

    def _compute_standstill_trajectory(self) -> TrajectorySample:
        """Artificial standstill sample when the vehicle is already stopped (reference :667-713)."""
        x_0 = self.x_0
        x_0_lon, x_0_lat = self.x_0_cl
        logger.info("Adding standstill trajectory")
        traj_lon = QuarticTrajectory(tau_0=0, delta_tau=self.horizon, x_0=np.asarray(x_0_lon), x_d=np.array([0, 0]))
        traj_lat = QuinticTrajectory(tau_0=0, delta_tau=self.horizon, x_0=np.asarray(x_0_lat),
                                     x_d=np.array([x_0_lat[0], 0, 0]))
        kappa_0 = np.tan(x_0.steering_angle) / self.vehicle_params.wheelbase
        p = TrajectorySample(self.horizon, self.dt, traj_lon, traj_lat)
        a = np.repeat(0.0, self.N)
        a[1] = - self.x_0.velocity / self.dt
        p.cartesian = CartesianSample(np.repeat(x_0.position[0], self.N), np.repeat(x_0.position[1], self.N),
                                      np.repeat(x_0.orientation, self.N), np.repeat(0.0, self.N), a,
                                      np.repeat(kappa_0, self.N), np.repeat(0.0, self.N), current_time_step=self.N)
        co = self._co
        s_idx = np.argmax(co.ref_pos > x_0_lon[0]) - 1
        ref_theta = np.unwrap(co.ref_theta)
        theta_cl = x_0.orientation - interpolate_angle(x_0_lon[0], co.ref_pos[s_idx], co.ref_pos[s_idx + 1],
                                                       ref_theta[s_idx], ref_theta[s_idx + 1])
        p.curvilinear = CurviLinearSample(np.repeat(x_0_lon[0], self.N), np.repeat(x_0_lat[0], self.N),
                                          np.repeat(theta_cl, self.N), dd=np.repeat(x_0_lat[1], self.N),
                                          ddd=np.repeat(x_0_lat[2], self.N), ss=np.repeat(x_0_lon[1], self.N),
                                          sss=np.repeat(x_0_lon[2], self.N), current_time_step=self.N)
        return p

    def convert_state_list_to_commonroad_object(self, state_list: List[ReactivePlannerState], obstacle_id: int = 42):
        """Planned state list -> CommonRoad DynamicObstacle of the ego vehicle (reference :1138-1159; needs
        commonroad-io)."""
        from commonroad.geometry.shape import Rectangle
        from commonroad.prediction.prediction import TrajectoryPrediction
        from commonroad.scenario.obstacle import DynamicObstacle, ObstacleType
        from commonroad.scenario.state import InitialState
        from commonroad.scenario.trajectory import Trajectory as CRTrajectory
        shifted = [st.shift_positions_to_center(self.vehicle_params.wb_rear_axle) for st in state_list]
        trajectory = CRTrajectory(initial_time_step=shifted[0].time_step, state_list=shifted)
        shape = Rectangle(self.vehicle_params.length, self.vehicle_params.width)
        init_state = trajectory.state_list[0].convert_state_to_state(InitialState())
        return DynamicObstacle(obstacle_id, ObstacleType.CAR, shape, init_state, TrajectoryPrediction(trajectory, shape))
