"""Sampling spaces -- same classes as the reference's ``commonroad_rp/sampling.py``.

The per-level sample tables are still Python ``set`` objects of ``np.float64`` built with the very same
expressions as the reference (:80-118): candidate enumeration order IS the iteration order of those
sets, so it is obtained by iterating them on the host every cycle and shipped to
the device as three small ordered arrays (``sample_grid``) -- "sample sets become device arrays".
``generate_trajectories_at_level`` keeps its signature and return type for user code and custom
pipelines; it solves all polynomial coefficients of the level in ONE device launch.
"""
from abc import ABC, abstractmethod
from typing import Dict, List, Optional

import numpy as np

from commonroad_rp_b200.polynomial_trajectory import (QUARTIC, QUINTIC, QuarticTrajectory, QuinticTrajectory,
                                                       solve_batch)
from commonroad_rp_b200.trajectories import TrajectorySample
from commonroad_rp_b200.utility.config import ReactivePlannerConfiguration

try:  # optional dependency of CorridorSampling, exactly as in the reference (:17-25)
    from commonroad_reach.data_structure.reach.driving_corridor import DrivingCorridor
    import commonroad_reach.utility.reach_operation as util_reach_operation
    cr_reach_installed = True
except ImportError:
    DrivingCorridor = None
    util_reach_operation = None
    cr_reach_installed = False


class Sampling(ABC):
    """One sampled dimension: a set of values per sampling level, built once by ``_sample`` (reference :28-69)."""

    def __init__(self, low: float, up: float, num_sampling_levels: int):
        if not np.greater_equal(up, low):
            raise AssertionError('<Sampling>: Upper sampling bound is not greater than lower bound! '
                                 'up = {} , low = {}'.format(up, low))
        if not (isinstance(num_sampling_levels, int) and num_sampling_levels > 0):
            raise AssertionError('<Sampling: number of samples must be positive integer>')
        self.low, self.up = low, up
        self._n_samples = num_sampling_levels
        self._dict_level_to_sample_set: Dict[int, set] = {}
        self._sample()

    @abstractmethod
    def _sample(self):
        """fill ``_dict_level_to_sample_set[level]`` for every level"""

    @property
    def num_sampling_levels(self) -> int:
        return self._n_samples

    def samples_at_level(self, sampling_level: int = 0) -> set:
        if not 0 <= sampling_level < self._n_samples:
            raise AssertionError('<Sampling>: Provided sampling level is incorrect! stage = {}'.format(sampling_level))
        return self._dict_level_to_sample_set[sampling_level]


def _nested_linspace_sets(low, up, levels):
    """3, 5, 9, 17, ... equidistant samples (reference :80-84, :95-99)."""
    out, n = {}, 3
    for level in range(levels):
        out[level] = set(np.linspace(low, up, n))
        n = (n * 2) - 1
    return out


class VelocitySampling(Sampling):
    def _sample(self):
        self._dict_level_to_sample_set = _nested_linspace_sets(self.low, self.up, self.num_sampling_levels)


class PositionSampling(Sampling):
    def _sample(self):
        self._dict_level_to_sample_set = _nested_linspace_sets(self.low, self.up, self.num_sampling_levels)


class TimeSampling(Sampling):
    """Time-horizon samples, denser with every level (reference :102-118)."""

    def __init__(self, low: float, up: float, num_sampling_levels: int, dt: float):
        self.dT = dt
        assert low >= 2 * self.dT, "<TimeSampling: lower bound of time sampling must be greater-equal than the given" \
                                   "time step>"
        super().__init__(low, up, num_sampling_levels)

    def _sample(self):
        for level in range(self.num_sampling_levels):
            step_size = int((1 / (level + 1)) / self.dT)
            samp = set(np.arange(self.low, round(self.up + self.dT, 2), step_size * self.dT))
            samp.discard(round(self.up + self.dT, 2))
            self._dict_level_to_sample_set[level] = samp


class SamplingSpace(ABC):
    """Base of all sampling spaces: holds the per-dimension ``Sampling`` objects (``samples_t``,
    ``samples_d``, ``samples_v``; subclasses may add more) and produces the trajectory set of a level
    (reference :121-175)."""

    def __init__(self, num_sampling_levels: int):
        self._num_sampling_levels = num_sampling_levels
        self._samples_t: Optional[TimeSampling] = None
        self._samples_d: Optional[PositionSampling] = None
        self._samples_v: Optional[VelocitySampling] = None
        self._samples_s: Optional[PositionSampling] = None

    @property
    def num_sampling_levels(self) -> int:
        return self._num_sampling_levels

    # plain attribute-backed properties; CorridorSampling overrides the d / v setters
    samples_t = property(lambda self: self._samples_t, lambda self, value: setattr(self, "_samples_t", value))
    samples_d = property(lambda self: self._samples_d, lambda self, value: setattr(self, "_samples_d", value))
    samples_v = property(lambda self: self._samples_v, lambda self, value: setattr(self, "_samples_v", value))

    @abstractmethod
    def generate_trajectories_at_level(self, level_sampling: int, x_0_lon: np.ndarray, x_0_lat: np.ndarray,
                                       longitudinal_mode: str, low_vel_mode: bool) -> List[TrajectorySample]:
        """List of TrajectorySample of the given level, in enumeration order."""


class FixedIntervalSampling(SamplingSpace):
    """Fixed-interval sampling in the t, v (or s) and d domains (reference :178-270)."""

    def __init__(self, config: ReactivePlannerConfiguration):
        num_sampling_levels = config.sampling.num_sampling_levels
        super().__init__(num_sampling_levels)
        cs = config.sampling
        self.dt = config.planning.dt
        self.horizon = config.planning.dt * config.planning.time_steps_computation
        self._longitudinal_mode = None
        self.samples_t = TimeSampling(cs.t_min, self.horizon, num_sampling_levels, self.dt)
        self.samples_d = PositionSampling(cs.d_min, cs.d_max, num_sampling_levels)
        self.samples_v = VelocitySampling(cs.v_min, cs.v_max, num_sampling_levels)
        self.samples_s = PositionSampling(cs.s_min, cs.s_max, num_sampling_levels)

    samples_s = property(lambda self: self._samples_s, lambda self, value: setattr(self, "_samples_s", value))

    def _get_lon_samples(self, level_sampling):
        if self._longitudinal_mode == "velocity_keeping":
            return self.samples_v.samples_at_level(level_sampling)
        elif self._longitudinal_mode == "stopping":
            return self.samples_s.samples_at_level(level_sampling)
        raise AttributeError(f"<FixedIntervalSampling>: specified longitudinal mode {self._longitudinal_mode} is"
                             f"invalid.")

    def sample_grid(self, level_sampling: int, x_0_lat, longitudinal_mode: str):
        """The three ORDERED sample lists of a level -- iteration order of the reference's sets
        (reference :218, :220, :226) -- as float64 arrays for ``rp_plan_grid``.  Enumeration index of a
        candidate = (i_t * n_lon + i_lon) * n_d + i_d."""
        self._longitudinal_mode = longitudinal_mode
        t_set = self.samples_t.samples_at_level(level_sampling)
        lon_set = self._get_lon_samples(level_sampling)
        # the t / lon arrays are re-read whenever a set object (or its size) changed; they are shared between cycles:
        # callers must not write to them (read-only arrays)
        memo = self.__dict__.get("_grid_memo")
        if memo is None:
            memo = self._grid_memo = {}
        ht = memo.get(level_sampling)                           # (t and lon separately: set_desired_velocity replaces
        if ht is None or ht[0] is not t_set or ht[1] != len(t_set):   # the velocity samples every cycle, the t samples stay)
            t = np.fromiter(t_set, dtype=np.float64)
            t.flags.writeable = False
            ht = memo[level_sampling] = (t_set, len(t_set), t)
        key = (level_sampling, longitudinal_mode)
        hl = memo.get(key)
        if hl is None or hl[0] is not lon_set or hl[1] != len(lon_set):
            lon = np.fromiter(lon_set, dtype=np.float64)
            lon.flags.writeable = False
            hl = memo[key] = (lon_set, len(lon_set), lon)
        # the union builds a new set whose order depends on d0: iterated afresh every cycle
        d = np.fromiter(self.samples_d.samples_at_level(level_sampling).union((x_0_lat[0],)), dtype=np.float64)
        return ht[2], hl[2], d

    def generate_trajectories_at_level(self, level_sampling: int, x_0_lon: np.ndarray, x_0_lat: np.ndarray,
                                       longitudinal_mode: str, low_vel_mode: bool) -> List[TrajectorySample]:
        """List of TrajectorySample in enumeration order (reference :202-242); all coefficient solves of the
        level run in one device launch."""
        t, lon, d = self.sample_grid(level_sampling, x_0_lat, longitudinal_mode)
        x0_lon = np.asarray(x_0_lon, dtype=np.float64)
        x0_lat = np.asarray(x_0_lat, dtype=np.float64)
        quartic = longitudinal_mode == "velocity_keeping"
        n_t, n_lon, n_d = len(t), len(lon), len(d)
        # longitudinal systems: (t, lon)
        tt, ll = np.meshgrid(t, lon, indexing="ij")
        n_l = tt.size
        xd_l = np.zeros((n_l, 3))
        xd_l[:, 0] = ll.ravel()
        c_lon = solve_batch(np.full(n_l, QUARTIC if quartic else QUINTIC, dtype=np.int32),
                            np.broadcast_to(x0_lon, (n_l, 3)), xd_l, tt.ravel()) if n_l else np.zeros((0, 6))
        lon_trajs = []
        for q in range(n_l):
            if quartic:
                lon_trajs.append(QuarticTrajectory(tau_0=0, delta_tau=tt.ravel()[q], x_0=x0_lon.copy(),
                                                   x_d=np.array([ll.ravel()[q], 0.0]), coeffs=c_lon[q]))
            else:
                lon_trajs.append(QuinticTrajectory(tau_0=0, delta_tau=tt.ravel()[q], x_0=x0_lon.copy(),
                                                   x_d=np.array([ll.ravel()[q], 0.0, 0.0]), coeffs=c_lon[q]))
        # lateral systems: (t, d) at high velocity, (t, lon, d) when sampling over the travelled distance
        if low_vel_mode:
            tau_lat = np.empty((n_t, n_lon))
            for q, traj in enumerate(lon_trajs):
                s_goal = traj.evaluate_state_at_tau(traj.delta_tau)[0] - x0_lon[0]
                tau_lat.ravel()[q] = traj.delta_tau if s_goal <= 0 else s_goal
            tau_all = np.repeat(tau_lat.ravel(), n_d)
            d_all = np.tile(d, n_t * n_lon)
        else:
            tau_all = np.repeat(t, n_d)
            d_all = np.tile(d, n_t)
        n_s = tau_all.size
        xd_t = np.zeros((n_s, 3))
        xd_t[:, 0] = d_all
        c_lat = solve_batch(np.full(n_s, QUINTIC, dtype=np.int32), np.broadcast_to(x0_lat, (n_s, 3)), xd_t,
                            tau_all) if n_s else np.zeros((0, 6))
        lat_trajs = [QuinticTrajectory(tau_0=0, delta_tau=tau_all[q], x_0=x0_lat.copy(),
                                       x_d=np.array([d_all[q], 0.0, 0.0]), coeffs=c_lat[q]) for q in range(n_s)]
        out = []
        for it in range(n_t):
            for il in range(n_lon):
                tl = lon_trajs[it * n_lon + il]
                for idd in range(n_d):
                    lat = lat_trajs[(it * n_lon + il) * n_d + idd] if low_vel_mode else lat_trajs[it * n_d + idd]
                    out.append(TrajectorySample(self.horizon, self.dt, tl, lat))
        return out


class IntervalCorridor(dict):
    """A driving corridor as plain interval arrays: ``time index -> (n, 6) array`` whose rows are the curvilinear
    bounding intervals (s_lo, s_hi, d_lo, d_hi, v_lo, v_hi) of the reach-set nodes at that time -- what
    commonroad_reach's ReachNode exposes as p_lon_min/max, p_lat_min/max, v_lon_min/max.  ``CorridorSampling`` accepts
    it in place of a commonroad_reach ``DrivingCorridor`` (which needs CommonRoad-Reach installed); the four reach
    operations the reference calls (sampling.py:312, :365-371) are then the interval versions below."""

    def __init__(self, intervals):
        super().__init__({int(k): np.asarray(v, dtype=np.float64).reshape(-1, 6) for k, v in dict(intervals).items()})


class _IntervalReachOperations:
    """commonroad_reach.utility.reach_operation on IntervalCorridor node arrays"""

    @staticmethod
    def lon_velocity_interval_connected_set(nodes):
        return (float(np.min(nodes[:, 4])), float(np.max(nodes[:, 5]))) if len(nodes) else (0.0, 0.0)

    @staticmethod
    def determine_overlapping_nodes_with_lon_pos(nodes, lon_pos):
        return nodes[(nodes[:, 0] <= lon_pos) & (lon_pos <= nodes[:, 1])]

    @staticmethod
    def determine_connected_components(nodes):
        """nodes whose lateral intervals overlap (transitively) form one connected set; ordered by lower bound"""
        nodes = np.asarray(nodes, dtype=np.float64).reshape(-1, 6)
        if len(nodes) == 0:
            return []
        nodes = nodes[np.argsort(nodes[:, 2], kind="stable")]
        groups, start, hi = [], 0, nodes[0, 3]
        for k in range(1, len(nodes)):
            if nodes[k, 2] > hi:
                groups.append(nodes[start:k])
                start, hi = k, nodes[k, 3]
            else:
                hi = max(hi, nodes[k, 3])
        groups.append(nodes[start:])
        return groups

    @staticmethod
    def lat_interval_connected_set(nodes):
        return float(np.min(nodes[:, 2])), float(np.max(nodes[:, 3]))


class CorridorSampling(SamplingSpace):
    """Adaptive sampling inside a precomputed collision-free driving corridor (reference :273-397): a
    commonroad_reach ``DrivingCorridor`` (needs CommonRoad-Reach, like the reference) or an ``IntervalCorridor`` of plain
    (s, d, v) interval arrays.  Candidates reach the GPU through the generic list form (``rp_plan_list``)."""

    def __init__(self, config: ReactivePlannerConfiguration):
        num_sampling_levels = config.sampling.num_sampling_levels
        super().__init__(num_sampling_levels)
        self.dt = config.planning.dt
        self.horizon = config.planning.dt * config.planning.time_steps_computation
        self.samples_t = TimeSampling(config.sampling.t_min, self.horizon, num_sampling_levels, self.dt)
        self._corridor = None
        self._velocity_constraints: Dict = dict()
        self._dict_level_to_num_samples: Dict[int, int] = dict()
        self.set_dict_number_of_samples()

    @property
    def driving_corridor(self):
        return self._corridor

    @driving_corridor.setter
    def driving_corridor(self, corridor):
        if isinstance(corridor, IntervalCorridor):
            self._ops = _IntervalReachOperations
        elif cr_reach_installed:
            self._ops = util_reach_operation
        else:
            raise ImportError("<CorridorSampling>: Please install CommonRoad-Reach to use adaptive corridor sampling "
                              "(or pass an IntervalCorridor of plain interval arrays)!")
        self._corridor = corridor
        self._velocity_constraints = dict()
        for time_idx, connected_reach_set in self._corridor.items():
            lo, hi = self._ops.lon_velocity_interval_connected_set(connected_reach_set)[:2]
            self._velocity_constraints[time_idx] = [lo, hi]

    @SamplingSpace.samples_d.setter
    def samples_d(self, pos_sampling: PositionSampling):
        self._d_min = pos_sampling.low
        self._d_max = pos_sampling.up

    @SamplingSpace.samples_v.setter
    def samples_v(self, vel_sampling: VelocitySampling):
        self._v_min = vel_sampling.low
        self._v_max = vel_sampling.up

    def set_dict_number_of_samples(self, n_min: int = 3, dict_level_to_num_samples: dict = None):
        if dict_level_to_num_samples is not None:
            for level in range(self.num_sampling_levels):
                assert level in dict_level_to_num_samples.keys(), \
                    f"<SamplingSpace.set_dict_number_of_samples()>: input dictionary does not contain sampling level:{level}"
        else:
            n = n_min
            for level in range(self.num_sampling_levels):
                self._dict_level_to_num_samples[level] = n
                n = (n * 2) - 1

    def generate_trajectories_at_level(self, level_sampling: int, x_0_lon: np.ndarray, x_0_lat: np.ndarray,
                                       longitudinal_mode: str, low_vel_mode: bool) -> List[TrajectorySample]:
        if self._corridor is None:
            raise AttributeError("<CorridorSampling>: Please set a driving corridor.")
        num_samples = self._dict_level_to_num_samples[level_sampling]
        x0_lon = np.asarray(x_0_lon, dtype=np.float64)
        x0_lat = np.asarray(x_0_lat, dtype=np.float64)
        # pass 1: longitudinal quartics for every (t, v) pair, one batched solve
        pairs = []
        for t in self.samples_t.samples_at_level(level_sampling):
            time_step = round(t / self.dt) + min(self._corridor.keys())
            low, up = self._velocity_constraints[time_step]
            for v in set(np.linspace(low, up, num_samples)):
                pairs.append((t, v, time_step))
        if not pairs:
            return []
        xd = np.zeros((len(pairs), 3))
        xd[:, 0] = [p[1] for p in pairs]
        c_lon = solve_batch(np.full(len(pairs), QUARTIC, dtype=np.int32), np.broadcast_to(x0_lon, (len(pairs), 3)), xd,
                            np.array([p[0] for p in pairs]))
        # pass 2: lateral samples from the corridor's connected sets, one batched solve
        lat_jobs = []
        lon_trajs = []
        for q, (t, v, time_step) in enumerate(pairs):
            traj_lon = QuarticTrajectory(tau_0=0, delta_tau=t, x_0=x0_lon.copy(), x_d=np.array([v, 0]), coeffs=c_lon[q])
            lon_trajs.append(traj_lon)
            end_pos_lon = traj_lon.calc_position(t, t ** 2, t ** 3, t ** 4, t ** 5)
            overlap = self._ops.determine_overlapping_nodes_with_lon_pos(self._corridor[time_step], end_pos_lon)
            if len(list(overlap)) == 0:
                continue
            for lat_con_set in self._ops.determine_connected_components(list(overlap)):
                lo, hi = self._ops.lat_interval_connected_set(lat_con_set)[:2]
                d_samples = set(np.linspace(lo, hi, num_samples))
                if lo < 0 < hi:
                    d_samples = d_samples.union({0})
                for d in d_samples:
                    lat_jobs.append((q, t, d))
        if not lat_jobs:
            return []
        xd = np.zeros((len(lat_jobs), 3))
        xd[:, 0] = [j[2] for j in lat_jobs]
        c_lat = solve_batch(np.full(len(lat_jobs), QUINTIC, dtype=np.int32),
                            np.broadcast_to(x0_lat, (len(lat_jobs), 3)), xd, np.array([j[1] for j in lat_jobs]))
        out = []
        for r, (q, t, d) in enumerate(lat_jobs):
            lat = QuinticTrajectory(tau_0=0, delta_tau=t, x_0=x0_lat.copy(), x_d=np.array([d, 0.0, 0.0]), coeffs=c_lat[r])
            out.append(TrajectorySample(self.horizon, self.dt, lon_trajs[q], lat))
        return out


def sampling_space_factory(config: ReactivePlannerConfiguration):
    """Factory function to select the SamplingSpace class (reference :400-408; the reference forgets to
    raise for an invalid method -- here it raises)."""
    sampling_method = config.sampling.sampling_method
    if sampling_method == 1:
        return FixedIntervalSampling(config)
    elif sampling_method == 2:
        return CorridorSampling(config)
    raise ValueError("Invalid sampling method specified")
