"""Trajectory containers -- the public classes of the reference's ``commonroad_rp/trajectories.py``
(``FeasibilityStatus, CartesianSample, CurviLinearSample, TrajectorySample, TrajectoryBundle``) with
the same attribute names, now backed by device results:

* a ``TrajectorySample`` produced by the GPU planner is a lazy VIEW: ``cartesian`` / ``curvilinear`` /
  ``cost`` / ``feasibility_label`` pull the candidate's 14 x (N+1) state block (``rp_fetch_states``) and
  verdict from the engine on first access, so a 131 072-candidate bundle never materialises Python objects
  unless the caller touches them;
* ``TrajectoryBundle.sort()`` on a device-backed bundle orders by the device costs; the planner itself
  never sorts -- it takes the feasible arg-min (csrc/rp_kernels.cuh ``finalize_kernel``), which is the
  element a stable ascending sort would visit first (reference :502-510, reactive_planner.py:1031).

Samples built by hand (user code, custom sampling spaces) behave exactly like the reference's.
"""
import math
from abc import ABC, abstractmethod
from enum import Enum
from typing import List, Optional, Union

import numpy as np

from commonroad_rp_b200.polynomial_trajectory import PolynomialTrajectory


class FeasibilityStatus(Enum):
    """Feasibility label of a TrajectorySample after checking (reference :18-22)."""
    FEASIBLE = 'feasible'
    INFEASIBLE_KINEMATIC = 'infeasible_kinematic'
    INFEASIBLE_COLLISION = "infeasible_collision"


class Sample(ABC):
    """A trajectory sample in some coordinate system."""

    def __init__(self, current_time_step: int):
        self.current_time_step = current_time_step

    @property
    def current_time_step(self):
        return self._current_time_step

    @current_time_step.setter
    def current_time_step(self, curr_time_step):
        self._current_time_step = curr_time_step

    @abstractmethod
    def length(self) -> int:
        pass

    @abstractmethod
    def enlarge(self, dt: float):
        pass


class CartesianSample(Sample):
    """x, y, theta, v, a, kappa, kappa_dot over the horizon (reference :61-197)."""

    def __init__(self, x: np.ndarray, y: np.ndarray, theta: np.ndarray, v: np.ndarray, a: np.ndarray,
                 kappa: np.ndarray, kappa_dot: np.ndarray, current_time_step: int):
        super().__init__(current_time_step)
        self.x = x
        self.y = y
        self.theta = theta
        self.v = v
        self.a = a
        self.kappa = kappa
        self.kappa_dot = kappa_dot

    def length(self) -> int:
        return len(self.x)

    def enlarge(self, dt: float):
        """Constant-acceleration, constant-heading extension up to the array length (reference :168-197).
        Device-produced samples arrive already extended; this serves hand-built samples."""
        last = self.current_time_step - 1
        steps = self.length() - self.current_time_step
        tau = np.arange(1, steps + 1, 1) * dt
        self.a[self.current_time_step:] = self.a[last]
        v_ext = self.v[last] + tau * self.a[-1]
        v_ext = v_ext * np.greater_equal(v_ext, 0)
        self.v[self.current_time_step:] = v_ext
        for arr in (self.theta, self.kappa, self.kappa_dot):
            arr[self.current_time_step:] = arr[last]
        self.x[self.current_time_step:] = self.x[last] + np.cumsum(dt * v_ext * math.cos(self.theta[last]))
        self.y[self.current_time_step:] = self.y[last] + np.cumsum(dt * v_ext * math.sin(self.theta[last]))
        self.current_time_step = self.length()


class CurviLinearSample(Sample):
    """s, d, theta, s_dot, s_ddot, d_dot, d_ddot over the horizon (reference :200-332)."""

    def __init__(self, s: np.ndarray, d: np.ndarray, theta: np.ndarray, current_time_step: int, dd=None, ddd=None,
                 ss=None, sss=None):
        super().__init__(current_time_step)
        self.s = s
        self.d = d
        self.theta = theta
        self.d_dot = dd
        self.d_ddot = ddd
        self.s_dot = ss
        self.s_ddot = sss

    def length(self) -> int:
        return len(self.s)

    def enlarge(self, dt: float):
        """Extension up to the array length (reference :302-332; the velocity extrapolation there reads the
        still-zero LAST acceleration entry, which is kept)."""
        last = self._current_time_step - 1
        steps = self.length() - self._current_time_step
        tau = np.arange(1, (steps + 1), 1) * dt
        s_dot_ext = self.s_dot[last] + tau * self.s_ddot[-1]
        self.s_dot[self.current_time_step:] = s_dot_ext * np.greater_equal(s_dot_ext, 0)
        self.d_dot[self.current_time_step:] = self.d_dot[last] + tau * self.d_ddot[-1]
        self.s_ddot[self.current_time_step:] = self.s_ddot[last]
        self.d_ddot[self.current_time_step:] = self.d_ddot[last]
        self.theta[self.current_time_step:] = self.theta[last]
        self.s[self.current_time_step:] = self.s[last] + tau * self.s_dot[last]
        self.d[self.current_time_step:] = self.d[last] + tau * self.d_dot[last]
        self.current_time_step = self.length()


class _DeviceBacking:
    """Where a lazy TrajectorySample finds its results: the engine that evaluated the bundle, the
    candidate's enumeration index and the per-bundle verdict arrays (shared by all views of a bundle)."""
    __slots__ = ("engine", "index", "bundle_arrays", "generation")

    def __init__(self, engine, index, bundle_arrays, generation):
        self.engine = engine
        self.index = index
        self.bundle_arrays = bundle_arrays
        self.generation = generation


_STATUS_TO_LABEL = {0: FeasibilityStatus.FEASIBLE, 1: FeasibilityStatus.INFEASIBLE_KINEMATIC,
                    2: FeasibilityStatus.INFEASIBLE_COLLISION, 3: None,
                    4: FeasibilityStatus.FEASIBLE}     # lazy collision pass never visited it (as in the reference)


class TrajectorySample(Sample):
    """Longitudinal + lateral polynomial with the evaluated Cartesian / curvilinear samples
    (reference :335-463)."""

    def __init__(self, horizon: float, dt: float, trajectory_long: PolynomialTrajectory,
                 trajectory_lat: PolynomialTrajectory):
        self.horizon = horizon
        self.dt = dt
        assert isinstance(trajectory_long, PolynomialTrajectory), \
            '<TrajectorySample/init>: Provided longitudinal trajectory is not valid! trajectory = {}'.format(trajectory_long)
        assert isinstance(trajectory_lat, PolynomialTrajectory), \
            '<TrajectorySample/init>: Provided lateral trajectory is not valid! trajectory = {}'.format(trajectory_lat)
        self._trajectory_long = trajectory_long
        self._trajectory_lat = trajectory_lat
        self._cost = 0
        self._cost_function = None
        self._cartesian: Optional[CartesianSample] = None
        self._curvilinear: Optional[CurviLinearSample] = None
        self._ext_cartesian = None
        self._ext_curvilinear = None
        self._label: Optional[FeasibilityStatus] = None
        self._backing: Optional[_DeviceBacking] = None

    # ---- device view plumbing ----
    def _attach(self, backing: _DeviceBacking):
        self._backing = backing
        return self

    def _materialise(self):
        """Pull this candidate's state block from the device (once)."""
        b = self._backing
        if self.__dict__.get("_pending_block") is not None:
            self._build_from_block()
        if b is None or self._cartesian is not None:
            return
        if b.engine.plan_generation == b.generation and hasattr(b.bundle_arrays, "select"):
            b.bundle_arrays.select()          # cycle launches hold several levels: address this bundle's one
        st = b.engine.fetch_states(b.index) if b.engine.plan_generation == b.generation else None
        if st is None:
            raise RuntimeError("<TrajectorySample>: the device bundle this sample belongs to has been replaced by a "
                               "newer plan() call; access states before re-planning or keep a deepcopy")
        n = st.shape[1]
        self._cartesian = CartesianSample(st[0], st[1], st[2], st[3], st[4], st[5], st[6], current_time_step=n)
        self._curvilinear = CurviLinearSample(st[7], st[8], st[9], current_time_step=n, ss=st[10], sss=st[11],
                                              dd=st[12], ddd=st[13])

    def _set_states(self, st: np.ndarray):
        """The candidate's 14 x (N + 1) device state block; the two sample objects over its rows are built when
        ``cartesian`` / ``curvilinear`` are first read (plan()'s output packing reads the block itself)."""
        self._pending_block = st
        self._state_block = None
        self._cartesian = self._curvilinear = None

    def _build_from_block(self):
        st = self.__dict__.get("_pending_block")
        if st is None:
            return
        self._pending_block = None
        n = st.shape[1]
        r = tuple(st)                     # the 14 row views handed to the two samples
        self._state_block = (st, r)       # output packing reads the block whole while the samples still hold these rows
        self._cartesian = CartesianSample(r[0], r[1], r[2], r[3], r[4], r[5], r[6], current_time_step=n)
        self._curvilinear = CurviLinearSample(r[7], r[8], r[9], current_time_step=n, ss=r[10], sss=r[11],
                                              dd=r[12], ddd=r[13])

    def _device_block(self):
        """The device state block while nobody replaced or re-bound the samples' rows, else None."""
        st = self.__dict__.get("_pending_block")
        if st is not None:
            return st
        return self._state_block[0] if self._rows_untouched() else None

    def _rows_untouched(self) -> bool:
        """True while cartesian / curvilinear still hold exactly the row views of the device block."""
        if self.__dict__.get("_pending_block") is not None:
            return True
        blk = getattr(self, "_state_block", None)
        if blk is None or self._cartesian is None or self._curvilinear is None:
            return False
        r, ca, cu = blk[1], self._cartesian, self._curvilinear
        return (ca.x is r[0] and ca.y is r[1] and ca.theta is r[2] and ca.v is r[3] and ca.a is r[4] and ca.kappa is r[5]
                and cu.s is r[7] and cu.d is r[8] and cu.s_dot is r[10] and cu.s_ddot is r[11] and cu.d_dot is r[12]
                and cu.d_ddot is r[13])

    def __deepcopy__(self, memo):
        import copy
        self._materialise_if_available()
        self.trajectory_long, self.trajectory_lat          # (lazy views build their polynomial objects now)
        new = TrajectorySample.__new__(TrajectorySample)
        for k, v in self.__dict__.items():
            setattr(new, k, None if k in ("_backing", "_state_block", "_poly_factory", "_pending_block") else copy.deepcopy(v, memo))
        if self._backing is not None:
            new._cost = self.cost
            new._label = self.feasibility_label
        return new

    def _materialise_if_available(self):
        b = self._backing
        if self.__dict__.get("_pending_block") is not None:
            self._build_from_block()
        if b is not None and self._cartesian is None and b.engine.plan_generation == b.generation:
            status = int(b.bundle_arrays["status"][b.index])
            if status in (0, 2, 4) or b.bundle_arrays.get("all_states", False):
                self._materialise()

    # ---- reference API ----
    @property
    def trajectory_long(self) -> PolynomialTrajectory:
        return self._trajectory_long

    @trajectory_long.setter
    def trajectory_long(self, trajectory_long):
        pass

    @property
    def trajectory_lat(self) -> PolynomialTrajectory:
        return self._trajectory_lat

    @trajectory_lat.setter
    def trajectory_lat(self, trajectory_lat):
        pass

    @property
    def cost(self) -> float:
        if self._backing is not None and self._cost_function is None:
            c = float(self._backing.bundle_arrays["cost"][self._backing.index])
            return c if not math.isnan(c) else 0
        return self._cost

    @cost.setter
    def cost(self, cost_function):
        """Evaluate and store the cost with the given cost function (reference :397-404)."""
        self._cost = cost_function.evaluate(self)
        self._cost_function = cost_function

    @property
    def curvilinear(self) -> CurviLinearSample:
        if self._curvilinear is None:
            if self.__dict__.get("_pending_block") is not None:
                self._build_from_block()
            elif self._backing is not None:
                self._materialise_if_available()
        return self._curvilinear

    @curvilinear.setter
    def curvilinear(self, curvilinear: CurviLinearSample):
        assert isinstance(curvilinear, CurviLinearSample)
        self._curvilinear = curvilinear

    @property
    def cartesian(self) -> CartesianSample:
        if self._cartesian is None:
            if self.__dict__.get("_pending_block") is not None:
                self._build_from_block()
            elif self._backing is not None:
                self._materialise_if_available()
        return self._cartesian

    @cartesian.setter
    def cartesian(self, cartesian: CartesianSample):
        assert isinstance(cartesian, CartesianSample)
        self._cartesian = cartesian

    @property
    def feasibility_label(self):
        if self._label is None and self._backing is not None:
            arr, k = self._backing.bundle_arrays, self._backing.index
            status = int(arr["status"][k])
            if status == 2:
                # the device may have collision-checked every candidate; the reference's lazy pass (:1031-1063) only
                # labels the colliders it met BEFORE the winner in (cost, index) order -- the others stay FEASIBLE
                key = arr.get("winner_key") if hasattr(arr, "get") else None
                if key is not None and (float(arr["cost"][k]), k) > key:
                    return FeasibilityStatus.FEASIBLE
            return _STATUS_TO_LABEL[status]
        return self._label

    @feasibility_label.setter
    def feasibility_label(self, feasbility_status: FeasibilityStatus):
        self._label = feasbility_status

    def length(self) -> int:
        return self.cartesian.length()

    def enlarge(self, dt: float):
        self.cartesian.enlarge(dt)
        self.curvilinear.enlarge(dt)


class DeviceTrajectorySample(TrajectorySample):
    """A view on one candidate of a device-evaluated grid bundle whose two polynomial OBJECTS are built on first access
    (``poly_factory() -> (trajectory_long, trajectory_lat)``): the planner creates one such view per cycle for the
    winner, and most callers only read its sampled states."""

    def __init__(self, horizon: float, dt: float, poly_factory):
        self.horizon = horizon
        self.dt = dt
        self._poly_factory = poly_factory
        self._trajectory_long = None
        self._trajectory_lat = None
        self._cost = 0
        self._cost_function = None
        self._cartesian = None
        self._curvilinear = None
        self._ext_cartesian = None
        self._ext_curvilinear = None
        self._label = None
        self._backing = None

    def _polys(self):
        if self._trajectory_long is None and self._poly_factory is not None:
            self._trajectory_long, self._trajectory_lat = self._poly_factory()
            self._poly_factory = None

    @property
    def trajectory_long(self) -> PolynomialTrajectory:
        self._polys()
        return self._trajectory_long

    @trajectory_long.setter
    def trajectory_long(self, trajectory_long):
        pass

    @property
    def trajectory_lat(self) -> PolynomialTrajectory:
        self._polys()
        return self._trajectory_lat

    @trajectory_lat.setter
    def trajectory_lat(self, trajectory_lat):
        pass


class TrajectoryBundle:
    """A collection of trajectory samples (reference :466-558).  ``trajectories`` may be given lazily as
    a callable returning the list (device-backed bundles)."""

    def __init__(self, trajectories, cost_function):
        if callable(trajectories):
            self._lazy = trajectories
            self._trajectory_bundle = None
        else:
            assert isinstance(trajectories, list) and all([isinstance(t, TrajectorySample) for t in trajectories]), \
                '<TrajectoryBundle/init>: Provided list of trajectory samples is not valid! List = {}'.format(trajectories)
            self._lazy = None
            self._trajectory_bundle = trajectories
        self._cost_function = cost_function
        self._is_sorted = False
        self.device = None          # filled by the planner: grid description / verdict arrays of the device bundle

    @property
    def trajectories(self) -> List[TrajectorySample]:
        if self._trajectory_bundle is None and self._lazy is not None:
            self._trajectory_bundle = self._lazy()
        return self._trajectory_bundle

    @trajectories.setter
    def trajectories(self, trajectories: List[TrajectorySample]):
        self._trajectory_bundle = trajectories
        self._lazy = None

    def sort(self):
        """Ascending, stable by cost (reference :502-510).  Device-backed samples already carry their cost;
        anything else is evaluated with the bundle's cost function."""
        if not self._is_sorted:
            for trajectory in self.trajectories:
                if trajectory._backing is None or trajectory._cost_function is not None:
                    trajectory.cost = self._cost_function
            self._trajectory_bundle.sort(key=lambda x: x.cost)
            self._is_sorted = True

    def optimal_trajectory(self) -> Union[TrajectorySample, None]:
        if not self.trajectories or not self._is_sorted:
            return None
        return min(self._trajectory_bundle, key=lambda x: x.cost)

    def min_costs(self) -> TrajectorySample:
        return self.trajectories[0] if self._is_sorted else None

    def max_costs(self) -> TrajectorySample:
        return self.trajectories[-1] if self._is_sorted else None

    def get_sorted_list(self) -> list:
        if not self._is_sorted:
            self.sort()
        return self._trajectory_bundle

    def filter_goals_behind(self):
        """Drop candidates whose longitudinal goal lies behind the start (reference :545-550)."""
        self.trajectories = [t for t in self.trajectories if t.trajectory_long.x_0[0] < t.trajectory_long.x_d[0]]

    @property
    def empty(self) -> bool:
        return len(self.trajectories) == 0
