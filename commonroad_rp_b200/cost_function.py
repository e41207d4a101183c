"""Cost functions -- same plugin interface as the reference's ``commonroad_rp/cost_function.py``
(``CostFunction.evaluate(trajectory) -> float``).

Inside ``ReactivePlanner.plan()`` the two built-in cost functions are evaluated for ALL candidates by
the fused CUDA kernel (csrc/rp_kernels.cuh, cost phase: numpy's pairwise summation order is reproduced
so that near-ties break like the reference).  ``device_spec()`` tells the planner how to configure that
kernel; a user subclass without it is evaluated through its own Python ``evaluate`` on materialised
states (generic plug-in path).  ``evaluate`` below is the public per-sample entry
point for user code; the planner does not call it for the built-in classes.
"""
from abc import ABC, abstractmethod
from typing import Optional

import numpy as np

from commonroad_rp_b200 import _lib


class CostFunction(ABC):
    """Abstract base class for new cost functions (reference :17-32)."""

    def __init__(self):
        pass

    @abstractmethod
    def evaluate(self, trajectory) -> float:
        """Computes the costs of a given trajectory sample."""


class DefaultCostFunction(CostFunction):
    """Default cost function for comfort driving (reference :35-71): squared weighted acceleration,
    velocity error (+ end and mid-horizon terms), optional longitudinal position error, lateral offset
    and curvilinear orientation (each with an end term)."""

    def __init__(self, desired_speed: Optional[float] = None, desired_d: float = 0.0,
                 desired_s: Optional[float] = None):
        super().__init__()
        self.desired_speed = desired_speed
        self.desired_d = desired_d
        self.desired_s = desired_s
        self.w_a = 5

    def device_spec(self) -> dict:
        return {"cost_kind": _lib.COST_DEFAULT, "desired_speed": self.desired_speed, "desired_s": self.desired_s,
                "desired_d": self.desired_d, "w_a": self.w_a}

    def evaluate(self, trajectory):
        ca, cu = trajectory.cartesian, trajectory.curvilinear
        total = 0.0
        total += np.sum((self.w_a * ca.a) ** 2)
        if self.desired_speed is not None:
            dv = ca.v - self.desired_speed
            total += np.sum((5 * dv) ** 2) + (50 * dv[-1] ** 2) + (100 * dv[int(len(ca.v) / 2)] ** 2)
        if self.desired_s is not None:
            ds = self.desired_s - cu.s
            total += np.sum((0.25 * ds) ** 2) + (20 * ds[-1]) ** 2
        dd = self.desired_d - cu.d
        total += np.sum((0.25 * dd) ** 2) + (20 * dd[-1]) ** 2
        th = np.abs(cu.theta)
        total += np.sum((0.25 * th) ** 2) + (5 * th[-1]) ** 2
        return total


class DefaultCostFunctionFailSafe(CostFunction):
    """Default cost function for fail-safe trajectory planning (reference :74-92)."""

    def __init__(self):
        super().__init__()

    def device_spec(self) -> dict:
        return {"cost_kind": _lib.COST_FAILSAFE, "desired_speed": None, "desired_s": None, "desired_d": 0.0,
                "w_a": 1.0}

    def evaluate(self, trajectory):
        ca, cu = trajectory.cartesian, trajectory.curvilinear
        total = np.sum((1 * ca.a) ** 2)
        total += np.sum((0.25 * cu.d) ** 2) + (20 * cu.d[-1]) ** 2
        th = np.abs(cu.theta)
        total += np.sum((0.25 * th) ** 2) + (5 * th[-1]) ** 2
        return total
