"""Planner state used at the API boundary (the reference's ``commonroad_rp/state.py``): a kinematic
single-track state whose POSITION REFERS TO THE REAR AXLE, extended by acceleration and yaw rate.
Host glue, not on the accelerated path."""
import math
from dataclasses import dataclass
from typing import Any

import numpy as np

from commonroad_rp_b200._compat import InitialState, KSState  # noqa: F401  (InitialState re-exported for users)


def _along_heading(theta: float, distance: float) -> np.ndarray:
    return np.array([distance * math.cos(theta), distance * math.sin(theta)])


@dataclass(eq=False)
class ReactivePlannerState(KSState):
    acceleration: Any = None
    yaw_rate: Any = None

    def __repr__(self):
        parts = ("time_step", "position", "steering_angle", "velocity", "orientation", "acceleration", "yaw_rate")
        return "(" + ", ".join("%s=%s" % (p, getattr(self, p)) for p in parts) + ")"

    def shift_positions_to_center(self, wb_rear_axle: float):
        """Copy of this state with the position moved from the rear axle to the vehicle centre."""
        return self.translate_rotate(_along_heading(self.orientation, wb_rear_axle), 0.0)

    @classmethod
    def create_from_initial_state(cls, initial_state, wheelbase: float, wb_rear_axle: float):
        """Planning-problem initial state (vehicle centre, slip angle) -> planner state: acceleration
        defaults to 0, the slip angle is dropped, the position moves back to the rear axle and the steering
        angle follows from the yaw rate (reference state.py:33-67)."""
        if getattr(initial_state, "acceleration", None) is None:
            initial_state.acceleration = 0.
        try:
            del initial_state.slip_angle
        except AttributeError:
            pass
        rear = initial_state.translate_rotate(_along_heading(initial_state.orientation, -wb_rear_axle), 0.0)
        planner_state = rear.convert_state_to_state(cls())
        planner_state.steering_angle = np.arctan2(wheelbase * planner_state.yaw_rate, planner_state.velocity)
        return planner_state
