"""``ReactivePlannerState`` -- same class as the reference's ``commonroad_rp/state.py`` (host glue,
out of scope to accelerate; positions refer to the rear axle)."""
from dataclasses import dataclass
from typing import Any

import numpy as np

from commonroad_rp_b200._compat import KSState, InitialState


@dataclass(eq=False)
class ReactivePlannerState(KSState):
    """KSState + acceleration and yaw rate; position is the REAR-AXLE position (reference state.py:7-20)."""
    acceleration: Any = None
    yaw_rate: Any = None

    def __repr__(self):
        return f"(time_step={self.time_step}, position={self.position},steering_angle={self.steering_angle}, " \
               f"velocity={self.velocity}, orientation={self.orientation}, acceleration={self.acceleration}, " \
               f"yaw_rate = {self.yaw_rate})"

    def shift_positions_to_center(self, wb_rear_axle: float):
        """rear axle -> vehicle centre (reference state.py:22-31)"""
        th = self.orientation
        return self.translate_rotate(np.array([wb_rear_axle * np.cos(th), wb_rear_axle * np.sin(th)]), 0.0)

    @classmethod
    def create_from_initial_state(cls, initial_state, wheelbase: float, wb_rear_axle: float):
        """InitialState (vehicle centre) -> planner state at the rear axle, steering angle from the yaw rate
        (reference state.py:33-67)."""
        if getattr(initial_state, "acceleration", None) is None:
            initial_state.acceleration = 0.
        if hasattr(initial_state, "slip_angle"):
            try:
                delattr(initial_state, "slip_angle")
            except AttributeError:
                pass
        th = initial_state.orientation
        shifted = initial_state.translate_rotate(np.array([-wb_rear_axle * np.cos(th), -wb_rear_axle * np.sin(th)]), 0.0)
        x0 = shifted.convert_state_to_state(cls())
        x0.steering_angle = np.arctan2(wheelbase * x0.yaw_rate, x0.velocity)
        return x0
