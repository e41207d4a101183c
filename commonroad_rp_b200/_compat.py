"""commonroad-io types used at the planner's API boundary.  When commonroad-io is installed its own
classes are used (so that results plug into the CommonRoad tool chain); otherwise light stand-ins with
the same attribute names keep ``ReactivePlanner.plan()`` self-contained (this image has no commonroad-io).
"""
import dataclasses
from dataclasses import dataclass
from typing import Any, List

import numpy as np

try:  # pragma: no cover - not installable in the build image
    from commonroad.scenario.state import KSState, CustomState, InputState, InitialState  # noqa: F401
    from commonroad.scenario.trajectory import Trajectory  # noqa: F401
    HAVE_COMMONROAD_IO = True
except ImportError:
    HAVE_COMMONROAD_IO = False

    @dataclass(eq=False)
    class KSState:
        time_step: Any = None
        position: Any = None
        steering_angle: Any = None
        velocity: Any = None
        orientation: Any = None

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

    class CustomState:
        def __init__(self, **kwargs):
            self.__dict__.update(kwargs)

        def __repr__(self):
            return "CustomState(%s)" % ", ".join("%s=%r" % kv for kv in self.__dict__.items())

    @dataclass(eq=False)
    class InputState:
        time_step: Any = None
        acceleration: Any = None
        steering_angle_speed: Any = None

    @dataclass(eq=False)
    class InitialState(KSState):
        yaw_rate: Any = None
        slip_angle: Any = None
        acceleration: Any = None

    class Trajectory:
        def __init__(self, initial_time_step: int, state_list: List[Any]):
            self.initial_time_step = initial_time_step
            self.state_list = state_list

        @property
        def final_state(self):
            return self.state_list[-1]
