"""Small host helpers of the reference's ``utility/general.py``: scenario loading (commonroad-io, or the package's
own XML reader), the desired velocity implied by a planning problem, orientation folding."""
from typing import Optional

import numpy as np

_TWO_PI = 2 * np.pi


def load_scenario_and_planning_problem(path_scenario, idx_planning_problem: Optional[int] = None):
    """(scenario, planning problem, planning problem set) from a CommonRoad XML file (reference utility/general.py:11-29):
    through commonroad-io when it is installed, else through the package's own reader (utility/scenario_io.py; the
    planning problem set is then None)."""
    try:
        from commonroad.common.file_reader import CommonRoadFileReader
    except ImportError:
        from commonroad_rp_b200.utility.scenario_io import read_commonroad_xml
        scenario, problem = read_commonroad_xml(path_scenario, idx_planning_problem)
        return scenario, problem, None
    scenario, problem_set = CommonRoadFileReader(path_scenario).open()
    if idx_planning_problem is None:
        problem = next(iter(problem_set.planning_problem_dict.values()))
    else:
        try:
            problem = problem_set.find_planning_problem_by_id(idx_planning_problem)
        except KeyError:
            raise KeyError(f"<ReactivePlannerConfiguration.update()>:"
                           f"Planning Problem with ID: {idx_planning_problem} does not exist!")
    return scenario, problem, problem_set


def retrieve_desired_velocity_from_pp(planning_problem):
    """Mid-point of the goal velocity interval (half the upper bound if the interval starts at 0); the
    initial velocity when the goal names no velocity (reference utility/general.py:32-46)."""
    goal_state = planning_problem.goal.state_list[0]
    if not hasattr(goal_state, 'velocity'):
        return planning_problem.initial_state.velocity
    lo, hi = goal_state.velocity.start, goal_state.velocity.end
    return (lo + hi) / 2 if lo > 0 else hi / 2


def shift_orientation(trajectory, interval_start=-np.pi, interval_end=np.pi):
    """Fold the orientation of every state into [interval_start, interval_end] by whole turns."""
    for state in trajectory.state_list:
        theta = state.orientation
        while theta < interval_start:
            theta += _TWO_PI
        while theta > interval_end:
            theta -= _TWO_PI
        state.orientation = theta
    return trajectory
