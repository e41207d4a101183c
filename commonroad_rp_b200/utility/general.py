"""Host glue of the reference's utility/general.py (scenario loading, goal velocity, orientation shift)."""
from typing import Optional

import numpy as np


def load_scenario_and_planning_problem(path_scenario, idx_planning_problem: Optional[int] = None):
    """CommonRoad XML -> (scenario, planning problem, planning problem set); needs commonroad-io
    (reference utility/general.py:11-29)."""
    try:
        from commonroad.common.file_reader import CommonRoadFileReader
    except ImportError as e:
        raise ImportError("loading CommonRoad XML needs commonroad-io; pass scenario objects / a "
                          "collision.CollisionChecker to the planner instead") from e
    scenario, pps = CommonRoadFileReader(path_scenario).open()
    if idx_planning_problem is not None:
        try:
            pp = pps.find_planning_problem_by_id(idx_planning_problem)
        except KeyError:
            raise KeyError(f"<ReactivePlannerConfiguration.update()>:"
                           f"Planning Problem with ID: {idx_planning_problem} does not exist!")
    else:
        pp = list(pps.planning_problem_dict.values())[0]
    return scenario, pp, pps


def retrieve_desired_velocity_from_pp(planning_problem):
    """average goal velocity, else the initial velocity (reference utility/general.py:32-46)"""
    goal_state = planning_problem.goal.state_list[0]
    if hasattr(goal_state, 'velocity'):
        if goal_state.velocity.start > 0:
            return (goal_state.velocity.start + goal_state.velocity.end) / 2
        return goal_state.velocity.end / 2
    return planning_problem.initial_state.velocity


def shift_orientation(trajectory, interval_start=-np.pi, interval_end=np.pi):
    """fold every state's orientation into [interval_start, interval_end] (reference utility/general.py:49-55)"""
    for state in trajectory.state_list:
        while state.orientation < interval_start:
            state.orientation += 2 * np.pi
        while state.orientation > interval_end:
            state.orientation -= 2 * np.pi
    return trajectory
