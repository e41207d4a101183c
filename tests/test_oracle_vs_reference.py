"""oracle/rp_oracle.py against the reference's own code executed live through oracle/ref_shims.py
(only where /root/reference exists).  Same machine, same numpy => bit-identical."""
import numpy as np
import pytest

from commonroad_rp_b200.utility import synthetic

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("seed,level,N,sdot,mode", [(11, 1, 20, 14.0, "velocity_keeping"), (12, 2, 20, 2.5, "velocity_keeping"),
                                                    (13, 1, 40, 9.0, "velocity_keeping"), (14, 1, 30, 7.0, "stopping")])
def test_port_is_bit_identical_to_reference(seed, level, N, sdot, mode):
    from oracle import ref_harness as H, rp_oracle as O
    scn = synthetic.make_scenario(seed=seed, amplitude=10.0 + seed, wavelength=50.0)
    p = H.build_planner(scn, N=N, longitudinal_mode=mode)
    s0 = float(p.coordinate_system.ref_pos[12])
    H.set_initial_state(p, [s0, sdot, 0.2], [0.25, 0.1, 0.0])
    if mode == "stopping":
        p.set_desired_lon_position(s0 + 15.0)
    else:
        p.set_desired_velocity(desired_velocity=sdot, current_speed=sdot)
    prob = H.problem_from_planner(p, level, scn)
    r = H.evaluate_level(p, level)
    o = O.plan_grid(prob, full_collision=False)
    assert np.array_equal(r["coeffs_lon"], o["coeffs_lon"]) and np.array_equal(r["coeffs_lat"], o["coeffs_lat"])
    assert np.array_equal(r["kin_feasible"], o["kin_feasible"])
    assert np.array_equal(r["cost"], o["cost"], equal_nan=True)
    assert r["winner"] == o["winner"]
    assert r["reasons"] == o["reasons"]
    assert r["n_infeasible_kinematics"] == o["n_infeasible_kinematics"]
    assert r["n_infeasible_collision"] == o["n_infeasible_collision"]
    m = r["kin_feasible"]
    assert np.array_equal(r["states"][m], o["states"][m])


def test_reference_fork_mode_matches_single_process():
    """the reference's own multiproc path (reactive_planner.py:1084-1111) selects the same trajectory"""
    from oracle import ref_harness as H
    scn = synthetic.make_scenario(seed=3)
    winners = []
    for multiproc in (False, True):
        p = H.build_planner(scn, N=20)
        p.config.debug.multiproc = multiproc
        p.config.debug.num_workers = 3
        H.set_initial_state(p, [float(p.coordinate_system.ref_pos[10]), 15.0, 0.0], [0.3, 0.0, 0.0])
        p.set_desired_velocity(desired_velocity=15.0, current_speed=15.0)
        out = p.plan()
        winners.append(np.array([[s.position[0], s.position[1]] for s in out[0].state_list]))
    assert np.allclose(winners[0], winners[1], rtol=0, atol=0)
