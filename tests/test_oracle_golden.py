"""The portable oracle (oracle/rp_oracle.py) against the fixtures generated from the REFERENCE's own
code (oracle/make_golden.py).  Runs on any machine: this is what pins the oracle where /root/reference
does not exist."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import rp_oracle as O
from tests import golden_io

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SYN = sorted(glob.glob(os.path.join(GOLDEN, "syn_*.npz")))
# same numpy build => bit identical; other machines may dispatch different SIMD transcendental kernels
RTOL = 1e-11


def _close(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.allclose(a, b, rtol=RTOL, atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("path", SYN, ids=[os.path.basename(p)[4:-4] for p in SYN])
def test_oracle_matches_reference_fixture(path):
    z = np.load(path)
    prob = golden_io.unpack_problem(z)
    o = O.plan_grid(prob, want_states=True, full_collision=False)
    assert o["n"] == len(z["r_cost"])
    assert _close(o["coeffs_lon"], z["r_coeffs_lon"]) and _close(o["coeffs_lat"], z["r_coeffs_lat"])
    assert _close(o["delta_tau_lat"], z["r_delta_tau_lat"])
    assert np.array_equal(o["kin_feasible"], z["r_kin_feasible"])           # flags: exact
    assert _close(o["cost"], z["r_cost"])
    assert o["winner"] == int(z["r_winner"])                                # selected index: exact
    assert o["n_infeasible_kinematics"] == int(z["r_n_inf_kin"])
    assert o["n_infeasible_collision"] == int(z["r_n_inf_col"])
    assert o["reasons"] == json.loads(str(z["r_reasons"]))
    idx = z["r_state_idx"]
    assert _close(o["states"][idx], z["r_states"])
    # labels the lazy collision pass of the reference leaves behind
    lab = z["r_label"]
    assert np.array_equal(lab == 3, (o["status"] == O.ST_COLLISION) & (lab == 3))
    assert np.all(o["status"][lab == 3] == O.ST_COLLISION)


def test_fixtures_cover_the_branches():
    names = {os.path.basename(p)[4:-4] for p in SYN}
    assert {"lowvel", "standstill_carry", "draw_all", "stopping", "lvl2_N60", "dense_small"} <= names


def test_oracle_initial_states_match_reference_vectors():
    """the oracle's restatement of _compute_initial_states (reactive_planner.py:446-512) against
    the reference's own outputs on the Cartesian states of every recorded replanning cycle (init_states.npz)"""
    g = np.load(os.path.join(GOLDEN, "init_states.npz"))
    wheelbase = O.vehicle_dict()["wheelbase"]
    n_checked = 0
    for path in sorted(glob.glob(os.path.join(GOLDEN, "cyc_*.npz"))):
        z = np.load(path)
        name = json.loads(str(z["meta"]))["name"]
        ref, ccosy, _ = O.reference_tables(z["ref_path_raw"])
        frame = O._ArrayCCosy(ccosy)
        for row, x in enumerate(g[name + "_x0"]):
            for tag, flag in (("hv", False), ("lv", True)):
                lon, lat = O.initial_states(x[[0, 1, 2, 3, 4, 6]], flag, ref, frame, wheelbase)
                assert _close(lon, g[name + "_lon_" + tag][row]) and _close(lat, g[name + "_lat_" + tag][row]), (name, row, tag)
                n_checked += 1
    assert n_checked >= 60
    with pytest.raises(ValueError):
        O.initial_states([1e6, 1e6, 0, 1, 0, 0], False, ref, frame, wheelbase)
