"""CPU-only probe (development aid): pure-Python overhead of ReactivePlanner.plan() with the engine call replaced by
canned results of the recorded ZAM_Over-1_1 cycles.   python tools/probe_python_local.py [profile]"""
import cProfile
import json
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from commonroad_rp_b200 import _lib, collision  # noqa: E402
from commonroad_rp_b200.reactive_planner import ReactivePlanner  # noqa: E402
from commonroad_rp_b200.state import ReactivePlannerState  # noqa: E402
from commonroad_rp_b200.utility.config import ReactivePlannerConfiguration  # noqa: E402
from commonroad_rp_b200.utility.utils_coordinate_system import CoordinateSystem  # noqa: E402
from tests import golden_io  # noqa: E402


class FakeEngine:
    plan_generation = 0
    _cyc_chosen = 0

    def __init__(self):
        self.res = (_lib.PlanResult * 4)()
        self.states = None

    def cycle_limits(self):
        return (4, 448, 128, 1 << 20)

    def plan_levels(self, inputs, levels):
        self.plan_generation += 1
        return self.res, 0

    def cycle_winner_states(self):
        return self.states.copy()

    def select_level(self, j):
        pass

    _cyc_selected = 0


def main():
    z = dict(np.load(os.path.join(ROOT, "tests", "golden", "cyc_ZAM_Over-1_1.npz")))
    meta = json.loads(str(z["meta"]))
    cfg = ReactivePlannerConfiguration()
    cfg.planning.time_steps_computation = meta["N"]
    cfg.sampling.t_min = meta["t_min"]

    class _Empty:
        static_obstacles, dynamic_obstacles = (), ()
        lanelet_network = type("LN", (), {"lanelets": ()})()

    cfg.update(scenario=_Empty(), planning_problem=None)
    planner = ReactivePlanner(cfg)
    fake = FakeEngine()
    planner._engine = fake
    if os.environ.get("RP_SPECULATION"):
        planner.speculation = os.environ["RP_SPECULATION"]
    planner._sync_device_tables = lambda: None
    cc = collision.checker_from_arrays(**golden_io.unpack_obstacles(z, "ob_"))
    co = CoordinateSystem(z["ref_path_raw"])
    times = []

    def run(reps):
        for rep in range(reps):
            for ci in range(meta["n_cycles"]):
                x = z["c%d_x0" % ci]
                x0 = ReactivePlannerState(time_step=int(x[7]), position=np.array([x[0], x[1]]), orientation=x[2],
                                          velocity=x[3], acceleration=x[4], yaw_rate=x[5], steering_angle=x[6])
                planner.reset(initial_state_cart=x0, initial_state_curv=(list(z["c%d_x0_lon" % ci]), list(z["c%d_x0_lat" % ci])),
                              collision_checker=cc, coordinate_system=co)
                planner.set_desired_velocity(desired_velocity=meta["desired_velocity"] if ci == 0 else None, current_speed=x0.velocity)
                lv = meta["cycles"][ci]["levels"][0]
                fake.res[0].winner = lv["winner"]
                fake.res[0].n_candidates = lv["n"]
                fake.res[0].winner_cost = 1.0
                fake.states = z["c%d_l0_winner_states" % ci]
                t0 = time.perf_counter()
                out = planner.plan()
                times.append(time.perf_counter() - t0)
                assert out is not None

    run(3)
    del times[:]
    if len(sys.argv) > 1 and sys.argv[1] == "profile":
        pr = cProfile.Profile()
        pr.enable()
        run(200)
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(28)
    else:
        run(200)
    t = np.array(times) * 1e6
    print("pure-Python plan() overhead: p10 %.1f us  p50 %.1f us  p95 %.1f us  (n=%d)" % (np.percentile(t, 10), np.percentile(t, 50), np.percentile(t, 95), len(t)))


if __name__ == "__main__":
    main()
