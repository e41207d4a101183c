"""TEST INFRASTRUCTURE ONLY -- the portable CPU oracle of the candidate-trajectory hot path.

A numpy / scalar-Python restatement of the reference algorithm over plain arrays (no
commonroad objects), so that it can travel to the GPU box where /root/reference does not
exist.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it; the product package never does.

Each function cites the reference lines it follows (paths relative to /root/reference).  The
scalar expressions deliberately keep the reference's operation order and its mix of numpy
scalars, ``math`` and ``**`` so that, on the same machine, results are bit-identical to the
reference executed through oracle/ref_shims.py (checked by tests/test_oracle_vs_reference.py
here, and pinned for other machines by the fixtures in tests/golden/).

Pinning status: everything that lives in commonroad_rp is pinned by the reference's own code;
the third-party pieces (CCosy projection, OBB SAT, make_valid_orientation) come from
oracle/third_party.py and are PARITY UNPINNED (no reference test or golden vector exists).
"""
import math

import numpy as np

from . import third_party as tp

_EPS = 1e-5                       # reactive_planner.py:49
CONSTRAINTS = ("velocity", "acceleration", "kappa", "kappa_dot", "yaw_rate")
# reason codes shared with include/rp_b200.h
R_NONE, R_VELOCITY, R_ACCELERATION, R_KAPPA, R_KAPPA_DOT, R_YAW_RATE, R_PROJECTION, R_REF_RANGE = range(8)
REASON_OF = {"velocity": R_VELOCITY, "acceleration": R_ACCELERATION, "kappa": R_KAPPA,
             "kappa_dot": R_KAPPA_DOT, "yaw_rate": R_YAW_RATE}
ST_FEASIBLE, ST_KINEMATIC, ST_COLLISION, ST_FILTERED = 0, 1, 2, 3
STATE_FIELDS = ("x", "y", "theta", "v", "a", "kappa", "kappa_dot",
                "s", "d", "theta_cl", "s_dot", "s_ddot", "d_dot", "d_ddot")


# ------------------------------------------------------------------------------------------------
# a3: coefficient solves  (polynomial_trajectory.py:292-320, :341-360)
# ------------------------------------------------------------------------------------------------
def solve_quintic(p0, p0d, p0dd, pf, pfd, pfdd, tau):
    t2 = np.power(tau, 2)
    t3 = t2 * tau
    t4 = t2 * t2
    t5 = t4 * tau
    a = np.array([[t3, t4, t5],
                  [3. * t2, 4. * t3, 5. * t4],
                  [6. * tau, 12. * t2, 20. * t3]])
    b = np.array([pf - (p0 + p0d * tau + .5 * p0dd * t2),
                  pfd - (p0d + p0dd * tau),
                  pfdd - p0dd])
    x = np.linalg.solve(a, b)
    return np.array([p0, p0d, .5 * p0dd, x[0], x[1], x[2]])


def solve_quartic(p0, p0d, p0dd, tau, v_des):
    t2 = np.power(tau, 2)
    t3 = t2 * tau
    a = np.array([[3. * t2, 4. * t3],
                  [6. * tau, 12. * t2]])
    b = np.array([v_des - p0d - p0dd * tau, - p0dd])
    x = np.linalg.solve(a, b)
    return np.array([p0, p0d, .5 * p0dd, x[0], x[1], 0.])


# a4: power-form evaluation (polynomial_trajectory.py:240-271)
def poly_pos(c, t, t2, t3, t4, t5):
    return c[0] + c[1] * t + c[2] * t2 + c[3] * t3 + c[4] * t4 + c[5] * t5


def poly_vel(c, t, t2, t3, t4):
    return c[1] + 2. * c[2] * t + 3. * c[3] * t2 + 4. * c[4] * t3 + 5. * c[5] * t4


def poly_acc(c, t, t2, t3):
    return 2 * c[2] + 6 * c[3] * t + 12 * c[4] * t2 + 20 * c[5] * t3


def position_at_tau(c, tau, delta_tau):
    """evaluate_state_at_tau(tau)[0] with tau_0 = 0 (polynomial_trajectory.py:192-227)."""
    if tau < 0:
        tau = 0
    elif tau > delta_tau:
        tau = delta_tau
    tau2 = np.power(tau, 2)
    tau3 = tau2 * tau
    tau4 = tau2 * tau2
    tau5 = tau3 * tau2
    return poly_pos(c, tau, tau2, tau3, tau4, tau5)


# ------------------------------------------------------------------------------------------------
# a2: enumeration t x lon x d in the HOST-GIVEN order (sampling.py:202-242)
# ------------------------------------------------------------------------------------------------
def enumerate_grid(t_list, lon_list, d_list, x0_lon, x0_lat, lon_mode, low_vel_mode):
    """Returns coeffs_lon [n,6], coeffs_lat [n,6], delta_tau_lon [n], delta_tau_lat [n],
    goal_behind [n] (the ``filter_goals_behind`` predicate, trajectories.py:545-550)."""
    cl, ct, dtl, dtt, behind = [], [], [], [], []
    lon_cache, lat_cache = {}, {}
    for t in t_list:
        for lon in lon_list:
            key = (float(t), float(lon))
            if key not in lon_cache:
                if lon_mode == "velocity_keeping":       # sampling.py:254-258
                    lon_cache[key] = solve_quartic(x0_lon[0], x0_lon[1], x0_lon[2], t, lon)
                elif lon_mode == "stopping":             # sampling.py:259-263
                    lon_cache[key] = solve_quintic(x0_lon[0], x0_lon[1], x0_lon[2], lon, 0.0, 0.0, t)
                else:
                    raise AttributeError("invalid longitudinal mode %r" % (lon_mode,))
            c_lon = lon_cache[key]
            for d in d_list:
                if low_vel_mode:                         # sampling.py:229-234
                    s_goal = position_at_tau(c_lon, t, t) - x0_lon[0]
                    if s_goal <= 0:
                        s_goal = t
                    tau_lat = s_goal
                else:
                    tau_lat = t
                k2 = (float(tau_lat), float(d))
                if k2 not in lat_cache:
                    lat_cache[k2] = solve_quintic(x0_lat[0], x0_lat[1], x0_lat[2], d, 0.0, 0.0, tau_lat)
                cl.append(c_lon)
                ct.append(lat_cache[k2])
                dtl.append(t)
                dtt.append(tau_lat)
                behind.append(bool(lon_mode == "stopping" and not (x0_lon[0] < lon)))
    n = len(cl)
    return (np.array(cl).reshape(n, 6), np.array(ct).reshape(n, 6), np.array(dtl, dtype=np.float64),
            np.array(dtt, dtype=np.float64), np.array(behind, dtype=bool))


def traj_len_of(delta_tau, dt):
    """reactive_planner.py:733, :748."""
    return len(np.arange(0, np.round(delta_tau + dt, 5), dt))


# ------------------------------------------------------------------------------------------------
# a8: interpolate_angle (utility/utils_coordinate_system.py:25-43)
# ------------------------------------------------------------------------------------------------
def interpolate_angle(x, x1, x2, y1, y2):
    delta = y2 - y1
    return tp.make_valid_orientation(delta * (x - x1) / (x2 - x1) + y1)


# ------------------------------------------------------------------------------------------------
# a7: constraint checks at step i (reactive_planner.py:971-1017); returns reason code or 0
# ------------------------------------------------------------------------------------------------
def initial_states(x0, low_vel_mode, ref, ccosy, wheelbase):
    """ReactivePlanner._compute_initial_states (reactive_planner.py:446-512): Cartesian rear-axle state
    x0 = (x, y, orientation, velocity, acceleration, steering_angle) -> ([s, s_dot, s_ddot], [d, d_dot, d_ddot]).
    ``ccosy`` restates pycrccosy's convert_to_curvilinear_coords (oracle/third_party.py); raises like the reference."""
    px, py, orientation, velocity, acceleration, steering_angle = [float(v) for v in x0]
    s, d = ccosy.convert_to_curvilinear_coords(px, py)                       # ValueError outside the domain (:459-461)
    ref_pos, ref_curv, ref_curv_d = ref["ref_pos"], ref["ref_curv"], ref["ref_curv_d"]
    s_idx = np.argmax(ref_pos > s) - 1                                       # :464
    s_lambda = (s - ref_pos[s_idx]) / (ref_pos[s_idx + 1] - ref_pos[s_idx])
    ref_theta = np.unwrap(ref["ref_theta"])                                  # :469
    theta_cl = orientation - interpolate_angle(s, ref_pos[s_idx], ref_pos[s_idx + 1], ref_theta[s_idx], ref_theta[s_idx + 1])
    kr = (ref_curv[s_idx + 1] - ref_curv[s_idx]) * s_lambda + ref_curv[s_idx]
    kr_d = (ref_curv_d[s_idx + 1] - ref_curv_d[s_idx]) * s_lambda + ref_curv_d[s_idx]
    kappa_0 = np.tan(steering_angle) / wheelbase                             # :480
    d_p = (1 - kr * d) * np.tan(theta_cl)                                    # :483
    d_pp = -(kr_d * d + kr * d_p) * np.tan(theta_cl) + ((1 - kr * d) / (math.cos(theta_cl) ** 2)) * (
        kappa_0 * (1 - kr * d) / math.cos(theta_cl) - kr)
    s_velocity = velocity * math.cos(theta_cl) / (1 - kr * d)                # :488
    if s_velocity < 0:
        raise Exception("Initial state or reference incorrect! Curvilinear velocity is negative")
    s_acceleration = acceleration                                            # :493-497
    s_acceleration -= (s_velocity ** 2 / math.cos(theta_cl)) * (
        (1 - kr * d) * np.tan(theta_cl) * (kappa_0 * (1 - kr * d) / (math.cos(theta_cl)) - kr) - (kr_d * d + kr * d_p))
    s_acceleration /= ((1 - kr * d) / (math.cos(theta_cl)))
    if low_vel_mode:                                                         # :500-507
        d_velocity, d_acceleration = d_p, d_pp
    else:
        d_velocity = velocity * math.sin(theta_cl)
        d_acceleration = s_acceleration * d_p + s_velocity ** 2 * d_pp
    return [s, s_velocity, s_acceleration], [d, d_velocity, d_acceleration]


def check_constraints(v, kappa_gl, theta_gl, a, i, veh, dt, constraints):
    if "velocity" in constraints:
        if v[i] < -_EPS:
            return R_VELOCITY
    kappa_max = np.tan(veh["delta_max"]) / veh["wheelbase"]
    if "kappa" in constraints:
        if abs(kappa_gl[i]) > kappa_max:
            return R_KAPPA
    if "yaw_rate" in constraints:
        yaw_rate = (theta_gl[i] - theta_gl[i - 1]) / dt if i > 0 else 0.
        theta_dot_max = kappa_max * v[i]
        if abs(round(yaw_rate, 5)) > theta_dot_max:
            return R_YAW_RATE
    if "kappa_dot" in constraints:
        steering_angle = np.arctan2(veh["wheelbase"] * kappa_gl[i], 1.0)
        kappa_dot_max = veh["v_delta_max"] / (veh["wheelbase"] * math.cos(steering_angle) ** 2)
        kappa_dot = (kappa_gl[i] - kappa_gl[i - 1]) / dt if i > 0 else 0.
        if abs(kappa_dot) > kappa_dot_max:
            return R_KAPPA_DOT
    if "acceleration" in constraints:
        v_switch = veh["v_switch"]
        a_max = veh["a_max"] * v_switch / v[i] if v[i] > v_switch else veh["a_max"]
        a_min = -veh["a_max"]
        if not a_min <= a[i] <= a_max:
            return R_ACCELERATION
    return R_NONE


# ------------------------------------------------------------------------------------------------
# a10: horizon extension (trajectories.py:168-197, :302-332)
# ------------------------------------------------------------------------------------------------
def enlarge_cartesian(x, y, theta, v, a, kappa, kappa_dot, traj_len, dt):
    last = traj_len - 1
    steps = len(x) - traj_len
    t = np.arange(1, steps + 1, 1) * dt
    a[traj_len:] = np.repeat(a[last], steps)
    v_temp = v[last] + t * a[-1]
    v_temp = v_temp * np.greater_equal(v_temp, 0)
    v[traj_len:] = v_temp
    theta[traj_len:] = np.repeat(theta[last], steps)
    kappa[traj_len:] = np.repeat(kappa[last], steps)
    kappa_dot[traj_len:] = np.repeat(kappa_dot[last], steps)
    x[traj_len:] = x[last] + np.cumsum(dt * v_temp * math.cos(theta[last]))
    y[traj_len:] = y[last] + np.cumsum(dt * v_temp * math.sin(theta[last]))


def enlarge_curvilinear(s, d, theta, s_dot, s_ddot, d_dot, d_ddot, traj_len, dt):
    last = traj_len - 1
    steps = len(s) - traj_len
    t = np.arange(1, (steps + 1), 1) * dt
    s_dot_temp = s_dot[last] + t * s_ddot[-1]          # s_ddot[-1] is still the zero tail (App. B#7)
    s_dot_temp = s_dot_temp * np.greater_equal(s_dot_temp, 0)
    s_dot[traj_len:] = s_dot_temp
    d_dot_temp = d_dot[last] + t * d_ddot[-1]
    d_dot[traj_len:] = d_dot_temp
    s_ddot[traj_len:] = np.repeat(s_ddot[last], steps)
    d_ddot[traj_len:] = np.repeat(d_ddot[last], steps)
    theta[traj_len:] = np.repeat(theta[last], steps)
    s[traj_len:] = s[last] + t * s_dot[last]
    d[traj_len:] = d[last] + t * d_dot[last]


# ------------------------------------------------------------------------------------------------
# a6: kinematic evaluation of ONE candidate (reactive_planner.py:731-960)
# ------------------------------------------------------------------------------------------------
def check_kinematics_one(c_lon, c_lat, delta_tau, prob, ccosy):
    """Returns (feasible, reason, first_bad_step, states[14, N+1] or None)."""
    dt = prob["dt"]
    N = prob["N"]
    veh = prob["vehicle"]
    ref = prob["ref"]
    P, TH, K, KD = ref["ref_pos"], ref["ref_theta"], ref["ref_curv"], ref["ref_curv_d"]
    low_vel = prob["low_vel_mode"]
    draw = prob.get("draw_all", False)
    constraints = prob["constraints"]

    t = np.arange(0, np.round(delta_tau + dt, 5), dt)
    t2 = np.square(t)
    t3 = t2 * t
    t4 = np.square(t2)
    t5 = t4 * t
    traj_len = len(t)

    s = np.zeros(N + 1)
    s_vel = np.zeros(N + 1)
    s_acc = np.zeros(N + 1)
    d = np.zeros(N + 1)
    d_vel = np.zeros(N + 1)
    d_acc = np.zeros(N + 1)
    s[:traj_len] = poly_pos(c_lon, t, t2, t3, t4, t5)
    s_vel[:traj_len] = poly_vel(c_lon, t, t2, t3, t4)
    s_acc[:traj_len] = poly_acc(c_lon, t, t2, t3)
    if not low_vel:
        d[:traj_len] = poly_pos(c_lat, t, t2, t3, t4, t5)
        d_vel[:traj_len] = poly_vel(c_lat, t, t2, t3, t4)
        d_acc[:traj_len] = poly_acc(c_lat, t, t2, t3)
    else:
        s1 = s[:traj_len] - s[0]
        s2 = np.square(s1)
        s3 = s2 * s1
        s4 = np.square(s2)
        s5 = s4 * s1
        d[:traj_len] = poly_pos(c_lat, s1, s2, s3, s4, s5)
        d_vel[:traj_len] = poly_vel(c_lat, s1, s2, s3, s4)
        d_acc[:traj_len] = poly_acc(c_lat, s1, s2, s3)
    s_vel[np.abs(s_vel) < _EPS] = 0.0
    d_vel[np.abs(d_vel) < _EPS] = 0.0

    x = np.zeros(N + 1)
    y = np.zeros(N + 1)
    v = np.zeros(N + 1)
    a = np.zeros(N + 1)
    theta_gl = np.zeros(N + 1)
    theta_cl = np.zeros(N + 1)
    kappa_gl = np.zeros(N + 1)

    if not draw:                                           # :796-805
        if np.any(np.abs(s_acc) > veh["a_max"]):
            return False, R_ACCELERATION, -1, None
        if np.any(s_vel < -_EPS):
            return False, R_VELOCITY, -1, None

    feasible = True
    reason = R_NONE
    bad_step = -1
    for i in range(0, traj_len):
        if not low_vel:                                    # :810-829
            if s_vel[i] > 0.001:
                dp = d_vel[i] / s_vel[i]
            else:
                dp = 0.
            ddot = d_acc[i] - dp * s_acc[i]
            if s_vel[i] > 0.001:
                dpp = ddot / (s_vel[i] ** 2)
            else:
                dpp = 0.
        else:
            dp = d_vel[i]
            dpp = d_acc[i]

        s_idx = np.argmax(P > s[i]) - 1                    # :835 (negative index wraps, App. B#8)
        if s_idx + 1 >= len(P):
            feasible = False
            if reason == R_NONE:
                reason, bad_step = R_REF_RANGE, i
            break
        s_lambda = (s[i] - P[s_idx]) / (P[s_idx + 1] - P[s_idx])

        if s_vel[i] > 0.001 or low_vel:                    # :842-863
            theta_cl[i] = np.arctan2(dp, 1.0)
            theta_gl[i] = theta_cl[i] + interpolate_angle(s[i], P[s_idx], P[s_idx + 1], TH[s_idx], TH[s_idx + 1])
        else:                                              # :866-873
            theta_gl[i] = prob["x0_orientation"] if i == 0 else theta_gl[i - 1]
            theta_cl[i] = theta_gl[i] - interpolate_angle(s[i], P[s_idx], P[s_idx + 1], TH[s_idx], TH[s_idx + 1])

        k_r = (K[s_idx + 1] - K[s_idx]) * s_lambda + K[s_idx]           # :876-880
        k_r_d = (KD[s_idx + 1] - KD[s_idx]) * s_lambda + KD[s_idx]

        oneKrD = (1 - k_r * d[i])                          # :883-888
        cosTheta = math.cos(theta_cl[i])
        tanTheta = np.tan(theta_cl[i])
        kappa_gl[i] = (dpp + (k_r * dp + k_r_d * d[i]) * tanTheta) * cosTheta * (cosTheta / oneKrD) ** 2 + (
                cosTheta / oneKrD) * k_r
        v[i] = s_vel[i] * (oneKrD / (math.cos(theta_cl[i])))             # :891
        a[i] = s_acc[i] * oneKrD / cosTheta + ((s_vel[i] ** 2) / cosTheta) * (          # :894-896
                oneKrD * tanTheta * (kappa_gl[i] * oneKrD / cosTheta - k_r) - (
                k_r_d * d[i] + k_r * dp))

        if feasible:                                       # :899-904
            r = check_constraints(v, kappa_gl, theta_gl, a, i, veh, dt, constraints)
            if r != R_NONE:
                feasible = False
                reason, bad_step = r, i
        if not feasible and not draw:
            break

    if not (feasible or draw):
        return False, reason, bad_step, None

    for i in range(0, traj_len):                           # :908-917
        try:
            pos = ccosy.convert_to_cartesian_coords(s[i], d[i])
        except Exception:
            pos = None
        if pos is not None:
            x[i] = pos[0]
            y[i] = pos[1]
        else:
            if feasible:
                reason, bad_step = R_PROJECTION, i
            feasible = False
            break

    if not feasible and not draw:
        return False, reason, bad_step, None

    kappa_dot = np.append([0], np.diff(kappa_gl))          # :923
    if N + 1 > traj_len:                                   # :933-934
        enlarge_cartesian(x, y, theta_gl, v, a, kappa_gl, kappa_dot, traj_len, dt)
        enlarge_curvilinear(s, d, theta_cl, s_vel, s_acc, d_vel, d_acc, traj_len, dt)
    states = np.stack([x, y, theta_gl, v, a, kappa_gl, kappa_dot, s, d, theta_cl, s_vel, s_acc, d_vel, d_acc])
    return feasible, reason, bad_step, states


# ------------------------------------------------------------------------------------------------
# a11: cost functions (cost_function.py:51-71, :85-92)
# ------------------------------------------------------------------------------------------------
def default_cost(states, cost):
    a, v = states[4], states[3]
    s, d, th = states[7], states[8], states[9]
    costs = 0.0
    costs += np.sum((cost["w_a"] * a) ** 2)
    if cost.get("desired_speed") is not None:
        vd = cost["desired_speed"]
        costs += np.sum((5 * (v - vd)) ** 2) + (50 * (v[-1] - vd) ** 2) + (100 * (v[int(len(v) / 2)] - vd) ** 2)
    if cost.get("desired_s") is not None:
        sd = cost["desired_s"]
        costs += np.sum((0.25 * (sd - s)) ** 2) + (20 * (sd - s[-1])) ** 2
    dd = cost.get("desired_d", 0.0)
    costs += np.sum((0.25 * (dd - d)) ** 2) + (20 * (dd - d[-1])) ** 2
    costs += np.sum((0.25 * np.abs(th)) ** 2) + (5 * (np.abs(th[-1]))) ** 2
    return costs


def failsafe_cost(states):
    a, d, th = states[4], states[8], states[9]
    costs = np.sum((1 * a) ** 2)
    costs += np.sum((0.25 * d) ** 2) + (20 * d[-1]) ** 2
    costs += np.sum((0.25 * np.abs(th)) ** 2) + (5 * (np.abs(th[-1]))) ** 2
    return costs


def evaluate_cost(states, cost):
    if cost.get("kind", "default") == "failsafe":
        return failsafe_cost(states)
    return default_cost(states, cost)


# ------------------------------------------------------------------------------------------------
# a13: ego-vs-obstacle check of one candidate (reactive_planner.py:1026-1046)
# ------------------------------------------------------------------------------------------------
def build_checker(obst, continuous=False):
    """pycrcc.CollisionChecker as set_collision_checker builds it (reactive_planner.py:234-251); with the
    continuous collision check every dynamic obstacle is replaced by its OBB-sum hulls (:240-241)."""
    cc = tp.CollisionChecker()
    for cx, cy, th, l, w in np.asarray(obst.get("static_boxes", np.zeros((0, 5)))).reshape(-1, 5):
        cc.add_collision_object(tp.RectOBB(0.5 * l, 0.5 * w, th, cx, cy))
    for t0, st, lw in zip(obst.get("dyn_t0", ()), obst.get("dyn_states", ()), obst.get("dyn_lw", ())):
        tvo = tp.TimeVariantCollisionObject(int(t0))
        for cx, cy, th in np.asarray(st).reshape(-1, 3):
            tvo.append_obstacle(tp.RectOBB(0.5 * lw[0], 0.5 * lw[1], th, cx, cy))
        if continuous:
            tvo, err = tp.trajectory_preprocess_obb_sum(tvo)
        cc.add_collision_object(tvo)
    sg = tp.ShapeGroup()
    for cx, cy, th, hl, hw in np.asarray(obst.get("boundary_boxes", np.zeros((0, 5)))).reshape(-1, 5):
        sg.add_shape(tp.RectOBB(hl, hw, th, cx, cy))
    for tri in np.asarray(obst.get("boundary_tris", np.zeros((0, 6)))).reshape(-1, 6):
        sg.add_shape(tp.Triangle(*tri))
    cc.add_collision_object(sg)
    return cc


def candidate_collides(states, prob, checker):
    """Returns the first colliding step index, or -1."""
    veh = prob["vehicle"]
    half_length = 0.5 * veh["length"]
    half_width = 0.5 * veh["width"]
    x, y, theta = states[0], states[1], states[2]
    pos1 = x + veh["wb_rear_axle"] * np.cos(theta)
    pos2 = y + veh["wb_rear_axle"] * np.sin(theta)
    for i in range(len(pos1)):
        ego = tp.TimeVariantCollisionObject(prob["x0_time_step"] + i * prob["factor"])
        ego.append_obstacle(tp.RectOBB(half_length, half_width, theta[i], pos1[i], pos2[i]))
        if checker.collide(ego):
            return i
    return -1


def candidate_collides_continuous(states, prob, checker):
    """The additional continuous check (reactive_planner.py:1049-1058): the OBB-sum hulls of consecutive ego boxes,
    time indices x_0.time_step + i (no ``factor`` here), against the checker."""
    veh = prob["vehicle"]
    x, y, theta = states[0], states[1], states[2]
    pos1 = x + veh["wb_rear_axle"] * np.cos(theta)
    pos2 = y + veh["wb_rear_axle"] * np.sin(theta)
    ego = tp.TimeVariantCollisionObject(prob["x0_time_step"])
    for i in range(len(pos1)):
        ego.append_obstacle(tp.RectOBB(0.5 * veh["length"], 0.5 * veh["width"], theta[i], pos1[i], pos2[i]))
    ego, err = tp.trajectory_preprocess_obb_sum(ego)
    return checker.collide(ego)


# ------------------------------------------------------------------------------------------------
# a12-a14: whole bundle (reactive_planner.py:1065-1136)
# ------------------------------------------------------------------------------------------------
def plan_candidates(coeffs_lon, coeffs_lat, delta_tau, prob, goal_behind=None, want_states=True,
                    full_collision=True):
    """Evaluate a bundle given per-candidate coefficients.  ``full_collision`` checks every feasible
    candidate (what the GPU does); the lazily-visited subset of the reference is reported through
    ``n_infeasible_collision`` (colliders ranked before the winner, App. B#12)."""
    n = len(delta_tau)
    N = prob["N"]
    ccosy = _ccosy_from(prob)
    continuous = bool(prob.get("continuous", False))
    checker = build_checker(prob["obstacles"], continuous)
    status = np.zeros(n, dtype=np.int32)
    reason = np.zeros(n, dtype=np.int32)
    bad_step = np.full(n, -1, dtype=np.int32)
    cost = np.full(n, np.nan)
    collide_step = np.full(n, -1, dtype=np.int32)
    states = np.full((n, 14, N + 1), np.nan) if want_states else None
    reasons = {c: 0 for c in prob["constraints"]}
    feasible_idx = []
    all_states = {}
    for k in range(n):
        if goal_behind is not None and goal_behind[k]:
            status[k] = ST_FILTERED
            continue
        ok, r, bs, st = check_kinematics_one(coeffs_lon[k], coeffs_lat[k], delta_tau[k], prob, ccosy)
        reason[k] = r
        bad_step[k] = bs
        for name, code in REASON_OF.items():
            if r == code and name in reasons:
                reasons[name] += 1
        if st is not None and want_states:
            states[k] = st
        if ok:
            feasible_idx.append(k)
            all_states[k] = st
        else:
            status[k] = ST_KINEMATIC
    n_considered = int(np.sum(status != ST_FILTERED))
    n_inf_kin = n_considered - len(feasible_idx)

    for k in feasible_idx:
        cost[k] = evaluate_cost(all_states[k], prob["cost"])
    order = sorted(feasible_idx, key=lambda k: cost[k])    # stable: ties -> lowest enum index

    winner = -1
    n_inf_col = 0
    for k in order:
        cs = candidate_collides(all_states[k], prob, checker)
        collide_step[k] = cs
        if cs >= 0:
            status[k] = ST_COLLISION
            if winner < 0:
                n_inf_col += 1
        elif winner < 0:
            winner = k
            if not full_collision:
                break
    continuous_hit = False
    if continuous and winner >= 0 and candidate_collides_continuous(all_states[winner], prob, checker):
        # the first discretely collision-free candidate fails the continuous check: the reference counts it, labels
        # it and BREAKS OUT OF THE CANDIDATE LOOP (:1054-1058) -- no trajectory at this level
        status[winner] = ST_COLLISION
        n_inf_col += 1
        winner = -1
        continuous_hit = True
    return {
        "continuous_hit": continuous_hit,
        "n": n, "status": status, "reason": reason, "bad_step": bad_step, "cost": cost,
        "collide_step": collide_step, "states": states, "winner": int(winner),
        "n_infeasible_kinematics": int(n_inf_kin), "n_infeasible_collision": int(n_inf_col),
        "reasons": reasons, "kin_feasible": np.isin(np.arange(n), feasible_idx),
    }


def plan_grid(prob, want_states=True, full_collision=True):
    """Grid form: enumerate t x lon x d (host order) and evaluate the bundle."""
    cl, ct, dtl, dtt, behind = enumerate_grid(prob["t"], prob["lon"], prob["d"], prob["x0_lon"], prob["x0_lat"],
                                              prob["lon_mode"], prob["low_vel_mode"])
    out = plan_candidates(cl, ct, dtl, prob, goal_behind=behind if prob["lon_mode"] == "stopping" else None,
                          want_states=want_states, full_collision=full_collision)
    out.update(coeffs_lon=cl, coeffs_lat=ct, delta_tau_lon=dtl, delta_tau_lat=dtt)
    return out


def _ccosy_from(prob):
    cc = prob.get("_ccosy_obj")
    if cc is None:
        cc = _ArrayCCosy(prob["ccosy"])
        prob["_ccosy_obj"] = cc
    return cc


class _ArrayCCosy(tp.CurvilinearCoordinateSystem):
    """third_party CCosy rebuilt from its packed tables (path incl. the two extension vertices)."""

    def __init__(self, tables):
        self._path = np.asarray(tables["path"], dtype=np.float64)
        self._S = np.asarray(tables["S"], dtype=np.float64)
        self._normal = np.asarray(tables["normals"], dtype=np.float64)
        self._limit = float(tables["limit"])


# ------------------------------------------------------------------------------------------------
# problem packing helpers (tests / bench)
# ------------------------------------------------------------------------------------------------
def vehicle_dict(vp=None):
    vp = vp or tp.vehicle_parameters(2)
    return {"length": vp.l, "width": vp.w, "wb_rear_axle": vp.b, "wheelbase": vp.a + vp.b,
            "a_max": vp.longitudinal.a_max, "v_switch": vp.longitudinal.v_switch,
            "delta_max": vp.steering.max, "v_delta_max": vp.steering.v_max}


def reference_tables(ref_path, smooth=True):
    """CoordinateSystem.__init__ (utility/utils_coordinate_system.py:88-118) over the third_party CCosy."""
    from scipy.interpolate import splprep, splev
    reference = np.asarray(ref_path, dtype=np.float64)
    _, idx = np.unique(reference, axis=0, return_index=True)
    reference = reference[np.sort(idx)]
    if smooth:
        tck, u = splprep(reference.T, u=None, k=3, s=0.0)
        u_new = np.linspace(u.min(), u.max(), 200)
        x_new, y_new = splev(u_new, tck, der=0)
        reference = tp.resample_polyline(np.array([x_new, y_new]).transpose(), 1.0)
        _, idx = np.unique(reference, axis=0, return_index=True)
        reference = reference[np.sort(idx)]
    cc = tp.CurvilinearCoordinateSystem(reference)
    path = np.asarray(cc.reference_path())
    ref_pos = tp.compute_pathlength_from_polyline(path)
    ref_curv = tp.compute_curvature_from_polyline(path)
    ref_theta = np.unwrap(tp.compute_orientation_from_polyline(path))
    ref_curv_d = np.gradient(ref_curv, ref_pos)
    ref = {"ref_pos": ref_pos, "ref_theta": ref_theta, "ref_curv": ref_curv, "ref_curv_d": ref_curv_d}
    ccosy = {"path": path, "S": cc.pathlength, "normals": cc.normals, "limit": cc.projection_domain_limit}
    return ref, ccosy, cc


# ------------------------------------------------------------------------------------------------
# the reference's only parallelism: contiguous candidate chunks checked in forked workers, results
# pickled back (reactive_planner.py:1084-1111).  Used by bench.py --impl reference.
# ------------------------------------------------------------------------------------------------
_PAR = {}


def _par_chunk(bounds):
    lo, hi = bounds
    cl, ct, dtl, prob = _PAR["cl"], _PAR["ct"], _PAR["dtl"], _PAR["prob"]
    ccosy = _ccosy_from(prob)
    out = []
    for k in range(lo, hi):
        ok, r, bs, st = check_kinematics_one(cl[k], ct[k], dtl[k], prob, ccosy)
        out.append((k, ok, r, st if ok else None))
    return out


def plan_grid_parallel(prob, workers):
    import math as _m
    import multiprocessing as mp
    cl, ct, dtl, dtt, behind = enumerate_grid(prob["t"], prob["lon"], prob["d"], prob["x0_lon"], prob["x0_lat"],
                                              prob["lon_mode"], prob["low_vel_mode"])
    n = len(dtl)
    _PAR.update(cl=cl, ct=ct, dtl=dtl, prob=prob)
    chunk = _m.ceil(n / workers)
    bounds = [(i * chunk, min(n, (i + 1) * chunk)) for i in range(workers) if i * chunk < n]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(bounds)) as pool:
        parts = pool.map(_par_chunk, bounds)
    feasible = {}
    for part in parts:
        for k, ok, r, st in part:
            if ok:
                feasible[k] = st
    cost = {k: evaluate_cost(st, prob["cost"]) for k, st in feasible.items()}
    order = sorted(feasible, key=lambda k: (cost[k], k))
    checker = build_checker(prob["obstacles"])
    winner, n_col = -1, 0
    for k in order:
        if candidate_collides(feasible[k], prob, checker) >= 0:
            n_col += 1
        else:
            winner = k
            break
    return {"n": n, "winner": winner, "n_infeasible_kinematics": n - len(feasible), "n_infeasible_collision": n_col,
            "winner_cost": cost.get(winner)}
