"""Synthetic workloads of the shapes BASELINE.json names.

Pure data generation (numpy only, no planner arithmetic): the same dictionaries feed the GPU
planner, the oracle and the reference-under-shims harness, so that all three see identical
inputs.  A "scenario dict" holds plain arrays:

    ref_path        (L, 2)  raw reference polyline handed to ``set_reference_path``
    static_boxes    (ns, 5) cx, cy, theta, length, width
    dyn_t0          (nd,)   first time index of each dynamic obstacle
    dyn_states      list of (K_i, 3) arrays: cx, cy, theta at time index t0 + k
    dyn_lw          (nd, 2) length, width
    boundary_boxes  (nb, 5) cx, cy, theta, half_length, half_width  (road boundary OBBs)
    boundary_tris   (nt, 6) x1, y1, x2, y2, x3, y3                   (road boundary triangles)
"""
import numpy as np


def sine_path(amplitude=20.0, wavelength=40.0, length=300):
    x = np.arange(0, length, 1.0)
    y = amplitude * np.sin(x / wavelength) if amplitude != 0.0 else np.zeros_like(x)
    return np.stack([x, y], axis=1)


def _frame_on_sine(x, amplitude, wavelength):
    """point, heading and unit normal of y = A sin(x / lambda) at abscissa x."""
    y = amplitude * np.sin(x / wavelength)
    slope = amplitude / wavelength * np.cos(x / wavelength)
    theta = np.arctan(slope)
    nx, ny = -np.sin(theta), np.cos(theta)
    return y, theta, nx, ny


def make_scenario(seed=0, amplitude=20.0, wavelength=40.0, length=300, n_dynamic=8, n_static=2,
                  n_dyn_steps=100, boundary_offset=5.25, boundary=True, static_x=(60.0, 140.0),
                  static_offset=0.0, boundary_kind="boxes", boundary_depth=1.0):
    """Config-4/5 style scenario: sinusoidal reference path, ``n_dynamic`` 5.0x2.0 cars on
    +-3.5 m lateral offsets at speeds U(5, 20) for ``n_dyn_steps`` steps, ``n_static`` boxes on
    the path, and a road boundary of thin OBBs at +-``boundary_offset``."""
    rng = np.random.default_rng(seed)
    ref_path = sine_path(amplitude, wavelength, length)
    dt = 0.1

    dyn_t0, dyn_states, dyn_lw = [], [], []
    for k in range(n_dynamic):
        off = 3.5 if k % 2 == 0 else -3.5
        speed = rng.uniform(5.0, 20.0)
        x_start = rng.uniform(10.0, 0.6 * length)
        steps = np.arange(n_dyn_steps + 1)
        xs = x_start + speed * dt * steps
        y, th, nx, ny = _frame_on_sine(xs, amplitude, wavelength)
        dyn_states.append(np.stack([xs + off * nx, y + off * ny, th], axis=1))
        dyn_t0.append(0)
        dyn_lw.append((5.0, 2.0))

    static_boxes = []
    for k in range(n_static):
        xs = static_x[k] if k < len(static_x) else rng.uniform(30.0, length - 30.0)
        y, th, nx, ny = _frame_on_sine(np.float64(xs), amplitude, wavelength)
        static_boxes.append((xs + static_offset * nx, y + static_offset * ny, th, 4.5, 2.0))

    boundary_boxes, boundary_tris = [], []
    if boundary and boundary_kind == "tris":
        # the same two border lines as a triangulated band: per segment the quad between the border and the line
        # ``boundary_depth`` farther out, split along its diagonal (road boundary as triangles)
        xs = np.arange(0, length, 1.0)
        y, th, nx, ny = _frame_on_sine(xs, amplitude, wavelength)
        for side in (+1.0, -1.0):
            ix, iy = xs + side * boundary_offset * nx, y + side * boundary_offset * ny
            ox, oy = xs + side * (boundary_offset + boundary_depth) * nx, y + side * (boundary_offset + boundary_depth) * ny
            for i in range(len(xs) - 1):
                boundary_tris.append((ix[i], iy[i], ix[i + 1], iy[i + 1], ox[i + 1], oy[i + 1]))
                boundary_tris.append((ix[i], iy[i], ox[i + 1], oy[i + 1], ox[i], oy[i]))
    elif boundary:
        xs = np.arange(0, length, 1.0)
        y, th, nx, ny = _frame_on_sine(xs, amplitude, wavelength)
        for side in (+1.0, -1.0):
            px = xs + side * boundary_offset * nx
            py = y + side * boundary_offset * ny
            for i in range(len(xs) - 1):
                dx, dy = px[i + 1] - px[i], py[i + 1] - py[i]
                boundary_boxes.append((0.5 * (px[i] + px[i + 1]), 0.5 * (py[i] + py[i + 1]),
                                       np.arctan2(dy, dx), 0.5 * np.hypot(dx, dy), 0.05))

    return {
        "ref_path": ref_path,
        "static_boxes": np.asarray(static_boxes, dtype=np.float64).reshape(-1, 5),
        "dyn_t0": np.asarray(dyn_t0, dtype=np.int64),
        "dyn_states": dyn_states,
        "dyn_lw": np.asarray(dyn_lw, dtype=np.float64).reshape(-1, 2),
        "boundary_boxes": np.asarray(boundary_boxes, dtype=np.float64).reshape(-1, 5),
        "boundary_tris": np.asarray(boundary_tris, dtype=np.float64).reshape(-1, 6),
    }


def dense_grid(n_t=32, n_v=64, n_d=64, v_lo=5.0, v_hi=25.0, d_lo=-3.0, d_hi=3.0, t_first_step=29, dt=0.1):
    """Config-4 sample sets: t = dt*(t_first_step + k); v, d linspaces.  d0 is chosen ON the d grid
    so that ``set(d) | {d0}`` keeps n_d entries."""
    t = dt * (t_first_step + np.arange(n_t))
    v = np.linspace(v_lo, v_hi, n_v)
    d = np.linspace(d_lo, d_hi, n_d)
    return t, v, d, float(d[n_d // 2])


def scenario_seeded(scenario_id):
    """Config-5 per-scenario draw: A~U(0,20), lambda~U(30,120), s_dot0~U(5,25), d0~U(-1,1)."""
    rng = np.random.default_rng(1000003 + int(scenario_id))
    amplitude = rng.uniform(0.0, 20.0)
    wavelength = rng.uniform(30.0, 120.0)
    s_dot0 = rng.uniform(5.0, 25.0)
    d0 = rng.uniform(-1.0, 1.0)
    scn = make_scenario(seed=int(scenario_id), amplitude=amplitude, wavelength=wavelength)
    return scn, float(s_dot0), float(d0)


def add_crossing_obstacle(scn, speed=40.0, phase=2.0, x_c=24.0, y_path=0.0, n_steps=40, size=(1.0, 1.0)):
    """A small box crossing the road along +y at x = x_c with ``speed`` m/s: consecutive discrete poses straddle the
    ego, so only the continuous collision check (OBB-sum hulls of consecutive poses) sees it."""
    k = np.arange(n_steps)
    y = y_path - 6.0 + speed * 0.1 * (k - 10) + phase
    st = np.stack([np.full(n_steps, x_c), y, np.full(n_steps, np.pi / 2)], axis=1)
    out = dict(scn)
    out["dyn_t0"] = list(scn["dyn_t0"]) + [0]
    out["dyn_states"] = list(scn["dyn_states"]) + [st]
    out["dyn_lw"] = [tuple(x) for x in scn["dyn_lw"]] + [tuple(size)]
    return out
