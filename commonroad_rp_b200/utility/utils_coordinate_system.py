"""Host mirror of ``commonroad_rp.utility.utils_coordinate_system`` (reference file of the same name).

Scenario-static set-up only (once per reference path, SURVEY.md section 2: path *construction* is out
of scope and stays host numpy); the per-(candidate, step) consumers -- angle interpolation, table
lookups and (s, d) -> (x, y) -- run on the device (csrc/rp_device.cuh).  ``CoordinateSystem`` keeps
the reference's attribute names (``reference, ref_pos, ref_theta, ref_curv, ref_curv_d, ccosy``) and
adds ``device_tables()`` for ``rp_ctx_set_reference``.

commonroad_dc is not installed here; ``PolylineFrame`` is this package's own curvilinear frame with
the pseudo-normal projection of pycrccosy (semantics per SURVEY.md App. D#1, parity unpinned).  A real
``pycrccosy.CurvilinearCoordinateSystem`` passed as ``ccosy=`` is accepted as long as it offers
``reference_path()``; its polyline is re-framed here so that host and device agree.
"""
import logging
import math

import numpy as np

logger = logging.getLogger("RP_LOGGER")

_TWO_PI = 2.0 * np.pi


def make_valid_orientation(angle: float) -> float:
    """commonroad.common.util.make_valid_orientation: fold into [-2pi, 2pi]."""
    while angle > _TWO_PI:
        angle -= _TWO_PI
    while angle < -_TWO_PI:
        angle += _TWO_PI
    return angle


def interpolate_angle(x: float, x1: float, x2: float, y1: float, y2: float) -> float:
    """Linear angle interpolation followed by make_valid_orientation (reference :25-43; the
    shortest-arc selection is commented out there, so none is applied here either)."""
    return make_valid_orientation((y2 - y1) * (x - x1) / (x2 - x1) + y1)


# ---- polyline helpers (commonroad_dc.geometry.util semantics) ---------------------------------------
def compute_pathlength_from_polyline(polyline: np.ndarray) -> np.ndarray:
    p = np.asarray(polyline, dtype=np.float64)
    seg = np.sqrt(np.sum(np.diff(p, axis=0) ** 2, axis=1))
    return np.concatenate([[0.0], np.cumsum(seg)])


def compute_orientation_from_polyline(polyline: np.ndarray) -> np.ndarray:
    p = np.asarray(polyline, dtype=np.float64)
    dxy = np.diff(p, axis=0)
    th = np.arctan2(dxy[:, 1], dxy[:, 0])
    return np.concatenate([th, th[-1:]])


def compute_curvature_from_polyline(polyline: np.ndarray) -> np.ndarray:
    p = np.asarray(polyline, dtype=np.float64)
    pl = compute_pathlength_from_polyline(p)
    xd = np.gradient(p[:, 0], pl)
    yd = np.gradient(p[:, 1], pl)
    xdd = np.gradient(xd, pl)
    ydd = np.gradient(yd, pl)
    return (xd * ydd - xdd * yd) / ((xd ** 2 + yd ** 2) ** 1.5)


def resample_polyline(polyline: np.ndarray, step: float = 2.0) -> np.ndarray:
    """Fixed arc-length resampling that keeps the end point."""
    p = np.asarray(polyline, dtype=np.float64)
    if len(p) < 2:
        return p.copy()
    out = [p[0]]
    want = step
    k = 0
    seg = float(np.linalg.norm(p[1] - p[0]))
    while k < len(p) - 1:
        if want >= seg:
            want -= seg
            k += 1
            if k > len(p) - 2:
                break
            seg = float(np.linalg.norm(p[k + 1] - p[k]))
        else:
            w = want / seg
            out.append((1.0 - w) * p[k] + w * p[k + 1])
            want += step
    if np.linalg.norm(out[-1] - p[-1]) >= 1e-6:
        out.append(p[-1])
    return np.array(out)


def chaikins_corner_cutting(polyline: np.ndarray, refinements: int = 1) -> np.ndarray:
    pts = np.asarray(polyline, dtype=np.float64)
    for _ in range(refinements):
        dup = pts.repeat(2, axis=0)
        nb = np.empty_like(dup)
        nb[0] = dup[0]
        nb[2::2] = dup[1:-1:2]
        nb[1:-1:2] = dup[2::2]
        nb[-1] = dup[-1]
        pts = dup * 0.75 + nb * 0.25
    return pts


def extrapolate_ref_path(reference_path: np.ndarray, resample_step: float = 2.0) -> np.ndarray:
    """Extend the end of the reference path linearly (reference :46-57)."""
    line = np.poly1d(np.polyfit(reference_path[-2:, 0], reference_path[-2:, 1], 1))
    x = 2.3 * reference_path[-1, 0] - reference_path[-2, 0]
    return resample_polyline(np.concatenate((reference_path, np.array([[x, line(x)]])), axis=0), step=resample_step)


def preprocess_ref_path(ref_path: np.ndarray, resample_step: float = 1.0, max_curv_desired: float = 0.01):
    """Corner cutting until the curvature bound holds (reference :60-71)."""
    path = np.array(ref_path, dtype=np.float64, copy=True)
    max_curv = max_curv_desired + 0.2
    while max_curv > max_curv_desired:
        path = resample_polyline(chaikins_corner_cutting(path), resample_step)
        max_curv = max(compute_curvature_from_polyline(path))
    return path


def smooth_ref_path(ref_path: np.ndarray, smoothing_factor=0.0, resample_step: float = 1.0):
    """Cubic-spline smoothing to 200 points, then 1 m resampling (reference :74-83)."""
    from scipy.interpolate import splev, splprep
    logger.info("Smoothing reference path...")
    tck, u = splprep(ref_path.T, u=None, k=3, s=smoothing_factor)
    u_new = np.linspace(u.min(), u.max(), 200)
    x_new, y_new = splev(u_new, tck, der=0)
    return resample_polyline(np.array([x_new, y_new]).transpose(), resample_step)


def _dedupe(points: np.ndarray) -> np.ndarray:
    _, idx = np.unique(points, axis=0, return_index=True)
    return points[np.sort(idx)]


class PolylineFrame:
    """Curvilinear frame over a polyline with pseudo-normal projection (pycrccosy semantics,
    SURVEY.md App. D#1).  The stored polyline carries one extra vertex at each end, ``eps2`` beyond
    the first / last segment; vertex pseudo-tangents are the normalised chords p[i+1]-p[i-1] (end
    vertices: adjacent segment)."""

    def __init__(self, reference_path, default_projection_domain_limit: float = 20.0, eps: float = 0.1,
                 eps2: float = 1e-4, extend: bool = True):
        ref = np.asarray(reference_path, dtype=np.float64)
        if ref.ndim != 2 or ref.shape[1] != 2 or ref.shape[0] < 3:
            raise ValueError("<PolylineFrame>: reference path must be an (n >= 3, 2) array")
        if extend:
            head = ref[1] - ref[0]
            tail = ref[-1] - ref[-2]
            ref = np.vstack([ref[0] - eps2 * head / np.linalg.norm(head), ref,
                             ref[-1] + eps2 * tail / np.linalg.norm(tail)])
        self._path = ref
        self._limit = float(default_projection_domain_limit)
        self._s = compute_pathlength_from_polyline(ref)
        chord = np.empty_like(ref)
        chord[0] = ref[1] - ref[0]
        chord[-1] = ref[-1] - ref[-2]
        chord[1:-1] = ref[2:] - ref[:-2]
        chord /= np.linalg.norm(chord, axis=1)[:, None]
        self._normals = np.stack([-chord[:, 1], chord[:, 0]], axis=1)

    def reference_path(self):
        return [row.copy() for row in self._path]

    @property
    def path(self) -> np.ndarray:
        return self._path

    @property
    def pathlength(self) -> np.ndarray:
        return self._s

    @property
    def normals(self) -> np.ndarray:
        return self._normals

    @property
    def projection_domain_limit(self) -> float:
        return self._limit

    def convert_to_cartesian_coords(self, s: float, d: float) -> np.ndarray:
        S = self._s
        if not (S[0] <= s <= S[-1]) or not (abs(d) <= self._limit):
            raise ValueError("<PolylineFrame>: (s, d) outside of the projection domain")
        j = min(int(np.searchsorted(S, s, side="right")) - 1, len(S) - 2)
        lam = (s - S[j]) / (S[j + 1] - S[j])
        base = self._path[j] + lam * (self._path[j + 1] - self._path[j])
        pn = self._normals[j] + lam * (self._normals[j + 1] - self._normals[j])
        return base + d * pn

    def convert_to_curvilinear_coords(self, x: float, y: float) -> np.ndarray:
        """Inverse pseudo-normal map: per segment the foot parameter solves a quadratic; the solution
        with the smallest |d| wins.  Raises ValueError outside the projection domain."""
        p = np.array([x, y], dtype=np.float64)
        p0 = self._path[:-1]
        e = self._path[1:] - p0
        m = self._normals[:-1]
        dn = self._normals[1:] - m
        a = p[None, :] - p0
        cross = lambda u, v: u[:, 0] * v[:, 1] - u[:, 1] * v[:, 0]
        qa, qb, qc = -cross(e, dn), cross(a, dn) - cross(e, m), cross(a, m)
        best = None
        for j in range(len(p0)):
            roots = []
            if abs(qa[j]) < 1e-14:
                if qb[j] != 0.0:
                    roots.append(-qc[j] / qb[j])
            else:
                disc = qb[j] * qb[j] - 4.0 * qa[j] * qc[j]
                if disc >= 0.0:
                    # cancellation-free form of (-qb +- sqrt(disc)) / (2 qa) (qa is tiny on gently curved paths)
                    r = math.sqrt(disc)
                    qq = -0.5 * (qb[j] + math.copysign(r, qb[j]))
                    if qq == 0.0:
                        roots += [0.0, 0.0]
                    elif qb[j] >= 0.0:
                        roots += [qc[j] / qq, qq / qa[j]]
                    else:
                        roots += [qq / qa[j], qc[j] / qq]
            for lam in roots:
                if -1e-12 <= lam <= 1.0 + 1e-12:
                    lam = min(max(lam, 0.0), 1.0)
                    pn = m[j] + lam * dn[j]
                    dist = float(np.dot(p - (p0[j] + lam * e[j]), pn) / np.dot(pn, pn))
                    if abs(dist) <= self._limit and (best is None or abs(dist) < abs(best[1])):
                        best = (float(self._s[j] + lam * (self._s[j + 1] - self._s[j])), dist)
        if best is None:
            raise ValueError("<PolylineFrame>: point outside of the projection domain")
        return np.array(best)


class CoordinateSystem:
    """Same surface as the reference's ``CoordinateSystem`` (:86-178)."""

    def __init__(self, reference: np.ndarray = None, ccosy=None, smooth_reference: bool = True):
        if ccosy is None:
            assert reference is not None, '<CoordinateSystem>: Please provide a reference path OR a ' \
                                          'CurvilinearCoordinateSystem object.'
            reference = _dedupe(np.asarray(reference, dtype=np.float64))
            if smooth_reference:
                reference = _dedupe(smooth_ref_path(reference))
            self.reference = reference
        else:
            self.ccosy = ccosy

        self._ref_pos = compute_pathlength_from_polyline(self.reference)
        self._ref_curv = compute_curvature_from_polyline(self.reference)
        self._ref_theta = np.unwrap(compute_orientation_from_polyline(self.reference))
        self._ref_curv_d = np.gradient(self._ref_curv, self._ref_pos)
        self._ref_curv_dd = np.gradient(self._ref_curv_d, self._ref_pos)

    @property
    def reference(self) -> np.ndarray:
        """reference polyline as the curvilinear frame stores it (with its two extension vertices)"""
        return self._reference

    @reference.setter
    def reference(self, reference):
        self._ccosy = PolylineFrame(reference)
        self._reference = np.asarray(self._ccosy.reference_path())

    @property
    def ccosy(self):
        return self._ccosy

    @ccosy.setter
    def ccosy(self, ccosy):
        if isinstance(ccosy, PolylineFrame):
            self._ccosy = ccosy
        else:
            # foreign frame object (e.g. pycrccosy): adopt its stored polyline without re-extending it
            self._ccosy = PolylineFrame(np.asarray(ccosy.reference_path()), extend=False)
        self._reference = np.asarray(self._ccosy.reference_path())

    @property
    def ref_pos(self) -> np.ndarray:
        return self._ref_pos

    @property
    def ref_curv(self) -> np.ndarray:
        return self._ref_curv

    @property
    def ref_curv_d(self) -> np.ndarray:
        return self._ref_curv_d

    @property
    def ref_cruv_dd(self) -> np.ndarray:      # (sic) the reference spells it this way (:158)
        return self._ref_curv_dd

    @property
    def ref_theta(self) -> np.ndarray:
        return self._ref_theta

    def convert_to_cartesian_coords(self, s: float, d: float):
        """(s, d) -> (x, y); None outside the projection domain (reference :167-174)."""
        try:
            return self._ccosy.convert_to_cartesian_coords(s, d)
        except Exception:
            return None

    def convert_to_curvilinear_coords(self, x: float, y: float) -> np.ndarray:
        return self._ccosy.convert_to_curvilinear_coords(x, y)

    def device_tables(self) -> dict:
        """Arguments of ``rp_ctx_set_reference`` (include/rp_b200.h)."""
        return {"ref_pos": self._ref_pos, "ref_theta": self._ref_theta, "ref_curv": self._ref_curv,
                "ref_curv_d": self._ref_curv_d, "path_xy": self._ccosy.path, "path_s": self._ccosy.pathlength,
                "path_normals": self._ccosy.normals, "proj_limit": self._ccosy.projection_domain_limit}

    def plot_reference_states(self):
        from matplotlib import pyplot as plt
        fig, axes = plt.subplots(4, 1, figsize=(7, 7.5))
        fig.suptitle("Reference path states")
        for ax, (vals, label) in zip(axes, ((self.ref_theta, "theta_ref"), (self.ref_curv, "kappa_ref"),
                                            (self.ref_curv_d, "kappa_dot_ref"), (self.ref_cruv_dd, "kappa_dot_dot_ref"))):
            ax.plot(self.ref_pos, vals, color="k")
            ax.set_xlabel("s")
            ax.set_ylabel(label)
        fig.tight_layout()
        plt.show()
