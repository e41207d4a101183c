"""Quartic / quintic end-state polynomials -- same classes and attributes as the reference's
``commonroad_rp/polynomial_trajectory.py``.

What changed underneath: the coefficient solve is the batched CUDA kernel (csrc/rp_kernels.cuh
``solve_kernel`` / ``coeff_kernel``; in-register LU with partial pivoting replacing
``np.linalg.solve`` at reference :315 and :355).  ``solve_batch`` solves any number of systems in one
launch; a single object constructed on its own solves lazily (n = 1) on first ``.coeffs`` access.
There is no host solver: without a CUDA device ``.coeffs`` raises.  The power-form evaluators
(``calc_position`` ...) are the public helpers user code calls on single objects; the planner itself
never calls them -- it evaluates all candidates in the fused kernel.
"""
import warnings
from abc import ABC, abstractmethod

import numpy as np

QUARTIC, QUINTIC = 0, 1


def solve_batch(kind, x0, xd, tau, engine=None) -> np.ndarray:
    """Solve n end-state problems on the device.  kind[i] = QUARTIC (xd[i][0] = target velocity, target
    acceleration 0) or QUINTIC (xd[i] = target position, velocity, acceleration).  Returns coeffs[n][6]."""
    if engine is None:
        from commonroad_rp_b200._device import default_engine
        engine = default_engine()
    return engine.solve_coeffs(kind, x0, xd, tau)


def _is_real(x):
    return isinstance(x, (int, float, np.integer, np.floating))


class PolynomialTrajectory(ABC):
    """Polynomial p(tau) = sum coeffs[k] * tau**k on [tau_0, tau_0 + delta_tau] (reference :17-271)."""

    _KIND = QUINTIC

    def __init__(self, tau_0=0, delta_tau=0, x_0=np.zeros([3, 1]), x_d=np.zeros([3, 1]), power=5, coeffs=None):
        super().__init__()
        self.tau_0 = tau_0
        self.delta_tau = delta_tau
        self.x_0 = x_0
        self.x_d = x_d
        self._cost = None
        assert isinstance(power, (int, np.integer)) and power >= 4, \
            '<PolynomialTrajectory/power>: power not valid! power={}'.format(power)
        self._power = power
        if power != 5 and power != 4:
            warnings.warn('Only power of 5 currently supported!')
        self._coeffs = None
        self._db = dict()
        if coeffs is not None:
            self.coeffs = coeffs

    @property
    def power(self) -> float:
        return self._power

    @property
    def coeffs(self) -> np.ndarray:
        if self._coeffs is None:
            self.coeffs = self.calc_coeffs()
        return self._coeffs

    @coeffs.setter
    def coeffs(self, co: np.ndarray):
        assert isinstance(co, np.ndarray) and len(co) == 6, \
            '<PolynomialTrajectory/coeffs>: coeffs length not valid! length={}'.format(len(co))
        self._coeffs = co
        self._db = dict()

    @property
    def delta_tau(self) -> float:
        return self._delta_tau

    @delta_tau.setter
    def delta_tau(self, tau: float):
        assert tau > 0, '<PolynomialTrajectory/delta_tau>: delta_tau not valid! delta_tau={}'.format(tau)
        self._delta_tau = tau

    @property
    def tau_0(self) -> float:
        return self._tau_0

    @tau_0.setter
    def tau_0(self, tau: float):
        assert _is_real(tau) and tau >= 0, '<PolynomialTrajectory/tau_0>: tau_0 not valid! tau_0={}'.format(tau)
        self._tau_0 = tau

    @property
    def x_0(self) -> np.ndarray:
        return self._x_0

    @x_0.setter
    def x_0(self, x: np.ndarray):
        self._x_0 = x

    @property
    def x_d(self) -> np.ndarray:
        return self._x_d

    @x_d.setter
    def x_d(self, x: np.ndarray):
        assert isinstance(x, np.ndarray)
        self._x_d = x

    @abstractmethod
    def calc_coeffs(self):
        """numpy array of the 6 coefficients (device solve)."""

    @property
    def cost(self) -> float:
        return self._cost

    @cost.setter
    def cost(self, cost):
        assert _is_real(cost) and cost >= 0, '<PolynomialTrajectory/cost>: cost not valid! cost={}'.format(cost)
        self._cost = cost

    # ---- evaluators (reference :171-271) ----
    def squared_jerk_integral(self, t):
        c = self.coeffs
        t2 = t * t
        t3 = t2 * t
        t4 = t3 * t
        t5 = t4 * t
        return (36 * c[3] * c[3] * t + 144 * c[3] * c[4] * t2 + 240 * c[3] * c[5] * t3 + 192 * c[4] * c[4] * t3 +
                720 * c[4] * c[5] * t4 + 720 * c[5] * c[5] * t5)

    def evaluate_state_at_tau(self, tau: float):
        """[p, p_dot, p_ddot] at tau, clamped into the definition interval."""
        if tau in self._db:
            return self._db[tau]
        key = tau
        rel = tau - self.tau_0
        if rel < 0:
            tau = self.tau_0
        elif rel > self.delta_tau:
            tau = self.delta_tau
        tau2 = np.power(tau, 2)
        tau3 = tau2 * tau
        tau4 = tau2 * tau2
        tau5 = tau3 * tau2
        result = np.array([self.calc_position(tau, tau2, tau3, tau4, tau5), self.calc_velocity(tau, tau2, tau3, tau4),
                           self.calc_acceleration(tau, tau2, tau3)])
        self._db[key] = result
        return result

    def calc_jerk(self, tau, tau2):
        c = self.coeffs
        return 6 * c[3] + 24 * c[4] * tau + 60 * c[5] * tau2

    def calc_acceleration(self, tau, tau2, tau3):
        c = self.coeffs
        return 2 * c[2] + 6 * c[3] * tau + 12 * c[4] * tau2 + 20 * c[5] * tau3

    def calc_velocity(self, tau, tau2, tau3, tau4):
        c = self.coeffs
        return c[1] + 2. * c[2] * tau + 3. * c[3] * tau2 + 4. * c[4] * tau3 + 5. * c[5] * tau4

    def calc_position(self, tau, tau2, tau3, tau4, tau5):
        c = self.coeffs
        return c[0] + c[1] * tau + c[2] * tau2 + c[3] * tau3 + c[4] * tau4 + c[5] * tau5


class QuinticTrajectory(PolynomialTrajectory):
    """Quintic with position / velocity / acceleration end conditions (reference :274-320)."""
    _KIND = QUINTIC

    def __init__(self, tau_0=0, delta_tau=0, x_0=np.zeros([3, 1]), x_d=np.zeros([3, 1]), coeffs=None):
        super().__init__(tau_0=tau_0, delta_tau=delta_tau, x_0=x_0, x_d=x_d, power=5, coeffs=coeffs)

    def calc_coeffs(self) -> np.ndarray:
        x0 = np.asarray(self.x_0, dtype=np.float64).reshape(3)
        xd = np.asarray(self.x_d, dtype=np.float64).reshape(3)
        return solve_batch([QUINTIC], x0[None, :], xd[None, :], [self.delta_tau])[0]


class QuarticTrajectory(PolynomialTrajectory):
    """Quartic with velocity end condition and zero end acceleration (reference :323-360)."""
    _KIND = QUARTIC

    def __init__(self, tau_0=0, delta_tau=0, x_0=np.zeros([3, 1]), x_d=np.zeros([2, 1]), coeffs=None):
        self._desired_velocity = x_d[0]
        super().__init__(tau_0=tau_0, delta_tau=delta_tau, x_0=x_0, x_d=x_d, power=4, coeffs=coeffs)

    def calc_coeffs(self) -> np.ndarray:
        x0 = np.asarray(self.x_0, dtype=np.float64).reshape(3)
        xd = np.array([float(np.asarray(self._desired_velocity).reshape(-1)[0]), 0.0, 0.0])
        return solve_batch([QUARTIC], x0[None, :], xd[None, :], [self.delta_tau])[0]
