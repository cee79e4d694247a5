"""IMU state + covariance propagation (oracle; test infrastructure only).  reference: src/x/ekf/propagator.cpp."""
from dataclasses import dataclass

import numpy as np

from .qd_poly import qd_poly
from .quat import omega, rot_raw, skew, qnormalized
from .state import K_CORE


@dataclass
class ImuNoise:
    """reference: include/x/common/types.h:65-85.  NOTE the in-tree default for n_ba is the octal literal
    `00013` (= 11.0); real runs overwrite it from YAML (vio.cpp:181-185).  The oracle default is the intended value."""
    n_w: float = 0.0083
    n_bw: float = 0.00083
    n_a: float = 0.0013
    n_ba: float = 0.00013


class Propagator:
    def __init__(self, g=(0.0, 0.0, -9.81), noise=None):
        self.g = np.array(g, dtype=float)
        self.noise = noise or ImuNoise()

    def propagate_state(self, s0, s1):
        """reference: propagator.cpp:30-51."""
        s1.set_static_states_from(s0)
        w1, a1 = s1.unbiased_imu()
        w0, a0 = s0.unbiased_imu()
        dt = s1.time - s0.time
        dq = self.quaternion_integrator(w0, w1, dt)
        s1.q = qnormalized(dq @ s0.q)  # coeffs (x,y,z,w) order, propagator.cpp:42-43
        dv = (rot_raw(s1.q) @ a1 + rot_raw(s0.q) @ a0) / 2.0
        s1.v = s0.v + (dv + self.g) * dt
        s1.p = s0.p + (s1.v + s0.v) / 2.0 * dt

    @staticmethod
    def quaternion_integrator(w0, w1, dt):
        """reference: propagator.cpp:74-98 (4th-order series + commutator term)."""
        om1, om0 = omega(w1), omega(w0)
        om_mean = omega((w1 + w0) / 2.0)
        a = om_mean * 0.5 * dt
        a_k = a.copy()
        mat_exp = np.eye(4)
        fac = 1
        for k in range(1, 5):
            fac *= k
            mat_exp = mat_exp + a_k / fac
            a_k = a_k @ a
        return mat_exp + 1.0 / 48.0 * (om1 @ om0 - om0 @ om1) * dt * dt

    @staticmethod
    def discrete_state_transition(dt, w, a, q):
        """reference: propagator.cpp:100-164."""
        w_x, a_x = skew(w), skew(a)
        eye3 = np.eye(3)
        c_q = rot_raw(q)
        dt2 = dt * dt * 0.5
        dt3 = dt2 * dt / 3.0
        dt4 = dt3 * dt * 0.25
        dt5 = dt4 * dt * 0.2
        cqa = c_q @ a_x
        ww = w_x @ w_x
        A = cqa @ (-dt2 * eye3 + dt3 * w_x - dt4 * ww)
        B = cqa @ (dt3 * eye3 - dt4 * w_x + dt5 * ww)
        D = -A
        E = eye3 - dt * w_x + dt2 * ww
        F = -dt * eye3 + dt2 * w_x - dt3 * ww
        Cm = cqa @ F
        f_d = np.eye(K_CORE)
        f_d[0:3, 3:6] = dt * eye3
        f_d[0:3, 6:9] = A
        f_d[0:3, 9:12] = B
        f_d[0:3, 12:15] = -c_q * dt2
        f_d[3:6, 6:9] = Cm
        f_d[3:6, 9:12] = D
        f_d[3:6, 12:15] = -c_q * dt
        f_d[6:9, 6:9] = E
        f_d[6:9, 9:12] = F
        return f_d

    def discrete_process_noise(self, dt, q, w, a):
        """reference: propagator.cpp:207-840 (symbolic polynomial; see oracle/tools/gen_qd.py)."""
        n = self.noise
        return qd_poly(dt, rot_raw(q), w, a, n.n_w, n.n_bw, n.n_a, n.n_ba)

    def propagate_covariance(self, s0, s1):
        """reference: propagator.cpp:53-72 and :166-205."""
        w1, a1 = s1.unbiased_imu()
        dt = s1.time - s0.time
        f_d = self.discrete_state_transition(dt, w1, a1, s1.q)
        q_d = self.discrete_process_noise(dt, s1.q, w1, a1)
        c0 = s0.cov
        c1 = np.empty_like(c0) if s1.cov.shape != c0.shape else s1.cov
        k = K_CORE
        c1[:k, :k] = f_d @ c0[:k, :k] @ f_d.T + q_d
        c1[:k, k:] = f_d @ c0[:k, k:]
        c1[k:, :k] = c0[k:, :k] @ f_d.T  # computed independently of P_iv, propagator.cpp:197-203
        c1[k:, k:] = c0[k:, k:]  # full P_vv copy, propagator.cpp:204
        s1.cov = c1
        return f_d, q_d
