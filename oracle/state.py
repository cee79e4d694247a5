"""State container (oracle; test infrastructure only).  reference: include/x/ekf/state.h:36-337, src/x/ekf/state.cpp."""
import copy

import numpy as np

from .quat import qmul, qnormalized, rot, small_angle_quat

K_CORE = 15  # reference: include/x/common/types.h:39-47 (kSizeCoreErr)
K_INVALID = -1.0  # reference: include/x/common/types.h:90


class State:
    """Error-state order [p v theta b_w b_a | p_array(3M) | theta_array(3M) | f_array(3F)] (state.cpp:201-214).

    Quaternions are (x,y,z,w).  `cov` is always the full N x N matrix (state.cpp:23-37).
    """

    def __init__(self, n_poses=0, n_features=0):
        self.time = K_INVALID
        self.seq = 0
        self.p = np.zeros(3)
        self.v = np.zeros(3)
        self.q = np.array([0.0, 0.0, 0.0, 1.0])
        self.b_w = np.zeros(3)
        self.b_a = np.zeros(3)
        self.p_array = np.zeros(3 * n_poses)
        self.q_array = np.zeros(4 * n_poses)
        self.f_array = np.zeros(3 * n_features)
        n = K_CORE + 6 * n_poses + 3 * n_features
        self.cov = np.eye(n)  # state.cpp:33
        self.q_ic = np.array([0.0, 0.0, 0.0, 1.0])
        self.p_ic = np.zeros(3)
        self.w_m = np.zeros(3)
        self.a_m = np.zeros(3)

    def copy(self):
        return copy.deepcopy(self)

    # state.cpp:167-175
    def n_poses_max(self):
        return self.p_array.size // 3

    def n_features_max(self):
        return self.f_array.size // 3

    def n_error_states(self):
        return K_CORE + self.p_array.size + (self.q_array.size // 4) * 3 + self.f_array.size

    def set_imu(self, time, seq, w_m, a_m):  # state.cpp:145-151
        self.time, self.seq = float(time), int(seq)
        self.w_m, self.a_m = np.array(w_m, dtype=float), np.array(a_m, dtype=float)

    def set_static_states_from(self, o):  # state.cpp:153-161
        self.b_w, self.b_a = o.b_w.copy(), o.b_a.copy()
        self.q_ic, self.p_ic = o.q_ic.copy(), o.p_ic.copy()
        self.p_array, self.q_array, self.f_array = o.p_array.copy(), o.q_array.copy(), o.f_array.copy()

    def unbiased_imu(self):  # state.cpp:177-182
        return self.w_m - self.b_w, self.a_m - self.b_a

    def camera_orientation(self):  # state.cpp:193-195
        return qmul(qnormalized(self.q), qnormalized(self.q_ic))

    def camera_position(self):  # state.cpp:189-191
        return self.p + rot(self.q) @ self.p_ic

    def dynamic_states(self):  # state.cpp:87-99
        return np.concatenate([self.p, self.v, self.q, self.b_w, self.b_a])

    def correct(self, d):
        """reference: src/x/ekf/state.cpp:197-249."""
        d = np.asarray(d, dtype=float).reshape(-1)
        assert d.size == self.n_error_states()
        n_p = self.p_array.size
        n_f = self.f_array.size
        self.p = self.p + d[0:3]
        self.v = self.v + d[3:6]
        self.b_w = self.b_w + d[9:12]
        self.b_a = self.b_a + d[12:15]
        self.p_array = self.p_array + d[K_CORE:K_CORE + n_p]
        self.f_array = self.f_array + d[K_CORE + 2 * n_p:K_CORE + 2 * n_p + n_f]
        self.q = qnormalized(qmul(self.q, small_angle_quat(d[6:9])))
        dth = d[K_CORE + n_p:K_CORE + 2 * n_p]
        for i in range(n_p // 3):
            qi = self.q_array[4 * i:4 * i + 4]
            qi = qmul(qi, small_angle_quat(dth[3 * i:3 * i + 3]))
            # Eigen normalize(): divides by norm (a zero quaternion of an unused slot stays NaN-free only
            # if delta is zero there; the reference has the same hazard) -- guard the 0/0 case.
            nrm = np.sqrt(qi @ qi)
            self.q_array[4 * i:4 * i + 4] = qi / nrm if nrm > 0.0 else qi
