"""EKF driver + ring buffer (oracle; test infrastructure only).  reference: src/x/ekf/ekf.cpp, src/x/ekf/state_buffer.cpp."""
import numpy as np

from .propagator import Propagator
from .state import K_INVALID

K_INVALID_INDEX = -1  # state_buffer.h:23


class StateBuffer:
    def __init__(self, size, default_state, time_margin=0.005):
        self.states = [default_state.copy() for _ in range(size)]
        self.tail = 0
        self.head = 0
        self.n_valid = 0
        self.time_margin = time_margin

    def __len__(self):
        return len(self.states)

    def next_idx(self, i):
        return (i + 1) % len(self.states)

    def prev_idx(self, i):
        return len(self.states) - 1 if i == 0 else i - 1

    def closest_idx(self, t):
        """reference: state_buffer.cpp:26-63."""
        if t > self.states[self.tail].time + self.time_margin:
            return K_INVALID_INDEX
        if t < self.states[self.head].time - self.time_margin:
            return K_INVALID_INDEX
        off = abs(t - self.states[self.tail].time)
        idx = self.prev_idx(self.tail)
        count = 1
        while abs(t - self.states[idx].time) < off and count < self.n_valid:
            off = abs(t - self.states[idx].time)
            idx = self.prev_idx(idx)
            count += 1
        return self.next_idx(idx)

    def enqueue_in_place(self):
        """reference: state_buffer.cpp:76-88."""
        self.tail = (self.tail + 1) % len(self.states)
        if self.n_valid < len(self.states):
            self.n_valid += 1
        else:
            self.head = (self.head + 1) % len(self.states)
        return self.states[self.tail]

    def reset_from_state(self, s):
        """reference: state_buffer.cpp:90-102."""
        for st in self.states:
            st.time = K_INVALID
        self.tail = self.head = 0
        self.n_valid = 1
        self.states[0] = s.copy()


class Ekf:
    """reference: src/x/ekf/ekf.cpp (single thread; lock()/unlock() are no-ops here)."""
    NOT_INIT, STANDBY, INIT = 0, 1, 2

    def __init__(self, updater, g, noise, buffer_sz, default_state, a_m_max=50.0, delta_seq_imu=1,
                 time_margin=0.005):
        self.updater = updater
        self.prop = Propagator(g, noise)
        self.buf = StateBuffer(buffer_sz, default_state, time_margin)
        self.a_m_max = a_m_max
        self.delta_seq_imu = delta_seq_imu
        self.status = self.NOT_INIT
        self.last_seq = 0

    def initialize_from_state(self, s):  # ekf.cpp:43-64
        if len(self.buf) == 0:
            raise RuntimeError("The EKF state buffer must have non-zero size.")
        b = self.buf.states[0]
        if (s.p_array.shape != b.p_array.shape or s.q_array.shape != b.q_array.shape
                or s.f_array.shape != b.f_array.shape or s.cov.shape != b.cov.shape):
            raise ValueError("init_bfr_mismatch")
        self.buf.reset_from_state(s)
        self.status = self.STANDBY

    def process_imu(self, t, seq, w_m, a_m):
        """reference: ekf.cpp:66-140."""
        if self.status == self.NOT_INIT:
            return None
        last = self.buf.states[self.buf.tail]
        a_m = np.asarray(a_m, dtype=float)
        if self.status == self.STANDBY:
            if np.linalg.norm(a_m) < self.a_m_max:
                last.set_imu(t, seq, w_m, a_m)
                self.last_seq = seq
                self.status = self.INIT
                return last
            return None
        if t <= last.time:
            return None
        self.last_seq = seq
        a_s = a_m if np.linalg.norm(a_m) < self.a_m_max else last.a_m
        nxt = self.buf.enqueue_in_place()
        nxt.set_imu(t, seq, w_m, a_s)
        self.prop.propagate_state(last, nxt)
        self.prop.propagate_covariance(last, nxt)
        return nxt

    def _repropagate(self, state, idx):
        """reference: ekf.cpp:227-255."""
        if self.buf.states[idx].time != state.time:
            return False
        self.buf.states[idx] = state
        cur = idx
        nxt = self.buf.next_idx(idx)
        while cur != self.buf.tail:
            self.prop.propagate_state(self.buf.states[cur], self.buf.states[nxt])
            self.prop.propagate_covariance(self.buf.states[cur], self.buf.states[nxt])
            cur = nxt
            nxt = self.buf.next_idx(nxt)
        return True

    def process_update_measurement(self):
        """reference: ekf.cpp:179-213."""
        if self.status == self.NOT_INIT:
            return None
        idx = self.buf.closest_idx(self.updater.get_time())
        if idx == K_INVALID_INDEX:
            return None
        s = self.buf.states[idx].copy()
        self.updater.update(s)
        return s if self._repropagate(s, idx) else None

    def process_others_measurement(self, t, collaborative_update):
        """reference: ekf.cpp:143-176; `collaborative_update(state)` restates Updater::collaborativeUpdate."""
        if self.status == self.NOT_INIT:
            return None
        idx = self.buf.closest_idx(t)
        if idx == K_INVALID_INDEX:
            return None
        s = self.buf.states[idx].copy()
        collaborative_update(s)
        return s if self._repropagate(s, idx) else None
