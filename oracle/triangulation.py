"""Feature triangulation (oracle; test infrastructure only).  reference: src/x/vision/triangulation.cpp."""
import numpy as np

from .quat import rot


def dlt_two_view(P1, P2, z1, z2):
    """cv::triangulatePoints for one point (third-party: OpenCV >= 3.3.1, unpinned, CMakeLists.txt:101-110;
    call site triangulation.cpp:93).  Published algorithm (calib3d/triangulate.cpp): stack
    x*P[2]-P[0], y*P[2]-P[1] of both views into a 4x4 A and take the right singular vector of the
    smallest singular value.  Cross-checked against cv2.triangulatePoints in tests/test_oracle.py."""
    A = np.empty((4, 4))
    for j, (P, z) in enumerate(((P1, z1), (P2, z2))):
        A[2 * j] = z[0] * P[2] - P[0]
        A[2 * j + 1] = z[1] * P[2] - P[1]
    _, _, vt = np.linalg.svd(A)
    return vt[3]


class Triangulation:
    """reference: triangulation.cpp:48-79 (ctor from attitude/translation lists)."""

    def __init__(self, quats, poss, max_iter=10, term=1e-5):
        self.n_poses = len(quats)
        self.max_iter = max_iter
        self.term = term
        self.rotations = [rot(q).T for q in quats]
        self.positions = [np.asarray(p, dtype=float) for p in poss]
        self.projs = [np.hstack([R, (-R @ p)[:, None]]) for R, p in zip(self.rotations, self.positions)]

    def triangulate_gn(self, track):
        """reference: triangulation.cpp:102-206.  `track` is an (L,2) array of normalised coordinates.

        Waived quirk: the reference indexes the *track* with the pose indices i1,i2 (`track[i1]`,
        `track[i2]`, :112-113), which is only in-bounds when the track spans the whole pose list
        (always true for MsckfUpdate, msckf_update.cpp:145-160; out-of-bounds UB for shorter
        MSCKF-SLAM tracks with the shared triangulator, vio_updater.cpp:280).  The oracle uses the
        first and last observation, the evident intent.
        """
        track = np.asarray(track, dtype=float)
        n_obs = track.shape[0]
        i2 = self.n_poses - 1
        i1 = i2 - n_obs + 1
        pt_h = dlt_two_view(self.projs[i1], self.projs[i2], track[0], track[-1])
        pt_xyz = pt_h[:3] / pt_h[3]
        pt_c2 = self.projs[i2] @ np.append(pt_xyz, 1.0)
        alpha = pt_c2[0] / pt_c2[2]
        beta = pt_c2[1] / pt_c2[2]
        rho = 1.0 / pt_c2[2]
        rot_a = self.rotations[i2]
        p_a = self.positions[i2]
        n_meas = 2 * self.n_poses
        r_norm_last, r_norm = 1000.0, 100.0
        it = 0
        while r_norm_last - r_norm > self.term:
            it += 1
            if it > self.max_iter:
                break
            r = np.zeros(n_meas)
            J = np.zeros((n_meas, 3))
            for i in range(i1, i2 + 1):
                R = self.rotations[i]
                dR = R @ rot_a.T
                dp = R @ p_a - R @ self.positions[i]
                k = i - i1
                h_i = dR @ np.array([alpha, beta, 1.0]) + rho * dp
                h = np.array([h_i[0] / h_i[2], h_i[1] / h_i[2]])
                r[2 * k:2 * k + 2] = track[k] - h
                j0 = np.column_stack([dR[:, 0], dR[:, 1], dp])
                j1 = np.array([[-1.0 / h_i[2], 0.0, h_i[0] / h_i[2] ** 2],
                               [0.0, -1.0 / h_i[2], h_i[1] / h_i[2] ** 2]])
                J[2 * k:2 * k + 2] = j1 @ j0
            delta = np.linalg.inv(J.T @ J) @ J.T @ r
            alpha -= delta[0]
            beta -= delta[1]
            rho -= delta[2]
            r_norm_last = r_norm
            r_norm = np.sqrt(r @ r)
        return np.array([alpha, beta, rho])
