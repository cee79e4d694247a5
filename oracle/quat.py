"""Quaternion / small-matrix helpers with Eigen semantics (oracle; test infrastructure only).

Quaternions are stored (x, y, z, w) -- the order of Eigen's `coeffs()` and of the reference's
`q_array_` (reference: src/x/ekf/state.cpp:235-247).  Hamilton product, body->global rotation.
"""
import numpy as np


def qnormalized(q):
    q = np.asarray(q, dtype=np.float64)
    return q / np.sqrt(q @ q)


def rot_raw(q):
    """Eigen::Quaterniond::toRotationMatrix() WITHOUT normalisation (used by propagator.cpp:46-47)."""
    x, y, z, w = q
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1.0 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1.0 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1.0 - (txx + tyy)]])


def rot(q):
    """`q.normalized().toRotationMatrix()` -- the form used everywhere in src/x/vio/*.cpp."""
    return rot_raw(qnormalized(q))


def qmul(a, b):
    """Eigen `a * b` (Hamilton), (x,y,z,w) storage."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def qconj(q):
    return np.array([-q[0], -q[1], -q[2], q[3]])


def small_angle_quat(dtheta):
    """reference: src/x/ekf/state.cpp:273-283 (errorQuatFromSmallAngles): exact angle-axis."""
    dtheta = np.asarray(dtheta, dtype=np.float64)
    n = np.sqrt(dtheta @ dtheta)
    if n == 0.0:
        return np.array([0.0, 0.0, 0.0, 1.0])
    axis = dtheta / n
    s = np.sin(0.5 * n)
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, np.cos(0.5 * n)])


def skew(v):
    """reference: include/x/common/eigen_matrix_base_plugin.h:32-41 (toCrossMatrix), tools.h:57-66 (Skew)."""
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def omega(v):
    """reference: include/x/common/eigen_matrix_base_plugin.h:43-52 (toOmegaMatrix)."""
    x, y, z = v
    return np.array([[0.0, z, -y, x], [-z, 0.0, x, y], [y, -x, 0.0, z], [-x, -y, -z, 0.0]])
