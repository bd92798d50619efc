"""Host-side float32 quaternion algebra for the world-frame overload of DirectMethod::TrackFeatures
(src/direct_method_tracker/direct_method_tracker.cpp:8-39).  Quaternions are (w, x, y, z).  Every operation is a single IEEE
binary32 operation in the order Eigen's scalar code paths use (the same semantics oracle/shim/basic_type.h freezes for the
CPU checker), so the composed result is bit-identical to the reference's."""
import numpy as np

f32 = np.float32


def _q(q):
    return np.asarray(q, dtype=np.float32).reshape(4)


def inverse(q):
    q = _q(q)
    n2 = f32(f32(f32(q[1] * q[1]) + f32(q[2] * q[2])) + f32(q[3] * q[3])) + f32(q[0] * q[0])
    if n2 > 0:
        return np.array([q[0] / n2, -q[1] / n2, -q[2] / n2, -q[3] / n2], np.float32)
    return np.zeros(4, np.float32)


def multiply(a, b):
    a, b = _q(a), _q(b)
    w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3]
    x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2]
    y = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3]
    z = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]
    return np.array([w, x, y, z], np.float32)


def rotate(q, v):
    """q * v for v [..., 3] (Eigen's _transformVector: uv = vec x v; uv += uv; v + w * uv + vec x uv)."""
    q = _q(q)
    v = np.asarray(v, dtype=np.float32)
    qw, qx, qy, qz = q[0], q[1], q[2], q[3]
    vx, vy, vz = v[..., 0], v[..., 1], v[..., 2]
    ux = qy * vz - qz * vy
    uy = qz * vx - qx * vz
    uz = qx * vy - qy * vx
    ux, uy, uz = ux + ux, uy + uy, uz + uz
    out = np.empty_like(v)
    out[..., 0] = vx + qw * ux + (qy * uz - qz * uy)
    out[..., 1] = vy + qw * uy + (qz * ux - qx * uz)
    out[..., 2] = vz + qw * uz + (qx * uy - qy * ux)
    return out
