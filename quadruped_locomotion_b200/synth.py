"""Synthetic simpledog states for tests and bench.

Counter-based generator: every value is a pure function of (seed, stream, instance index), so any
slice of a batch can be regenerated on any rank without communication and the host oracle and the
GPU see bit-identical inputs.  All arrays are SoA, component-major / batch-minor, like the C ABI.

Configs (BASELINE.json `configs`):
  C1  one state, four-leg stance, zero perturbation
  C2  trot: 65 536 states, stance mask alternating {RF,LH} / {LF,RH}
  C3  2^20 random states, stance mix 60 % four / 25 % diagonal pair / 15 % three legs
  C4  = C3 inputs (FP32 variant, sharded)
  C5  Monte-Carlo: 2^14 nominal states x 2^10 perturbations of base attitude, mu and wrench scale
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0x5EED0000
LF, RF, RH, LH = 0, 1, 2, 3
KNEE_SIGN = np.array([+1.0, -1.0, +1.0, -1.0])  # LF, RF, RH, LH (generator choice)
TOTAL_MASS = 51.0  # 27 kg torso + 4 x 6 kg legs (quadruped_state.cpp:28,36-41)
GRAVITY = 9.8

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _bits(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        key = _mix64(np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03))
        return _mix64(key ^ _mix64(idx.astype(np.uint64)))


def uniform(seed: int, stream: int, idx: np.ndarray, lo=0.0, hi=1.0) -> np.ndarray:
    u = (_bits(seed, stream, idx) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return lo + (hi - lo) * u


# ---- elementary functions from IEEE-exact operations only (+, -, *, /, sqrt, rint, frexp; one rounding per
# operation, no fused multiply-add): the device generator (csrc/qlb_gen.cuh) repeats the same operation sequence with
# the __d*_rn intrinsics, so host and device produce BIT-IDENTICAL states (tests/test_generator.py).  Accuracy ~1e-16.
_PIO2_HI = 1.57079632673412561417e+00
_PIO2_LO = 6.07710050650619224932e-11
_TWO_OVER_PI = 6.36619772367581382433e-01
_SIN_C = (1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
          -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01)
_COS_C = (-1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
          2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02)
_LN2_HI = 6.93147180369123816490e-01
_LN2_LO = 1.90821492927058770002e-10
_SQRT_HALF = 0.70710678118654752440


def sincos_exact(x):
    """(sin x, cos x) for |x| up to ~1e4: Cody-Waite reduction by pi/2 and the fdlibm kernel polynomials, every
    multiply and add rounded separately."""
    x = np.asarray(x, dtype=np.float64)
    kd = np.rint(x * _TWO_OVER_PI)
    k = kd.astype(np.int64)
    r = (x - kd * _PIO2_HI) - kd * _PIO2_LO
    z = r * r
    ps = np.full_like(z, _SIN_C[0])
    for c in _SIN_C[1:]:
        ps = ps * z + c
    sr = r + (z * r) * ps
    pc = np.full_like(z, _COS_C[0])
    for c in _COS_C[1:]:
        pc = pc * z + c
    cr = (1.0 - 0.5 * z) + (z * z) * pc
    odd = (k & 1) != 0
    s0 = np.where(odd, cr, sr)
    c0 = np.where(odd, sr, cr)
    sn = np.where((k & 2) != 0, -s0, s0)
    cs = np.where(((k + 1) & 2) != 0, -c0, c0)
    return sn, cs


def log_exact(x):
    """log x for 0 < x <= 1 (what Box-Muller needs): frexp, then 2 atanh((m-1)/(m+1)) as an odd series in
    s = (m-1)/(m+1), |s| <= 0.172, through s^21."""
    x = np.asarray(x, dtype=np.float64)
    m, e = np.frexp(x)
    small = m < _SQRT_HALF
    m = np.where(small, m * 2.0, m)
    ed = (e - small.astype(e.dtype)).astype(np.float64)
    s = (m - 1.0) / (m + 1.0)
    z = s * s
    p = np.full_like(z, 1.0 / 21.0)
    for n in (19, 17, 15, 13, 11, 9, 7, 5, 3, 1):
        p = p * z + 1.0 / n
    return ed * _LN2_HI + (ed * _LN2_LO + (2.0 * s) * p)


def normal(seed: int, stream: int, idx: np.ndarray, sigma=1.0) -> np.ndarray:
    u1 = uniform(seed, 2 * stream + 1000, idx)
    u2 = uniform(seed, 2 * stream + 1001, idx)
    _, c = sincos_exact((2.0 * np.pi) * u2)
    return (sigma * np.sqrt(-2.0 * log_exact(1.0 - u1))) * c


def quat_from_ypr(yaw, pitch, roll):
    """(w,x,y,z) of R = Rz(yaw) Ry(pitch) Rx(roll), base->world."""
    sy, cy = sincos_exact(0.5 * np.asarray(yaw, dtype=np.float64))
    sp, cp = sincos_exact(0.5 * np.asarray(pitch, dtype=np.float64))
    sr, cr = sincos_exact(0.5 * np.asarray(roll, dtype=np.float64))
    w = cr * cp * cy + sr * sp * sy
    x = sr * cp * cy - cr * sp * sy
    y = cr * sp * cy + sr * cp * sy
    z = cr * cp * sy - sr * sp * cy
    return np.stack([w, x, y, z])


def rot_from_quat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def _joints(seed, idx):
    q = np.empty((12, idx.size))
    for leg in range(4):
        s = KNEE_SIGN[leg]
        q[3 * leg + 0] = uniform(seed, 10 + 3 * leg, idx, -0.25, 0.25)
        q[3 * leg + 1] = s * (0.7 + uniform(seed, 11 + 3 * leg, idx, -0.3, 0.3))
        q[3 * leg + 2] = -s * (1.4 + uniform(seed, 12 + 3 * leg, idx, -0.4, 0.4))
    return q


def _wrench(seed, idx, quat, scale=None):
    R = rot_from_quat(quat)  # base->world, shape (3,3,B)
    w = np.empty((6, idx.size))
    weight = TOTAL_MASS * GRAVITY
    if scale is not None:
        weight = weight * scale
    for a in range(3):
        # R_wb (0,0,mg) = third row of R_bw
        w[a] = R[2, a] * weight + normal(seed, 30 + a, idx, 60.0)
        w[3 + a] = normal(seed, 33 + a, idx, 25.0)
    return w


def make_states(config: str, B: int | None = None, start: int = 0, seed: int | None = None) -> dict:
    """Return dict(q[12,B], quat[4,B], wrench[6,B], mask[B] uint8, mu[4,B], normals[12,B]).

    `start` offsets the instance index (rank sharding: rank r of N takes start = r*B_local).
    """
    config = config.upper()
    cid = {"C1": 1, "C2": 2, "C3": 3, "C4": 3, "C5": 5}[config]
    if seed is None:
        seed = SEED_BASE + cid
    if B is None:
        B = {"C1": 1, "C2": 65536, "C3": 1 << 20, "C4": 1 << 20, "C5": 1 << 24}[config]
    idx = np.arange(start, start + B, dtype=np.uint64)
    mu = np.full((4, B), 0.6)
    normals = np.zeros((12, B))
    normals[2::3] = 1.0

    if config == "C1":
        q = np.tile(np.array([0.0, 0.7, -1.4, 0.0, -0.7, 1.4, 0.0, 0.7, -1.4, 0.0, -0.7, 1.4])[:, None], (1, B))
        quat = np.zeros((4, B)); quat[0] = 1.0
        wrench = np.zeros((6, B)); wrench[2] = TOTAL_MASS * GRAVITY
        mask = np.full(B, 0xF, dtype=np.uint8)
        return dict(q=q, quat=quat, wrench=wrench, mask=mask, mu=mu, normals=normals)

    if config == "C5":
        nominal = idx >> np.uint64(10)   # 2^10 perturbations per nominal state
        q = _joints(seed, nominal)
        yaw = uniform(seed, 1, nominal, -np.pi, np.pi) + normal(seed, 40, idx, 0.3)
        pitch = uniform(seed, 2, nominal, -0.25, 0.25) + normal(seed, 41, idx, 0.1)
        roll = uniform(seed, 3, nominal, -0.25, 0.25) + normal(seed, 42, idx, 0.1)
        quat = quat_from_ypr(yaw, pitch, roll)
        for leg in range(4):
            mu[leg] = uniform(seed, 50 + leg, idx, 0.2, 1.0)
        wrench = _wrench(seed, nominal, quat, scale=uniform(seed, 60, idx, 0.8, 1.2))
        sel = uniform(seed, 4, nominal)
    else:
        q = _joints(seed, idx)
        yaw = uniform(seed, 1, idx, -np.pi, np.pi)
        pitch = uniform(seed, 2, idx, -0.25, 0.25)
        roll = uniform(seed, 3, idx, -0.25, 0.25)
        quat = quat_from_ypr(yaw, pitch, roll)
        wrench = _wrench(seed, idx, quat)
        sel = uniform(seed, 4, idx)

    if config == "C2":
        mask = np.where(idx % np.uint64(2) == 0, (1 << RF) | (1 << LH), (1 << LF) | (1 << RH)).astype(np.uint8)
    else:
        # 60 % four-stance, 25 % diagonal pairs, 15 % three-stance
        pick = (_bits(seed, 5, idx if config != "C5" else (idx >> np.uint64(10))) >> np.uint64(40)).astype(np.int64)
        diag = np.where(pick % 2 == 0, (1 << RF) | (1 << LH), (1 << LF) | (1 << RH))
        three = 0xF & ~(1 << (pick % 4))
        mask = np.where(sel < 0.60, 0xF, np.where(sel < 0.85, diag, three)).astype(np.uint8)
    return dict(q=np.ascontiguousarray(q), quat=np.ascontiguousarray(quat), wrench=wrench, mask=mask,
                mu=mu, normals=normals)
