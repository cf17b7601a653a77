"""Synthetic planar push scenes (the reference ships none) and seeded inputs.

SURVEY.md §8(d): workspace x, y in [-1, 1] m bounded by four static wall boxes; robot = SE(2)
floating end-effector box 0.10 x 0.05 m; objects = boxes 0.12 x 0.12 m (every 4th a disc r = 0.06 m),
m = 0.35 kg, I = 0.002 kg m^2, ground mu = 0.19, body-body mu = 0.3 (the magnitudes the reference
mentions at oracle/LearnedOracle.cpp:314-318).  Control limits: |v| <= 0.6 m/s in a ball, |omega| <=
1.5 rad/s, accelerations 2 / 2 / 4, plateau U[0.1, 1.0] s (pushing/OraclePushPlanner.cpp:70-71).

Collision policy defaults follow PlanningProblem (pushing/OraclePushPlanner.cpp:90-91): static
collisions allowed, robot-static forbidden.

Everything here is host-side input generation; no arithmetic of the hot path.
"""
from __future__ import annotations

import math

import numpy as np

from . import abi

# BASELINE.json configs -> (movable objects, K, L)
CONFIGS = {
    1: dict(n_objects=3, K=10, L=1, name="NaiveRRT robot+3 objects, K=10"),
    2: dict(n_objects=6, K=1024, L=4, name="HybridActionRRT 6 objects, K=1024 chains (L<=4)"),
    3: dict(n_objects=10, K=1, L=3, name="SliceOracleRRT 10 objects, 1M-node tree"),
    4: dict(n_objects=16, K=65536, L=1, name="OracleRRT (batched) 16 objects, K=65536"),
    5: dict(n_objects=8, K=10, L=1, name="4096 x NaiveRRT 8 objects"),
}

V_LIMIT = 0.6
W_LIMIT = 1.5
DURATION_LIMITS = (0.1, 1.0)


def default_scene() -> abi.Scene:
    """The values of mps_scene_defaults (csrc/abi.cu; DESIGN.md §3 / SURVEY.md Appendix A) without loading the CUDA
    library, so that the CPU-only arms of bench.py never map the product .so;
    tests/test_abi_and_host.py checks the two byte for byte."""
    sc = abi.Scene()
    sc.vel_iters = 8
    for i in range(abi.MPS_MAX_BODIES):
        sc.shape[i] = abi.MPS_SHAPE_BOX
        sc.hx[i] = 0.06
        sc.hy[i] = 0.06
        sc.mass[i] = 0.35
        sc.inertia[i] = 0.002
        sc.mu_ground[i] = 0.19
        sc.fr_radius[i] = 0.0459
        sc.mu_contact[i] = 0.3
    for i in range(abi.MPS_MAX_STATICS):
        sc.s_shape[i] = abi.MPS_SHAPE_BOX
        sc.s_hx[i] = 0.05
        sc.s_hy[i] = 0.05
        sc.s_mu_contact[i] = 0.3
    flt_max = 3.4028234663852886e38
    sc.x_min, sc.x_max, sc.y_min, sc.y_max = -flt_max, flt_max, -flt_max, flt_max
    sc.accel_limits[0], sc.accel_limits[1], sc.accel_limits[2] = 2.0, 2.0, 4.0
    sc.position_weight = 1.0
    sc.orientation_weight = 0.08
    sc.dt = 0.01
    sc.t_max = 8.0
    sc.gravity = 9.81
    sc.slop = 0.002
    sc.baumgarte = 0.2
    sc.max_bias_vel = 0.5
    sc.pen_max = 0.01
    sc.v_rest = 0.001
    sc.w_rest = 0.01
    sc.max_manifolds = abi.MPS_MAX_MANIFOLDS
    sc.strict_active_contacts = 0
    sc.jam_steps = 25
    return sc


def make_scene(n_objects: int, seed: int = 0, workspace: float = 1.0, walls: bool = True,
               disc_every: int = 4) -> tuple[abi.Scene, np.ndarray]:
    """Returns (scene, start_state[3B]) with the robot as body 0 and non-overlapping seeded poses."""
    sc = default_scene()
    B = n_objects + 1
    if B > abi.MPS_MAX_BODIES:
        raise ValueError("too many bodies")
    sc.n_bodies = B
    sc.robot_id = 0
    # robot: 0.10 x 0.05 m box
    sc.shape[0] = abi.MPS_SHAPE_BOX
    sc.hx[0] = 0.05
    sc.hy[0] = 0.025
    sc.mass[0] = 1.0
    sc.inertia[0] = 1.0
    sc.mu_ground[0] = 0.0
    sc.fr_radius[0] = 0.0
    sc.mu_contact[0] = 0.3
    for i in range(1, B):
        disc = disc_every > 0 and (i % disc_every == 0)
        sc.shape[i] = abi.MPS_SHAPE_DISC if disc else abi.MPS_SHAPE_BOX
        sc.hx[i] = 0.06
        sc.hy[i] = 0.06
        sc.mass[i] = 0.35
        sc.inertia[i] = 0.002
        sc.mu_ground[i] = 0.19
        # lever arm of the support friction torque: 2/3 r for a disc, 0.3826 a for a square of side a
        sc.fr_radius[i] = (2.0 / 3.0) * 0.06 if disc else 0.3826 * 0.12
        sc.mu_contact[i] = 0.3
    w = float(workspace)
    sc.x_min, sc.x_max, sc.y_min, sc.y_max = -w, w, -w, w
    if walls:
        t = 0.05  # wall half thickness
        specs = [(-w - t, 0.0, t, w + 2 * t), (w + t, 0.0, t, w + 2 * t),
                 (0.0, -w - t, w + 2 * t, t), (0.0, w + t, w + 2 * t, t)]
        sc.n_statics = 4
        for k, (x, y, hx, hy) in enumerate(specs):
            sc.s_shape[k] = abi.MPS_SHAPE_BOX
            sc.s_x[k], sc.s_y[k], sc.s_theta[k] = x, y, 0.0
            sc.s_hx[k], sc.s_hy[k] = hx, hy
            sc.s_mu_contact[k] = 0.3
    else:
        sc.n_statics = 0
    # policy: robot must not touch statics
    sc.static_forbidden = 1 << 0
    sc.max_manifolds = int(min(abi.MPS_MAX_MANIFOLDS, max(16, 3 * B)))
    start = place_bodies(sc, seed)
    return sc, start


def bounding_radius(sc: abi.Scene, i: int) -> float:
    if sc.shape[i] == abi.MPS_SHAPE_DISC:
        return float(sc.hx[i])
    return math.hypot(sc.hx[i], sc.hy[i])


def place_bodies(sc: abi.Scene, seed: int, margin: float = 0.01) -> np.ndarray:
    """Rejection-sampled poses whose bounding circles do not overlap each other or the walls."""
    rng = np.random.RandomState(seed)
    B = sc.n_bodies
    state = np.zeros(3 * B, dtype=np.float32)
    placed: list[tuple[float, float, float]] = []
    for i in range(B):
        r = bounding_radius(sc, i)
        for _ in range(100000):
            x = rng.uniform(sc.x_min + r + margin, sc.x_max - r - margin)
            y = rng.uniform(sc.y_min + r + margin, sc.y_max - r - margin)
            if all((x - px) ** 2 + (y - py) ** 2 > (r + pr + margin) ** 2 for px, py, pr in placed):
                break
        else:
            raise RuntimeError("could not place bodies")
        placed.append((x, y, r))
        state[3 * i] = x
        state[3 * i + 1] = y
        state[3 * i + 2] = rng.uniform(-math.pi, math.pi)
    return state


def sample_controls(rng: np.random.RandomState, K: int, L: int = 1) -> np.ndarray:
    """RampVelocityControlSampler::sample (ompl/control/RampVelocityControl.cpp:353-394): (vx, vy)
    uniform in a ball of radius V_LIMIT, omega and the plateau uniform, rest time 0.  [K][L][5] float32."""
    n = K * L
    # uniformInBall: direction from normalised gaussians, radius r * u^(1/dim)
    g = rng.normal(size=(n, 2))
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    rad = V_LIMIT * np.sqrt(rng.uniform(size=(n, 1)))
    v = g * rad
    om = rng.uniform(-W_LIMIT, W_LIMIT, size=(n, 1))
    dur = rng.uniform(DURATION_LIMITS[0], DURATION_LIMITS[1], size=(n, 1))
    c = np.concatenate([v, om, dur, np.zeros((n, 1))], axis=1).astype(np.float32)
    return c.reshape(K, L, 5)


def sample_state(rng: np.random.RandomState, sc: abi.Scene) -> np.ndarray:
    """Uniform state inside the workspace bounds (not validity-checked)."""
    B = sc.n_bodies
    s = np.zeros(3 * B, dtype=np.float32)
    s[0::3] = rng.uniform(sc.x_min, sc.x_max, size=B)
    s[1::3] = rng.uniform(sc.y_min, sc.y_max, size=B)
    s[2::3] = rng.uniform(-math.pi, math.pi, size=B)
    return s


def config_scene(cfg: int) -> tuple[abi.Scene, np.ndarray]:
    """Scene of BASELINE.json config `cfg` with seed 1000 + cfg."""
    c = CONFIGS[cfg]
    return make_scene(c["n_objects"], seed=1000 + cfg)


def cluttered_start(sc: abi.Scene, seed: int, radius: float | None = None, margin: float = 0.004) -> np.ndarray:
    """Contact-rich start: all bodies rejection-sampled inside a disc just large enough to hold them,
    so pushes hit objects within a few steps (used by the parity tests and the rollout benchmark)."""
    rng = np.random.RandomState(seed)
    B = sc.n_bodies
    rs = [bounding_radius(sc, i) for i in range(B)]
    if radius is None:
        radius = 1.9 * math.sqrt(sum(r * r for r in rs))
    radius = min(radius, min(sc.x_max, sc.y_max) - max(rs) - 0.02)
    state = np.zeros(3 * B, dtype=np.float32)
    for attempt in range(200):
        placed: list[tuple[float, float, float]] = []
        ok = True
        for i in range(B):
            r = rs[i]
            for _ in range(4000):
                a = rng.uniform(0, 2 * math.pi)
                d = radius * math.sqrt(rng.uniform())
                x, y = d * math.cos(a), d * math.sin(a)
                if all((x - px) ** 2 + (y - py) ** 2 > (r + pr + margin) ** 2 for px, py, pr in placed):
                    break
            else:
                ok = False
                break
            placed.append((x, y, r))
            state[3 * i], state[3 * i + 1] = x, y
            state[3 * i + 2] = rng.uniform(-math.pi, math.pi)
        if ok:
            return state
        radius *= 1.05
    raise RuntimeError("could not build a cluttered start")


def sample_controls_towards(rng: np.random.RandomState, sc: abi.Scene, start: np.ndarray, K: int, L: int = 1,
                            noise: float = 0.5) -> np.ndarray:
    """Controls whose translation points from the robot towards a random object (+- noise rad):
    a stand-in for the oracle-generated pushes of HybridActionRRT / OracleRRT (RRT.cpp:491-529)."""
    B = sc.n_bodies
    r = sc.robot_id
    c = sample_controls(rng, K, L)
    objs = [i for i in range(B) if i != r]
    for k in range(K):
        for l in range(L):
            j = objs[rng.randint(len(objs))]
            ang = math.atan2(start[3 * j + 1] - start[3 * r + 1], start[3 * j] - start[3 * r]) + rng.uniform(-noise, noise)
            sp = rng.uniform(0.2, V_LIMIT)
            c[k, l, 0] = sp * math.cos(ang)
            c[k, l, 1] = sp * math.sin(ang)
    return c
