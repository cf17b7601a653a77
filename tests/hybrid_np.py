"""numpy-float32 restatement of the closed forms behind HybridActionRRT::sampleActionSequence, written from the
reference's sources (RampComputer.cpp:40-92, HumanOracle.cpp:111-159, OracleControlSampler.cpp:77-207, RRT.cpp:491-529,
RampVelocityControl.cpp:98-118,353-394) and the random-number hash of csrc/forest_step.cu.  TEST INFRASTRUCTURE: the
checker of the device-side generator (mps_extend_upload_generated)."""
import math

import numpy as np

f32 = np.float32
M64 = (1 << 64) - 1
TAG_CTRL, TAG_H_DIE, TAG_H_OBJ, TAG_H_OBJ_UNIFORM, TAG_H_TGT, TAG_H_GAUSS, TAG_H_ZERO, TAG_H_RANDOM = 7, 11, 12, 13, 14, 15, 16, 17


def mix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def u01(seed, iteration, gid, tag, idx):
    h = mix64(seed ^ mix64((iteration * 0x100000001B3 + tag) & M64))
    h = mix64(h ^ mix64(gid & M64))
    h = mix64(h ^ (idx & M64))
    return f32(h >> 40) * f32(1.0 / 16777216.0)


def sample_control(seed, iteration, gid, tag, idx, v_limit, w_limit, dmin, dmax):
    two_pi = f32(6.28318530717958647692)
    r = f32(v_limit) * np.sqrt(u01(seed, iteration, gid, tag, idx * 4))
    a = two_pi * u01(seed, iteration, gid, tag, idx * 4 + 1)
    return np.array([r * np.cos(a), r * np.sin(a), (f32(2) * u01(seed, iteration, gid, tag, idx * 4 + 2) - f32(1)) * f32(w_limit),
                     f32(dmin) + (f32(dmax) - f32(dmin)) * u01(seed, iteration, gid, tag, idx * 4 + 3), 0], np.float32)


def shortest_so2(a, b):
    v = f32(b) - f32(a)
    if abs(v) > f32(math.pi):
        v = v - f32(2 * math.pi) if v > 0 else v + f32(2 * math.pi)
    return f32(v)


def accel_time(v, acc):
    return max(abs(f32(v[i]) / f32(acc[i])) for i in range(3))


def steer(cur, des, vlim, acc, dur_max):
    """RampComputer::steer -> list of (vx, vy, w, duration, 0)"""
    d = np.array([f32(des[0]) - f32(cur[0]), f32(des[1]) - f32(cur[1]), shortest_so2(cur[2], des[2])], np.float32)
    distance = np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
    out = []
    if not distance > 0:
        return out
    d = d / distance
    vabs = min(f32(vlim[i]) / abs(d[i]) if d[i] != 0 else f32(np.inf) for i in range(3))
    vel = vabs * d
    ta = accel_time(vel, acc)
    norm3 = lambda v, s: np.sqrt(s * v[0] * s * v[0] + s * v[1] * s * v[1] + s * v[2] * s * v[2])  # noqa: E731
    dist_accel = norm3(vel, ta)
    while distance > 0 and len(out) < 64:
        if dist_accel > distance:
            vabs = distance / ta
            vel = vabs * d
            ta = accel_time(vel, acc)
            dist_accel = norm3(vel, ta)
        distance = distance - dist_accel
        plateau = min(norm3(vel, f32(dur_max)), distance)
        out.append(np.array([vel[0], vel[1], vel[2], plateau / vabs, 0], np.float32))
        distance = distance - plateau
    return out


def gauss01(u1, u2):
    return np.sqrt(f32(-2) * np.log(f32(1) - u1)) * np.cos(f32(6.28318530717958647692) * u2)


def normalize_d(v):
    v = math.fmod(v, 2.0 * math.pi)
    if v < -math.pi:
        v += 2.0 * math.pi
    elif v >= math.pi:
        v -= 2.0 * math.pi
    return v


def feasible_state(cur_o, des_o, pushing_dist, angle, zero_draws=(0.5, 0.5)):
    """HumanOracle::sampleFeasibleState with the two gaussian draws given"""
    rel = [f32(des_o[0]) - f32(cur_o[0]), f32(des_o[1]) - f32(cur_o[1]), shortest_so2(cur_o[2], des_o[2])]
    if np.sqrt(rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2]) == 0:
        rel[0] = f32(0.0001) * (f32(0.5) - f32(zero_draws[0]))
        rel[1] = f32(0.0001) * (f32(0.5) - f32(zero_draws[1]))
    n = np.sqrt(rel[0] * rel[0] + rel[1] * rel[1])
    pd = (f32(1), f32(0)) if not n > 0 else (rel[0] / n, rel[1] / n)
    ca, sa = np.cos(f32(angle)), np.sin(f32(angle))
    x = f32(cur_o[0]) - f32(pushing_dist) * (ca * pd[0] - sa * pd[1])
    y = f32(cur_o[1]) - f32(pushing_dist) * (sa * pd[0] + ca * pd[1])
    th = f32(np.arctan2(pd[1], pd[0])) - f32(1.57)
    return np.array([x, y, f32(normalize_d(float(th)))], np.float32)


def predict_action(cur_r, cur_o, des_o, size, vlim, acc, dur_max):
    """HumanOracle::predictAction -> one control"""
    rx, ry = f32(des_o[0]) - f32(cur_o[0]), f32(des_o[1]) - f32(cur_o[1])
    n = np.sqrt(rx * rx + ry * ry)
    if n == 0:
        return np.zeros(5, np.float32)
    half = f32(size) / f32(2)
    nxt = [f32(des_o[0]) - rx / n * half, f32(des_o[1]) - ry / n * half, cur_r[2]]
    c = steer(cur_r, nxt, vlim, acc, dur_max)
    return c[0] if c else np.zeros(5, np.float32)


def action_sequence(g, k, start, dest, B, robot, acc, L):
    """chain k of mps_extend_upload_generated: (controls [n][5] (n <= L), n, truncated, kind)"""
    gid = g.k_offset + k
    vlim = [g.velocity_limits[i] for i in range(3)]
    cur_r = start[3 * robot:3 * robot + 3]
    die = u01(g.seed, g.iteration, gid, TAG_H_DIE, 0)
    if die < f32(g.action_randomness):
        seq = [sample_control(g.seed, g.iteration, gid, TAG_H_RANDOM, 0, vlim[0], vlim[2], g.duration_min, g.duration_max)]
        kind = "random"
    else:
        obj = g.fixed_object
        if obj < 0:
            uo = u01(g.seed, g.iteration, gid, TAG_H_OBJ, 0)
            obj = min(int(u01(g.seed, g.iteration, gid, TAG_H_OBJ_UNIFORM, 0) * f32(B)), B - 1)
            if uo < f32(g.target_bias) and g.n_targets > 0:
                obj = g.targets[min(int(u01(g.seed, g.iteration, gid, TAG_H_TGT, 0) * f32(g.n_targets)), g.n_targets - 1)]
            elif uo < f32(g.target_bias) + f32(g.robot_bias):
                obj = robot
        if obj == robot:
            seq = steer(cur_r, dest[3 * robot:3 * robot + 3], vlim, acc, g.duration_max)
            kind = "transit"
        else:
            cur_o, des_o = start[3 * obj:3 * obj + 3], dest[3 * obj:3 * obj + 3]
            pd = f32(g.optimal_push_distance) + f32(g.push_distance_tolerance) * gauss01(
                u01(g.seed, g.iteration, gid, TAG_H_GAUSS, 0), u01(g.seed, g.iteration, gid, TAG_H_GAUSS, 1))
            an = f32(g.push_angle_tolerance) * gauss01(u01(g.seed, g.iteration, gid, TAG_H_GAUSS, 2),
                                                       u01(g.seed, g.iteration, gid, TAG_H_GAUSS, 3))
            feas = feasible_state(cur_o, des_o, pd, an, (u01(g.seed, g.iteration, gid, TAG_H_ZERO, 0),
                                                         u01(g.seed, g.iteration, gid, TAG_H_ZERO, 1)))
            seq = steer(cur_r, feas, vlim, acc, g.duration_max)
            seq.append(predict_action(feas, cur_o, des_o, g.object_size[obj], vlim, acc, g.duration_max))
            kind = f"transfer{obj}"
    n = len(seq)
    return np.array(seq[:L], np.float32).reshape(-1, 5), min(n, L), n > L, kind
