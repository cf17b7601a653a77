"""Action sequences of HybridActionRRT / OracleRRT generated on the device (mps_extend_upload_generated) against
(a) hand-derived known answers of the closed forms (RampComputer::steer, HumanOracle::
predictAction / sampleFeasibleState), (b) the numpy restatement tests/hybrid_np.py chain by chain, (c) the oracle: the
rollouts of the generated chains are bit-identical to the CPU propagation of the same controls; and the distribution
of the device's random-control sampler against RampVelocityControlSampler::sample (RampVelocityControl.cpp:353-394)."""
import math

import numpy as np
import pytest

from manipulation_planning_suite_b200 import abi, api, scenes

import hybrid_np as hn
import oracle_binding as ob

pytestmark = pytest.mark.gpu


def _scene():
    sc, _ = scenes.make_scene(3, seed=8, walls=False)
    sc.x_min = sc.y_min = -100.0
    sc.x_max = sc.y_max = 100.0
    return sc


def test_closed_forms_known_answers():
    sc = _scene()
    ctx = api.Context(sc)
    far = [50.0, 50.0, 0.0, -50.0, 50.0, 0.0]
    # (1) transit: RampComputer::steer from (0,0,0) to (1,0,0), limits (0.6,0.6,1.5), accelerations (2,2,4), t_max 1:
    # v = 0.6, t_acc = 0.3, the ramps cover 0.18 -> plateau 0.6 m in 1 s; the rest 0.22 = 0.18 + 0.04 -> 0.04 / 0.6 s
    start = np.array([0, 0, 0, 10, 10, 0] + far, np.float32)
    dest = np.array([1, 0, 0, 10, 10, 0] + far, np.float32)
    g = api.Context.hybrid_gen(sc, seed=1, fixed_object=0)
    ctx.extend_upload_generated(start, 4, 6, g, dest=dest)
    c, n, tr = ctx.generated_controls()
    assert tr == 0 and (n == 2).all()
    assert np.allclose(c[0, 0], [0.6, 0, 0, 1.0, 0], atol=1e-6) and np.allclose(c[0, 1], [0.6, 0, 0, 0.04 / 0.6, 0], atol=2e-6)
    assert np.array_equal(c[0], c[3])   # no randomness in a transit
    # a short move, 0.08 m: the ramps of the fastest control (0.18 m) would overshoot -> plateau velocity d / t_acc =
    # 0.08 / 0.3 (RampComputer.cpp:73-80), whose own ramps cover v^2 / 2 = 0.0356 m; the rest is plateau
    dest2 = dest.copy()
    dest2[0] = 0.08
    ctx.extend_upload_generated(start, 1, 6, g, dest=dest2)
    c, n, _ = ctx.generated_controls()
    v = 0.08 / 0.3
    assert n[0] == 1 and c[0, 0, 0] == pytest.approx(v, rel=1e-6) and c[0, 0, 3] == pytest.approx((0.08 - v * v / 2) / v, rel=1e-5)
    # a pure rotation by -pi/2: the shortest direction, omega limit 1.5, t_acc = 0.375
    dest3 = start.copy()
    dest3[2] = -math.pi / 2
    ctx.extend_upload_generated(start, 1, 6, g, dest=dest3)
    c, n, _ = ctx.generated_controls()
    assert n[0] == 1 and c[0, 0, 2] == pytest.approx(-1.5) and c[0, 0, 0] == 0 and c[0, 0, 1] == 0
    assert c[0, 0, 3] == pytest.approx((math.pi / 2 - 1.5 * 0.375) / 1.5, rel=1e-5)
    # (2) transfer without noise: object 1 at (0.3, 0.1) to be pushed to (0.3, 0.6): the pushing pose is 0.2 m behind
    # the object on the line of the push, facing it (atan2 - 1.57); the push is the first control of the steer
    # towards the destination minus half an object size (0.06)
    start = np.array([0, 0, 0, 0.3, 0.1, 0] + far, np.float32)
    dest = np.array([0, 0, 0, 0.3, 0.6, 0] + far, np.float32)
    g = api.Context.hybrid_gen(sc, seed=1, fixed_object=1, push_distance_tolerance=0.0, push_angle_tolerance=0.0)
    ctx.extend_upload_generated(start, 2, 6, g, dest=dest)
    c, n, _ = ctx.generated_controls()
    feas = np.array([0.3, -0.1, math.pi / 2 - 1.57])
    want = hn.steer(start[:3], feas.astype(np.float32), (0.6, 0.6, 1.5), (2.0, 2.0, 4.0), 1.0)
    assert n[0] == len(want) + 1
    for i, w in enumerate(want):
        assert np.allclose(c[0, i], w, atol=2e-6)
    # the push: from (0.3, -0.1) towards (0.3, 0.54): 0.64 m straight up at 0.6 m/s -> plateau (0.64 - 0.18) / 0.6 s
    assert np.allclose(c[0, n[0] - 1], [0.0, 0.6, 0.0, (0.64 - 0.18) / 0.6, 0.0], atol=3e-6)
    # an object that only has to turn: predictAction gives the null control (HumanOracle.cpp:128-132)
    dest4 = start.copy()
    dest4[5] = 1.0
    ctx.extend_upload_generated(start, 1, 6, g, dest=dest4)
    c, n, _ = ctx.generated_controls()
    assert np.array_equal(c[0, n[0] - 1], np.zeros(5, np.float32))
    ctx.close()


def test_generated_chains_equal_the_numpy_restatement_and_replay_on_the_oracle():
    sc, _ = scenes.make_scene(6, seed=1002)
    start = scenes.cluttered_start(sc, 1002)
    rng = np.random.RandomState(4)
    dest = scenes.sample_state(rng, sc)
    K, L = 256, 6
    ctx = api.Context(sc)
    g = api.Context.hybrid_gen(sc, seed=77, iteration=3, k_offset=1000, action_randomness=0.15, target_bias=0.3, robot_bias=0.1,
                               targets=(1, 2))
    goals = [(1, 0.5, 0.5, 0.1)]
    ctx.extend_upload_generated(start, K, L, g, dest=dest, compare_f32=True, goals=goals, k_offset=1000)
    c, n, tr = ctx.generated_controls()
    kinds = {}
    acc = [sc.accel_limits[i] for i in range(3)]
    close = 0
    for k in range(K):
        wc, wn, wtr, kind = hn.action_sequence(g, k, start, dest, sc.n_bodies, sc.robot_id, acc, L)
        kinds[kind[:8]] = kinds.get(kind[:8], 0) + 1
        if wn == n[k] and np.allclose(c[k, :wn], wc, rtol=2e-5, atol=2e-6):
            close += 1
        assert (c[k, n[k]:] == 0).all()
    # libm differences of a few ulp can flip the steer loop's overshoot test on a rare chain: all but a handful agree
    assert close >= K - 3, f"{K - close} of {K} chains differ from the restatement"
    assert kinds.get("random", 0) > 15 and kinds.get("transit", 0) > 5 and kinds.get("transfer", 0) > 100
    # the extend on the generated chains == the oracle on the same controls, bit for bit
    ctx.extend_launch()
    got = ctx.extend_fetch(dump=True)
    want = ob.extend(sc, start, c, n_ctrl=n, dest=dest, compare_f32=True, goals=goals, k_offset=1000)
    assert np.array_equal(got.all_flags, want["all_flags"]) and np.array_equal(got.all_states, want["all_states"])
    assert (got.k, got.n_ok, got.distance) == (want["k"], want["n_ok"], want["distance"])
    assert got.k >= 1000
    # sharding: chains [128, 256) generated alone are the same chains (k_offset shifts the random streams)
    g2 = api.Context.hybrid_gen(sc, seed=77, iteration=3, k_offset=1128, action_randomness=0.15, target_bias=0.3, robot_bias=0.1,
                                targets=(1, 2))
    ctx.extend_upload_generated(start, 128, L, g2, dest=dest, compare_f32=True, goals=goals, k_offset=1128)
    c2, n2, _ = ctx.generated_controls()
    assert np.array_equal(c2, c[128:]) and np.array_equal(n2, n[128:])
    # chains longer than L are cut and counted
    ctx.extend_upload_generated(start, K, 1, g, dest=dest)
    c1, n1, tr1 = ctx.generated_controls()
    assert (n1 <= 1).all() and tr1 == int((n > 1).sum()) + 0 * tr
    ctx.close()


def test_device_control_sampler_has_the_reference_distribution():
    """RampVelocityControlSampler::sample: (vx, vy) uniform in the ball of radius v_limit (uniformInBall), omega and
    the plateau uniform, rest time 0 - Kolmogorov-Smirnov against the exact laws"""
    from scipy import stats

    sc = _scene()
    ctx = api.Context(sc)
    n = 200_000
    c = ctx.sample_controls(seed=12345, n=n)
    ctx.close()
    r = np.hypot(c[:, 0], c[:, 1]).astype(np.float64)
    assert r.max() <= 0.6 + 1e-6
    # uniform in the disc: (r / R)^2 ~ U(0, 1), the angle ~ U(-pi, pi), the two independent
    assert stats.kstest((r / 0.6) ** 2, "uniform").pvalue > 1e-3
    ang = np.arctan2(c[:, 1], c[:, 0]).astype(np.float64)
    assert stats.kstest((ang + np.pi) / (2 * np.pi), "uniform").pvalue > 1e-3
    assert abs(np.corrcoef(r, ang)[0, 1]) < 0.01
    assert stats.kstest((c[:, 2].astype(np.float64) + 1.5) / 3.0, "uniform").pvalue > 1e-3
    assert stats.kstest((c[:, 3].astype(np.float64) - 0.1) / 0.9, "uniform").pvalue > 1e-3
    assert (c[:, 4] == 0).all()
    # moments of a uniform disc: E[x] = 0, E[x^2] = R^2 / 4
    assert abs(c[:, 0].mean()) < 3e-3 and abs((c[:, 0].astype(np.float64) ** 2).mean() - 0.09) < 1e-3
    # and the numpy restatement of the same hash gives the same draws
    for i in (0, 1, 77, 4095):
        w = hn.sample_control(12345, 0, 0, hn.TAG_CTRL, i, 0.6, 1.5, 0.1, 1.0)
        assert np.allclose(c[i], w, atol=2e-6)
