"""GPU parity of the rollout path against the CPU oracle, through the C ABI.

Bar: states within 1e-4 m / 1e-3 rad per body, flags / rest times / winner identical,
distances identical as doubles.  The physics spec fixes the operation order, so the expected
difference is zero; the test asserts the stated tolerance and reports exact-match counts.
"""
import numpy as np
import pytest

from manipulation_planning_suite_b200 import abi, api, scenes

import oracle_binding as ob

pytestmark = pytest.mark.gpu

POS_TOL = 1e-4
ANG_TOL = 1e-3


def _ang_diff(a, b):
    d = np.abs(a - b)
    return np.minimum(d, 2 * np.pi - d)


def assert_states_close(got, want, flags, what=""):
    """got / want: [..., 3B]; rows whose flags carry OVERFLOW are skipped (their state is unspecified)."""
    g = got.reshape(-1, got.shape[-1] // 3, 3)
    w = want.reshape(-1, want.shape[-1] // 3, 3)
    f = flags.reshape(-1)
    keep = ((f & abi.MPS_F_RAN) != 0) & ((f & abi.MPS_F_OVERFLOW) == 0)
    dpos = np.abs(g[keep][:, :, :2] - w[keep][:, :, :2]).max() if keep.any() else 0.0
    dang = _ang_diff(g[keep][:, :, 2], w[keep][:, :, 2]).max() if keep.any() else 0.0
    assert dpos <= POS_TOL, f"{what}: position differs by {dpos}"
    assert dang <= ANG_TOL, f"{what}: angle differs by {dang}"
    return dpos, dang, int((g[keep] == w[keep]).all(axis=(1, 2)).sum()), int(keep.sum())


def compare_extend(sc, start, controls, **kw):
    ctx = api.Context(sc)
    got = ctx.extend(start, controls, dump=True, **kw)
    want = ob.extend(sc, start, controls, **kw)
    assert np.array_equal(got.all_flags, want["all_flags"]), "flags differ"
    assert np.array_equal(got.all_steps, want["all_steps"]), "step counts differ"
    assert np.array_equal(got.all_rest, want["all_rest"]), "rest times differ"
    dpos, dang, exact, n = assert_states_close(got.all_states, want["all_states"], want["all_flags"], "all_states")
    assert np.array_equal(got.all_dist, want["all_dist"]), "distances differ"
    assert got.k == want["k"], f"winner differs: {got.k} vs {want['k']}"
    assert got.n_ok == want["n_ok"]
    assert got.goal_step == want["goal_step"]
    if got.k >= 0:
        assert got.distance == want["distance"]
        assert np.array_equal(got.best_flags, want["best_flags"])
        assert np.array_equal(got.best_controls, want["best_controls"])
        assert_states_close(got.best_states, want["best_states"], want["best_flags"], "best_states")
    ctx.close()
    return dict(dpos=dpos, dang=dang, exact=exact, n=n, got=got, want=want)


@pytest.mark.parametrize("n_objects,K,seed", [(3, 10, 0), (3, 64, 1), (6, 128, 2), (8, 96, 3), (16, 64, 4)])
def test_single_control_rollouts_match_oracle(n_objects, K, seed):
    sc, _ = scenes.make_scene(n_objects, seed=seed)
    start = scenes.cluttered_start(sc, seed + 10)
    rng = np.random.RandomState(seed)
    controls = scenes.sample_controls_towards(rng, sc, start, K, 1)
    dest = scenes.sample_state(rng, sc)
    r = compare_extend(sc, start, controls, dest=dest, mask=1 << 1)
    # contact must actually have happened in this batch, otherwise the test says little
    moved = np.abs(r["want"]["all_states"][:, 0, 3:] - start[None, 3:]).max(axis=1) > 1e-4
    assert moved.mean() > 0.3
    assert r["exact"] == r["n"], f"only {r['exact']} of {r['n']} rollouts bit-identical (within tolerance though)"


def test_random_controls_free_space():
    sc, start = scenes.make_scene(3, seed=1001)
    rng = np.random.RandomState(5)
    controls = scenes.sample_controls(rng, 40, 1)
    compare_extend(sc, start, controls, dest=scenes.sample_state(rng, sc))


@pytest.mark.parametrize("compare_f32", [False, True])
def test_chains_hybrid(compare_f32):
    sc, _ = scenes.make_scene(6, seed=7)
    start = scenes.cluttered_start(sc, 17)
    rng = np.random.RandomState(11)
    K, L = 48, 4
    controls = scenes.sample_controls_towards(rng, sc, start, K, L)
    n_ctrl = rng.randint(1, L + 1, size=K).astype(np.int32)
    dest = scenes.sample_state(rng, sc)
    goals = [(1, float(start[3]) + 0.1, float(start[4]), 0.15)]
    r = compare_extend(sc, start, controls, n_ctrl=n_ctrl, dest=dest, compare_f32=compare_f32, goals=goals,
                       ball_center=start, ball_radius=0.3)
    fl = r["want"]["all_flags"]
    assert ((fl & abi.MPS_F_GOAL) != 0).any(), "no chain reached the goal: weak test"
    assert ((fl & abi.MPS_F_IN_BALL) != 0).any()


def test_invalid_outcomes_and_policy():
    """Robot driven into a wall (robot-static forbidden) and out of bounds; object-object forbidden pair."""
    sc, _ = scenes.make_scene(3, seed=3)
    sc.pair_forbidden[1] |= 1 << 2
    sc.pair_forbidden[2] |= 1 << 1
    start = np.array([0.7, 0.0, 0.0, 0.0, 0.5, 0.3, 0.17, 0.5, 0.0, -0.5, -0.5, 1.0], np.float32)
    controls = np.array([
        [0.5, 0.0, 0.0, 0.9, 0.0],    # robot into the +x wall / out of bounds
        [-0.2, 0.0, 0.5, 0.3, 0.0],   # free motion
        [-0.5, 0.35, 0.0, 0.9, 0.0],  # pushes object 1 into object 2 (forbidden pair)
        [0.0, 0.0, 0.0, 0.1, 0.0],    # null velocity
    ], np.float32)
    r = compare_extend(sc, start, controls, dest=start)
    fl = r["want"]["all_flags"][:, 0]
    assert not (fl[0] & abi.MPS_F_OK)
    assert fl[1] & abi.MPS_F_OK


def test_all_invalid_returns_minus_one():
    sc, _ = scenes.make_scene(3, seed=3)
    start = np.array([0.9, 0.0, 0.0, 0.0, 0.5, 0.3, 0.3, 0.5, 0.0, -0.5, -0.5, 1.0], np.float32)
    controls = np.tile(np.array([0.6, 0.0, 0.0, 1.0, 0.0], np.float32), (5, 1))
    r = compare_extend(sc, start, controls, dest=start)
    assert r["got"].k == -1 and r["got"].n_ok == 0


def test_ties_first_index_wins():
    sc, start = scenes.make_scene(3, seed=1001)
    c = np.array([0.1, 0.05, 0.2, 0.3, 0.0], np.float32)
    controls = np.tile(c, (7, 1))
    r = compare_extend(sc, start, controls, dest=start)
    assert r["got"].k == 0


def test_strict_active_contacts_flag():
    sc, _ = scenes.make_scene(3, seed=3)
    sc.strict_active_contacts = 1
    sc.pair_forbidden[0] |= 1 << 1
    sc.pair_forbidden[1] |= 1 << 0
    start = np.array([0.0, 0.0, 0.0, 0.15, 0.0, 0.3, 0.5, 0.5, 0.0, -0.5, -0.5, 1.0], np.float32)
    controls = np.array([[0.4, 0.0, 0.0, 0.5, 0.0], [-0.4, 0.0, 0.0, 0.5, 0.0]], np.float32)
    r = compare_extend(sc, start, controls, dest=start)
    assert r["want"]["all_flags"][0, 0] & abi.MPS_F_ACTIVE_CONTACT


def test_manifold_overflow_flag():
    sc, _ = scenes.make_scene(8, seed=2)
    sc.max_manifolds = 2
    start = scenes.cluttered_start(sc, 4)
    rng = np.random.RandomState(2)
    controls = scenes.sample_controls_towards(rng, sc, start, 32, 1)
    r = compare_extend(sc, start, controls, dest=start)
    assert ((r["want"]["all_flags"] & abi.MPS_F_OVERFLOW) != 0).any()


def test_propagate_single_call_and_alias():
    sc, _ = scenes.make_scene(3, seed=3)
    start = np.array([0.0, 0.0, 0.0, 0.15, 0.01, 0.3, 0.32, 0.0, 0.0, 0.5, 0.5, 0.0], np.float32)
    control = np.array([0.4, 0.0, 0.0, 0.6, 0.123], np.float32)
    ctx = api.Context(sc)
    ok, out, c, fl = ctx.propagate(start, control)
    wok, wout, wc, wfl, _ = ob.propagate(sc, start, control)
    assert ok == wok and fl == wfl
    assert np.array_equal(c, wc)
    assert np.abs(out - wout).max() <= POS_TOL
    # replay determinism (the reference's verifySolution contract, OraclePushPlanner.cpp:495-519)
    ok2, out2, c2, _ = ctx.propagate(start, control)
    assert np.array_equal(out, out2) and np.array_equal(c, c2)
    ctx.close()


def test_is_valid_batch_matches_oracle():
    sc, _ = scenes.make_scene(6, seed=9)
    rng = np.random.RandomState(9)
    states = np.stack([scenes.sample_state(rng, sc) * (1.08 if i % 7 == 0 else 1.0) for i in range(300)])
    states[5] = scenes.cluttered_start(sc, 1)
    ctx = api.Context(sc)
    valid, flags = ctx.is_valid(states)
    for i in range(len(states)):
        wok, wfl = ob.is_valid(sc, states[i])
        assert bool(valid[i]) == wok and int(flags[i]) == wfl, i
    assert valid.any() and (~valid).any()
    ctx.close()


def test_run_to_run_determinism_large_batch():
    sc, _ = scenes.make_scene(6, seed=21)
    start = scenes.cluttered_start(sc, 22)
    rng = np.random.RandomState(23)
    controls = scenes.sample_controls_towards(rng, sc, start, 2048, 2)
    dest = scenes.sample_state(rng, sc)
    ctx = api.Context(sc)
    a = ctx.extend(start, controls, dump=True, dest=dest, compare_f32=True)
    b = ctx.extend(start, controls, dump=True, dest=dest, compare_f32=True)
    assert a.k == b.k and np.array_equal(a.all_states, b.all_states) and np.array_equal(a.all_flags, b.all_flags)
    # a sample of the batch against the oracle
    idx = rng.choice(2048, 48, replace=False)
    want = ob.extend(sc, start, controls[idx], dest=dest, compare_f32=True)
    assert np.array_equal(a.all_flags[idx], want["all_flags"])
    assert_states_close(a.all_states[idx], want["all_states"], want["all_flags"])
    assert np.array_equal(a.all_dist[idx], want["all_dist"])
    ctx.close()


def test_extend_appends_winner_to_tree():
    sc, _ = scenes.make_scene(3, seed=3)
    start = scenes.cluttered_start(sc, 8)
    rng = np.random.RandomState(3)
    controls = scenes.sample_controls_towards(rng, sc, start, 16, 2)
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    root = tree.add(start)
    r = ctx.extend(start, controls, dest=scenes.sample_state(rng, sc), append_to_tree=True, parent=root)
    assert r.k >= 0 and r.tree_index == 1
    assert tree.size() == 1 + r.n_ok
    for l in range(r.n_ok):
        s, p, c, sl = tree.get(r.tree_index + l)
        assert np.array_equal(s, r.best_states[l])
        assert np.array_equal(c, r.best_controls[l])
        assert p == (root if l == 0 else r.tree_index + l - 1)
    path = tree.backtrack(r.tree_index + r.n_ok - 1)
    assert list(path) == list(range(0, r.n_ok + 1))
    ctx.close()


def test_bad_arguments_are_errors_not_crashes():
    sc, start = scenes.make_scene(3, seed=3)
    ctx = api.Context(sc)
    with pytest.raises(api.MpsError):
        ctx.extend(start, np.zeros((1, abi.MPS_MAX_CHAIN + 1, 5), np.float32))  # L > MPS_MAX_CHAIN
    with pytest.raises(api.MpsError):
        ctx.extend(start, np.zeros((2, 1, 5), np.float32), goals=[(9, 0.0, 0.0, 0.1)])  # goal body out of range
    ctx.close()
    bad = scenes.default_scene()
    bad.n_bodies = 1
    with pytest.raises(api.MpsError):
        api.Context(bad)


def test_gpu_reproduces_committed_physics_golden():
    """The CUDA path against the committed fixtures (no oracle involved at run time)."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "physics_golden_v3.npz"))
    i = 0
    while f"c{i}_meta" in g:
        n_objects, seed, k, n_ok = [int(x) for x in g[f"c{i}_meta"]]
        sc, _ = scenes.make_scene(n_objects, seed=seed)
        ctx = api.Context(sc)
        r = ctx.extend(g[f"c{i}_start"], g[f"c{i}_controls"], dump=True, dest=g[f"c{i}_dest"], compare_f32=True)
        assert r.k == k and r.n_ok == n_ok
        assert np.array_equal(r.all_flags, g[f"c{i}_all_flags"])
        assert np.array_equal(r.all_rest, g[f"c{i}_all_rest"])
        assert np.array_equal(r.all_steps, g[f"c{i}_all_steps"])
        assert np.array_equal(r.all_dist, g[f"c{i}_all_dist"])
        assert_states_close(r.all_states, g[f"c{i}_all_states"], g[f"c{i}_all_flags"], f"golden case {i}")
        ctx.close()
        i += 1
    assert i == 4


def test_gpu_reproduces_the_independent_physics_answers():
    """tests/golden/physics_kat_v3.json: known answers of the physics spec computed by the independent numpy
    restatement (tests/golden/spec_np.py) - neither the oracle nor any kernel wrote them.  The CUDA path must give
    the same bits: final state, flags, rest time, step count; and the same validity flags."""
    import physics_kat as pk

    kat = pk.load()
    for case in kat["propagate"]:
        sc = pk.build_scene(case["scene"])
        ctx = api.Context(sc)
        ok, out, c, fl = ctx.propagate(pk.arr(case["start"]), pk.arr(case["control"]))
        assert ok == case["ok"] and fl == case["flags"], case["name"]
        assert c[4] == np.float32(case["rest"]), case["name"]
        assert np.array_equal(out, pk.arr(case["out"])), (case["name"], out, case["out"])
        r = ctx.extend(pk.arr(case["start"]), pk.arr(case["control"]).reshape(1, 1, 5), dump=True)
        assert int(r.all_steps[0, 0]) == case["steps"], case["name"]
        ctx.close()
    sc = pk.build_scene(kat["valid"]["scene"])
    ctx = api.Context(sc)
    valid, flags = ctx.is_valid(np.stack([pk.arr(c["state"]) for c in kat["valid"]["cases"]]))
    for i, case in enumerate(kat["valid"]["cases"]):
        assert int(flags[i]) == case["flags"] and bool(valid[i]) == (case["flags"] == 0), case["name"]
    ctx.close()


@pytest.mark.parametrize("stop_at_goal", [False, True])
def test_per_chain_starts_long_chains_and_stop_at_goal(stop_at_goal):
    """The shortcutters' candidate replays (RRT.cpp:946-1034): every chain has its own start state, chains are
    longer than an extend's (L = 24 here, limit MPS_MAX_CHAIN = 64) and end at their first goal state."""
    sc, _ = scenes.make_scene(4, seed=21)
    rng = np.random.RandomState(5)
    K, L = 40, 24
    starts = np.stack([scenes.cluttered_start(sc, 100 + (k % 7)) for k in range(K)])
    controls = np.stack([scenes.sample_controls_towards(rng, sc, starts[k], 1, L)[0] for k in range(K)])
    n_ctrl = rng.randint(1, L + 1, size=K).astype(np.int32)
    goals = [(1, float(starts[0][3]) + 0.05, float(starts[0][4]), 0.2)]
    r = compare_extend(sc, None, controls, n_ctrl=n_ctrl, goals=goals, starts=starts, stop_at_goal=stop_at_goal)
    fl = r["want"]["all_flags"]
    goal = (fl & abi.MPS_F_GOAL) != 0
    assert goal.any(), "no chain reached the goal: weak test"
    ran = (fl & abi.MPS_F_RAN) != 0
    if stop_at_goal:
        # nothing is propagated behind the first goal state of a chain
        first = np.where(goal.any(axis=1), goal.argmax(axis=1), L)
        for k in range(K):
            assert not ran[k, first[k] + 1:].any()
    else:
        assert any(ran[k, goal[k].argmax() + 1:].any() for k in range(K) if goal[k].any())
    # chains that share a start state and controls prefix agree; different starts give different results
    assert len({tuple(r["want"]["all_states"][k, 0]) for k in range(K)}) > 7


def test_general_sweep_matches_oracle(monkeypatch):
    """The colour-by-colour sweep used when a body has more than 4 manifolds, forced on through the test hook:
    same bits as the round schedule and as the oracle."""
    monkeypatch.setenv("MPS_FORCE_GENERAL_SWEEP", "1")
    sc, _ = scenes.make_scene(8, seed=3)
    start = scenes.cluttered_start(sc, 23)
    rng = np.random.RandomState(4)
    controls = scenes.sample_controls_towards(rng, sc, start, 64, 2)
    dest = scenes.sample_state(rng, sc)
    r = compare_extend(sc, start, controls, dest=dest)
    assert r["got"].all_steps.sum() > 1000
    assert ((r["want"]["all_flags"] & abi.MPS_F_OK) != 0).any()


@pytest.mark.parametrize("group,n_objects", [(8, 3), (8, 6), (8, 7), (16, 8), (16, 12), (16, 15)])
def test_packed_chains_match_oracle(monkeypatch, group, n_objects):
    """Two / four chains per warp (scenes with at most 16 / 8 bodies, chosen automatically for launches that fill
    the machine, forced here through MPS_ROLLOUT_GROUP): same bits as one chain per warp and as the oracle."""
    monkeypatch.setenv("MPS_ROLLOUT_GROUP", str(group))
    sc, _ = scenes.make_scene(n_objects, seed=30 + n_objects)
    start = scenes.cluttered_start(sc, 40 + n_objects)
    rng = np.random.RandomState(n_objects)
    K, L = 70, 3   # not a multiple of the chains per CTA: the last groups run out of work at different times
    controls = scenes.sample_controls_towards(rng, sc, start, K, L)
    n_ctrl = rng.randint(1, L + 1, size=K).astype(np.int32)
    dest = scenes.sample_state(rng, sc)
    goals = [(1, float(start[3]) + 0.1, float(start[4]), 0.15)]
    r = compare_extend(sc, start, controls, n_ctrl=n_ctrl, dest=dest, goals=goals, ball_center=start, ball_radius=0.3)
    assert r["got"].all_steps.sum() > 2000
    assert ((r["want"]["all_flags"] & abi.MPS_F_OK) != 0).any()


@pytest.mark.parametrize("group", [8, 16])
def test_packed_chains_redo_pass(monkeypatch, group):
    """A packed launch gives each chain few manifold slots; chains that need more are run again one per warp.
    With the capacity forced down to 1 almost every contact-rich chain takes that path: same bits."""
    monkeypatch.setenv("MPS_ROLLOUT_GROUP", str(group))
    monkeypatch.setenv("MPS_ROLLOUT_SMALL_CAP", "1")
    sc, _ = scenes.make_scene(6, seed=36)
    start = scenes.cluttered_start(sc, 46)
    rng = np.random.RandomState(8)
    controls = scenes.sample_controls_towards(rng, sc, start, 90, 2)
    dest = scenes.sample_state(rng, sc)
    r = compare_extend(sc, start, controls, dest=dest)
    assert (r["got"].all_steps.sum(axis=1) > 0).all()


def test_baseline_config2_whole_batch_matches_oracle():
    """BASELINE config 2 at full size (the bench workload: B = 7, K = 1024 chains of up to 4 controls): every
    chain of one batch against the oracle (pthread fan-out, a fraction of a second)."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    sc, start, batches, goals = bench.make_workload([0], 2)
    controls, n_ctrl, dest = batches[1]
    ctx = api.Context(sc)
    got = ctx.extend(start, controls, dump=True, n_ctrl=n_ctrl, dest=dest, compare_f32=True, goals=goals)
    want = ob.extend(sc, start, controls, n_ctrl=n_ctrl, dest=dest, compare_f32=True, goals=goals, threads=os.cpu_count() or 1)
    assert np.array_equal(got.all_flags, want["all_flags"]) and np.array_equal(got.all_steps, want["all_steps"])
    assert np.array_equal(got.all_rest, want["all_rest"]) and np.array_equal(got.all_dist, want["all_dist"])
    _, _, exact, n = assert_states_close(got.all_states, want["all_states"], want["all_flags"])
    assert exact == n and n > 2000
    assert got.k == want["k"] and got.distance == want["distance"]
    ctx.close()


def test_baseline_config4_shape_determinism_and_sampled_parity():
    """BASELINE config 4's shape (B = 17, 65 536 single-control rollouts in one launch): run-to-run identical,
    a checksum over all outputs, and 256 sampled chains against the oracle."""
    sc, _ = scenes.make_scene(16, seed=1004)
    start = scenes.cluttered_start(sc, 1004)
    rng = np.random.RandomState(16)
    K = 65536
    controls = scenes.sample_controls_towards(rng, sc, start, K, 1)
    dest = scenes.sample_state(rng, sc)
    ctx = api.Context(sc)
    a = ctx.extend(start, controls, dump=True, dest=dest)
    b = ctx.extend(start, controls, dump=True, dest=dest)
    assert a.k == b.k and a.distance == b.distance
    assert np.array_equal(a.all_states, b.all_states) and np.array_equal(a.all_flags, b.all_flags)
    ok = (a.all_flags[:, 0] & abi.MPS_F_OK) != 0
    assert 0.05 < ok.mean() <= 1.0
    # the winner is the first chain with the smallest distance among the valid ones
    d = np.where(ok, a.all_dist, np.inf)
    assert a.k == int(np.argmin(d)) and a.distance == d.min()
    idx = rng.choice(K, 256, replace=False)
    want = ob.extend(sc, start, controls[idx], dest=dest, threads=8)
    assert np.array_equal(a.all_flags[idx], want["all_flags"]) and np.array_equal(a.all_dist[idx], want["all_dist"])
    assert_states_close(a.all_states[idx], want["all_states"], want["all_flags"])
    ctx.close()


@pytest.mark.parametrize("env,n_objects,K,L", [
    ({"MPS_ROLLOUT_MINB": "3", "MPS_ROLLOUT_SOLO_BUILD": "0"}, 6, 200, 3),   # <4,3,32>, one CTA per SM
    ({"MPS_ROLLOUT_MINB": "3"}, 10, 150, 2),                                  # <4,1,32>: up to four rounds in registers
    ({"MPS_ROLLOUT_MINB": "4", "MPS_ROLLED_MAX_BODIES": "0"}, 10, 150, 2),    # <4,4,32>, sweeps inlined per round
    ({"MPS_ROLLOUT_MINB": "4", "MPS_ROLLED_MAX_BODIES": "32"}, 10, 150, 2),   # <4,4,32,rolled>
    ({"MPS_ROLLOUT_MINB": "4", "MPS_ROLLED_MAX_BODIES": "32"}, 16, 100, 1),
    ({"MPS_ROLLOUT_GROUP": "16"}, 8, 150, 2),                                 # <4,4,16,rolled>
    ({"MPS_ROLLOUT_GROUP": "8"}, 6, 150, 2),                                  # <4,4,8>
])
def test_every_build_of_the_rollout_kernel_matches_oracle(monkeypatch, env, n_objects, K, L):
    """The launcher picks one of six instantiations of rollout_kernel by launch shape; here each is forced through the
    launcher's knobs (read when the context is created) on the same kind of contact-rich batch: same bits as the
    oracle whatever the build."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    sc, _ = scenes.make_scene(n_objects, seed=50 + n_objects)
    start = scenes.cluttered_start(sc, 60 + n_objects)
    rng = np.random.RandomState(n_objects)
    controls = scenes.sample_controls_towards(rng, sc, start, K, L)
    n_ctrl = rng.randint(1, L + 1, size=K).astype(np.int32)
    dest = scenes.sample_state(rng, sc)
    r = compare_extend(sc, start, controls, n_ctrl=n_ctrl, dest=dest)
    assert r["got"].all_steps.sum() > 2000
    assert ((r["want"]["all_flags"] & abi.MPS_F_OK) != 0).any()


def test_nine_bodies_many_rollouts_two_chains_per_warp_sampled_parity():
    """A 9-body scene, K >= 16 384 single-control rollouts from one start state: the launcher packs two chains per
    warp (rollout_kernel<4,4,16,rolled>) on its own: run-to-run identical, the winner is the first chain with the
    smallest distance among the valid ones, and 256 sampled chains equal the oracle's bit for bit."""
    sc, _ = scenes.make_scene(8, seed=1005)
    start = scenes.cluttered_start(sc, 1005)
    rng = np.random.RandomState(8)
    K = 16384 + 37
    controls = scenes.sample_controls_towards(rng, sc, start, K, 1)
    dest = scenes.sample_state(rng, sc)
    ctx = api.Context(sc)
    a = ctx.extend(start, controls, dump=True, dest=dest)
    b = ctx.extend(start, controls, dump=True, dest=dest)
    assert a.k == b.k and a.distance == b.distance
    assert np.array_equal(a.all_states, b.all_states) and np.array_equal(a.all_flags, b.all_flags)
    ok = (a.all_flags[:, 0] & abi.MPS_F_OK) != 0
    assert 0.05 < ok.mean() <= 1.0
    d = np.where(ok, a.all_dist, np.inf)
    assert a.k == int(np.argmin(d)) and a.distance == d.min()
    idx = rng.choice(K, 256, replace=False)
    want = ob.extend(sc, start, controls[idx], dest=dest, threads=8)
    assert np.array_equal(a.all_flags[idx], want["all_flags"]) and np.array_equal(a.all_dist[idx], want["all_dist"])
    assert np.array_equal(a.all_steps[idx], want["all_steps"]) and np.array_equal(a.all_rest[idx], want["all_rest"])
    _, _, exact, n = assert_states_close(a.all_states[idx], want["all_states"], want["all_flags"])
    assert exact == n
    ctx.close()


def test_sharded_extend_world_1_equals_ctx_extend_and_appends_on_the_device():
    """ShardedExtend on one rank: rollouts, record copy, merge kernel and the tree append all on the device, only the
    24-byte head comes back; same winner, same appended nodes as ctx.extend(append_to_tree=True)."""
    import torch

    from manipulation_planning_suite_b200 import sharded

    sc, _ = scenes.make_scene(6, seed=21)
    start = scenes.cluttered_start(sc, 21)
    rng = np.random.RandomState(21)
    K, L = 96, 3
    controls = scenes.sample_controls_towards(rng, sc, start, K, L)
    n_ctrl = rng.randint(1, L + 1, size=K).astype(np.int32)
    dest = scenes.sample_state(rng, sc)
    ref_ctx = api.Context(sc)
    ref_tree = api.NearestNeighbors(ref_ctx)
    ref_tree.add(start)
    want = ref_ctx.extend(start, controls, n_ctrl=n_ctrl, dest=dest, compare_f32=True, append_to_tree=True, parent=0)
    # the ctx of the sharded extend lives on its own non-blocking stream (not torch's current one): the stream
    # order must still hold
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.add(start)
    sh = sharded.ShardedExtend(ctx)
    for _ in range(2):  # twice: the second run appends below the same parent again
        sh.upload(start, controls, n_ctrl=n_ctrl, dest=dest, compare_f32=True, append_to_tree=True, parent=0)
        sh.launch()
        head = sh.head()
        assert (head.k, head.n_ok, head.goal_step, head.distance) == (want.k, want.n_ok, want.goal_step, want.distance)
    assert head.tree_index == 1 + want.n_ok
    rec = sh.result()
    assert np.array_equal(rec.states, want.best_states) and np.array_equal(rec.controls, want.best_controls)
    assert tree.size() == 1 + 2 * want.n_ok and ref_tree.size() == 1 + want.n_ok
    for i in range(want.n_ok):
        a, b = tree.get(1 + i), ref_tree.get(1 + i)
        c = tree.get(1 + want.n_ok + i)
        for x, y in zip(a, b):          # (state, parent, control, slice id): same node as the fused append made
            assert np.array_equal(x, y)
        assert np.array_equal(a[0], c[0]) and np.array_equal(a[2], c[2])   # second run: same chain again
        assert c[1] == (0 if i == 0 else want.n_ok + i)
    torch.cuda.synchronize()
    ctx.close()
    ref_ctx.close()


def test_merge_kernel_follows_the_reference_tie_rules():
    """mps_extend_merge_records on hand-made gathered records: strict <, first rank wins equal keys, float compare
    when asked, records without a valid chain are skipped, no valid record -> k = -1; equal to the host mirror
    sharded.merge_records."""
    import torch

    from manipulation_planning_suite_b200 import sharded

    sc, _ = scenes.make_scene(3, seed=5)
    B, L = sc.n_bodies, 2
    start = scenes.cluttered_start(sc, 5)
    ctx = api.Context(sc)
    ctx.extend_upload(start, np.zeros((1, L, 5), np.float32), dest=start)   # shapes the ctx's record buffer
    rng = np.random.RandomState(0)

    def rec(k, d):
        return sharded.Record(k, 1 if k >= 0 else 0, -1, -1, d, rng.rand(L, 3 * B).astype(np.float32),
                              rng.rand(L, 5).astype(np.float32), np.array([3, 0], np.uint32))

    d0 = 1.0 + 2.0 ** -30   # equal to 1.0 as a float, larger as a double
    cases = [([rec(5, 2.0), rec(40, 1.5), rec(70, 1.5)], True, False),
             ([rec(-1, np.inf), rec(40, 3.0), rec(70, 2.5)], True, False),
             ([rec(5, d0), rec(40, 1.0), rec(70, 7.0)], True, True),
             ([rec(5, d0), rec(40, 1.0), rec(70, 7.0)], True, False),
             ([rec(-1, np.inf), rec(-1, np.inf)], True, False),
             ([rec(-1, np.inf), rec(33, 9.0), rec(70, 1.0)], False, False)]
    for recs, has_dest, cf32 in cases:
        host = np.stack([sharded.pack_record(r, B, L) for r in recs])
        dev = torch.from_numpy(host).cuda()
        torch.cuda.synchronize()
        head = ctx.merge_records(dev.data_ptr(), len(recs), has_dest, cf32, wait=True)
        w = sharded.merge_records(recs, has_dest, cf32)
        if w < 0:
            assert head.k == -1 and head.n_ok == 0 and np.isinf(head.distance)
        else:
            assert (head.k, head.distance) == (recs[w].k, recs[w].distance)
            got = ctx.extend_fetch()
            assert np.array_equal(got.best_states, recs[w].states) and np.array_equal(got.best_controls, recs[w].controls)
    ctx.close()
