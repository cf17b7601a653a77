"""The CUDA kernels' SOURCE against the CPU oracle, without a GPU: rollout_kernel / select_kernel /
validity_kernel are compiled for the host through the SIMT emulator of tests/emu (test infrastructure, see
tests/emu/cuda_emu.h) and must reproduce the oracle bit for bit, in every instantiation the launcher uses
(one chain per warp with either register budget, packed lane groups, packed with the small manifold
capacity + redo pass).  The `-m gpu` tests make the same comparison through libmps_b200.so on the device."""
import numpy as np
import pytest

import emu_binding as eb
import oracle_binding as ob
from manipulation_planning_suite_b200 import scenes

VARIANTS = [(32, 1, 0), (32, 3, 0), (32, 4, 0), (32, 5, 0), (8, 4, 14), (8, 4, 2), (16, 4, 14), (16, 4, 2)]


def _same(r, w):
    for key in ("all_states", "all_flags", "all_rest", "all_steps", "all_dist", "best_states", "best_controls",
                "best_flags"):
        assert np.array_equal(r[key], w[key]), key
    assert (r["k"], r["n_ok"], r["goal_step"]) == (w["k"], w["n_ok"], w["goal_step"])
    assert r["distance"] == w["distance"] or (np.isinf(r["distance"]) and np.isinf(w["distance"]))


@pytest.mark.parametrize("group,minb,mcap_small", VARIANTS)
def test_cfg2_shape_chains_match_oracle(group, minb, mcap_small):
    sc, _ = scenes.make_scene(6, seed=1002)
    start = scenes.cluttered_start(sc, 1002)
    rng = np.random.RandomState(3)
    K, L = 24, 4
    c = scenes.sample_controls_towards(rng, sc, start, K, L)
    n = rng.randint(1, L + 1, size=K).astype(np.int32)
    d = scenes.sample_state(rng, sc)
    kw = dict(n_ctrl=n, dest=d, compare_f32=True, goals=[(1, 0.5, 0.5, 0.1)])
    w = ob.extend(sc, start, c, **kw)
    r = eb.extend(sc, start, c, group=group, minb=minb, mcap_small=mcap_small, **kw)
    _same(r, w)
    assert r["counters"]["steps"] >= int(w["all_steps"].sum())


@pytest.mark.parametrize("n_objects,group,minb", [(3, 32, 4), (3, 8, 4), (10, 32, 4), (16, 32, 4), (12, 16, 4), (8, 16, 4), (9, 16, 4),
                                                  (10, 32, 1), (16, 32, 1), (16, 32, 3), (10, 32, 5), (16, 32, 5)])
def test_single_control_rollouts_match_oracle(n_objects, group, minb):
    sc, _ = scenes.make_scene(n_objects, seed=1000 + n_objects)
    start = scenes.cluttered_start(sc, 1000 + n_objects)
    rng = np.random.RandomState(n_objects)
    K = 12
    c = scenes.sample_controls_towards(rng, sc, start, K, 1)
    d = scenes.sample_state(rng, sc)
    w = ob.extend(sc, start, c, dest=d, mask=1 << 1)
    r = eb.extend(sc, start, c, group=group, minb=minb, dest=d, mask=1 << 1)
    _same(r, w)


def test_validity_kernel_matches_oracle():
    sc, _ = scenes.make_scene(6, seed=1002)
    rng = np.random.RandomState(5)
    states = np.stack([scenes.sample_state(rng, sc) for _ in range(40)])
    states[::3, 3:5] = states[::3, 0:2]  # object 1 on top of the robot: touching / infeasible states
    valid, flags = eb.is_valid_batch(sc, states)
    for i in range(len(states)):
        ok, f = ob.is_valid(sc, states[i])
        assert (bool(valid[i]), int(flags[i])) == (ok, f)


@pytest.mark.parametrize("group", [32, 8])
def test_kernel_source_reproduces_the_independent_physics_answers(group):
    """tests/golden/physics_kat_v3.json (from the numpy restatement of the spec, tests/golden/spec_np.py) against the
    rollout / validity kernels' source run through the emulator."""
    import physics_kat as pk

    kat = pk.load()
    for case in kat["propagate"]:
        sc = pk.build_scene(case["scene"])
        r = eb.extend(sc, pk.arr(case["start"]), pk.arr(case["control"]).reshape(1, 1, 5), group=group, minb=4)
        assert int(r["all_flags"][0, 0]) == case["flags"] and int(r["all_steps"][0, 0]) == case["steps"], case["name"]
        assert r["all_rest"][0, 0] == np.float32(case["rest"]), case["name"]
        assert np.array_equal(r["all_states"][0, 0], pk.arr(case["out"])), case["name"]
    sc = pk.build_scene(kat["valid"]["scene"])
    valid, flags = eb.is_valid_batch(sc, np.stack([pk.arr(c["state"]) for c in kat["valid"]["cases"]]))
    for i, case in enumerate(kat["valid"]["cases"]):
        assert int(flags[i]) == case["flags"] and bool(valid[i]) == (case["flags"] == 0), case["name"]


@pytest.mark.parametrize("group", [32, 8])
def test_sparse_scene_free_flight_and_approach_match_oracle(group):
    """objects spread over the workspace: most steps are free flight (only the robot's lane steps while the candidate
    list is empty and nothing else moves), with approaches, contacts and separations in between"""
    sc, start = scenes.make_scene(5, seed=1001)
    rng = np.random.RandomState(8)
    K, L = 20, 3
    c = scenes.sample_controls(rng, K, L)
    c[::2] = scenes.sample_controls_towards(rng, sc, start, K, L, noise=0.1)[::2]   # half of them aimed at objects
    d = scenes.sample_state(rng, sc)
    w = ob.extend(sc, start, c, dest=d, compare_f32=True)
    r = eb.extend(sc, start, c, group=group, minb=4, dest=d, compare_f32=True)
    _same(r, w)
    moved = np.abs(w["all_states"][:, :, 3:] - start[None, None, 3:]).max(axis=(1, 2)) > 1e-4
    assert 2 <= moved.sum() <= K - 2   # both kinds of chains are present
    assert r["counters"]["steps"] == int(w["all_steps"].sum())
