"""Multi-instance batches (api.Forest) against the oracle, and the multi-seed NaiveRRT sweep."""
import numpy as np
import pytest

from manipulation_planning_suite_b200 import abi, api, multi_seed, scenes

import oracle_binding as ob

pytestmark = pytest.mark.gpu


def test_forest_extend_and_nearest_match_oracle_per_instance():
    sc, _ = scenes.make_scene(6, seed=31)
    B = sc.n_bodies
    rng = np.random.RandomState(31)
    n_inst, M, K, L = 12, 7, 9, 2
    roots = np.stack([scenes.cluttered_start(sc, 100 + i) for i in range(n_inst)])
    ctx = api.Context(sc)
    forest = api.Forest(ctx, n_inst, 40)
    forest.reset(roots)
    assert (forest.sizes() == 1).all()
    inst = rng.choice(n_inst, M, replace=False).astype(np.int32)
    for round_ in range(3):
        sizes = forest.sizes()
        parents = np.array([rng.randint(0, sizes[i]) for i in inst], np.int32)
        controls = np.stack([scenes.sample_controls_towards(rng, sc, roots[i], K, L) for i in inst])
        n_ctrl = rng.randint(1, L + 1, size=(M, K)).astype(np.int32)
        dests = np.stack([scenes.sample_state(rng, sc) for _ in inst])
        masks = np.array([1 << rng.randint(0, B) for _ in inst], np.uint32)
        before = [forest.read(int(i), 0, int(sizes[i])) for i in inst]
        res = forest.extend(inst, parents, controls, n_ctrl=n_ctrl, dests=dests, masks=masks, compare_f32=True,
                            goals=[(1, 0.0, 0.0, 0.2)], append=True)
        new_sizes = forest.sizes()
        for m, i in enumerate(inst):
            start = before[m][0][parents[m]]
            want = ob.extend(sc, start, controls[m], n_ctrl=n_ctrl[m], dest=dests[m], mask=int(masks[m]),
                             compare_f32=True, goals=[(1, 0.0, 0.0, 0.2)])
            assert res["k"][m] == want["k"], (round_, m)
            assert res["n_ok"][m] == want["n_ok"] and res["goal_step"][m] == want["goal_step"]
            if want["k"] >= 0:
                assert res["distance"][m] == want["distance"]
                assert np.array_equal(res["states"][m], want["best_states"])
                assert np.array_equal(res["controls"][m], want["best_controls"])
                assert np.array_equal(res["flags"][m], want["best_flags"])
                n_app = want["goal_step"] + 1 if want["goal_step"] >= 0 else want["n_ok"]
                assert new_sizes[i] == sizes[i] + n_app and res["tree_index"][m] == sizes[i]
                st, pa, co = forest.read(int(i), int(sizes[i]), n_app)
                assert np.array_equal(st, want["best_states"][:n_app])
                assert pa[0] == parents[m] and all(pa[j] == sizes[i] + j - 1 for j in range(1, n_app))
            else:
                assert new_sizes[i] == sizes[i]
        # untouched trees stay untouched
        others = np.setdiff1d(np.arange(n_inst), inst)
        assert (new_sizes[others] == 1).all()
        # nearest per instance over its own tree
        qs = np.stack([scenes.sample_state(rng, sc) for _ in inst])
        gi, gd = forest.nearest(inst, qs)
        for m, i in enumerate(inst):
            st, _, _ = forest.read(int(i), 0, int(new_sizes[i]))
            wi, wd = ob.nn_linear(sc, st, qs[m])
            assert gi[m] == wi and gd[m] == wd
    ctx.close()


@pytest.mark.parametrize("device_sampling", [False, True])
def test_multi_seed_naive_rrt_sweep_solutions_replay(device_sampling):
    """Lockstep multi-seed NaiveRRT (BASELINE config 5), with the random draws on the host or the whole iteration
    on the device (mps_forest_naive_step): solutions replay on the oracle, the split over devices is invisible."""
    sc, start = scenes.make_scene(3, seed=1001)
    goals = [(1, float(start[3]) + 0.2, float(start[4]) + 0.1, 0.12)]
    ctx = api.Context(sc)
    res, forest = multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=64, num_control_samples=10,
                                             max_iterations=150, seed=7, goal_bias=0.2, target_bias=0.4,
                                             device_sampling=device_sampling)
    assert res.rollouts > 0 and res.iterations >= 1
    solved = np.flatnonzero(res.solved_iteration >= 0)
    assert len(solved) >= 4, f"only {len(solved)} of 64 instances solved"
    for i in solved[:6]:
        st, co = multi_seed.extract_path(forest, int(i), int(res.goal_nodes[i]))
        assert np.array_equal(st[0], start)
        cur = st[0]
        for j in range(1, len(st)):  # replay through the CPU oracle: bit-identical states and rest times
            ok, out, c, fl, _ = ob.propagate(sc, cur, co[j])
            assert ok and np.array_equal(out, st[j]) and c[4] == co[j][4]
            cur = out
        g = goals[0]
        assert np.hypot(cur[3 * g[0]] - g[1], cur[3 * g[0] + 1] - g[2]) <= g[3] + 1e-6
    # sharding invariance: instances [32, 64) run alone give the same outcome as inside the full sweep
    ctx2 = api.Context(sc)
    res2, _ = multi_seed.naive_rrt_sweep(ctx2, start, goals, n_instances=32, num_control_samples=10, max_iterations=150,
                                         seed=7, goal_bias=0.2, target_bias=0.4, instance_offset=32,
                                         device_sampling=device_sampling)
    assert np.array_equal(res2.solved_iteration, res.solved_iteration[32:])
    ctx.close()
    ctx2.close()


def test_forest_calls_validate_what_the_device_would_dereference():
    """a forest that was reserved but never reset has no roots; parents / goal bodies outside their ranges and a
    workspace without finite bounds are argument errors, not out-of-bounds reads or NaN samples"""
    from manipulation_planning_suite_b200 import abi

    sc, start = scenes.make_scene(3, seed=1001)
    ctx = api.Context(sc)
    forest = api.Forest(ctx, 4, 16)
    ids = np.arange(4, dtype=np.int32)
    controls = np.zeros((4, 2, 1, 5), np.float32)
    controls[..., 3] = 0.1
    with pytest.raises(api.MpsError) as e:
        forest.extend(ids, np.zeros(4, np.int32), controls)
    assert e.value.code == abi.MPS_ERR_STATE
    with pytest.raises(api.MpsError) as e:
        forest.nearest(ids, np.tile(start, (4, 1)))
    assert e.value.code == abi.MPS_ERR_STATE
    forest.reset(np.tile(start, (4, 1)))
    for bad in (-1, 10 ** 6):
        parents = np.zeros(4, np.int32)
        parents[2] = bad
        with pytest.raises(api.MpsError) as e:
            forest.extend(ids, parents, controls)
        assert e.value.code == abi.MPS_ERR_ARG
    with pytest.raises(api.MpsError) as e:
        forest.extend(ids, np.zeros(4, np.int32), controls, goals=[(17, 0.0, 0.0, 0.1)])
    assert e.value.code == abi.MPS_ERR_ARG
    forest.extend(ids, np.zeros(4, np.int32), controls)   # and the valid call still runs
    ctx.close()
    # +-FLT_MAX workspace (the scene default): sampling states from it is refused
    sc2, start2 = scenes.make_scene(3, seed=1001)
    sc2.x_min, sc2.x_max = -3.4e38, 3.4e38
    ctx2 = api.Context(sc2)
    f2 = api.Forest(ctx2, 2, 8)
    f2.reset(np.tile(start2, (2, 1)))
    with pytest.raises(api.MpsError) as e:
        f2.naive_step(np.arange(2, dtype=np.int32), 0, 1, [(1, 0.0, 0.0, 0.1)])
    assert e.value.code == abi.MPS_ERR_ARG
    with pytest.raises(ValueError):
        multi_seed.naive_rrt_sweep(ctx2, start2, [(1, 0.0, 0.0, 0.1)], n_instances=2, max_iterations=2)
    ctx2.close()
