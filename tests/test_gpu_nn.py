"""GPU parity of the nearest-neighbour scan against the oracle's linear double-precision scan.
Bar: index bit-exact (ties -> lowest index) and distance identical as a double."""
import numpy as np
import pytest

from manipulation_planning_suite_b200 import api, scenes

import oracle_binding as ob

pytestmark = pytest.mark.gpu


def _masks(B):
    full = (1 << B) - 1
    return [full, full & ~1, 1]  # all, all-but-robot (ObjectArrangementDistanceFn), robot only (RobotStateDistanceFn)


@pytest.mark.parametrize("n_objects,N", [(3, 1), (3, 5), (6, 1023), (10, 4096), (10, 50001), (16, 20000)])
def test_nearest_matches_linear_scan(n_objects, N):
    sc, _ = scenes.make_scene(n_objects, seed=N)
    B = sc.n_bodies
    rng = np.random.RandomState(N)
    states = np.stack([scenes.sample_state(rng, sc) for _ in range(N)])
    slices = rng.randint(0, 5, size=N).astype(np.int32)
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.add_batch(states, parents=np.arange(-1, N - 1), slice_ids=slices)
    assert tree.size() == N
    weights = rng.uniform(0.5, 2.0, size=B).astype(np.float32)
    for qi in range(6):
        q = scenes.sample_state(rng, sc)
        for mask in _masks(B):
            for w in (None, weights):
                for filt in (-1, 2):
                    gi, gd = tree.nearest(q, weights=w, mask=mask, slice_filter=filt)
                    wi, wd = ob.nn_linear(sc, states, q, weights=w, mask=mask, slice_ids=slices, slice_filter=filt)
                    assert gi == wi, (qi, mask, filt, gi, wi, gd, wd)
                    if wi >= 0:
                        assert gd == wd
    ctx.close()


def test_ties_resolve_to_lowest_index_and_exact_hit():
    sc, _ = scenes.make_scene(10, seed=77)
    rng = np.random.RandomState(77)
    base = np.stack([scenes.sample_state(rng, sc) for _ in range(3000)])
    states = np.concatenate([base, base[100:200], base[:50]])  # duplicates at higher indices
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.add_batch(states)
    for j in (0, 17, 49, 100, 150, 199, 2500):
        gi, gd = tree.nearest(states[j])
        assert gi == j and gd == 0.0
    # robot-only metric: many nodes share the robot sub-state -> massive ties
    states2 = states.copy()
    states2[:, :3] = states2[7, :3]
    tree.clear()
    tree.add_batch(states2)
    gi, gd = tree.nearest(states2[1234], mask=1)
    assert gi == 0 and gd == 0.0
    ctx.close()


def test_empty_tree_and_filter_without_members():
    sc, _ = scenes.make_scene(3, seed=5)
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.clear()
    q = np.zeros(3 * sc.n_bodies, np.float32)
    gi, gd = tree.nearest(q)
    assert gi == -1 and np.isinf(gd)
    tree.add(q, slice_id=3)
    gi, gd = tree.nearest(q, slice_filter=4)
    assert gi == -1
    gi, gd = tree.nearest(q, slice_filter=3)
    assert gi == 0
    ctx.close()


def test_growth_preserves_nodes_and_batch_queries():
    sc, _ = scenes.make_scene(6, seed=6)
    rng = np.random.RandomState(6)
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx, capacity=8)
    all_states = []
    for chunk in (3, 200, 5000, 9000):
        st = np.stack([scenes.sample_state(rng, sc) for _ in range(chunk)])
        first = tree.add_batch(st)
        assert first == sum(len(s) for s in all_states)
        all_states.append(st)
    states = np.concatenate(all_states)
    assert tree.size() == len(states)
    assert np.array_equal(tree.read_states(0, len(states)), states)
    Q = 33
    qs = np.stack([scenes.sample_state(rng, sc) for _ in range(Q)])
    B = sc.n_bodies
    masks = np.array([_masks(B)[i % 3] for i in range(Q)], np.uint32)
    gi, gd = tree.nearest_batch(qs, masks=masks)
    for i in range(Q):
        wi, wd = ob.nn_linear(sc, states, qs[i], mask=int(masks[i]))
        assert gi[i] == wi and gd[i] == wd
    ctx.close()


def test_adversarial_descending_distances():
    """Every node improves on the previous one: the fp32 pre-filter passes everything to the double
    path; the answer must still be exact."""
    sc, _ = scenes.make_scene(3, seed=1)
    B = sc.n_bodies
    N = 20000
    q = np.zeros(3 * B, np.float32)
    states = np.zeros((N, 3 * B), np.float32)
    states[:, 0] = np.linspace(0.9, 1e-6, N).astype(np.float32)
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.add_batch(states)
    gi, gd = tree.nearest(q)
    wi, wd = ob.nn_linear(sc, states, q)
    assert gi == wi and gd == wd
    ctx.close()


def test_million_node_properties():
    """BASELINE config 3 size: 1M nodes, B = 11.  Size-independent properties + an oracle check on
    the real data (the oracle's scan of 1M x 11 bodies takes well under a second per query)."""
    sc, _ = scenes.config_scene(3)
    N = 1_000_000
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.fill_random(N, seed=1003, n_slices=64)
    assert tree.size() == N
    states = tree.read_states(0, N)
    assert np.isfinite(states).all()
    rng = np.random.RandomState(3)
    # exact hit: querying a stored node returns it (distance 0) unless an earlier duplicate exists
    for j in (0, 1, 999_999, 123_456):
        gi, gd = tree.nearest(states[j])
        assert gd == 0.0 and np.array_equal(states[gi], states[j]) and gi <= j
    B = sc.n_bodies
    for mask in _masks(B):
        q = scenes.sample_state(rng, sc)
        gi, gd = tree.nearest(q, mask=mask)
        wi, wd = ob.nn_linear(sc, states, q, mask=mask)
        assert gi == wi and gd == wd
    ctx.close()


def test_single_query_paths_agree():
    """One query can travel three ways: in the kernel parameters with the result written to pinned host memory
    (mps_tree_nearest), in the kernel parameters with the result fetched by a copy (upload / launch / fetch with
    Q = 1), and through the device query buffers (a batch).  Same index and the same double every time."""
    sc, _ = scenes.make_scene(8, seed=77)
    B = sc.n_bodies
    rng = np.random.RandomState(77)
    N = 30011
    states = np.stack([scenes.sample_state(rng, sc) for _ in range(N)])
    slices = rng.randint(0, 7, size=N).astype(np.int32)
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.add_batch(states, parents=np.arange(-1, N - 1), slice_ids=slices)
    weights = rng.uniform(0.5, 2.0, size=B).astype(np.float32)
    for mask in _masks(B):
        for w in (None, weights):
            for filt in (-1, 3):
                q = scenes.sample_state(rng, sc)
                wi, wd = ob.nn_linear(sc, states, q, weights=w, mask=mask, slice_ids=slices, slice_filter=filt)
                i1, d1 = tree.nearest(q, weights=w, mask=mask, slice_filter=filt)
                tree.nearest_upload(q[None, :], weights=w, masks=np.array([mask], np.uint32),
                                    slice_filters=np.array([filt], np.int32))
                tree.nearest_launch()
                tree.nearest_launch()  # relaunching a staged query is allowed
                i2, d2 = tree.nearest_fetch()
                q2 = np.stack([q, scenes.sample_state(rng, sc)])
                i3, d3 = tree.nearest_batch(q2, weights=w, masks=np.array([mask, mask], np.uint32),
                                            slice_filters=np.array([filt, filt], np.int32))
                assert i1 == wi and int(i2[0]) == wi and int(i3[0]) == wi
                assert d1 == wd and float(d2[0]) == wd and float(d3[0]) == wd
    ctx.close()


@pytest.mark.parametrize("n_objects,N,Q", [(10, 60_011, 37), (6, 1_003, 8), (16, 20_000, 70)])
def test_tiled_batches_match_linear_scan(n_objects, N, Q):
    """Batches of >= 8 queries take the tiled scan (eight queries per CTA share every load): per-query masks, slice
    filters, weights, duplicate nodes (ties to the lowest index), queries that sit on nodes, a ragged last tile."""
    sc, _ = scenes.make_scene(n_objects, seed=N)
    B = sc.n_bodies
    rng = np.random.RandomState(N)
    states = np.stack([scenes.sample_state(rng, sc) for _ in range(N)])
    states[N // 2:N // 2 + 50] = states[:50]  # duplicates at higher indices
    slices = rng.randint(0, 5, size=N).astype(np.int32)
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.add_batch(states, slice_ids=slices)
    qs = np.stack([scenes.sample_state(rng, sc) for _ in range(Q)])
    qs[3] = states[17]
    qs[Q - 1] = states[N // 2 + 7]  # an exact hit on a duplicated node: the lower index wins
    masks = np.array([_masks(B)[i % 3] for i in range(Q)], np.uint32)
    filters = np.array([(-1, 2, 4, 9)[i % 4] for i in range(Q)], np.int32)  # slice 9 has no members
    weights = rng.uniform(0.5, 2.0, size=B).astype(np.float32)
    for w in (None, weights):
        for m, f in ((None, None), (masks, None), (masks, filters)):
            gi, gd = tree.nearest_batch(qs, weights=w, masks=m, slice_filters=f)
            for i in range(Q):
                wi, wd = ob.nn_linear(sc, states, qs[i], weights=w, mask=None if m is None else int(m[i]),
                                      slice_ids=slices, slice_filter=-1 if f is None else int(f[i]))
                assert gi[i] == wi, (i, gi[i], wi)
                if wi >= 0:
                    assert gd[i] == wd
    ctx.close()


def test_tiled_batch_with_overflowing_candidate_lists():
    """Every node improves on the previous one for every query of the tile: the candidate lists overflow and the
    CTAs take their second, exact pass."""
    sc, _ = scenes.make_scene(3, seed=1)
    B = sc.n_bodies
    N = 40000
    states = np.zeros((N, 3 * B), np.float32)
    states[:, 0] = np.linspace(0.9, 1e-6, N).astype(np.float32)
    qs = np.zeros((11, 3 * B), np.float32)
    qs[:, 1] = np.linspace(0.0, 0.01, 11).astype(np.float32)
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.add_batch(states)
    gi, gd = tree.nearest_batch(qs)
    for i in range(len(qs)):
        wi, wd = ob.nn_linear(sc, states, qs[i])
        assert gi[i] == wi and gd[i] == wd
    ctx.close()


def test_back_to_back_scans_overlap_without_changing_results():
    """Scans enqueued back to back for one context may overlap (programmatic dependent launch: the next scan streams
    while the previous one's last CTAs finish; both share the context's scratch and tickets).  Every launch of such a
    stream must return what a lone launch returns - also when a tree append sits between two launches (the chain is
    then broken and the scan is ordered behind the append), and when two contexts alternate on one stream."""
    import torch

    sc, _ = scenes.make_scene(10, seed=5)
    rng = np.random.RandomState(5)
    N = 400_000
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctxs = [api.Context(sc, stream=stream.cuda_stream) for _ in range(2)]
        trees = [api.NearestNeighbors(c) for c in ctxs]
        for j, t in enumerate(trees):
            t.fill_random(N, seed=50 + j, n_slices=8)
        states = [t.read_states(0, N) for t in trees]
        for rep in range(3):
            q = scenes.sample_state(rng, sc)
            want = [ob.nn_linear(sc, states[j], q) for j in range(2)]
            for j, t in enumerate(trees):
                t.nearest_upload(q[None, :])
            for i in range(24):  # alternating contexts, every launch overlapping the previous one
                trees[i % 2].nearest_launch()
                if i % 5 == 4:
                    gi, gd = trees[i % 2].nearest_fetch()
                    assert (int(gi[0]), float(gd[0])) == want[i % 2]
            for j, t in enumerate(trees):
                gi, gd = t.nearest_fetch()
                assert (int(gi[0]), float(gd[0])) == want[j]
            # a node exactly at the query, appended between two launches of the stream: the next launch must see it
            t = trees[0]
            t.nearest_launch()
            first = t.add_batch(q[None, :])
            t.nearest_launch()
            t.nearest_launch()
            gi, gd = t.nearest_fetch()
            assert int(gi[0]) == N + rep and float(gd[0]) == 0.0
            assert t.nearest(q) == (N + rep, 0.0)
            states[0] = np.concatenate([states[0], q[None, :]])
            del first
        for c in ctxs:
            c.close()


def _all_distances(sc, states, q, weights=None, mask=None):
    """orc_distance (oracle/mps_oracle.c) for every node, vectorised in numpy with the same operations in the same
    order (double; body order) - checked against the oracle itself on a sample below"""
    B = sc.n_bodies
    full = (1 << B) - 1
    mask = full if mask is None else mask
    st = states.astype(np.float64)
    qq = q.astype(np.float64)
    dist = np.zeros(len(states))
    for i in range(B):
        if (mask >> i) & 1:
            dx = st[:, 3 * i] - qq[3 * i]
            dy = st[:, 3 * i + 1] - qq[3 * i + 1]
            dpos = np.sqrt(dx * dx + dy * dy)
            d = np.abs(st[:, 3 * i + 2] - qq[3 * i + 2])
            arc = np.where(d > np.pi, 2.0 * np.pi - d, d)
            cfg = sc.position_weight * dpos + sc.orientation_weight * arc
            w = 1.0 if weights is None else float(weights[i])
            dist = dist + w * cfg
    return dist


def _sorted_scan(d, k, keep=None):
    idx = np.arange(len(d))
    if keep is not None:
        d, idx = d[keep], idx[keep]
    order = np.lexsort((idx, d))[:k]
    return idx[order], d[order]


@pytest.mark.parametrize("n_objects,N", [(3, 7), (6, 5000), (10, 200_000)])
def test_nearest_k_equals_sorted_double_scan(n_objects, N):
    """mps_tree_nearest_k (ompl nearestK): ascending (double distance, index), bit-exact"""
    sc, _ = scenes.make_scene(n_objects, seed=N)
    B = sc.n_bodies
    rng = np.random.RandomState(N)
    states = np.empty((N, 3 * B), np.float32)
    states[:, 0::3] = rng.uniform(sc.x_min, sc.x_max, size=(N, B))
    states[:, 1::3] = rng.uniform(sc.y_min, sc.y_max, size=(N, B))
    states[:, 2::3] = rng.uniform(-np.pi, np.pi, size=(N, B))
    slices = rng.randint(0, 5, size=N).astype(np.int32)
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.add_batch(states, slice_ids=slices)
    weights = rng.uniform(0.5, 2.0, size=B).astype(np.float32)
    for qi in range(4):
        q = scenes.sample_state(rng, sc)
        for mask in _masks(B):
            w = weights if qi % 2 else None
            d = _all_distances(sc, states, q, weights=w, mask=mask)
            for j in rng.randint(0, N, size=3):   # the vectorised reference is the oracle's arithmetic
                assert d[j] == ob.distance(sc, states[j], q, weights=w, mask=mask)
            for k in (1, 4, 32):
                gi, gd = tree.nearest_k(q, k, weights=w, mask=mask)
                wi, wd = _sorted_scan(d, k)
                assert np.array_equal(gi, wi) and np.array_equal(gd, wd), (qi, mask, k)
            gi, gd = tree.nearest_k(q, 8, weights=w, mask=mask, slice_filter=3)
            wi, wd = _sorted_scan(d, 8, keep=slices == 3)
            assert np.array_equal(gi, wi) and np.array_equal(gd, wd)
            n1, d1 = tree.nearest(q, weights=w, mask=mask)
            gi, gd = tree.nearest_k(q, 1, weights=w, mask=mask)
            assert (gi[0], gd[0]) == (n1, d1)
    ctx.close()


def test_nearest_k_on_an_empty_tree_and_a_slice_without_members():
    """nearestK with nothing to return: an empty tree, a slice filter no node carries; then fewer nodes than k."""
    sc, _ = scenes.make_scene(3, seed=5)
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.clear()
    q = np.zeros(3 * sc.n_bodies, np.float32)
    gi, gd = tree.nearest_k(q, 4)
    assert len(gi) == 0 and len(gd) == 0
    tree.add(q, slice_id=3)
    gi, gd = tree.nearest_k(q, 4, slice_filter=4)
    assert len(gi) == 0
    gi, gd = tree.nearest_k(q, 4, slice_filter=3)
    assert list(gi) == [0] and gd[0] == 0.0
    q2 = q.copy()
    q2[0] = 0.25
    tree.add(q2, slice_id=3)
    gi, gd = tree.nearest_k(q, 32)
    assert list(gi) == [0, 1] and gd[0] == 0.0 and gd[1] == ob.distance(sc, q2, q)
    assert tree.nearest_k_fallbacks() == 0
    ctx.close()


def test_nearest_k_at_one_million_nodes_and_with_duplicates():
    sc, _ = scenes.make_scene(10, seed=1003)
    B = sc.n_bodies
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    N = 1_000_000
    tree.fill_random(N, seed=1003, n_slices=64)
    states = tree.read_states(0, N)
    rng = np.random.RandomState(5)
    for _ in range(3):
        q = scenes.sample_state(rng, sc)
        d = _all_distances(sc, states, q)
        gi, gd = tree.nearest_k(q, 32)
        wi, wd = _sorted_scan(d, 32)
        assert np.array_equal(gi, wi) and np.array_equal(gd, wd)
    assert tree.nearest_k_fallbacks() == 0
    # duplicates crowd one warp's list: the launch flags it and the successor scans answer exactly
    small = states[:3000].copy()
    q = scenes.sample_state(rng, sc)
    d = _all_distances(sc, small, q)
    j = int(np.argmin(d))
    crowded = np.concatenate([small[:64], np.tile(small[j], (40, 1)), small[64:]])
    tree.clear()
    tree.add_batch(crowded)
    gi, gd = tree.nearest_k(q, 6)
    wi, wd = _sorted_scan(_all_distances(sc, crowded, q), 6)
    assert np.array_equal(gi, wi) and np.array_equal(gd, wd)
    assert tree.nearest_k_fallbacks() == 1
    gi, gd = tree.nearest_k(q, 32)   # more results asked for than the duplicates: still exact
    wi, wd = _sorted_scan(_all_distances(sc, crowded, q), 32)
    assert np.array_equal(gi, wi) and np.array_equal(gd, wd)
    ctx.close()


def test_two_level_query_equals_two_separate_queries():
    """mps_tree_nearest_two_level: nearest slice representative, then the nearest member of that slice, chained on the
    device (SliceBasedOracleRRT::selectTreeNode RRT.cpp:783-827 without the host round trip)"""
    sc, _ = scenes.make_scene(6, seed=9)
    B = sc.n_bodies
    rng = np.random.RandomState(9)
    n_slices, N = 40, 6000
    reps = np.stack([scenes.sample_state(rng, sc) for _ in range(n_slices)])
    states = np.stack([scenes.sample_state(rng, sc) for _ in range(N)])
    slices = rng.randint(0, n_slices - 1, size=N).astype(np.int32)   # the last slice has no member
    ctx, rctx = api.Context(sc), api.Context(sc)
    tree, rtree = api.NearestNeighbors(ctx), api.NearestNeighbors(rctx)
    tree.add_batch(states, slice_ids=slices)
    rtree.add_batch(reps)
    arr_mask, robot_mask = ((1 << B) - 1) & ~1, 1
    for _ in range(8):
        q = scenes.sample_state(rng, sc)
        (ri, rd), (i, d) = tree.nearest_two_level(rtree, q, arr_mask, robot_mask)
        wri, wrd = rtree.nearest(q, mask=arr_mask)
        wi, wd = tree.nearest(q, mask=robot_mask, slice_filter=wri)
        assert (ri, rd, i, d) == (wri, wrd, wi, wd)
    q = reps[n_slices - 1]   # the empty slice is the nearest: no member
    (ri, rd), (i, d) = tree.nearest_two_level(rtree, q, arr_mask, robot_mask)
    assert ri == n_slices - 1 and rd == 0.0 and i == -1
    ctx.close()
    rctx.close()


def test_nearest_with_near_duplicates_at_the_angle_wrap():
    """under a one-body mask the whole distance can be tiny while the fp32 arc of a wrapped angle difference,
    2 pi - |dtheta|, carries an absolute error of a few 1e-7 rad; the candidate slack has to cover it, or the true
    double-precision minimum is filtered out.  Nodes that differ from the query only by angles just across the wrap."""
    sc, _ = scenes.make_scene(3, seed=4)
    B = sc.n_bodies
    rng = np.random.RandomState(3)
    N = 4096
    states = np.stack([scenes.sample_state(rng, sc) for _ in range(N)])
    q = scenes.sample_state(rng, sc)
    q[2] = np.float32(np.pi - 3e-5)
    for i in range(0, N, 2):                       # robot sub-state: same position, angle just beyond -pi
        states[i, 0:2] = q[0:2]
        states[i, 2] = np.float32(-np.pi + 3e-5 + 2e-7 * (i % 97))
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.add_batch(states)
    for mask in (1, (1 << B) - 1):
        gi, gd = tree.nearest(q, mask=mask)
        wi, wd = ob.nn_linear(sc, states, q, mask=mask)
        assert (gi, gd) == (wi, wd)
        ki, kd = tree.nearest_k(q, 16, mask=mask)
        d = _all_distances(sc, states, q, mask=mask)
        wi16, wd16 = _sorted_scan(d, 16)
        assert np.array_equal(ki, wi16) and np.array_equal(kd, wd16)
        if mask == 1:
            assert gd < 1e-4   # the distances in play really are tiny
    ctx.close()
