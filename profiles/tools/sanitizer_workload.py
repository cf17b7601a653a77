import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from manipulation_planning_suite_b200 import api, scenes
sc, _ = scenes.make_scene(6, seed=9)
B = sc.n_bodies
rng = np.random.RandomState(1)
ctx = api.Context(sc)
tree = api.NearestNeighbors(ctx)
N = 60000
st = np.stack([scenes.sample_state(rng, sc) for _ in range(2000)])
st = np.tile(st, (30, 1)) + rng.normal(scale=1e-3, size=(N, 3 * B)).astype(np.float32)
tree.add_batch(st.astype(np.float32), slice_ids=rng.randint(0, 7, size=N).astype(np.int32))
q = scenes.sample_state(rng, sc)
for k in (1, 8, 32):
    print(k, tree.nearest_k(q, k)[0][:3])
print(tree.nearest(q), tree.nearest_k(q, 5, slice_filter=3)[0])
rctx = api.Context(sc); rt = api.NearestNeighbors(rctx); rt.add_batch(st[:7])
print(tree.nearest_two_level(rt, q, ((1 << B) - 1) & ~1, 1))
start = scenes.cluttered_start(sc, 9)
g = api.Context.hybrid_gen(sc, seed=3, action_randomness=0.2, target_bias=0.3, robot_bias=0.1, targets=(1,))
ctx.extend_upload_generated(start, 128, 6, g, dest=q, compare_f32=True)
ctx.extend_launch(); r = ctx.extend_fetch(); print("gen extend", r.k, r.n_ok)
import torch
from manipulation_planning_suite_b200 import sharded
sh = sharded.ShardedExtend(ctx)
c = scenes.sample_controls_towards(rng, sc, start, 64, 2)
tree2 = api.NearestNeighbors(ctx)
sh.upload(start, c, dest=q, append_to_tree=True, parent=0)
sh.launch(); h = sh.head(); print("sharded", h.k, h.tree_index)
# packed launches: two chains per warp with the rolled sweeps (9 bodies), four per warp (7 bodies), forced per context
import os
for grp, nobj in (("16", 8), ("8", 6)):
    os.environ["MPS_ROLLOUT_GROUP"] = grp
    s2, _ = scenes.make_scene(nobj, seed=30 + nobj)
    st2 = scenes.cluttered_start(s2, 40 + nobj)
    c2 = scenes.sample_controls_towards(rng, s2, st2, 70, 2)
    cx = api.Context(s2)
    r2 = cx.extend(st2, c2, dest=scenes.sample_state(rng, s2))
    print("packed", grp, r2.k, r2.n_ok)
    cx.close()
os.environ.pop("MPS_ROLLOUT_GROUP")
