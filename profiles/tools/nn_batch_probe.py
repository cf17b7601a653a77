import sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
from manipulation_planning_suite_b200 import api, scenes
sc, _ = scenes.make_scene(bench.NN_CFG["n_objects"], seed=bench.NN_CFG["seed"])
B = sc.n_bodies
stream = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.cuda.stream(stream):
    ctx = api.Context(sc, stream=stream.cuda_stream)
    tree = api.NearestNeighbors(ctx)
    N = bench.NN_CFG["nodes"]
    tree.fill_random(N, seed=bench.NN_CFG["seed"], n_slices=64)
    rng = np.random.RandomState(11)
    for Q in (1, 2, 4, 8, 16, 32, 64, 128):
        qs = np.stack([scenes.sample_state(rng, sc) for _ in range(Q)])
        tree.nearest_upload(qs)
        tree.nearest_launch(); torch.cuda.synchronize()
        ts = []
        for rep in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); tree.nearest_launch(); e1.record(stream); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        print(f"Q={Q}: {ms*1e3:.1f} us per launch, {ms/Q*1e3:.2f} us per query, {Q/ms*1e3:.0f} queries/s", flush=True)
