"""NN scan (cfg3 shape: B = 11, 1 M nodes) with a given build of libmps_b200.so: three trees scanned in rotation
(every launch reads its 132 MB from HBM), timed (a) one CUDA event pair per launch, (b) one pair around the whole
stream of launches (the kernel's average launch duration when queries are enqueued back to back), and the C loop of
single host-to-host queries.  usage: nn_pipeline_probe.py [lib.so]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import abi  # noqa: E402

if len(sys.argv) > 1:
    abi._lib = abi.load_library(sys.argv[1])
from manipulation_planning_suite_b200 import api, scenes  # noqa: E402

name = os.path.basename(sys.argv[1]) if len(sys.argv) > 1 else "default"
for n_objects, N in ((10, 1_000_000), (10, 4_000_000), (6, 100_000)):
    sc, _ = scenes.make_scene(n_objects, seed=1003)
    B = sc.n_bodies
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        trees = []
        rng = np.random.RandomState(11)
        q = scenes.sample_state(rng, sc)
        for j in range(3 if N <= 1_000_000 else 1):
            ctx = api.Context(sc, stream=stream.cuda_stream)
            t = api.NearestNeighbors(ctx)
            t.fill_random(N, seed=1003 + j, n_slices=64)
            t.nearest_upload(q[None, :])
            t.nearest_launch()
            trees.append(t)
        torch.cuda.synchronize()
        nt = len(trees)
        iters = 48
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for i in range(iters):
            ev[i][0].record(stream)
            trees[i % nt].nearest_launch()
            ev[i][1].record(stream)
        torch.cuda.synchronize()
        per = float(np.mean([a.elapsed_time(b) for a, b in ev[6:]])) * 1e3
        best = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            trees[0].nearest_launch()
            e0.record(stream)
            for i in range(iters):
                trees[(i + 1) % nt].nearest_launch()
            e1.record(stream)
            torch.cuda.synchronize()
            best.append(e0.elapsed_time(e1) / iters * 1e3)
        pipe = float(np.median(best))
        r0 = trees[0].nearest_fetch()
        qs = np.stack([scenes.sample_state(rng, sc) for _ in range(200)])
        idx, dist, sec = trees[0].nearest_timed(qs)
        # the pipelined launches must return what a lone launch returns
        trees[0].nearest_upload(qs[:1])
        trees[0].nearest_launch()
        a = trees[0].nearest_fetch()
        for _ in range(8):
            trees[0].nearest_launch()
        b = trees[0].nearest_fetch()
        same = (a[0][0] == b[0][0]) and (a[1][0] == b[1][0]) and idx[0] == a[0][0]
    gb = N * 12 * B / 1e3
    print(f"{name:24s} B={B} N={N}: per-launch {per:6.2f} us ({gb / per / 6554.2:.3f})  stream of launches {pipe:6.2f} us "
          f"({gb / pipe / 6554.2:.3f})  e2e C loop {sec / 200 * 1e6:6.2f} us  consistent={same}", flush=True)
    for t in trees:
        t.ctx.close()
