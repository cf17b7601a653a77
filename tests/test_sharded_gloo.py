"""world_size-2 (and 3) gloo test of the multi-GPU extend's host logic: every rank propagates its slice
(here with the CPU oracle standing in for the device), the winner records are all-gathered and merged;
the result must equal the single-process extend over all K chains."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_binding as ob
    from manipulation_planning_suite_b200 import scenes, sharded

    sc, _ = scenes.make_scene(6, seed=7)
    start = scenes.cluttered_start(sc, 17)
    rng = np.random.RandomState(11)
    K, L = 40, 3
    controls = scenes.sample_controls_towards(rng, sc, start, K, L)
    n_ctrl = rng.randint(1, L + 1, size=K).astype(np.int32)
    dest = scenes.sample_state(rng, sc)
    k0, k1 = sharded.shard_range(K, rank, world)
    r = ob.extend(sc, start, controls[k0:k1], n_ctrl=n_ctrl[k0:k1], dest=dest, compare_f32=True, k_offset=k0)
    rec = sharded.Record(r["k"], r["n_ok"], r["goal_step"], -1, r["distance"], r["best_states"], r["best_controls"],
                         r["best_flags"])
    local = torch.from_numpy(sharded.pack_record(rec, sc.n_bodies, L))
    gathered = sharded.all_gather_records(local).numpy()
    recs = [sharded.unpack_record(gathered[i], sc.n_bodies, L) for i in range(world)]
    w = sharded.merge_records(recs, has_dest=True, compare_f32=True)
    full = ob.extend(sc, start, controls, n_ctrl=n_ctrl, dest=dest, compare_f32=True)
    ok = (recs[w].k == full["k"] and recs[w].distance == full["distance"]
          and np.array_equal(recs[w].states, full["best_states"]) and recs[w].n_ok == full["n_ok"])
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_extend_merge_equals_single_process(world):
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    port = 29000 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret[r] for r in range(world)), dict(ret)
