"""Generates tests/golden/physics_golden_v3.npz: rollouts of the physics spec mps-planar-v3 produced by the
CPU oracle on seeded inputs.  The reference's physics (sim_env / Box2D) is not in /root/reference, so
these vectors pin OUR spec (regression + CPU/GPU agreement), not the reference: "parity unpinned".
They are a REGRESSION file (the oracle wrote them); the independent pin of the spec is physics_kat_v3.json
(make_physics_kat.py, from the numpy restatement spec_np.py).  One file per spec version: an existing file is
never regenerated - a spec change gets a new version and a new file.

Run: python tests/golden/make_physics_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle_binding as ob  # noqa: E402
from manipulation_planning_suite_b200 import scenes  # noqa: E402


def cases():
    for n_objects, K, L, seed in [(3, 24, 1, 101), (6, 32, 2, 102), (10, 16, 3, 103), (16, 12, 1, 104)]:
        sc, _ = scenes.make_scene(n_objects, seed=seed)
        start = scenes.cluttered_start(sc, seed)
        rng = np.random.RandomState(seed)
        controls = scenes.sample_controls_towards(rng, sc, start, K, L)
        dest = scenes.sample_state(rng, sc)
        yield n_objects, seed, sc, start, controls, dest


def main():
    out = {}
    for i, (n_objects, seed, sc, start, controls, dest) in enumerate(cases()):
        r = ob.extend(sc, start, controls, dest=dest, compare_f32=True)
        out[f"c{i}_meta"] = np.array([n_objects, seed, r["k"], r["n_ok"]], np.int64)
        out[f"c{i}_start"] = start
        out[f"c{i}_controls"] = controls
        out[f"c{i}_dest"] = dest
        for k in ("all_states", "all_rest", "all_flags", "all_dist", "all_steps", "best_states"):
            out[f"c{i}_{k}"] = r[k]
    path = os.path.join(HERE, "physics_golden_v3.npz")
    if os.path.exists(path):
        raise SystemExit(path + " exists: a committed spec version is never regenerated")
    np.savez_compressed(path, **out)
    print("wrote", path, sum(v.nbytes for v in out.values()), "bytes raw")


if __name__ == "__main__":
    main()
