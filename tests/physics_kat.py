"""Reads tests/golden/physics_kat_v3.json (written by tests/golden/make_physics_kat.py from the independent numpy
restatement of the physics spec, tests/golden/spec_np.py) and rebuilds the scenes its cases are stated on.
TEST INFRASTRUCTURE: shared by the CPU test of the oracle and the GPU parity test."""
from __future__ import annotations

import json
import os

import numpy as np

from manipulation_planning_suite_b200 import abi, scenes

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "physics_kat_v3.json")


def load():
    with open(PATH) as f:
        return json.load(f)


def build_scene(spec: dict) -> abi.Scene:
    """spec: {"base": {"n_objects", "seed", "walls", "disc_every"}, "set": {field: value | {index: value}}}"""
    b = spec["base"]
    sc, _ = scenes.make_scene(b["n_objects"], seed=b["seed"], walls=b["walls"], disc_every=b["disc_every"])
    for name, val in spec.get("set", {}).items():
        if isinstance(val, dict):
            arr = getattr(sc, name)
            for i, v in val.items():
                arr[int(i)] = v
        else:
            setattr(sc, name, val)
    return sc


def arr(x):
    return np.array(x, np.float32)
