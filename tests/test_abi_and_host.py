"""CPU tests of the boundary itself: the library loads without a GPU, exports every symbol the header
declares, and fails loudly (no fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from manipulation_planning_suite_b200 import abi, api, scenes, sharded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mps_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mps_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"libmps_b200.so does not export {n}"
    # and the ctypes table covers the header
    assert set(names) == set(abi.SYMBOLS.keys())


def test_struct_sizes_match_the_header(tmp_path):
    """sizeof / offsetof as the C compiler sees include/mps_b200.h vs the ctypes mirror."""
    import subprocess

    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "mps_b200.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(mps_hybrid_gen), offsetof(mps_hybrid_gen, object_size), sizeof(mps_scene), sizeof(mps_goal),"
        " sizeof(mps_extend_in), sizeof(mps_extend_best), sizeof(mps_extend_out), offsetof(mps_scene, position_weight),"
        " offsetof(mps_scene, dt), offsetof(mps_scene, jam_steps), sizeof(mps_forest_extend_in), sizeof(mps_forest_naive_in),"
        " offsetof(mps_forest_naive_in, goals), offsetof(mps_forest_naive_in, duration_max), offsetof(mps_extend_in, starts));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(abi.HybridGen), abi.HybridGen.object_size.offset, C.sizeof(abi.Scene), C.sizeof(abi.Goal), C.sizeof(abi.ExtendIn), C.sizeof(abi.ExtendBest),
            C.sizeof(abi.ExtendOut), abi.Scene.position_weight.offset, abi.Scene.dt.offset, abi.Scene.jam_steps.offset,
            C.sizeof(abi.ForestExtendIn), C.sizeof(abi.ForestNaiveIn), abi.ForestNaiveIn.goals.offset,
            abi.ForestNaiveIn.duration_max.offset, abi.ExtendIn.starts.offset]
    assert got == want


def test_scene_defaults_follow_the_reference_defaults():
    sc = scenes.default_scene()
    assert sc.position_weight == 1.0 and sc.orientation_weight == 0.08   # SimEnvState.h:224-231
    assert sc.t_max == 8.0                                                # OraclePushPlanner.cpp:52
    assert sc.x_min < -1e30 and sc.x_max > 1e30                           # OraclePushPlanner.cpp:61-66
    assert sc.strict_active_contacts == 0                                 # SimEnvStatePropagator.cpp:79 quirk kept


def test_python_scene_defaults_equal_the_library_defaults():
    lib_sc = abi.Scene()
    abi.load_library().mps_scene_defaults(lib_sc)
    assert bytes(lib_sc) == bytes(scenes.default_scene())


def test_no_device_is_a_hard_error():
    lib = abi.load_library()
    n = C.c_int(0)
    lib.mps_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a CUDA device is present")
    sc, _ = scenes.make_scene(3, seed=0)
    with pytest.raises(api.MpsError) as e:
        api.Context(sc)
    assert e.value.code == abi.MPS_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_invalid_scene_is_rejected_before_touching_the_device():
    sc = scenes.default_scene()
    sc.n_bodies = 40
    with pytest.raises(api.MpsError) as e:
        api.Context(sc)
    assert e.value.code == abi.MPS_ERR_ARG


def test_missing_library_fails_loudly():
    with pytest.raises(abi.LibraryMissing):
        abi.load_library("/nonexistent/libmps_b200.so")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "manipulation_planning_suite_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle_binding" not in text and "mps_oracle" not in text, f


def test_record_pack_roundtrip_and_merge_rule():
    B, L = 7, 4
    rng = np.random.RandomState(0)
    recs = []
    for r in range(4):
        recs.append(sharded.Record(k=100 * r + 5, n_ok=2, goal_step=-1, tree_index=-1, distance=float(rng.rand()),
                                   states=rng.rand(L, 3 * B).astype(np.float32),
                                   controls=rng.rand(L, 5).astype(np.float32), flags=np.arange(L, dtype=np.uint32)))
    for rec in recs:
        back = sharded.unpack_record(sharded.pack_record(rec, B, L), B, L)
        assert back.k == rec.k and back.distance == rec.distance
        assert np.array_equal(back.states, rec.states) and np.array_equal(back.flags, rec.flags)
    w = sharded.merge_records(recs)
    assert w == int(np.argmin([r.distance for r in recs]))
    # ties -> lowest rank (lowest global index); invalid shards are skipped
    recs[1].distance = recs[3].distance = 0.0
    assert sharded.merge_records(recs) == 1
    recs[1].k = -1
    assert sharded.merge_records(recs) == 3
    # float compare (HybridActionRRT): two doubles that collapse to the same float tie
    recs[0].distance, recs[2].distance = 1.0 + 1e-12, 1.0
    recs[3].distance = 5.0
    assert sharded.merge_records(recs, compare_f32=True) == 0
    assert sharded.merge_records(recs, compare_f32=False) == 2
    assert sharded.shard_range(10, 0, 3) == (0, 3) and sharded.shard_range(10, 2, 3) == (6, 10)
