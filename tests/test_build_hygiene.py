"""Build hygiene of libmps_b200.so, checked on the CPU box with cuobjdump: the library carries sm_100a code only,
and no kernel of it uses a stack frame or local memory (a spill or a per-thread copy of the parameter block in a
hot kernel is a silent slowdown: the NN scan once lost 1.3 us per launch to one)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "manipulation_planning_suite_b200", "csrc", "libmps_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not os.path.exists(LIB) or not os.path.exists(CUOBJDUMP),
                                reason="needs the built library and cuobjdump")


def _resource_usage():
    out = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function\s+(\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+)\s+STACK:(\d+)\s+SHARED:(\d+)\s+LOCAL:(\d+)", line)
        if m and name:
            usage[name] = dict(reg=int(m.group(1)), stack=int(m.group(2)), shared=int(m.group(3)), local=int(m.group(4)))
            name = None
    return usage


def test_only_sm_100a_code():
    out = subprocess.run([CUOBJDUMP, "-lelf", LIB], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_kernel_uses_stack_or_local_memory():
    usage = _resource_usage()
    kernels = {k: v for k, v in usage.items() if "kernel" in k}
    assert len(kernels) >= 15, sorted(kernels)
    bad = {k: v for k, v in kernels.items() if v["stack"] != 0 or v["local"] != 0}
    assert not bad, bad


def test_register_budgets_of_the_hot_kernels():
    usage = _resource_usage()
    rollout = {k: v["reg"] for k, v in usage.items() if "rollout_kernel" in k}
    assert len(rollout) == 6, rollout
    for k, reg in rollout.items():
        # <4, 1, G>: one CTA per SM -> up to 255 registers; <4, 3, G>: three CTAs of 128 threads per SM -> at most
        # 168; <4, 4, G>: four -> at most 128
        limit = 255 if "ILi4ELi1E" in k else 168 if "ILi4ELi3E" in k else 128
        assert reg <= limit, (k, reg)
    scan = [v["reg"] for k, v in usage.items() if "nn_scan_kernel" in k]
    assert len(scan) == 2 and max(scan) <= 128, scan  # two CTAs of 256 threads per SM
