"""Solution files in the reference's format (OraclePushPlanner::saveSolution / loadSolution,
src/mps/planner/pushing/OraclePushPlanner.cpp:260-385):

    line 1   <#objects>,<has_velocities>,<dofs obj 0>,<dofs obj 1>,...
    line 2   <#control parameters>                         (5 for the ramp control incl. rest time)
    line 3   PlanningStatistics::printCVS                  (algorithm/Interfaces.h:73-84)
    line 4   <#motions>
    then per motion: one line with the state (x, y, theta per object, ", "-separated) and one with the
    control parameters (vx, vy, omega, t_plateau, t_rest).

The reference prints floats with the stream's default precision (6 significant digits, Eigen::StreamPrecision in
util/Serialize.cpp:13); this writer uses 9 so that a reloaded path replays bit-identically — the reference's
reader (std::stof per comma-separated token) accepts both.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

STAT_FIELDS = ["num_iterations", "num_state_propagations", "num_samples", "num_nearest_neighbor_queries", "runtime",
               "success", "reproducible", "cost_before_shortcut", "cost_after_shortcut", "reproducible_after_shortcut"]


@dataclass
class Solution:
    states: np.ndarray       # [n, 3B]
    controls: np.ndarray     # [n, 5]
    stats: dict = field(default_factory=dict)


def _fmt(v) -> str:
    return ", ".join(f"{float(x):.9g}" for x in v)


def save_solution(filename: str, states, controls, stats: dict | None = None) -> None:
    states = np.asarray(states, np.float32)
    controls = np.asarray(controls, np.float32)
    n, dim = states.shape
    assert dim % 3 == 0 and controls.shape == (n, 5)
    B = dim // 3
    st = {k: 0 for k in STAT_FIELDS}
    st.update(stats or {})
    with open(filename, "w") as f:
        f.write(",".join([str(B), "0"] + ["3"] * B) + "\n")   # position-only states: has_velocities = 0
        f.write("5\n")
        f.write(", ".join(str(int(st[k])) if k not in ("runtime", "cost_before_shortcut", "cost_after_shortcut")
                          else f"{float(st[k]):g}" for k in STAT_FIELDS) + "\n")
        f.write(f"{n}\n")
        for i in range(n):
            f.write(_fmt(states[i]) + "\n")
            f.write(_fmt(controls[i]) + "\n")


def load_solution(filename: str, n_objects: int | None = None) -> Solution:
    with open(filename) as f:
        lines = [ln.rstrip("\n") for ln in f]
    head = [x.strip() for x in lines[0].split(",")]
    if len(head) < 3:
        raise ValueError("Could not load solution. Invalid header!")
    B = int(head[0])
    if n_objects is not None and B != n_objects:
        raise ValueError("Could not load solution. Invalid number of objects in state space.")
    if int(head[1]) != 0:
        raise ValueError("Could not load solution. Semi-dynamic flag is incompatible.")
    if len(head) != B + 2:
        raise ValueError("Could not load solution. Incorrect number of object DoFs in header")
    if any(int(d) != 3 for d in head[2:]):
        raise ValueError("Could not load solution. Dimension for an object is incompatible.")
    if int(lines[1]) != 5:
        raise ValueError("Could not load solution. Action space is incompatible")
    sv = [x.strip() for x in lines[2].split(",")]
    if len(sv) != 10:
        raise ValueError("Could not read stats.")
    stats = {k: (float(v) if k in ("runtime", "cost_before_shortcut", "cost_after_shortcut") else int(v))
             for k, v in zip(STAT_FIELDS, sv)}
    n = int(lines[3])
    states = np.zeros((n, 3 * B), np.float32)
    controls = np.zeros((n, 5), np.float32)
    for i in range(n):
        states[i] = [np.float32(x) for x in lines[4 + 2 * i].split(",")]
        controls[i] = [np.float32(x) for x in lines[5 + 2 * i].split(",")]
    return Solution(states, controls, stats)
