"""Multi-GPU extend: K rollouts sharded over the ranks of one node, one exchange step.

SURVEY.md §8(e): the iterations of NaiveControlSampler::sampleTo (ompl/control/NaiveControlSampler.cpp:46-64)
and of HybridActionRRT::extend (pushing/algorithm/RRT.cpp:458-477) are independent, so rank r propagates the
contiguous slice [k0, k1) of the K chains with k_offset = k0.  The only data-path collective is one
all-gather of the fixed-size winner record per rank (a few hundred bytes over NVLink); every rank then
applies the same strict-< / lowest-global-index rule, so the result is independent of the GPU count.

Works with torch.distributed backend "nccl" (records stay on the device until after the gather) and
"gloo" (CPU tensors; used by the CPU tests of the merge logic).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import abi


def shard_range(K: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice of the K chains owned by `rank`."""
    return K * rank // world, K * (rank + 1) // world


@dataclass
class Record:
    k: int
    n_ok: int
    goal_step: int
    tree_index: int
    distance: float
    states: np.ndarray    # [L, 3B]
    controls: np.ndarray  # [L, 5]
    flags: np.ndarray     # [L]


def record_size(B: int, L: int) -> int:
    n = C.sizeof(abi.ExtendBest) + 4 * L * (3 * B + abi.MPS_CONTROL_DIM) + 4 * L
    return (n + 7) // 8 * 8


def pack_record(rec: Record, B: int, L: int) -> np.ndarray:
    """Same byte layout as the device record (include/mps_b200.h, mps_extend_record_to)."""
    buf = np.zeros(record_size(B, L), np.uint8)
    head = abi.ExtendBest(rec.k, rec.n_ok, rec.goal_step, rec.tree_index, rec.distance)
    hb = C.sizeof(abi.ExtendBest)
    buf[:hb] = np.frombuffer(bytes(head), np.uint8)
    o = hb
    for arr, dt, n in ((rec.states, np.float32, L * 3 * B), (rec.controls, np.float32, L * 5), (rec.flags, np.uint32, L)):
        a = np.ascontiguousarray(arr, dt).reshape(-1)
        assert a.size == n
        buf[o:o + 4 * n] = a.view(np.uint8)
        o += 4 * n
    return buf


def unpack_record(buf: np.ndarray, B: int, L: int) -> Record:
    buf = np.ascontiguousarray(buf, np.uint8)
    hb = C.sizeof(abi.ExtendBest)
    head = abi.ExtendBest.from_buffer_copy(buf[:hb].tobytes())
    o = hb
    states = buf[o:o + 4 * L * 3 * B].view(np.float32).reshape(L, 3 * B).copy()
    o += 4 * L * 3 * B
    controls = buf[o:o + 4 * L * 5].view(np.float32).reshape(L, 5).copy()
    o += 4 * L * 5
    flags = buf[o:o + 4 * L].view(np.uint32).copy()
    return Record(head.k, head.n_ok, head.goal_step, head.tree_index, head.distance, states, controls, flags)


def merge_records(records: list[Record], has_dest: bool = True, compare_f32: bool = False) -> int:
    """Index (into `records`, i.e. the rank) of the global winner, or -1.

    Ranks hold ascending contiguous slices, so walking them in rank order with a strict `<` reproduces
    the sequential loop of the reference: first index wins ties (NaiveControlSampler.cpp:55, RRT.cpp:467).
    """
    best = -1
    best_key = None
    for r, rec in enumerate(records):
        if rec.k < 0:
            continue
        if not has_dest:
            return r
        key = float(np.float32(rec.distance)) if compare_f32 else float(rec.distance)
        if best < 0 or key < best_key:
            best, best_key = r, key
    return best


def all_gather_records(local: "torch.Tensor", group=None) -> "torch.Tensor":  # noqa: F821
    """One all-gather of the per-rank record bytes (uint8 [rec_bytes]) -> [world, rec_bytes]."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = torch.empty((world, local.numel()), dtype=torch.uint8, device=local.device)
    if dist.get_backend(group) == "gloo":
        chunks = [out[i] for i in range(world)]
        dist.all_gather(chunks, local.contiguous(), group=group)
    else:
        dist.all_gather_into_tensor(out.view(-1), local.contiguous(), group=group)
    return out


class ShardedExtend:
    """One extend over all ranks: local rollouts on this rank's device, ONE all-gather of the packed winner
    records, and the merge + tree append on the device (mps_extend_merge_records).  Only the 24-byte head of the
    winner returns to the host.

    Stream order: the rollout, the record copy, the all-gather and the merge are all enqueued on the ctx's own
    stream (torch sees it as an ExternalStream; NCCL orders its work against the stream that is current when the
    collective is called), so nothing reads a record before the select kernel has written it - whichever stream
    the ctx was created on.

    `owner` = the rank whose ctx holds the tree (north_star: the single-tree NN scan stays on one device); only
    that rank appends the winner chain.
    """

    def __init__(self, ctx, group=None, owner: int = 0):
        import torch
        import torch.distributed as dist

        self.ctx = ctx
        self.group = group
        self.dist = dist
        self.torch = torch
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.owner = owner
        self.device = torch.device("cuda", ctx.device)
        self.stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=self.device)
        self._buf = None
        self._gathered = None
        self._empty = False

    def upload(self, start, controls_all, **kw):
        """controls_all: the full [K][L][5] batch (same on every rank); this rank keeps its slice.
        append_to_tree / parent: the winner of the WHOLE extend is appended on the owner rank (never the local one)."""
        controls_all = np.asarray(controls_all, np.float32)
        if controls_all.ndim == 2:
            controls_all = controls_all[:, None, :]
        K = controls_all.shape[0]
        k0, k1 = shard_range(K, self.rank, self.world)
        n_ctrl = kw.pop("n_ctrl", None)
        if n_ctrl is not None:
            n_ctrl = np.asarray(n_ctrl, np.int32)
        self.append = bool(kw.pop("append_to_tree", False))
        self.parent = int(kw.pop("parent", -1))
        self.has_dest = kw.get("dest") is not None
        self.compare_f32 = bool(kw.get("compare_f32", False))
        self.L = controls_all.shape[1]
        # a rank whose slice is empty (K < world) still takes part in the gather with a k = -1 record; the ctx is
        # given chain 0 so that its record buffer has the shape of this extend, but nothing is launched
        self._empty = k1 <= k0
        if self._empty:
            k0, k1 = 0, 1
        self.ctx.extend_upload(start, controls_all[k0:k1], n_ctrl=None if n_ctrl is None else n_ctrl[k0:k1], k_offset=k0, **kw)
        self._ensure_buffers()

    def _ensure_buffers(self):
        n = self.ctx.record_size()
        if self._buf is None or self._buf.numel() != n:
            self._buf = self.torch.empty(n, dtype=self.torch.uint8, device=self.device)
            self._gbuf = self.torch.empty((self.world, n), dtype=self.torch.uint8, device=self.device)
            none = Record(-1, 0, -1, -1, float("inf"), np.zeros((self.L, 3 * self.ctx.B), np.float32),
                          np.zeros((self.L, 5), np.float32), np.zeros(self.L, np.uint32))
            self._none = self.torch.from_numpy(pack_record(none, self.ctx.B, self.L)).to(self.device)

    def upload_local(self, start, controls_local, k_offset: int, **kw):
        """This rank's slice only (each rank generated / received just its own chains): controls_local [k][L][5] are
        the chains [k_offset, k_offset + k) of the extend."""
        controls_local = np.asarray(controls_local, np.float32)
        if controls_local.ndim == 2:
            controls_local = controls_local[:, None, :]
        self.append = bool(kw.pop("append_to_tree", False))
        self.parent = int(kw.pop("parent", -1))
        self.has_dest = kw.get("dest") is not None
        self.compare_f32 = bool(kw.get("compare_f32", False))
        self.L = controls_local.shape[1]
        self._empty = False
        self.ctx.extend_upload(start, controls_local, k_offset=int(k_offset), **kw)
        self._ensure_buffers()

    def launch(self):
        """Enqueue rollouts + select + record copy + all-gather + merge (+ append on the owner) on the ctx stream;
        no host synchronisation."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            if self._empty:
                self._buf.copy_(self._none, non_blocking=True)
            else:
                self.ctx.extend_launch()
                self.ctx.record_to(self._buf.data_ptr())
            if self.world > 1:
                self.dist.all_gather_into_tensor(self._gbuf.view(-1), self._buf, group=self.group)
                gathered = self._gbuf
            else:
                gathered = self._buf.view(1, -1)
            self._gathered = gathered
            self.ctx.merge_records(gathered.data_ptr(), self.world, self.has_dest, self.compare_f32,
                                   append_to_tree=self.append and self.rank == self.owner, parent=self.parent)
        return gathered

    def head(self) -> abi.ExtendBest:
        """The 24-byte head of the merged winner (k is the global chain index; tree_index is set on the owner)."""
        return self.ctx.extend_head()

    def result(self) -> Record | None:
        """Head + winner chain of the merged record (device -> host copy of one record)."""
        K, L = self.ctx._ext_shape
        r = self.ctx.extend_fetch()
        if r.k < 0:
            return None
        return Record(r.k, r.n_ok, r.goal_step, r.tree_index, r.distance, r.best_states, r.best_controls, r.best_flags)
