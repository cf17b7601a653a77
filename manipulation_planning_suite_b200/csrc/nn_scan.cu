// nn_scan.cu — exact weighted SE(2)^n nearest neighbour over the device-resident tree as one
// coalesced, float4-vectorised HBM scan (sm_100a), plus the tree's append / fill / gather kernels.
//
// What it replaces in the reference (paths relative to the reference root):
//   _tree->nearest(sample) on ompl::NearestNeighborsGNAT<MotionPtr>      pushing/algorithm/RRT.cpp:46-51,192-204
//   the slice-level and in-slice GNATs of SliceBasedOracleRRT             RRT.cpp:766-884, algorithm/Interfaces.cpp:52-124
//   with the metric of SimEnvWorldStateDistanceMeasure::distance          ompl/state/SimEnvWorldStateDistanceMeasure.cpp:34-51
//
// Exactness.  The reference evaluates the metric in double over float-valued states.  The scan
// evaluates it in fp32 for every node (HBM-bound: 12 bytes and ~15 flops per body per node) and keeps a
// per-CTA upper bound; only nodes whose fp32 value is within a relative 1e-4 (+1e-15 absolute) of that
// bound — far wider than the fp32 evaluation error of < 3e-6 relative — are re-evaluated with the
// reference's double expression.  The winner is the lexicographic minimum of (double distance, index),
// i.e. exactly what a sequential double-precision scan with strict `<` returns: bit-exact index and
// distance, ties to the lowest index.
#include "mps_device.cuh"
#include "nn_scan.h"

#include <cstdlib>

#define FULL 0xffffffffu
// launch shape (tuned on B200, profiles/tools/nn_tune.sh): threads per CTA, CTAs per SM, bodies whose
// 3 rows are loaded before the arithmetic starts (3 * NN_BODY_BATCH LDG.128 in flight per thread)
#ifndef NN_THREADS
#define NN_THREADS 256
#endif
#ifndef NN_CTAS_PER_SM
#define NN_CTAS_PER_SM 2
#endif
#ifndef NN_BODY_BATCH
#define NN_BODY_BATCH 6
#endif
#define NN_MIN_BLOCKS NN_CTAS_PER_SM
// Candidate filter of the fp32 pass: relative slack far above the evaluation error of a body term (< 3e-6), plus an
// absolute slack set per query (NNParams::abs_eps = 1e-15 + 1.5e-6 * ow * sum of the active weights): the arc of a
// wrapped angle difference, 2 pi - |dtheta|, carries an ABSOLUTE error of a few 1e-7 rad whatever its size, which a
// relative bound does not cover when the whole distance is tiny (near-duplicate nodes under a one-body mask).
#define NN_REL_EPS 1.0e-4f
#define NN_CAND_CAP 512

namespace {

// Programmatic dependent launch (sm_90+): a scan is launched with the programmatic-stream-serialization attribute, so
// its CTAs may become resident while the previous kernel of the stream is still draining.  grid_dep_wait() blocks
// until that kernel has completed and its writes are visible; grid_dep_trigger() lets the NEXT kernel of the stream
// start its own CTAs as soon as every CTA of this one has called it (or exited).
__device__ __forceinline__ void grid_dep_wait()
{
#ifndef MPS_CPU_EMU
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void grid_dep_trigger()
{
#ifndef MPS_CPU_EMU
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// NN_TRACE (profiles/tools/nn_trace_probe.cu): per-CTA timestamps of the scan's phases
#ifdef NN_TRACE
__device__ unsigned long long g_nn_trace[8 * 4096];
__device__ __forceinline__ void nn_trace(int slot)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g_nn_trace[8 * blockIdx.y + slot] = t;
    }
}
#else
#define nn_trace(slot)
#endif

__device__ __forceinline__ float sqrt_approx(float x)
{
#ifdef MPS_CPU_EMU
    return sqrtf(x);  // the fp32 pass only filters; every reported distance is the double expression
#else
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

__device__ __forceinline__ float body_term_f32(float x, float y, float th, float qx, float qy, float qth, float pwf,
                                               float owf)
{
    // (the library is compiled without implicit fusion: the fused multiply-adds of this filter are spelled out - it
    // only has to stay within the slack of the exact metric, and it is a fifth of the instructions of a batched scan)
    float dx = x - qx;
    float dy = y - qy;
    float d = fabsf(th - qth);
    float arc = fminf(d, MPS_TWO_PI_F - d);
    return fmaf(pwf, sqrt_approx(fmaf(dx, dx, dy * dy)), owf * arc);
}

// the reference's double expression for one node (rare path)
__device__ __noinline__ double exact_distance(const TreeView& t, long long n, const float* __restrict__ q,
                                              const float* __restrict__ weights, uint32_t mask, double pw, double ow)
{
    double dist = 0.0;
    for (int i = 0; i < t.B; ++i) {
        if ((mask >> i) & 1u) {
            float x = t.states[(size_t)(3 * i) * t.cap + n];
            float y = t.states[(size_t)(3 * i + 1) * t.cap + n];
            float th = t.states[(size_t)(3 * i + 2) * t.cap + n];
            double cfg = mps_body_distance(x, y, th, q[3 * i], q[3 * i + 1], q[3 * i + 2], pw, ow);
            double w = weights ? (double)weights[i] : 1.0;
            dist += w * cfg;
        }
    }
    return dist;
}

__device__ __forceinline__ void take_min(double& bd, long long& bi, double od, long long oi)
{
    if (oi >= 0 && (bi < 0 || od < bd || (od == bd && oi < bi))) {
        bd = od;
        bi = oi;
    }
}

// One warp evaluates the reference's double expression for one node: lane i computes body i's term
// (double sqrt is the expensive part), the terms are then added in body order so the sum has exactly the
// rounding of the sequential loop in SimEnvWorldStateDistanceMeasure.cpp:41-46.
__device__ __forceinline__ double warp_exact_distance(const TreeView& t, long long n, const float* s_q,
                                                      const float* __restrict__ weights, uint32_t mask, double pw,
                                                      double ow, int lane)
{
    double term = 0.0;
    if (lane < t.B && ((mask >> lane) & 1u)) {
        float x = t.states[(size_t)(3 * lane) * t.cap + n];
        float y = t.states[(size_t)(3 * lane + 1) * t.cap + n];
        float th = t.states[(size_t)(3 * lane + 2) * t.cap + n];
        double cfg = mps_body_distance(x, y, th, s_q[3 * lane], s_q[3 * lane + 1], s_q[3 * lane + 2], pw, ow);
        double w = weights ? (double)weights[lane] : 1.0;
        term = w * cfg;
    }
    double dist = 0.0;
    for (int i = 0; i < t.B; ++i) {
        double ti = __shfl_sync(FULL, term, i);
        if ((mask >> i) & 1u) dist += ti;
    }
    return dist;
}

// fp32 metric of the four nodes n0 .. n0 + 3 (one aligned LDG.128 per body row), +inf for nodes beyond N or outside
// the slice filter
template <bool STREAM>
__device__ __forceinline__ void group_distances(const NNParams& p, const int* s_act, const float* s_q, const float* s_w,
                                                int nact, size_t cap, size_t col0, long long n0, long long N, int filter,
                                                float pwf, float owf, float& d0, float& d1, float& d2, float& d3)
{
    const float inf = __uint_as_float(0x7f800000u);
    d0 = 0.0f;
    d1 = 0.0f;
    d2 = 0.0f;
    d3 = 0.0f;
    for (int a0 = 0; a0 < nact; a0 += NN_BODY_BATCH) {
        float4 X[NN_BODY_BATCH], Y[NN_BODY_BATCH], T[NN_BODY_BATCH];
#pragma unroll
        for (int u = 0; u < NN_BODY_BATCH; ++u) {  // all loads of the batch first: 3 * NN_BODY_BATCH LDG.128 in flight
            const int a = a0 + u < nact ? a0 + u : nact - 1;
            const float* row = p.tree.states + (size_t)(3 * s_act[a]) * cap + col0 + n0;
            X[u] = STREAM ? __ldcs(reinterpret_cast<const float4*>(row)) : __ldg(reinterpret_cast<const float4*>(row));
            Y[u] = STREAM ? __ldcs(reinterpret_cast<const float4*>(row + cap))
                          : __ldg(reinterpret_cast<const float4*>(row + cap));
            T[u] = STREAM ? __ldcs(reinterpret_cast<const float4*>(row + 2 * cap))
                          : __ldg(reinterpret_cast<const float4*>(row + 2 * cap));
        }
#pragma unroll
        for (int u = 0; u < NN_BODY_BATCH; ++u) {
            if (a0 + u < nact) {
                const int i = s_act[a0 + u];
                const float qx = s_q[3 * i], qy = s_q[3 * i + 1], qth = s_q[3 * i + 2];
                const float w = s_w[a0 + u];
                d0 += w * body_term_f32(X[u].x, Y[u].x, T[u].x, qx, qy, qth, pwf, owf);
                d1 += w * body_term_f32(X[u].y, Y[u].y, T[u].y, qx, qy, qth, pwf, owf);
                d2 += w * body_term_f32(X[u].z, Y[u].z, T[u].z, qx, qy, qth, pwf, owf);
                d3 += w * body_term_f32(X[u].w, Y[u].w, T[u].w, qx, qy, qth, pwf, owf);
            }
        }
    }
    if (filter >= 0) {
        const int4 sl = __ldg(reinterpret_cast<const int4*>(p.tree.slice + col0 + n0));
        if (sl.x != filter) d0 = inf;
        if (sl.y != filter) d1 = inf;
        if (sl.z != filter) d2 = inf;
        if (sl.w != filter) d3 = inf;
    }
    if (n0 + 3 >= N) {  // ragged tail
        if (n0 + 1 >= N) d1 = inf;
        if (n0 + 2 >= N) d2 = inf;
        if (n0 + 3 >= N) d3 = inf;
    }
}

#ifndef MPS_CPU_EMU
// STREAM: load the state rows with the evict-first hint (ld.global.cs).  Measured on B200: faster when the
// scanned rows are about the size of L2 (1 M nodes x 132 B: 29.2 us vs 32.9 us), slower when they are several
// times larger (4 M nodes: 96 us vs 88 us) - the host picks per launch (mps_launch_nn_scan).
template <bool STREAM>
__global__ void __launch_bounds__(NN_THREADS, NN_MIN_BLOCKS) nn_scan_kernel(const __grid_constant__ NNParams p)
{
    __shared__ float s_q[3 * MPS_MAX_BODIES];
    __shared__ float s_w[MPS_MAX_BODIES];   // weight of the a-th active body
    __shared__ float s_wb[MPS_MAX_BODIES];  // weight of body i (1 when the call has none)
    __shared__ int s_act[MPS_MAX_BODIES];
    __shared__ int s_nact;
    __shared__ unsigned int s_bound;
    __shared__ unsigned int s_ncand;
    __shared__ unsigned int s_cand_d[NN_CAND_CAP];   // fp32 distance bits of the deferred candidates
    __shared__ long long s_cand_i[NN_CAND_CAP];      // their node indices
    __shared__ double s_rd[NN_THREADS / 32];
    __shared__ long long s_ri[NN_THREADS / 32];
    __shared__ bool s_last;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    // grid = (queries, CTAs per query): the CTAs that scan the same part of the tree for different queries of a batch are
    // neighbours in launch order, so a part of the tree is fetched from HBM once and served to the other queries from
    // L2 (with the queries in grid.y every query streamed the whole tree on its own: 18.4 us per query at Q = 32,
    // the cost of a lone scan)
    const int q = blockIdx.x;
    const unsigned int cta = blockIdx.y, ncta = gridDim.y;
    const int B = p.tree.B;
    // Ordering against the previous kernel of the stream (see grid_dep_wait).  overlap_prev: the host knows that
    // kernel is another scan of this context that wrote nothing this one reads (no tree append, no chained filter) -
    // the streaming pass then starts while that scan's last CTAs finish, and only the hand-off below (shared
    // scratch, result block) waits for it.  Otherwise everything waits here.
    nn_trace(0);
    if (!p.overlap_prev) grid_dep_wait();
    grid_dep_trigger();
    nn_trace(2);
    // the query, its mask / slice filter and the weights come from device buffers, or (single query through
    // mps_tree_nearest) from the kernel parameters themselves
    const bool inl = p.inline_q != 0;
    uint32_t mask = inl ? p.inline_mask : (p.masks ? p.masks[q] : 0xffffffffu);
    if (B < 32) mask &= (1u << B) - 1u;
    const int filter = inl ? (p.filter_from ? (int)*p.filter_from : p.inline_filter) : (p.filters ? p.filters[q] : -1);
    // the rows of this thread's first group are requested before the prologue's barrier (they depend on nothing the
    // prologue computes): the first LDG.128 of the streaming pass then finds them on their way
    if (!p.inst_ids && (tid & 7) == 0) {
        const long long g0 = (long long)cta * NN_THREADS + tid;
        if ((g0 << 2) < p.n) {
            const float* base = p.tree.states + (g0 << 2);
            for (int i = 0; i < B; ++i) {
                if ((mask >> i) & 1u) {
#pragma unroll
                    for (int r = 0; r < 3; ++r)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)(3 * i + r) * p.tree.cap));
                }
            }
        }
    }
    nn_trace(7);
    if (tid < 3 * B) s_q[tid] = inl ? p.q_inline[tid] : p.queries[(size_t)q * 3 * B + tid];
    if (tid < B) {
        const float w = inl ? (p.inline_has_w ? p.w_inline[tid] : 1.0f) : (p.weights ? p.weights[tid] : 1.0f);
        s_wb[tid] = w;
        // active-body list, built in parallel: body i goes to position popc(mask below bit i)
        if ((mask >> tid) & 1u) {
            const int a = __popc(mask & ((1u << tid) - 1u));
            s_act[a] = tid;
            s_w[a] = w;
        }
    }
    if (tid == 0) {
        s_nact = __popc(mask);
        s_bound = 0x7f800000u;  // +inf
        s_ncand = 0u;
    }
    __syncthreads();
    nn_trace(1);
    const int nact = s_nact;
    const float pwf = (float)p.pw, owf = (float)p.ow;
    const size_t cap = p.tree.cap;
    long long N = p.n;
    size_t col0 = 0;  // first column of the scanned segment
    if (p.inst_ids) {
        const int inst = p.inst_ids[q];
        col0 = (size_t)inst * p.seg;
        N = p.seg_sizes[inst];
    }
    const long long groups = (N + 3) >> 2;
    const float inf = __uint_as_float(0x7f800000u);

    double best_d = __longlong_as_double(0x7ff0000000000000ll);
    long long best_i = -1;

    // ---- streaming pass: fp32 metric for every node, candidates deferred ------------------------
    // grid-stride: at any moment the grid reads one contiguous window of every row.  Measured on B200 and dropped: one
    // contiguous range of groups per CTA (296 x 33 separate DRAM streams: the pass 17 % slower), and dealing the
    // groups of the partial last trip to all CTAs in warp-sized units (every CTA then stays to the end with two or
    // three of its eight warps busy: 21.7 us against 20.9 us, and 24.2 us against 20.7 us per scan in a stream of
    // scans, where the CTAs that leave early make room for the next scan's).  The grid is sized so that all CTAs
    // make the same number of trips instead (mps_nn_grid_x).
    for (long long g = (long long)cta * NN_THREADS + tid; g < groups; g += (long long)ncta * NN_THREADS) {
        const long long n0 = g << 2;
        float d0, d1, d2, d3;
        group_distances<STREAM>(p, s_act, s_q, s_w, nact, cap, col0, n0, N, filter, pwf, owf, d0, d1, d2, d3);
        // Tighten the CTA bound with this warp's fp32 minimum BEFORE testing against it.
        const float dm = fminf(fminf(d0, d1), fminf(d2, d3));
        const unsigned am = __activemask();
        const unsigned wmin = __reduce_min_sync(am, __float_as_uint(dm));  // d >= 0: uint order == float order
        if (lane == (__ffs(am) - 1)) atomicMin(&s_bound, wmin);
        __syncwarp(am);
        const float b = fminf(__uint_as_float(*reinterpret_cast<volatile unsigned int*>(&s_bound)),
                              __uint_as_float(wmin));
        const float thr = b * (1.0f + NN_REL_EPS) + p.abs_eps;
        if (dm <= thr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float dj = j == 0 ? d0 : j == 1 ? d1 : j == 2 ? d2 : d3;
                if (dj <= thr && dj < inf) {
                    const unsigned slot = atomicAdd(&s_ncand, 1u);
                    if (slot < NN_CAND_CAP) {
                        s_cand_d[slot] = __float_as_uint(dj);
                        s_cand_i[slot] = n0 + j;
                    } else {
                        // list full (adversarial orderings): evaluate on the spot
                        const double e = exact_distance(p.tree, (long long)col0 + n0 + j, s_q, s_wb, mask, p.pw, p.ow);
                        take_min(best_d, best_i, e, n0 + j);
                    }
                }
            }
        }
    }
    __syncthreads();
    nn_trace(3);

    // ---- deferred exact pass: only candidates still within the final CTA bound ---------------------
    {
        const float bf = __uint_as_float(s_bound);
        const float thr = bf * (1.0f + NN_REL_EPS) + p.abs_eps;
        const int ncand = (int)min(s_ncand, (unsigned)NN_CAND_CAP);
        for (int c = warp; c < ncand; c += NN_THREADS / 32) {
            if (__uint_as_float(s_cand_d[c]) <= thr) {
                const long long n = s_cand_i[c];
                const double e = warp_exact_distance(p.tree, (long long)col0 + n, s_q, s_wb, mask, p.pw, p.ow, lane);
                if (lane == 0) take_min(best_d, best_i, e, n);
            }
        }
    }

    // CTA reduction on (distance, index)
    for (int off = 16; off > 0; off >>= 1) {
        double od = __shfl_down_sync(FULL, best_d, off);
        long long oi = __shfl_down_sync(FULL, best_i, off);
        take_min(best_d, best_i, od, oi);
    }
    if (lane == 0) {
        s_rd[warp] = best_d;
        s_ri[warp] = best_i;
    }
    __syncthreads();
    nn_trace(4);
    if (tid < 32) {
        best_d = tid < NN_THREADS / 32 ? s_rd[tid] : 0.0;
        best_i = tid < NN_THREADS / 32 ? s_ri[tid] : -1;
        for (int off = 16; off > 0; off >>= 1) {
            double od = __shfl_down_sync(FULL, best_d, off);
            long long oi = __shfl_down_sync(FULL, best_i, off);
            take_min(best_d, best_i, od, oi);
        }
        // everything this kernel shares with the previous one (scratch, tickets, result block) is touched from here on
        if (tid == 0 && p.overlap_prev) grid_dep_wait();
        if (p.host_parts) {
            // single query the host is waiting for (mps_tree_nearest): every CTA posts its (distance, index) straight
            // into pinned host memory as ONE 16-byte store - distance bits | sequence number << 40 | index (40 bits,
            // all ones = none) - and the host folds the grid's records as they arrive.  No ticket, no fence, no last
            // CTA: the cross-CTA hand-off took 3.9 us of a 25 us launch (profiles/tools/nn_trace_probe.cu).
            if (tid == 0) {
                const unsigned long long w1 = ((unsigned long long)(p.host_seq & 0xffffffu) << 40) |
                                              ((unsigned long long)best_i & 0xffffffffffull);
                *reinterpret_cast<ulonglong2*>(p.host_parts + 2 * (size_t)cta) =
                    make_ulonglong2((unsigned long long)__double_as_longlong(best_d), w1);
                s_last = false;
            }
        } else if (tid == 0 && ncta == 1) {
            // small trees: one CTA, no cross-CTA hand-off (saves ~4 dependent global round trips)
            p.out_idx[q] = best_i;
            p.out_dist[q] = best_i < 0 ? __longlong_as_double(0x7ff0000000000000ll) : best_d;
            if (p.mirror_idx && q == 0) {
                *p.mirror_idx = best_i;
                *p.mirror_dist = best_i < 0 ? __longlong_as_double(0x7ff0000000000000ll) : best_d;
            }
            if (p.done_flag) {
                __threadfence_system();
                *p.done_flag = p.done_seq;
            }
            s_last = false;
        } else if (tid == 0) {
            p.part_d[(size_t)q * p.grid_x + cta] = best_d;
            p.part_i[(size_t)q * p.grid_x + cta] = best_i;
            __threadfence();
            unsigned int t = atomicAdd(&p.tickets[q], 1u);
            s_last = (t == ncta - 1);
        }
    }
    __syncthreads();
    nn_trace(5);
    if (s_last) {
        // the last CTA folds the per-CTA results: every thread fetches its entries at once (one round trip to L2 for
        // the usual grid of 296; a single warp walking the list took ten dependent ones), then warp and CTA reduction
        __threadfence();
        double fd = 0.0;
        long long fi = -1;
        for (int c = tid; c < (int)ncta; c += NN_THREADS) {
            double od = __ldcg(&p.part_d[(size_t)q * p.grid_x + c]);
            long long oi = __ldcg(&p.part_i[(size_t)q * p.grid_x + c]);
            take_min(fd, fi, od, oi);
        }
        for (int off = 16; off > 0; off >>= 1) {
            double od = __shfl_down_sync(FULL, fd, off);
            long long oi = __shfl_down_sync(FULL, fi, off);
            take_min(fd, fi, od, oi);
        }
        if (lane == 0) {
            s_rd[warp] = fd;
            s_ri[warp] = fi;
        }
        __syncthreads();
        if (tid < 32) {
            fd = tid < NN_THREADS / 32 ? s_rd[tid] : 0.0;
            fi = tid < NN_THREADS / 32 ? s_ri[tid] : -1;
            for (int off = 16; off > 0; off >>= 1) {
                double od = __shfl_down_sync(FULL, fd, off);
                long long oi = __shfl_down_sync(FULL, fi, off);
                take_min(fd, fi, od, oi);
            }
            if (tid == 0) {
                p.out_idx[q] = fi;
                p.out_dist[q] = fi < 0 ? __longlong_as_double(0x7ff0000000000000ll) : fd;
                if (p.mirror_idx && q == 0) {
                    *p.mirror_idx = fi;
                    *p.mirror_dist = fi < 0 ? __longlong_as_double(0x7ff0000000000000ll) : fd;
                }
                p.tickets[q] = 0u;
                if (p.done_flag) {
                    __threadfence_system();
                    *p.done_flag = p.done_seq;
                }
                nn_trace(6);
            }
        }
    }
}


// ---- batches: a tile of NN_QT queries per CTA ----------------------------------------------------------------------
// mps_tree_nearest_batch with many queries (Q >= NN_QT, plain tree).  A thread loads its four nodes of a body's rows
// ONCE and evaluates the fp32 metric for the NN_QT queries of the CTA's tile, the accumulators of all of them in
// registers: the scan of a batch is bound by L2 bandwidth when every query reads the rows for itself (13.4 us per
// query), by arithmetic when a tile shares them.  Per query everything else is the single-query scan's: CTA bound,
// deferred candidates re-evaluated in double, (double, index) minimum, ticket hand-off, last CTA folds.  A candidate
// list that overflows (adversarial orderings) sends the query to a second pass of the CTA over its nodes with the
// final bound, evaluated exactly on the spot.
#define NN_QT 8
#define NN_TILE_CAND 128

// Packed fp32 (sm_100a FFMA2 / FADD2 / FMUL2): two IEEE operations per instruction, bit for bit the scalar ones.  The tiled
// scan is bound by issue slots once its loads are pipelined (72 % busy, FMA pipe 50 %), and a packed instruction takes one
// slot for two nodes (profiles/r02_ffma2_probe.txt: same flops, half the instructions).
#ifndef NN_TILE_F32X2
#define NN_TILE_F32X2 1  // 0: the scalar arithmetic (same bits; 5.71 against 5.38 us per query at Q = 128)
#endif
#if NN_TILE_F32X2
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float lo, float hi)
{
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2_t v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b)
{
    f32x2_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b)
{
    f32x2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c)
{
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// two nodes of one body against one query: the operations of body_term_f32 and of the accumulation, pairwise
__device__ __forceinline__ f32x2_t body_term_pair(f32x2_t acc, float x0, float x1, float y0, float y1, float t0, float t1,
                                                  const float4 qw, f32x2_t pw2, f32x2_t ow2)
{
    const f32x2_t dx = sub2(pk2(x0, x1), pk2(qw.x, qw.x));
    const f32x2_t dy = sub2(pk2(y0, y1), pk2(qw.y, qw.y));
    float s0, s1;
    upk2(fma2(dx, dx, mul2(dy, dy)), s0, s1);
    const float d0 = fabsf(t0 - qw.z), d1 = fabsf(t1 - qw.z);
    const float a0 = fminf(d0, MPS_TWO_PI_F - d0), a1 = fminf(d1, MPS_TWO_PI_F - d1);
    const f32x2_t term = fma2(pw2, pk2(sqrt_approx(s0), sqrt_approx(s1)), mul2(ow2, pk2(a0, a1)));
    return fma2(pk2(qw.w, qw.w), term, acc);
}
#endif

template <bool STREAM>
__global__ void __launch_bounds__(NN_THREADS, NN_MIN_BLOCKS) nn_scan_tile_kernel(const __grid_constant__ NNParams p)
{
    __shared__ float4 s_qw[NN_QT][MPS_MAX_BODIES];  // per query and body: (qx, qy, qtheta, weight or 0 when masked out)
    __shared__ float s_qf[NN_QT][3 * MPS_MAX_BODIES];  // the queries as the exact pass reads them
    __shared__ float s_wb[MPS_MAX_BODIES];
    __shared__ uint32_t s_mask[NN_QT];
    __shared__ int s_filter[NN_QT];
    __shared__ int s_act[MPS_MAX_BODIES];
    __shared__ int s_nact, s_any_filter;
    __shared__ unsigned int s_bound[NN_QT], s_ncand[NN_QT];
    __shared__ unsigned int s_cand_d[NN_QT][NN_TILE_CAND];
    __shared__ long long s_cand_i[NN_QT][NN_TILE_CAND];
    __shared__ double s_rd[NN_THREADS / 32];
    __shared__ long long s_ri[NN_THREADS / 32];
    __shared__ bool s_last[NN_QT];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int q0 = blockIdx.x * NN_QT;
    const int nq = min(NN_QT, p.Q - q0);
    const unsigned int cta = blockIdx.y, ncta = gridDim.y;
    const int B = p.tree.B;
    grid_dep_wait();
    grid_dep_trigger();
    const uint32_t full = B < 32 ? (1u << B) - 1u : 0xffffffffu;
    if (tid < NN_QT) {
        const bool live = tid < nq;
        s_mask[tid] = live ? ((p.masks ? p.masks[q0 + tid] : 0xffffffffu) & full) : 0u;
        s_filter[tid] = live && p.filters ? p.filters[q0 + tid] : -1;
        s_bound[tid] = 0x7f800000u;
        s_ncand[tid] = 0u;
    }
    if (tid < B) s_wb[tid] = p.weights ? p.weights[tid] : 1.0f;
    __syncthreads();
    for (int e = tid; e < NN_QT * B; e += NN_THREADS) {
        const int qq = e / B, i = e - qq * B;
        float qx = 0.0f, qy = 0.0f, qth = 0.0f;
        if (qq < nq) {
            const float* src = p.queries + (size_t)(q0 + qq) * 3 * B + 3 * i;
            qx = src[0];
            qy = src[1];
            qth = src[2];
        }
        s_qf[qq][3 * i] = qx;
        s_qf[qq][3 * i + 1] = qy;
        s_qf[qq][3 * i + 2] = qth;
        s_qw[qq][i] = make_float4(qx, qy, qth, ((s_mask[qq] >> i) & 1u) ? s_wb[i] : 0.0f);
    }
    if (tid == 0) {
        uint32_t any = 0u;
        int anyf = 0;
        for (int qq = 0; qq < nq; ++qq) {
            any |= s_mask[qq];
            anyf |= s_filter[qq] >= 0 ? 1 : 0;
        }
        int n = 0;
        for (int i = 0; i < B; ++i)
            if ((any >> i) & 1u) s_act[n++] = i;
        s_nact = n;
        s_any_filter = anyf;
    }
    __syncthreads();
    const int nact = s_nact;
    const bool any_filter = s_any_filter != 0;
    const float pwf = (float)p.pw, owf = (float)p.ow;
    const size_t cap = p.tree.cap;
    const long long N = p.n;
    const long long groups = (N + 3) >> 2;
    const float inf = __uint_as_float(0x7f800000u);

    // ---- streaming pass ----------------------------------------------------------------------------------------------
    for (long long g = (long long)cta * NN_THREADS + tid; g < groups; g += (long long)ncta * NN_THREADS) {
        const long long n0 = g << 2;
#if NN_TILE_F32X2
        f32x2_t dp[NN_QT][2];
        const f32x2_t pw2 = pk2(pwf, pwf), ow2 = pk2(owf, owf);
#pragma unroll
        for (int qq = 0; qq < NN_QT; ++qq) dp[qq][0] = dp[qq][1] = pk2(0.0f, 0.0f);
#else
        float d[NN_QT][4];
#pragma unroll
        for (int qq = 0; qq < NN_QT; ++qq) d[qq][0] = d[qq][1] = d[qq][2] = d[qq][3] = 0.0f;
#endif
        // the rows of the next two bodies are on their way while the current two are evaluated for the eight queries
        // (the pass was waiting for its loads half of the time: 7.2 -> 6.5 us per query at Q = 32, 6.2 -> 5.7 at Q = 128)
        auto load2 = [&](int a0, float4 (&X)[2], float4 (&Y)[2], float4 (&T)[2]) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int a = a0 + u < nact ? a0 + u : nact - 1;
                const float* row = p.tree.states + (size_t)(3 * s_act[a]) * cap + n0;
                X[u] = STREAM ? __ldcs(reinterpret_cast<const float4*>(row)) : __ldg(reinterpret_cast<const float4*>(row));
                Y[u] = STREAM ? __ldcs(reinterpret_cast<const float4*>(row + cap))
                              : __ldg(reinterpret_cast<const float4*>(row + cap));
                T[u] = STREAM ? __ldcs(reinterpret_cast<const float4*>(row + 2 * cap))
                              : __ldg(reinterpret_cast<const float4*>(row + 2 * cap));
            }
        };
        float4 X[2], Y[2], T[2];
        load2(0, X, Y, T);
        for (int a0 = 0; a0 < nact; a0 += 2) {
            float4 Xn[2], Yn[2], Tn[2];
            load2(a0 + 2 < nact ? a0 + 2 : a0, Xn, Yn, Tn);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (a0 + u < nact) {
                    const int i = s_act[a0 + u];
#pragma unroll
                    for (int qq = 0; qq < NN_QT; ++qq) {
                        const float4 qw = s_qw[qq][i];
#if NN_TILE_F32X2
                        dp[qq][0] = body_term_pair(dp[qq][0], X[u].x, X[u].y, Y[u].x, Y[u].y, T[u].x, T[u].y, qw, pw2, ow2);
                        dp[qq][1] = body_term_pair(dp[qq][1], X[u].z, X[u].w, Y[u].z, Y[u].w, T[u].z, T[u].w, qw, pw2, ow2);
#else
                        d[qq][0] = fmaf(qw.w, body_term_f32(X[u].x, Y[u].x, T[u].x, qw.x, qw.y, qw.z, pwf, owf), d[qq][0]);
                        d[qq][1] = fmaf(qw.w, body_term_f32(X[u].y, Y[u].y, T[u].y, qw.x, qw.y, qw.z, pwf, owf), d[qq][1]);
                        d[qq][2] = fmaf(qw.w, body_term_f32(X[u].z, Y[u].z, T[u].z, qw.x, qw.y, qw.z, pwf, owf), d[qq][2]);
                        d[qq][3] = fmaf(qw.w, body_term_f32(X[u].w, Y[u].w, T[u].w, qw.x, qw.y, qw.z, pwf, owf), d[qq][3]);
#endif
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                X[u] = Xn[u];
                Y[u] = Yn[u];
                T[u] = Tn[u];
            }
        }
        int4 sl = make_int4(0, 0, 0, 0);
        if (any_filter) sl = __ldg(reinterpret_cast<const int4*>(p.tree.slice + n0));
        const bool t1 = n0 + 1 >= N, t2 = n0 + 2 >= N, t3 = n0 + 3 >= N;
        const unsigned am = __activemask();
#pragma unroll
        for (int qq = 0; qq < NN_QT; ++qq) {
            if (qq >= nq) break;  // uniform
#if NN_TILE_F32X2
            float d0, d1, d2, d3;
            upk2(dp[qq][0], d0, d1);
            upk2(dp[qq][1], d2, d3);
#else
            float d0 = d[qq][0], d1 = d[qq][1], d2 = d[qq][2], d3 = d[qq][3];
#endif
            const int f = s_filter[qq];
            if (f >= 0) {
                if (sl.x != f) d0 = inf;
                if (sl.y != f) d1 = inf;
                if (sl.z != f) d2 = inf;
                if (sl.w != f) d3 = inf;
            }
            if (t1) d1 = inf;
            if (t2) d2 = inf;
            if (t3) d3 = inf;
            const float dm = fminf(fminf(d0, d1), fminf(d2, d3));
            const unsigned wmin = __reduce_min_sync(am, __float_as_uint(dm));
            if (lane == (__ffs(am) - 1)) atomicMin(&s_bound[qq], wmin);
            __syncwarp(am);
            const float b = fminf(__uint_as_float(*reinterpret_cast<volatile unsigned int*>(&s_bound[qq])),
                                  __uint_as_float(wmin));
            const float thr = b * (1.0f + NN_REL_EPS) + p.abs_eps;
            if (dm <= thr) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float dj = j == 0 ? d0 : j == 1 ? d1 : j == 2 ? d2 : d3;
                    if (dj <= thr && dj < inf) {
                        const unsigned slot = atomicAdd(&s_ncand[qq], 1u);  // beyond the capacity: second pass below
                        if (slot < NN_TILE_CAND) {
                            s_cand_d[qq][slot] = __float_as_uint(dj);
                            s_cand_i[qq][slot] = n0 + j;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- per query: exact pass, CTA minimum, hand-off ---------------------------------------------------------------
    for (int qq = 0; qq < nq; ++qq) {
        const int q = q0 + qq;
        const uint32_t mask = s_mask[qq];
        const float thr = __uint_as_float(s_bound[qq]) * (1.0f + NN_REL_EPS) + p.abs_eps;
        double best_d = __longlong_as_double(0x7ff0000000000000ll);
        long long best_i = -1;
        if (s_ncand[qq] <= NN_TILE_CAND) {
            const int ncand = (int)s_ncand[qq];
            for (int c = warp; c < ncand; c += NN_THREADS / 32) {
                if (__uint_as_float(s_cand_d[qq][c]) <= thr) {
                    const long long n = s_cand_i[qq][c];
                    const double e = warp_exact_distance(p.tree, n, s_qf[qq], s_wb, mask, p.pw, p.ow, lane);
                    if (lane == 0) take_min(best_d, best_i, e, n);
                }
            }
        } else {
            // the list overflowed: once more over this CTA's nodes for this query with the final bound, every node
            // within the slack evaluated exactly on the spot
            const int f = s_filter[qq];
            for (long long g = (long long)cta * NN_THREADS + tid; g < groups; g += (long long)ncta * NN_THREADS) {
                const long long n0 = g << 2;
                float dd[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                for (int a = 0; a < nact; ++a) {
                    const int i = s_act[a];
                    const float4 qw = s_qw[qq][i];
                    const float* row = p.tree.states + (size_t)(3 * i) * cap + n0;
                    const float4 X = __ldg(reinterpret_cast<const float4*>(row)), Y = __ldg(reinterpret_cast<const float4*>(row + cap));
                    const float4 T = __ldg(reinterpret_cast<const float4*>(row + 2 * cap));
                    dd[0] = fmaf(qw.w, body_term_f32(X.x, Y.x, T.x, qw.x, qw.y, qw.z, pwf, owf), dd[0]);
                    dd[1] = fmaf(qw.w, body_term_f32(X.y, Y.y, T.y, qw.x, qw.y, qw.z, pwf, owf), dd[1]);
                    dd[2] = fmaf(qw.w, body_term_f32(X.z, Y.z, T.z, qw.x, qw.y, qw.z, pwf, owf), dd[2]);
                    dd[3] = fmaf(qw.w, body_term_f32(X.w, Y.w, T.w, qw.x, qw.y, qw.z, pwf, owf), dd[3]);
                }
                int4 sl = make_int4(f, f, f, f);
                if (f >= 0) sl = __ldg(reinterpret_cast<const int4*>(p.tree.slice + n0));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int sj = j == 0 ? sl.x : j == 1 ? sl.y : j == 2 ? sl.z : sl.w;
                    if (n0 + j < N && sj == f && dd[j] <= thr) {
                        const double e = exact_distance(p.tree, n0 + j, s_qf[qq], s_wb, mask, p.pw, p.ow);
                        take_min(best_d, best_i, e, n0 + j);
                    }
                }
            }
        }
        for (int off = 16; off > 0; off >>= 1) {
            double od = __shfl_down_sync(FULL, best_d, off);
            long long oi = __shfl_down_sync(FULL, best_i, off);
            take_min(best_d, best_i, od, oi);
        }
        if (lane == 0) {
            s_rd[warp] = best_d;
            s_ri[warp] = best_i;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < NN_THREADS / 32; ++w) take_min(best_d, best_i, s_rd[w], s_ri[w]);
            if (ncta == 1) {
                p.out_idx[q] = best_i;
                p.out_dist[q] = best_i < 0 ? __longlong_as_double(0x7ff0000000000000ll) : best_d;
                s_last[qq] = false;
            } else {
                p.part_d[(size_t)q * p.grid_x + cta] = best_d;
                p.part_i[(size_t)q * p.grid_x + cta] = best_i;
                __threadfence();
                s_last[qq] = atomicAdd(&p.tickets[q], 1u) == ncta - 1;
            }
        }
        __syncthreads();
    }
    // the last CTA of a query folds its per-CTA results
    for (int qq = 0; qq < nq; ++qq) {
        if (!s_last[qq]) continue;  // uniform
        const int q = q0 + qq;
        __threadfence();
        double fd = 0.0;
        long long fi = -1;
        for (int c = tid; c < (int)ncta; c += NN_THREADS) {
            double od = __ldcg(&p.part_d[(size_t)q * p.grid_x + c]);
            long long oi = __ldcg(&p.part_i[(size_t)q * p.grid_x + c]);
            take_min(fd, fi, od, oi);
        }
        for (int off = 16; off > 0; off >>= 1) {
            double od = __shfl_down_sync(FULL, fd, off);
            long long oi = __shfl_down_sync(FULL, fi, off);
            take_min(fd, fi, od, oi);
        }
        if (lane == 0) {
            s_rd[warp] = fd;
            s_ri[warp] = fi;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < NN_THREADS / 32; ++w) take_min(fd, fi, s_rd[w], s_ri[w]);
            p.out_idx[q] = fi;
            p.out_dist[q] = fi < 0 ? __longlong_as_double(0x7ff0000000000000ll) : fd;
            p.tickets[q] = 0u;
        }
        __syncthreads();
    }
}

#endif  // MPS_CPU_EMU

// ------------------------------------------------------------------------------------------------------------
// Exact k nearest neighbours (k <= 32): ompl::NearestNeighbors<MotionPtr>::nearestK of the interface the reference
// binds (include/mps/planner/pushing/algorithm/RRT.h:149), under the same metric and tie rule as the 1-NN scan:
// ascending (double distance, node index).
//
// One streaming pass.  Every warp keeps the k smallest keys it has seen, key = fp32 distance bits << 32 | node index,
// one key per lane in ascending order (a warp-level insertion: ballot for the position, shuffle-up to make room), and
// the smallest fp32 distance of anything it dropped or never admitted.  The CTA merges its warps' lists by ranking
// (each thread counts the keys below its own), the last CTA does the same over all CTAs' lists.  fp32 order is not
// the order of the reference's double metric, so the result is fixed up exactly: every node whose fp32 distance is
// within the slack of the k-th smallest is re-evaluated with the reference's double expression and the candidates are
// ranked by (double, index).  That is exact as long as no node that could belong to the answer was dropped before the
// last stage, which the kernel checks: the smallest dropped distance must lie above the slack (it does unless k + 1
// near-equal distances meet in one warp, e.g. duplicate nodes); otherwise status = 1 and the host falls back to k
// successor scans in double (nn_next_kernel).
#define NN_KEY_SENTINEL 0x7f800000ffffffffull

__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src)
{
    return __shfl_sync(FULL, v, src);
}

// insert `key` (warp-uniform) into the warp's ascending list (one entry per lane, lanes >= k hold the sentinel);
// returns the entry that fell off the end (sentinel if the list was not full)
__device__ __forceinline__ unsigned long long warp_list_insert(unsigned long long& mine, unsigned long long key, int k,
                                                               int lane)
{
    const int pos = __popc(__ballot_sync(FULL, mine < key));
    const unsigned long long up = __shfl_up_sync(FULL, mine, 1);
    const unsigned long long dropped = shfl_u64(mine, k - 1);
    if (lane == pos)
        mine = key;
    else if (lane > pos)
        mine = up;
    if (lane >= k) mine = NN_KEY_SENTINEL;
    return dropped;
}

// the 32 smallest keys of two ascending 32-entry lists (one entry per lane), ascending
__device__ __forceinline__ unsigned long long warp_merge32(unsigned long long a, unsigned long long b, int lane)
{
    const unsigned long long rb = shfl_u64(b, 31 - lane);
    unsigned long long m = a < rb ? a : rb;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const unsigned long long o = __shfl_xor_sync(FULL, m, s);
        const bool up = (lane & s) != 0;
        m = (up == (m < o)) ? o : m;  // lower lane keeps the smaller, upper lane the larger
    }
    return m;
}

// one key per lane -> ascending over the lanes (bitonic network, 15 compare-exchange stages)
__device__ __forceinline__ unsigned long long warp_sort32(unsigned long long v, int lane)
{
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            const unsigned long long o = __shfl_xor_sync(FULL, v, j);
            const bool take_max = ((lane & j) != 0) != ((lane & k2) != 0);  // upper lane of the pair, block ascending
            v = (take_max == (v < o)) ? o : v;
        }
    }
    return v;
}

// more candidates than this in one offer: sort them and merge the two lists instead of inserting one by one (an
// insertion is ~4 dependent shuffles, the sort + merge 21 stages whatever the number of candidates)
#ifndef NN_OFFER_SERIAL_MAX
#define NN_OFFER_SERIAL_MAX 5
#endif

// the keys a warp holds (one per lane) are offered to the list; rej collects the fp32 bits of what does not stay
__device__ __forceinline__ void warp_list_offer(unsigned long long& mine, unsigned long long& kth, unsigned int& rej,
                                                unsigned long long key, bool valid, int k, int lane)
{
    const bool cand = valid && key < kth;
    unsigned bal = __ballot_sync(FULL, cand);
    if (valid && !cand) rej = min(rej, (unsigned int)(key >> 32));
    if (__popc(bal) > NN_OFFER_SERIAL_MAX) {
        // many candidates (the first trips of a warp, before its list has tightened)
        const unsigned long long c = warp_sort32(cand ? key : NN_KEY_SENTINEL, lane);
        const unsigned long long old = mine;
        mine = warp_merge32(mine, c, lane);
        if (lane >= k) mine = NN_KEY_SENTINEL;
        kth = shfl_u64(mine, k - 1);
        // what did not stay: list entries and offered keys above the new k-th (keys are unique)
        if (old != NN_KEY_SENTINEL && old > kth) rej = min(rej, (unsigned int)(old >> 32));
        if (cand && key > kth) rej = min(rej, (unsigned int)(key >> 32));
        return;
    }
    while (bal) {
        const int src = __ffs(bal) - 1;
        bal &= bal - 1u;
        const unsigned long long k64 = shfl_u64(key, src);
        if (k64 < kth) {
            const unsigned long long dropped = warp_list_insert(mine, k64, k, lane);
            if (dropped != NN_KEY_SENTINEL) rej = min(rej, (unsigned int)(dropped >> 32));
            kth = shfl_u64(mine, k - 1);
        } else {
            rej = min(rej, (unsigned int)(k64 >> 32));  // the list tightened since the ballot
        }
    }
}

// CTA merge of the warps' lists (s_keys[w * 32 + l] = entry l of warp w's ascending list, sentinels at the end):
// a key's rank is its position in its own list plus, for every other list, the number of entries below it (binary
// search; keys are unique).  s_out[0 .. k) receives the k smallest in ascending order; the smallest fp32 bits among
// the valid keys that did not make it goes to *s_rej (atomicMin).
__device__ __forceinline__ void cta_rank_merge(const unsigned long long* s_keys, unsigned long long* s_out, int k,
                                               unsigned int* s_rej, int tid)
{
    for (int i = tid; i < k; i += NN_THREADS) s_out[i] = NN_KEY_SENTINEL;
    __syncthreads();
    const unsigned long long mine = s_keys[tid];
    if (mine != NN_KEY_SENTINEL) {
        const int w = tid >> 5;
        int rank = tid & 31;
        for (int o = 0; o < NN_THREADS / 32 && rank < k; ++o) {
            if (o == w) continue;
            const unsigned long long* L = s_keys + o * 32;
            int lo = 0, hi = 32;  // first entry of list o that is not below `mine`
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const int mid = (lo + hi) >> 1;
                if (L[mid] < mine)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            if (lo < 32 && L[lo] < mine) ++lo;
            rank += lo;
        }
        if (rank < k)
            s_out[rank] = mine;
        else
            atomicMin(s_rej, (unsigned int)(mine >> 32));
    }
    __syncthreads();
}

template <bool STREAM>
__global__ void __launch_bounds__(NN_THREADS, NN_MIN_BLOCKS) nn_topk_kernel(const __grid_constant__ TopKParams tp)
{
    const NNParams& p = tp.nn;
    __shared__ float s_q[3 * MPS_MAX_BODIES];
    __shared__ float s_w[MPS_MAX_BODIES];
    __shared__ float s_wb[MPS_MAX_BODIES];
    __shared__ int s_act[MPS_MAX_BODIES];
    __shared__ int s_nact;
    __shared__ unsigned long long s_keys[NN_THREADS];  // the warps' lists
    __shared__ unsigned long long s_out[NN_TOPK_MAX];
    __shared__ unsigned int s_rej;
    __shared__ bool s_last;
    __shared__ unsigned int s_ncand;
    __shared__ long long s_ci[NN_TOPK_CAND];
    __shared__ double s_cd[NN_TOPK_CAND];
    __shared__ unsigned int s_nhot, s_nbuf;
    __shared__ unsigned short s_hot[NN_TOPK_HOT];            // lists whose head lies within the bound (last CTA)
    __shared__ unsigned long long s_buf[NN_TOPK_BUF];        // their keys within the bound
    __shared__ double s_term[NN_TOPK_CAND * MPS_MAX_BODIES];  // body terms of the candidates' exact distances

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int B = p.tree.B;
    const int k = tp.k;
    uint32_t mask = p.inline_mask;
    if (B < 32) mask &= (1u << B) - 1u;
    const int filter = p.filter_from ? (int)*p.filter_from : p.inline_filter;
    if (tid < 3 * B) s_q[tid] = p.q_inline[tid];
    if (tid < B) {
        const float w = p.inline_has_w ? p.w_inline[tid] : 1.0f;
        s_wb[tid] = w;
        if ((mask >> tid) & 1u) {
            const int a = __popc(mask & ((1u << tid) - 1u));
            s_act[a] = tid;
            s_w[a] = w;
        }
    }
    if (tid == 0) {
        s_nact = __popc(mask);
        s_rej = 0x7f800000u;
        s_ncand = 0u;
    }
    __syncthreads();
    const int nact = s_nact;
    const float pwf = (float)p.pw, owf = (float)p.ow;
    const size_t cap = p.tree.cap;
    const long long N = p.n;
    const long long groups = (N + 3) >> 2;
    const float inf = __uint_as_float(0x7f800000u);

    // ---- streaming pass: the whole warp stays in the loop (the list operations are warp collectives) ----------
    unsigned long long mine = NN_KEY_SENTINEL, kth = NN_KEY_SENTINEL;
    unsigned int rej = 0x7f800000u;
    for (long long gw = (long long)blockIdx.x * NN_THREADS + (tid & ~31); gw < groups; gw += (long long)gridDim.x * NN_THREADS) {
        const long long g = gw + lane;
        float d0 = inf, d1 = inf, d2 = inf, d3 = inf;
        const long long n0 = g << 2;
        if (g < groups) group_distances<STREAM>(p, s_act, s_q, s_w, nact, cap, 0, n0, N, filter, pwf, owf, d0, d1, d2, d3);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float dj = j == 0 ? d0 : j == 1 ? d1 : j == 2 ? d2 : d3;
            const unsigned long long key = ((unsigned long long)__float_as_uint(dj) << 32) | (unsigned long long)(unsigned int)(n0 + j);
#if !defined(NN_TOPK_PROBE) || NN_TOPK_PROBE < 3 || NN_TOPK_PROBE >= 4
            warp_list_offer(mine, kth, rej, key, dj < inf, k, lane);
#else
            rej = min(rej, (unsigned int)(key >> 32));
#endif
        }
    }
    rej = __reduce_min_sync(FULL, rej);
    s_keys[tid] = mine;
    if (lane == 0) atomicMin(&s_rej, rej);
    __syncthreads();
#if !defined(NN_TOPK_PROBE) || NN_TOPK_PROBE < 2 || NN_TOPK_PROBE >= 4
    cta_rank_merge(s_keys, s_out, k, &s_rej, tid);
#endif
    if (tid < k) {
        tp.part_keys[(size_t)blockIdx.x * k + tid] = s_out[tid];
        __threadfence();  // every writer of the list fences its own store before the CTA takes its ticket
    }
    __syncthreads();
    if (tid == 0) {
        tp.part_rej[blockIdx.x] = s_rej;
        __threadfence();
        const unsigned int t = atomicAdd(&p.tickets[0], 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
#if defined(NN_TOPK_PROBE) && NN_TOPK_PROBE >= 1 && NN_TOPK_PROBE <= 3
    if (tid == 0) p.tickets[0] = 0u;
    return;
#endif

    // ---- last CTA: the k smallest fp32 keys of all CTAs' lists, then the exact fix-up ---------------------------------
    // The lists are ascending, so the k-th smallest of their HEADS, U, bounds the k-th smallest key of the whole tree from
    // above (k keys <= U exist), and every key that can matter - the k smallest and whatever lies within the slack of the
    // k-th - is <= U' = U inflated by the slack and sits in a list whose head is <= U' (usually k lists or a few more out
    // of 245).  Three dependent trips to L2 (heads, the hot lists' keys, the dropped minima) instead of a merge chain
    // over all lists.
    __threadfence();
    const int lists = (int)gridDim.x;
    const unsigned long long* keys = tp.part_keys;  // read through L2 (__ldcg): written by other SMs before their tickets
    {
        unsigned long long hmine = NN_KEY_SENTINEL, hkth = NN_KEY_SENTINEL;
        unsigned int hrej = 0u;
        for (int l0 = warp * 32; l0 < lists; l0 += NN_THREADS) {
            const int l = l0 + lane;
            const unsigned long long h = l < lists ? __ldcg(keys + (size_t)l * k) : NN_KEY_SENTINEL;
            warp_list_offer(hmine, hkth, hrej, h, h != NN_KEY_SENTINEL, k, lane);
        }
        s_keys[tid] = hmine;
        if (tid < NN_TOPK_MAX) s_out[tid] = NN_KEY_SENTINEL;
        if (tid == 0) {
            s_rej = 0x7f800000u;
            s_nhot = 0u;
            s_nbuf = 0u;
        }
        __syncthreads();
        for (int stride = 1; stride < NN_THREADS / 32; stride <<= 1) {
            if ((warp & (2 * stride - 1)) == 0) {
                hmine = warp_merge32(hmine, s_keys[(warp + stride) * 32 + lane], lane);
                s_keys[tid] = hmine;
            }
            __syncthreads();
        }
    }
    const unsigned long long U = s_keys[k - 1];  // sentinel: fewer than k non-empty lists, every key matters
    const float thrU = U == NN_KEY_SENTINEL ? inf : __uint_as_float((unsigned int)(U >> 32)) * (1.0f + NN_REL_EPS) + p.abs_eps;
    for (int l = tid; l < lists; l += NN_THREADS) {
        const unsigned long long h = __ldcg(keys + (size_t)l * k);
        if (h != NN_KEY_SENTINEL && __uint_as_float((unsigned int)(h >> 32)) <= thrU) {
            const unsigned int slot = atomicAdd(&s_nhot, 1u);
            if (slot < NN_TOPK_HOT) s_hot[slot] = (unsigned short)l;
        }
    }
    __syncthreads();
    int status = 0;
    if (s_nhot > NN_TOPK_HOT || lists > 65535) status = 1;
    const int nhot = (int)min(s_nhot, (unsigned int)NN_TOPK_HOT);
    for (int i = tid; i < nhot * k; i += NN_THREADS) {
        const int hl = i / k;
        const unsigned long long key = __ldcg(keys + (size_t)s_hot[hl] * k + (i - hl * k));
        if (key != NN_KEY_SENTINEL && __uint_as_float((unsigned int)(key >> 32)) <= thrU) {
            const unsigned int slot = atomicAdd(&s_nbuf, 1u);
            if (slot < NN_TOPK_BUF) s_buf[slot] = key;
        }
    }
    // smallest distance any CTA dropped: it must lie beyond the slack
    unsigned int gr = 0x7f800000u;
    for (int c = tid; c < lists; c += NN_THREADS) gr = min(gr, __ldcg(tp.part_rej + c));
    gr = __reduce_min_sync(FULL, gr);
    if (lane == 0) atomicMin(&s_rej, gr);
    __syncthreads();
#if defined(NN_TOPK_PROBE) && NN_TOPK_PROBE == 4
    if (tid == 0) p.tickets[0] = 0u;
    return;
#endif
    if (s_nbuf > NN_TOPK_BUF) status = 1;  // more near-equal distances than the buffer holds: the caller falls back
    const int m = (int)min(s_nbuf, (unsigned int)NN_TOPK_BUF);
    // the k smallest of the buffer by ranking (keys are unique)
    for (int i = tid; i < m; i += NN_THREADS) {
        const unsigned long long mk = s_buf[i];
        int rank = 0;
        for (int j = 0; j < m; ++j) rank += s_buf[j] < mk ? 1 : 0;
        if (rank < k) s_out[rank] = mk;
    }
    __syncthreads();
    const int have = m < k ? m : k;
    const unsigned int lo = have > 0 ? (unsigned int)(s_out[have - 1] >> 32) : 0x7f800000u;
    float thr = inf;
    if (have > 0) thr = __uint_as_float(lo) * (1.0f + NN_REL_EPS) + p.abs_eps;
    // candidates: every buffered key within the slack of the k-th (thr <= thrU: all of them are in the buffer)
    for (int i = tid; i < m; i += NN_THREADS) {
        const unsigned long long key = s_buf[i];
        if (__uint_as_float((unsigned int)(key >> 32)) <= thr) {
            const unsigned int slot = atomicAdd(&s_ncand, 1u);
            if (slot < NN_TOPK_CAND) s_ci[slot] = (long long)(unsigned int)(key & 0xffffffffull);
        }
    }
    __syncthreads();
#if defined(NN_TOPK_PROBE) && NN_TOPK_PROBE == 5
    if (tid == 0) p.tickets[0] = 0u;
    return;
#endif
    if (have == k && __uint_as_float(s_rej) <= thr) status = 1;
    if (s_ncand > NN_TOPK_CAND) status = 1;
    const int nc = (int)min(s_ncand, (unsigned int)NN_TOPK_CAND);
    // the reference's double expression for every candidate: thread = (candidate, body), all loads in flight at once;
    // the terms are then added in body order (the rounding of the sequential loop, as in warp_exact_distance)
    for (int w = tid; w < nc * B; w += NN_THREADS) {
        const int c = w / B, b = w - c * B;
        double term = 0.0;
        if ((mask >> b) & 1u) {
            const long long n = s_ci[c];
            const float x = p.tree.states[(size_t)(3 * b) * cap + n];
            const float y = p.tree.states[(size_t)(3 * b + 1) * cap + n];
            const float th = p.tree.states[(size_t)(3 * b + 2) * cap + n];
            term = (double)s_wb[b] * mps_body_distance(x, y, th, s_q[3 * b], s_q[3 * b + 1], s_q[3 * b + 2], p.pw, p.ow);
        }
        s_term[w] = term;
    }
    __syncthreads();
    if (tid < nc) {
        double dist = 0.0;
        for (int b = 0; b < B; ++b)
            if ((mask >> b) & 1u) dist += s_term[tid * B + b];
        s_cd[tid] = dist;
    }
    __syncthreads();
    if (tid < nc) {
        const double md = s_cd[tid];
        const long long mi = s_ci[tid];
        int rank = 0;
        for (int j = 0; j < nc; ++j) rank += (s_cd[j] < md || (s_cd[j] == md && s_ci[j] < mi)) ? 1 : 0;
        if (rank < k) {
            if (tp.out_rec) {
                *reinterpret_cast<ulonglong2*>(tp.out_rec + 2 * rank) =
                    make_ulonglong2((unsigned long long)__double_as_longlong(md),
                                    ((unsigned long long)(tp.done_seq & 0xffffffu) << 40) | ((unsigned long long)mi & 0xffffffffffull));
            } else {
                tp.out_idx[rank] = mi;
                tp.out_dist[rank] = md;
            }
        }
    }
    if (tid == 0) {
        if (tp.out_rec) {
            *reinterpret_cast<ulonglong2*>(tp.out_rec + 2 * NN_TOPK_MAX) =
                make_ulonglong2((unsigned long long)(unsigned int)(nc < k ? nc : k) | ((unsigned long long)(unsigned int)status << 32),
                                (unsigned long long)(tp.done_seq & 0xffffffu) << 40);
        } else {
            tp.out_meta[0] = nc < k ? nc : k;
            tp.out_meta[1] = status;
        }
        p.tickets[0] = 0u;
    }
    if (tp.done_flag && !tp.out_rec) {
        __threadfence_system();  // every writer of the result block, before the CTA signs it
        __syncthreads();
        if (tid == 0) *tp.done_flag = tp.done_seq;
    }
}

// successor scan in double: the smallest (distance, index) above (after_d, after_i).  Plain: one thread per node,
// the reference's expression for every node (a few hundred microseconds per million nodes).
__global__ void __launch_bounds__(NN_THREADS) nn_next_kernel(const __grid_constant__ NextParams np)
{
    const NNParams& p = np.nn;
    __shared__ float s_q[3 * MPS_MAX_BODIES];
    __shared__ float s_wb[MPS_MAX_BODIES];
    __shared__ double s_rd[NN_THREADS / 32];
    __shared__ long long s_ri[NN_THREADS / 32];
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int B = p.tree.B;
    uint32_t mask = p.inline_mask;
    if (B < 32) mask &= (1u << B) - 1u;
    const int filter = p.inline_filter;
    if (tid < 3 * B) s_q[tid] = p.q_inline[tid];
    if (tid < B) s_wb[tid] = p.inline_has_w ? p.w_inline[tid] : 1.0f;
    __syncthreads();
    double best_d = 0.0;
    long long best_i = -1;
    for (long long n = (long long)blockIdx.x * NN_THREADS + tid; n < p.n; n += (long long)gridDim.x * NN_THREADS) {
        if (filter >= 0 && p.tree.slice[n] != filter) continue;
        const double e = exact_distance(p.tree, n, s_q, s_wb, mask, p.pw, p.ow);
        if (e > np.after_d || (e == np.after_d && n > np.after_i)) take_min(best_d, best_i, e, n);
    }
    for (int off = 16; off > 0; off >>= 1) {
        double od = __shfl_down_sync(FULL, best_d, off);
        long long oi = __shfl_down_sync(FULL, best_i, off);
        take_min(best_d, best_i, od, oi);
    }
    if (lane == 0) {
        s_rd[warp] = best_d;
        s_ri[warp] = best_i;
    }
    __syncthreads();
    if (tid < 32) {
        best_d = tid < NN_THREADS / 32 ? s_rd[tid] : 0.0;
        best_i = tid < NN_THREADS / 32 ? s_ri[tid] : -1;
        for (int off = 16; off > 0; off >>= 1) {
            double od = __shfl_down_sync(FULL, best_d, off);
            long long oi = __shfl_down_sync(FULL, best_i, off);
            take_min(best_d, best_i, od, oi);
        }
        if (tid == 0) {
            p.part_d[blockIdx.x] = best_d;
            p.part_i[blockIdx.x] = best_i;
            __threadfence();
            s_last = atomicAdd(&p.tickets[0], 1u) == gridDim.x - 1;
        }
    }
    __syncthreads();
    if (s_last && tid < 32) {
        __threadfence();
        double fd = 0.0;
        long long fi = -1;
        for (int c = tid; c < (int)gridDim.x; c += 32)
            take_min(fd, fi, *reinterpret_cast<volatile double*>(&p.part_d[c]), *reinterpret_cast<volatile long long*>(&p.part_i[c]));
        for (int off = 16; off > 0; off >>= 1) {
            double od = __shfl_down_sync(FULL, fd, off);
            long long oi = __shfl_down_sync(FULL, fi, off);
            take_min(fd, fi, od, oi);
        }
        if (tid == 0) {
            p.out_idx[0] = fi;
            p.out_dist[0] = fi < 0 ? __longlong_as_double(0x7ff0000000000000ll) : fd;
            p.tickets[0] = 0u;
        }
    }
}

#ifndef MPS_CPU_EMU
// staged node-major rows -> SoA columns [first, first + n)
__global__ void tree_append_kernel(TreeView t, long long first, long long n, const float* __restrict__ st,
                                   const int32_t* __restrict__ parents, const float* __restrict__ ctrl,
                                   const int32_t* __restrict__ slices, long long* size_out)
{
    if (size_out && blockIdx.x == 0 && threadIdx.x == 0) *size_out = first + n;
    const int dims = 3 * t.B;
    const long long total = n * dims;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        long long node = e / dims;
        int r = (int)(e % dims);
        t.states[(size_t)r * t.cap + first + node] = st[e];
    }
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n * MPS_CONTROL_DIM;
         e += (long long)gridDim.x * blockDim.x) {
        long long node = e / MPS_CONTROL_DIM;
        int c = (int)(e % MPS_CONTROL_DIM);
        t.controls[(size_t)c * t.cap + first + node] = ctrl ? ctrl[e] : 0.0f;
    }
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        t.parent[first + e] = parents ? parents[e] : -1;
        t.slice[first + e] = slices ? slices[e] : -1;
    }
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// seeded uniform states written straight into the SoA (benchmarks: no 132 MB host copy)
__global__ void tree_fill_kernel(TreeView t, long long n, unsigned long long seed, int n_slices, float x_min,
                                 float x_max, float y_min, float y_max, long long* size_out)
{
    if (size_out && blockIdx.x == 0 && threadIdx.x == 0) *size_out = n;
    const int dims = 3 * t.B;
    for (long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x; node < n;
         node += (long long)gridDim.x * blockDim.x) {
        for (int r = 0; r < dims; ++r) {
            unsigned long long h = splitmix64(seed ^ splitmix64((unsigned long long)node * 131ull + (unsigned)r));
            float u = (float)(h >> 40) * (1.0f / 16777216.0f);  // [0, 1)
            int c = r % 3;
            float v;
            if (c == 0)
                v = x_min + u * (x_max - x_min);
            else if (c == 1)
                v = y_min + u * (y_max - y_min);
            else
                v = -MPS_PI_F + u * MPS_TWO_PI_F;
            if (c == 2 && v >= MPS_PI_F) v = -MPS_PI_F;
            t.states[(size_t)r * t.cap + node] = v;
        }
        for (int c = 0; c < MPS_CONTROL_DIM; ++c) t.controls[(size_t)c * t.cap + node] = 0.0f;
        t.parent[node] = node == 0 ? -1 : (int32_t)(node - 1);
        unsigned long long hs = splitmix64(seed ^ splitmix64(0xABCDull + (unsigned long long)node));
        t.slice[node] = n_slices > 0 ? (int32_t)(hs % (unsigned long long)n_slices) : -1;
    }
}

// SoA columns [first, first + n) -> node-major rows
__global__ void tree_gather_kernel(TreeView t, long long first, long long n, float* __restrict__ out)
{
    const int dims = 3 * t.B;
    const long long total = n * dims;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        long long node = e / dims;
        int r = (int)(e % dims);
        out[e] = t.states[(size_t)r * t.cap + first + node];
    }
}

// capacity growth: row-wise copy into a wider SoA
__global__ void tree_copy_kernel(TreeView src, TreeView dst, long long n)
{
    const int rows = 3 * src.B;
    const long long total = n * rows;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        int r = (int)(e / n);
        long long node = e % n;
        dst.states[(size_t)r * dst.cap + node] = src.states[(size_t)r * src.cap + node];
    }
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n * MPS_CONTROL_DIM;
         e += (long long)gridDim.x * blockDim.x) {
        int c = (int)(e / n);
        long long node = e % n;
        dst.controls[(size_t)c * dst.cap + node] = src.controls[(size_t)c * src.cap + node];
    }
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        dst.parent[e] = src.parent[e];
        dst.slice[e] = src.slice[e];
    }
}

// every tree of the forest := { root }
__global__ void forest_roots_kernel(TreeView t, size_t seg, int n_inst, const float* __restrict__ roots, long long* sizes)
{
    const int dims = 3 * t.B;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)n_inst * dims;
         e += (long long)gridDim.x * blockDim.x) {
        int inst = (int)(e / dims), r = (int)(e % dims);
        t.states[(size_t)r * t.cap + (size_t)inst * seg] = roots[e];
    }
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_inst; e += (long long)gridDim.x * blockDim.x) {
        const size_t col = (size_t)e * seg;
        for (int c = 0; c < MPS_CONTROL_DIM; ++c) t.controls[(size_t)c * t.cap + col] = 0.0f;
        t.parent[col] = -1;
        t.slice[col] = -1;
        sizes[e] = 1;
    }
}

int grid_for(long long work, int threads, int max_blocks)
{
    long long b = (work + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

}  // namespace

// CTAs of the scan: enough to cover the node groups, at most 4 per SM (a multiple of the SM count
// for large trees; every thread then streams >= 2 groups of 3B float4 loads)
int mps_nn_grid_x(long long n, int num_sms)
{
    long long groups = (n + 3) / 4;
    long long want = (groups + NN_THREADS - 1) / NN_THREADS;
    long long cap = (long long)num_sms * NN_CTAS_PER_SM;
    if (want > cap) {
        // more than one trip per thread: the fewest CTAs that make the same number of trips as the full grid would,
        // so that every CTA makes ALL of them (1 M nodes: 245 CTAs x 4 trips instead of 89 x 4 + 207 x 3)
        static const int mode = getenv("MPS_NN_GRID_MODE") ? atoi(getenv("MPS_NN_GRID_MODE")) : 1;
        const long long per_trip = cap * NN_THREADS;
        const long long trips = (groups + per_trip - 1) / per_trip;
        long long even = (groups + trips * NN_THREADS - 1) / (trips * NN_THREADS);
        want = mode == 1 && even <= cap ? even : cap;
    }
    if (want < 1) want = 1;
    return (int)want;
}

cudaError_t mps_launch_nn_scan(const NNParams& p_in, cudaStream_t stream)
{
    NNParams p = p_in;
    if (p.filter_from) p.overlap_prev = 0;  // the filter is the previous scan's result
    // launched with programmatic stream serialization: the CTAs may become resident (and run their prologue, or - with
    // overlap_prev - their whole streaming pass) before the previous kernel of the stream has drained; the kernel
    // orders itself with griddepcontrol.wait
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p.Q, (unsigned)p.grid_x, 1);  // see nn_scan_kernel
    cfg.blockDim = dim3(NN_THREADS, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // evict-first loads while the scanned rows are within ~2x the 126 MB L2, default policy beyond
    const double scanned = (double)p.n * 12.0 * p.tree.B;
    static const int tile_min_q = getenv("MPS_NN_TILE_MIN_Q") ? atoi(getenv("MPS_NN_TILE_MIN_Q")) : NN_QT;
    if (p.Q >= tile_min_q && !p.inline_q && !p.inst_ids && !p.host_parts && p.queries) {
        // a batch on the plain tree: tiles of NN_QT queries share every load (nn_scan_tile_kernel).  CTAs per tile: as
        // many as make the whole grid one wave of the machine (the single scan's grid per tile would be 3.3 waves at
        // Q = 32, the last one a third full, and the warps of a CTA waited at the barrier behind the pass for those
        // that had one trip more: 1.4 barrier stalls per issued instruction)
        const int tiles = (p.Q + NN_QT - 1) / NN_QT;
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            sms = 148;
        int gx = sms * NN_CTAS_PER_SM / tiles;
        if (gx < 1) gx = 1;
        if (gx < p.grid_x) p.grid_x = gx;
        cfg.gridDim = dim3((unsigned)tiles, (unsigned)p.grid_x, 1);
        if (scanned < 252.0e6) return cudaLaunchKernelEx(&cfg, nn_scan_tile_kernel<true>, p);
        return cudaLaunchKernelEx(&cfg, nn_scan_tile_kernel<false>, p);
    }
    if (scanned < 252.0e6) return cudaLaunchKernelEx(&cfg, nn_scan_kernel<true>, p);
    return cudaLaunchKernelEx(&cfg, nn_scan_kernel<false>, p);
}

cudaError_t mps_launch_nn_topk(const TopKParams& p, cudaStream_t stream)
{
    dim3 grid((unsigned)p.nn.grid_x, 1, 1);
    const double scanned = (double)p.nn.n * 12.0 * p.nn.tree.B;
    if (scanned < 252.0e6)
        nn_topk_kernel<true><<<grid, NN_THREADS, 0, stream>>>(p);
    else
        nn_topk_kernel<false><<<grid, NN_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t mps_launch_nn_next(const NextParams& p, cudaStream_t stream)
{
    nn_next_kernel<<<(unsigned)p.nn.grid_x, NN_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t mps_launch_tree_append(const TreeView& t, long long first, long long n, const float* staged_states,
                                   const int32_t* staged_parents, const float* staged_controls,
                                   const int32_t* staged_slices, long long* size_out, cudaStream_t stream)
{
    int blocks = grid_for(n * 3 * t.B, 256, 1184);
    tree_append_kernel<<<blocks, 256, 0, stream>>>(t, first, n, staged_states, staged_parents, staged_controls,
                                                   staged_slices, size_out);
    return cudaGetLastError();
}

cudaError_t mps_launch_tree_fill(const TreeView& t, long long n, unsigned long long seed, int n_slices, float x_min,
                                 float x_max, float y_min, float y_max, long long* size_out, cudaStream_t stream)
{
    int blocks = grid_for(n, 256, 1184);
    tree_fill_kernel<<<blocks, 256, 0, stream>>>(t, n, seed, n_slices, x_min, x_max, y_min, y_max, size_out);
    return cudaGetLastError();
}

cudaError_t mps_launch_tree_gather(const TreeView& t, long long first, long long n, float* out_states,
                                   cudaStream_t stream)
{
    int blocks = grid_for(n * 3 * t.B, 256, 1184);
    tree_gather_kernel<<<blocks, 256, 0, stream>>>(t, first, n, out_states);
    return cudaGetLastError();
}

cudaError_t mps_launch_tree_copy(const TreeView& src, const TreeView& dst, long long n, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    int blocks = grid_for(n * 3 * src.B, 256, 1184);
    tree_copy_kernel<<<blocks, 256, 0, stream>>>(src, dst, n);
    return cudaGetLastError();
}

cudaError_t mps_launch_forest_roots(const TreeView& forest, size_t seg, int n_inst, const float* roots, long long* sizes,
                                    cudaStream_t stream)
{
    int blocks = grid_for((long long)n_inst * 3 * forest.B, 256, 1184);
    forest_roots_kernel<<<blocks, 256, 0, stream>>>(forest, seg, n_inst, roots, sizes);
    return cudaGetLastError();
}
#else
}  // namespace (the launchers above are not part of the emulator build)
#endif  // MPS_CPU_EMU
