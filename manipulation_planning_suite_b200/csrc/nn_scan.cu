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
#define NN_REL_EPS 1.0e-4f
#define NN_ABS_EPS 1.0e-15f
#define NN_CAND_CAP 512

namespace {

__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float body_term_f32(float x, float y, float th, float qx, float qy, float qth, float pwf,
                                               float owf)
{
    float dx = x - qx;
    float dy = y - qy;
    float d = fabsf(th - qth);
    float arc = d > MPS_PI_F ? MPS_TWO_PI_F - d : d;
    return pwf * sqrt_approx(dx * dx + dy * dy) + owf * arc;
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
    const int q = blockIdx.y;
    const int B = p.tree.B;
    // the query, its mask / slice filter and the weights come from device buffers, or (single query through
    // mps_tree_nearest) from the kernel parameters themselves
    const bool inl = p.inline_q != 0;
    uint32_t mask = inl ? p.inline_mask : (p.masks ? p.masks[q] : 0xffffffffu);
    if (B < 32) mask &= (1u << B) - 1u;
    const int filter = inl ? p.inline_filter : (p.filters ? p.filters[q] : -1);
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
    for (long long g = (long long)blockIdx.x * NN_THREADS + tid; g < groups; g += (long long)gridDim.x * NN_THREADS) {
        const long long n0 = g << 2;
        float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
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
        // Tighten the CTA bound with this warp's fp32 minimum BEFORE testing against it.
        const float dm = fminf(fminf(d0, d1), fminf(d2, d3));
        const unsigned am = __activemask();
        const unsigned wmin = __reduce_min_sync(am, __float_as_uint(dm));  // d >= 0: uint order == float order
        if (lane == (__ffs(am) - 1)) atomicMin(&s_bound, wmin);
        __syncwarp(am);
        const float b = fminf(__uint_as_float(*reinterpret_cast<volatile unsigned int*>(&s_bound)),
                              __uint_as_float(wmin));
        const float thr = b * (1.0f + NN_REL_EPS) + NN_ABS_EPS;
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

    // ---- deferred exact pass: only candidates still within the final CTA bound ---------------------
    {
        const float bf = __uint_as_float(s_bound);
        const float thr = bf * (1.0f + NN_REL_EPS) + NN_ABS_EPS;
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
    if (tid < 32) {
        best_d = tid < NN_THREADS / 32 ? s_rd[tid] : 0.0;
        best_i = tid < NN_THREADS / 32 ? s_ri[tid] : -1;
        for (int off = 16; off > 0; off >>= 1) {
            double od = __shfl_down_sync(FULL, best_d, off);
            long long oi = __shfl_down_sync(FULL, best_i, off);
            take_min(best_d, best_i, od, oi);
        }
        if (tid == 0 && gridDim.x == 1) {
            // small trees: one CTA, no cross-CTA hand-off (saves ~4 dependent global round trips)
            p.out_idx[q] = best_i;
            p.out_dist[q] = best_i < 0 ? __longlong_as_double(0x7ff0000000000000ll) : best_d;
            s_last = false;
        } else if (tid == 0) {
            p.part_d[(size_t)q * p.grid_x + blockIdx.x] = best_d;
            p.part_i[(size_t)q * p.grid_x + blockIdx.x] = best_i;
            __threadfence();
            unsigned int t = atomicAdd(&p.tickets[q], 1u);
            s_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (s_last && tid < 32) {
        __threadfence();
        double fd = 0.0;
        long long fi = -1;
        for (int c = tid; c < (int)gridDim.x; c += 32) {
            double od = *reinterpret_cast<volatile double*>(&p.part_d[(size_t)q * p.grid_x + c]);
            long long oi = *reinterpret_cast<volatile long long*>(&p.part_i[(size_t)q * p.grid_x + c]);
            take_min(fd, fi, od, oi);
        }
        for (int off = 16; off > 0; off >>= 1) {
            double od = __shfl_down_sync(FULL, fd, off);
            long long oi = __shfl_down_sync(FULL, fi, off);
            take_min(fd, fi, od, oi);
        }
        if (tid == 0) {
            p.out_idx[q] = fi;
            p.out_dist[q] = fi < 0 ? __longlong_as_double(0x7ff0000000000000ll) : fd;
            p.tickets[q] = 0u;
        }
    }
}

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
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

cudaError_t mps_launch_nn_scan(const NNParams& p, cudaStream_t stream)
{
    dim3 grid((unsigned)p.grid_x, (unsigned)p.Q, 1);
    // evict-first loads while the scanned rows are within ~2x the 126 MB L2, default policy beyond
    const double scanned = (double)p.n * 12.0 * p.tree.B;
    if (scanned < 252.0e6)
        nn_scan_kernel<true><<<grid, NN_THREADS, 0, stream>>>(p);
    else
        nn_scan_kernel<false><<<grid, NN_THREADS, 0, stream>>>(p);
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
