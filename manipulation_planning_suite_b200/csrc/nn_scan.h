// nn_scan.h — launch interface of nn_scan.cu (internal to libmps_b200.so)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// Device-resident tree, dimension-major SoA: row r = 3*body + {0:x, 1:y, 2:theta}
// holds that coordinate of every node, rows are `cap` floats apart (cap % 4 == 0, base 256 B aligned),
// so four consecutive nodes of one row are one aligned 16-byte load.
struct TreeView {
    float* states;    // [3B][cap]
    float* controls;  // [5][cap]
    int32_t* parent;  // [cap]
    int32_t* slice;   // [cap]
    size_t cap;
    int B;
};

struct NNParams {
    TreeView tree;
    long long n;            // nodes to scan
    const float* queries;   // [Q][3B]
    const uint32_t* masks;  // [Q] or nullptr (all bodies)
    const int32_t* filters; // [Q] or nullptr (no slice filter)
    const float* weights;   // [B] or nullptr (1.0)
    double pw, ow;
    int Q;
    long long* out_idx;  // [Q]
    double* out_dist;    // [Q]
    double* part_d;      // [Q][grid_x]
    long long* part_i;   // [Q][grid_x]
    unsigned int* tickets;  // [Q], zero on entry, zero on exit
    // forest queries (multi-instance): query q scans segment inst_ids[q] = columns [inst * seg, inst * seg + seg_sizes[inst])
    const int32_t* inst_ids;      // [Q] or nullptr (plain tree: columns [0, n))
    const long long* seg_sizes;   // [segments]
    size_t seg;
    int grid_x;
    // single query carried in the kernel parameters (no host -> device copy on the query path): used when
    // inline_q != 0, Q == 1; queries / masks / filters / weights above are then ignored
    int inline_q;
    int inline_has_w;
    uint32_t inline_mask;
    int32_t inline_filter;
    float q_inline[3 * 32];
    float w_inline[32];
    // chained queries (SliceBasedOracleRRT's two-level search): when set, the slice filter of an inline query is
    // the node index another scan wrote to device memory (index of the nearest slice representative = slice id)
    const long long* filter_from;
    // completion flag for the single-query path: after (index, distance) the last CTA writes `done_seq` here
    // (pinned host memory) so that the host can wait for the result without a stream synchronise
    volatile unsigned int* done_flag;
    unsigned int done_seq;
    // host-folded single query: per-CTA records {distance bits, seq << 40 | index} in pinned host memory, 2 words per
    // CTA (see nn_scan_kernel); when set, out_idx / out_dist / tickets / done_flag are not used
    unsigned long long* host_parts;
    unsigned int host_seq;
    // optional second copy of (index, distance) of query 0, e.g. in pinned host memory for a scan whose result also
    // feeds the next scan on the device (device-side hand-off only)
    long long* mirror_idx;
    double* mirror_dist;
    float abs_eps;  // absolute slack of the fp32 candidate filter (covers the rounding of the angle wrap)
    // programmatic dependent launch: 1 = the previous kernel of the stream is a scan of the same context and nothing
    // it writes is read by this one before the hand-off (the streaming pass may overlap its tail); 0 = wait for the
    // previous kernel before anything else (always safe; mps_launch_nn_scan forces it for chained filters)
    int overlap_prev;
};

// exact k nearest neighbours of one query (k <= 32), see nn_topk_kernel
#define NN_TOPK_MAX 32
#define NN_TOPK_CAND 64
#define NN_TOPK_HOT 512   // lists the last CTA looks into (grid <= 2 CTAs x 148 SMs)
#define NN_TOPK_BUF 1024  // keys within the bound of the k-th smallest list head (k hot lists x k keys at most, without ties)
struct TopKParams {
    NNParams nn;                    // tree, n, the query (inline), pw / ow, grid_x, tickets
    int k;
    unsigned long long* part_keys;  // [grid_x][k] per-CTA smallest keys (fp32 distance bits << 32 | node index)
    unsigned int* part_rej;         // [grid_x] fp32 bits of the smallest distance the CTA dropped
    long long* out_idx;             // [k] ascending by (double distance, index)
    double* out_dist;               // [k]
    int* out_meta;                  // [0] results written (<= k), [1] status: 0 exact, 1 ambiguous (caller falls back)
    // results in pinned host memory: the last CTA writes done_seq here behind them (system-scope fence), the host waits
    // for the number instead of synchronising the stream
    volatile unsigned int* done_flag;
    unsigned int done_seq;
    // or (out_rec != nullptr; out_idx / out_dist / out_meta / done_flag unused): self-signed 16-byte records in pinned host
    // memory, one store each - [r] = {distance bits, done_seq << 40 | index} for result r, [NN_TOPK_MAX] = {results
    // written | status << 32, done_seq << 40}; a record that carries the sequence number is complete, so nothing has to
    // be fenced (the system-scope fence before the flag cost ~2 us of the last CTA's tail)
    unsigned long long* out_rec;
};

// lexicographic successor scan in double (rare fallback of the top-k path, and its checker):
// min over nodes of (distance, index) > (after_d, after_i)
struct NextParams {
    NNParams nn;
    double after_d;
    long long after_i;
};

int mps_nn_grid_x(long long n, int num_sms);
cudaError_t mps_launch_nn_scan(const NNParams& p, cudaStream_t stream);
cudaError_t mps_launch_nn_topk(const TopKParams& p, cudaStream_t stream);
cudaError_t mps_launch_nn_next(const NextParams& p, cudaStream_t stream);
cudaError_t mps_launch_tree_append(const TreeView& t, long long first, long long n, const float* staged_states,
                                   const int32_t* staged_parents, const float* staged_controls,
                                   const int32_t* staged_slices, long long* size_out, cudaStream_t stream);
cudaError_t mps_launch_tree_fill(const TreeView& t, long long n, unsigned long long seed, int n_slices, float x_min,
                                 float x_max, float y_min, float y_max, long long* size_out, cudaStream_t stream);
cudaError_t mps_launch_tree_gather(const TreeView& t, long long first, long long n, float* out_states,
                                   cudaStream_t stream);
cudaError_t mps_launch_tree_copy(const TreeView& src, const TreeView& dst, long long n, cudaStream_t stream);
cudaError_t mps_launch_forest_roots(const TreeView& forest, size_t seg, int n_inst, const float* roots, long long* sizes,
                                    cudaStream_t stream);
