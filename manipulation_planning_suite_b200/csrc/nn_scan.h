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
};

int mps_nn_grid_x(long long n, int num_sms);
cudaError_t mps_launch_nn_scan(const NNParams& p, cudaStream_t stream);
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
