// cuda_emu.h — a minimal SIMT emulator for the repo's own .cu sources.
//
// TEST INFRASTRUCTURE.  Lives under tests/, is built into tests/emu/libmps_emu.so and is loaded by the
// `-m "not gpu"` tests only, to check the CUDA kernels' SOURCE against the CPU oracle on a machine without
// a GPU (the development container has none).  It is not a CPU fallback: nothing under
// manipulation_planning_suite_b200/ builds, links, loads or knows about it, and libmps_b200.so fails
// with MPS_ERR_NO_DEVICE when there is no device.
//
// How: every CUDA thread of a CTA is a fiber (own stack, cooperative switch, one OS thread); CTAs run one
// after the other.  Warp collectives (`__shfl_sync`, `__ballot_sync`, votes, `__reduce_*_sync`,
// `__syncwarp`) are rendezvous points keyed by their member mask, so lane groups of one warp may be in
// different places of the code exactly as on the device; `__syncthreads` is a CTA-wide rendezvous.
// Arithmetic: the including translation unit is compiled with -ffp-contract=off -mfma, so every +,-,*,/,
// sqrtf is one IEEE operation and fmaf is a fused multiply-add, as in the device build (-fmad=false).
// Not emulated: `__activemask` (kernels under test name their masks), memory-model races (fibers switch
// only at collectives), anything asynchronous (TMA, cp.async, clusters).
#pragma once

#include <cuda_runtime.h>  // vector types, cudaError_t, cudaStream_t (host-compiler view of the headers)

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#ifndef __x86_64__
#error "tests/emu: the fiber switch is written for x86-64"
#endif

#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __noinline__
#undef __launch_bounds__
#undef __shared__
#undef __align__
#undef __grid_constant__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define __grid_constant__

namespace emu {

struct Group {
    uint32_t mask = 0, arrived = 0, left = 0;
    bool leaving = false;
    uint64_t vals[32];
};

struct Warp {
    Group groups[8];
    int n_groups = 0;
};

struct Fiber {
    void* sp = nullptr;
    unsigned char* stack = nullptr;
    uint3 tid;
    int lane = 0;
    Warp* warp = nullptr;
    bool done = false;
};

extern Fiber* cur;
extern dim3 g_block_idx, g_block_dim, g_grid_dim;
extern unsigned char* g_smem;

void yield();
void gather(uint32_t mask, uint64_t v, uint64_t out[32]);
void cta_barrier();
// runs body() once per thread of every CTA of the grid
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);

inline unsigned char* dyn_smem() { return g_smem; }

template <class T>
inline uint64_t to_bits(T v)
{
    static_assert(sizeof(T) <= 8, "collective payload too wide");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T>
inline T from_bits(uint64_t b)
{
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::g_block_idx)
#define blockDim (emu::g_block_dim)
#define gridDim (emu::g_grid_dim)

// ---- warp collectives -------------------------------------------------------------------------
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    (void)width;
    uint64_t out[32];
    emu::gather(mask, emu::to_bits(v), out);
    src &= 31;
    if (!((mask >> src) & 1u)) return v;  // undefined on the device; keep it deterministic here
    return emu::from_bits<T>(out[src]);
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    (void)width;
    uint64_t out[32];
    emu::gather(mask, emu::to_bits(v), out);
    const int src = emu::cur->lane + (int)delta;
    if (src > 31 || !((mask >> src) & 1u)) return v;
    return emu::from_bits<T>(out[src]);
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    (void)width;
    uint64_t out[32];
    emu::gather(mask, emu::to_bits(v), out);
    const int src = emu::cur->lane - (int)delta;
    if (src < 0 || !((mask >> src) & 1u)) return v;
    return emu::from_bits<T>(out[src]);
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32)
{
    (void)width;
    uint64_t out[32];
    emu::gather(mask, emu::to_bits(v), out);
    const int src = emu::cur->lane ^ lane_mask;
    if (src > 31 || !((mask >> src) & 1u)) return v;
    return emu::from_bits<T>(out[src]);
}
inline unsigned __ballot_sync(unsigned mask, int pred)
{
    uint64_t out[32];
    emu::gather(mask, pred ? 1u : 0u, out);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if (((mask >> i) & 1u) && out[i]) r |= 1u << i;
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
inline void __syncwarp(unsigned mask = 0xffffffffu)
{
    uint64_t out[32];
    emu::gather(mask, 0, out);
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v)
{
    uint64_t out[32];
    emu::gather(mask, v, out);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if ((mask >> i) & 1u) r |= (unsigned)out[i];
    return r;
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v)
{
    uint64_t out[32];
    emu::gather(mask, v, out);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if (((mask >> i) & 1u) && (unsigned)out[i] > r) r = (unsigned)out[i];
    return r;
}
inline int __reduce_max_sync(unsigned mask, int v)
{
    uint64_t out[32];
    emu::gather(mask, (uint64_t)(uint32_t)v, out);
    int r = INT32_MIN;
    for (int i = 0; i < 32; ++i)
        if (((mask >> i) & 1u) && (int)(uint32_t)out[i] > r) r = (int)(uint32_t)out[i];
    return r;
}
inline unsigned __reduce_min_sync(unsigned mask, unsigned v)
{
    uint64_t out[32];
    emu::gather(mask, v, out);
    unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; ++i)
        if (((mask >> i) & 1u) && (unsigned)out[i] < r) r = (unsigned)out[i];
    return r;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v)
{
    uint64_t out[32];
    emu::gather(mask, v, out);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if ((mask >> i) & 1u) r += (unsigned)out[i];
    return r;
}
inline void __syncthreads() { emu::cta_barrier(); }
inline void __threadfence() {}
inline void __threadfence_system() {}

// ---- scalar intrinsics ------------------------------------------------------------------------
inline unsigned __float_as_uint(float f) { return emu::from_bits<unsigned>(emu::to_bits(f)); }
inline int __float_as_int(float f) { return emu::from_bits<int>(emu::to_bits(f)); }
inline float __uint_as_float(unsigned u) { return emu::from_bits<float>(emu::to_bits(u)); }
inline float __int_as_float(int u) { return emu::from_bits<float>(emu::to_bits(u)); }
inline double __longlong_as_double(long long v) { return emu::from_bits<double>(emu::to_bits(v)); }
inline long long __double_as_longlong(double v) { return emu::from_bits<long long>(emu::to_bits(v)); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline long long clock64() { return 0; }
template <class T>
inline T __ldg(const T* p) { return *p; }
template <class T>
inline T __ldcg(const T* p) { return *p; }
template <class T>
inline T __ldcs(const T* p) { return *p; }
template <class T>
inline T atomicAdd(T* p, T v)
{
    T o = *p;
    *p = o + v;
    return o;
}
template <class T>
inline T atomicMin(T* p, T v)
{
    T o = *p;
    if (v < o) *p = v;
    return o;
}
template <class T>
inline T atomicMax(T* p, T v)
{
    T o = *p;
    if (v > o) *p = v;
    return o;
}
template <class T>
inline T atomicExch(T* p, T v)
{
    T o = *p;
    *p = v;
    return o;
}
template <class T>
inline T atomicCAS(T* p, T cmp, T v)
{
    T o = *p;
    if (o == cmp) *p = v;
    return o;
}
template <class T>
inline T min(T a, T b) { return b < a ? b : a; }
template <class T>
inline T max(T a, T b) { return a < b ? b : a; }
inline unsigned min(unsigned a, int b) { return (unsigned)b < a ? (unsigned)b : a; }
inline void __sincosf(float a, float* s, float* c)
{
    *s = sinf(a);
    *c = cosf(a);
}
