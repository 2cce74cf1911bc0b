// Shared helpers for the iNMF kernels (sm_100a).  See include/inmf_b200.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "inmf_b200.h"

#define INMF_FULL_MASK 0xffffffffu

// Checked build (python -m pyliger_b200.build -DINMF_CHECKED): device-side bounds checks on the data-dependent
// addresses of the gather kernels and the work lists; a failed check prints its location and traps the kernel.
#ifdef INMF_CHECKED
#define INMF_CHECK(cond)                                                                          \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            printf("INMF_CHECK failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, \
                   (int)blockIdx.x, (int)threadIdx.x);                                            \
            __trap();                                                                             \
        }                                                                                         \
    } while (0)
#else
#define INMF_CHECK(cond) ((void)0)
#endif

extern thread_local char g_inmf_err[512];

#define INMF_CUDA_CHECK(call)                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            snprintf(g_inmf_err, sizeof(g_inmf_err), "%s:%d: %s -> %s", __FILE__, __LINE__, #call, \
                     cudaGetErrorString(e_));                                                   \
            return -1;                                                                          \
        }                                                                                       \
    } while (0)

#define INMF_REQUIRE(cond, msg)                                                                 \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            snprintf(g_inmf_err, sizeof(g_inmf_err), "%s:%d: invalid argument: %s", __FILE__,     \
                     __LINE__, msg);                                                            \
            return -2;                                                                          \
        }                                                                                       \
    } while (0)

extern long long g_inmf_launches;  // kernels launched by this library (host-side count)
#define INMF_LAUNCH_CHECK()                      \
    do {                                         \
        __atomic_add_fetch(&g_inmf_launches, 1, __ATOMIC_RELAXED); \
        INMF_CUDA_CHECK(cudaGetLastError());     \
    } while (0)

#define INMF_MAX_DEVICES 64
static inline int inmf_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= INMF_MAX_DEVICES) dev = 0;
    return dev;
}

// SM count of the CURRENT device (a process may drive several GPUs through the Python API's device= argument).
static inline int inmf_num_sms() {
    static int sms[INMF_MAX_DEVICES];
    const int dev = inmf_device();
    if (sms[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        sms[dev] = n;
    }
    return sms[dev];
}

// Opt in to more than 48 KB of dynamic shared memory, once per (kernel instantiation, device).
#define INMF_OPTIN_SMEM(kernel, bytes)                                                                     \
    do {                                                                                                   \
        static bool done_[INMF_MAX_DEVICES];                                                               \
        const int d_ = inmf_device();                                                                      \
        if (!done_[d_]) {                                                                                  \
            INMF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); \
            done_[d_] = true;                                                                              \
        }                                                                                                  \
    } while (0)

template <typename T>
__device__ __forceinline__ T inmf_rsqrt(T x);
template <>
__device__ __forceinline__ float inmf_rsqrt<float>(float x) {
    float y;  // MUFU.RSQ without the denormal pre-scaling sequence (Gram pivots are never denormal)
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <>
__device__ __forceinline__ double inmf_rsqrt<double>(double x) {
    return 1.0 / sqrt(x);
}

template <typename T>
__device__ __forceinline__ T inmf_abs(T x) {
    return x < T(0) ? -x : x;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(INMF_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(INMF_FULL_MASK, v, o);
    return v;
}

// Segment window descriptor (see header): seg_desc[s] = {out_off, row_base, win_start, win_mod}.
struct SegWin {
    int64_t out_off, row_base, win_start, win_mod, nrows;
};
__device__ __forceinline__ SegWin load_seg(const int64_t* __restrict__ seg_desc, int s) {
    SegWin w;
    w.out_off = seg_desc[4 * s + 0];
    w.row_base = seg_desc[4 * s + 1];
    w.win_start = seg_desc[4 * s + 2];
    w.win_mod = seg_desc[4 * s + 3];
    w.nrows = seg_desc[4 * (s + 1)] - w.out_off;
    return w;
}
__device__ __forceinline__ int64_t seg_row(const SegWin& w, int64_t j) {
    int64_t r = w.win_start + j;
    if (r >= w.win_mod) r %= w.win_mod;
    return w.row_base + r;
}

// =================================================================================================
// Tall-skinny Gram helper: out(k x k, double) += tile' * tile for a row tile staged in shared memory.
// Each thread owns a 4x4 block of the output.
// =================================================================================================
#define GRAM_TILE_ROWS 64

// Sum into a MULTICAST address: the add is performed on the copy of every GPU bound to the multicast object (NVLink
// switch reduction, PTX multimem.red); the caller provides the cross-GPU ordering (barriers around the kernel).
__device__ __forceinline__ void mc_red_add_f64(double* p, double v) {
    asm volatile("multimem.red.relaxed.sys.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
#define INMF_GRAM_ATOMIC 0
#define INMF_GRAM_PLAIN 1
#define INMF_GRAM_MULTICAST 2

// mode = INMF_GRAM_PLAIN: `out` belongs to this CTA alone (a partial buffer of the deterministic mode): every entry is
// owned by one thread, which adds to it without an atomic.  INMF_GRAM_MULTICAST: `out` is a multicast address.
template <typename T>
__device__ __forceinline__ void tile_gram_accumulate(const T* __restrict__ tile, int ldt, int rows, int k,
                                                     double* __restrict__ out, double scale, int mode = INMF_GRAM_ATOMIC) {
    const int nb = (k + 3) >> 2;
    for (int blk = threadIdx.x; blk < nb * nb; blk += blockDim.x) {
        const int bi = blk / nb, bj = blk % nb;
        T acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = T(0);
        for (int r = 0; r < rows; ++r) {
            const T* row = tile + (size_t)r * ldt;
            T va[4], vb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                va[a] = row[4 * bi + a];
                vb[a] = row[4 * bj + a];
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] += va[a] * vb[b];
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int i = 4 * bi + a, j = 4 * bj + b;
                if (i < k && j < k) {
                    if (mode == INMF_GRAM_PLAIN) out[(size_t)i * k + j] += scale * (double)acc[a][b];
                    else if (mode == INMF_GRAM_MULTICAST) mc_red_add_f64(&out[(size_t)i * k + j], scale * (double)acc[a][b]);
                    else atomicAdd(&out[(size_t)i * k + j], scale * (double)acc[a][b]);
                }
            }
    }
}

