// Blocked sparse x shared-memory-resident dense tile products: the two X-streaming passes of an ALS
// iteration.
//
//   pass 1 (H-update right-hand side):  R[cell,:]  += X[cell, gene block] * Mt[gene block, :]
//   pass 2 (statistics):                St[gene,:] += X[cell block, gene]' * H[cell block, :]   (+ HtH)
//
// Both are "sparse row times a dense (rows x KP) tile" with k <= 64 columns, so they are bound by the
// shared-memory operand bandwidth (4*KP bytes per non-zero at 128 B/clk/SM), not by tensor math: the
// dense tile is brought in ONCE per CTA with a TMA bulk copy (cp.async.bulk, mbarrier completion), the
// packed (index, value) pairs of X are streamed from HBM exactly once per pass with 64/128-bit loads,
// each lane owns 4 (fp32) / 2 (fp64) consecutive factors and reads them with one 128-bit LDS, and
// accumulation is in registers.  X is stored twice in blocked form (gene-blocked CSR for pass 1,
// cell-blocked CSC for pass 2) so that in both passes the gathered operand is the shared tile.
#pragma once
#include "inmf_common.cuh"

#define INMF_WORK_FIELDS 8
// work[w] = { ptr_base, dense_row0, tile_rows, row_begin, row_end, out_row_base, gram_seg (-1: none), unused }

template <typename T>
struct PairTraits;
template <>
struct PairTraits<float> {
    typedef int2 Pair;  // {idx, float bits}
    typedef float4 Vec;
    static constexpr int VEC = 4;
    __device__ static __forceinline__ void unpack(const Pair& p, int& idx, float& v) {
        idx = p.x;
        v = __int_as_float(p.y);
    }
    __device__ static __forceinline__ Pair zero() { return make_int2(0, 0); }
};
template <>
struct PairTraits<double> {
    typedef longlong2 Pair;  // {idx (low 32 bits), double bits}
    typedef double2 Vec;
    static constexpr int VEC = 2;
    __device__ static __forceinline__ void unpack(const Pair& p, int& idx, double& v) {
        idx = (int)p.x;
        v = __longlong_as_double(p.y);
    }
    __device__ static __forceinline__ Pair zero() { return make_longlong2(0, 0); }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    }
}
// TMA 1-D bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <typename T>
__device__ __forceinline__ typename PairTraits<T>::Pair ld_pair(const typename PairTraits<T>::Pair* p);
template <>
__device__ __forceinline__ int2 ld_pair<float>(const int2* p) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
template <>
__device__ __forceinline__ longlong2 ld_pair<double>(const longlong2* p) {
    longlong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}

#define BLK_UNROLL 4

// One warp: out[0..KP) = sum over pairs [beg, end) of val * tile[idx][:].  Lane (grp, sub): grp = which of
// the NPI pairs in flight, sub = which VEC-wide factor chunk.  Result (all groups summed) is returned in
// acc[] of the lanes with grp == 0.
template <typename T>
__device__ __forceinline__ void warp_sparse_dot(const typename PairTraits<T>::Pair* __restrict__ pairs, int64_t beg,
                                                int64_t end, const T* __restrict__ tile, int KP, int lpn, int npi,
                                                int grp, int sub, bool active, T (&acc)[PairTraits<T>::VEC]) {
    typedef PairTraits<T> PT;
    typedef typename PT::Pair Pair;
    typedef typename PT::Vec Vec;
    constexpr int VEC = PT::VEC;
#pragma unroll
    for (int u = 0; u < VEC; ++u) acc[u] = T(0);
    const T* __restrict__ tsub = tile + sub * VEC;
    const int64_t step = (int64_t)npi * BLK_UNROLL;
    // software pipeline: the loads of batch i+1 are in flight while batch i is consumed
    Pair cur[BLK_UNROLL], nxt[BLK_UNROLL];
#pragma unroll
    for (int t = 0; t < BLK_UNROLL; ++t) {
        const int64_t q = beg + (int64_t)t * npi + grp;
        cur[t] = (active && q < end) ? ld_pair<T>(pairs + q) : PT::zero();
    }
    for (int64_t base = beg; base < end; base += step) {  // warp-uniform trip count
#pragma unroll
        for (int t = 0; t < BLK_UNROLL; ++t) {
            const int64_t q = base + step + (int64_t)t * npi + grp;
            nxt[t] = (active && q < end) ? ld_pair<T>(pairs + q) : PT::zero();
        }
#pragma unroll
        for (int t = 0; t < BLK_UNROLL; ++t) {
            int idx;
            T v;
            PT::unpack(cur[t], idx, v);
            const Vec m = *reinterpret_cast<const Vec*>(tsub + (size_t)idx * KP);
            const T* pm = reinterpret_cast<const T*>(&m);
#pragma unroll
            for (int u = 0; u < VEC; ++u) acc[u] = fma(v, pm[u], acc[u]);
        }
#pragma unroll
        for (int t = 0; t < BLK_UNROLL; ++t) cur[t] = nxt[t];
    }
    // fold the npi groups onto group 0
    for (int gq = 1; gq < npi; ++gq) {
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
            const T o = __shfl_down_sync(INMF_FULL_MASK, acc[u], gq * lpn);
            if (grp == 0) acc[u] += o;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(1024, 1)
blocked_spmm_kernel(const int64_t* __restrict__ work, const int64_t* __restrict__ ptr,
                    const typename PairTraits<T>::Pair* __restrict__ pairs, const T* __restrict__ dense, int KP,
                    T* __restrict__ out, int64_t ldo, int k, double* __restrict__ gram_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    T* tile = reinterpret_cast<T*>(smem_raw + 128);
    const int64_t* wd = work + (size_t)blockIdx.x * INMF_WORK_FIELDS;
    const int64_t ptr_base = wd[0], dense_row0 = wd[1];
    const int tile_rows = (int)wd[2];
    const int64_t row_begin = wd[3], row_end = wd[4], out_row_base = wd[5];
    const int gram_seg = (int)wd[6];
    constexpr int VEC = PairTraits<T>::VEC;

    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t total = (size_t)tile_rows * KP * sizeof(T);
        mbar_expect_tx(bar, (uint32_t)total);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(dense + (size_t)dense_row0 * KP);
        unsigned char* dst = reinterpret_cast<unsigned char*>(tile);
        for (size_t off = 0; off < total; off += 32768) {
            const size_t n = (total - off < 32768) ? (total - off) : 32768;
            bulk_g2s(dst + off, src + off, (uint32_t)n, bar);
        }
    }
    mbar_wait(bar, 0);

    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lpn = KP / VEC;         // lanes per non-zero
    const int npi = 32 / lpn;         // non-zeros in flight per warp instruction
    const int grp = lane / lpn, sub = lane - grp * lpn;
    const bool active = grp < npi;
    int64_t r = row_begin + warp;
    int64_t beg = 0, end = 0;
    if (r < row_end) {
        beg = ptr[ptr_base + r];
        end = ptr[ptr_base + r + 1];
    }
    while (r < row_end) {
        // prefetch the next row's extent while this one is processed
        const int64_t rn = r + nwarps;
        int64_t nbeg = 0, nend = 0;
        if (rn < row_end) {
            nbeg = ptr[ptr_base + rn];
            nend = ptr[ptr_base + rn + 1];
        }
        if (beg != end) {
            T acc[VEC];
            warp_sparse_dot<T>(pairs, beg, end, tile, KP, lpn, npi, active ? grp : 0, active ? sub : 0, active, acc);
            if (grp == 0) {
                T* o = out + (out_row_base + r) * ldo + sub * VEC;
#pragma unroll
                for (int u = 0; u < VEC; ++u)
                    if (sub * VEC + u < k && acc[u] != T(0)) atomicAdd(o + u, acc[u]);
            }
        }
        r = rn;
        beg = nbeg;
        end = nend;
    }
    if (gram_seg >= 0 && gram_out != nullptr) {
        // H'H of this cell block straight from the resident tile
        tile_gram_accumulate<T>(tile, KP, tile_rows, k, gram_out + (size_t)gram_seg * k * k, 1.0);
    }
}
