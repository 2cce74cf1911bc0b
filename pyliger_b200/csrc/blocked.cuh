// Blocked sparse x shared-memory-resident dense tile products: the two X-streaming passes of an ALS
// iteration.
//
//   pass 1 (H-update right-hand side):  R[cell,:]  += X[cell, gene block] * Mt[gene block, :]
//   pass 2 (statistics):                St[gene,:] += X[cell block, gene]' * H[cell block, :]   (+ HtH)
//
// Both are "sparse row times a dense (rows x KP) tile" with k <= 64 columns, so they are bound by the
// shared-memory operand bandwidth (4*KP bytes per non-zero at 128 B/clk/SM), not by tensor math: the
// dense tile is brought in ONCE per CTA with a TMA bulk copy (cp.async.bulk, mbarrier completion), the
// packed (offset, value) pairs of X are streamed from HBM exactly once per pass with 64/128-bit loads,
// each lane owns 4 (fp32) / 2 (fp64) consecutive factors and reads them with one 128-bit LDS, and
// accumulation is in registers.  X is stored twice in blocked form (gene-blocked CSR for pass 1,
// cell-blocked CSC for pass 2) so that in both passes the gathered operand is the shared tile.
//
// Pair layout: f32 {int32 off, float val}, f64 {int64 off, double val}; off = (row of the tile) * KP *
// sizeof(T) is the BYTE offset of the gathered tile row, premultiplied when the format is built so the
// inner loop is LDG.64 -> IADD -> LDS.128 -> 4 FFMA per lane group.
#pragma once
#include "inmf_common.cuh"

#define INMF_WORK_FIELDS 8
// work[w] = { ptr_base, dense_row0, tile_rows, row_begin, row_end, out_row_base, gram_seg (-1: none), unused }

template <typename T>
struct PairTraits;
template <>
struct PairTraits<float> {
    typedef int2 Pair;
    static constexpr int VEC = 4;
    __device__ static __forceinline__ Pair zero() { return make_int2(0, 0); }
    __device__ static __forceinline__ int bits(float v) { return __float_as_int(v); }
    __device__ static __forceinline__ Pair ld(const Pair* p) {
        int2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
        return r;
    }
    // acc[0..3] += val * tile_row[sub*4 .. sub*4+3]
    __device__ static __forceinline__ void fma_row(const Pair& p, uint32_t tsub, float (&acc)[4]) {
        float m0, m1, m2, m3;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(m0), "=f"(m1), "=f"(m2), "=f"(m3)
                     : "r"(tsub + (uint32_t)p.x));
        const float v = __int_as_float(p.y);
        acc[0] = fmaf(v, m0, acc[0]);
        acc[1] = fmaf(v, m1, acc[1]);
        acc[2] = fmaf(v, m2, acc[2]);
        acc[3] = fmaf(v, m3, acc[3]);
    }
};
template <>
struct PairTraits<double> {
    typedef longlong2 Pair;
    static constexpr int VEC = 2;
    __device__ static __forceinline__ Pair zero() { return make_longlong2(0, 0); }
    __device__ static __forceinline__ long long bits(double v) { return __double_as_longlong(v); }
    __device__ static __forceinline__ Pair ld(const Pair* p) {
        longlong2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
        return r;
    }
    __device__ static __forceinline__ void fma_row(const Pair& p, uint32_t tsub, double (&acc)[2]) {
        double m0, m1;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(m0), "=d"(m1) : "r"(tsub + (uint32_t)p.x));
        const double v = __longlong_as_double(p.y);
        acc[0] = fma(v, m0, acc[0]);
        acc[1] = fma(v, m1, acc[1]);
    }
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

#define BLK_U 4  // quads (groups of npi non-zeros) per pipeline stage

// One warp, one sparse row of m pairs starting at pl0 (= pairs + beg).  Lane (grp, sub): grp = which of
// the npi pairs handled per warp instruction, sub = which VEC-wide factor chunk.  The sum over all groups
// ends up in acc[] of the lanes with grp == 0.  LPN_T > 0 fixes lanes-per-non-zero at compile time.
template <typename T, int LPN_T>
__device__ __forceinline__ void warp_row_dot(const typename PairTraits<T>::Pair* __restrict__ pl0, int m, int lpn_rt,
                                             int grp, bool active, uint32_t tsub,
                                             T (&acc)[PairTraits<T>::VEC]) {
    typedef PairTraits<T> PT;
    typedef typename PT::Pair Pair;
    constexpr int VEC = PT::VEC;
    const int lpn = LPN_T > 0 ? LPN_T : lpn_rt;
    const int npi = 32 / lpn;
#pragma unroll
    for (int u = 0; u < VEC; ++u) acc[u] = T(0);
    // lanes beyond npi*lpn shadow group 0 (their sums are never folded in)
    const Pair* __restrict__ pl = pl0 + (active ? grp : 0);
    const int nfull = m / npi;              // quads in which every group has a pair
    const int nq = (m + npi - 1) / npi;     // all quads
    int j = 0;
    if (nfull >= BLK_U) {
        Pair A[BLK_U], B[BLK_U];
#pragma unroll
        for (int t = 0; t < BLK_U; ++t) A[t] = PT::ld(pl + t * npi);
        // ping-pong: the loads of the next stage are in flight while the current one is consumed
        while (true) {
            const Pair* __restrict__ pn = pl + (size_t)(j + BLK_U) * npi;
            if (j + 2 * BLK_U <= nfull) {
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) B[t] = PT::ld(pn + t * npi);
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) PT::fma_row(A[t], tsub, acc);
                j += BLK_U;
            } else {
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) PT::fma_row(A[t], tsub, acc);
                j += BLK_U;
                break;
            }
            const Pair* __restrict__ pm = pl + (size_t)(j + BLK_U) * npi;
            if (j + 2 * BLK_U <= nfull) {
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) A[t] = PT::ld(pm + t * npi);
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) PT::fma_row(B[t], tsub, acc);
                j += BLK_U;
            } else {
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) PT::fma_row(B[t], tsub, acc);
                j += BLK_U;
                break;
            }
        }
    }
    // tail quads (predicated)
    {
        Pair C[BLK_U];  // at most (nfull mod BLK_U) + 1 <= BLK_U quads are left
#pragma unroll
        for (int t = 0; t < BLK_U; ++t) {
            const int q = (j + t) * npi + grp;
            C[t] = (active && (j + t) < nq && q < m) ? PT::ld(pl0 + q) : PT::zero();
        }
#pragma unroll
        for (int t = 0; t < BLK_U; ++t)
            if (j + t < nq) PT::fma_row(C[t], tsub, acc);
    }
    // fold the npi groups onto group 0
#pragma unroll 4
    for (int gq = 1; gq < npi; ++gq) {
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
            const T o = __shfl_down_sync(INMF_FULL_MASK, acc[u], gq * lpn);
            acc[u] += (grp == 0) ? o : T(0);
        }
    }
}

template <typename T, int LPN_T>
__global__ void __launch_bounds__(1024, 1)
blocked_spmm_kernel(const int64_t* __restrict__ work, const int64_t* __restrict__ ptr,
                    const typename PairTraits<T>::Pair* __restrict__ pairs, const T* __restrict__ dense, int KP,
                    T* __restrict__ out, int64_t ldo, int k, double* __restrict__ gram_out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
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
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lpn = LPN_T > 0 ? LPN_T : KP / VEC;  // lanes per non-zero
    const int npi = 32 / lpn;                      // non-zeros per warp instruction
    const int grp = lane / lpn, sub = lane - grp * lpn;
    const bool active = grp < npi;
    const uint32_t tsub = smem_u32(tile) + (uint32_t)((active ? sub : 0) * VEC * sizeof(T));
    int64_t r = row_begin + warp;
    int64_t beg = 0, end = 0;
    if (r < row_end) {  // overlaps the tile copy
        beg = ptr[ptr_base + r];
        end = ptr[ptr_base + r + 1];
    }
    mbar_wait(bar, 0);

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
            warp_row_dot<T, LPN_T>(pairs + beg, (int)(end - beg), lpn, grp, active, tsub, acc);
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

// =====================================================================================================
// Split-tile variant for 32 < KP <= 64 (fp32): the dense tile is held as a [rows x 32] part (128-byte rows) plus a
// [rows x RS] remainder (RS = 8, 16 or 32 floats per row, RL = RS / 8 per lane).
//
// Why: a 128-bit shared load is served in four phases of 8 lanes.  With KP = 40 (10 lanes per non-zero) a phase mixes
// the tail of one gathered row with the head of another, and one warp instruction (3 non-zeros) costs ~6 wavefronts
// instead of 4.  Here 8 lanes own one non-zero: every phase of the main load is one contiguous 128-byte row (4
// non-zeros per instruction, 4 wavefronts), and the remainder is one 32/64/128-bit load per lane.
// Pairs carry the MAIN byte offset (row * 128); the remainder offset is a shift of it.
// The tile arrives by per-row bulk copies (two per row) that all complete on one mbarrier.
// =====================================================================================================
template <int RL>
struct SplitRem;
template <>
struct SplitRem<1> {
    static constexpr int SHIFT = 2;  // row * 128 -> row * 32 bytes
    __device__ static __forceinline__ void fma(float v, uint32_t addr, float (&acc)[1]) {
        float m;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(m) : "r"(addr));
        acc[0] = fmaf(v, m, acc[0]);
    }
};
template <>
struct SplitRem<2> {
    static constexpr int SHIFT = 1;  // row * 64 bytes
    __device__ static __forceinline__ void fma(float v, uint32_t addr, float (&acc)[2]) {
        float m0, m1;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(m0), "=f"(m1) : "r"(addr));
        acc[0] = fmaf(v, m0, acc[0]);
        acc[1] = fmaf(v, m1, acc[1]);
    }
};
template <>
struct SplitRem<4> {
    static constexpr int SHIFT = 0;  // row * 128 bytes
    __device__ static __forceinline__ void fma(float v, uint32_t addr, float (&acc)[4]) {
        float m0, m1, m2, m3;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(m0), "=f"(m1), "=f"(m2), "=f"(m3) : "r"(addr));
        acc[0] = fmaf(v, m0, acc[0]);
        acc[1] = fmaf(v, m1, acc[1]);
        acc[2] = fmaf(v, m2, acc[2]);
        acc[3] = fmaf(v, m3, acc[3]);
    }
};

template <int RL>
__device__ __forceinline__ void split_fma(const int2& p, uint32_t tmain, uint32_t trem, float (&acc)[4],
                                          float (&accr)[RL]) {
    float m0, m1, m2, m3;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(m0), "=f"(m1), "=f"(m2), "=f"(m3)
                 : "r"(tmain + (uint32_t)p.x));
    const float v = __int_as_float(p.y);
    acc[0] = fmaf(v, m0, acc[0]);
    acc[1] = fmaf(v, m1, acc[1]);
    acc[2] = fmaf(v, m2, acc[2]);
    acc[3] = fmaf(v, m3, acc[3]);
    SplitRem<RL>::fma(v, trem + ((uint32_t)p.x >> SplitRem<RL>::SHIFT), accr);
}

// One warp, one sparse row of m pairs: lane (grp = lane / 8, sub = lane % 8); sums end up in the lanes of group 0.
template <int RL>
__device__ __forceinline__ void warp_row_dot_split(const int2* __restrict__ pl0, int m, int grp, uint32_t tmain,
                                                   uint32_t trem, float (&acc)[4], float (&accr)[RL]) {
    typedef PairTraits<float> PT;
    constexpr int NPI = 4;
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] = 0.f;
#pragma unroll
    for (int u = 0; u < RL; ++u) accr[u] = 0.f;
    const int2* __restrict__ pl = pl0 + grp;
    const int nfull = m / NPI;
    const int nq = (m + NPI - 1) / NPI;
    int j = 0;
    if (nfull >= BLK_U) {
        int2 A[BLK_U], B[BLK_U];
#pragma unroll
        for (int t = 0; t < BLK_U; ++t) A[t] = PT::ld(pl + t * NPI);
        while (true) {
            const int2* __restrict__ pn = pl + (size_t)(j + BLK_U) * NPI;
            if (j + 2 * BLK_U <= nfull) {
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) B[t] = PT::ld(pn + t * NPI);
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) split_fma<RL>(A[t], tmain, trem, acc, accr);
                j += BLK_U;
            } else {
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) split_fma<RL>(A[t], tmain, trem, acc, accr);
                j += BLK_U;
                break;
            }
            const int2* __restrict__ pm = pl + (size_t)(j + BLK_U) * NPI;
            if (j + 2 * BLK_U <= nfull) {
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) A[t] = PT::ld(pm + t * NPI);
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) split_fma<RL>(B[t], tmain, trem, acc, accr);
                j += BLK_U;
            } else {
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) split_fma<RL>(B[t], tmain, trem, acc, accr);
                j += BLK_U;
                break;
            }
        }
    }
    {
        int2 C[BLK_U];
#pragma unroll
        for (int t = 0; t < BLK_U; ++t) {
            const int q = (j + t) * NPI + grp;
            C[t] = ((j + t) < nq && q < m) ? PT::ld(pl0 + q) : PT::zero();  // zero pair: offset 0, value 0
        }
#pragma unroll
        for (int t = 0; t < BLK_U; ++t)
            if (j + t < nq) split_fma<RL>(C[t], tmain, trem, acc, accr);
    }
#pragma unroll
    for (int gq = 1; gq < NPI; ++gq) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float o = __shfl_down_sync(INMF_FULL_MASK, acc[u], gq * 8);
            acc[u] += (grp == 0) ? o : 0.f;
        }
#pragma unroll
        for (int u = 0; u < RL; ++u) {
            const float o = __shfl_down_sync(INMF_FULL_MASK, accr[u], gq * 8);
            accr[u] += (grp == 0) ? o : 0.f;
        }
    }
}

// H'H of a split tile (columns < 32 in `tmain`, the others in `trem`), same blocking as tile_gram_accumulate.
__device__ __forceinline__ void split_gram_accumulate(const float* __restrict__ tmain, const float* __restrict__ trem,
                                                      int rs, int rows, int k, double* __restrict__ out,
                                                      int mode = INMF_GRAM_ATOMIC) {
    const int nb = (k + 3) >> 2;
    for (int blk = threadIdx.x; blk < nb * nb; blk += blockDim.x) {
        const int bi = blk / nb, bj = blk % nb;
        const float* pa = (4 * bi < 32) ? tmain + 4 * bi : trem + (4 * bi - 32);
        const float* pb = (4 * bj < 32) ? tmain + 4 * bj : trem + (4 * bj - 32);
        const int sa = (4 * bi < 32) ? 32 : rs, sb = (4 * bj < 32) ? 32 : rs;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        for (int r = 0; r < rows; ++r) {
            const float4 va = *reinterpret_cast<const float4*>(pa + (size_t)r * sa);
            const float4 vb = *reinterpret_cast<const float4*>(pb + (size_t)r * sb);
            const float xa[4] = {va.x, va.y, va.z, va.w}, xb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] += xa[a] * xb[b];
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int i = 4 * bi + a, j = 4 * bj + b;
                if (i < k && j < k) {
                    if (mode == INMF_GRAM_PLAIN) out[(size_t)i * k + j] += (double)acc[a][b];
                    else if (mode == INMF_GRAM_MULTICAST) mc_red_add_f64(&out[(size_t)i * k + j], (double)acc[a][b]);
                    else atomicAdd(&out[(size_t)i * k + j], (double)acc[a][b]);
                }
            }
    }
}

template <int RL>
__global__ void __launch_bounds__(1024, 1)
blocked_split_kernel(const int64_t* __restrict__ work, const int64_t* __restrict__ ptr, const int2* __restrict__ pairs,
                     const float* __restrict__ dense, int KP, float* __restrict__ out, int64_t ldo, int k,
                     int max_tile_rows, double* __restrict__ gram_out) {
    constexpr int RS = 8 * RL;  // remainder floats per row in shared memory
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    float* tmain = reinterpret_cast<float*>(smem_raw + 128);
    float* trem = tmain + (size_t)max_tile_rows * 32;
    const int64_t* wd = work + (size_t)blockIdx.x * INMF_WORK_FIELDS;
    const int64_t ptr_base = wd[0], dense_row0 = wd[1];
    const int tile_rows = (int)wd[2];
    const int64_t row_begin = wd[3], row_end = wd[4], out_row_base = wd[5];
    const int gram_seg = (int)wd[6];
    const int R = KP - 32;  // remainder floats per row in global memory (multiple of 4, zero padded beyond k)

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)((size_t)tile_rows * KP * sizeof(float)));
    }
    __syncthreads();
    {
        const float* src = dense + (size_t)dense_row0 * KP;
        for (int r = threadIdx.x; r < tile_rows; r += blockDim.x) {
            bulk_g2s(tmain + (size_t)r * 32, src + (size_t)r * KP, 128u, bar);
            bulk_g2s(trem + (size_t)r * RS, src + (size_t)r * KP + 32, (uint32_t)(R * sizeof(float)), bar);
        }
    }
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, sub = lane & 7;
    const uint32_t ta = smem_u32(tmain) + (uint32_t)(sub * 16);
    const uint32_t tr = smem_u32(trem) + (uint32_t)(sub * RL * 4);
    int64_t r = row_begin + warp;
    int64_t beg = 0, end = 0;
    if (r < row_end) {
        beg = ptr[ptr_base + r];
        end = ptr[ptr_base + r + 1];
    }
    mbar_wait(bar, 0);

    while (r < row_end) {
        const int64_t rn = r + nwarps;
        int64_t nbeg = 0, nend = 0;
        if (rn < row_end) {
            nbeg = ptr[ptr_base + rn];
            nend = ptr[ptr_base + rn + 1];
        }
        if (beg != end) {
            float acc[4], accr[RL];
            warp_row_dot_split<RL>(pairs + beg, (int)(end - beg), grp, ta, tr, acc, accr);
            if (grp == 0) {
                float* o = out + (out_row_base + r) * ldo;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (acc[u] != 0.f) atomicAdd(o + sub * 4 + u, acc[u]);
#pragma unroll
                for (int u = 0; u < RL; ++u)
                    if (32 + sub * RL + u < k && accr[u] != 0.f) atomicAdd(o + 32 + sub * RL + u, accr[u]);
            }
        }
        r = rn;
        beg = nbeg;
        end = nend;
    }
    if (gram_seg >= 0 && gram_out != nullptr)
        split_gram_accumulate(tmain, trem, RS, tile_rows, k, gram_out + (size_t)gram_seg * k * k);
}

// =====================================================================================================
// Wide variant for KP == 32 (fp32): 4 lanes per non-zero, 8 factors per lane (two 128-bit loads), 8 non-zeros per warp
// instruction -- one pair load now feeds 16 FMAs per lane instead of 8, which halves the pair-delivery share of the
// L1 data pipe (0.72 -> 0.36 of ~1.7 wavefronts per non-zero).  To keep every 8-lane phase of a 128-bit load
// conflict free, even lane groups read the first 64 bytes of their row with the first load and the second 64 bytes
// with the second one, odd groups the other way round (a phase = one even + one odd group = banks 0-15 + 16-31).
// Same tile (row-major, 128-byte rows) and same pairs as the plain kernel.
// =====================================================================================================
__device__ __forceinline__ void wide_fma(const int2& p, uint32_t tA, uint32_t tB, float (&a)[4], float (&b)[4]) {
    float m0, m1, m2, m3, n0, n1, n2, n3;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(m0), "=f"(m1), "=f"(m2), "=f"(m3)
                 : "r"(tA + (uint32_t)p.x));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(n0), "=f"(n1), "=f"(n2), "=f"(n3)
                 : "r"(tB + (uint32_t)p.x));
    const float v = __int_as_float(p.y);
    a[0] = fmaf(v, m0, a[0]);
    a[1] = fmaf(v, m1, a[1]);
    a[2] = fmaf(v, m2, a[2]);
    a[3] = fmaf(v, m3, a[3]);
    b[0] = fmaf(v, n0, b[0]);
    b[1] = fmaf(v, n1, b[1]);
    b[2] = fmaf(v, n2, b[2]);
    b[3] = fmaf(v, n3, b[3]);
}

// One warp, one sparse row of m pairs.  lane = (grp = lane / 4, sub = lane % 4).  On return the lanes of group 0 hold
// the row sums: a[] = factors 4*sub .. 4*sub+3, b[] = factors 16 + 4*sub .. 16 + 4*sub + 3.
__device__ __forceinline__ void warp_row_dot_wide(const int2* __restrict__ pl0, int m, int grp, uint32_t tA,
                                                  uint32_t tB, float (&a)[4], float (&b)[4]) {
    typedef PairTraits<float> PT;
    constexpr int NPI = 8;
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = b[u] = 0.f;
    const int2* __restrict__ pl = pl0 + grp;
    const int nfull = m / NPI;
    const int nq = (m + NPI - 1) / NPI;
    int j = 0;
    if (nfull >= BLK_U) {
        int2 A[BLK_U], B[BLK_U];
#pragma unroll
        for (int t = 0; t < BLK_U; ++t) A[t] = PT::ld(pl + t * NPI);
        while (true) {
            const int2* __restrict__ pn = pl + (size_t)(j + BLK_U) * NPI;
            if (j + 2 * BLK_U <= nfull) {
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) B[t] = PT::ld(pn + t * NPI);
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) wide_fma(A[t], tA, tB, a, b);
                j += BLK_U;
            } else {
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) wide_fma(A[t], tA, tB, a, b);
                j += BLK_U;
                break;
            }
            const int2* __restrict__ pm = pl + (size_t)(j + BLK_U) * NPI;
            if (j + 2 * BLK_U <= nfull) {
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) A[t] = PT::ld(pm + t * NPI);
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) wide_fma(B[t], tA, tB, a, b);
                j += BLK_U;
            } else {
#pragma unroll
                for (int t = 0; t < BLK_U; ++t) wide_fma(B[t], tA, tB, a, b);
                j += BLK_U;
                break;
            }
        }
    }
    {
        int2 C[BLK_U];
#pragma unroll
        for (int t = 0; t < BLK_U; ++t) {
            const int q = (j + t) * NPI + grp;
            C[t] = ((j + t) < nq && q < m) ? PT::ld(pl0 + q) : PT::zero();
        }
#pragma unroll
        for (int t = 0; t < BLK_U; ++t)
            if (j + t < nq) wide_fma(C[t], tA, tB, a, b);
    }
    // odd groups hold the halves the other way round: fold them onto their even neighbours with a swap
    {
        float oa[4], ob[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            oa[u] = __shfl_down_sync(INMF_FULL_MASK, a[u], 4);
            ob[u] = __shfl_down_sync(INMF_FULL_MASK, b[u], 4);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a[u] += ob[u];
            b[u] += oa[u];
        }
    }
    // even groups 4, 6 onto 0, 2; then 2 onto 0 (only the lanes of group 0 are read afterwards)
#pragma unroll
    for (int d = 16; d >= 8; d >>= 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a[u] += __shfl_down_sync(INMF_FULL_MASK, a[u], d);
            b[u] += __shfl_down_sync(INMF_FULL_MASK, b[u], d);
        }
    }
}

__global__ void __launch_bounds__(1024, 1)
blocked_wide_kernel(const int64_t* __restrict__ work, const int64_t* __restrict__ ptr, const int2* __restrict__ pairs,
                    const float* __restrict__ dense, float* __restrict__ out, int64_t ldo, int k,
                    double* __restrict__ gram_out) {
    constexpr int KP = 32;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    float* tile = reinterpret_cast<float*>(smem_raw + 128);
    const int64_t* wd = work + (size_t)blockIdx.x * INMF_WORK_FIELDS;
    const int64_t ptr_base = wd[0], dense_row0 = wd[1];
    const int tile_rows = (int)wd[2];
    const int64_t row_begin = wd[3], row_end = wd[4], out_row_base = wd[5];
    const int gram_seg = (int)wd[6];

    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t total = (size_t)tile_rows * KP * sizeof(float);
        mbar_expect_tx(bar, (uint32_t)total);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(dense + (size_t)dense_row0 * KP);
        unsigned char* dst = reinterpret_cast<unsigned char*>(tile);
        for (size_t off = 0; off < total; off += 32768) {
            const size_t n = (total - off < 32768) ? (total - off) : 32768;
            bulk_g2s(dst + off, src + off, (uint32_t)n, bar);
        }
    }
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 2, sub = lane & 3, par = grp & 1;
    const uint32_t tA = smem_u32(tile) + (uint32_t)((sub + 4 * par) * 16);
    const uint32_t tB = smem_u32(tile) + (uint32_t)((sub + 4 * (1 - par)) * 16);
    int64_t r = row_begin + warp;
    int64_t beg = 0, end = 0;
    if (r < row_end) {
        beg = ptr[ptr_base + r];
        end = ptr[ptr_base + r + 1];
    }
    mbar_wait(bar, 0);

    while (r < row_end) {
        const int64_t rn = r + nwarps;
        int64_t nbeg = 0, nend = 0;
        if (rn < row_end) {
            nbeg = ptr[ptr_base + rn];
            nend = ptr[ptr_base + rn + 1];
        }
        if (beg != end) {
            float a[4], b[4];
            warp_row_dot_wide(pairs + beg, (int)(end - beg), grp, tA, tB, a, b);
            if (grp == 0) {  // group 0 is even: a = first half (factors 4*sub ..), b = second half (16 + 4*sub ..)
                float* o = out + (out_row_base + r) * ldo + sub * 4;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (sub * 4 + u < k && a[u] != 0.f) atomicAdd(o + u, a[u]);
                    if (16 + sub * 4 + u < k && b[u] != 0.f) atomicAdd(o + 16 + u, b[u]);
                }
            }
        }
        r = rn;
        beg = nbeg;
        end = nend;
    }
    if (gram_seg >= 0 && gram_out != nullptr)
        tile_gram_accumulate<float>(tile, KP, tile_rows, k, gram_out + (size_t)gram_seg * k * k, 1.0);
}
