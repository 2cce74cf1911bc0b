// Sliced (SELL-8) form of the two X-streaming passes for fp32 (k <= 64): the default of an ALS iteration.
//
//   pass 1:  R[cell,:]  += X[cell, gene block] * Mt[gene block, :]      "row" = (cell, gene block) segment
//   pass 2:  St[gene,:] += X[cell block, gene]' * H[cell block, :]      "row" = (gene, cell block) segment
//
// Same tiles and the same work split as blocked.cuh (one persistent CTA per SM holds one dense tile in shared
// memory, brought in by TMA bulk copies), but the sparse rows of a CTA are sorted by length and packed in slices of
// 8 rows stored step-interleaved:
//
//     slice = [step pair jj = 0 .. L2)[row i = 0 .. 8) { off(2jj), val(2jj), off(2jj+1), val(2jj+1) }   (int4)
//
// A warp owns a slice, a group of 4 lanes owns ONE ROW of it and each lane 8 factors (+ its share of the factors
// beyond 32), so that
//   * one coalesced 128-bit load per lane (8 distinct 16-byte pieces = one 128-byte line per warp instruction)
//     delivers 16 non-zeros: 1/16 L1 wavefront per non-zero for the pair stream instead of ~1/4;
//   * there is no cross-lane fold at the end of a row (the old kernels spend 24 shuffles per ~120 non-zeros on it),
//     no per-row pointer chasing and no ragged row tails: rows of a slice have (almost) the same length;
//   * the gathered operand is read with two 128-bit loads per lane whose 8-lane phases are conflict free (even
//     groups read the first 64 bytes of their tile row first, odd groups the second 64 bytes).
// What remains is the operand traffic itself: 4 * KP bytes of shared memory per non-zero.
//
// Pairs: {int32 off, float val}, off = (tile row) * 128 = byte offset of the row in the [rows x 32] main part of the
// tile; factors beyond 32 live in a [rows x REM] remainder part (REM = 8, 16, 32 floats) addressed by a shift of it.
// Padding pairs are {0, 0.0f}.  Slices are handed out dynamically, longest first.
#pragma once
#include "blocked.cuh"

#define SELL_WORK_FIELDS 8
// work[w] = { slice_begin, slice_end, dense_row0, tile_rows, out_row_base, gram_seg (-1: none), part, unused }
// Deterministic mode (part_stride > 0): `out` is an array of partial buffers part_stride elements apart; work item w
// STORES its rows into partial `part` (rows of one partial are written by exactly one work item, each by one lane
// group, so no atomics and no dependence on the order in which CTAs run); the caller then adds the partials in a fixed
// order (inmf_reduce_parts_*).  gram_out likewise (gram_part_stride doubles apart).
#define SELL_U 2  // 128-bit pair loads per pipeline stage

__device__ __forceinline__ int4 sell_ld(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void lds128(uint32_t addr, float (&m)[4]) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(m[0]), "=f"(m[1]), "=f"(m[2]), "=f"(m[3]) : "r"(addr));
}
__device__ __forceinline__ void lds64(uint32_t addr, float (&m)[2]) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(m[0]), "=f"(m[1]) : "r"(addr));
}
__device__ __forceinline__ void red_add(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add4(float* p, const float (&v)[4]) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
                 : "memory");
}

// Per-lane accumulators of one row: a/b = the two 4-factor chunks of the main part, r = remainder share.
template <int REM>
struct SellAcc {
    static constexpr int NR = REM == 0 ? 1 : REM / 4;  // (REM == 32: two chunks of 4)
    float a[4], b[4], r[NR];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = b[u] = 0.f;
#pragma unroll
        for (int u = 0; u < NR; ++u) r[u] = 0.f;
    }
};

template <int REM>
__device__ __forceinline__ void sell_fma(int off, int valbits, uint32_t tA, uint32_t tB, uint32_t tR, uint32_t tR2,
                                         SellAcc<REM>& c) {
    float m[4], n[4];
    lds128(tA + (uint32_t)off, m);
    lds128(tB + (uint32_t)off, n);
    const float v = __int_as_float(valbits);
#pragma unroll
    for (int u = 0; u < 4; ++u) c.a[u] = fmaf(v, m[u], c.a[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) c.b[u] = fmaf(v, n[u], c.b[u]);
    if constexpr (REM == 8) {
        float q[2];
        lds64(tR + ((uint32_t)off >> 2), q);
        c.r[0] = fmaf(v, q[0], c.r[0]);
        c.r[1] = fmaf(v, q[1], c.r[1]);
    } else if constexpr (REM == 16) {
        float q[4];
        lds128(tR + ((uint32_t)off >> 1), q);
#pragma unroll
        for (int u = 0; u < 4; ++u) c.r[u] = fmaf(v, q[u], c.r[u]);
    } else if constexpr (REM == 32) {
        float q[4], s[4];
        lds128(tR + (uint32_t)off, q);
        lds128(tR2 + (uint32_t)off, s);
#pragma unroll
        for (int u = 0; u < 4; ++u) c.r[u] = fmaf(v, q[u], c.r[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) c.r[4 + u] = fmaf(v, s[u], c.r[4 + u]);
    }
}

template <int REM>
__device__ __forceinline__ void sell_consume(const int4& q, uint32_t tA, uint32_t tB, uint32_t tR, uint32_t tR2,
                                             SellAcc<REM>& c) {
    sell_fma<REM>(q.x, q.y, tA, tB, tR, tR2, c);
    sell_fma<REM>(q.z, q.w, tA, tB, tR, tR2, c);
}

__device__ __forceinline__ void red_add2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// The same adds on a multicast address (every bound GPU's copy receives them: pass 2 fused with its all-reduce).
__device__ __forceinline__ void mc_red_add(float* p, float v) {
    asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void mc_red_add2(float* p, float a, float b) {
    asm volatile("multimem.red.relaxed.sys.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mc_red_add4(float* p, const float (&v)[4]) {
    asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]),
                 "f"(v[2]), "f"(v[3])
                 : "memory");
}

// ---- pass 2 fused with its all-reduce (MC kernels) --------------------------------------------------------------
// Partial sums are accumulated in the rank's own `out` exactly as in the single-GPU pass (L2 atomics).  The work items
// that add into one (dataset, gene range) share an arrival counter; the CTA that arrives LAST knows the range is
// final on this rank and adds it -- contiguous, 16 bytes per lane -- into the multicast copy, i.e. into every rank's
// sum buffer through the NVLink switch, while other CTAs are still computing.  (Adding every partial sum through the
// switch instead was measured at ~37 GB/s per GPU, 7x slower than the pass itself: the switch wants few large adds.)
struct SellMcArgs {
    float* out_mc;         // multicast address of the sum buffer, same layout as `out` (ldo == k)
    double* gram_mc;       // multicast address of the k x k sums, same layout as gram_out
    int* counters;         // one per (dataset, gene range), zero before the first launch (the last arriver resets it)
    const int* expected;   // work items per counter
};

__device__ __forceinline__ void sell_mc_push(const int64_t* __restrict__ wd, const float* __restrict__ out, int64_t ldo,
                                             int k, const double* __restrict__ gram_out,
                                             const SellMcArgs* __restrict__ mc, int* flag_smem) {
    __threadfence();  // this thread's adds are performed before the arrival is counted
    __syncthreads();
    if (threadIdx.x == 0) {
        const int id = (int)wd[6];
        const int old = atomicAdd(&mc->counters[id], 1);
        const int last = old == mc->expected[id] - 1;
        if (last) mc->counters[id] = 0;
        *flag_smem = last;
    }
    __syncthreads();
    if (!*flag_smem) return;
    __threadfence();
    const int64_t lo = wd[7] >> 32, hi = wd[7] & 0xffffffffll;
    const size_t beg = (size_t)(wd[4] + lo) * ldo, n = (size_t)(hi - lo) * ldo;
    const float* src = out + beg;
    float* dst = mc->out_mc + beg;
    size_t head = ((16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15) >> 2;
    if (head > n) head = n;
    for (size_t i = threadIdx.x; i < head; i += blockDim.x) {
        const float v = __ldcg(src + i);
        if (v != 0.f) mc_red_add(dst + i, v);
    }
    const size_t n4 = (n - head) >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src + head);
    float* d4 = dst + head;
    for (size_t i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 t = __ldcg(s4 + i);
        const float v[4] = {t.x, t.y, t.z, t.w};
        if (t.x != 0.f || t.y != 0.f || t.z != 0.f || t.w != 0.f) mc_red_add4(d4 + 4 * i, v);
    }
    for (size_t i = head + 4 * n4 + threadIdx.x; i < n; i += blockDim.x) {
        const float v = __ldcg(src + i);
        if (v != 0.f) mc_red_add(dst + i, v);
    }
    if (wd[5] >= 0 && gram_out != nullptr) {  // (the items that add a tile Gram are exactly those of the first range)
        const size_t gb = (size_t)wd[5] * k * k;
        for (int i = threadIdx.x; i < k * k; i += blockDim.x) {
            const double v = __ldcg(gram_out + gb + i);
            if (v != 0.0) mc_red_add_f64(mc->gram_mc + gb + i, v);
        }
    }
}

// out[col0 .. col0+3] += v[0..3] for the columns < k.  vec = 4: rows are 16-byte aligned, 2: 8-byte aligned, else 1;
// vec = 0: plain stores (deterministic mode, zeros included).
__device__ __forceinline__ void sell_out4(float* __restrict__ o, int col0, int k, const float (&v)[4], int vec) {
    if (vec == 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (col0 + u < k) o[col0 + u] = v[u];
        return;
    }
    if (vec == 4 && col0 + 4 <= k) {
        if (v[0] != 0.f || v[1] != 0.f || v[2] != 0.f || v[3] != 0.f) red_add4(o + col0, v);
    } else if (vec >= 2 && col0 + 4 <= k) {
        if (v[0] != 0.f || v[1] != 0.f) red_add2(o + col0, v[0], v[1]);
        if (v[2] != 0.f || v[3] != 0.f) red_add2(o + col0 + 2, v[2], v[3]);
    } else {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (col0 + u < k && v[u] != 0.f) red_add(o + col0 + u, v[u]);
    }
}

template <int REM, bool MC>
__global__ void __launch_bounds__(1024, 1)
sell_pass_kernel(const int64_t* __restrict__ work, const int64_t* __restrict__ slice_ptr,
                 const int32_t* __restrict__ rowid, const int4* __restrict__ pairs, const float* __restrict__ dense,
                 int KP, float* __restrict__ out, int64_t ldo, int k, int max_tile_rows,
                 double* __restrict__ gram_out, int64_t part_stride, int64_t gram_part_stride,
                 const SellMcArgs* __restrict__ mc) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    int* next_slice = reinterpret_cast<int*>(smem_raw + 16);
    float* tmain = reinterpret_cast<float*>(smem_raw + 128);
    float* trem = tmain + (size_t)max_tile_rows * 32;
    const int64_t* wd = work + (size_t)blockIdx.x * SELL_WORK_FIELDS;
    const int slice_begin = (int)wd[0], slice_end = (int)wd[1];
    const int64_t dense_row0 = wd[2];
    const int tile_rows = (int)wd[3];
    const int64_t out_row_base = wd[4];
    const int gram_seg = (int)wd[5];
    if (part_stride > 0) {
        out += (size_t)wd[6] * part_stride;
        if (gram_out != nullptr) gram_out += (size_t)wd[6] * gram_part_stride;
    }
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    (void)nwarps;
    (void)warp;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)((size_t)tile_rows * KP * sizeof(float)));
        *next_slice = slice_begin;
    }
    __syncthreads();
    {
        const float* src = dense + (size_t)dense_row0 * KP;
        if constexpr (REM == 0) {  // KP == 32: the tile is one contiguous block
            if (threadIdx.x == 0) {
                const size_t total = (size_t)tile_rows * 128;
                const unsigned char* s = reinterpret_cast<const unsigned char*>(src);
                unsigned char* d = reinterpret_cast<unsigned char*>(tmain);
                for (size_t off = 0; off < total; off += 32768) {
                    const size_t n = (total - off < 32768) ? (total - off) : 32768;
                    bulk_g2s(d + off, s + off, (uint32_t)n, bar);
                }
            }
        } else {
            const int R = KP - 32;  // remainder floats per row in global memory (multiple of 4)
            for (int r = threadIdx.x; r < tile_rows; r += blockDim.x) {
                bulk_g2s(tmain + (size_t)r * 32, src + (size_t)r * KP, 128u, bar);
                bulk_g2s(trem + (size_t)r * REM, src + (size_t)r * KP + 32, (uint32_t)(R * sizeof(float)), bar);
            }
        }
    }
    const int grp = lane >> 2, sub = lane & 3, par = grp & 1;
    const uint32_t tA = smem_u32(tmain) + (uint32_t)((sub + 4 * par) * 16);
    const uint32_t tB = smem_u32(tmain) + (uint32_t)((sub + 4 * (1 - par)) * 16);
    uint32_t tR = 0, tR2 = 0;
    if constexpr (REM == 8) tR = smem_u32(trem) + (uint32_t)(sub * 8);
    if constexpr (REM == 16) tR = smem_u32(trem) + (uint32_t)(sub * 16);
    if constexpr (REM == 32) {
        tR = smem_u32(trem) + (uint32_t)((sub + 4 * par) * 16);
        tR2 = smem_u32(trem) + (uint32_t)((sub + 4 * (1 - par)) * 16);
    }
    // part_stride > 0: partial buffers (plain stores)
    const int vec_ok = part_stride > 0 ? 0
                       : (((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) ? 4
                       : (((ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0)) ? 2 : 1;
    const int gram_mode = part_stride > 0 ? INMF_GRAM_PLAIN : INMF_GRAM_ATOMIC;

    mbar_wait(bar, 0);
    // H'H of the resident tile (pass 2): done FIRST by the few warps whose threads own an output block, while all
    // other warps already stream slices -- the slices are handed out dynamically, so those warps simply join later
    // and the latency-bound Gram loop costs no tail.
    if (gram_seg >= 0 && gram_out != nullptr) {
        if constexpr (REM == 0)
            tile_gram_accumulate<float>(tmain, 32, tile_rows, k, gram_out + (size_t)gram_seg * k * k, 1.0, gram_mode);
        else
            split_gram_accumulate(tmain, trem, REM, tile_rows, k, gram_out + (size_t)gram_seg * k * k, gram_mode);
    }
    int s = 0;
    if (lane == 0) s = atomicAdd(next_slice, 1);
    s = __shfl_sync(INMF_FULL_MASK, s, 0);
    int64_t base = 0, lim = 0;
    if (s < slice_end) {
        base = slice_ptr[s];
        lim = slice_ptr[s + 1];
    }

    while (s < slice_end) {
        const int L2 = (int)((lim - base) >> 3);
        const int4* __restrict__ p = pairs + base + grp;
        const int rid = rowid[(size_t)s * 8 + grp];
        SellAcc<REM> c;
        c.clear();
        const int4 z4 = make_int4(0, 0, 0, 0);
        int4 A[SELL_U], B[SELL_U];
#pragma unroll
        for (int t = 0; t < SELL_U; ++t) A[t] = (t < L2) ? sell_ld(p + t * 8) : z4;
        int j = 0;
        while (j < L2) {
#pragma unroll
            for (int t = 0; t < SELL_U; ++t) B[t] = (j + SELL_U + t < L2) ? sell_ld(p + (size_t)(j + SELL_U + t) * 8) : z4;
#pragma unroll
            for (int t = 0; t < SELL_U; ++t)
                if (j + t < L2) {
                    INMF_CHECK((uint32_t)A[t].x < (uint32_t)tile_rows * 128u && (uint32_t)A[t].z < (uint32_t)tile_rows * 128u);
                    sell_consume<REM>(A[t], tA, tB, tR, tR2, c);
                }
            j += SELL_U;
            if (j >= L2) break;
#pragma unroll
            for (int t = 0; t < SELL_U; ++t) A[t] = (j + SELL_U + t < L2) ? sell_ld(p + (size_t)(j + SELL_U + t) * 8) : z4;
#pragma unroll
            for (int t = 0; t < SELL_U; ++t)
                if (j + t < L2) {
                    INMF_CHECK((uint32_t)B[t].x < (uint32_t)tile_rows * 128u && (uint32_t)B[t].z < (uint32_t)tile_rows * 128u);
                    sell_consume<REM>(B[t], tA, tB, tR, tR2, c);
                }
            j += SELL_U;
        }
        // next slice (dynamic, longest first): fetch before the output so the latencies overlap
        int sn = 0;
        if (lane == 0) sn = atomicAdd(next_slice, 1);
        sn = __shfl_sync(INMF_FULL_MASK, sn, 0);
        int64_t nbase = 0, nlim = 0;
        if (sn < slice_end) {
            nbase = slice_ptr[sn];
            nlim = slice_ptr[sn + 1];
        }
        if (rid >= 0) {
            float* o = out + (out_row_base + rid) * ldo;
            // even groups: a = factors 4*sub.., b = 16 + 4*sub..; odd groups the other way round
            sell_out4(o, 4 * sub + 16 * par, k, c.a, vec_ok);
            sell_out4(o, 4 * sub + 16 * (1 - par), k, c.b, vec_ok);
            if constexpr (REM == 8) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (32 + 2 * sub + u < k) {
                        if (vec_ok == 0) o[32 + 2 * sub + u] = c.r[u];
                        else if (c.r[u] != 0.f) {
                            red_add(o + 32 + 2 * sub + u, c.r[u]);
                        }
                    }
                }
            } else if constexpr (REM == 16) {
                float q[4] = {c.r[0], c.r[1], c.r[2], c.r[3]};
                sell_out4(o, 32 + 4 * sub, k, q, vec_ok);
            } else if constexpr (REM == 32) {
                float q[4] = {c.r[0], c.r[1], c.r[2], c.r[3]};
                float t4[4] = {c.r[4], c.r[5], c.r[6], c.r[7]};
                sell_out4(o, 32 + 4 * sub + 16 * par, k, q, vec_ok);
                sell_out4(o, 32 + 4 * sub + 16 * (1 - par), k, t4, vec_ok);
            }
        }
        s = sn;
        base = nbase;
        lim = nlim;
    }
    if constexpr (MC) sell_mc_push(wd, out, ldo, k, gram_out, mc, next_slice + 1);
}

// =====================================================================================================
// One lane per row, 32 rows per slice, KP == 32 (k <= 32): the pair stream is not replicated across lanes at all.
//
// ncu on the 4-lanes-per-row kernel above (C2, k = 30): 1.03 shared-memory wavefronts per non-zero for the operand --
// the floor -- but another 0.30 for the pairs, because a 128-bit load returns 512 bytes to the register file even when
// groups of 4 lanes ask for the same 16 bytes; the L1 data pipe is 91-98 % busy, so that is 23 % of the kernel.
// Here a lane owns a whole row: one 128-bit load per lane brings two non-zeros of ITS row (512 distinct bytes per
// warp instruction, 1/16 wavefront per non-zero) and the lane reads all eight 16-byte chunks of the gathered tile
// row itself.  Lanes of a phase (8 lanes) would all hit the same 4 banks, so lane l walks the chunks in the rotated
// order (l + j) mod 8: its j-th load always goes to accumulator chunk j (static register index), and the 8 lanes of a
// phase always touch 8 different bank groups, whatever rows they gather.  32 accumulators per lane: 512 threads per
// CTA with up to 128 registers.
// Slices: [step pair jj][row i = 0 .. 32) int4; rowid int32[nslices * 32].
// =====================================================================================================
#define SELL32_THREADS 512

template <bool MC>
__global__ void __launch_bounds__(SELL32_THREADS, 1)
sell32_pass_kernel(const int64_t* __restrict__ work, const int64_t* __restrict__ slice_ptr,
                   const int32_t* __restrict__ rowid, const int4* __restrict__ pairs, const float* __restrict__ dense,
                   float* __restrict__ out, int64_t ldo, int k, double* __restrict__ gram_out, int64_t part_stride,
                   int64_t gram_part_stride, const SellMcArgs* __restrict__ mc) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    int* next_slice = reinterpret_cast<int*>(smem_raw + 16);
    float* tile = reinterpret_cast<float*>(smem_raw + 128);
    const int64_t* wd = work + (size_t)blockIdx.x * SELL_WORK_FIELDS;
    const int slice_begin = (int)wd[0], slice_end = (int)wd[1];
    const int64_t dense_row0 = wd[2];
    const int tile_rows = (int)wd[3];
    const int64_t out_row_base = wd[4];
    const int gram_seg = (int)wd[5];
    if (part_stride > 0) {
        out += (size_t)wd[6] * part_stride;
        if (gram_out != nullptr) gram_out += (size_t)wd[6] * gram_part_stride;
    }
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        *next_slice = slice_begin;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t total = (size_t)tile_rows * 128;
        mbar_expect_tx(bar, (uint32_t)total);
        const unsigned char* sp = reinterpret_cast<const unsigned char*>(dense + (size_t)dense_row0 * 32);
        unsigned char* d = reinterpret_cast<unsigned char*>(tile);
        for (size_t off = 0; off < total; off += 32768) {
            const size_t n = (total - off < 32768) ? (total - off) : 32768;
            bulk_g2s(d + off, sp + off, (uint32_t)n, bar);
        }
    }
    uint32_t tb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) tb[j] = smem_u32(tile) + (uint32_t)(((lane + j) & 7) * 16);
    // part_stride > 0: partial buffers (plain stores)
    const int vec_ok = part_stride > 0 ? 0
                       : (((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) ? 4
                       : (((ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0)) ? 2 : 1;
    const int gram_mode = part_stride > 0 ? INMF_GRAM_PLAIN : INMF_GRAM_ATOMIC;

    mbar_wait(bar, 0);
    if (gram_seg >= 0 && gram_out != nullptr)  // first, by the warps that own output blocks (see sell_pass_kernel)
        tile_gram_accumulate<float>(tile, 32, tile_rows, k, gram_out + (size_t)gram_seg * k * k, 1.0, gram_mode);
    int s = 0;
    if (lane == 0) s = atomicAdd(next_slice, 1);
    s = __shfl_sync(INMF_FULL_MASK, s, 0);
    int64_t base = 0, lim = 0;
    if (s < slice_end) {
        base = slice_ptr[s];
        lim = slice_ptr[s + 1];
    }

    while (s < slice_end) {
        const int L2 = (int)((lim - base) >> 5);
        const int4* __restrict__ p = pairs + base + lane;
        const int rid = rowid[(size_t)s * 32 + lane];
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[j][u] = 0.f;
        const int4 z4 = make_int4(0, 0, 0, 0);
        int4 A = (0 < L2) ? sell_ld(p) : z4;
        int4 B = (1 < L2) ? sell_ld(p + 32) : z4;
        for (int j = 0; j < L2; j += 2) {
            const int4 A2 = (j + 2 < L2) ? sell_ld(p + (size_t)(j + 2) * 32) : z4;
            const int4 B2 = (j + 3 < L2) ? sell_ld(p + (size_t)(j + 3) * 32) : z4;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int off = h == 0 ? A.x : h == 1 ? A.z : h == 2 ? B.x : B.z;
                const float v = __int_as_float(h == 0 ? A.y : h == 1 ? A.w : h == 2 ? B.y : B.w);
                INMF_CHECK((uint32_t)off < (uint32_t)tile_rows * 128u && (off & 127) == 0);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float m[4];
                    lds128(tb[c] + (uint32_t)off, m);
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc[c][u] = fmaf(v, m[u], acc[c][u]);
                }
            }
            A = A2;
            B = B2;
        }
        int sn = 0;
        if (lane == 0) sn = atomicAdd(next_slice, 1);
        sn = __shfl_sync(INMF_FULL_MASK, sn, 0);
        int64_t nbase = 0, nlim = 0;
        if (sn < slice_end) {
            nbase = slice_ptr[sn];
            nlim = slice_ptr[sn + 1];
        }
        if (rid >= 0) {
            float* o = out + (out_row_base + rid) * ldo;
#pragma unroll
            for (int c = 0; c < 8; ++c) sell_out4(o, ((lane + c) & 7) * 4, k, acc[c], vec_ok);
        }
        s = sn;
        base = nbase;
        lim = nlim;
    }
    if constexpr (MC) sell_mc_push(wd, out, ldo, k, gram_out, mc, next_slice + 1);
}

// =====================================================================================================
// One lane per row for 32 < k <= 40 (the k = 40 target shape): [rows x 32] main tile + [rows x 8] remainder.
//
// The main part is read as in sell32_pass_kernel (rotated chunk order, conflict free).  A remainder row is 32 bytes =
// two 16-byte chunks, four rows per 128-byte bank line: its bank position is (2 * (row mod 4) + chunk), i.e. data
// dependent, and the four-lanes-per-row kernel pays ~2-way conflicts for it (plus 0.25 wavefront per non-zero for the
// replicated pair stream).  Here even lanes read chunk 0 first and chunk 1 second, odd lanes the other way round, so
// in each of the two loads the even lanes hit even positions and the odd lanes odd ones; and the builder
// (sell_fill_residue_kernel) orders every row's non-zeros so that at step j lane l gathers a tile row with
// (row mod 4) == (l / 2 + j) mod 4 while its supply of that residue class lasts -- the four even (odd) lanes of a
// phase then hit four different positions.  Conflicts remain only where a row runs out of a class early.
// =====================================================================================================
template <bool MC>
__global__ void __launch_bounds__(SELL32_THREADS, 1)
sell32r_pass_kernel(const int64_t* __restrict__ work, const int64_t* __restrict__ slice_ptr,
                    const int32_t* __restrict__ rowid, const int4* __restrict__ pairs, const float* __restrict__ dense,
                    int KP, float* __restrict__ out, int64_t ldo, int k, int max_tile_rows,
                    double* __restrict__ gram_out, int64_t part_stride, int64_t gram_part_stride,
                    const SellMcArgs* __restrict__ mc) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    int* next_slice = reinterpret_cast<int*>(smem_raw + 16);
    float* tmain = reinterpret_cast<float*>(smem_raw + 128);
    float* trem = tmain + (size_t)max_tile_rows * 32;
    const int64_t* wd = work + (size_t)blockIdx.x * SELL_WORK_FIELDS;
    const int slice_begin = (int)wd[0], slice_end = (int)wd[1];
    const int64_t dense_row0 = wd[2];
    const int tile_rows = (int)wd[3];
    const int64_t out_row_base = wd[4];
    const int gram_seg = (int)wd[5];
    if (part_stride > 0) {
        out += (size_t)wd[6] * part_stride;
        if (gram_out != nullptr) gram_out += (size_t)wd[6] * gram_part_stride;
    }
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)((size_t)tile_rows * KP * sizeof(float)));
        *next_slice = slice_begin;
    }
    __syncthreads();
    {
        const float* src = dense + (size_t)dense_row0 * KP;
        const int R = KP - 32;  // remainder floats per row in global memory (4 or 8)
        for (int r = threadIdx.x; r < tile_rows; r += blockDim.x) {
            bulk_g2s(tmain + (size_t)r * 32, src + (size_t)r * KP, 128u, bar);
            bulk_g2s(trem + (size_t)r * 8, src + (size_t)r * KP + 32, (uint32_t)(R * sizeof(float)), bar);
        }
    }
    uint32_t tb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) tb[j] = smem_u32(tmain) + (uint32_t)(((lane + j) & 7) * 16);
    const uint32_t rA = smem_u32(trem) + (uint32_t)((lane & 1) * 16);
    const uint32_t rB = smem_u32(trem) + (uint32_t)((1 - (lane & 1)) * 16);
    // part_stride > 0: partial buffers (plain stores)
    const int vec_ok = part_stride > 0 ? 0
                       : (((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) ? 4
                       : (((ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0)) ? 2 : 1;
    const int gram_mode = part_stride > 0 ? INMF_GRAM_PLAIN : INMF_GRAM_ATOMIC;
    mbar_wait(bar, 0);
    if (gram_seg >= 0 && gram_out != nullptr)
        split_gram_accumulate(tmain, trem, 8, tile_rows, k, gram_out + (size_t)gram_seg * k * k, gram_mode);
    int s = 0;
    if (lane == 0) s = atomicAdd(next_slice, 1);
    s = __shfl_sync(INMF_FULL_MASK, s, 0);
    int64_t base = 0, lim = 0;
    if (s < slice_end) {
        base = slice_ptr[s];
        lim = slice_ptr[s + 1];
    }
    while (s < slice_end) {
        const int L2 = (int)((lim - base) >> 5);
        const int4* __restrict__ p = pairs + base + lane;
        const int rid = rowid[(size_t)s * 32 + lane];
        float acc[8][4], ra[4], rb[4];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[j][u] = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) ra[u] = rb[u] = 0.f;
        const int4 z4 = make_int4(0, 0, 0, 0);
        int4 A = (0 < L2) ? sell_ld(p) : z4;
        for (int j = 0; j < L2; ++j) {
            const int4 A2 = (j + 1 < L2) ? sell_ld(p + (size_t)(j + 1) * 32) : z4;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int off = h == 0 ? A.x : A.z;
                const float v = __int_as_float(h == 0 ? A.y : A.w);
                INMF_CHECK((uint32_t)off < (uint32_t)tile_rows * 128u && (off & 127) == 0);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float m[4];
                    lds128(tb[c] + (uint32_t)off, m);
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc[c][u] = fmaf(v, m[u], acc[c][u]);
                }
                float m[4], n[4];
                lds128(rA + ((uint32_t)off >> 2), m);
                lds128(rB + ((uint32_t)off >> 2), n);
#pragma unroll
                for (int u = 0; u < 4; ++u) ra[u] = fmaf(v, m[u], ra[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) rb[u] = fmaf(v, n[u], rb[u]);
            }
            A = A2;
        }
        int sn = 0;
        if (lane == 0) sn = atomicAdd(next_slice, 1);
        sn = __shfl_sync(INMF_FULL_MASK, sn, 0);
        int64_t nbase = 0, nlim = 0;
        if (sn < slice_end) {
            nbase = slice_ptr[sn];
            nlim = slice_ptr[sn + 1];
        }
        if (rid >= 0) {
            float* o = out + (out_row_base + rid) * ldo;
#pragma unroll
            for (int c = 0; c < 8; ++c) sell_out4(o, ((lane + c) & 7) * 4, k, acc[c], vec_ok);
            sell_out4(o, 32 + 4 * (lane & 1), k, ra, vec_ok);        // even lanes: ra = factors 32..35, rb = 36..39
            sell_out4(o, 32 + 4 * (1 - (lane & 1)), k, rb, vec_ok);
        }
        s = sn;
        base = nbase;
        lim = nlim;
    }
    if constexpr (MC) sell_mc_push(wd, out, ldo, k, gram_out, mc, next_slice + 1);
}

// =====================================================================================================
// Mixed precision: the gathered tile in bf16 (64-byte rows), fp32 pairs and fp32 accumulation.
//
// The fp32 kernels above sit on the shared-memory operand bandwidth (128 B per non-zero at 128 B/clk/SM: ncu 1.03
// wavefronts per non-zero, data pipe 91 % busy).  Halving the operand halves that bound; X itself (the values of the
// pairs) and the accumulators stay fp32, so only the dense factor (W + V_s, or H) is rounded, to 8 significant bits.
// One lane per row as in sell32_pass_kernel: a tile row is four 16-byte chunks, lane l reads them in the order
// (l + j) mod 4 and converts each bf16 pair with one shift / one mask.  Two rows share a 128-byte bank line, so lanes
// l and l + 4 of a phase collide when their rows have the same parity: the builder orders every row's non-zeros
// "matching parity first" for lanes with bit 2 clear and "other parity first" for lanes with bit 2 set, which leaves
// a conflict only where the two halves of the rows differ in length (~5 % of the steps).
// Pairs: {int32 off = tile row * 64, float val}.  This is the dtype="mixed" mode of the engine, never the fp32 path.
// =====================================================================================================
__device__ __forceinline__ void lds128u(uint32_t addr, uint32_t (&m)[4]) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(m[0]), "=r"(m[1]), "=r"(m[2]), "=r"(m[3]) : "r"(addr));
}

__global__ void __launch_bounds__(SELL32_THREADS, 1)
sell32h_pass_kernel(const int64_t* __restrict__ work, const int64_t* __restrict__ slice_ptr,
                    const int32_t* __restrict__ rowid, const int4* __restrict__ pairs,
                    const uint16_t* __restrict__ dense, float* __restrict__ out, int64_t ldo, int k) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    int* next_slice = reinterpret_cast<int*>(smem_raw + 16);
    unsigned char* tile = smem_raw + 128;
    const int64_t* wd = work + (size_t)blockIdx.x * SELL_WORK_FIELDS;
    const int slice_begin = (int)wd[0], slice_end = (int)wd[1];
    const int64_t dense_row0 = wd[2];
    const int tile_rows = (int)wd[3];
    const int64_t out_row_base = wd[4];
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        *next_slice = slice_begin;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t total = (size_t)tile_rows * 64;
        mbar_expect_tx(bar, (uint32_t)total);
        const unsigned char* sp = reinterpret_cast<const unsigned char*>(dense + (size_t)dense_row0 * 32);
        for (size_t off = 0; off < total; off += 32768) {
            const size_t n = (total - off < 32768) ? (total - off) : 32768;
            bulk_g2s(tile + off, sp + off, (uint32_t)n, bar);
        }
    }
    uint32_t tb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tb[j] = smem_u32(tile) + (uint32_t)(((lane + j) & 3) * 16);
    const int vec_ok = (((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) ? 4
                       : (((ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0)) ? 2 : 1;
    mbar_wait(bar, 0);
    int s = 0;
    if (lane == 0) s = atomicAdd(next_slice, 1);
    s = __shfl_sync(INMF_FULL_MASK, s, 0);
    int64_t base = 0, lim = 0;
    if (s < slice_end) {
        base = slice_ptr[s];
        lim = slice_ptr[s + 1];
    }
    while (s < slice_end) {
        const int L2 = (int)((lim - base) >> 5);
        const int4* __restrict__ p = pairs + base + lane;
        const int rid = rowid[(size_t)s * 32 + lane];
        float acc[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[j][u] = 0.f;
        const int4 z4 = make_int4(0, 0, 0, 0);
        int4 A = (0 < L2) ? sell_ld(p) : z4;
        int4 B = (1 < L2) ? sell_ld(p + 32) : z4;
        for (int j = 0; j < L2; j += 2) {
            const int4 A2 = (j + 2 < L2) ? sell_ld(p + (size_t)(j + 2) * 32) : z4;
            const int4 B2 = (j + 3 < L2) ? sell_ld(p + (size_t)(j + 3) * 32) : z4;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int off = h == 0 ? A.x : h == 1 ? A.z : h == 2 ? B.x : B.z;
                const float v = __int_as_float(h == 0 ? A.y : h == 1 ? A.w : h == 2 ? B.y : B.w);
                INMF_CHECK((uint32_t)off < (uint32_t)tile_rows * 64u && (off & 63) == 0);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t m[4];
                    lds128u(tb[c] + (uint32_t)off, m);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {  // bf16 pair -> two fp32: byte permute / mask, both on the ALU pipe
                        acc[c][2 * u] = fmaf(v, __uint_as_float(__byte_perm(m[u], 0u, 0x1044)), acc[c][2 * u]);
                        acc[c][2 * u + 1] = fmaf(v, __uint_as_float(m[u] & 0xffff0000u), acc[c][2 * u + 1]);
                    }
                }
            }
            A = A2;
            B = B2;
        }
        int sn = 0;
        if (lane == 0) sn = atomicAdd(next_slice, 1);
        sn = __shfl_sync(INMF_FULL_MASK, sn, 0);
        int64_t nbase = 0, nlim = 0;
        if (sn < slice_end) {
            nbase = slice_ptr[sn];
            nlim = slice_ptr[sn + 1];
        }
        if (rid >= 0) {
            float* o = out + (out_row_base + rid) * ldo;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int col0 = ((lane + c) & 3) * 8;
                const float lo4[4] = {acc[c][0], acc[c][1], acc[c][2], acc[c][3]};
                const float hi4[4] = {acc[c][4], acc[c][5], acc[c][6], acc[c][7]};
                sell_out4(o, col0, k, lo4, vec_ok);
                sell_out4(o, col0 + 4, k, hi4, vec_ok);
            }
        }
        s = sn;
        base = nbase;
        lim = nlim;
    }
}

// H (fp32, cells x ld) -> bf16 rows of 32 AND H'H of the unrounded values, per segment (dataset), one streaming pass.
// A warp walks rows; lane l owns column l: it rounds its value (even lanes pack a pair), keeps the row in a per-warp
// shared line, and accumulates row l of the Gram, G[l][j] += h_l h_j, from 8 broadcast 128-bit loads per row.  The
// warps of a CTA are folded through shared memory before one double-precision atomicAdd per entry.
#define GRAMCVT_WARPS 8
__global__ void __launch_bounds__(GRAMCVT_WARPS * 32) to_bf16_gram_kernel(const float* __restrict__ H, int64_t ld,
                                                                          const int64_t* __restrict__ seg_off, int k,
                                                                          uint32_t* __restrict__ dst,
                                                                          double* __restrict__ gram) {
    __shared__ __align__(16) float rowbuf[GRAMCVT_WARPS][2][32];
    __shared__ float fold[32][33];
    const int seg = blockIdx.y;
    const int64_t r0 = seg_off[seg], r1 = seg_off[seg + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    int buf = 0;
    for (int64_t r = r0 + (int64_t)blockIdx.x * GRAMCVT_WARPS + warp; r < r1; r += (int64_t)gridDim.x * GRAMCVT_WARPS) {
        const float h = lane < k ? H[r * ld + lane] : 0.f;
        const uint32_t u = __float_as_uint(h);
        const uint32_t rb = (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;   // finite inputs: round to nearest even
        const uint32_t nb = __shfl_down_sync(INMF_FULL_MASK, rb, 1);
        if ((lane & 1) == 0) dst[r * 16 + (lane >> 1)] = rb | (nb << 16);
        rowbuf[warp][buf][lane] = h;
        __syncwarp();
        const float4* rv = reinterpret_cast<const float4*>(rowbuf[warp][buf]);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 v = rv[q];
            acc[4 * q] = fmaf(h, v.x, acc[4 * q]);
            acc[4 * q + 1] = fmaf(h, v.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(h, v.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(h, v.w, acc[4 * q + 3]);
        }
        buf ^= 1;
    }
    for (int w = 0; w < GRAMCVT_WARPS; ++w) {
        if (warp == w) {
#pragma unroll
            for (int j = 0; j < 32; ++j) fold[lane][j] = (w == 0 ? 0.f : fold[lane][j]) + acc[j];
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
        const int i = e / k, j = e - i * k;
        atomicAdd(&gram[(size_t)seg * k * k + e], (double)fold[i][j]);
    }
}

// fp32 rows (ld floats apart) -> bf16 rows of 32 (round to nearest even), zero padded beyond k.
__global__ void __launch_bounds__(256) to_bf16_rows_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int k,
                                                           uint32_t* __restrict__ dst) {
    const int64_t total = rows * 16;  // one thread per bf16 PAIR
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t >> 4;
        const int c = (int)(t & 15) * 2;
        const float a = c < k ? src[r * ld + c] : 0.f;
        const float b = c + 1 < k ? src[r * ld + c + 1] : 0.f;
        const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
        const uint32_t ra = (ua + 0x7fffu + ((ua >> 16) & 1u)) >> 16;   // finite inputs: RNE on the dropped 16 bits
        const uint32_t rb = (ub + 0x7fffu + ((ub >> 16) & 1u)) >> 16;
        dst[t] = ra | (rb << 16);
    }
}

// Builder: copies the rows listed in slice_rows (ptr indices of the row-ordered blocked format, -1 = no row) into
// the step-interleaved slices of HEIGHT rows.  One warp per slice; HEIGHT = 8: lane = (row i = lane / 4, step-pair
// phase c = lane % 4); HEIGHT = 32: lane = row.
// Parity-ordered variant for the mixed-precision kernel (64-byte tile rows, HEIGHT = 32): lane = row; the row's
// non-zeros whose tile row has parity == bit 2 of the lane come first, the others after, each in their original order.
__global__ void __launch_bounds__(256) sell_fill_parity_kernel(const int64_t* __restrict__ ptr,
                                                               const int2* __restrict__ src,
                                                               const int64_t* __restrict__ slice_rows,
                                                               const int64_t* __restrict__ slice_ptr, int64_t nslices,
                                                               int4* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int want = (lane >> 2) & 1;
    for (int64_t s = warp0; s < nslices; s += nwarp) {
        const int64_t base = slice_ptr[s];
        const int L2 = (int)((slice_ptr[s + 1] - base) >> 5);
        const int64_t row = slice_rows[s * 32 + lane];
        int64_t beg = 0;
        int len = 0;
        if (row >= 0) {
            beg = ptr[row];
            len = (int)(ptr[row + 1] - beg);
        }
        int outpos = 0;
        int2 hold = make_int2(0, 0);
        for (int pass = 0; pass < 2; ++pass) {
            for (int t = 0; t < len; ++t) {
                const int2 a = src[beg + t];
                const int par = (a.x >> 6) & 1;
                if ((par == want) != (pass == 0)) continue;
                if (outpos & 1) dst[base + (int64_t)(outpos >> 1) * 32 + lane] = make_int4(hold.x, hold.y, a.x, a.y);
                else hold = a;
                ++outpos;
            }
        }
        if (outpos & 1) {
            dst[base + (int64_t)(outpos >> 1) * 32 + lane] = make_int4(hold.x, hold.y, 0, 0);
            ++outpos;
        }
        for (int jj = outpos >> 1; jj < L2; ++jj) dst[base + (int64_t)jj * 32 + lane] = make_int4(0, 0, 0, 0);
    }
}

// Residue-ordered variant for sell32r_pass_kernel (HEIGHT = 32, lane = row): at step j the lane emits a non-zero whose
// tile row satisfies (row mod 4) == (lane / 2 + j) mod 4 if its row still has one (in original order within the
// class), else the next one of the class with the most left.  off = tile row * 128.
__global__ void __launch_bounds__(256) sell_fill_residue_kernel(const int64_t* __restrict__ ptr,
                                                                const int2* __restrict__ src,
                                                                const int64_t* __restrict__ slice_rows,
                                                                const int64_t* __restrict__ slice_ptr, int64_t nslices,
                                                                int4* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = warp0; s < nslices; s += nwarp) {
        const int64_t base = slice_ptr[s];
        const int L2 = (int)((slice_ptr[s + 1] - base) >> 5);
        const int64_t row = slice_rows[s * 32 + lane];
        int64_t beg = 0;
        int len = 0;
        if (row >= 0) {
            beg = ptr[row];
            len = (int)(ptr[row + 1] - beg);
        }
        int left[4] = {0, 0, 0, 0}, cur[4] = {0, 0, 0, 0};
        for (int t = 0; t < len; ++t) ++left[(src[beg + t].x >> 7) & 3];
        int2 hold = make_int2(0, 0);
        for (int j = 0; j < len; ++j) {
            int c = ((lane >> 1) + j) & 3;
            if (left[c] == 0) {
                int best = 0;
#pragma unroll
                for (int q = 1; q < 4; ++q)
                    if (left[q] > left[best]) best = q;
                c = best;
            }
            int t = cur[c];
            while (((src[beg + t].x >> 7) & 3) != c) ++t;  // next non-zero of class c (left[c] > 0 guarantees one)
            const int2 a = src[beg + t];
            cur[c] = t + 1;
            --left[c];
            if (j & 1) dst[base + (int64_t)(j >> 1) * 32 + lane] = make_int4(hold.x, hold.y, a.x, a.y);
            else hold = a;
        }
        int outpos = len;
        if (outpos & 1) {
            dst[base + (int64_t)(outpos >> 1) * 32 + lane] = make_int4(hold.x, hold.y, 0, 0);
            ++outpos;
        }
        for (int jj = outpos >> 1; jj < L2; ++jj) dst[base + (int64_t)jj * 32 + lane] = make_int4(0, 0, 0, 0);
    }
}

template <int HEIGHT>
__global__ void __launch_bounds__(256) sell_fill_kernel(const int64_t* __restrict__ ptr, const int2* __restrict__ src,
                                                        const int64_t* __restrict__ slice_rows,
                                                        const int64_t* __restrict__ slice_ptr, int64_t nslices,
                                                        int4* __restrict__ dst) {
    constexpr int LPR = 32 / HEIGHT;  // lanes per row
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int i = lane / LPR, c = lane % LPR;
    for (int64_t s = warp0; s < nslices; s += nwarp) {
        const int64_t base = slice_ptr[s];
        const int L2 = (int)((slice_ptr[s + 1] - base) / HEIGHT);
        const int64_t row = slice_rows[s * HEIGHT + i];
        int64_t beg = 0;
        int len = 0;
        if (row >= 0) {
            beg = ptr[row];
            len = (int)(ptr[row + 1] - beg);
        }
        for (int jj = c; jj < L2; jj += LPR) {
            int4 q = make_int4(0, 0, 0, 0);
            if (2 * jj < len) {
                const int2 a = src[beg + 2 * jj];
                q.x = a.x;
                q.y = a.y;
            }
            if (2 * jj + 1 < len) {
                const int2 b = src[beg + 2 * jj + 1];
                q.z = b.x;
                q.w = b.y;
            }
            dst[base + (int64_t)jj * HEIGHT + i] = q;
        }
    }
}
