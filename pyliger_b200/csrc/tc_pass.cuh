// Dense-tile (tcgen05) form of the two X-streaming passes, fp32 data with 3xTF32 (fp32-grade) or single
// TF32 arithmetic.
//
//   out[tile rows, 0..k) (+)= X_tile (128 x contraction) * D (contraction x k)
//
// pass 1: rows = cells, contraction = genes,  D = Mt = (W+V)'  ->  R   (plain stores, no atomics)
// pass 2: rows = genes, contraction = cells,  D = H            ->  St  (RED.ADD when a tile's cell range is split)
//
// The CUDA-core form of these passes (blocked.cuh) is bound by the shared-memory operand bandwidth (each
// non-zero gathers k operands).  Here the sparse tile is DENSIFIED in shared memory instead: per chunk of
// 64 contraction entities the non-zeros of a 128 x 64 tile are scattered into a zeroed 32 KB operand tile
// (one 4-byte store per non-zero and precision term, and one more later to clear it again -- the tile is
// never re-zeroed wholesale), and the contraction itself runs on the tensor core with the accumulator in
// TMEM, so the per-non-zero shared-memory traffic drops from 4k + 8 bytes to ~16 bytes.
//
// Warp roles (320 threads, one CTA per SM, persistent over a static list of tiles):
//   warp 0 lane 0     MMA issuer: waits for a filled A stage and its B chunk, issues 8 (x3) tcgen05.mma,
//                     commits to the stage's "empty" barrier; after the last chunk commits to "accum"
//   warps 1..8        producers: wait "empty", clear the stage's previous non-zeros, scatter the new chunk
//                     (hi / lo TF32 split), fence.proxy.async, arrive on "full"; after the tile's last chunk
//                     warps 1..4 read the accumulator (tcgen05.ld) and write the output rows
//   warp 9 lane 0     B loader: TMA bulk copies of the pre-laid-out dense-operand chunks (hi and lo)
//
// Precision: v = hi + lo with hi = v with the low 13 mantissa bits cleared (exactly a TF32 number);
// D += A_hi B_hi + A_lo B_hi + A_hi B_lo  (3xTF32).  Measured error of the 128 x 64 x 32 self test: 1e-6.
#pragma once
#include "tc_common.cuh"

#define TC_STAGES 2                   // A (scatter) stages
#define TC_MAX_BSTAGES 4              // B (dense operand chunk) stages, fewer when the chunks are wide
#define TC_NPW 8                      // producer warps
#define TC_THREADS ((TC_NPW + 2) * 32)
#define TC_WORK_FIELDS 8
#ifdef TC_DEBUG_TIMERS   // per-role cycle counters dumped through inmf_tc_debug_buffer (profiling builds only)
#define TC_T(...) __VA_ARGS__
#define TC_DBG_ON 1
#else
#define TC_T(...)
#define TC_DBG_ON 0
#endif
#define TC_PREF 6                     // pairs per lane kept in registers per (band, chunk): 192 per band
#define TC_SLOTS 4                    // register slots: pairs are loaded TC_SLOTS chunk periods ahead of use
// work[w] = { ptr_index0, dense_chunk0, nchunks, out_row0, tile_rows, atomic, unused, unused }
// ptr has TC_NPW entries per (tile, chunk): the pairs of the 16-row bands 0..7 (plus one final entry).

struct TcSmemLayout {
    uint32_t a_hi[TC_STAGES], a_lo[TC_STAGES], b_hi[TC_MAX_BSTAGES], b_lo[TC_MAX_BSTAGES];  // byte offsets
    uint32_t bars, slot, total;
    int nbs;
};
__host__ __device__ inline TcSmemLayout tc_smem_layout(int npad) {
    TcSmemLayout L;
    uint32_t off = 0;
    const uint32_t bbytes = (uint32_t)npad * TC_KC * 4;
    L.nbs = npad <= 32 ? 4 : (npad <= 48 ? 3 : 2);
    for (int s = 0; s < TC_STAGES; ++s) { L.a_hi[s] = off; off += TC_A_BYTES; }
    for (int s = 0; s < TC_STAGES; ++s) { L.a_lo[s] = off; off += TC_A_BYTES; }
    for (int s = 0; s < TC_MAX_BSTAGES; ++s) { L.b_hi[s] = off; if (s < L.nbs) off += bbytes; }
    for (int s = 0; s < TC_MAX_BSTAGES; ++s) { L.b_lo[s] = off; if (s < L.nbs) off += bbytes; }
    L.bars = off; off += 8 * (2 * TC_STAGES + 2 * TC_MAX_BSTAGES + 1);
    L.slot = off; off += 16;
    L.total = off;
    return L;
}

// Re-layout of the dense operand into per-chunk UMMA B tiles (K-major, no swizzle), split into TF32 hi / lo.
// segd[s] = { first row of the segment in D, rows in the segment, first chunk index of the segment }
__global__ void tc_relayout_kernel(const float* __restrict__ D, int64_t ldd, const int64_t* __restrict__ segd,
                                   float* __restrict__ out_hi, float* __restrict__ out_lo, int k, int npad) {
    const int s = blockIdx.y;
    const int64_t row0 = segd[3 * s], rows = segd[3 * s + 1], chunk0 = segd[3 * s + 2];
    const int64_t nch = (rows + TC_KC - 1) / TC_KC;
    const int per = npad * TC_KC;
    for (int64_t c = blockIdx.x; c < nch; c += gridDim.x) {
        float* oh = out_hi + (chunk0 + c) * per;
        float* ol = out_lo + (chunk0 + c) * per;
        for (int e = threadIdx.x; e < per; e += blockDim.x) {
            // e enumerates the destination: [n/8][k/4][n%8][k%4]
            const int k4 = e & 3, n8 = (e >> 2) & 7, kq = (e >> 5) % (TC_KC / 4), ng = e / (32 * (TC_KC / 4));
            const int n = ng * 8 + n8, kk = kq * 4 + k4;
            const int64_t r = c * TC_KC + kk;
            float v = 0.f;
            if (r < rows && n < k) v = D[(row0 + r) * ldd + n];
            const float h = tf32_hi(v);
            oh[e] = h;
            ol[e] = v - h;
        }
    }
}

// Flattened (tile, chunk) sequence of one CTA.
struct TcSeq {
    const int64_t* work;
    int nwork, w, c, nch;
    int64_t pidx0;
    __device__ __forceinline__ void init(const int64_t* wk, int nw) {
        work = wk; nwork = nw; w = blockIdx.x; c = 0; load();
    }
    __device__ __forceinline__ void load() {
        if (w < nwork) { pidx0 = work[(size_t)w * TC_WORK_FIELDS]; nch = (int)work[(size_t)w * TC_WORK_FIELDS + 2]; }
        else { pidx0 = 0; nch = 0; }
    }
    __device__ __forceinline__ bool valid() const { return w < nwork; }
    __device__ __forceinline__ void next() {
        if (++c >= nch) { c = 0; w += gridDim.x; load(); }
    }
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_pass_kernel(const int64_t* __restrict__ work, int nwork, const int64_t* __restrict__ ptr,
               const int2* __restrict__ pairs, const float* __restrict__ dc_hi, const float* __restrict__ dc_lo,
               float* __restrict__ out, int64_t ldo, int k, int npad, int three_pass, long long* __restrict__ dbg) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const TcSmemLayout L = tc_smem_layout(npad);
    const int NBS = L.nbs;
    TC_T(long long t_a = 0, t_b = 0, t_c = 0, t_d = 0;)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bars);
    uint64_t* full = bars;                                        // [S]   producers -> MMA
    uint64_t* empty = bars + TC_STAGES;                           // [S]   MMA (commit) -> producers
    uint64_t* bfull = bars + 2 * TC_STAGES;                       // [NBS] TMA -> MMA
    uint64_t* bempty = bars + 2 * TC_STAGES + TC_MAX_BSTAGES;     // [NBS] MMA (commit) -> B loader
    uint64_t* accum = bars + 2 * TC_STAGES + 2 * TC_MAX_BSTAGES;  //       MMA (commit) -> epilogue
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem_raw + L.slot);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = smem_u32(smem_raw);
    const uint32_t bbytes = (uint32_t)npad * TC_KC * 4;

    // one-time setup: zero the A stages, barriers, TMEM
    for (uint32_t e = tid; e < 2u * TC_STAGES * TC_A_BYTES / 16; e += blockDim.x)
        reinterpret_cast<uint4*>(smem_raw + L.a_hi[0])[e] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full[s], TC_NPW);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < TC_MAX_BSTAGES; ++s) {
            mbar_init(&bfull[s], 1);
            mbar_init(&bempty[s], 1);
        }
        mbar_init(accum, 1);
    }
    fence_async_smem();
    if (warp == 0) tmem_alloc(slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;

    if (warp == 0) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32_m128(npad);
            // descriptors of every stage at k-step 0, built once; a k-step advances the 16-byte-unit start
            // address field (low word) by 256 B / 16 -- the issue loop is a handful of instructions per MMA
            uint64_t dA_hi[TC_STAGES], dA_lo[TC_STAGES], dB_hi[TC_MAX_BSTAGES], dB_lo[TC_MAX_BSTAGES];
#pragma unroll
            for (int s = 0; s < TC_STAGES; ++s) {
                dA_hi[s] = umma_smem_desc(smem_base + L.a_hi[s], TC_A_LBO, TC_A_SBO);
                dA_lo[s] = umma_smem_desc(smem_base + L.a_lo[s], TC_A_LBO, TC_A_SBO);
            }
#pragma unroll
            for (int s = 0; s < TC_MAX_BSTAGES; ++s) {
                dB_hi[s] = umma_smem_desc(smem_base + L.b_hi[s], TC_B_LBO, TC_B_SBO);
                dB_lo[s] = umma_smem_desc(smem_base + L.b_lo[s], TC_B_LBO, TC_B_SBO);
            }
            uint32_t it = 0;
            for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
                const int nch = (int)work[(size_t)w * TC_WORK_FIELDS + 2];
                for (int c = 0; c < nch; ++c, ++it) {
                    const int s = it % TC_STAGES, bs = it % NBS;
                    uint64_t ah = 0, al = 0, bh = 0, bl = 0;
#pragma unroll
                    for (int q = 0; q < TC_STAGES; ++q)
                        if (q == s) { ah = dA_hi[q]; al = dA_lo[q]; }
#pragma unroll
                    for (int q = 0; q < TC_MAX_BSTAGES; ++q)
                        if (q == bs) { bh = dB_hi[q]; bl = dB_lo[q]; }
                    TC_T(long long t0 = clock64();)
                    mbar_wait(&bfull[bs], (it / NBS) & 1);
                    TC_T(long long t1 = clock64();)
                    mbar_wait(&full[s], (it / TC_STAGES) & 1);
                    TC_T(long long t2 = clock64();)
                    TC_T(t_a += t1 - t0;)
                    TC_T(t_b += t2 - t1;)
                    tc_fence_after();
                    if (three_pass) {
#pragma unroll
                        for (int ks = 0; ks < TC_KC / 8; ++ks) {
                            umma_tf32(tmem, ah + ks * 16, bh + ks * 16, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                            umma_tf32(tmem, al + ks * 16, bh + ks * 16, idesc, 1u);
                            umma_tf32(tmem, ah + ks * 16, bl + ks * 16, idesc, 1u);
                        }
                    } else {
#pragma unroll
                        for (int ks = 0; ks < TC_KC / 8; ++ks)
                            umma_tf32(tmem, ah + ks * 16, bh + ks * 16, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);
                    umma_commit(&bempty[bs]);
                    TC_T(t_c += clock64() - t2;)
                }
                umma_commit(accum);
            }
            TC_T(if (dbg) { dbg[blockIdx.x * 64 + 0] = t_a; dbg[blockIdx.x * 64 + 1] = t_b; dbg[blockIdx.x * 64 + 2] = t_c; })
        }
    } else if (warp == TC_NPW + 1) {
        // ===================== B loader =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
                const int64_t dchunk0 = work[(size_t)w * TC_WORK_FIELDS + 1];
                const int nch = (int)work[(size_t)w * TC_WORK_FIELDS + 2];
                for (int c = 0; c < nch; ++c, ++it) {
                    const int bs = it % NBS;
                    const uint32_t u = it / NBS;
                    TC_T(long long t0 = clock64();)
                    if (u >= 1) mbar_wait(&bempty[bs], (u - 1) & 1);
                    TC_T(t_a += clock64() - t0;)
                    mbar_expect_tx(&bfull[bs], three_pass ? 2 * bbytes : bbytes);
                    const size_t goff = (size_t)(dchunk0 + c) * ((size_t)npad * TC_KC);
                    bulk_g2s(smem_raw + L.b_hi[bs], dc_hi + goff, bbytes, &bfull[bs]);
                    if (three_pass) bulk_g2s(smem_raw + L.b_lo[bs], dc_lo + goff, bbytes, &bfull[bs]);
                }
            }
            TC_T(if (dbg) dbg[blockIdx.x * 64 + 3] = t_a;)
        }
    } else {
        // ===================== producers / epilogue =====================
        // Warp pw owns the 16-row band pw of every tile: no two warps ever touch the same shared address,
        // so clearing the previous chunk and scattering the next need no cross-warp barrier.  The pairs of
        // the NEXT chunk are loaded into registers before the current chunk's barrier is waited on, and
        // the offsets to clear are remembered in registers, so the loop body never waits on global memory.
        const int pw = warp - 1;
        uint32_t it = 0, tile_no = 0;
        uint32_t oldoff[TC_STAGES][TC_PREF];
        int64_t old_xbeg[TC_STAGES], old_xend[TC_STAGES];  // overflow ranges (more than 32*TC_PREF pairs)
#pragma unroll
        for (int s = 0; s < TC_STAGES; ++s) {
            old_xbeg[s] = old_xend[s] = 0;
#pragma unroll
            for (int j = 0; j < TC_PREF; ++j) oldoff[s][j] = 0xffffffffu;
        }
        // Register pipeline with STATIC slots (the chunk loop is unrolled by the number of slots, so no
        // register is ever moved): slot j holds the pairs of the chunk that will be scattered when the
        // unrolled body reaches j again, i.e. TC_SLOTS chunk periods after the loads were issued; the band
        // pointers those loads need were fetched another TC_SLOTS periods earlier.
        TcSeq cur, far;  // far runs 2 * TC_SLOTS chunks ahead of cur (pointer prefetch)
        cur.init(work, nwork);
        far = cur;
        int2 qp[TC_SLOTS][TC_PREF];
        int64_t qb[TC_SLOTS], qe[TC_SLOTS], pb[TC_SLOTS], pe[TC_SLOTS];
        auto load_ptr = [&](const TcSeq& q, int64_t& b, int64_t& e) {
            if (q.valid()) {
                b = ptr[(q.pidx0 + q.c) * TC_NPW + pw];
                e = ptr[(q.pidx0 + q.c) * TC_NPW + pw + 1];
            } else {
                b = e = 0;
            }
        };
        auto load_pairs = [&](int2 (&dst)[TC_PREF], int64_t b, int64_t e) {
#pragma unroll
            for (int j = 0; j < TC_PREF; ++j) {
                const int64_t i = b + lane + 32 * j;
                dst[j] = (i < e) ? PairTraits<float>::ld(pairs + i) : make_int2(-1, 0);
            }
        };
#pragma unroll
        for (int j = 0; j < TC_SLOTS; ++j) {
            load_ptr(far, qb[j], qe[j]);
            far.next();
        }
#pragma unroll
        for (int j = 0; j < TC_SLOTS; ++j) {
            load_pairs(qp[j], qb[j], qe[j]);
            load_ptr(far, pb[j], pe[j]);
            far.next();
        }
        TC_T(long long t_e = 0, t_f = 0, t_last = clock64();)
        while (cur.valid()) {
#pragma unroll
            for (int j = 0; j < TC_SLOTS; ++j) {
                if (!cur.valid()) break;
                const int64_t* wd = work + (size_t)cur.w * TC_WORK_FIELDS;
                const int64_t xbeg = qb[j] + 32 * TC_PREF, xend = qe[j];
                const int s = it % TC_STAGES;
                const uint32_t u = it / TC_STAGES;
                TC_T(long long t0 = clock64();)
                TC_T(t_f += t0 - t_last;)
                if (u >= 1) mbar_wait(&empty[s], (u - 1) & 1);
                TC_T(long long t1 = clock64();)
                TC_T(t_a += t1 - t0;)
                unsigned char* ah = smem_raw + L.a_hi[s];
                unsigned char* al = smem_raw + L.a_lo[s];
#pragma unroll
                for (int q = 0; q < TC_STAGES; ++q) {
                    if (q == s) {
                        // clear the previous occupant's entries
#pragma unroll
                        for (int t = 0; t < TC_PREF; ++t) {
                            const uint32_t off = oldoff[q][t];
                            if (off != 0xffffffffu) {
                                *reinterpret_cast<float*>(ah + off) = 0.f;
                                if (three_pass) *reinterpret_cast<float*>(al + off) = 0.f;
                            }
                        }
                        for (int64_t i = old_xbeg[q] + lane; i < old_xend[q]; i += 32) {
                            const uint32_t off = (uint32_t)pairs[i].x;
                            *reinterpret_cast<float*>(ah + off) = 0.f;
                            if (three_pass) *reinterpret_cast<float*>(al + off) = 0.f;
                        }
                        // scatter this chunk and remember where
#pragma unroll
                        for (int t = 0; t < TC_PREF; ++t) {
                            const uint32_t off = (uint32_t)qp[j][t].x;
                            oldoff[q][t] = off;
                            if (off != 0xffffffffu) {
                                const float v = __int_as_float(qp[j][t].y), h = tf32_hi(v);
                                *reinterpret_cast<float*>(ah + off) = h;
                                if (three_pass) *reinterpret_cast<float*>(al + off) = v - h;
                            }
                        }
                        for (int64_t i = xbeg + lane; i < xend; i += 32) {
                            const int2 pr = PairTraits<float>::ld(pairs + i);
                            const float v = __int_as_float(pr.y), h = tf32_hi(v);
                            *reinterpret_cast<float*>(ah + (uint32_t)pr.x) = h;
                            if (three_pass) *reinterpret_cast<float*>(al + (uint32_t)pr.x) = v - h;
                        }
                        old_xbeg[q] = xbeg < xend ? xbeg : 0;
                        old_xend[q] = xbeg < xend ? xend : 0;
                    }
                }
                TC_T(long long t2 = clock64();)
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
                TC_T(long long t3 = clock64();)
                TC_T(t_b += t2 - t1;)
                TC_T(t_c += t3 - t2;)
                ++it;
                // refill this slot: pairs of the chunk TC_SLOTS ahead, pointers of the one 2*TC_SLOTS ahead
                qb[j] = pb[j];
                qe[j] = pe[j];
                load_pairs(qp[j], qb[j], qe[j]);
                load_ptr(far, pb[j], pe[j]);
                far.next();
                TC_T(t_e += clock64() - t3;)

                if (cur.c == cur.nch - 1) {
                    // ---- last chunk of the tile: accumulator -> output rows ----
                    const int64_t out_row0 = wd[3];
                    const int tile_rows = (int)wd[4];
                    const int atomic = (int)wd[5];
                    TC_T(long long t4 = clock64();)
                    mbar_wait(accum, tile_no & 1);
                    TC_T(t_d += clock64() - t4;)
                    tc_fence_after();
                    if (pw < 4) {
                        const int q4 = warp & 3;  // TMEM lane quarter this warp may access
                        const int r = q4 * 32 + lane;
                        uint32_t v[32];
                        for (int c0 = 0; c0 < npad; c0 += 32) {
                            tmem_ld_32x32(tmem + ((uint32_t)(q4 * 32) << 16) + c0, v);
                            if (r < tile_rows) {
                                float* o = out + (out_row0 + r) * ldo + c0;
                                if (atomic) {
#pragma unroll
                                    for (int t = 0; t < 32; ++t)
                                        if (c0 + t < k && v[t] != 0u) atomicAdd(o + t, __uint_as_float(v[t]));
                                } else {
#pragma unroll
                                    for (int t = 0; t < 32; ++t)
                                        if (c0 + t < k) o[t] = __uint_as_float(v[t]);
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    ++tile_no;
                }
                cur.next();
                TC_T(t_last = clock64();)
            }
        }
        if (TC_DBG_ON && dbg && lane == 0) {
            TC_T(dbg[blockIdx.x * 64 + 16 + pw] = t_b;)
            TC_T(dbg[blockIdx.x * 64 + 24 + pw] = t_e;)
            TC_T(dbg[blockIdx.x * 64 + 32 + pw] = t_f;)
            TC_T(dbg[blockIdx.x * 64 + 40 + pw] = t_a;)
        }
        if (TC_DBG_ON && dbg && pw == 0 && lane == 0) {
            TC_T(dbg[blockIdx.x * 64 + 4] = t_a; dbg[blockIdx.x * 64 + 5] = t_b; dbg[blockIdx.x * 64 + 6] = t_c;)
            TC_T(dbg[blockIdx.x * 64 + 7] = t_d; dbg[blockIdx.x * 64 + 8] = it;)
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}
