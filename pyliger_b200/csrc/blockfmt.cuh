// Device builders of the two blocked copies of X (see blocked.cuh / pyliger_b200/blocked.py):
//   format 1  gene-blocked CSR   pairs ordered by (dataset, gene block, cell, gene)
//   format 2  cell-blocked CSC   pairs ordered by (dataset, cell block, gene, cell)
// Both are stable counting sorts of the CSR non-zeros, done as  count -> (exclusive scan by the caller) -> fill,
// deterministic (no order-dependent atomics), each pass reading X once.  The generic torch sort path in
// blocked.py produces bit-identical arrays and remains the fallback (and the test reference).
#pragma once
#include "inmf_common.cuh"

// ---- format 1 ----------------------------------------------------------------------------------------
// seg1[s] = { own_off, n_own, csr_row0, unused }: owned rows [own_off, own_off + n_own) of the H buffer are CSR
// rows csr_row0 + local_row.  ptr index of (s, gene block b, local_row) = NB * own_off + b * n_own + local_row.
__device__ __forceinline__ int find_seg(const int64_t* __restrict__ seg, int nseg, int64_t i) {
    int s = 0;
    while (s + 1 < nseg && i >= seg[4 * (s + 1)]) ++s;
    return s;
}

template <bool FILL, typename T>
__global__ void __launch_bounds__(256) blockfmt1_kernel(const int64_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ colidx,
                                                        const T* __restrict__ vals,
                                                        const int64_t* __restrict__ seg1, int nseg, int64_t n_rows,
                                                        int NB, int gbs, int row_bytes,
                                                        int64_t* __restrict__ cnt_or_ptr,
                                                        typename PairTraits<T>::Pair* __restrict__ pairs) {
    typedef typename PairTraits<T>::Pair Pair;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < n_rows; i += nwarp) {
        const int s = find_seg(seg1, nseg, i);
        const int64_t own_off = seg1[4 * s], n_own = seg1[4 * s + 1], lr = i - own_off;
        const int64_t r = seg1[4 * s + 2] + lr;
        const int64_t key0 = (int64_t)NB * own_off + lr;
        const int64_t b0 = rowptr[r], e0 = rowptr[r + 1];
        for (int64_t p0 = b0; p0 < e0; p0 += 32) {
            const int64_t p = p0 + lane;
            const bool on = p < e0;
            const int c = on ? colidx[p] : 0;
            const int b = on ? c / gbs : -1;
            if (!FILL) {
                // sorted columns: the blocks present in this chunk are bmin..bmax
                const int bmin = __shfl_sync(INMF_FULL_MASK, b, 0);
                const int nact = (int)((e0 - p0 < 32) ? (e0 - p0) : 32);
                const int bmax = __shfl_sync(INMF_FULL_MASK, b, nact - 1);
                for (int bb = bmin; bb <= bmax; ++bb) {
                    const int m = __popc(__ballot_sync(INMF_FULL_MASK, b == bb));
                    if (lane == 0 && m) cnt_or_ptr[key0 + (int64_t)bb * n_own] += m;  // only this warp owns the row
                }
            } else if (on) {
                int64_t below = 0;  // non-zeros of this row in earlier gene blocks
                for (int bb = 0; bb < b; ++bb) {
                    const int64_t kk = key0 + (int64_t)bb * n_own;
                    below += cnt_or_ptr[kk + 1] - cnt_or_ptr[kk];
                }
                const int64_t q = cnt_or_ptr[key0 + (int64_t)b * n_own] + (p - b0 - below);
                Pair pr;
                pr.x = (decltype(pr.x))((int64_t)(c - b * gbs) * row_bytes);
                pr.y = PairTraits<T>::bits(vals[p]);
                pairs[q] = pr;
            }
        }
    }
}

// ---- format 2 ----------------------------------------------------------------------------------------
// blk2[j] = { csr_row0, rows, key_base (= g * j), unused } for cell block j.  One CTA per cell block, W warps;
// warp w owns the contiguous row range w of the block and walks it IN ORDER, so that with per-(warp, gene)
// counters the cells of every gene list come out in increasing order without any inter-warp ordering.
// Shared memory: uint16 counters [W][g].
template <bool FILL, typename T>
__global__ void blockfmt2_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                 const T* __restrict__ vals, const int64_t* __restrict__ blk2, int g, int row_bytes,
                                 int64_t* __restrict__ cnt, const int64_t* __restrict__ ptr,
                                 uint16_t* __restrict__ offs, typename PairTraits<T>::Pair* __restrict__ pairs) {
    typedef typename PairTraits<T>::Pair Pair;
    extern __shared__ uint16_t bf2_cnt[];
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t j = blockIdx.x;
    const int64_t row0 = blk2[4 * j], rows = blk2[4 * j + 1], key_base = blk2[4 * j + 2];
    uint16_t* __restrict__ mine = bf2_cnt + (size_t)warp * g;
    uint16_t* __restrict__ goffs = offs + (size_t)j * W * g;
    if (!FILL) {
        for (int t = threadIdx.x; t < W * g; t += blockDim.x) bf2_cnt[t] = 0;
    } else {
        for (int t = threadIdx.x; t < W * g; t += blockDim.x) bf2_cnt[t] = goffs[t];
    }
    __syncthreads();
    const int64_t per = (rows + W - 1) / W;
    const int64_t lo = warp * per, hi = (lo + per < rows) ? lo + per : rows;
    for (int64_t lr = lo; lr < hi; ++lr) {
        const int64_t b0 = rowptr[row0 + lr], e0 = rowptr[row0 + lr + 1];
        for (int64_t p0 = b0; p0 < e0; p0 += 32) {
            const int64_t p = p0 + lane;
            const bool on = p < e0;
            const int c = on ? colidx[p] : -1 - lane;  // inactive lanes never match anything
            // duplicate column entries in one row (non-canonical CSR) would collide: rank them
            const unsigned same = __match_any_sync(INMF_FULL_MASK, c);
            const int rank = __popc(same & ((1u << lane) - 1u));
            if (on) {
                const int base = mine[c];
                if (FILL) {
                    const int64_t q = ptr[key_base + c] + base + rank;
                    Pair pr;
                    pr.x = (decltype(pr.x))(lr * row_bytes);
                    pr.y = PairTraits<T>::bits(vals[p]);
                    pairs[q] = pr;
                }
                if (rank == 0) mine[c] = (uint16_t)(base + __popc(same));
            }
            __syncwarp();
        }
    }
    if (!FILL) {
        __syncthreads();
        // exclusive scan over the warps for every gene; the total goes to the global histogram
        for (int c = threadIdx.x; c < g; c += blockDim.x) {
            int run = 0;
            for (int w = 0; w < W; ++w) {
                const int t = bf2_cnt[(size_t)w * g + c];
                goffs[(size_t)w * g + c] = (uint16_t)run;
                run += t;
            }
            cnt[key_base + c] = run;
        }
    }
}
