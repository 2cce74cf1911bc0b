// The O(nnz) passes that produce scale_data from raw counts (the stage just before the factorisation):
//   normalize          /root/reference/src/pyliger/preprocessing/_normalization.py:91-99
//   scale_not_center   /root/reference/src/pyliger/preprocessing/_scale.py:92-102
// CSR in, CSR out, one warp per cell (row).  HBM-bound streaming kernels: every non-zero is read once and
// written once per pass; the per-gene sums are double atomics (genes << non-zeros, they stay in L2).
#pragma once
#include "inmf_common.cuh"

// Row-L1 normalisation (sklearn.preprocessing.normalize(norm="l1"): v / sum|v|, rows with zero norm are left
// as they are) fused with the per-gene sum and sum of squares of the normalised values, and the count of rows
// whose plain sum is zero (_remove_missing_obs, /root/reference/src/pyliger/_utilities.py:74-75).
template <typename T>
__global__ void __launch_bounds__(256) normalize_rows_kernel(const int64_t* __restrict__ rowptr,
                                                             const int32_t* __restrict__ colidx,
                                                             const T* __restrict__ vals, int64_t n_rows,
                                                             T* __restrict__ out, double* __restrict__ col_sum,
                                                             double* __restrict__ col_sumsq,
                                                             unsigned long long* __restrict__ n_missing) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += nwarp) {
        const int64_t b = rowptr[r], e = rowptr[r + 1];
        T sa = T(0), ss = T(0);
        for (int64_t p = b + lane; p < e; p += 32) {
            const T v = vals[p];
            sa += inmf_abs(v);
            ss += v;
        }
        sa = warp_sum(sa);
        ss = warp_sum(ss);
        if (lane == 0 && ss == T(0) && n_missing != nullptr) atomicAdd(n_missing, 1ull);
        const T norm = (sa == T(0)) ? T(1) : sa;
        for (int64_t p = b + lane; p < e; p += 32) {
            const T o = vals[p] / norm;
            out[p] = o;
            const int32_t c = colidx[p];
            atomicAdd(col_sum + c, (double)o);
            atomicAdd(col_sumsq + c, (double)o * (double)o);
        }
    }
}

// Gene selection, step 1: non-zeros per row that survive the column map (colmap[c] = new index or -1).
__global__ void __launch_bounds__(256) select_count_kernel(const int64_t* __restrict__ rowptr,
                                                           const int32_t* __restrict__ colidx, int64_t n_rows,
                                                           const int32_t* __restrict__ colmap,
                                                           int64_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += nwarp) {
        const int64_t b = rowptr[r], e = rowptr[r + 1];
        int cnt = 0;
        for (int64_t p = b + lane; p < e; p += 32) cnt += colmap[colidx[p]] >= 0;
        cnt = __reduce_add_sync(INMF_FULL_MASK, cnt);
        if (lane == 0) counts[r] = cnt;
    }
}

// Step 2: order-preserving compaction of the kept columns, values multiplied by the per-gene scaler
// (sklearn inplace_column_scale: data *= scale.take(indices)).
template <typename T>
__global__ void __launch_bounds__(256) select_scale_kernel(const int64_t* __restrict__ rowptr,
                                                           const int32_t* __restrict__ colidx,
                                                           const T* __restrict__ vals, int64_t n_rows,
                                                           const int32_t* __restrict__ colmap,
                                                           const double* __restrict__ scaler,
                                                           const int64_t* __restrict__ new_rowptr,
                                                           int32_t* __restrict__ new_colidx,
                                                           T* __restrict__ new_vals) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += nwarp) {
        const int64_t b = rowptr[r], e = rowptr[r + 1];
        int64_t o = new_rowptr[r];
        for (int64_t p0 = b; p0 < e; p0 += 32) {
            const int64_t p = p0 + lane;
            int32_t nc = -1;
            T v = T(0);
            if (p < e) {
                nc = colmap[colidx[p]];
                v = vals[p];
            }
            const unsigned keep = __ballot_sync(INMF_FULL_MASK, nc >= 0);
            if (nc >= 0) {
                const int64_t q = o + __popc(keep & lt);
                new_colidx[q] = nc;
                new_vals[q] = (T)((double)v * scaler[nc]);
            }
            o += __popc(keep);
        }
    }
}
