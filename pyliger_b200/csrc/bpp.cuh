// Warp-per-column block principal pivoting NNLS (device side).
//
// Algorithm: Kim & Park 2011 as used by pyliger (/root/reference/src/pyliger/factorization/
// _utilities.py:226-299).  Columns are independent (the reference's passive-set grouping,
// _utilities.py:341-357, is only a CPU cost optimisation), so one warp owns one column:
//   * the passive set is a 64-bit mask held uniformly by the warp (k <= 64),
//   * infeasible indices are found with __ballot_sync,
//   * the passive-set system G_PP z = r_P is solved by a left-looking Cholesky on the compacted
//     (p+1) x p augmented matrix [G_PP ; r_P'] kept in a per-warp shared-memory scratch, the extra
//     row giving the forward substitution for free; back substitution is right-looking,
//   * y = G x - r is formed for the non-passive rows from the shared copy of G.
// The reference solves with LAPACK LU (numpy.linalg.solve); for symmetric positive definite G_PP the
// Cholesky solution is the same up to round-off.
#pragma once
#include "inmf_common.cuh"

template <typename T>
struct BppWarpScratch {
    T* L;          // (k+1) x ldl
    T* r;          // k   right-hand side of this column
    T* z;          // k   compacted solution (thresholded)
    T* dinv;       // k   1 / L[c][c]
    uint8_t* idx;  // 64  compacted position -> factor index
    int ldl;
};

template <typename T>
__host__ __device__ inline size_t bpp_warp_scratch_elems(int k) {
    int ldl = k | 1;
    return (size_t)(k + 1) * ldl + 3 * (size_t)k + 64 / sizeof(T) + 2;
}

template <typename T>
__device__ __forceinline__ BppWarpScratch<T> bpp_carve(T* base, int k) {
    BppWarpScratch<T> s;
    s.ldl = k | 1;
    s.L = base;
    s.r = s.L + (size_t)(k + 1) * s.ldl;
    s.z = s.r + k;
    s.dinv = s.z + k;
    s.idx = reinterpret_cast<uint8_t*>(s.dinv + k);
    return s;
}

struct BppCounters {
    int rounds;
    int nfact;
    int nbackup;
    int nfail;
    bool ok;
};

__device__ __forceinline__ bool mask_bit(uint64_t m, int j) { return (m >> j) & 1ull; }

// Solve G_PP z = r_P for the warp's column and update x (lane j, j+32), y.  P != 0 required.
template <typename T>
__device__ __forceinline__ void bpp_solve_passive(const T* __restrict__ Gs, int ldg, BppWarpScratch<T>& sc,
                                                  int k, uint64_t P, T r0, T r1, T& x0, T& x1, T& y0,
                                                  T& y1, BppCounters& cnt) {
    const int lane = threadIdx.x & 31;
    const int j0 = lane, j1 = lane + 32;
    const bool has0 = j0 < k, has1 = j1 < k;
    const int p = __popcll(P);
    const int ldl = sc.ldl;
    const uint64_t below0 = (1ull << j0) - 1ull;
    const uint64_t below1 = (1ull << j1) - 1ull;
    const bool pas0 = has0 && mask_bit(P, j0);
    const bool pas1 = has1 && mask_bit(P, j1);
    const int pos0 = __popcll(P & below0);
    const int pos1 = __popcll(P & below1);
    if (pas0) sc.idx[pos0] = (uint8_t)j0;
    if (pas1) sc.idx[pos1] = (uint8_t)j1;
    __syncwarp();

    // ---- left-looking Cholesky of the augmented system, column by column ----
    for (int c = 0; c < p; ++c) {
        const int ic = sc.idx[c];
        const T* Lc = sc.L + (size_t)c * ldl;
        int i = c + lane;
        T s = T(0);
        if (i <= p) {
            s = (i < p) ? Gs[(int)sc.idx[i] * ldg + ic] : sc.r[ic];
            const T* Li = sc.L + (size_t)i * ldl;
            for (int j = 0; j < c; ++j) s -= Li[j] * Lc[j];
        }
        const T d = __shfl_sync(INMF_FULL_MASK, s, 0);
        T di = T(0);
        if (d > T(0)) {
            di = inmf_rsqrt<T>(d);
        } else if (lane == 0) {
            cnt.nfail++;
        }
        if (i <= p) sc.L[(size_t)i * ldl + c] = s * di;
        if (lane == 0) sc.dinv[c] = di;
        for (i = c + lane + 32; i <= p; i += 32) {
            T s2 = (i < p) ? Gs[(int)sc.idx[i] * ldg + ic] : sc.r[ic];
            const T* Li = sc.L + (size_t)i * ldl;
            for (int j = 0; j < c; ++j) s2 -= Li[j] * Lc[j];
            sc.L[(size_t)i * ldl + c] = s2 * di;
        }
        __syncwarp();
    }

    // ---- right-looking back substitution: z = L^-T w, w = augmented row ----
    const T* Lw = sc.L + (size_t)p * ldl;
    T w0 = (lane < p) ? Lw[lane] : T(0);
    T w1 = (lane + 32 < p) ? Lw[lane + 32] : T(0);
    for (int c = p - 1; c >= 0; --c) {
        const T wc = __shfl_sync(INMF_FULL_MASK, (c < 32) ? w0 : w1, c & 31);
        const T zc = wc * sc.dinv[c];
        const T* Lc = sc.L + (size_t)c * ldl;
        if (lane < c) w0 -= Lc[lane] * zc;
        if (lane + 32 < c) w1 -= Lc[lane + 32] * zc;
        if (lane == (c & 31)) {
            if (c < 32) w0 = zc; else w1 = zc;
        }
    }
    // threshold as _utilities.py:290 and publish the compacted solution
    if (lane < p) sc.z[lane] = (inmf_abs(w0) < T(1e-16)) ? T(0) : w0;
    if (lane + 32 < p) sc.z[lane + 32] = (inmf_abs(w1) < T(1e-16)) ? T(0) : w1;
    __syncwarp();

    x0 = pas0 ? sc.z[pos0] : T(0);
    x1 = pas1 ? sc.z[pos1] : T(0);
    // y = G x - r on the non-passive rows (zero on the passive set), thresholded as :292
    T a0 = T(0), a1 = T(0);
    const T* G0 = Gs + (size_t)(has0 ? j0 : 0) * ldg;
    const T* G1 = Gs + (size_t)(has1 ? j1 : 0) * ldg;
    for (int c = 0; c < p; ++c) {
        const int ic = sc.idx[c];
        const T zc = sc.z[c];
        a0 += G0[ic] * zc;
        a1 += G1[ic] * zc;
    }
    a0 -= r0;
    a1 -= r1;
    y0 = (!has0 || pas0 || inmf_abs(a0) < T(1e-16)) ? T(0) : a0;
    y1 = (!has1 || pas1 || inmf_abs(a1) < T(1e-16)) ? T(0) : a1;
    __syncwarp();
    if (lane == 0) cnt.nfact++;
}

// Full BPP for one column.  r0/r1: this lane's entries of r (factor lane, lane+32; 0 beyond k).
// P: in = warm-start passive set (ignored unless warm), out = final passive set.
template <typename T>
__device__ __forceinline__ void bpp_column(const T* __restrict__ Gs, int ldg, BppWarpScratch<T>& sc, int k,
                                           T r0, T r1, uint64_t& P, bool warm, T& x0, T& x1, T& y0, T& y1,
                                           BppCounters& cnt) {
    const int lane = threadIdx.x & 31;
    const bool has0 = lane < k, has1 = lane + 32 < k;
    const uint64_t kmask = (k >= 64) ? ~0ull : ((1ull << k) - 1ull);
    if (has0) sc.r[lane] = r0;
    if (has1) sc.r[lane + 32] = r1;
    __syncwarp();
    x0 = T(0);
    x1 = T(0);
    y0 = has0 ? -r0 : T(0);
    y1 = has1 ? -r1 : T(0);
    P = warm ? (P & kmask) : 0ull;
    if (P != 0ull) bpp_solve_passive<T>(Gs, ldg, sc, k, P, r0, r1, x0, x1, y0, y1, cnt);
    int ninf = k + 1;
    int pbar = 3;
    const int max_rounds = 5 * k;  // _utilities.py:220
    int rounds = 0;
    while (true) {
        const bool bad0 = has0 && (mask_bit(P, lane) ? (x0 < T(0)) : (y0 < T(0)));
        const bool bad1 = has1 && (mask_bit(P, lane + 32) ? (x1 < T(0)) : (y1 < T(0)));
        const uint64_t I = (uint64_t)__ballot_sync(INMF_FULL_MASK, bad0) |
                           ((uint64_t)__ballot_sync(INMF_FULL_MASK, bad1) << 32);
        const int bad = __popcll(I);
        if (bad == 0) break;
        ++rounds;
        if (rounds > max_rounds) {
            cnt.ok = false;
            --rounds;
            break;
        }
        if (bad < ninf) {  // improving: full exchange, reset the patience counter (:264-270)
            ninf = bad;
            pbar = 3;
            P ^= I;
        } else if (pbar >= 1) {  // not improving: spend patience, still full exchange (:271-277)
            --pbar;
            P ^= I;
        } else {  // backup rule: flip only the largest infeasible index (:278-283)
            P ^= 1ull << (63 - __clzll(I));
            if (lane == 0) cnt.nbackup++;
        }
        if (P != 0ull) {
            bpp_solve_passive<T>(Gs, ldg, sc, k, P, r0, r1, x0, x1, y0, y1, cnt);
        } else {
            x0 = T(0);
            x1 = T(0);
            y0 = has0 ? -r0 : T(0);
            y1 = has1 ? -r1 : T(0);
        }
    }
    cnt.rounds = rounds;
}
