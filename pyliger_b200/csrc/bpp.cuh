// Warp-per-column block principal pivoting NNLS (device side).
//
// Algorithm: Kim & Park 2011 as used by pyliger (/root/reference/src/pyliger/factorization/
// _utilities.py:226-299).  Columns are independent (the reference's passive-set grouping,
// _utilities.py:341-357, is only a CPU cost optimisation), so one warp owns one column:
//   * the passive set is a 64-bit mask held uniformly by the warp (k <= 64),
//   * infeasible indices are found with __ballot_sync,
//   * the passive-set system G_PP z = r_P is solved by a left-looking Cholesky of the compacted
//     augmented matrix [G_PP ; r_P'] (the extra row is the forward substitution, for free):
//       - fast path (|P| <= 31): lane i keeps row i of L in REGISTERS (fully unrolled), the pivot
//         row is broadcast from a per-warp shared copy with 128-bit loads, back substitution is
//         right-looking with one shuffle per step;
//       - complement path (|P| > 31, k - |P| <= 31, G^-1 available): the same register routine on
//         (G^-1)_FF for the active set F, see bpp_solve_passive;
//       - general path (anything else, |P| up to 64): L lives in the per-warp shared scratch.
//   * y = G x - r is formed for the non-passive rows from the shared copy of G.
// The reference solves with LAPACK LU (numpy.linalg.solve); for symmetric positive definite G_PP the
// Cholesky solution is the same up to round-off.
//
// Zero thresholds: the reference zeroes |x|, |y| < 1e-16 (_utilities.py:290,292).  That is kept
// verbatim for T = double.  For T = float the same absolute constant is meaningless, so values
// within a few ulps of the terms they were summed from are treated as zero (see zero_x / zero_y);
// without it near-degenerate columns cycle until the 5k round limit.
#pragma once
#include "inmf_common.cuh"

#define BPP_FAST_MAX_P 31

template <typename T>
struct BppTraits;
template <>
struct BppTraits<float> {
    static constexpr int VEC = 4;
    static constexpr int LDL_FAST = 36;  // 32 + 4: 16-byte aligned rows, 4-way (not 32-way) column stores
    typedef float4 Vec;
    __device__ static __forceinline__ float rel_tol() { return 9.5367431640625e-7f; }  // 16 ulp
};
template <>
struct BppTraits<double> {
    static constexpr int VEC = 2;
    static constexpr int LDL_FAST = 34;
    typedef double2 Vec;
    __device__ static __forceinline__ double rel_tol() { return 0.0; }  // reference behaviour: absolute 1e-16 only
};

template <typename T>
struct BppWarpScratch {
    T* L;          // max((k+1) x ldl, 32 x LDL_FAST)
    T* r;          // k   right-hand side of this column
    T* z;          // k   compacted solution (general path)
    T* dinv;       // k   1 / L[c][c]         (general path)
    T* xf;         // k   unconstrained solution G^-1 r of this column (complement path)
    uint8_t* idx;  // 64  compacted position -> factor index
    int ldl;       // general-path leading dimension
};

template <typename T>
__host__ __device__ inline size_t bpp_warp_scratch_elems(int k) {
    const int ldl = k | 1;
    size_t l_elems = (size_t)(k + 1) * ldl;
    const size_t fast = (size_t)32 * BppTraits<T>::LDL_FAST;
    if (l_elems < fast) l_elems = fast;
    l_elems = (l_elems + 3) & ~(size_t)3;
    const size_t kp = (k + 3) & ~3;
    const size_t zk = kp < 36 ? 36 : kp;  // the fast path writes z for all 32 lanes (+ padded reads)
    size_t total = l_elems + 3 * kp + zk + 64 / sizeof(T);
    return (total + 3) & ~(size_t)3;
}

template <typename T>
__device__ __forceinline__ BppWarpScratch<T> bpp_carve(T* base, int k) {
    BppWarpScratch<T> s;
    s.ldl = k | 1;
    size_t l_elems = (size_t)(k + 1) * s.ldl;
    const size_t fast = (size_t)32 * BppTraits<T>::LDL_FAST;
    if (l_elems < fast) l_elems = fast;
    l_elems = (l_elems + 3) & ~(size_t)3;
    const int kp = (k + 3) & ~3;
    s.L = base;
    s.r = s.L + l_elems;
    s.z = s.r + kp;
    s.dinv = s.z + (kp < 36 ? 36 : kp);
    s.xf = s.dinv + kp;
    s.idx = reinterpret_cast<uint8_t*>(s.xf + kp);
    return s;
}

struct BppCounters {
    int rounds;
    int nfact;
    int nbackup;
    int nfail;
    bool ok;
    int h[6];  // experiment counters (BPP_HIST builds only)
};

__device__ __forceinline__ bool mask_bit(uint64_t m, int j) { return (m >> j) & 1ull; }

// x: solution component, scale: largest |z| of the column's solve
template <typename T>
__device__ __forceinline__ T zero_x(T x, T scale) {
    return (inmf_abs(x) < T(1e-16) || inmf_abs(x) <= BppTraits<T>::rel_tol() * scale) ? T(0) : x;
}
// y = a - r with a = (G x)_j
template <typename T>
__device__ __forceinline__ T zero_y(T a, T r) {
    const T y = a - r;
    return (inmf_abs(y) < T(1e-16) || inmf_abs(y) <= BppTraits<T>::rel_tol() * (inmf_abs(a) + inmf_abs(r))) ? T(0) : y;
}

template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T t = __shfl_xor_sync(INMF_FULL_MASK, v, o);
        v = t > v ? t : v;
    }
    return v;
}

#include "bpp_fast.cuh"
#include "bpp16.cuh"

// ------------------------------------------------------------------------------------------------
// General path: any |P| <= 64, L in the shared scratch.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __noinline__ void bpp_solve_passive_general(const T* __restrict__ Gs, int ldg, BppWarpScratch<T>& sc,
                                                       int k, uint64_t P, int p, T r0, T r1, T& x0, T& x1, T& y0,
                                                       T& y1, BppCounters& cnt) {
    const int lane = threadIdx.x & 31;
    const int j0 = lane, j1 = lane + 32;
    const bool has0 = j0 < k, has1 = j1 < k;
    const int ldl = sc.ldl;
    const bool pas0 = has0 && mask_bit(P, j0);
    const bool pas1 = has1 && mask_bit(P, j1);
    const int pos0 = __popcll(P & ((1ull << j0) - 1ull));
    const int pos1 = __popcll(P & ((1ull << j1) - 1ull));
    if (pas0) sc.idx[pos0] = (uint8_t)j0;
    if (pas1) sc.idx[pos1] = (uint8_t)j1;
    __syncwarp();

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
        } else {
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
    if (lane >= p) w0 = T(0);
    if (lane + 32 >= p) w1 = T(0);
    const T zscale = warp_max(inmf_abs(w0) > inmf_abs(w1) ? inmf_abs(w0) : inmf_abs(w1));
    if (lane < p) sc.z[lane] = zero_x<T>(w0, zscale);
    if (lane + 32 < p) sc.z[lane + 32] = zero_x<T>(w1, zscale);
    __syncwarp();

    x0 = pas0 ? sc.z[pos0] : T(0);
    x1 = pas1 ? sc.z[pos1] : T(0);
    T a0 = T(0), a1 = T(0);
    const T* G0 = Gs + (size_t)(has0 ? j0 : 0) * ldg;
    const T* G1 = Gs + (size_t)(has1 ? j1 : 0) * ldg;
    for (int c = 0; c < p; ++c) {
        const int ic = sc.idx[c];
        const T zc = sc.z[c];
        a0 += G0[ic] * zc;
        a1 += G1[ic] * zc;
    }
    y0 = (!has0 || pas0) ? T(0) : zero_y<T>(a0, r0);
    y1 = (!has1 || pas1) ? T(0) : zero_y<T>(a1, r1);
    __syncwarp();
    cnt.nfact++;
}

// Unconstrained solution x_full = G^-1 r from the shared copy of the inverse (once per column).
template <typename T>
__device__ __forceinline__ void bpp_unconstrained(const T* __restrict__ Gis, int ldg, BppWarpScratch<T>& sc, int k) {
    const int lane = threadIdx.x & 31;
    const bool has0 = lane < k, has1 = lane + 32 < k;
    const T* __restrict__ I0 = Gis + (has0 ? lane : 0) * ldg;
    const T* __restrict__ I1 = Gis + (has1 ? lane + 32 : 0) * ldg;
    const T* __restrict__ r = sc.r;
    T a0 = T(0), b0 = T(0), a1 = T(0), b1 = T(0);
    int c = 0;
    for (; c + 1 < k; c += 2) {
        const T rc = r[c], rd = r[c + 1];
        a0 = fma(I0[c], rc, a0);
        b0 = fma(I0[c + 1], rd, b0);
        if (k > 32) {
            a1 = fma(I1[c], rc, a1);
            b1 = fma(I1[c + 1], rd, b1);
        }
    }
    if (c < k) {
        a0 = fma(I0[c], r[c], a0);
        if (k > 32) a1 = fma(I1[c], r[c], a1);
    }
    if (has0) sc.xf[lane] = a0 + b0;  // kept in shared memory: no registers held across the rounds
    if (has1) sc.xf[lane + 32] = a1 + b1;
    __syncwarp();
}

// Passive-set solve for the partition P (0 < |P|; |P| < k whenever the inverse is available).
//
// |P| <= 31: Cholesky of G_PP in registers.
// |P| >  31 and |F| = k - |P| <= 31 (every case for k <= 62), inverse available: solve on the complement.
//   With x = G^-1 (r + y), y_P = 0, x_F = 0:   (G^-1)_FF y_F = -xfull_F,   x_P = xfull_P + (G^-1)_PF y_F.
//   That is the same routine on M = G^-1, S = F, q = xfull: it returns u = -y_F on F and (M u - q) = -x_P on P.
// otherwise: general path.
// `prefer`: take the complement whenever it is the smaller system (|F| < |P|) -- for well-conditioned G (the
// H-update's (W+V)(W+V)' + lambda VV') with mostly-passive solutions this replaces a |P| ~ 0.8k factorisation by
// a |F| ~ 0.2k one.
// COMP = false (k <= 31 without `prefer`, where |P| <= 31 always) compiles the complement path away.
template <typename T, bool COMP>
__device__ __forceinline__ void bpp_solve_passive(const T* __restrict__ Gs, const T* __restrict__ Gis, int ldg,
                                                  BppWarpScratch<T>& sc, int k, uint64_t kmask, uint64_t P,
                                                  bool prefer, T r0, T r1, T& x0, T& x1, T& y0, T& y1,
                                                  BppCounters& cnt) {
    const int p = __popcll(P);
    const bool comp = COMP && Gis != nullptr && k - p <= BPP_FAST_MAX_P &&
                      (p > BPP_FAST_MAX_P || (prefer && k - p < p));
    if (!COMP || p <= BPP_FAST_MAX_P || comp) {  // !COMP: k <= 31, so is p
        T u0, u1, v0, v1;
        if (comp) {
            const int lane = threadIdx.x & 31;
            r0 = (lane < k) ? sc.xf[lane] : T(0);
            r1 = (lane + 32 < k) ? sc.xf[lane + 32] : T(0);
        }
        bpp_solve_passive_fast<T>(comp ? Gis : Gs, ldg, sc, comp ? sc.xf : sc.r, k, comp ? (kmask & ~P) : P,
                                  comp ? k - p : p, r0, r1, u0, u1, v0, v1, cnt);
        x0 = comp ? T(0) - v0 : u0;
        x1 = comp ? T(0) - v1 : u1;
        y0 = comp ? T(0) - u0 : v0;
        y1 = comp ? T(0) - u1 : v1;
    } else {
        bpp_solve_passive_general<T>(Gs, ldg, sc, k, P, p, r0, r1, x0, x1, y0, y1, cnt);
    }
}

// Full BPP for one column.  r0/r1: this lane's entries of r (factor lane, lane+32; 0 beyond k).
// P: in = warm-start passive set (ignored unless warm), out = final passive set.
template <typename T, bool COMP>
__device__ __forceinline__ void bpp_column(const T* __restrict__ Gs, const T* __restrict__ Gis, int ldg,
                                           BppWarpScratch<T>& sc, int k,
                                           T r0, T r1, uint64_t& P, bool warm, bool prefer, T& x0, T& x1, T& y0,
                                           T& y1, BppCounters& cnt) {
    const int lane = threadIdx.x & 31;
    const bool has0 = lane < k, has1 = lane + 32 < k;
    const uint64_t kmask = (k >= 64) ? ~0ull : ((1ull << k) - 1ull);
    if (has0) sc.r[lane] = r0;
    if (has1) sc.r[lane + 32] = r1;
    __syncwarp();
    P = warm ? (P & kmask) : 0ull;
    int ninf = k + 1;
    int pbar = 3;
    const int max_rounds = 5 * k;  // _utilities.py:220
    int rounds = 0;
    bool have_xf = false;  // sc.xf = G^-1 r, formed the first time a partition with |P| > 31 comes up
    while (true) {
        // evaluate the current partition (single call site keeps the unrolled solver in one copy)
        if (Gis != nullptr && !have_xf &&
            (P == kmask || (COMP && (__popcll(P) > BPP_FAST_MAX_P || (prefer && 2 * __popcll(P) > k))))) {
            bpp_unconstrained<T>(Gis, ldg, sc, k);
            have_xf = true;
        }
        if (P == kmask && have_xf) {  // all indices passive: x = G^-1 r, y = 0
            const T xf0 = has0 ? sc.xf[lane] : T(0);
            const T xf1 = has1 ? sc.xf[lane + 32] : T(0);
            const T m = inmf_abs(xf0) > inmf_abs(xf1) ? inmf_abs(xf0) : inmf_abs(xf1);
            const T zscale = warp_max(m);
            x0 = zero_x<T>(xf0, zscale);
            x1 = zero_x<T>(xf1, zscale);
            y0 = T(0);
            y1 = T(0);
            cnt.nfact++;
        } else if (P != 0ull) {
            bpp_solve_passive<T, COMP>(Gs, Gis, ldg, sc, k, kmask, P, prefer, r0, r1, x0, x1, y0, y1, cnt);
        } else {
            x0 = T(0);
            x1 = T(0);
            y0 = has0 ? -r0 : T(0);
            y1 = has1 ? -r1 : T(0);
        }
        const bool bad0 = has0 && (mask_bit(P, lane) ? (x0 < T(0)) : (y0 < T(0)));
        const bool bad1 = has1 && (mask_bit(P, lane + 32) ? (x1 < T(0)) : (y1 < T(0)));
        const uint64_t I = (uint64_t)__ballot_sync(INMF_FULL_MASK, bad0) |
                           ((uint64_t)__ballot_sync(INMF_FULL_MASK, bad1) << 32);
        const int bad = __popcll(I);
        if (bad == 0) break;
#ifdef BPP_HIST  // experiment: size of the first exchange set
        if (rounds == 0) {
            const int rem = __popcll(I & P);
            cnt.h[0] = (bad == 1 && rem == 1);   // one removal
            cnt.h[1] = (bad == 1 && rem == 0);   // one addition
            cnt.h[2] = (bad == 2);
            cnt.h[3] = (bad == 3 || bad == 4);
            cnt.h[4] = (bad >= 5);
        }
        if (rounds == 1) cnt.h[5] = 1;           // needed a third solve
#endif
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
            cnt.nbackup++;
        }
    }
    cnt.rounds = rounds;
}
