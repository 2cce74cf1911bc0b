// BPP fast path: passive-set solve for |P| = p <= 31 with the Cholesky rows in registers.
// Included by bpp.cuh (needs BppTraits, BppWarpScratch, BppCounters, zero_x, zero_y, warp_max).
//
// Lane i owns compacted row i of the augmented matrix [G_PP ; r_P']: lanes 0..p-1 hold the rows of
// G_PP, lane p holds r_P (its row of L is the forward-substituted right-hand side w = L^-1 r_P).
// Per step c (fully unrolled, static register indices):
//   s_i  = A[i][c] - sum_{j<c} L[i][j] L[c][j]      own row from registers, pivot row c from the shared
//                                                    copy with 128-bit broadcast loads; the j = c-1 term
//                                                    (produced by the previous step) comes by shuffle, so
//                                                    the rest of the dot product is prefetched one step
//                                                    ahead and overlaps the pivot chain
//   d    = s_c (shuffle),  L[i][c] = s_i / sqrt(d)  for i > c, 0 otherwise (keeps the shared copy
//                                                    strictly lower triangular: no masks later)
// Back substitution is right-looking: one shuffle + one shared load + one FMA per step.
// Lanes beyond p run the same instructions on finite dummy data; nothing they produce is read.
//
// The routine is written for a generic SPD system M_SS z = q_S with the residual a - q formed off S.  The
// caller uses it two ways (bpp_solve_passive): directly (M = G, S = P, q = r), and -- for |P| > 31 --
// on the COMPLEMENT F = ~P with M = G^-1, q = G^-1 r, which yields -y_F and -x_P (see there).
#pragma once

template <typename T>
__device__ __forceinline__ void bpp_solve_passive_fast(const T* __restrict__ Gs, int ldg, BppWarpScratch<T>& sc,
                                                       const T* __restrict__ rhs, int k, uint64_t P, int p, T r0,
                                                       T r1, T& x0, T& x1, T& y0, T& y1, BppCounters& cnt) {
    typedef BppTraits<T> TR;
    typedef typename TR::Vec Vec;
    constexpr int VEC = TR::VEC;
    constexpr int LDL = TR::LDL_FAST;
    const int lane = threadIdx.x & 31;
    const int j0 = lane, j1 = lane + 32;
    const bool has0 = j0 < k, has1 = j1 < k;
    const bool pas0 = has0 && mask_bit(P, j0);
    const bool pas1 = has1 && mask_bit(P, j1);
    const int pos0 = __popcll(P & ((1ull << j0) - 1ull));
    const int pos1 = __popcll(P & ((1ull << j1) - 1ull));
    if (pas0) sc.idx[pos0] = (uint8_t)j0;
    if (pas1) sc.idx[pos1] = (uint8_t)j1;
    __syncwarp();
    const bool is_g = lane < p;
    const int myidx = is_g ? (int)sc.idx[lane] : 0;
    // one gather base per lane: G row of my factor, or the right-hand side for the augmented row
    const T* __restrict__ gbase = is_g ? (Gs + myidx * ldg) : rhs;
    T* __restrict__ Ls = sc.L;
    T* __restrict__ Lmine = Ls + lane * LDL;
    T L[32];
    T mydinv = T(0);
    T dmin = T(1);
    T s_part = gbase[__shfl_sync(INMF_FULL_MASK, myidx, 0)];
#pragma unroll
    for (int c = 0; c < BPP_FAST_MAX_P; ++c) {
        if (c >= p) break;
        __syncwarp();  // step c-1's column store is visible to the prefetch below
        // ---- prefetch for step c+1: everything that does not depend on this step's pivot ----
        T s_next = gbase[__shfl_sync(INMF_FULL_MASK, myidx, (c + 1) & 31)];
        {
            T s2 = T(0);
            const T* __restrict__ Ln = Ls + (c + 1) * LDL;
#pragma unroll
            for (int jb = 0; jb < c; jb += VEC) {
                const Vec v = *reinterpret_cast<const Vec*>(Ln + jb);
                const T* pv = reinterpret_cast<const T*>(&v);
#pragma unroll
                for (int u = 0; u < VEC; ++u) {
                    if (jb + u < c) {
                        if (u & 1) s2 = fma(-L[jb + u], pv[u], s2);
                        else s_next = fma(-L[jb + u], pv[u], s_next);
                    }
                }
            }
            s_next += s2;
        }
        // ---- step c ----
        T s = s_part;
        if (c > 0) {
            const T last = __shfl_sync(INMF_FULL_MASK, L[c > 0 ? c - 1 : 0], c);
            s = fma(-L[c > 0 ? c - 1 : 0], last, s);
        }
        const T d = __shfl_sync(INMF_FULL_MASK, s, c);
        dmin = d < dmin ? d : dmin;
        const T di = (d > T(0)) ? inmf_rsqrt<T>(d) : T(0);
        const T l = (lane > c) ? s * di : T(0);
        L[c] = l;
        Lmine[c] = l;
        mydinv = (lane == c) ? di : mydinv;
        s_part = s_next;
    }
    __syncwarp();
    // w = forward-substituted right-hand side = row p of L
    T w = is_g ? Ls[p * LDL + lane] : T(0);
    // right-looking back substitution: z_c = w_c / L[c][c], then w_j -= L[c][j] z_c for j < c
    // (entries of the shared copy with column >= row are zero, so no lane masks are needed)
    for (int c = p - 1; c >= 0; --c) {
        const T zc = __shfl_sync(INMF_FULL_MASK, w * mydinv, c);
        w = fma(-Ls[c * LDL + lane], zc, w);
    }
    T z = is_g ? w * mydinv : T(0);
    const T zscale = warp_max(inmf_abs(z));
    z = zero_x<T>(z, zscale);
    const T t0 = __shfl_sync(INMF_FULL_MASK, z, pos0 & 31);
    const T t1 = __shfl_sync(INMF_FULL_MASK, z, pos1 & 31);
    x0 = pas0 ? t0 : T(0);
    x1 = pas1 ? t1 : T(0);
    // y = G x - r on the non-passive rows: z and idx are broadcast from shared memory, 4 terms per trip
    sc.z[lane] = z;  // lanes >= p write zeros: the padded terms below add nothing
    __syncwarp();
    T a0 = T(0), a1 = T(0);
    const T* __restrict__ G0 = Gs + (has0 ? j0 : 0) * ldg;
    const T* __restrict__ G1 = Gs + (has1 ? j1 : 0) * ldg;
    const uint32_t* __restrict__ idx4 = reinterpret_cast<const uint32_t*>(sc.idx);
    for (int cb = 0; cb < p; cb += 4) {
        const uint32_t ii = idx4[cb >> 2];
        const T z0 = sc.z[cb], z1 = sc.z[cb + 1], z2 = sc.z[cb + 2], z3 = sc.z[cb + 3];
        const int i0 = ii & 0xff, i1 = (ii >> 8) & 0xff, i2 = (ii >> 16) & 0xff, i3 = ii >> 24;
        a0 = fma(G0[i0], z0, a0);
        a0 = fma(G0[i1], z1, a0);
        a0 = fma(G0[i2], z2, a0);
        a0 = fma(G0[i3], z3, a0);
        if (k > 32) {
            a1 = fma(G1[i0], z0, a1);
            a1 = fma(G1[i1], z1, a1);
            a1 = fma(G1[i2], z2, a1);
            a1 = fma(G1[i3], z3, a1);
        }
    }
    y0 = (!has0 || pas0) ? T(0) : zero_y<T>(a0, r0);
    y1 = (!has1 || pas1) ? T(0) : zero_y<T>(a1, r1);
    __syncwarp();
    cnt.nfail += (dmin > T(0)) ? 0 : 1;
    cnt.nfact++;
}
