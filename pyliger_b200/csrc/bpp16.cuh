// Half-warp block principal pivoting: two columns per warp, 16 lanes each (included by bpp.cuh).
//
// Why: in the H-update the partitions are solved on whichever of the passive set P / its complement F is smaller
// (bpp.cuh: complement solves through G^-1), so the systems have min(|P|, |F|) <= k/2 unknowns -- measured mean 7 of
// k = 30 at C2, 9 of 40 at C4 -- and the gene-side (V, W) solves have sparse solutions (|P| small).  A 32-lane warp per
// column then idles 3/4 of its lanes in the factorisation and spends the same ~1 000 instructions per column whatever
// the size.  Here a TILE of 16 lanes owns a column: lane t holds factors t, t + 16, ... (NF per lane), row t of the
// augmented Cholesky factor [M_SS ; q_S'] lives in 16 registers, and every shuffle / ballot / barrier is confined to
// the tile (width-16 shuffles).  The two tiles of a warp run in LOCKSTEP: the control flow is warp-uniform (a loop
// runs while either tile needs it, a tile that does not is predicated off and works on dummy data), so every
// collective uses the full-warp mask and compiles to a plain SHFL / VOTE -- with per-tile member masks the compiler
// wraps each of them in a WARPSYNC / ENDCOLLECTIVE sequence, which cost a fifth of the instructions of the first
// version of this file and, with the solver inlined at two call sites, pushed the kernel to 90 KB of code: ncu showed
// "no instruction" (instruction-cache misses) as its largest stall.  Same algorithm, rules, thresholds and counters
// as bpp_column / bpp_solve_passive_fast.
//
// A partition that needs more than 15 unknowns on both sides (or on the passive side when no valid inverse exists)
// cannot be factorised in a tile: the column is DEFERRED -- nothing is written for it, its index goes to a per-segment
// work list -- and the full-warp kernel (bpp_kernel, list mode) solves the listed columns from their original start
// set afterwards, so the exchange-rule sequence (and the counters) of every column are exactly the reference's.
#pragma once

#define BPP16_MAX_P 15

template <typename T>
struct Bpp16Traits;
template <>
struct Bpp16Traits<float> {
    static constexpr int LDL = 20;  // 16 + 4: 16-byte aligned rows
};
template <>
struct Bpp16Traits<double> {
    static constexpr int LDL = 18;
};

template <typename T>
struct Bpp16Scratch {
    T* L;          // 16 x LDL   shared copy of the Cholesky rows (pivot-row broadcasts)
    T* r;          // 16 * NF    right-hand side of this column
    T* xf;         // 16 * NF    unconstrained solution G^-1 r
    T* z;          // 20         compacted solution (+ zero padding read by the residual loop)
    uint8_t* idx;  // 32         compacted position -> factor index
};

template <typename T>
__host__ __device__ inline size_t bpp16_tile_scratch_elems(int nf) {
    size_t total = (size_t)16 * Bpp16Traits<T>::LDL + 2 * (size_t)16 * nf + 20 + 32 / sizeof(T);
    return (total + 3) & ~(size_t)3;
}

template <typename T>
__device__ __forceinline__ Bpp16Scratch<T> bpp16_carve(T* base, int nf) {
    Bpp16Scratch<T> s;
    s.L = base;
    s.r = s.L + 16 * Bpp16Traits<T>::LDL;
    s.xf = s.r + 16 * nf;
    s.z = s.xf + 16 * nf;
    s.idx = reinterpret_cast<uint8_t*>(s.z + 20);
    return s;
}

// Passive-set masks: 32 bits are enough (and a third of the integer instructions) for k <= 32.
template <int NF>
struct Bpp16Mask {
    typedef uint64_t type;
};
template <>
struct Bpp16Mask<1> {
    typedef uint32_t type;
};
template <>
struct Bpp16Mask<2> {
    typedef uint32_t type;
};
__device__ __forceinline__ int m_popc(uint32_t m) { return __popc(m); }
__device__ __forceinline__ int m_popc(uint64_t m) { return __popcll(m); }
__device__ __forceinline__ bool m_bit(uint32_t m, int j) { return (m >> j) & 1u; }
__device__ __forceinline__ bool m_bit(uint64_t m, int j) { return (m >> j) & 1ull; }
__device__ __forceinline__ uint32_t m_below(uint32_t m, int j) { return m & ((1u << j) - 1u); }       // j < 32
__device__ __forceinline__ uint64_t m_below(uint64_t m, int j) { return m & ((1ull << j) - 1ull); }   // j < 64
__device__ __forceinline__ uint32_t m_top(uint32_t m) { return 1u << (31 - __clz(m)); }
__device__ __forceinline__ uint64_t m_top(uint64_t m) { return 1ull << (63 - __clzll(m)); }

template <typename T>
__device__ __forceinline__ T tile16_max(T v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        const T t = __shfl_xor_sync(INMF_FULL_MASK, v, o, 16);
        v = t > v ? t : v;
    }
    return v;
}

// Solve M_SS z = q_S (|S| = p <= 15) and form the residual M z - q off S, for both tiles of the warp at once.
//   u[f] = z on the factors of S held by this lane (0 elsewhere), v[f] = (M u - q) off S (0 on S).
// p = 0 for a tile that has nothing to solve (it runs the same instructions on dummy data); pmax = the larger p of
// the two tiles (warp-uniform trip counts).
template <typename T, int NF>
__device__ __forceinline__ void bpp16_solve(const T* __restrict__ M, int ldg, Bpp16Scratch<T>& sc,
                                            const T* __restrict__ rhs, int k, typename Bpp16Mask<NF>::type S, int p,
                                            int pmax, const T (&q)[NF], T (&u)[NF], T (&v)[NF], int t,
                                            BppCounters& cnt) {
    typedef BppTraits<T> TR;
    typedef typename TR::Vec Vec;
    constexpr int VEC = TR::VEC;
    constexpr int LDL = Bpp16Traits<T>::LDL;
    bool has[NF], in[NF];
    int pos[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const int j = t + 16 * f;
        has[f] = j < k;
        in[f] = has[f] && m_bit(S, j);
        pos[f] = m_popc(m_below(S, j));
        if (in[f]) sc.idx[pos[f]] = (uint8_t)j;
    }
    __syncwarp();
    const bool is_g = t < p;
    const int myidx = is_g ? (int)sc.idx[t] : 0;
    const T* __restrict__ gbase = is_g ? (M + myidx * ldg) : rhs;
    T* __restrict__ Ls = sc.L;
    T* __restrict__ Lmine = Ls + t * LDL;
    T L[16];
    T mydinv = T(0);
    T dmin = T(1);
    T s_part = gbase[__shfl_sync(INMF_FULL_MASK, myidx, 0, 16)];
#pragma unroll
    for (int c = 0; c < BPP16_MAX_P; ++c) {
        if (c >= pmax) break;
        __syncwarp();  // step c-1's column store is visible to the prefetch below
        // prefetch for step c+1: everything that does not depend on this step's pivot
        T s_next = gbase[__shfl_sync(INMF_FULL_MASK, myidx, (c + 1) & 15, 16)];
        {
            T s2 = T(0);
            const T* __restrict__ Ln = Ls + (c + 1) * LDL;
#pragma unroll
            for (int jb = 0; jb < c; jb += VEC) {
                const Vec vv = *reinterpret_cast<const Vec*>(Ln + jb);
                const T* pv = reinterpret_cast<const T*>(&vv);
#pragma unroll
                for (int w = 0; w < VEC; ++w) {
                    if (jb + w < c) {
                        if (w & 1) s2 = fma(-L[jb + w], pv[w], s2);
                        else s_next = fma(-L[jb + w], pv[w], s_next);
                    }
                }
            }
            s_next += s2;
        }
        T s = s_part;
        if (c > 0) {
            const T last = __shfl_sync(INMF_FULL_MASK, L[c > 0 ? c - 1 : 0], c, 16);
            s = fma(-L[c > 0 ? c - 1 : 0], last, s);
        }
        const T d = __shfl_sync(INMF_FULL_MASK, s, c, 16);
        dmin = (c < p && d < dmin) ? d : dmin;
        const T di = (d > T(0)) ? inmf_rsqrt<T>(d) : T(0);
        const T l = (t > c) ? s * di : T(0);
        L[c] = l;
        Lmine[c] = l;
        mydinv = (t == c) ? di : mydinv;
        s_part = s_next;
    }
    __syncwarp();
    T w = is_g ? Ls[p * LDL + t] : T(0);
    for (int c = pmax - 1; c >= 0; --c) {
        const T zc = __shfl_sync(INMF_FULL_MASK, w * mydinv, c, 16);
        if (c < p) w = fma(-Ls[c * LDL + t], zc, w);
    }
    T z = is_g ? w * mydinv : T(0);
    const T zscale = tile16_max(inmf_abs(z));
    z = zero_x<T>(z, zscale);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const T tz = __shfl_sync(INMF_FULL_MASK, z, pos[f] & 15, 16);
        u[f] = in[f] ? tz : T(0);
    }
    sc.z[t] = z;  // lanes >= p write zeros: the padded terms below add nothing
    __syncwarp();
    T a[NF];
    const T* __restrict__ Mrow[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        a[f] = T(0);
        Mrow[f] = M + (has[f] ? t + 16 * f : 0) * ldg;
    }
    const uint32_t* __restrict__ idx4 = reinterpret_cast<const uint32_t*>(sc.idx);
    for (int cb = 0; cb < pmax; cb += 4) {
        const uint32_t ii = idx4[cb >> 2];
        const T z0 = sc.z[cb], z1 = sc.z[cb + 1], z2 = sc.z[cb + 2], z3 = sc.z[cb + 3];
        const int i0 = ii & 0xff, i1 = (ii >> 8) & 0xff, i2 = (ii >> 16) & 0xff, i3 = ii >> 24;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            if (f * 16 < k) {
                a[f] = fma(Mrow[f][i0], z0, a[f]);
                a[f] = fma(Mrow[f][i1], z1, a[f]);
                a[f] = fma(Mrow[f][i2], z2, a[f]);
                a[f] = fma(Mrow[f][i3], z3, a[f]);
            }
        }
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) v[f] = (!has[f] || in[f]) ? T(0) : zero_y<T>(a[f], q[f]);
    __syncwarp();
    if (p > 0) {
        cnt.nfail += (dmin > T(0)) ? 0 : 1;
        cnt.nfact++;
    }
}

// xf = G^-1 r for this tile's column (rows t + 16 f of the shared inverse), kept in shared memory.
template <typename T, int NF>
__device__ __forceinline__ void bpp16_unconstrained(const T* __restrict__ Gis, int ldg, Bpp16Scratch<T>& sc, int k,
                                                    int t) {
    T a[NF], b[NF];
    const T* __restrict__ row[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        a[f] = b[f] = T(0);
        row[f] = Gis + ((t + 16 * f < k) ? t + 16 * f : 0) * ldg;
    }
    const T* __restrict__ r = sc.r;
    int c = 0;
    for (; c + 1 < k; c += 2) {
        const T rc = r[c], rd = r[c + 1];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            if (f * 16 < k) {
                a[f] = fma(row[f][c], rc, a[f]);
                b[f] = fma(row[f][c + 1], rd, b[f]);
            }
        }
    }
    if (c < k) {
        const T rc = r[c];
#pragma unroll
        for (int f = 0; f < NF; ++f)
            if (f * 16 < k) a[f] = fma(row[f][c], rc, a[f]);
    }
#pragma unroll
    for (int f = 0; f < NF; ++f)
        if (t + 16 * f < k) sc.xf[t + 16 * f] = a[f] + b[f];
    __syncwarp();
}

// Full BPP for the two columns of a warp (one per tile) in lockstep.  `valid`: this tile has a column.
// Returns 1 = solved, 2 = has to be deferred to the full-warp solver (x, y, P are then meaningless and nothing may be
// written for the column), 0 = no column.
template <typename T, int NF>
__device__ __forceinline__ int bpp16_column(const T* __restrict__ Gs, const T* __restrict__ Gis, int ldg,
                                            Bpp16Scratch<T>& sc, int k, bool valid, const T (&r)[NF], uint64_t& P64,
                                            bool warm, bool prefer, T (&x)[NF], T (&y)[NF], BppCounters& cnt, int t,
                                            int half) {
    typedef typename Bpp16Mask<NF>::type Mask;
    const Mask kmask = (k >= (int)(8 * sizeof(Mask))) ? ~Mask(0) : (Mask)((Mask(1) << k) - Mask(1));
    bool has[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        has[f] = t + 16 * f < k;
        if (has[f]) sc.r[t + 16 * f] = r[f];
        x[f] = y[f] = T(0);
    }
    __syncwarp();
    Mask P = warm ? ((Mask)P64 & kmask) : Mask(0);
    int ninf = k + 1;
    int pbar = 3;
    const int max_rounds = 5 * k;  // _utilities.py:220
    int rounds = 0;
    bool have_xf = false;
    int status = valid ? 0 : 3;  // 0 running, 1 solved, 2 deferred, 3 no column
    while (true) {
        bool run = status == 0;
        const int p = m_popc(P);
        const int nf = k - p;
        const bool comp = Gis != nullptr && p > 0 && nf <= BPP16_MAX_P && (p > BPP16_MAX_P || (prefer && nf < p));
        const bool all_passive = P == kmask && Gis != nullptr;
        if (run && p > BPP16_MAX_P && !comp && !all_passive) {  // too large for a tile
            status = 2;
            run = false;
        }
        if (!__any_sync(INMF_FULL_MASK, run)) break;
        if (__any_sync(INMF_FULL_MASK, run && !have_xf && (all_passive || comp))) {
            bpp16_unconstrained<T, NF>(Gis, ldg, sc, k, t);  // (for both tiles: harmless where it is not needed)
            have_xf = true;
        }
        const bool solve = run && !all_passive && P != Mask(0);
        if (__any_sync(INMF_FULL_MASK, solve)) {
            // single call site: the unrolled solver exists once in the kernel (instruction cache)
            T u[NF], v[NF], q[NF];
#pragma unroll
            for (int f = 0; f < NF; ++f) q[f] = comp ? (has[f] ? sc.xf[t + 16 * f] : T(0)) : r[f];
            const int pe = solve ? (comp ? nf : p) : 0;
            const int po = __shfl_xor_sync(INMF_FULL_MASK, pe, 16);
            bpp16_solve<T, NF>(comp ? Gis : Gs, ldg, sc, comp ? sc.xf : sc.r, k,
                               solve ? (comp ? (Mask)(kmask & ~P) : P) : Mask(0), pe, pe > po ? pe : po, q, u, v, t, cnt);
            if (solve) {
#pragma unroll
                for (int f = 0; f < NF; ++f) {  // complement: u = -y_F, v = -x_P
                    x[f] = comp ? T(0) - v[f] : u[f];
                    y[f] = comp ? T(0) - u[f] : v[f];
                }
            }
        }
        if (__any_sync(INMF_FULL_MASK, run && all_passive)) {  // all indices passive: x = G^-1 r, y = 0
            T xs[NF];
            T m = T(0);
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                xs[f] = has[f] ? sc.xf[t + 16 * f] : T(0);
                m = inmf_abs(xs[f]) > m ? inmf_abs(xs[f]) : m;
            }
            const T zscale = tile16_max(m);
            if (run && all_passive) {
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    x[f] = zero_x<T>(xs[f], zscale);
                    y[f] = T(0);
                }
                cnt.nfact++;
            }
        }
        if (run && P == Mask(0)) {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                x[f] = T(0);
                y[f] = has[f] ? -r[f] : T(0);
            }
        }
        Mask I = Mask(0);
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const bool bad = run && has[f] && (m_bit(P, t + 16 * f) ? (x[f] < T(0)) : (y[f] < T(0)));
            const unsigned b = (__ballot_sync(INMF_FULL_MASK, bad) >> (16 * half)) & 0xffffu;
            I |= (Mask)b << (16 * f);
        }
        if (run) {
            const int bad = m_popc(I);
            if (bad == 0) {
                status = 1;
            } else if (rounds + 1 > max_rounds) {
                cnt.ok = false;
                status = 1;
            } else {
                ++rounds;
                if (bad < ninf) {  // improving: full exchange, reset the patience counter (:264-270)
                    ninf = bad;
                    pbar = 3;
                    P ^= I;
                } else if (pbar >= 1) {  // not improving: spend patience, still full exchange (:271-277)
                    --pbar;
                    P ^= I;
                } else {  // backup rule: flip only the largest infeasible index (:278-283)
                    P ^= m_top(I);
                    cnt.nbackup++;
                }
            }
        }
    }
    cnt.rounds = rounds;
    P64 = (uint64_t)P;
    return status == 3 ? 0 : status;
}
