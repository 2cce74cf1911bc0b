// libinmf_b200.so -- kernels + C ABI for pyliger's iNMF hot path on B200 (sm_100a).
// ABI and reference citations: include/inmf_b200.h.  Design notes: DESIGN.md.
#include "inmf_common.cuh"
#include "bpp.cuh"
#include "blocked.cuh"
#include "sell.cuh"
#include "tc_common.cuh"
#include "tc_pass.cuh"
#include "quantile.cuh"

thread_local char g_inmf_err[512] = "";

long long g_inmf_launches = 0;

// Deterministic launch configurations (see include/inmf_b200.h, inmf_set_deterministic).
static int g_inmf_deterministic = 0;
extern "C" int32_t inmf_set_deterministic(int32_t on) {
    const int old = __atomic_exchange_n(&g_inmf_deterministic, on ? 1 : 0, __ATOMIC_RELAXED);
    return old;
}
static inline bool inmf_deterministic() { return __atomic_load_n(&g_inmf_deterministic, __ATOMIC_RELAXED) != 0; }

extern "C" int32_t inmf_version(void) { return 200; }
extern "C" int64_t inmf_launch_count(void) { return (int64_t)__atomic_load_n(&g_inmf_launches, __ATOMIC_RELAXED); }
// Kernels of this library that ran through a CUDA-graph replay (the host-side counter above only sees the launches made
// while the graph was captured): the caller reports how many kernel nodes each replay contained.
extern "C" void inmf_note_graph_launches(int64_t n) { __atomic_add_fetch(&g_inmf_launches, (long long)n, __ATOMIC_RELAXED); }
extern "C" const char* inmf_last_error(void) { return g_inmf_err; }

// =================================================================================================
// K3: batched BPP
#ifndef BPP_THREADS
#define BPP_THREADS 128
#define BPP_MIN_BLOCKS 6
#endif
// =================================================================================================
// In-place Gauss-Jordan inverse of every segment's SPD normal matrix (double, one CTA per segment).
// Used for the "all indices passive" solve x = G^-1 r, which is what the first round of a cold start
// always asks for when the right-hand sides are non-negative.  A pivot that is non-positive or smaller than
// rcond_min times the largest diagonal entry (the Gauss-Jordan pivots of an SPD matrix are its successive Schur
// complements: a cheap conditioning proxy) marks the inverse invalid (NaN in element 0) and the solver falls
// back to the Cholesky paths, whose accuracy depends on cond(G_PP) only.
// Every thread keeps its (up to 4) elements in registers; an elimination step only needs row p and column p of the
// current matrix, which their owners publish (double-buffered) while they compute step p - 1: one barrier per step.
__global__ void __launch_bounds__(1024) ginv_kernel(const double* __restrict__ G, double* __restrict__ Ginv, int k,
                                                    double rcond_min) {
    __shared__ double rowbuf[2][INMF_MAX_K], colbuf[2][INMF_MAX_K], diag[INMF_MAX_K];
    const int seg = blockIdx.x;
    const int t = threadIdx.x, nt = blockDim.x;
    const int kk = k * k;
    int ei[4], ej[4];
    double a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int e = t + u * nt;
        ei[u] = e < kk ? e / k : -1;
        ej[u] = e < kk ? e - (e / k) * k : 0;
        a[u] = e < kk ? G[(size_t)seg * kk + e] : 0.0;
        if (ei[u] >= 0) {
            if (ei[u] == ej[u]) diag[ei[u]] = a[u];
            if (ei[u] == 0) rowbuf[0][ej[u]] = a[u];
            if (ej[u] == 0) colbuf[0][ei[u]] = a[u];
        }
    }
    __syncthreads();
    double dmax = 0.0;
    for (int p = 0; p < k; ++p) dmax = fmax(dmax, diag[p]);
    const double floor_piv = rcond_min * dmax;
    bool ok = true;
    for (int p = 0; p < k; ++p) {
        const int b = p & 1;
        const double piv = rowbuf[b][p];  // uniform: all threads see the same value, so the break is too
        if (!(piv > floor_piv)) {
            ok = false;
            break;
        }
        const double rp = 1.0 / piv;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = ei[u], j = ej[u];
            if (i < 0) continue;
            const double prow = ((j == p) ? 1.0 : rowbuf[b][j]) * rp;
            const double nv = (i == p) ? prow : ((j == p) ? 0.0 : a[u]) - colbuf[b][i] * prow;
            a[u] = nv;
            if (i == p + 1) rowbuf[b ^ 1][j] = nv;
            if (j == p + 1) colbuf[b ^ 1][i] = nv;
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (ei[u] >= 0)
            Ginv[(size_t)seg * kk + t + u * nt] = ok ? a[u] : __longlong_as_double(0x7ff8000000000000ll);
}

// COMP kernels (k > 31) are limited to <= 5 CTAs per SM by their shared-memory scratch anyway: give them the registers.
#ifndef BPP_COMP_MIN_BLOCKS
#define BPP_COMP_MIN_BLOCKS (BPP_MIN_BLOCKS - 1)
#endif
template <typename T, bool COMP>
__global__ void __launch_bounds__(BPP_THREADS, COMP ? BPP_COMP_MIN_BLOCKS : BPP_MIN_BLOCKS) bpp_kernel(const double* __restrict__ G, const double* __restrict__ Ginv, const T* R, int64_t ldr, T* X, int64_t ldx,
                           T* __restrict__ Y, int64_t ldy, uint64_t* __restrict__ mask, int warm, const int64_t* __restrict__ seg_off, int k,
                           long long* __restrict__ stats, const int* __restrict__ list_count,
                           const int32_t* __restrict__ list, int64_t list_cap) {
    // list mode (list != nullptr): only the columns the half-warp solver deferred, list[seg * list_cap + i]
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int seg = blockIdx.y;
    const int64_t c_begin = seg_off[seg];
    const int64_t ncols = list ? (int64_t)list_count[seg] : seg_off[seg + 1] - c_begin;
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if ((int64_t)blockIdx.x * nwarps >= ncols) return;

    // the first column's right-hand side and start set are requested before the staging of G, and every next column's
    // before the current one is solved: the global-load latency otherwise sits in front of every column
    const int64_t stride = (int64_t)gridDim.x * nwarps;
    int64_t c = (int64_t)blockIdx.x * nwarps + warp;
    int64_t col_n = 0;
    T r0_n = T(0), r1_n = T(0);
    uint64_t P_n = 0ull;
    auto fetch = [&](int64_t cc) {
        col_n = c_begin + (list ? (int64_t)list[(size_t)seg * list_cap + cc] : cc);
        const T* r = R + col_n * ldr;
        r0_n = (lane < k) ? r[lane] : T(0);
        r1_n = (lane + 32 < k) ? r[lane + 32] : T(0);
        P_n = ((warm & 1) && mask) ? mask[col_n] : 0ull;
    };
    if (c < ncols) fetch(c);

    const int ldg = k | 1;
    T* Gs = smem;
    const double* Gseg = G + (size_t)seg * k * k;
#pragma unroll 4
    for (int row = warp; row < k; row += nwarps) {
        if (lane < k) Gs[row * ldg + lane] = (T)Gseg[row * k + lane];
        if (lane + 32 < k) Gs[row * ldg + lane + 32] = (T)Gseg[row * k + lane + 32];
    }
    size_t goff = ((size_t)k * ldg + 3) & ~(size_t)3;
    const T* Gis = nullptr;  // shared copy of G^-1 (full-passive-set solve), if provided and valid
    if (Ginv != nullptr) {
        const double* Iseg = Ginv + (size_t)seg * k * k;
        T* Gi = smem + goff;
#pragma unroll 4
        for (int row = warp; row < k; row += nwarps) {
            if (lane < k) Gi[row * ldg + lane] = (T)Iseg[row * k + lane];
            if (lane + 32 < k) Gi[row * ldg + lane + 32] = (T)Iseg[row * k + lane + 32];
        }
        if (Iseg[0] == Iseg[0]) Gis = Gi;
        goff *= 2;
    }
    __syncthreads();
    BppWarpScratch<T> sc = bpp_carve<T>(smem + goff + (size_t)warp * bpp_warp_scratch_elems<T>(k), k);
    if (lane < 16) reinterpret_cast<uint32_t*>(sc.idx)[lane] = 0u;  // padded index reads stay in range
    __syncwarp();

    long long n_notconv = 0, n_fact = 0, n_backup = 0, n_fail = 0, n_rounds = 0, max_rounds = 0, n_cols = 0;
    for (; c < ncols; c += stride) {
        const int64_t col = col_n;
        const T r0 = r0_n, r1 = r1_n;
        uint64_t P = P_n;
        if (c + stride < ncols) fetch(c + stride);
        T x0, x1, y0, y1;
        BppCounters cnt = {0, 0, 0, 0, true, {0, 0, 0, 0, 0, 0}};
        bpp_column<T, COMP>(Gs, Gis, ldg, sc, k, r0, r1, P, (warm & 1) != 0, (warm & 2) != 0, x0, x1, y0, y1, cnt);
        T* x = X + col * ldx;
        if (lane < k) x[lane] = x0;
        if (lane + 32 < k) x[lane + 32] = x1;
        if (Y) {  // warm bit 2: Y is a scratch buffer to clear (the H-update's spare right-hand-side buffer)
            T* y = Y + col * ldy;
            const bool clr = (warm & 4) != 0;
            if (lane < k) y[lane] = clr ? T(0) : y0;
            if (lane + 32 < k) y[lane + 32] = clr ? T(0) : y1;
        }
        if (mask && lane == 0) mask[col] = P;
        if (lane == 0) {
            n_cols++;
#ifndef BPP_HIST
            n_notconv += cnt.ok ? 0 : 1;
            n_fact += cnt.nfact;
            n_backup += cnt.nbackup;
            n_fail += cnt.nfail;
            n_rounds += cnt.rounds;
            max_rounds = max_rounds > cnt.rounds ? max_rounds : cnt.rounds;
#else
            n_notconv += cnt.h[0];
            n_fact += cnt.h[1];
            n_backup += cnt.h[2];
            n_fail += cnt.h[3];
            n_rounds += cnt.h[4];
            max_rounds += cnt.h[5];
#endif
        }
    }
    if (stats && lane == 0 && n_cols) {
        atomicAdd((unsigned long long*)&stats[0], (unsigned long long)n_notconv);
        atomicAdd((unsigned long long*)&stats[1], (unsigned long long)n_fact);
        atomicAdd((unsigned long long*)&stats[2], (unsigned long long)n_backup);
        atomicAdd((unsigned long long*)&stats[3], (unsigned long long)n_fail);
        atomicAdd((unsigned long long*)&stats[4], (unsigned long long)n_rounds);
        atomicMax(&stats[5], max_rounds);
        atomicAdd((unsigned long long*)&stats[6], (unsigned long long)n_cols);
    }
}

// Half-warp solver (bpp16.cuh): 8 tiles of 16 lanes per CTA, one column per tile.
#define BPP16_THREADS 128
template <typename T, int NF>
__global__ void __launch_bounds__(BPP16_THREADS, sizeof(T) == 8 ? 4 : (NF <= 2 ? 7 : 6))
bpp16_kernel(const double* __restrict__ G, const double* __restrict__ Ginv, const T* R, int64_t ldr, T* X, int64_t ldx,
             T* __restrict__ Y, int64_t ldy, uint64_t* __restrict__ mask, int warm,
             const int64_t* __restrict__ seg_off, int k, long long* __restrict__ stats, int* __restrict__ defer_count,
             int32_t* __restrict__ defer_list, int64_t list_cap) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int seg = blockIdx.y;
    const int64_t c_begin = seg_off[seg];
    const int64_t ncols = seg_off[seg + 1] - c_begin;
    constexpr int TILES = BPP16_THREADS / 16;
    const int tile = threadIdx.x >> 4, t = threadIdx.x & 15;
    const int half = (threadIdx.x >> 4) & 1;
    if ((int64_t)blockIdx.x * TILES >= ncols) return;

    const int ldg = k | 1;
    T* Gs = smem;
    const double* Gseg = G + (size_t)seg * k * k;
    for (int e = threadIdx.x; e < k * k; e += blockDim.x) Gs[(e / k) * ldg + (e % k)] = (T)Gseg[e];
    size_t goff = ((size_t)k * ldg + 3) & ~(size_t)3;
    const T* Gis = nullptr;
    if (Ginv != nullptr) {
        const double* Iseg = Ginv + (size_t)seg * k * k;
        T* Gi = smem + goff;
        for (int e = threadIdx.x; e < k * k; e += blockDim.x) Gi[(e / k) * ldg + (e % k)] = (T)Iseg[e];
        if (Iseg[0] == Iseg[0]) Gis = Gi;
        goff *= 2;
    }
    __syncthreads();
    Bpp16Scratch<T> sc = bpp16_carve<T>(smem + goff + (size_t)tile * bpp16_tile_scratch_elems<T>(NF), NF);
    if (t < 8) reinterpret_cast<uint32_t*>(sc.idx)[t] = 0u;  // padded index reads stay in range
    if (t < 4) sc.z[16 + t] = T(0);
    __syncwarp();

    long long n_notconv = 0, n_fact = 0, n_backup = 0, n_fail = 0, n_rounds = 0, max_rounds = 0, n_cols = 0, n_defer = 0;
    // warp-uniform trip count: the two tiles of a warp work in lockstep, a tile past the end is predicated off
    for (int64_t cb = (int64_t)blockIdx.x * TILES; cb < ncols; cb += (int64_t)gridDim.x * TILES) {
        const int64_t c = cb + tile;
        const bool valid = c < ncols;
        const int64_t col = c_begin + (valid ? c : 0);
        const T* rp = R + col * ldr;
        T r[NF], x[NF], y[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) r[f] = (valid && t + 16 * f < k) ? rp[t + 16 * f] : T(0);
        uint64_t P = (valid && (warm & 1) && mask) ? mask[col] : 0ull;
        BppCounters cnt = {0, 0, 0, 0, true, {0, 0, 0, 0, 0, 0}};
        const int status = bpp16_column<T, NF>(Gs, Gis, ldg, sc, k, valid, r, P, (warm & 1) != 0, (warm & 2) != 0, x, y,
                                               cnt, t, half);
        if (status == 2) {  // too large for a tile: the full-warp kernel redoes this column from its original start set
            if (t == 0) {
                const int slot = atomicAdd(&defer_count[seg], 1);
                INMF_CHECK(slot >= 0 && slot < list_cap);
                defer_list[(size_t)seg * list_cap + slot] = (int32_t)c;
                n_defer++;
            }
        } else if (status == 1) {
            T* xo = X + col * ldx;
#pragma unroll
            for (int f = 0; f < NF; ++f)
                if (t + 16 * f < k) xo[t + 16 * f] = x[f];
            if (Y) {
                T* yo = Y + col * ldy;
                const bool clr = (warm & 4) != 0;
#pragma unroll
                for (int f = 0; f < NF; ++f)
                    if (t + 16 * f < k) yo[t + 16 * f] = clr ? T(0) : y[f];
            }
            if (mask && t == 0) mask[col] = P;
            if (t == 0) {
                n_cols++;
                n_notconv += cnt.ok ? 0 : 1;
                n_fact += cnt.nfact;
                n_backup += cnt.nbackup;
                n_fail += cnt.nfail;
                n_rounds += cnt.rounds;
                max_rounds = max_rounds > cnt.rounds ? max_rounds : cnt.rounds;
            }
        }
    }
    if (stats && t == 0 && (n_cols || n_defer)) {
        atomicAdd((unsigned long long*)&stats[0], (unsigned long long)n_notconv);
        atomicAdd((unsigned long long*)&stats[1], (unsigned long long)n_fact);
        atomicAdd((unsigned long long*)&stats[2], (unsigned long long)n_backup);
        atomicAdd((unsigned long long*)&stats[3], (unsigned long long)n_fail);
        atomicAdd((unsigned long long*)&stats[4], (unsigned long long)n_rounds);
        atomicMax(&stats[5], max_rounds);
        atomicAdd((unsigned long long*)&stats[6], (unsigned long long)n_cols);
        atomicAdd((unsigned long long*)&stats[7], (unsigned long long)n_defer);
    }
}

// Workspace layout of inmf_bpp_*: [G^-1: nseg*k*k doubles][deferral counters: nseg int32, padded to 16 bytes]
// [deferral lists: nseg * max_seg_cols int32].
static inline size_t bpp_ws_ginv_bytes(int nseg, int k) { return (size_t)nseg * k * k * sizeof(double); }
static inline size_t bpp_ws_count_bytes(int nseg) { return (((size_t)nseg * sizeof(int)) + 15) & ~(size_t)15; }
extern "C" int64_t inmf_bpp_workspace_bytes(int64_t max_seg_cols, int32_t nseg, int32_t k) {
    if (max_seg_cols < 0 || nseg < 1 || k < 1) return -1;
    return (int64_t)(bpp_ws_ginv_bytes(nseg, k) + bpp_ws_count_bytes(nseg) + (size_t)nseg * max_seg_cols * sizeof(int32_t));
}

template <typename T, int NF>
static int32_t bpp16_launch_one(const double* G, const double* ginv, const T* R, int64_t ldr, T* X, int64_t ldx, T* Y,
                                int64_t ldy, uint64_t* mask, int32_t warm, const int64_t* seg_off, int32_t nseg,
                                int64_t max_seg_cols, int32_t k, int64_t* stats, int* defer_count, int32_t* defer_list,
                                cudaStream_t st) {
    const int ldg = k | 1;
    const size_t goff = (((size_t)k * ldg + 3) & ~(size_t)3) * (ginv ? 2 : 1);
    const size_t smem = (goff + (BPP16_THREADS / 16) * bpp16_tile_scratch_elems<T>(NF)) * sizeof(T);
    INMF_OPTIN_SMEM((bpp16_kernel<T, NF>), 227 * 1024);
    constexpr int TILES = BPP16_THREADS / 16;
    int64_t blocks = (max_seg_cols + TILES - 1) / TILES;
    int64_t cap = ((int64_t)inmf_num_sms() * 2 * 8 + nseg - 1) / nseg;
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    bpp16_kernel<T, NF><<<dim3((unsigned)blocks, (unsigned)nseg), BPP16_THREADS, smem, st>>>(
        G, ginv, R, ldr, X, ldx, Y, ldy, mask, warm, seg_off, k, (long long*)stats, defer_count, defer_list, max_seg_cols);
    INMF_LAUNCH_CHECK();
    return 0;
}

template <typename T>
static int32_t bpp_launch(const double* G, const T* R, int64_t ldr, T* X, int64_t ldx, T* Y, int64_t ldy,
                          uint64_t* mask, int32_t warm, const int64_t* seg_off, int32_t nseg,
                          int64_t max_seg_cols, int32_t k, int64_t* stats, void* ws, int64_t ws_bytes, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K, "k must be in [1, 64]");
    INMF_REQUIRE(nseg >= 1 && nseg <= 65535, "nseg out of range");
    INMF_REQUIRE(ldr >= k && ldx >= k && (Y == nullptr || ldy >= k), "leading dimensions must be >= k");
    if (max_seg_cols <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    // workspace: the inverse needs the first part, the half-warp solver's deferral lists the rest
    double* ginv_ws = (ws != nullptr && ws_bytes >= (int64_t)bpp_ws_ginv_bytes(nseg, k)) ? (double*)ws : nullptr;
    INMF_REQUIRE(ws == nullptr || ginv_ws != nullptr, "workspace smaller than nseg * k * k doubles");
    if (warm & 16) ginv_ws = nullptr;  // the caller does not want the inverse formed (small warm-started solves)
    const bool have_lists = ws != nullptr && ws_bytes >= inmf_bpp_workspace_bytes(max_seg_cols, nseg, k) &&
                            max_seg_cols < ((int64_t)1 << 31);
    const int ldg = k | 1;
    const size_t goff = (((size_t)k * ldg + 3) & ~(size_t)3) * (ginv_ws ? 2 : 1);
    const size_t per_warp = bpp_warp_scratch_elems<T>(k) * sizeof(T);
    int nwarps = BPP_THREADS / 32;
    while (nwarps > 1 && goff * sizeof(T) + nwarps * per_warp > 64 * 1024) nwarps >>= 1;
    const size_t smem = goff * sizeof(T) + nwarps * per_warp;
    INMF_REQUIRE(smem <= 227 * 1024, "shared memory budget exceeded");
    INMF_OPTIN_SMEM((bpp_kernel<T, false>), 227 * 1024);
    INMF_OPTIN_SMEM((bpp_kernel<T, true>), 227 * 1024);
    // Grid: about two waves of resident CTAs over ALL segments, each warp striding over its share of the columns,
    // so that the per-CTA staging of G (and G^-1) into shared memory is amortised over many columns.
    // Small batches (the gene-side solves: a few columns per warp) get ONE wave: with two, half of the CTAs stage G only
    // after the first half has finished, and the staging is as long as the few columns a warp then solves (ncu, C2 V
    // solve, 12 000 columns: 22 % of the stall samples on the fp64 -> T conversion of the staged G).
    int64_t blocks = (max_seg_cols + nwarps - 1) / nwarps;
    const int64_t one_wave = (int64_t)inmf_num_sms() * BPP_MIN_BLOCKS;
    const int waves = ((int64_t)nseg * max_seg_cols >= one_wave * nwarps * 16) ? 2 : 1;
    int64_t cap = (one_wave * waves + nseg - 1) / nseg;
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    if (ginv_ws != nullptr) {
        // the inverse is used in T arithmetic: accept it only while cond(G) * eps(T) stays small
        const double rcond_min = sizeof(T) == 4 ? 1e-4 : 1e-10;
        ginv_kernel<<<(unsigned)nseg, 1024, 0, st>>>(G, ginv_ws, k, rcond_min);
        INMF_LAUNCH_CHECK();
    }
    const bool comp_kernel = k > BPP_FAST_MAX_P || ((warm & 2) && ginv_ws != nullptr);
    // Half-warp solver first (two columns per warp) when its partitions can be expected to fit 15 unknowns: always
    // for k <= 32 (one side of every partition has <= 16), for larger k only warm-started and with the inverse
    // (complement solves): measured at 10 M columns, a cold start at k = 40 / 60 defers 94 / 100 % of the columns
    // (99.6 vs 167 M columns/s with the full-warp kernel alone), a warm start 0.2 / 1 % (1 281 vs 625, 828 vs 268).
    // warm bit 3 switches it off.  What it defers is solved by the full-warp kernel in list mode.
    // Small batches (the gene-side V / W solves: 3 000 - 40 000 columns) are latency-bound either way and lose more to
    // the extra launches than they gain (C2: V 0.075 -> 0.102 ms, W 0.057 -> 0.067 ms): full-warp kernel only.
    const bool use16 = have_lists && !(warm & 8) && (k <= 32 || (ginv_ws != nullptr && (warm & 1))) &&
                       (int64_t)nseg * max_seg_cols >= 65536;
    int* defer_count = nullptr;
    int32_t* defer_list = nullptr;
    if (use16) {
        defer_count = reinterpret_cast<int*>((unsigned char*)ws + bpp_ws_ginv_bytes(nseg, k));
        defer_list = reinterpret_cast<int32_t*>((unsigned char*)defer_count + bpp_ws_count_bytes(nseg));
        INMF_CUDA_CHECK(cudaMemsetAsync(defer_count, 0, bpp_ws_count_bytes(nseg), st));
#define INMF_BPP16_ARGS G, ginv_ws, R, ldr, X, ldx, Y, ldy, mask, warm, seg_off, nseg, max_seg_cols, k, stats, defer_count, defer_list, st
        int32_t rc;
        if (k <= 32) rc = bpp16_launch_one<T, 2>(INMF_BPP16_ARGS);
        else if (k <= 48) rc = bpp16_launch_one<T, 3>(INMF_BPP16_ARGS);
        else rc = bpp16_launch_one<T, 4>(INMF_BPP16_ARGS);
#undef INMF_BPP16_ARGS
        if (rc != 0) return rc;
        // the deferred columns: the same grid as a full solve -- CTAs without work leave at once, before staging G, and
        // at k = 40 a real warm-started H solve defers a quarter of its columns (C4: 0.78 M of 2.75 M), which a small
        // grid (2 CTAs per SM) solved at a third of the kernel's throughput
    }
    dim3 grid((unsigned)blocks, (unsigned)nseg);
    if (comp_kernel)
        bpp_kernel<T, true><<<grid, nwarps * 32, smem, st>>>(G, ginv_ws, R, ldr, X, ldx, Y, ldy, mask, warm, seg_off, k,
                                                            (long long*)stats, defer_count, defer_list, max_seg_cols);
    else
        bpp_kernel<T, false><<<grid, nwarps * 32, smem, st>>>(G, ginv_ws, R, ldr, X, ldx, Y, ldy, mask, warm, seg_off, k,
                                                             (long long*)stats, defer_count, defer_list, max_seg_cols);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_bpp_f32(const double* G, const float* R, int64_t ldr, float* X, int64_t ldx, float* Y,
                                int64_t ldy, uint64_t* mask, int32_t warm, const int64_t* seg_off,
                                int32_t nseg, int64_t max_seg_cols, int32_t k, int64_t* stats, void* ws,
                                int64_t ws_bytes, void* stream) {
    return bpp_launch<float>(G, R, ldr, X, ldx, Y, ldy, mask, warm, seg_off, nseg, max_seg_cols, k, stats, ws,
                             ws_bytes, stream);
}
extern "C" int32_t inmf_bpp_f64(const double* G, const double* R, int64_t ldr, double* X, int64_t ldx,
                                double* Y, int64_t ldy, uint64_t* mask, int32_t warm, const int64_t* seg_off,
                                int32_t nseg, int64_t max_seg_cols, int32_t k, int64_t* stats, void* ws,
                                int64_t ws_bytes, void* stream) {
    return bpp_launch<double>(G, R, ldr, X, ldx, Y, ldy, mask, warm, seg_off, nseg, max_seg_cols, k, stats, ws,
                              ws_bytes, stream);
}

// =================================================================================================
// K1: H-update preparation (Mt, Gwv, Gvv) and G = Gwv + lambda Gvv
// =================================================================================================
template <typename T>
__global__ void prep_kernel(const T* __restrict__ Wt, const T* __restrict__ Vt, T* __restrict__ Mt, int64_t ldm,
                            double* __restrict__ Gwv, double* __restrict__ Gvv, int64_t g, int k) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    T* tm = reinterpret_cast<T*>(smem_raw);
    const int ldt = (int)ldm;  // multiple of 4, >= k
    T* tv = tm + (size_t)GRAM_TILE_ROWS * ldt;
    const int s = blockIdx.y;
    const T* V = Vt ? Vt + (size_t)s * g * k : nullptr;
    T* M = Mt + (size_t)s * g * ldm;
    for (int64_t r0 = (int64_t)blockIdx.x * GRAM_TILE_ROWS; r0 < g; r0 += (int64_t)gridDim.x * GRAM_TILE_ROWS) {
        const int rows = (int)((g - r0 < GRAM_TILE_ROWS) ? (g - r0) : GRAM_TILE_ROWS);
        for (int t = threadIdx.x; t < GRAM_TILE_ROWS * ldt; t += blockDim.x) {
            const int r = t / ldt, j = t % ldt;
            T w = T(0), v = T(0);
            if (r < rows && j < k) {
                w = Wt[(r0 + r) * k + j];
                if (V) v = V[(r0 + r) * k + j];
            }
            tm[t] = w + v;
            tv[t] = v;
            if (r < rows) M[(r0 + r) * ldm + j] = w + v;
        }
        __syncthreads();
        tile_gram_accumulate<T>(tm, ldt, rows, k, Gwv + (size_t)s * k * k, 1.0);
        if (V) tile_gram_accumulate<T>(tv, ldt, rows, k, Gvv + (size_t)s * k * k, 1.0);
        __syncthreads();
    }
}

__global__ void combine_g_kernel(const double* __restrict__ Gwv, const double* __restrict__ Gvv,
                                 double* __restrict__ G, double lambda, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) G[i] = Gwv[i] + lambda * Gvv[i];
}

template <typename T>
static int32_t prep_launch(const T* Wt, const T* Vt, T* Mt, int64_t ldm, double* G, double* Gwv, double* Gvv,
                           double lambda, int32_t nseg, int64_t g, int32_t k, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K, "k must be in [1, 64]");
    INMF_REQUIRE(ldm >= k && ldm % 4 == 0 && ldm <= 64, "ldm must be a multiple of 4 in [k, 64]");
    INMF_REQUIRE(nseg >= 1 && g >= 1, "empty problem");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t kk = (size_t)nseg * k * k * sizeof(double);
    INMF_CUDA_CHECK(cudaMemsetAsync(Gwv, 0, kk, st));
    INMF_CUDA_CHECK(cudaMemsetAsync(Gvv, 0, kk, st));
    int64_t tiles = (g + GRAM_TILE_ROWS - 1) / GRAM_TILE_ROWS;
    if (tiles > 1024) tiles = 1024;
    if (inmf_deterministic()) tiles = 1;  // one CTA per segment: every Gram entry is summed by one thread, in row order
    const size_t smem = 2 * (size_t)GRAM_TILE_ROWS * ldm * sizeof(T);
    INMF_OPTIN_SMEM((prep_kernel<T>), 100 * 1024);
    prep_kernel<T><<<dim3((unsigned)tiles, (unsigned)nseg), 256, smem, st>>>(Wt, Vt, Mt, ldm, Gwv, Gvv, g, k);
    INMF_LAUNCH_CHECK();
    const int n = nseg * k * k;
    combine_g_kernel<<<(n + 255) / 256, 256, 0, st>>>(Gwv, Gvv, G, lambda, n);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_prep_f32(const float* Wt, const float* Vt, float* Mt, int64_t ldm, double* G,
                                 double* Gwv, double* Gvv, double lambda, int32_t nseg, int64_t g, int32_t k,
                                 void* stream) {
    return prep_launch<float>(Wt, Vt, Mt, ldm, G, Gwv, Gvv, lambda, nseg, g, k, stream);
}
extern "C" int32_t inmf_prep_f64(const double* Wt, const double* Vt, double* Mt, int64_t ldm, double* G,
                                 double* Gwv, double* Gvv, double lambda, int32_t nseg, int64_t g, int32_t k,
                                 void* stream) {
    return prep_launch<double>(Wt, Vt, Mt, ldm, G, Gwv, Gvv, lambda, nseg, g, k, stream);
}

// =================================================================================================
// K2: CSR SpMM  R[q,:] = X[row(q),:] * Mt[s]      (warp per cell, lanes over factors)
// =================================================================================================
template <typename T>
__global__ void spmm_rows_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                 const T* __restrict__ vals, const T* __restrict__ Mt, int64_t ldm, int64_t g,
                                 T* __restrict__ R, int64_t ldr, const int64_t* __restrict__ seg_desc, int k) {
    const int s = blockIdx.y;
    const SegWin w = load_seg(seg_desc, s);
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const T* __restrict__ M = Mt + (size_t)s * g * ldm;
    const bool has0 = lane < k, has1 = lane + 32 < k;
    for (int64_t j = (int64_t)blockIdx.x * nwarps + warp; j < w.nrows; j += (int64_t)gridDim.x * nwarps) {
        const int64_t row = seg_row(w, j);
        const int64_t beg = rowptr[row], end = rowptr[row + 1];
        T acc0 = T(0), acc1 = T(0);
        for (int64_t p = beg; p < end; p += 32) {
            const bool mine = p + lane < end;
            const int32_t c = mine ? colidx[p + lane] : 0;
            const T v = mine ? vals[p + lane] : T(0);
            const int cnt = (end - p < 32) ? (int)(end - p) : 32;
            for (int t = 0; t < cnt; ++t) {
                const int32_t cc = __shfl_sync(INMF_FULL_MASK, c, t);
                const T vv = __shfl_sync(INMF_FULL_MASK, v, t);
                const T* m = M + (size_t)cc * ldm;
                if (has0) acc0 += vv * m[lane];
                if (has1) acc1 += vv * m[lane + 32];
            }
        }
        T* out = R + (w.out_off + j) * ldr;
        if (has0) out[lane] = acc0;
        if (has1) out[lane + 32] = acc1;
    }
}

template <typename T>
static int32_t spmm_launch(const int64_t* rowptr, const int32_t* colidx, const T* vals, const T* Mt, int64_t ldm,
                           int64_t g, T* R, int64_t ldr, const int64_t* seg_desc, int32_t nseg,
                           int64_t max_seg_rows, int32_t k, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K, "k must be in [1, 64]");
    INMF_REQUIRE(ldm >= k && ldr >= k, "leading dimensions must be >= k");
    INMF_REQUIRE(nseg >= 1 && nseg <= 65535, "nseg out of range");
    if (max_seg_rows <= 0) return 0;
    const int nwarps = 8;
    int64_t blocks = (max_seg_rows + nwarps - 1) / nwarps;
    const int64_t cap = (int64_t)inmf_num_sms() * 32;
    if (blocks > cap) blocks = cap;
    spmm_rows_kernel<T><<<dim3((unsigned)blocks, (unsigned)nseg), nwarps * 32, 0, (cudaStream_t)stream>>>(
        rowptr, colidx, vals, Mt, ldm, g, R, ldr, seg_desc, k);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_spmm_f32(const int64_t* rowptr, const int32_t* colidx, const float* vals,
                                 const float* Mt, int64_t ldm, int64_t g, float* R, int64_t ldr,
                                 const int64_t* seg_desc, int32_t nseg, int64_t max_seg_rows, int32_t k,
                                 void* stream) {
    return spmm_launch<float>(rowptr, colidx, vals, Mt, ldm, g, R, ldr, seg_desc, nseg, max_seg_rows, k, stream);
}
extern "C" int32_t inmf_spmm_f64(const int64_t* rowptr, const int32_t* colidx, const double* vals,
                                 const double* Mt, int64_t ldm, int64_t g, double* R, int64_t ldr,
                                 const int64_t* seg_desc, int32_t nseg, int64_t max_seg_rows, int32_t k,
                                 void* stream) {
    return spmm_launch<double>(rowptr, colidx, vals, Mt, ldm, g, R, ldr, seg_desc, nseg, max_seg_rows, k, stream);
}

// =================================================================================================
// K4/K5/K11: statistics  St[s] += X' H,  HtH[s] += H' H
// =================================================================================================
template <typename T>
__global__ void stats_scatter_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                     const T* __restrict__ vals, const T* __restrict__ H, int64_t ldh,
                                     T* __restrict__ St, int64_t g, const int64_t* __restrict__ seg_desc, int k) {
    const int s = blockIdx.y;
    const SegWin w = load_seg(seg_desc, s);
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    T* __restrict__ S = St + (size_t)s * g * k;
    const bool has0 = lane < k, has1 = lane + 32 < k;
    for (int64_t j = (int64_t)blockIdx.x * nwarps + warp; j < w.nrows; j += (int64_t)gridDim.x * nwarps) {
        const int64_t row = seg_row(w, j);
        const int64_t beg = rowptr[row], end = rowptr[row + 1];
        const T* h = H + (w.out_off + j) * ldh;
        const T h0 = has0 ? h[lane] : T(0);
        const T h1 = has1 ? h[lane + 32] : T(0);
        for (int64_t p = beg; p < end; p += 32) {
            const bool mine = p + lane < end;
            const int32_t c = mine ? colidx[p + lane] : 0;
            const T v = mine ? vals[p + lane] : T(0);
            const int cnt = (end - p < 32) ? (int)(end - p) : 32;
            for (int t = 0; t < cnt; ++t) {
                const int32_t cc = __shfl_sync(INMF_FULL_MASK, c, t);
                const T vv = __shfl_sync(INMF_FULL_MASK, v, t);
                T* dst = S + (size_t)cc * k;
                if (has0 && h0 != T(0)) atomicAdd(dst + lane, vv * h0);
                if (has1 && h1 != T(0)) atomicAdd(dst + lane + 32, vv * h1);
            }
        }
    }
}

template <typename T>
__global__ void gram_rows_kernel(const T* __restrict__ A, int64_t lda, const int64_t* __restrict__ seg_desc,
                                 double* __restrict__ out, int k) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    const int ldt = ((k + 3) & ~3) + 4;
    const int s = blockIdx.y;
    const int64_t r_begin = seg_desc[4 * s];
    const int64_t nrows = seg_desc[4 * (s + 1)] - r_begin;
    for (int64_t r0 = (int64_t)blockIdx.x * GRAM_TILE_ROWS; r0 < nrows; r0 += (int64_t)gridDim.x * GRAM_TILE_ROWS) {
        const int rows = (int)((nrows - r0 < GRAM_TILE_ROWS) ? (nrows - r0) : GRAM_TILE_ROWS);
        for (int t = threadIdx.x; t < GRAM_TILE_ROWS * ldt; t += blockDim.x) {
            const int r = t / ldt, j = t % ldt;
            tile[t] = (r < rows && j < k) ? A[(r_begin + r0 + r) * lda + j] : T(0);
        }
        __syncthreads();
        tile_gram_accumulate<T>(tile, ldt, rows, k, out + (size_t)s * k * k, 1.0);
        __syncthreads();
    }
}

template <typename T>
static int32_t stats_launch(const int64_t* rowptr, const int32_t* colidx, const T* vals, const T* H, int64_t ldh,
                            T* St, double* HtH, int64_t g, const int64_t* seg_desc, int32_t nseg,
                            int64_t max_seg_rows, int32_t k, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K, "k must be in [1, 64]");
    INMF_REQUIRE(ldh >= k, "ldh must be >= k");
    INMF_REQUIRE(nseg >= 1 && nseg <= 65535, "nseg out of range");
    if (max_seg_rows <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int nwarps = 8;
    int64_t blocks = (max_seg_rows + nwarps - 1) / nwarps;
    const int64_t cap = (int64_t)inmf_num_sms() * 32;
    if (blocks > cap) blocks = cap;
    stats_scatter_kernel<T><<<dim3((unsigned)blocks, (unsigned)nseg), nwarps * 32, 0, st>>>(
        rowptr, colidx, vals, H, ldh, St, g, seg_desc, k);
    INMF_LAUNCH_CHECK();
    int64_t tiles = (max_seg_rows + GRAM_TILE_ROWS - 1) / GRAM_TILE_ROWS;
    const int64_t tcap = (int64_t)inmf_num_sms() * 8;
    if (tiles > tcap) tiles = tcap;
    const int ldt = ((k + 3) & ~3) + 4;
    const size_t smem = (size_t)GRAM_TILE_ROWS * ldt * sizeof(T);
    gram_rows_kernel<T><<<dim3((unsigned)tiles, (unsigned)nseg), 256, smem, st>>>(H, ldh, seg_desc, HtH, k);
    INMF_LAUNCH_CHECK();
    return 0;
}
template <typename T>
static int32_t gram_rows_launch(const T* A, int64_t lda, const int64_t* seg_desc, int32_t nseg, int64_t max_seg_rows,
                                double* out, int32_t k, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K && lda >= k, "bad k / lda");
    INMF_REQUIRE(nseg >= 1 && nseg <= 65535, "nseg out of range");
    if (max_seg_rows <= 0) return 0;
    int64_t tiles = (max_seg_rows + GRAM_TILE_ROWS - 1) / GRAM_TILE_ROWS;
    const int64_t tcap = (int64_t)inmf_num_sms() * 8;
    if (tiles > tcap) tiles = tcap;
    const int ldt = ((k + 3) & ~3) + 4;
    gram_rows_kernel<T><<<dim3((unsigned)tiles, (unsigned)nseg), 256, (size_t)GRAM_TILE_ROWS * ldt * sizeof(T),
                          (cudaStream_t)stream>>>(A, lda, seg_desc, out, k);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_gram_rows_f32(const float* A, int64_t lda, const int64_t* seg_desc, int32_t nseg,
                                      int64_t max_seg_rows, double* out, int32_t k, void* stream) {
    return gram_rows_launch<float>(A, lda, seg_desc, nseg, max_seg_rows, out, k, stream);
}
extern "C" int32_t inmf_gram_rows_f64(const double* A, int64_t lda, const int64_t* seg_desc, int32_t nseg,
                                      int64_t max_seg_rows, double* out, int32_t k, void* stream) {
    return gram_rows_launch<double>(A, lda, seg_desc, nseg, max_seg_rows, out, k, stream);
}

extern "C" int32_t inmf_stats_f32(const int64_t* rowptr, const int32_t* colidx, const float* vals, const float* H,
                                  int64_t ldh, float* St, double* HtH, int64_t g, const int64_t* seg_desc,
                                  int32_t nseg, int64_t max_seg_rows, int32_t k, void* stream) {
    return stats_launch<float>(rowptr, colidx, vals, H, ldh, St, HtH, g, seg_desc, nseg, max_seg_rows, k, stream);
}
extern "C" int32_t inmf_stats_f64(const int64_t* rowptr, const int32_t* colidx, const double* vals,
                                  const double* H, int64_t ldh, double* St, double* HtH, int64_t g,
                                  const int64_t* seg_desc, int32_t nseg, int64_t max_seg_rows, int32_t k,
                                  void* stream) {
    return stats_launch<double>(rowptr, colidx, vals, H, ldh, St, HtH, g, seg_desc, nseg, max_seg_rows, k, stream);
}

// =================================================================================================
// ||X_s||^2
// =================================================================================================
template <typename T>
__global__ void sqnorm_kernel(const int64_t* __restrict__ rowptr, const T* __restrict__ vals,
                              const int64_t* __restrict__ seg_desc, double* __restrict__ out) {
    const int s = blockIdx.y;
    const int64_t row_base = seg_desc[4 * s + 1], n = seg_desc[4 * s + 3];
    const int64_t beg = rowptr[row_base], end = rowptr[row_base + n];
    double acc = 0.0;
    for (int64_t p = beg + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < end;
         p += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)vals[p];
        acc += v * v;
    }
    acc = warp_sum(acc);
    __shared__ double part[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        acc = (lane < (int)(blockDim.x >> 5)) ? part[lane] : 0.0;
        acc = warp_sum(acc);
        if (lane == 0) atomicAdd(&out[s], acc);
    }
}
template <typename T>
static int32_t sqnorm_launch(const int64_t* rowptr, const T* vals, const int64_t* seg_desc, int32_t nseg,
                             double* out, void* stream) {
    INMF_REQUIRE(nseg >= 1 && nseg <= 65535, "nseg out of range");
    cudaStream_t st = (cudaStream_t)stream;
    INMF_CUDA_CHECK(cudaMemsetAsync(out, 0, nseg * sizeof(double), st));
    // deterministic mode: one CTA per segment (fixed summation tree)
    sqnorm_kernel<T><<<dim3(inmf_deterministic() ? 1 : inmf_num_sms() * 2, nseg), inmf_deterministic() ? 1024 : 256, 0,
                       st>>>(rowptr, vals, seg_desc, out);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_sqnorm_f32(const int64_t* rowptr, const float* vals, const int64_t* seg_desc,
                                   int32_t nseg, double* out, void* stream) {
    return sqnorm_launch<float>(rowptr, vals, seg_desc, nseg, out, stream);
}
extern "C" int32_t inmf_sqnorm_f64(const int64_t* rowptr, const double* vals, const int64_t* seg_desc,
                                   int32_t nseg, double* out, void* stream) {
    return sqnorm_launch<double>(rowptr, vals, seg_desc, nseg, out, stream);
}

// =================================================================================================
// V / W right-hand sides from statistics (warp per gene)
// =================================================================================================
// out[lane], out[lane+32] of  row(1 x k) * C(k x k)   with row distributed over lanes; C staged in shared memory as T
// (the global fp64 matrix was read once per gene before: 19 / 34 us per call at C2 for a few hundred thousand FMAs).
template <typename T>
__device__ __forceinline__ void row_times_sym(const T* __restrict__ C, int k, T v0, T v1, T& o0, T& o1) {
    const int lane = threadIdx.x & 31;
    const int c0 = lane < k ? lane : 0, c1 = lane + 32 < k ? lane + 32 : 0;
    T a0 = T(0), a1 = T(0);
    const int k32 = k < 32 ? k : 32;
#pragma unroll 4
    for (int a = 0; a < k32; ++a) {
        const T va = __shfl_sync(INMF_FULL_MASK, v0, a);
        a0 = fma(va, C[a * k + c0], a0);
        if (k > 32) a1 = fma(va, C[a * k + c1], a1);
    }
    for (int a = 32; a < k; ++a) {
        const T va = __shfl_sync(INMF_FULL_MASK, v1, a - 32);
        a0 = fma(va, C[a * k + c0], a0);
        a1 = fma(va, C[a * k + c1], a1);
    }
    o0 = a0;
    o1 = a1;
}

template <typename T>
__global__ void __launch_bounds__(256) rhs_v_kernel(const T* __restrict__ St, const T* __restrict__ Wt,
                                                    const double* __restrict__ HtH, T* __restrict__ Rv,
                                                    double* __restrict__ Gv, double lambda, int64_t g, int k) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    T* Cs = reinterpret_cast<T*>(smem_raw);
    const int s = blockIdx.y;
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* C = HtH + (size_t)s * k * k;
    for (int t = threadIdx.x; t < k * k; t += blockDim.x) {
        const double c = C[t];
        Cs[t] = (T)c;
        if (blockIdx.x == 0) Gv[(size_t)s * k * k + t] = (1.0 + lambda) * c;
    }
    __syncthreads();
    for (int64_t c = (int64_t)blockIdx.x * nwarps + warp; c < g; c += (int64_t)gridDim.x * nwarps) {
        const T* w = Wt + c * k;
        const T w0 = (lane < k) ? w[lane] : T(0);
        const T w1 = (lane + 32 < k) ? w[lane + 32] : T(0);
        T o0, o1;
        row_times_sym<T>(Cs, k, w0, w1, o0, o1);
        const T* sr = St + ((size_t)s * g + c) * k;
        T* out = Rv + ((size_t)s * g + c) * k;
        if (lane < k) out[lane] = sr[lane] - o0;
        if (lane + 32 < k) out[lane + 32] = sr[lane + 32] - o1;
    }
}

// One gene per warp; the H'H of up to `stage` datasets are staged in shared memory at a time, and a gene's rows of
// V_s / St_s for all staged datasets are requested before any of them is used (the loads are independent).
#define RHS_W_MAX_STAGE 8
template <typename T>
__global__ void __launch_bounds__(256) rhs_w_kernel(const T* __restrict__ St, const T* __restrict__ Vt,
                                                    const double* __restrict__ HtH, T* __restrict__ Rw,
                                                    double* __restrict__ Gw, int nseg, int64_t g, int k, int stage) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    T* Cs = reinterpret_cast<T*>(smem_raw);
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kk = k * k;
    if (blockIdx.x == gridDim.x - 1)  // (the last CTA has the fewest genes)
        for (int t = threadIdx.x; t < kk; t += blockDim.x) {
            double acc = 0.0;
            for (int s = 0; s < nseg; ++s) acc += HtH[(size_t)s * kk + t];
            Gw[t] = acc;
        }
    const bool has0 = lane < k, has1 = lane + 32 < k;
    for (int64_t base = (int64_t)blockIdx.x * nwarps; base < g; base += (int64_t)gridDim.x * nwarps) {
        const int64_t c = base + warp;
        T r0 = T(0), r1 = T(0);
        for (int s0 = 0; s0 < nseg; s0 += stage) {
            const int ns = (nseg - s0 < stage) ? (nseg - s0) : stage;
            T v0[RHS_W_MAX_STAGE], v1[RHS_W_MAX_STAGE], q0[RHS_W_MAX_STAGE], q1[RHS_W_MAX_STAGE];
#pragma unroll
            for (int s = 0; s < RHS_W_MAX_STAGE; ++s) {
                v0[s] = v1[s] = q0[s] = q1[s] = T(0);
                if (s < ns && c < g) {
                    const size_t o = ((size_t)(s0 + s) * g + c) * k;
                    if (has0) { v0[s] = Vt[o + lane]; q0[s] = St[o + lane]; }
                    if (has1) { v1[s] = Vt[o + lane + 32]; q1[s] = St[o + lane + 32]; }
                }
            }
            __syncthreads();
            for (int t = threadIdx.x; t < ns * kk; t += blockDim.x) Cs[t] = (T)HtH[(size_t)s0 * kk + t];
            __syncthreads();
#pragma unroll
            for (int s = 0; s < RHS_W_MAX_STAGE; ++s) {
                if (s < ns) {
                    T o0, o1;
                    row_times_sym<T>(Cs + (size_t)s * kk, k, v0[s], v1[s], o0, o1);
                    r0 += q0[s] - o0;
                    r1 += q1[s] - o1;
                }
            }
        }
        if (c < g) {
            T* out = Rw + c * k;
            if (has0) out[lane] = r0;
            if (has1) out[lane + 32] = r1;
        }
    }
}

template <typename T>
static int32_t rhs_v_launch(const T* St, const T* Wt, const double* HtH, T* Rv, double* Gv, double lambda,
                            int32_t nseg, int64_t g, int32_t k, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K && nseg >= 1 && g >= 1, "bad shape");
    int64_t blocks = (g + 15) / 16;  // two genes per warp: the staging of C is amortised, the loads stay parallel
    if (blocks > inmf_num_sms() * 8) blocks = inmf_num_sms() * 8;
    rhs_v_kernel<T><<<dim3((unsigned)blocks, (unsigned)nseg), 256, (size_t)k * k * sizeof(T), (cudaStream_t)stream>>>(
        St, Wt, HtH, Rv, Gv, lambda, g, k);
    INMF_LAUNCH_CHECK();
    return 0;
}
template <typename T>
static int32_t rhs_w_launch(const T* St, const T* Vt, const double* HtH, T* Rw, double* Gw, int32_t nseg,
                            int64_t g, int32_t k, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K && nseg >= 1 && g >= 1, "bad shape");
    int64_t blocks = (g + 7) / 8;
    if (blocks > inmf_num_sms() * 8) blocks = inmf_num_sms() * 8;
    int stage = (int)((size_t)(64 * 1024) / ((size_t)k * k * sizeof(T)));  // datasets staged at a time (<= 64 KB)
    if (stage < 1) stage = 1;
    if (stage > RHS_W_MAX_STAGE) stage = RHS_W_MAX_STAGE;
    if (stage > nseg) stage = nseg;
    const size_t smem = (size_t)stage * k * k * sizeof(T);
    INMF_OPTIN_SMEM((rhs_w_kernel<T>), 64 * 1024);
    rhs_w_kernel<T><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(St, Vt, HtH, Rw, Gw, nseg, g, k, stage);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_rhs_v_f32(const float* St, const float* Wt, const double* HtH, float* Rv, double* Gv,
                                  double lambda, int32_t nseg, int64_t g, int32_t k, void* stream) {
    return rhs_v_launch<float>(St, Wt, HtH, Rv, Gv, lambda, nseg, g, k, stream);
}
extern "C" int32_t inmf_rhs_v_f64(const double* St, const double* Wt, const double* HtH, double* Rv, double* Gv,
                                  double lambda, int32_t nseg, int64_t g, int32_t k, void* stream) {
    return rhs_v_launch<double>(St, Wt, HtH, Rv, Gv, lambda, nseg, g, k, stream);
}
extern "C" int32_t inmf_rhs_w_f32(const float* St, const float* Vt, const double* HtH, float* Rw, double* Gw,
                                  int32_t nseg, int64_t g, int32_t k, void* stream) {
    return rhs_w_launch<float>(St, Vt, HtH, Rw, Gw, nseg, g, k, stream);
}
extern "C" int32_t inmf_rhs_w_f64(const double* St, const double* Vt, const double* HtH, double* Rw, double* Gw,
                                  int32_t nseg, int64_t g, int32_t k, void* stream) {
    return rhs_w_launch<double>(St, Vt, HtH, Rw, Gw, nseg, g, k, stream);
}

// =================================================================================================
// K8: objective from statistics
// =================================================================================================
template <typename T>
__global__ void objective_kernel(const T* __restrict__ St, const T* __restrict__ Mt, int64_t ldm,
                                 const double* __restrict__ HtH, const double* __restrict__ Gwv,
                                 const double* __restrict__ Gvv, const double* __restrict__ xnorm2, double lambda,
                                 int64_t g, int k, int nseg, double* __restrict__ obj) {
    double acc = 0.0;
    // (grid.y == nseg: one segment per block row; grid = (1, 1) in deterministic mode: one CTA walks all segments,
    // so the only cross-thread sums are the fixed-order reductions below)
    for (int s = blockIdx.y; s < nseg; s += gridDim.y) {
        const T* S = St + (size_t)s * g * k;
        const T* M = Mt + (size_t)s * g * ldm;
        double a2 = 0.0;
        const int64_t total = g * k;
        for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
             t += (int64_t)gridDim.x * blockDim.x) {
            const int64_t c = t / k;
            const int j = (int)(t - c * k);
            a2 += (double)S[t] * (double)M[c * ldm + j];
        }
        acc += -2.0 * a2;
        if (blockIdx.x == 0) {
            const double* C = HtH + (size_t)s * k * k;
            const double* A = Gwv + (size_t)s * k * k;
            const double* B = Gvv + (size_t)s * k * k;
            for (int t = threadIdx.x; t < k * k; t += blockDim.x) acc += C[t] * (A[t] + lambda * B[t]);
            if (threadIdx.x == 0) acc += xnorm2[s];
        }
    }
    acc = warp_sum(acc);
    __shared__ double part[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        acc = (lane < (int)(blockDim.x >> 5)) ? part[lane] : 0.0;
        acc = warp_sum(acc);
        if (lane == 0) atomicAdd(obj, acc);
    }
}
template <typename T>
static int32_t objective_launch(const T* St, const T* Mt, int64_t ldm, const double* HtH, const double* Gwv,
                                const double* Gvv, const double* xnorm2, double lambda, int32_t nseg, int64_t g,
                                int32_t k, double* obj, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K && nseg >= 1 && g >= 1, "bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    INMF_CUDA_CHECK(cudaMemsetAsync(obj, 0, sizeof(double), st));
    int64_t blocks = (g * k + 255) / 256;
    if (blocks > 64) blocks = 64;
    const bool det = inmf_deterministic();
    objective_kernel<T><<<dim3(det ? 1u : (unsigned)blocks, det ? 1u : (unsigned)nseg), det ? 1024 : 256, 0, st>>>(
        St, Mt, ldm, HtH, Gwv, Gvv, xnorm2, lambda, g, k, nseg, obj);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_objective_f32(const float* St, const float* Mt, int64_t ldm, const double* HtH,
                                      const double* Gwv, const double* Gvv, const double* xnorm2, double lambda,
                                      int32_t nseg, int64_t g, int32_t k, double* obj, void* stream) {
    return objective_launch<float>(St, Mt, ldm, HtH, Gwv, Gvv, xnorm2, lambda, nseg, g, k, obj, stream);
}
extern "C" int32_t inmf_objective_f64(const double* St, const double* Mt, int64_t ldm, const double* HtH,
                                      const double* Gwv, const double* Gvv, const double* xnorm2, double lambda,
                                      int32_t nseg, int64_t g, int32_t k, double* obj, void* stream) {
    return objective_launch<double>(St, Mt, ldm, HtH, Gwv, Gvv, xnorm2, lambda, nseg, g, k, obj, stream);
}

// =================================================================================================
// K11 epilogue: _update_A_B
// =================================================================================================
template <typename T>
__global__ void update_ab_kernel(double* __restrict__ A, double* __restrict__ A_old, T* __restrict__ B,
                                 T* __restrict__ B_old, const double* __restrict__ HHt, const T* __restrict__ XHt,
                                 const double* __restrict__ inv_mb, double scale, int rollover, int64_t g, int k) {
    const int s = blockIdx.y;
    const double im = inv_mb[s];
    const int64_t nb = g * k;
    const int kk = k * k;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nb + kk;
         t += (int64_t)gridDim.x * blockDim.x) {
        if (t < nb) {
            const size_t o = (size_t)s * nb + t;
            double b = (double)B[o], bo = (double)B_old[o];
            if (rollover) {
                b = b - bo;
                bo = scale * b;
            } else {
                bo = scale * bo;
            }
            b = scale * b + (double)XHt[o] * im;
            B[o] = (T)b;
            B_old[o] = (T)bo;
        } else {
            const int e = (int)(t - nb);
            const size_t o = (size_t)s * kk + e;
            double a = A[o], ao = A_old[o];
            if (rollover) {
                a = a - ao;
                ao = scale * a;
            } else {
                ao = scale * ao;
            }
            a = scale * a + HHt[o] * im;
            if ((e / k) == (e % k) && a == 0.0) a = 1e-15;  // _online_iNMF.py:596-599
            A[o] = a;
            A_old[o] = ao;
        }
    }
}
template <typename T>
static int32_t update_ab_launch(double* A, double* A_old, T* B, T* B_old, const double* HHt, const T* XHt,
                                const double* inv_mb, double scale, int32_t rollover, int32_t nseg, int64_t g,
                                int32_t k, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K && nseg >= 1 && g >= 1, "bad shape");
    int64_t blocks = (g * k + k * k + 255) / 256;
    if (blocks > inmf_num_sms() * 4) blocks = inmf_num_sms() * 4;
    update_ab_kernel<T><<<dim3((unsigned)blocks, (unsigned)nseg), 256, 0, (cudaStream_t)stream>>>(
        A, A_old, B, B_old, HHt, XHt, inv_mb, scale, rollover, g, k);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_update_ab_f32(double* A, double* A_old, float* B, float* B_old, const double* HHt,
                                      const float* XHt, const double* inv_mb, double scale, int32_t rollover,
                                      int32_t nseg, int64_t g, int32_t k, void* stream) {
    return update_ab_launch<float>(A, A_old, B, B_old, HHt, XHt, inv_mb, scale, rollover, nseg, g, k, stream);
}
extern "C" int32_t inmf_update_ab_f64(double* A, double* A_old, double* B, double* B_old, const double* HHt,
                                      const double* XHt, const double* inv_mb, double scale, int32_t rollover,
                                      int32_t nseg, int64_t g, int32_t k, void* stream) {
    return update_ab_launch<double>(A, A_old, B, B_old, HHt, XHt, inv_mb, scale, rollover, nseg, g, k, stream);
}

// =================================================================================================
// K12/K13: HALS sweeps, warp per gene row
// =================================================================================================
#define HALS_MAX_SETS 16  // datasets whose V sweeps are reduced together (register arrays)

// A is staged TRANSPOSED in shared memory (At[i][j][a] = A_i[a][j]) so that the per-factor dot products
// read consecutive addresses; when it does not fit, the same index is read from global memory.
template <typename T, bool A_IN_SMEM>
__global__ void hals_kernel(const double* __restrict__ A, const T* __restrict__ B, T* __restrict__ Wt,
                            T* __restrict__ Vt, double lambda, int n_prev, int nall, int64_t g, int k, int n_inner) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kk = k * k;
    T* At = reinterpret_cast<T*>(smem_raw);  // [nall][k][k] transposed (only if A_IN_SMEM)
    const size_t a_elems = A_IN_SMEM ? (((size_t)nall * kk + 3) & ~(size_t)3) : 0;
    if (A_IN_SMEM) {
        for (int e = threadIdx.x; e < nall * kk; e += blockDim.x) {
            const int i = e / kk, rem = e - i * kk, j = rem / k, a = rem - j * k;
            At[e] = (T)A[(size_t)i * kk + (size_t)a * k + j];
        }
        __syncthreads();
    }
    const int per_warp = (1 + 2 * nall) * k;
    T* ws = reinterpret_cast<T*>(smem_raw) + a_elems + (size_t)warp * per_warp;
    T* vs = ws + k;                 // [nall][k]
    T* bs = vs + (size_t)nall * k;  // [nall][k]
    const T eps = T(1e-16);         // nonneg(), _utilities.py:10-13
    const T one_lam = (T)(1.0 + lambda);
    const int ncur = nall - n_prev;
    auto a_t = [&](int i, int j, int a) -> T {  // A_i[a][j]
        return A_IN_SMEM ? At[(size_t)i * kk + j * k + a] : (T)A[(size_t)i * kk + (size_t)a * k + j];
    };
    for (int64_t w = (int64_t)blockIdx.x * nwarps + warp; w < g; w += (int64_t)gridDim.x * nwarps) {
        for (int j = lane; j < k; j += 32) ws[j] = Wt[w * k + j];
        for (int i = 0; i < nall; ++i)
            for (int j = lane; j < k; j += 32) {
                vs[i * k + j] = Vt[((size_t)i * g + w) * k + j];
                bs[i * k + j] = B[((size_t)i * g + w) * k + j];
            }
        __syncwarp();
        for (int it = 0; it < n_inner; ++it) {
            // ---- W sweep (all datasets, previous ones included) ----
            for (int j = 0; j < k; ++j) {
                T part = T(0);
                for (int a = lane; a < k; a += 32) {
                    const T wa = ws[a];
                    for (int i = 0; i < nall; ++i) part = fma(wa + vs[i * k + a], a_t(i, j, a), part);
                }
                part = warp_sum(part);
                if (lane == 0) {
                    T num = -part;
                    T den = T(0);
                    for (int i = 0; i < nall; ++i) {
                        num += bs[i * k + j];
                        den += a_t(i, j, j);
                    }
                    const T nw = ws[j] + num / den;
                    ws[j] = nw < eps ? eps : nw;
                }
                __syncwarp();
            }
            // ---- V sweep (current datasets only): j outer, the datasets' reductions run together ----
            for (int j = 0; j < k; ++j) {
                for (int i0 = 0; i0 < ncur; i0 += HALS_MAX_SETS) {
                    const int cnt = (ncur - i0 < HALS_MAX_SETS) ? (ncur - i0) : HALS_MAX_SETS;
                    T part[HALS_MAX_SETS];
#pragma unroll
                    for (int u = 0; u < HALS_MAX_SETS; ++u) part[u] = T(0);
                    for (int a = lane; a < k; a += 32) {
                        const T wa = ws[a];
#pragma unroll
                        for (int u = 0; u < HALS_MAX_SETS; ++u) {
                            if (u < cnt) {
                                const int i = n_prev + i0 + u;
                                part[u] = fma(wa + one_lam * vs[i * k + a], a_t(i, j, a), part[u]);
                            }
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int u = 0; u < HALS_MAX_SETS; ++u)
                            if (u < cnt) part[u] += __shfl_xor_sync(INMF_FULL_MASK, part[u], o);
                    }
#pragma unroll
                    for (int u = 0; u < HALS_MAX_SETS; ++u) {
                        if (u < cnt && lane == u) {  // lane u finishes dataset u
                            const int i = n_prev + i0 + u;
                            const T nv = vs[i * k + j] + (bs[i * k + j] - part[u]) / (one_lam * a_t(i, j, j));
                            vs[i * k + j] = nv < eps ? eps : nv;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        for (int j = lane; j < k; j += 32) Wt[w * k + j] = ws[j];
        for (int i = n_prev; i < nall; ++i)
            for (int j = lane; j < k; j += 32) Vt[((size_t)i * g + w) * k + j] = vs[i * k + j];
        __syncwarp();
    }
}
// Incremental form of the same Gauss-Seidel sweeps (used when the statistics fit in shared memory and
// there are at most HALS_MAX_SETS current datasets).  Per gene the warp keeps, distributed over lanes
// (lane owns factors lane and lane+32),
//   D[j']   = sum_i ( B_i[j'] - sum_a (W_a + V_ia) A_i[a][j'] )        (all datasets; drives the W sweep)
//   Q_i[j'] = sum_a ( W_a + (1+lambda) V_ia ) A_i[a][j']               (current datasets; drives the V sweep)
// and applies rank-1 corrections with row j of A_i whenever W_j or V_ij changes, instead of recomputing a
// k-term dot product per (factor, dataset).  Same update order as _update_W_HALS / _update_V_HALS.
template <typename T>
struct HalsIncCfg {  // CTA shape of hals_inc_kernel: two CTAs per SM, as many warps as the register file allows
    static constexpr int MAX_THREADS = sizeof(T) == 4 ? 576 : 320;
    static constexpr int CTAS_PER_SM = 2;
};

// NS = compiled bound on the number of current datasets (register arrays).  Per-lane state during the sweeps: its two
// entries of W, D and of every Q_i; W_j, D_j are broadcast by shuffles, so the sweeps touch shared memory only for the
// rows of A_i they apply and run without warp barriers.  Reciprocals of the diagonals are tabulated per CTA.
template <typename T, int NS>
__global__ void __launch_bounds__(HalsIncCfg<T>::MAX_THREADS, HalsIncCfg<T>::CTAS_PER_SM)
hals_inc_kernel(const double* __restrict__ A, const T* __restrict__ B, T* __restrict__ Wt, T* __restrict__ Vt,
                double lambda, int n_prev, int nall, int64_t g, int k, int n_inner) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kk = k * k;
    T* As = reinterpret_cast<T*>(smem_raw);              // [nall][k][k]   A_i (row-major, as given)
    T* SAs = As + (size_t)nall * kk;                     // [k][k]         sum_i A_i
    T* inv_sa = SAs + kk;                                // [k]            1 / (sum_i A_i)[j][j]
    T* inv_a = inv_sa + k;                               // [nall][k]      1 / ((1 + lambda) A_i[j][j])
    const T one_lam = (T)(1.0 + lambda), lam = (T)lambda;
    for (int e = threadIdx.x; e < nall * kk; e += blockDim.x) As[e] = (T)A[e];
    for (int e = threadIdx.x; e < kk; e += blockDim.x) {
        double acc = 0.0;
        for (int i = 0; i < nall; ++i) acc += A[(size_t)i * kk + e];
        SAs[e] = (T)acc;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < k; e += blockDim.x) inv_sa[e] = T(1) / SAs[e * k + e];
    for (int e = threadIdx.x; e < nall * k; e += blockDim.x) {
        const int i = e / k, j = e - i * k;
        inv_a[e] = T(1) / (one_lam * As[(size_t)i * kk + j * k + j]);
    }
    __syncthreads();
    const size_t shared_elems = (((size_t)(nall + 1) * kk + (size_t)(nall + 1) * k) + 3) & ~(size_t)3;
    const int per_warp = (1 + 2 * nall) * k;
    T* ws = reinterpret_cast<T*>(smem_raw) + shared_elems + (size_t)warp * per_warp;
    T* vs = ws + k;                 // [nall][k]
    T* bs = vs + (size_t)nall * k;  // [nall][k]
    const T eps = T(1e-16);
    const int ncur = nall - n_prev;
    const int j0 = lane, j1 = lane + 32;
    const bool has0 = j0 < k, has1 = j1 < k;
    const int c0 = has0 ? j0 : 0, c1 = has1 ? j1 : 0;  // clamped column indices for loads
    const T* Acur = As + (size_t)n_prev * kk;
    for (int64_t w = (int64_t)blockIdx.x * nwarps + warp; w < g; w += (int64_t)gridDim.x * nwarps) {
        T W0 = has0 ? Wt[w * k + j0] : T(0), W1 = has1 ? Wt[w * k + j1] : T(0);
        if (has0) ws[j0] = W0;
        if (has1) ws[j1] = W1;
        for (int i = 0; i < nall; ++i)
            for (int j = lane; j < k; j += 32) {
                vs[i * k + j] = Vt[((size_t)i * g + w) * k + j];
                bs[i * k + j] = B[((size_t)i * g + w) * k + j];
            }
        __syncwarp();
        for (int it = 0; it < n_inner; ++it) {
            // ---- phase 0: D and Q_i from scratch ----
            T D0 = T(0), D1 = T(0);
            for (int i = 0; i < n_prev; ++i) {  // previous datasets only enter D
                T p0 = T(0), p1 = T(0);
                const T* Ai = As + (size_t)i * kk;
                for (int a = 0; a < k; ++a) {
                    const T u = ws[a] + vs[i * k + a];
                    p0 = fma(u, Ai[a * k + c0], p0);
                    p1 = fma(u, Ai[a * k + c1], p1);
                }
                D0 += bs[i * k + c0] - p0;
                D1 += bs[i * k + c1] - p1;
            }
            T Q0[NS], Q1[NS];
            {
                T P0[NS], P1[NS], VA0[NS], VA1[NS];
#pragma unroll
                for (int u = 0; u < NS; ++u) P0[u] = P1[u] = VA0[u] = VA1[u] = T(0);
                for (int a = 0; a < k; ++a) {
                    const T wa = ws[a];
                    const T* Ar = Acur + a * k;
#pragma unroll
                    for (int u = 0; u < NS; ++u) {
                        if (u < ncur) {
                            const T va = vs[(n_prev + u) * k + a];
                            const T a0 = Ar[(size_t)u * kk + c0], a1 = Ar[(size_t)u * kk + c1];
                            P0[u] = fma(wa + va, a0, P0[u]);
                            P1[u] = fma(wa + va, a1, P1[u]);
                            VA0[u] = fma(va, a0, VA0[u]);
                            VA1[u] = fma(va, a1, VA1[u]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < NS; ++u) {
                    if (u < ncur) {
                        D0 += bs[(n_prev + u) * k + c0] - P0[u];
                        D1 += bs[(n_prev + u) * k + c1] - P1[u];
                    }
                    Q0[u] = fma(lam, VA0[u], P0[u]);  // Q_i = P_i + lambda V_i A_i; the W sweep corrects the P_i part
                    Q1[u] = fma(lam, VA1[u], P1[u]);
                }
            }
            // ---- W sweep ----
            for (int j = 0; j < k; ++j) {
                const int src = j & 31;
                const T d = __shfl_sync(INMF_FULL_MASK, (j < 32) ? D0 : D1, src);
                const T wj = __shfl_sync(INMF_FULL_MASK, (j < 32) ? W0 : W1, src);
                T nw = fma(d, inv_sa[j], wj);
                nw = nw < eps ? eps : nw;
                const T delta = nw - wj;
                if (lane == src) {
                    if (j < 32) W0 = nw; else W1 = nw;
                }
                D0 = fma(-delta, SAs[j * k + c0], D0);
                D1 = fma(-delta, SAs[j * k + c1], D1);
                const T* Ar = Acur + j * k;
#pragma unroll
                for (int u = 0; u < NS; ++u) {
                    if (u < ncur) {
                        Q0[u] = fma(delta, Ar[(size_t)u * kk + c0], Q0[u]);
                        Q1[u] = fma(delta, Ar[(size_t)u * kk + c1], Q1[u]);
                    }
                }
            }
            // ---- V sweep: lane u finishes dataset u, all lanes apply the rank-1 corrections ----
            for (int j = 0; j < k; ++j) {
                T q = T(0);
#pragma unroll
                for (int u = 0; u < NS; ++u) {
                    if (u < ncur) {
                        const T t = __shfl_sync(INMF_FULL_MASK, (j < 32) ? Q0[u] : Q1[u], j & 31);
                        q = (lane == u) ? t : q;
                    }
                }
                T delta = T(0);
                if (lane < ncur) {
                    const int i = n_prev + lane;
                    const T vij = vs[i * k + j];
                    T nv = fma(bs[i * k + j] - q, inv_a[i * k + j], vij);
                    nv = nv < eps ? eps : nv;
                    delta = one_lam * (nv - vij);
                    vs[i * k + j] = nv;
                }
                const T* Ar = Acur + j * k;
#pragma unroll
                for (int u = 0; u < NS; ++u) {
                    if (u < ncur) {
                        const T du = __shfl_sync(INMF_FULL_MASK, delta, u);
                        Q0[u] = fma(du, Ar[(size_t)u * kk + c0], Q0[u]);
                        Q1[u] = fma(du, Ar[(size_t)u * kk + c1], Q1[u]);
                    }
                }
            }
            if (it + 1 < n_inner) {  // the next inner iteration rebuilds D, Q from the stored W, V
                if (has0) ws[j0] = W0;
                if (has1) ws[j1] = W1;
            }
            __syncwarp();
        }
        if (has0) Wt[w * k + j0] = W0;
        if (has1) Wt[w * k + j1] = W1;
        for (int i = n_prev; i < nall; ++i)
            for (int j = lane; j < k; j += 32) Vt[((size_t)i * g + w) * k + j] = vs[i * k + j];
        __syncwarp();
    }
}

template <typename T, int NS>
static int32_t hals_inc_launch(const double* A, const T* B, T* Wt, T* Vt, double lambda, int32_t n_prev, int32_t nall,
                               int64_t g, int32_t k, int32_t n_inner, int nw, size_t smem, void* stream) {
    INMF_OPTIN_SMEM((hals_inc_kernel<T, NS>), 227 * 1024);
    // one gene per warp and round; the warps per CTA are chosen so that the last round is as full as the others
    const int64_t ctas = (int64_t)inmf_num_sms() * HalsIncCfg<T>::CTAS_PER_SM;
    const int64_t rounds = (g + ctas * nw - 1) / (ctas * nw);
    int64_t w_needed = (g + rounds * ctas - 1) / (rounds * ctas);  // warps per CTA that cover g in `rounds` rounds
    if (w_needed < 1) w_needed = 1;
    if (w_needed < nw) nw = (int)w_needed;
    int64_t blocks = (g + nw - 1) / nw;
    if (blocks > ctas) blocks = ctas;
    hals_inc_kernel<T, NS><<<(unsigned)blocks, nw * 32, smem, (cudaStream_t)stream>>>(A, B, Wt, Vt, lambda, n_prev, nall,
                                                                                    g, k, n_inner);
    INMF_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Grouped form of the incremental sweeps: the warp's lanes are split into NS groups of LPD = 32 / NS lanes, one group
// per current dataset; lane `sub` of a group owns the factors c = s * LPD + sub (slot s = 0 .. SLOTS-1) of
//   D[c]   = sum_i ( B_i[c] - sum_a (W_a + V_ia) A_i[a][c] )          (held by every group; drives the W sweep)
//   R_d[c] = B_d[c] - sum_a ( W_a + (1+lambda) V_da ) A_d[a][c]        (its own dataset d; drives the V sweep)
// so a step of either sweep is one in-group broadcast, a handful of scalar operations and SLOTS fused multiply-adds
// against a row of A_d that the lane reads as 16-byte vectors (the rows are stored permuted, [sub][slot], with the
// factor dimension padded to kp = SLOTS * LPD by zeros).  No work is spent on missing factors beyond that padding and
// none on "the other 32 columns" as in the column-per-lane kernels; 5.5x fewer instructions per gene at 8 datasets, k = 40.
// ------------------------------------------------------------------------------------------------
template <typename T, int SLOTS>
__device__ __forceinline__ void hals_ld_row(const T* __restrict__ p, T (&v)[SLOTS]) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int q = 0; q < SLOTS / 4; ++q) {
            const float4 t = reinterpret_cast<const float4*>(p)[q];
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < SLOTS / 2; ++q) {
            const double2 t = reinterpret_cast<const double2*>(p)[q];
            v[2 * q] = t.x; v[2 * q + 1] = t.y;
        }
    }
}

template <typename T, int SLOTS>
struct HalsGroupedCfg {  // threads per CTA that leave ~5 register arrays of SLOTS values per lane
    static constexpr int WORDS = SLOTS * (int)(sizeof(T) / 4);
    static constexpr int MAX_THREADS = WORDS <= 8 ? 1024 : WORDS <= 12 ? 896 : WORDS <= 16 ? 576
                                       : WORDS <= 24 ? 448 : 320;
    static constexpr int CTAS_PER_SM = 1;  // one large CTA: the matrices are staged once per SM
};

template <typename T, int NS, int SLOTS>
__global__ void __launch_bounds__((HalsGroupedCfg<T, SLOTS>::MAX_THREADS), (HalsGroupedCfg<T, SLOTS>::CTAS_PER_SM))
hals_grouped_kernel(const double* __restrict__ A, const T* __restrict__ B, T* __restrict__ Wt, T* __restrict__ Vt,
                    double lambda, int n_prev, int nall, int64_t g, int k, int n_inner, int stride_a) {
    constexpr int LPD = 32 / NS;
    constexpr int KP = SLOTS * LPD;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsets = n_prev + NS;                       // datasets with a (possibly all-zero) matrix in shared memory
    T* Ap = reinterpret_cast<T*>(smem_raw);              // [nsets][stride_a]: rows j < k of A_i, permuted + padded
    T* SAp = Ap + (size_t)nsets * stride_a;              // [k][KP]           sum_i A_i, same row layout
    T* inv_sa = SAp + (size_t)k * KP;                    // [k]               1 / (sum_i A_i)[j][j]
    T* inv_a = inv_sa + k;                               // [nsets][k]        1 / ((1 + lambda) A_i[j][j]); 0 for pad sets
    const T one_lam = (T)(1.0 + lambda);
    const int kk = k * k;
    for (int row = warp; row < nsets * k; row += nwarps) {  // rows of the A_i (zero rows for the pad datasets)
        const int i = row / k, j = row - i * k;
        for (int pos = lane; pos < KP; pos += 32) {
            const int c = (pos % SLOTS) * LPD + pos / SLOTS;  // pos = sub * SLOTS + slot
            Ap[(size_t)i * stride_a + j * KP + pos] = (i < nall && c < k) ? (T)A[(size_t)i * kk + j * k + c] : T(0);
        }
    }
    for (int j = warp; j < k; j += nwarps)
        for (int pos = lane; pos < KP; pos += 32) {
            const int c = (pos % SLOTS) * LPD + pos / SLOTS;
            double acc = 0.0;
            if (c < k)
                for (int i = 0; i < nall; ++i) acc += A[(size_t)i * kk + j * k + c];
            SAp[j * KP + pos] = (T)acc;
        }
    for (int e = threadIdx.x; e < k; e += blockDim.x) {
        double acc = 0.0;
        for (int i = 0; i < nall; ++i) acc += A[(size_t)i * kk + e * k + e];
        inv_sa[e] = T(1) / (T)acc;
    }
    for (int e = threadIdx.x; e < nsets * k; e += blockDim.x) {
        const int i = e / k, j = e - i * k;
        inv_a[e] = (i < nall) ? T(1) / (one_lam * (T)A[(size_t)i * kk + j * k + j]) : T(0);
    }
    const size_t shared_elems = ((size_t)nsets * stride_a + (size_t)k * KP + (size_t)(nsets + 1) * k + 3) & ~(size_t)3;
    const int per_warp = (1 + nsets) * k;
    T* ws = reinterpret_cast<T*>(smem_raw) + shared_elems + (size_t)warp * per_warp;  // [k]
    T* vs = ws + k;                                                                   // [nsets][k], pad sets stay zero
    for (int e = lane; e < per_warp; e += 32) ws[e] = T(0);
    __syncthreads();

    const T eps = T(1e-16);
    const int ncur = nall - n_prev;
    const int d = lane / LPD, sub = lane % LPD;  // my dataset (n_prev + d) and my lane in its group
    const bool live = d < ncur;
    const T* Ad = Ap + (size_t)(n_prev + d) * stride_a + sub * SLOTS;  // + j * KP: my part of row j of A_d
    const T* SAd = SAp + sub * SLOTS;
    T* vd = vs + (size_t)(n_prev + d) * k;
    const T* inv_ad = inv_a + (size_t)(n_prev + d) * k;
    for (int64_t w = (int64_t)blockIdx.x * nwarps + warp; w < g; w += (int64_t)gridDim.x * nwarps) {
        for (int j = lane; j < k; j += 32) ws[j] = Wt[w * k + j];
        for (int i = 0; i < nall; ++i)
            for (int j = lane; j < k; j += 32) vs[i * k + j] = Vt[((size_t)i * g + w) * k + j];
        __syncwarp();
        for (int it = 0; it < n_inner; ++it) {
            // ---- phase 0: D and R_d from scratch ----
            T D[SLOTS], R[SLOTS];
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) D[s] = T(0);
            for (int i = d; i < n_prev; i += NS) {  // previous datasets only enter D: spread over the groups
                T P[SLOTS], a[SLOTS];
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) P[s] = T(0);
                const T* Ai = Ap + (size_t)i * stride_a + sub * SLOTS;
                for (int r = 0; r < k; ++r) {
                    const T u = ws[r] + vs[i * k + r];
                    hals_ld_row<T, SLOTS>(Ai + r * KP, a);
#pragma unroll
                    for (int s = 0; s < SLOTS; ++s) P[s] = fma(u, a[s], P[s]);
                }
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int c = s * LPD + sub;
                    D[s] += ((c < k) ? B[((size_t)i * g + w) * k + c] : T(0)) - P[s];
                }
            }
            {
                T WA[SLOTS], VA[SLOTS], a[SLOTS];
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {  // starts from -B_d: the sums below need B_d - WA - ... only
                    const int c = s * LPD + sub;
                    WA[s] = (live && c < k) ? -B[((size_t)(n_prev + d) * g + w) * k + c] : T(0);
                    VA[s] = T(0);
                }
                for (int r = 0; r < k; ++r) {
                    const T wa = ws[r], va = vd[r];
                    hals_ld_row<T, SLOTS>(Ad + r * KP, a);
#pragma unroll
                    for (int s = 0; s < SLOTS; ++s) {
                        WA[s] = fma(wa, a[s], WA[s]);
                        VA[s] = fma(va, a[s], VA[s]);
                    }
                }
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    D[s] -= WA[s] + VA[s];
                    R[s] = -fma(one_lam, VA[s], WA[s]);
                }
            }
            if (NS > 1) {  // D summed over the groups (every group ends with the full sum)
#pragma unroll
                for (int o = LPD; o < 32; o <<= 1) {
#pragma unroll
                    for (int s = 0; s < SLOTS; ++s) D[s] += __shfl_xor_sync(INMF_FULL_MASK, D[s], o);
                }
            }
            // ---- W sweep ----
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                if (s * LPD < k) {
                    const int nsub = (k - s * LPD < LPD) ? (k - s * LPD) : LPD;
                    for (int sj = 0; sj < nsub; ++sj) {
                        const int j = s * LPD + sj;
                        const T dj = __shfl_sync(INMF_FULL_MASK, D[s], sj, LPD);
                        const T wj = ws[j];
                        T nw = fma(dj, inv_sa[j], wj);
                        nw = nw < eps ? eps : nw;
                        const T delta = nw - wj;
                        if (lane == 0) ws[j] = nw;
                        T a[SLOTS];
                        hals_ld_row<T, SLOTS>(SAd + j * KP, a);
#pragma unroll
                        for (int t = 0; t < SLOTS; ++t) D[t] = fma(-delta, a[t], D[t]);
                        hals_ld_row<T, SLOTS>(Ad + j * KP, a);
#pragma unroll
                        for (int t = 0; t < SLOTS; ++t) R[t] = fma(-delta, a[t], R[t]);
                    }
                }
            }
            // ---- V sweep: every group sweeps its own dataset ----
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                if (s * LPD < k) {
                    const int nsub = (k - s * LPD < LPD) ? (k - s * LPD) : LPD;
                    for (int sj = 0; sj < nsub; ++sj) {
                        const int j = s * LPD + sj;
                        const T rj = __shfl_sync(INMF_FULL_MASK, R[s], sj, LPD);
                        const T vij = vd[j];
                        T nv = fma(rj, inv_ad[j], vij);
                        nv = nv < eps ? eps : nv;
                        const T delta = one_lam * (nv - vij);
                        if (sub == 0 && live) vd[j] = nv;
                        T a[SLOTS];
                        hals_ld_row<T, SLOTS>(Ad + j * KP, a);
#pragma unroll
                        for (int t = 0; t < SLOTS; ++t) R[t] = fma(-delta, a[t], R[t]);
                    }
                }
            }
            __syncwarp();
        }
        for (int j = lane; j < k; j += 32) Wt[w * k + j] = ws[j];
        for (int i = n_prev; i < nall; ++i)
            for (int j = lane; j < k; j += 32) Vt[((size_t)i * g + w) * k + j] = vs[i * k + j];
        __syncwarp();
    }
}

// Stride (in elements) between the datasets' matrices in shared memory: >= k * KP, a multiple of 4, chosen so that the
// 16-byte row reads of the groups of a quarter warp fall into different bank groups where possible.
static int hals_grouped_stride(int k, int KP, int SLOTS, int LPD, int esize) {
    const int base = (k * KP + 3) & ~3;
    int best = base, best_cost = 1 << 30;
    const int per_phase = 128 / 16;  // lanes served per wavefront of a 16-byte access
    for (int pad = 0; pad < 64; pad += 4) {
        const int stride = base + pad;
        int cost = 0;
        for (int q = 0; q < SLOTS * esize / 16; ++q)
            for (int l0 = 0; l0 < 32; l0 += per_phase) {
                int seen[8] = {0, 0, 0, 0, 0, 0, 0, 0}, worst = 0;
                long long addr_seen[8][8];
                for (int l = l0; l < l0 + per_phase; ++l) {
                    const long long byte = ((long long)(l / LPD) * stride + (long long)(l % LPD) * SLOTS) * esize + 16 * q;
                    const int bank = (int)((byte / 16) % 8);
                    bool dup = false;
                    for (int t = 0; t < seen[bank]; ++t) dup = dup || addr_seen[bank][t] == byte;
                    if (!dup) addr_seen[bank][seen[bank]++] = byte;
                    if (seen[bank] > worst) worst = seen[bank];
                }
                cost += worst;
            }
        if (cost < best_cost) {
            best_cost = cost;
            best = stride;
        }
    }
    return best;
}

template <typename T, int NS, int SLOTS>
static int32_t hals_grouped_launch(const double* A, const T* B, T* Wt, T* Vt, double lambda, int32_t n_prev,
                                   int32_t nall, int64_t g, int32_t k, int32_t n_inner, void* stream, bool* done) {
    constexpr int LPD = 32 / NS, KP = SLOTS * LPD;
    const int nsets = n_prev + NS;
    const int stride_a = hals_grouped_stride(k, KP, SLOTS, LPD, (int)sizeof(T));
    const size_t shared_elems = ((size_t)nsets * stride_a + (size_t)k * KP + (size_t)(nsets + 1) * k + 3) & ~(size_t)3;
    const size_t per_warp = (size_t)(1 + nsets) * k * sizeof(T);
    typedef HalsGroupedCfg<T, SLOTS> Cfg;
    const size_t cta_budget = (size_t)(224 * 1024) / Cfg::CTAS_PER_SM;
    *done = false;
    if (shared_elems * sizeof(T) + per_warp > cta_budget) return 0;
    int nw = Cfg::MAX_THREADS / 32;
    while (nw > 1 && shared_elems * sizeof(T) + nw * per_warp > cta_budget) --nw;
    const size_t smem = shared_elems * sizeof(T) + nw * per_warp;
    INMF_OPTIN_SMEM((hals_grouped_kernel<T, NS, SLOTS>), 227 * 1024);
    // one gene per warp and round; the warps per CTA are chosen so that the last round is as full as the others
    const int64_t ctas = (int64_t)inmf_num_sms() * Cfg::CTAS_PER_SM;
    const int64_t rounds = (g + ctas * nw - 1) / (ctas * nw);
    int64_t w_needed = (g + rounds * ctas - 1) / (rounds * ctas);
    if (w_needed < 1) w_needed = 1;
    if (w_needed < nw) nw = (int)w_needed;
    int64_t blocks = (g + nw - 1) / nw;
    if (blocks > ctas) blocks = ctas;
    hals_grouped_kernel<T, NS, SLOTS><<<(unsigned)blocks, nw * 32, smem, (cudaStream_t)stream>>>(
        A, B, Wt, Vt, lambda, n_prev, nall, g, k, n_inner, stride_a);
    INMF_LAUNCH_CHECK();
    *done = true;
    return 0;
}

template <typename T, int NS>
static int32_t hals_grouped_slots(const double* A, const T* B, T* Wt, T* Vt, double lambda, int32_t n_prev,
                                  int32_t nall, int64_t g, int32_t k, int32_t n_inner, void* stream, bool* done) {
    constexpr int LPD = 32 / NS;
    const int slots = (k + LPD - 1) / LPD;
    *done = false;
    if (slots <= 4) return hals_grouped_launch<T, NS, 4>(A, B, Wt, Vt, lambda, n_prev, nall, g, k, n_inner, stream, done);
    if (slots <= 8) return hals_grouped_launch<T, NS, 8>(A, B, Wt, Vt, lambda, n_prev, nall, g, k, n_inner, stream, done);
    if (slots <= 12) return hals_grouped_launch<T, NS, 12>(A, B, Wt, Vt, lambda, n_prev, nall, g, k, n_inner, stream, done);
    if (slots <= 16) return hals_grouped_launch<T, NS, 16>(A, B, Wt, Vt, lambda, n_prev, nall, g, k, n_inner, stream, done);
    return 0;
}

template <typename T>
static int32_t hals_launch(const double* A, const T* B, T* Wt, T* Vt, double lambda, int32_t n_prev, int32_t nall,
                           int64_t g, int32_t k, int32_t n_inner, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K && nall >= 1 && g >= 1, "bad shape");
    INMF_REQUIRE(n_prev >= 0 && n_prev <= nall && n_inner >= 0, "bad n_prev / n_inner");
    const size_t per_warp = (size_t)(1 + 2 * nall) * k * sizeof(T);
    {   // grouped kernel: up to 8 current datasets, one lane group each
        const int ncur = nall - n_prev;
        bool done = false;
        int32_t rc = 0;
        if (ncur <= 1) rc = hals_grouped_slots<T, 1>(A, B, Wt, Vt, lambda, n_prev, nall, g, k, n_inner, stream, &done);
        else if (ncur <= 2) rc = hals_grouped_slots<T, 2>(A, B, Wt, Vt, lambda, n_prev, nall, g, k, n_inner, stream, &done);
        else if (ncur <= 4) rc = hals_grouped_slots<T, 4>(A, B, Wt, Vt, lambda, n_prev, nall, g, k, n_inner, stream, &done);
        else if (ncur <= 8) rc = hals_grouped_slots<T, 8>(A, B, Wt, Vt, lambda, n_prev, nall, g, k, n_inner, stream, &done);
        if (rc != 0 || done) return rc;
    }
    {   // incremental kernel: A_i, their sum and the reciprocal diagonals resident in shared memory
        const size_t shared_bytes = ((((size_t)(nall + 1) * k * k + (size_t)(nall + 1) * k) + 3) & ~(size_t)3) * sizeof(T);
        const int ncur = nall - n_prev;
        const size_t cta_budget = (size_t)(220 * 1024) / HalsIncCfg<T>::CTAS_PER_SM;
        if (ncur <= HALS_MAX_SETS && shared_bytes + per_warp <= cta_budget) {
            int nw = HalsIncCfg<T>::MAX_THREADS / 32;
            while (nw > 1 && shared_bytes + nw * per_warp > cta_budget) --nw;
            const size_t smem_inc = shared_bytes + nw * per_warp;
            if (ncur <= 4)
                return hals_inc_launch<T, 4>(A, B, Wt, Vt, lambda, n_prev, nall, g, k, n_inner, nw, smem_inc, stream);
            if (ncur <= 8)
                return hals_inc_launch<T, 8>(A, B, Wt, Vt, lambda, n_prev, nall, g, k, n_inner, nw, smem_inc, stream);
            return hals_inc_launch<T, HALS_MAX_SETS>(A, B, Wt, Vt, lambda, n_prev, nall, g, k, n_inner, nw, smem_inc,
                                                     stream);
        }
    }
    const size_t a_bytes = ((((size_t)nall * k * k) + 3) & ~(size_t)3) * sizeof(T);
    const bool a_smem = a_bytes <= 112 * 1024;
    const size_t budget = 220 * 1024 - (a_smem ? a_bytes : 0);
    int nwarps = 8;
    while (nwarps > 1 && nwarps * per_warp > budget) nwarps >>= 1;
    const size_t smem = (a_smem ? a_bytes : 0) + nwarps * per_warp;
    INMF_REQUIRE(smem <= 227 * 1024, "too many datasets for the HALS shared-memory staging");
    INMF_OPTIN_SMEM((hals_kernel<T, true>), 227 * 1024);
    INMF_OPTIN_SMEM((hals_kernel<T, false>), 227 * 1024);
    int64_t blocks = (g + nwarps - 1) / nwarps;
    if (blocks > inmf_num_sms() * 8) blocks = inmf_num_sms() * 8;
    if (a_smem)
        hals_kernel<T, true><<<(unsigned)blocks, nwarps * 32, smem, (cudaStream_t)stream>>>(A, B, Wt, Vt, lambda, n_prev,
                                                                                          nall, g, k, n_inner);
    else
        hals_kernel<T, false><<<(unsigned)blocks, nwarps * 32, smem, (cudaStream_t)stream>>>(A, B, Wt, Vt, lambda, n_prev,
                                                                                           nall, g, k, n_inner);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_hals_f32(const double* A, const float* B, float* Wt, float* Vt, double lambda,
                                 int32_t n_prev, int32_t nall, int64_t g, int32_t k, int32_t n_inner, void* stream) {
    return hals_launch<float>(A, B, Wt, Vt, lambda, n_prev, nall, g, k, n_inner, stream);
}
extern "C" int32_t inmf_hals_f64(const double* A, const double* B, double* Wt, double* Vt, double lambda,
                                 int32_t n_prev, int32_t nall, int64_t g, int32_t k, int32_t n_inner, void* stream) {
    return hals_launch<double>(A, B, Wt, Vt, lambda, n_prev, nall, g, k, n_inner, stream);
}

// =================================================================================================
// Dense normal equations for the raw-input operator form
// =================================================================================================
template <typename T>
__global__ void gram_dense_kernel(const T* __restrict__ A, int64_t m, int n, double* __restrict__ G) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    const int ldt = ((n + 3) & ~3) + 4;
    for (int64_t r0 = (int64_t)blockIdx.x * GRAM_TILE_ROWS; r0 < m; r0 += (int64_t)gridDim.x * GRAM_TILE_ROWS) {
        const int rows = (int)((m - r0 < GRAM_TILE_ROWS) ? (m - r0) : GRAM_TILE_ROWS);
        for (int t = threadIdx.x; t < GRAM_TILE_ROWS * ldt; t += blockDim.x) {
            const int r = t / ldt, j = t % ldt;
            tile[t] = (r < rows && j < n) ? A[(r0 + r) * n + j] : T(0);
        }
        __syncthreads();
        tile_gram_accumulate<T>(tile, ldt, rows, n, G, 1.0);
        __syncthreads();
    }
}
// Rt[c][j] = sum_r B[r][c] * A[r][j]; block = 64 (j) x 4 (c)
template <typename T>
__global__ void atb_dense_kernel(const T* __restrict__ A, const T* __restrict__ B, int64_t m, int n, int64_t cols,
                                 T* __restrict__ Rt) {
    const int j = threadIdx.x;
    const int64_t c = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (j >= n || c >= cols) return;
    double acc = 0.0;
    for (int64_t r = 0; r < m; ++r) acc += (double)B[r * cols + c] * (double)A[r * n + j];
    Rt[c * n + j] = (T)acc;
}
template <typename T>
static int32_t normal_eq_launch(const T* A, const T* B, int64_t m, int32_t n, int64_t cols, double* G, T* Rt,
                                void* stream) {
    INMF_REQUIRE(n >= 1 && n <= INMF_MAX_K && m >= 1 && cols >= 0, "bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    INMF_CUDA_CHECK(cudaMemsetAsync(G, 0, (size_t)n * n * sizeof(double), st));
    int64_t tiles = (m + GRAM_TILE_ROWS - 1) / GRAM_TILE_ROWS;
    if (tiles > inmf_num_sms() * 4) tiles = inmf_num_sms() * 4;
    const int ldt = ((n + 3) & ~3) + 4;
    gram_dense_kernel<T><<<(unsigned)tiles, 256, (size_t)GRAM_TILE_ROWS * ldt * sizeof(T), st>>>(A, m, n, G);
    INMF_LAUNCH_CHECK();
    if (cols > 0) {
        atb_dense_kernel<T><<<(unsigned)((cols + 3) / 4), dim3(64, 4), 0, st>>>(A, B, m, n, cols, Rt);
        INMF_LAUNCH_CHECK();
    }
    return 0;
}
extern "C" int32_t inmf_normal_eq_f32(const float* A, const float* B, int64_t m, int32_t n, int64_t cols, double* G,
                                      float* Rt, void* stream) {
    return normal_eq_launch<float>(A, B, m, n, cols, G, Rt, stream);
}
extern "C" int32_t inmf_normal_eq_f64(const double* A, const double* B, int64_t m, int32_t n, int64_t cols,
                                      double* G, double* Rt, void* stream) {
    return normal_eq_launch<double>(A, B, m, n, cols, G, Rt, stream);
}

// =================================================================================================
// HALS update of H (iNMF_HALS): one Gauss-Seidel sweep over the k factors of every cell
// =================================================================================================
// Reference: _update_H_HALS (_utilities.py:132-148): for j = 0..k-1, in place,
//   H[j,:] = nonneg(H[j,:] + ((W+V)[:,j]' X - (W+V)[:,j]'(W+V) H - lambda V'V[j,:] H) / ((W+V)'(W+V)[j,j] + lambda V'V[j,j]))
// i.e. per cell c, with r = X[c,:](W+V) (pass 1) and G = (W+V)'(W+V) + lambda V'V (prep):
//   h_j <- max(1e-16, h_j + (r_j - G[j,:] h) / G[j,j]),   h updated in place.
// One THREAD per cell (cells are independent, the sweep over j is sequential): G is read as shared-memory broadcasts,
// the cell's h lives in a [k][threads] shared array (conflict free), 2 k^2 / 32 warp instructions per cell.
template <typename T>
__global__ void __launch_bounds__(128) hals_h_kernel(const double* __restrict__ G, const T* __restrict__ R, int64_t ldr,
                                                     T* __restrict__ H, int64_t ldh,
                                                     const int64_t* __restrict__ seg_off, int k) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    T* Gs = reinterpret_cast<T*>(smem_raw);
    T* hs = Gs + (((size_t)k * k + 3) & ~(size_t)3);
    const int seg = blockIdx.y;
    const int64_t c_begin = seg_off[seg], ncols = seg_off[seg + 1] - c_begin;
    if ((int64_t)blockIdx.x * blockDim.x >= ncols) return;
    const double* Gseg = G + (size_t)seg * k * k;
    for (int e = threadIdx.x; e < k * k; e += blockDim.x) Gs[e] = (T)Gseg[e];
    __syncthreads();
    const int nt = blockDim.x, tid = threadIdx.x;
    for (int64_t c = (int64_t)blockIdx.x * nt + tid; c < ncols; c += (int64_t)gridDim.x * nt) {
        T* h = H + (c_begin + c) * ldh;
        const T* r = R + (c_begin + c) * ldr;
        for (int l = 0; l < k; ++l) hs[l * nt + tid] = h[l];
        for (int j = 0; j < k; ++j) {
            T s = r[j];
            const T* g = Gs + j * k;
            for (int l = 0; l < k; ++l) s = fma(-g[l], hs[l * nt + tid], s);
            T hj = hs[j * nt + tid] + s / g[j];
            hj = hj < T(1e-16) ? T(1e-16) : hj;  // nonneg(): x[x < eps] = eps (_utilities.py:10-13)
            hs[j * nt + tid] = hj;
        }
        for (int l = 0; l < k; ++l) h[l] = hs[l * nt + tid];
    }
}
template <typename T>
static int32_t hals_h_launch(const double* G, const T* R, int64_t ldr, T* H, int64_t ldh, const int64_t* seg_off,
                             int32_t nseg, int64_t max_seg_cols, int32_t k, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K, "k must be in [1, 64]");
    INMF_REQUIRE(nseg >= 1 && nseg <= 65535 && ldr >= k && ldh >= k, "bad shape");
    if (max_seg_cols <= 0) return 0;
    const size_t smem = ((((size_t)k * k + 3) & ~(size_t)3) + (size_t)k * 128) * sizeof(T);
    INMF_OPTIN_SMEM((hals_h_kernel<T>), 100 * 1024);
    int64_t blocks = (max_seg_cols + 127) / 128;
    const int64_t cap = (int64_t)inmf_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    hals_h_kernel<T><<<dim3((unsigned)blocks, (unsigned)nseg), 128, smem, (cudaStream_t)stream>>>(G, R, ldr, H, ldh,
                                                                                                 seg_off, k);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_hals_h_f32(const double* G, const float* R, int64_t ldr, float* H, int64_t ldh,
                                   const int64_t* seg_off, int32_t nseg, int64_t max_seg_cols, int32_t k, void* stream) {
    return hals_h_launch<float>(G, R, ldr, H, ldh, seg_off, nseg, max_seg_cols, k, stream);
}
extern "C" int32_t inmf_hals_h_f64(const double* G, const double* R, int64_t ldr, double* H, int64_t ldh,
                                   const int64_t* seg_off, int32_t nseg, int64_t max_seg_cols, int32_t k, void* stream) {
    return hals_h_launch<double>(G, R, ldr, H, ldh, seg_off, nseg, max_seg_cols, k, stream);
}

// =================================================================================================
// v2 hot kernels: blocked sparse x shared dense tile (blocked.cuh)
// =================================================================================================
template <typename T, int LPN_T>
static int32_t blocked_launch_one(const int64_t* work, int32_t nwork, const int64_t* ptr, const void* pairs,
                                  const T* dense, int32_t kp, T* out, int64_t ldo, int32_t k, size_t smem,
                                  double* gram_out, void* stream) {
    INMF_OPTIN_SMEM((blocked_spmm_kernel<T, LPN_T>), 227 * 1024);
    blocked_spmm_kernel<T, LPN_T><<<(unsigned)nwork, 1024, smem, (cudaStream_t)stream>>>(
        work, ptr, reinterpret_cast<const typename PairTraits<T>::Pair*>(pairs), dense, kp, out, ldo, k, gram_out);
    INMF_LAUNCH_CHECK();
    return 0;
}

template <typename T>
static int32_t blocked_launch(const int64_t* work, int32_t nwork, const int64_t* ptr, const void* pairs, const T* dense,
                              int32_t kp, T* out, int64_t ldo, int32_t k, int32_t max_tile_rows, double* gram_out,
                              void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K, "k must be in [1, 64]");
    INMF_REQUIRE(kp >= k && kp % 4 == 0 && kp <= 64, "kp must be a multiple of 4 in [k, 64]");
    INMF_REQUIRE(ldo >= k, "ldo must be >= k");
    if (nwork <= 0) return 0;
    const size_t smem = 128 + (size_t)max_tile_rows * kp * sizeof(T);
    INMF_REQUIRE(smem <= 227 * 1024, "dense tile does not fit in shared memory");
    const int lpn = kp / PairTraits<T>::VEC;
#define INMF_BLK_CASE(L)                                                                                          \
    case L:                                                                                                       \
        return blocked_launch_one<T, L>(work, nwork, ptr, pairs, dense, kp, out, ldo, k, smem, gram_out, stream)
    switch (lpn) {  // lanes per non-zero fixed at compile time for the common factor counts
        INMF_BLK_CASE(5);   // f32 k <= 20
        INMF_BLK_CASE(8);   // f32 k <= 32
        INMF_BLK_CASE(10);  // f32 k <= 40, f64 k <= 20
        INMF_BLK_CASE(15);  // f32 k <= 60
        INMF_BLK_CASE(16);  // f32 k <= 64, f64 k <= 32
        INMF_BLK_CASE(20);  // f64 k <= 40
        INMF_BLK_CASE(30);  // f64 k <= 60
        INMF_BLK_CASE(32);  // f64 k <= 64
        default:
            return blocked_launch_one<T, 0>(work, nwork, ptr, pairs, dense, kp, out, ldo, k, smem, gram_out, stream);
    }
#undef INMF_BLK_CASE
}
extern "C" int32_t inmf_blocked_spmm_f32(const int64_t* work, int32_t nwork, const int64_t* ptr, const void* pairs,
                                         const float* dense, int32_t kp, float* out, int64_t ldo, int32_t k,
                                         int32_t max_tile_rows, double* gram_out, void* stream) {
    return blocked_launch<float>(work, nwork, ptr, pairs, dense, kp, out, ldo, k, max_tile_rows, gram_out, stream);
}
extern "C" int32_t inmf_blocked_spmm_f64(const int64_t* work, int32_t nwork, const int64_t* ptr, const void* pairs,
                                         const double* dense, int32_t kp, double* out, int64_t ldo, int32_t k,
                                         int32_t max_tile_rows, double* gram_out, void* stream) {
    return blocked_launch<double>(work, nwork, ptr, pairs, dense, kp, out, ldo, k, max_tile_rows, gram_out, stream);
}

template <int RL>
static int32_t blocked_split_one(const int64_t* work, int32_t nwork, const int64_t* ptr, const void* pairs,
                                 const float* dense, int32_t kp, float* out, int64_t ldo, int32_t k,
                                 int32_t max_tile_rows, double* gram_out, void* stream) {
    const size_t smem = 128 + (size_t)max_tile_rows * (32 + 8 * RL) * sizeof(float);
    INMF_REQUIRE(smem <= 227 * 1024, "dense tile does not fit in shared memory");
    INMF_OPTIN_SMEM((blocked_split_kernel<RL>), 227 * 1024);
    blocked_split_kernel<RL><<<(unsigned)nwork, 1024, smem, (cudaStream_t)stream>>>(
        work, ptr, reinterpret_cast<const int2*>(pairs), dense, kp, out, ldo, k, max_tile_rows, gram_out);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_blocked_split_f32(const int64_t* work, int32_t nwork, const int64_t* ptr, const void* pairs,
                                          const float* dense, int32_t kp, float* out, int64_t ldo, int32_t k,
                                          int32_t max_tile_rows, double* gram_out, void* stream) {
    INMF_REQUIRE(k > 32 && k <= INMF_MAX_K, "the split-tile kernel is for 32 < k <= 64");
    INMF_REQUIRE(kp >= k && kp % 4 == 0 && kp <= 64, "kp must be a multiple of 4 in [k, 64]");
    INMF_REQUIRE(ldo >= k, "ldo must be >= k");
    if (nwork <= 0) return 0;
    const int rem = kp - 32;
    if (rem <= 8) return blocked_split_one<1>(work, nwork, ptr, pairs, dense, kp, out, ldo, k, max_tile_rows, gram_out, stream);
    if (rem <= 16) return blocked_split_one<2>(work, nwork, ptr, pairs, dense, kp, out, ldo, k, max_tile_rows, gram_out, stream);
    return blocked_split_one<4>(work, nwork, ptr, pairs, dense, kp, out, ldo, k, max_tile_rows, gram_out, stream);
}

extern "C" int32_t inmf_blocked_wide_f32(const int64_t* work, int32_t nwork, const int64_t* ptr, const void* pairs,
                                         const float* dense, int32_t kp, float* out, int64_t ldo, int32_t k,
                                         int32_t max_tile_rows, double* gram_out, void* stream) {
    INMF_REQUIRE(kp == 32 && k >= 1 && k <= 32, "the wide kernel is for kp == 32");
    INMF_REQUIRE(ldo >= k, "ldo must be >= k");
    if (nwork <= 0) return 0;
    const size_t smem = 128 + (size_t)max_tile_rows * 32 * sizeof(float);
    INMF_REQUIRE(smem <= 227 * 1024, "dense tile does not fit in shared memory");
    INMF_OPTIN_SMEM((blocked_wide_kernel), 227 * 1024);
    blocked_wide_kernel<<<(unsigned)nwork, 1024, smem, (cudaStream_t)stream>>>(
        work, ptr, reinterpret_cast<const int2*>(pairs), dense, out, ldo, k, gram_out);
    INMF_LAUNCH_CHECK();
    return 0;
}

// =================================================================================================
// Sliced (SELL-8) passes (sell.cuh)
// =================================================================================================
template <int REM, bool MC>
static int32_t sell_pass_one(const int64_t* work, int32_t nwork, const int64_t* slice_ptr, const int32_t* rowid,
                             const void* pairs4, const float* dense, int32_t kp, float* out, int64_t ldo, int32_t k,
                             int32_t max_tile_rows, double* gram_out, int64_t part_stride, int64_t gram_part_stride,
                             const SellMcArgs* mc, void* stream) {
    const size_t smem = 128 + (size_t)max_tile_rows * (32 + REM) * sizeof(float);
    INMF_REQUIRE(smem <= 227 * 1024, "dense tile does not fit in shared memory");
    INMF_OPTIN_SMEM((sell_pass_kernel<REM, MC>), 227 * 1024);
    sell_pass_kernel<REM, MC><<<(unsigned)nwork, 1024, smem, (cudaStream_t)stream>>>(
        work, slice_ptr, rowid, reinterpret_cast<const int4*>(pairs4), dense, kp, out, ldo, k, max_tile_rows, gram_out,
        part_stride, gram_part_stride, mc);
    INMF_LAUNCH_CHECK();
    return 0;
}
// MC: finished gene ranges are added into the multicast copies (inmf_sell_pass_mc_f32)
template <bool MC>
static int32_t sell_pass_dispatch(const int64_t* work, int32_t nwork, const int64_t* slice_ptr, const int32_t* rowid,
                                  const void* pairs4, int32_t height, const float* dense, int32_t kp, float* out,
                                  int64_t ldo, int32_t k, int32_t max_tile_rows, double* gram_out, int64_t part_stride,
                                  int64_t gram_part_stride, const SellMcArgs* mc, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K, "k must be in [1, 64]");
    INMF_REQUIRE(kp >= 32 && kp >= k && kp % 4 == 0 && kp <= 64, "kp must be a multiple of 4 in [max(k, 32), 64]");
    INMF_REQUIRE(ldo >= k, "ldo must be >= k");
    INMF_REQUIRE((reinterpret_cast<uintptr_t>(pairs4) & 15) == 0, "pairs4 must be 16-byte aligned");
    INMF_REQUIRE(height == 8 || (height == 32 && kp == 32) || (height == 33 && kp > 32 && kp <= 40),
                 "slice height must be 8, 32 with kp == 32, or 33 (32 residue-ordered rows) with 32 < kp <= 40");
    if (nwork <= 0) return 0;
    if (height == 33) {
        const size_t smem = 128 + (size_t)max_tile_rows * 40 * sizeof(float);
        INMF_REQUIRE(smem <= 227 * 1024, "dense tile does not fit in shared memory");
        INMF_OPTIN_SMEM((sell32r_pass_kernel<MC>), 227 * 1024);
        sell32r_pass_kernel<MC><<<(unsigned)nwork, SELL32_THREADS, smem, (cudaStream_t)stream>>>(
            work, slice_ptr, rowid, reinterpret_cast<const int4*>(pairs4), dense, kp, out, ldo, k, max_tile_rows, gram_out,
            part_stride, gram_part_stride, mc);
        INMF_LAUNCH_CHECK();
        return 0;
    }
    if (height == 32) {
        const size_t smem = 128 + (size_t)max_tile_rows * 128;
        INMF_REQUIRE(smem <= 227 * 1024, "dense tile does not fit in shared memory");
        INMF_OPTIN_SMEM((sell32_pass_kernel<MC>), 227 * 1024);
        sell32_pass_kernel<MC><<<(unsigned)nwork, SELL32_THREADS, smem, (cudaStream_t)stream>>>(
            work, slice_ptr, rowid, reinterpret_cast<const int4*>(pairs4), dense, out, ldo, k, gram_out, part_stride,
            gram_part_stride, mc);
        INMF_LAUNCH_CHECK();
        return 0;
    }
    const int rem = kp - 32;
#define INMF_SELL_ARGS \
    work, nwork, slice_ptr, rowid, pairs4, dense, kp, out, ldo, k, max_tile_rows, gram_out, part_stride, gram_part_stride, mc, stream
    if (rem == 0) return sell_pass_one<0, MC>(INMF_SELL_ARGS);
    if (rem <= 8) return sell_pass_one<8, MC>(INMF_SELL_ARGS);
    if (rem <= 16) return sell_pass_one<16, MC>(INMF_SELL_ARGS);
    return sell_pass_one<32, MC>(INMF_SELL_ARGS);
#undef INMF_SELL_ARGS
}
extern "C" int32_t inmf_sell_pass_f32(const int64_t* work, int32_t nwork, const int64_t* slice_ptr,
                                      const int32_t* rowid, const void* pairs4, int32_t height, const float* dense,
                                      int32_t kp, float* out, int64_t ldo, int32_t k, int32_t max_tile_rows,
                                      double* gram_out, int64_t part_stride, int64_t gram_part_stride, void* stream) {
    INMF_REQUIRE(part_stride >= 0 && gram_part_stride >= 0, "partial strides must be >= 0");
    return sell_pass_dispatch<false>(work, nwork, slice_ptr, rowid, pairs4, height, dense, kp, out, ldo, k, max_tile_rows,
                                     gram_out, part_stride, gram_part_stride, nullptr, stream);
}
extern "C" int32_t inmf_sell_pass_mc_f32(const int64_t* work, int32_t nwork, const int64_t* slice_ptr,
                                         const int32_t* rowid, const void* pairs4, int32_t height, const float* dense,
                                         int32_t kp, float* out, int32_t k, int32_t max_tile_rows, double* gram_out,
                                         const void* mc_args, void* stream) {
    INMF_REQUIRE(mc_args != nullptr, "mc_args (device struct of multicast addresses and counters) is required");
    return sell_pass_dispatch<true>(work, nwork, slice_ptr, rowid, pairs4, height, dense, kp, out, k, k, max_tile_rows,
                                    gram_out, 0, 0, reinterpret_cast<const SellMcArgs*>(mc_args), stream);
}

// out[i] = part[0][i] + part[1][i] + ... in that order (the deterministic mode's ordered sum of partial buffers).
template <typename T>
__global__ void __launch_bounds__(256) reduce_parts_kernel(const T* __restrict__ part, int nparts, int64_t stride,
                                                           int64_t n, T* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        T acc = part[i];
        for (int p = 1; p < nparts; ++p) acc += part[(size_t)p * stride + i];
        out[i] = acc;
    }
}
template <typename T>
static int32_t reduce_parts_launch(const T* part, int32_t nparts, int64_t stride, int64_t n, T* out, void* stream) {
    INMF_REQUIRE(nparts >= 1 && stride >= n && n >= 0, "bad partial layout");
    if (n == 0) return 0;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)inmf_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    reduce_parts_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(part, nparts, stride, n, out);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_reduce_parts_f32(const float* part, int32_t nparts, int64_t stride, int64_t n, float* out,
                                         void* stream) {
    return reduce_parts_launch<float>(part, nparts, stride, n, out, stream);
}
extern "C" int32_t inmf_reduce_parts_f64(const double* part, int32_t nparts, int64_t stride, int64_t n, double* out,
                                         void* stream) {
    return reduce_parts_launch<double>(part, nparts, stride, n, out, stream);
}

extern "C" int32_t inmf_sell_pass_mixed(const int64_t* work, int32_t nwork, const int64_t* slice_ptr,
                                        const int32_t* rowid, const void* pairs4, const void* dense_bf16, float* out,
                                        int64_t ldo, int32_t k, int32_t max_tile_rows, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= 32, "the mixed-precision passes are for k <= 32");
    INMF_REQUIRE(ldo >= k, "ldo must be >= k");
    INMF_REQUIRE((reinterpret_cast<uintptr_t>(pairs4) & 15) == 0 && (reinterpret_cast<uintptr_t>(dense_bf16) & 15) == 0,
                 "pairs4 and the bf16 operand must be 16-byte aligned");
    if (nwork <= 0) return 0;
    const size_t smem = 128 + (size_t)max_tile_rows * 64;
    INMF_REQUIRE(smem <= 227 * 1024, "dense tile does not fit in shared memory");
    INMF_OPTIN_SMEM((sell32h_pass_kernel), 227 * 1024);
    sell32h_pass_kernel<<<(unsigned)nwork, SELL32_THREADS, smem, (cudaStream_t)stream>>>(
        work, slice_ptr, rowid, reinterpret_cast<const int4*>(pairs4), reinterpret_cast<const uint16_t*>(dense_bf16), out,
        ldo, k);
    INMF_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t inmf_to_bf16_gram(const float* H, int64_t ld, const int64_t* seg_off, int32_t nseg,
                                     int64_t max_seg_rows, int32_t k, void* dst, double* gram, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= 32 && ld >= k && nseg >= 1 && nseg <= 65535, "bad shape (k <= 32)");
    if (max_seg_rows <= 0) return 0;
    int64_t blocks = (max_seg_rows + GRAMCVT_WARPS * 16 - 1) / (GRAMCVT_WARPS * 16);  // >= 16 rows per warp
    int64_t cap = ((int64_t)inmf_num_sms() * 8 + nseg - 1) / nseg;
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    to_bf16_gram_kernel<<<dim3((unsigned)blocks, (unsigned)nseg), GRAMCVT_WARPS * 32, 0, (cudaStream_t)stream>>>(
        H, ld, seg_off, k, reinterpret_cast<uint32_t*>(dst), gram);
    INMF_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t inmf_to_bf16_rows(const float* src, int64_t ld, int64_t rows, int32_t k, void* dst, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= 32 && ld >= k, "bad shape (k <= 32)");
    if (rows <= 0) return 0;
    int64_t blocks = (rows * 16 + 255) / 256;
    const int64_t cap = (int64_t)inmf_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    to_bf16_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, ld, rows, k,
                                                                            reinterpret_cast<uint32_t*>(dst));
    INMF_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t inmf_sell_fill_f32(const int64_t* ptr, const void* pairs, const int64_t* slice_rows,
                                      const int64_t* slice_ptr, int64_t nslices, int32_t height, void* pairs4,
                                      void* stream) {
    if (nslices <= 0) return 0;
    INMF_REQUIRE(height == 8 || height == 32 || height == -32 || height == 33,
                 "slice height must be 8 or 32 (-32: parity ordered, 33: residue ordered)");
    INMF_REQUIRE((reinterpret_cast<uintptr_t>(pairs4) & 15) == 0, "pairs4 must be 16-byte aligned");
    int64_t blocks = (nslices + 7) / 8;
    const int64_t cap = (int64_t)inmf_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (height == 33)
        sell_fill_residue_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            ptr, reinterpret_cast<const int2*>(pairs), slice_rows, slice_ptr, nslices, reinterpret_cast<int4*>(pairs4));
    else if (height == -32)
        sell_fill_parity_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            ptr, reinterpret_cast<const int2*>(pairs), slice_rows, slice_ptr, nslices, reinterpret_cast<int4*>(pairs4));
    else if (height == 8)
        sell_fill_kernel<8><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            ptr, reinterpret_cast<const int2*>(pairs), slice_rows, slice_ptr, nslices, reinterpret_cast<int4*>(pairs4));
    else
        sell_fill_kernel<32><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            ptr, reinterpret_cast<const int2*>(pairs), slice_rows, slice_ptr, nslices, reinterpret_cast<int4*>(pairs4));
    INMF_LAUNCH_CHECK();
    return 0;
}

// =================================================================================================
// tcgen05 diagnostics
// =================================================================================================
extern "C" int32_t inmf_tc_selftest(const float* A, const float* B, float* C, int32_t n, int32_t three_pass,
                                    void* stream) {
    INMF_REQUIRE(n >= 16 && n <= 64 && n % 16 == 0, "n must be 16, 32, 48 or 64");
    const size_t smem = 2 * TC_A_BYTES + 2 * TC_KC * 64 * 4 + 64;
    INMF_OPTIN_SMEM((tc_selftest_kernel), 227 * 1024);
    tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, C, n, three_pass);
    INMF_LAUNCH_CHECK();
    return 0;
}

// =================================================================================================
// Dense-tile (tcgen05) passes, fp32 only
// =================================================================================================
static long long* g_tc_dbg = nullptr;  // optional cycle-counter dump (set through inmf_tc_debug_buffer)
extern "C" void inmf_tc_debug_buffer(void* p) { g_tc_dbg = (long long*)p; }

extern "C" int32_t inmf_tc_relayout_f32(const float* D, int64_t ldd, const int64_t* segd, int32_t nseg,
                                        int64_t max_seg_chunks, float* out_hi, float* out_lo, int32_t k, int32_t npad,
                                        void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K && npad >= k && npad % 16 == 0 && npad <= 64, "bad k / npad");
    INMF_REQUIRE(nseg >= 1 && nseg <= 65535, "nseg out of range");
    if (max_seg_chunks <= 0) return 0;
    int64_t blocks = max_seg_chunks;
    if (blocks > inmf_num_sms() * 16) blocks = inmf_num_sms() * 16;
    tc_relayout_kernel<<<dim3((unsigned)blocks, (unsigned)nseg), 256, 0, (cudaStream_t)stream>>>(D, ldd, segd, out_hi,
                                                                                               out_lo, k, npad);
    INMF_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t inmf_tc_pass_f32(const int64_t* work, int32_t nwork, const int64_t* ptr, const void* pairs,
                                    const float* dc_hi, const float* dc_lo, float* out, int64_t ldo, int32_t k,
                                    int32_t npad, int32_t three_pass, int32_t max_ctas, void* stream) {
    INMF_REQUIRE(k >= 1 && k <= INMF_MAX_K && npad >= k && npad % 16 == 0 && npad <= 64, "bad k / npad");
    INMF_REQUIRE(ldo >= k, "ldo must be >= k");
    if (nwork <= 0) return 0;
    const TcSmemLayout L = tc_smem_layout(npad);
    INMF_REQUIRE(L.total + 1024 <= 227 * 1024, "shared memory budget exceeded");
    INMF_OPTIN_SMEM((tc_pass_kernel), 227 * 1024);
    int grid = inmf_num_sms();
    if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
    if (grid > nwork) grid = nwork;
    tc_pass_kernel<<<grid, TC_THREADS, L.total + 1024, (cudaStream_t)stream>>>(
        work, nwork, ptr, reinterpret_cast<const int2*>(pairs), dc_hi, dc_lo, out, ldo, k, npad, three_pass,
        g_tc_dbg);
    INMF_LAUNCH_CHECK();
    return 0;
}

// =================================================================================================
// Preprocessing passes: normalize + scale_not_center (the producers of scale_data)
// =================================================================================================
#include "preprocess.cuh"

static inline unsigned rows_grid(int64_t n_rows) {
    int64_t blocks = (n_rows + 7) / 8;  // 8 warps (rows) per 256-thread CTA
    const int64_t cap = (int64_t)inmf_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

template <typename T>
static int32_t normalize_rows_launch(const int64_t* rowptr, const int32_t* colidx, const T* vals, int64_t n_rows,
                                     T* norm_vals, double* col_sum, double* col_sumsq, int64_t* n_missing,
                                     void* stream) {
    INMF_REQUIRE(n_rows >= 0, "n_rows must be >= 0");
    if (n_rows == 0) return 0;
    normalize_rows_kernel<T><<<rows_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(
        rowptr, colidx, vals, n_rows, norm_vals, col_sum, col_sumsq, (unsigned long long*)n_missing);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_normalize_rows_f32(const int64_t* rowptr, const int32_t* colidx, const float* vals,
                                           int64_t n_rows, float* norm_vals, double* col_sum, double* col_sumsq,
                                           int64_t* n_missing, void* stream) {
    return normalize_rows_launch<float>(rowptr, colidx, vals, n_rows, norm_vals, col_sum, col_sumsq, n_missing,
                                        stream);
}
extern "C" int32_t inmf_normalize_rows_f64(const int64_t* rowptr, const int32_t* colidx, const double* vals,
                                           int64_t n_rows, double* norm_vals, double* col_sum, double* col_sumsq,
                                           int64_t* n_missing, void* stream) {
    return normalize_rows_launch<double>(rowptr, colidx, vals, n_rows, norm_vals, col_sum, col_sumsq, n_missing,
                                         stream);
}

extern "C" int32_t inmf_select_count(const int64_t* rowptr, const int32_t* colidx, int64_t n_rows,
                                     const int32_t* colmap, int64_t* row_counts, void* stream) {
    INMF_REQUIRE(n_rows >= 0, "n_rows must be >= 0");
    if (n_rows == 0) return 0;
    select_count_kernel<<<rows_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(rowptr, colidx, n_rows, colmap,
                                                                             row_counts);
    INMF_LAUNCH_CHECK();
    return 0;
}

template <typename T>
static int32_t select_scale_launch(const int64_t* rowptr, const int32_t* colidx, const T* vals, int64_t n_rows,
                                   const int32_t* colmap, const double* scaler, const int64_t* new_rowptr,
                                   int32_t* new_colidx, T* new_vals, void* stream) {
    INMF_REQUIRE(n_rows >= 0, "n_rows must be >= 0");
    if (n_rows == 0) return 0;
    select_scale_kernel<T><<<rows_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(
        rowptr, colidx, vals, n_rows, colmap, scaler, new_rowptr, new_colidx, new_vals);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_select_scale_f32(const int64_t* rowptr, const int32_t* colidx, const float* vals,
                                         int64_t n_rows, const int32_t* colmap, const double* scaler,
                                         const int64_t* new_rowptr, int32_t* new_colidx, float* new_vals,
                                         void* stream) {
    return select_scale_launch<float>(rowptr, colidx, vals, n_rows, colmap, scaler, new_rowptr, new_colidx, new_vals,
                                      stream);
}
extern "C" int32_t inmf_select_scale_f64(const int64_t* rowptr, const int32_t* colidx, const double* vals,
                                         int64_t n_rows, const int32_t* colmap, const double* scaler,
                                         const int64_t* new_rowptr, int32_t* new_colidx, double* new_vals,
                                         void* stream) {
    return select_scale_launch<double>(rowptr, colidx, vals, n_rows, colmap, scaler, new_rowptr, new_colidx,
                                       new_vals, stream);
}

// =================================================================================================
// Seeded initial factors on the device: the stream of numpy's legacy global RNG,
//   np.random.seed(seed); np.random.uniform(0, scale, n)      (_iNMF_ANLS.py:103-108, _online_iNMF.py:147-160)
// = MT19937 seeded by init_genrand, two 32-bit outputs per double ((a >> 5) * 2^26 + (b >> 6)) / 2^53.
// The twist of a 624-word block is 227-wide parallel in three dependent phases, so one CTA walks the stream;
// every arithmetic step is exact in double, which makes the result bit-identical to numpy's.
// =================================================================================================
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// 512 threads: threads 0..255 twist block b into the other state buffer (three phases, one barrier each) while
// threads 256..511 temper and write block b-1 from the current one, so a 624-word block costs three barriers.
__global__ void __launch_bounds__(512, 1) mt19937_uniform_kernel(uint32_t seed, int64_t n, double scale,
                                                                 double* __restrict__ out) {
    __shared__ uint32_t st[2][624];
    const int t = threadIdx.x;
    if (t == 0) {
        uint32_t s = seed;
        st[0][0] = s;
        for (int i = 1; i < 624; ++i) {
            s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i;
            st[0][i] = s;
        }
    }
    __syncthreads();
    const int64_t nblocks = (n + 311) / 312;
    int cur = 0;
    for (int64_t b = 0; b <= nblocks; ++b) {
        const uint32_t* __restrict__ C = st[cur];
        uint32_t* __restrict__ N = st[cur ^ 1];
        const bool twist = t < 227 && b < nblocks;
        const int e = t - 256;
        const bool emit = e >= 0 && e < 104 && b >= 1;
#pragma unroll
        for (int ph = 0; ph < 3; ++ph) {
            const int i = ph * 227 + t;
            if (twist && i < 624) {
                const uint32_t x1 = (i + 1 == 624) ? N[0] : C[i + 1];
                const uint32_t xm = (i + 397 < 624) ? C[i + 397] : N[i + 397 - 624];
                const uint32_t y = (C[i] & 0x80000000u) | (x1 & 0x7fffffffu);
                N[i] = xm ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            if (emit) {
                const int d = ph * 104 + e;
                const int64_t o = (b - 1) * 312 + d;
                if (o < n) {
                    const uint32_t a = mt_temper(C[2 * d]) >> 5, bb = mt_temper(C[2 * d + 1]) >> 6;
                    const double u = ((double)a * 67108864.0 + (double)bb) / 9007199254740992.0;
                    out[o] = 0.0 + scale * u;
                }
            }
            __syncthreads();
        }
        cur ^= 1;
    }
}

extern "C" int32_t inmf_mt19937_uniform(uint32_t seed, int64_t n, double scale, double* out, void* stream) {
    INMF_REQUIRE(n >= 0, "n must be >= 0");
    if (n == 0) return 0;
    mt19937_uniform_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(seed, n, scale, out);
    INMF_LAUNCH_CHECK();
    return 0;
}

// =================================================================================================
// Device builders of the blocked copies of X (count / fill around a scan done by the caller)
// =================================================================================================
#include "blockfmt.cuh"

template <typename T>
static int32_t blockfmt1_launch(int32_t fill, const int64_t* rowptr, const int32_t* colidx, const T* vals,
                                const int64_t* seg1, int32_t nseg, int64_t n_rows, int32_t NB, int32_t gbs,
                                int32_t row_bytes, int64_t* cnt_or_ptr, void* pairs, void* stream) {
    INMF_REQUIRE(nseg >= 1 && NB >= 1 && gbs >= 1 && row_bytes >= 1, "bad format-1 geometry");
    if (n_rows <= 0) return 0;
    typedef typename PairTraits<T>::Pair Pair;
    if (fill)
        blockfmt1_kernel<true, T><<<rows_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(
            rowptr, colidx, vals, seg1, nseg, n_rows, NB, gbs, row_bytes, cnt_or_ptr, (Pair*)pairs);
    else
        blockfmt1_kernel<false, T><<<rows_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(
            rowptr, colidx, vals, seg1, nseg, n_rows, NB, gbs, row_bytes, cnt_or_ptr, (Pair*)pairs);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_blockfmt1_f32(int32_t fill, const int64_t* rowptr, const int32_t* colidx, const float* vals,
                                      const int64_t* seg1, int32_t nseg, int64_t n_rows, int32_t NB, int32_t gbs,
                                      int32_t row_bytes, int64_t* cnt_or_ptr, void* pairs, void* stream) {
    return blockfmt1_launch<float>(fill, rowptr, colidx, vals, seg1, nseg, n_rows, NB, gbs, row_bytes, cnt_or_ptr,
                                   pairs, stream);
}
extern "C" int32_t inmf_blockfmt1_f64(int32_t fill, const int64_t* rowptr, const int32_t* colidx, const double* vals,
                                      const int64_t* seg1, int32_t nseg, int64_t n_rows, int32_t NB, int32_t gbs,
                                      int32_t row_bytes, int64_t* cnt_or_ptr, void* pairs, void* stream) {
    return blockfmt1_launch<double>(fill, rowptr, colidx, vals, seg1, nseg, n_rows, NB, gbs, row_bytes, cnt_or_ptr,
                                    pairs, stream);
}

template <typename T>
static int32_t blockfmt2_launch(int32_t fill, const int64_t* rowptr, const int32_t* colidx, const T* vals,
                                const int64_t* blk2, int32_t nblocks, int32_t g, int32_t row_bytes, int32_t W,
                                int64_t* cnt, const int64_t* ptr, uint16_t* offs, void* pairs, void* stream) {
    INMF_REQUIRE(g >= 1 && row_bytes >= 1 && W >= 1 && W <= 32, "bad format-2 geometry");
    const size_t smem = (size_t)W * g * sizeof(uint16_t);
    INMF_REQUIRE(smem <= 200 * 1024, "format-2 builder: W * genes counters exceed shared memory");
    if (nblocks <= 0) return 0;
    typedef typename PairTraits<T>::Pair Pair;
    INMF_OPTIN_SMEM((blockfmt2_kernel<true, T>), 200 * 1024);
    INMF_OPTIN_SMEM((blockfmt2_kernel<false, T>), 200 * 1024);
    if (fill)
        blockfmt2_kernel<true, T><<<(unsigned)nblocks, W * 32, smem, (cudaStream_t)stream>>>(
            rowptr, colidx, vals, blk2, g, row_bytes, cnt, ptr, offs, (Pair*)pairs);
    else
        blockfmt2_kernel<false, T><<<(unsigned)nblocks, W * 32, smem, (cudaStream_t)stream>>>(
            rowptr, colidx, vals, blk2, g, row_bytes, cnt, ptr, offs, (Pair*)pairs);
    INMF_LAUNCH_CHECK();
    return 0;
}
extern "C" int32_t inmf_blockfmt2_f32(int32_t fill, const int64_t* rowptr, const int32_t* colidx, const float* vals,
                                      const int64_t* blk2, int32_t nblocks, int32_t g, int32_t row_bytes, int32_t W,
                                      int64_t* cnt, const int64_t* ptr, uint16_t* offs, void* pairs, void* stream) {
    return blockfmt2_launch<float>(fill, rowptr, colidx, vals, blk2, nblocks, g, row_bytes, W, cnt, ptr, offs, pairs,
                                   stream);
}
extern "C" int32_t inmf_blockfmt2_f64(int32_t fill, const int64_t* rowptr, const int32_t* colidx, const double* vals,
                                      const int64_t* blk2, int32_t nblocks, int32_t g, int32_t row_bytes, int32_t W,
                                      int64_t* cnt, const int64_t* ptr, uint16_t* offs, void* pairs, void* stream) {
    return blockfmt2_launch<double>(fill, rowptr, colidx, vals, blk2, nblocks, g, row_bytes, W, cnt, ptr, offs, pairs,
                                    stream);
}

// =================================================================================================
// quantile_norm (quantile.cuh)
// =================================================================================================
extern "C" int32_t inmf_qn_max_factor(const double* H, int64_t ld, int64_t n, const int32_t* dims, int32_t m,
                                      int32_t center, double* stats_ws, int32_t* cluster, void* stream) {
    INMF_REQUIRE(m >= 1 && m <= 64 && ld >= 1 && n >= 2, "bad shape (1 <= factors <= 64, at least 2 cells)");
    cudaStream_t st = (cudaStream_t)stream;
    INMF_CUDA_CHECK(cudaMemsetAsync(stats_ws, 0, 2 * (size_t)m * sizeof(double), st));
    int64_t blocks = (n + 63) / 64;
    if (blocks > (int64_t)inmf_num_sms() * 8) blocks = (int64_t)inmf_num_sms() * 8;
    qn_colstats_kernel<<<(unsigned)blocks, 256, 0, st>>>(H, ld, n, dims, m, stats_ws);
    INMF_LAUNCH_CHECK();
    blocks = (n + 255) / 256;
    if (blocks > (int64_t)inmf_num_sms() * 8) blocks = (int64_t)inmf_num_sms() * 8;
    qn_max_factor_kernel<<<(unsigned)blocks, 256, 0, st>>>(H, ld, n, dims, m, stats_ws, center, cluster);
    INMF_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t inmf_qn_knn(const double* H, int64_t ld, int64_t n, const int32_t* dims, int32_t m, int32_t K,
                               int32_t* knn, void* stream) {
    INMF_REQUIRE(m >= 1 && m <= 64 && K >= 1 && K <= QN_KNN_MAXK && n >= 1 && n < ((int64_t)1 << 31), "bad shape");
    const size_t smem = ((size_t)m * QN_KNN_THREADS + (size_t)QN_KNN_TILE * m) * sizeof(double);
    INMF_OPTIN_SMEM((qn_knn_kernel), 200 * 1024);
    int64_t blocks = (n + QN_KNN_THREADS - 1) / QN_KNN_THREADS;
    if (blocks > (int64_t)inmf_num_sms() * 4) blocks = (int64_t)inmf_num_sms() * 4;
    qn_knn_kernel<<<(unsigned)blocks, QN_KNN_THREADS, smem, (cudaStream_t)stream>>>(H, ld, n, dims, m, K, knn);
    INMF_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t inmf_qn_vote(const int32_t* knn, int32_t K, int64_t n, int32_t* cluster, void* stream) {
    INMF_REQUIRE(K >= 1 && K <= 64 && n >= 0, "bad shape");
    if (n == 0) return 0;
    const int use_smem = n <= QN_VOTE_SMEM_CELLS;
    INMF_OPTIN_SMEM((qn_vote_kernel), 227 * 1024 - 1024);
    qn_vote_kernel<<<1, 32, use_smem ? (size_t)((n + 15) & ~(int64_t)15) : 0, (cudaStream_t)stream>>>(knn, K, n, cluster,
                                                                                                 use_smem);
    INMF_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t inmf_qn_align(const int64_t* group, int32_t ngroups, double* Hk, int64_t ldk, const double* Href,
                                 int64_t ldr, const int32_t* memb2, const int32_t* off2, const int32_t* memb1,
                                 const int32_t* off1, const int32_t* samp, int32_t quantiles, int32_t max_sample,
                                 void* stream) {
    INMF_REQUIRE(quantiles >= 1 && quantiles <= QN_MAX_QUANTILES, "1 <= quantiles <= 512");
    INMF_REQUIRE(max_sample >= 1 && max_sample <= 8192, "1 <= max_sample <= 8192");
    if (ngroups <= 0) return 0;
    int spow2 = 2;
    while (spow2 < max_sample) spow2 <<= 1;
    const size_t smem = (2 * (size_t)spow2 + 2 * (size_t)(quantiles + 1)) * sizeof(double);
    INMF_OPTIN_SMEM((qn_align_kernel), 200 * 1024);
    qn_align_kernel<<<(unsigned)ngroups, QN_ALIGN_THREADS, smem, (cudaStream_t)stream>>>(
        group, Hk, ldk, Href, ldr, memb2, off2, memb1, off1, samp, quantiles, spow2);
    INMF_LAUNCH_CHECK();
    return 0;
}
