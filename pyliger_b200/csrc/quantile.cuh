// quantile_norm on the device: the consumer of H.
// Reference: /root/reference/src/pyliger/tools/_quantile_norm.py:9-195, clustering/_utilities.py:11-95.
//
//   qn_colstats / qn_max_factor   column scaling of H + arg-max factor per cell        (_quantile_norm.py:110-125)
//   qn_knn                        exact k nearest neighbours in factor space, brute force over shared-memory tiles
//                                 (run_knn: sklearn NearestNeighbors, kd_tree, clustering/_utilities.py:11-17)
//   qn_vote                       the in-place, cell-order-dependent majority vote (cluster_vote, :57-79): a single
//                                 warp per dataset walks the cells in order, its labels in shared memory
//   qn_align                      per (cluster, factor): quantiles of both datasets (mquantiles alphap = betap = 1),
//                                 tie handling and linear warp of the values          (_quantile_norm.py:133-183)
// Everything is double precision (obsm["H"] is float64): these kernels are integer / selection work plus a few flops per
// value, bound by memory traffic and, for the vote, by latency -- not tensor work.
#pragma once
#include "inmf_common.cuh"

// ---------------------------------------------------------------------------------------------------------
// column sums / sums of squares of the selected factors:  out[0][m] = sum_c H[c, dims[m]],  out[1][m] = sum of squares
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qn_colstats_kernel(const double* __restrict__ H, int64_t ld, int64_t n,
                                                          const int32_t* __restrict__ dims, int m,
                                                          double* __restrict__ out) {
    // thread = (row lane, factor): blockDim.x = 256 = 8 row lanes x 32 factor slots (m <= 64: two passes)
    const int fs = threadIdx.x & 31, rl = threadIdx.x >> 5;
    for (int f = fs; f < m; f += 32) {
        const int col = dims[f];
        double s = 0.0, q = 0.0;
        for (int64_t r = (int64_t)blockIdx.x * 8 + rl; r < n; r += (int64_t)gridDim.x * 8) {
            const double v = H[r * ld + col];
            s += v;
            q += v * v;
        }
        atomicAdd(&out[f], s);
        atomicAdd(&out[m + f], q);
    }
}

// cluster[c] = argmax_m scaled(H[c, dims[m]]) with numpy's semantics (first maximum; a NaN beats everything).
// center = 0: H / sqrt(sumsq / (n-1));  center = 1: (H - mean) / std(ddof = 1).
__global__ void __launch_bounds__(256) qn_max_factor_kernel(const double* __restrict__ H, int64_t ld, int64_t n,
                                                            const int32_t* __restrict__ dims, int m,
                                                            const double* __restrict__ stats, int center,
                                                            int32_t* __restrict__ cluster) {
    __shared__ double mean_s[64], scale_s[64];
    if (threadIdx.x < m) {
        const double s = stats[threadIdx.x], q = stats[m + threadIdx.x];
        const double mu = s / (double)n;
        mean_s[threadIdx.x] = center ? mu : 0.0;
        // np.std(ddof=1) = sqrt(sum((x - mean)^2) / (n - 1)); the uncentred scale is sqrt(sum(x^2) / (n - 1))
        scale_s[threadIdx.x] = center ? sqrt((q - (double)n * mu * mu) / (double)(n - 1)) : sqrt(q / (double)(n - 1));
    }
    __syncthreads();
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        double best = 0.0;
        int bi = 0;
        for (int f = 0; f < m; ++f) {
            const double v = (H[r * ld + dims[f]] - mean_s[f]) / scale_s[f];
            if (f == 0 || v > best || (v != v && best == best)) {
                best = v;
                bi = f;
            }
        }
        cluster[r] = bi;
    }
}

// ---------------------------------------------------------------------------------------------------------
// exact kNN (squared Euclidean distance on the selected factors), self included, K <= 64
// ---------------------------------------------------------------------------------------------------------
#define QN_KNN_THREADS 128
#define QN_KNN_TILE 64
#define QN_KNN_MAXK 64

__global__ void __launch_bounds__(QN_KNN_THREADS) qn_knn_kernel(const double* __restrict__ H, int64_t ld, int64_t n,
                                                                const int32_t* __restrict__ dims, int m, int K,
                                                                int32_t* __restrict__ knn) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* qs = reinterpret_cast<double*>(smem_raw);            // [m][QN_KNN_THREADS]  queries, factor-major
    double* cs = qs + (size_t)m * QN_KNN_THREADS;                // [QN_KNN_TILE][m]     candidates, point-major
    const int tid = threadIdx.x;
    for (int64_t q0 = (int64_t)blockIdx.x * QN_KNN_THREADS; q0 < n; q0 += (int64_t)gridDim.x * QN_KNN_THREADS) {
        const int64_t q = q0 + tid;
        const bool live = q < n;
        __syncthreads();
        for (int f = 0; f < m; ++f) qs[f * QN_KNN_THREADS + tid] = live ? H[q * ld + dims[f]] : 0.0;
        double bd[QN_KNN_MAXK];
        int bi[QN_KNN_MAXK];
        int cnt = 0;
        double worst = 0.0;  // bd[K-1] once the list is full
        for (int64_t c0 = 0; c0 < n; c0 += QN_KNN_TILE) {
            const int nc = (int)((n - c0 < QN_KNN_TILE) ? (n - c0) : QN_KNN_TILE);
            __syncthreads();
            for (int e = tid; e < nc * m; e += QN_KNN_THREADS) cs[e] = H[(c0 + e / m) * ld + dims[e % m]];
            __syncthreads();
            if (!live) continue;
            for (int c = 0; c < nc; ++c) {
                const double* cp = cs + c * m;
                double d = 0.0;
                for (int f = 0; f < m; ++f) {
                    const double t = qs[f * QN_KNN_THREADS + tid] - cp[f];
                    d = fma(t, t, d);
                }
                const int id = (int)(c0 + c);
                if (cnt == K && !(d < worst)) continue;  // ties at the boundary keep the earlier (smaller) index
                int pos = cnt < K ? cnt : K - 1;         // insert, keeping (distance, index) order
                while (pos > 0 && bd[pos - 1] > d) {
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = d;
                bi[pos] = id;
                if (cnt < K) ++cnt;
                if (cnt == K) worst = bd[K - 1];
            }
        }
        if (live)
            for (int j = 0; j < K; ++j) knn[q * K + j] = j < cnt ? bi[j] : (int)q;  // n < K: pad with the cell itself
    }
}

// ---------------------------------------------------------------------------------------------------------
// cluster_vote: for i = 0 .. n-1 IN ORDER, cluster[i] = the most frequent label among its K neighbours (ties: the
// largest label), reading labels that earlier cells of the same sweep have already replaced.  One warp.
// ---------------------------------------------------------------------------------------------------------
#define QN_VOTE_SMEM_CELLS (200 * 1024)

__global__ void __launch_bounds__(32) qn_vote_kernel(const int32_t* __restrict__ knn, int K, int64_t n,
                                                     int32_t* __restrict__ cluster, int use_smem) {
    extern __shared__ unsigned char lab_s[];  // labels < 64 fit a byte (n <= QN_VOTE_SMEM_CELLS), else global memory
    __shared__ int cnt[64];
    const int lane = threadIdx.x;
    volatile int32_t* vcl = cluster;
    if (use_smem)
        for (int64_t i = lane; i < n; i += 32) lab_s[i] = (unsigned char)cluster[i];
    __syncwarp();
    // neighbour lists do not depend on the sweep: the next cell's are loaded while this one is voted on
    int nb0 = (lane < K && n > 0) ? knn[lane] : 0;
    int nb1 = (lane + 32 < K && n > 0) ? knn[lane + 32] : 0;
    for (int64_t i = 0; i < n; ++i) {
        int nx0 = 0, nx1 = 0;
        if (i + 1 < n) {
            if (lane < K) nx0 = knn[(i + 1) * K + lane];
            if (lane + 32 < K) nx1 = knn[(i + 1) * K + lane + 32];
        }
        cnt[lane] = 0;
        cnt[lane + 32] = 0;
        __syncwarp();
        if (lane < K) atomicAdd(&cnt[use_smem ? (int)lab_s[nb0] : (vcl[nb0] & 63)], 1);
        if (lane + 32 < K) atomicAdd(&cnt[use_smem ? (int)lab_s[nb1] : (vcl[nb1] & 63)], 1);
        __syncwarp();
        const int c0 = cnt[lane], c1 = cnt[lane + 32];
        const int k0 = c0 > 0 ? ((c0 << 8) | lane) : -1;
        const int k1 = c1 > 0 ? ((c1 << 8) | (lane + 32)) : -1;
        int best = k0 > k1 ? k0 : k1;  // most frequent label; among equals the LARGEST label (:66-78)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int t = __shfl_xor_sync(INMF_FULL_MASK, best, o);
            best = t > best ? t : best;
        }
        if (lane == 0) {
            if (use_smem) lab_s[i] = (unsigned char)(best & 0xff);
            else vcl[i] = best & 0xff;
        }
        __syncwarp();
        nb0 = nx0;
        nb1 = nx1;
    }
    if (use_smem) {
        __syncwarp();
        for (int64_t i = lane; i < n; i += 32) cluster[i] = (int32_t)lab_s[i];
    }
}

// ---------------------------------------------------------------------------------------------------------
// quantile alignment of one dataset against the reference dataset, one CTA per (cluster, factor) group
// ---------------------------------------------------------------------------------------------------------
// group[w] = { cluster j, factor column i, samp2_off, samp2_n, samp1_off, samp1_n }  (int64 x 6)
//   samp*_n = 0: use all members of the cluster (count <= max_sample); otherwise samp[off .. off+n) are positions
//   inside the cluster's member list (the prefix of the reference's np.random.permutation, drawn on the host).
// memb2 / off2: cells of THIS dataset sorted by cluster (off2[j] .. off2[j+1]); memb1 / off1: the reference dataset.
#define QN_GROUP_FIELDS 6
#define QN_ALIGN_THREADS 256
#define QN_MAX_QUANTILES 512

__device__ __forceinline__ void qn_bitonic_sort(double* a, int npow2) {
    for (int size = 2; size <= npow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (npow2 >> 1); t += blockDim.x) {
                const int lo = 2 * stride * (t / stride) + (t % stride);
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const double x = a[lo], y = a[hi];
                if ((x > y) == up) {
                    a[lo] = y;
                    a[hi] = x;
                }
            }
        }
    }
    __syncthreads();
}

// scipy.stats.mstats.mquantiles(x, linspace(0, 1, Q + 1), alphap = 1, betap = 1) of the sorted sample s[0 .. n)
__device__ __forceinline__ double qn_quantile(const double* s, int n, int mq, int Q) {
    const double p = (double)mq / (double)Q;
    const double aleph = (double)n * p + (1.0 - p);
    double kk = floor(fmin(fmax(aleph, 1.0), (double)(n - 1)));   // np.clip(aleph, 1, n-1): the upper bound wins
    if ((double)(n - 1) < 1.0) kk = (double)(n - 1);
    const double gamma = fmin(fmax(aleph - kk, 0.0), 1.0);
    const int ki = (int)kk;
    const double lo = s[ki - 1 < 0 ? n - 1 : ki - 1];             // x[k-1] with numpy's wrap-around for k = 0
    return (1.0 - gamma) * lo + gamma * s[ki];
}

__global__ void __launch_bounds__(QN_ALIGN_THREADS)
qn_align_kernel(const int64_t* __restrict__ group, double* Hk, int64_t ldk, const double* Href, int64_t ldr,
                const int32_t* __restrict__ memb2, const int32_t* __restrict__ off2,
                const int32_t* __restrict__ memb1, const int32_t* __restrict__ off1,
                const int32_t* __restrict__ samp, int Q, int spow2) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* s2 = reinterpret_cast<double*>(smem_raw);   // [spow2]
    double* s1 = s2 + spow2;                            // [spow2]
    double* q2 = s1 + spow2;                            // [Q + 1]
    double* q1 = q2 + (Q + 1);                          // [Q + 1]
    __shared__ double red[QN_ALIGN_THREADS];
    __shared__ double mm[2], tie_mean;
    __shared__ int degenerate;
    const int64_t* gd = group + (size_t)blockIdx.x * QN_GROUP_FIELDS;
    const int j = (int)gd[0], col = (int)gd[1];
    const int64_t so2 = gd[2], so1 = gd[4];
    const int sn2 = (int)gd[3], sn1 = (int)gd[5];
    const int32_t* m2 = memb2 + off2[j];
    const int32_t* m1 = memb1 + off1[j];
    const int n2 = off2[j + 1] - off2[j], n1 = off1[j + 1] - off1[j];
    const int c2 = sn2 ? sn2 : n2, c1 = sn1 ? sn1 : n1;           // sample sizes
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    for (int t = threadIdx.x; t < spow2; t += blockDim.x) {
        s2[t] = t < c2 ? Hk[(int64_t)m2[sn2 ? samp[so2 + t] : t] * ldk + col] : inf;
        s1[t] = t < c1 ? Href[(int64_t)m1[sn1 ? samp[so1 + t] : t] * ldr + col] : inf;
    }
    // min / max over ALL cells of the cluster (not just the sample), _quantile_norm.py:158-163
    double lo = inf, hi = -inf;
    for (int t = threadIdx.x; t < n2; t += blockDim.x) {
        const double v = Hk[(int64_t)m2[t] * ldk + col];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
    red[threadIdx.x] = lo;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] = fmin(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) mm[0] = red[0];
    __syncthreads();
    red[threadIdx.x] = hi;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) mm[1] = red[0];
    __syncthreads();
    qn_bitonic_sort(s2, spow2);
    qn_bitonic_sort(s1, spow2);
    for (int t = threadIdx.x; t <= Q; t += blockDim.x) {
        q2[t] = qn_quantile(s2, c2, t, Q);
        q1[t] = qn_quantile(s1, c1, t, Q);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (q2[Q] < mm[1]) q2[Q] = mm[1];
        if (q2[0] > mm[0]) q2[0] = mm[0];
        double sum1 = 0.0, sum2 = 0.0, zs = 0.0;
        int u1 = 1, u2 = 1, nz = 0;
        for (int t = 0; t <= Q; ++t) {
            sum1 += q1[t];
            sum2 += q2[t];
            if (t > 0 && q1[t] != q1[t - 1]) ++u1;   // q1, q2 are non-decreasing: distinct values = change points + 1
            if (t > 0 && q2[t] != q2[t - 1]) ++u2;
            if (q2[t] == 0.0) {
                zs += q1[t];
                ++nz;
            }
        }
        degenerate = (sum1 == 0.0 || sum2 == 0.0 || u1 < 2 || u2 < 2) ? 1 : 0;
        tie_mean = nz ? zs / (double)nz : 0.0;   // _mean_ties: y[x == 0] = mean(y[x == 0])
    }
    __syncthreads();
    if (!degenerate) {
        for (int t = threadIdx.x; t <= Q; t += blockDim.x)
            if (q2[t] == 0.0) q1[t] = tie_mean;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n2; t += blockDim.x) {
        double* hp = Hk + (int64_t)m2[t] * ldk + col;
        double out = 0.0;
        if (!degenerate) {
            // np.interp(x, q2, q1): j = last index with q2[j] <= x
            const double x = *hp;
            int a = 0, b = Q + 1;
            while (a < b) {
                const int mid = (a + b) >> 1;
                if (x >= q2[mid]) a = mid + 1;
                else b = mid;
            }
            const int jj = a - 1;
            if (jj < 0) out = q1[0];
            else if (jj >= Q) out = q1[Q];
            else if (q2[jj] == x) out = q1[jj];
            else {
                const double slope = (q1[jj + 1] - q1[jj]) / (q2[jj + 1] - q2[jj]);
                out = slope * (x - q2[jj]) + q1[jj];
                if (out != out) {
                    out = slope * (x - q2[jj + 1]) + q1[jj + 1];
                    if (out != out && q1[jj] == q1[jj + 1]) out = q1[jj];
                }
            }
        }
        *hp = out;
    }
}
