// tcgen05 / TMEM helpers for the dense-tile (tensor-core) form of the X-streaming passes (sm_100a).
//
// Shared-memory operand layouts are the canonical no-swizzle ("interleave") UMMA layouts, built from
// 8 x 16-byte core matrices:
//   A  (M x K, K-major, 4-byte elements):  (m, k) -> (m/8)*SBO + (m%8)*16 + (k/4)*LBO + (k%4)*4
//   B  (N x K, K-major, same form):        (n, k) -> (n/8)*SBO + (n%8)*16 + (k/4)*LBO + (k%4)*4
// (MN-major B with kind::tf32 and no swizzle returned zeros on B200; both operands are K-major.)
// One kind::tf32 MMA consumes K = 8 (32 bytes).  Descriptor bit fields as in CUTLASS
// cute/arch/mma_sm100_desc.hpp (SmemDescriptor / InstrDescriptor).
#pragma once
#include "blocked.cuh"

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for sm_100
    return d;                // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE
}

// kind::tf32, D = f32, A and B K-major, M = 128, N = n (multiple of 16)
__host__ __device__ constexpr uint32_t umma_idesc_tf32_m128(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp receives TMEM lane (base lane + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
        "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
        "%23, %24, %25, %26, %27, %28, %29, %30, %31};\n" ::"r"(r[0]),
        "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
        "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

#define TC_M 128           // rows of X per tile (UMMA M)
#define TC_KC 64           // contraction entities per chunk
#define TC_A_LBO 128u      // bytes between the two 16-byte K columns of one MMA
#define TC_A_SBO ((TC_KC / 4) * 128u)   // bytes between 8-row groups of the A tile
#define TC_A_BYTES (TC_M * TC_KC * 4)   // 32 KB

// byte offset of element (m, k) inside an A tile
__host__ __device__ constexpr uint32_t tc_a_offset(uint32_t m, uint32_t k) {
    return (m >> 3) * TC_A_SBO + (m & 7u) * 16u + (k >> 2) * 128u + (k & 3u) * 4u;
}
// byte offset of element (n, k) inside a B chunk (N x TC_KC, K-major): [n/8][k/4][n%8][k%4]
__host__ __device__ constexpr uint32_t tc_b_offset(uint32_t n, uint32_t k) {
    return (n >> 3) * ((TC_KC / 4) * 128u) + (n & 7u) * 16u + (k >> 2) * 128u + (k & 3u) * 4u;
}
#define TC_B_SBO ((TC_KC / 4) * 128u)   // bytes between groups of 8 columns (n)
#define TC_B_LBO 128u                    // bytes between the two 16-byte K columns of one MMA

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// ------------------------------------------------------------------------------------------------
// Self test: C (128 x n) = A (128 x 64) * B (64 x n) through one CTA, 1 or 3 TF32 passes.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int n,
                   int three_pass) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* a_hi = smem_raw;
    unsigned char* a_lo = a_hi + TC_A_BYTES;
    unsigned char* b_hi = a_lo + TC_A_BYTES;
    unsigned char* b_lo = b_hi + TC_KC * 64 * 4;
    uint64_t* bar = reinterpret_cast<uint64_t*>(b_lo + TC_KC * 64 * 4);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < TC_M * TC_KC; e += blockDim.x) {
        const int m = e / TC_KC, k = e % TC_KC;
        const float v = A[e], h = tf32_hi(v);
        *reinterpret_cast<float*>(a_hi + tc_a_offset(m, k)) = h;
        *reinterpret_cast<float*>(a_lo + tc_a_offset(m, k)) = v - h;
    }
    for (int e = tid; e < TC_KC * n; e += blockDim.x) {
        const int k = e / n, c = e % n;
        const float v = B[e], h = tf32_hi(v);
        *reinterpret_cast<float*>(b_hi + tc_b_offset(c, k)) = h;
        *reinterpret_cast<float*>(b_lo + tc_b_offset(c, k)) = v - h;
    }
    if (three_pass & 256) {  // debug: every operand byte pattern = bf16 1.0
        for (int e = tid; e < (2 * TC_A_BYTES + 2 * TC_KC * 64 * 4) / 4; e += blockDim.x)
            reinterpret_cast<uint32_t*>(smem_raw)[e] = 0x3f803f80u;
    }
    fence_async_smem();
    if (tid == 0) mbar_init(bar, 1);
    if (warp == 0) tmem_alloc(slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (three_pass & 2) {  // debug: prefill the accumulator with lane*100 + column
        uint32_t v[32];
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint((float)((warp * 32 + lane) * 100 + j));
        tmem_st_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (three_pass & 8) C[0] = __uint_as_float(tmem);  // debug: raw TMEM base address (overwritten below)
    if (tid == 0 && !(three_pass & 4)) {
        const uint32_t idesc = umma_idesc_tf32_m128(n);
        for (int ks = 0; ks < TC_KC / 8; ++ks) {
            const uint32_t alb = (three_pass & 32) ? TC_A_SBO : TC_A_LBO, asb = (three_pass & 32) ? TC_A_LBO : TC_A_SBO;
            const uint32_t blb = (three_pass & 64) ? TC_B_SBO : TC_B_LBO, bsb = (three_pass & 64) ? TC_B_LBO : TC_B_SBO;
            const uint64_t dah = umma_smem_desc(smem_u32(a_hi) + ks * 256, alb, asb);
            const uint64_t dal = umma_smem_desc(smem_u32(a_lo) + ks * 256, alb, asb);
            const uint64_t dbh = umma_smem_desc(smem_u32(b_hi) + ks * 256, blb, bsb);
            const uint64_t dbl = umma_smem_desc(smem_u32(b_lo) + ks * 256, blb, bsb);
            if ((three_pass & 128) && ks == 0) {  // debug: dump descriptors
                C[1] = __uint_as_float((uint32_t)dah); C[2] = __uint_as_float((uint32_t)(dah >> 32));
                C[3] = __uint_as_float((uint32_t)dbh); C[4] = __uint_as_float((uint32_t)(dbh >> 32));
                C[5] = __uint_as_float(idesc); C[6] = __uint_as_float(smem_u32(a_hi));
            }
            if (three_pass & 256) {
                const uint32_t id16 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem),
                    "l"(dah), "l"(dbh), "r"(id16), "r"((uint32_t)(ks > 0)), "r"(0u)
                    : "memory");
                continue;
            }
            if (three_pass & 512) {  // debug: tf32 with B K-major
                umma_tf32(tmem, dah, dbh, idesc & ~(1u << 16), ks > 0);
                continue;
            }
            umma_tf32(tmem, dah, dbh, idesc, (ks > 0) || ((three_pass & 2) && !(three_pass & 16)));
            if (three_pass & 1) {
                umma_tf32(tmem, dal, dbh, idesc, 1);
                umma_tf32(tmem, dah, dbl, idesc, 1);
            }
        }
        umma_commit(bar);
    }
    if (tid == 0 && (three_pass & 4)) mbar_arrive(bar);
    mbar_wait(bar, 0);
    tc_fence_after();
    uint32_t r[32];
    for (int c0 = 0; c0 < n; c0 += 32) {
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        for (int j = 0; j < 32 && c0 + j < n; ++j)
            if (!((three_pass & 128) && warp == 0 && lane == 0))
                C[(size_t)(warp * 32 + lane) * n + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}
