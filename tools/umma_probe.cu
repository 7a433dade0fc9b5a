// tcgen05 probe: validates, on a real B200, the encodings the tensor-core kernel relies on --
// shared-memory matrix descriptors for the un-swizzled K-major "[k/8][row][8]" operand image,
// the kind::f16 instruction descriptor, A-from-TMEM packing, tcgen05.ld/st 32x32b, tcgen05.commit ->
// mbarrier, 1-D bulk (TMA) loads with complete_tx and bulk stores.
//
//   umma_probe <variant>      one variant per process (a fault in one must not poison the next)
//
// Each variant computes D[128,N] = A[128,K] * B[N,K]^T with exactly representable inputs and
// compares with a host result bit for bit.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e = (x);                                                                    \
        if (e != cudaSuccess) {                                                                 \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);     \
            return 2;                                                                           \
        }                                                                                       \
    } while (0)

struct Variant {
    int N, K;
    int swap_lbo_sbo;  // 0: LBO = k-group stride, SBO = 8-row-group stride (expected); 1: swapped
    int a_tmem;        // 1: A operand from TMEM (packed pairs), 0: from smem descriptor
    int fp16;          // 0 bf16, 1 fp16
    int bulk_b;        // 1: B image via cp.async.bulk + mbarrier tx, 0: thread copies
    int bulk_out;      // 1: D staged in smem and written with a bulk store
    int passes;        // number of accumulate passes over K (3 = hi/lo split style accumulate)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;  // descriptor version 1 (Blackwell)
    return d;         // base offset 0, lbo mode 0, layout type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* flag, int code) {
    uint32_t addr = smem_u32(bar);
    for (long long it = 0; it < (1ll << 24); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    if (flag) atomicExch(flag, code);
    return false;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ Bimg, float* __restrict__ D, Variant v, int* flag) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int N = v.N, K = v.K;
    uint16_t* sA = reinterpret_cast<uint16_t*>(smem);
    uint16_t* sB = sA + 128 * K;
    float* sOut = reinterpret_cast<float*>(sB + N * K);  // [128][N] staging for the bulk store
    __shared__ __align__(8) uint64_t bar_mma, bar_tma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int e = tid; e < 128 * K; e += 128) {
        const int m = e / K, k = e - m * K;
        sA[((k >> 3) * 128 + m) * 8 + (k & 7)] = A[e];
    }
    if (!v.bulk_b)
        for (int e = tid; e < N * K; e += 128) sB[e] = Bimg[e];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar_mma)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar_tma)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tm = tmem_base_s;

    if (v.bulk_b) {
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(N * K * 2);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar_tma)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(sB)),
                         "l"(Bimg), "r"(bytes), "r"(smem_u32(&bar_tma))
                         : "memory");
        }
        if (!mbar_wait(&bar_tma, 0, flag, 11)) return;
    }

    const uint32_t tmemA = 256;  // column offset of the packed A operand
    if (v.a_tmem) {
        // thread = row; pack k pairs: low half = even k
        for (int c = 0; c < K / 2; c += 8) {
            uint32_t r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = 2 * (c + i);
                r[i] = (uint32_t)A[tid * K + k] | ((uint32_t)A[tid * K + k + 1] << 16);
            }
            const uint32_t addr = tm + ((uint32_t)(warp * 32) << 16) + tmemA + c;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(addr), "r"(r[0]),
                         "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
    }

    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t idesc = (1u << 4) | ((v.fp16 ? 0u : 1u) << 7) | ((v.fp16 ? 0u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_kg = 128 * 16, b_kg = (uint32_t)N * 16;  // bytes between k-groups of 8
        uint32_t acc = 0;
        for (int pass = 0; pass < v.passes; ++pass) {
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint64_t db = v.swap_lbo_sbo ? make_desc(smem_u32(sB) + ks * 2 * b_kg, 128, b_kg)
                                                   : make_desc(smem_u32(sB) + ks * 2 * b_kg, b_kg, 128);
                if (v.a_tmem) {
                    const uint32_t ta = tm + tmemA + ks * 8;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tm),
                        "r"(ta), "l"(db), "r"(idesc), "r"(acc)
                        : "memory");
                } else {
                    const uint64_t da = v.swap_lbo_sbo ? make_desc(smem_u32(sA) + ks * 2 * a_kg, 128, a_kg)
                                                       : make_desc(smem_u32(sA) + ks * 2 * a_kg, a_kg, 128);
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm),
                        "l"(da), "l"(db), "r"(idesc), "r"(acc)
                        : "memory");
                }
                acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    if (!mbar_wait(&bar_mma, 0, flag, 12)) return;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    for (int c = 0; c < N; c += 8) {
        uint32_t r[8];
        const uint32_t addr = tm + ((uint32_t)(warp * 32) << 16) + c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (v.bulk_out)
                sOut[tid * N + c + i] = __uint_as_float(r[i]);
            else
                D[tid * N + c + i] = __uint_as_float(r[i]);
        }
    }
    if (v.bulk_out) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(D), "r"(smem_u32(sOut)),
                         "r"((uint32_t)(128 * N * 4))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
    (void)lane;
}


// ---------------------------------------------------------------------------------------------
// Timing mode: cycles per tcgen05.mma for a long dependent chain (operand contents irrelevant).
//   umma_probe t <N> <a_tmem> <layout: 0 none, 2 sw128> <alt_d: 1 = alternate two accumulators> <per_commit>
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
time_kernel(int N, int a_tmem, int layout, int alt_d, int per_commit, int nmma, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 48 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u + e;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tm = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 16384;
        uint64_t da, db;
        if (layout == 0) {
            da = make_desc(sA, 2048, 128);
            db = make_desc(sB, (uint32_t)N * 16, 128);
        } else {
            da = make_desc(sA, 16, 1024) | (2ull << 61);
            db = make_desc(sB, 16, 1024) | (2ull << 61);
        }
        uint32_t parity = 0;
        const long long t0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            const uint32_t d = tm + ((alt_d && (i & 1)) ? 256u : 0u);
            // walk through a few k-steps worth of operand addresses like the real kernel does
            const uint32_t step = (uint32_t)(i & 3);
            const uint64_t dbi = db + (layout == 0 ? (uint64_t)((step * 2 * (uint32_t)N * 16) >> 4) : (uint64_t)((step * 32) >> 4));
            if (a_tmem) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                             "r"(tm + 480u), "l"(dbi), "r"(idesc), "r"(1u)
                             : "memory");
            } else {
                const uint64_t dai = da + (layout == 0 ? (uint64_t)((step * 4096) >> 4) : (uint64_t)((step * 32) >> 4));
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                             "l"(dai), "l"(dbi), "r"(idesc), "r"(1u)
                             : "memory");
            }
            if (per_commit && ((i + 1) % per_commit == 0 || i == nmma - 1)) {
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
                if (i == nmma - 1 || per_commit < 0) {
                }
                // only the final commit is waited on; earlier ones just arrive (phase flips are harmless here
                // because we wait for each in order)
                mbar_wait(&bar, parity, nullptr, 0);
                parity ^= 1u;
            }
        }
        if (!per_commit) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
            mbar_wait(&bar, 0, nullptr, 0);
        }
        const long long t1 = clock64();
        out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
}


// Timing v2: the whole warp runs the (warp-uniform) issue loop, only the MMA itself is predicated on one
// elected lane (CUTLASS style) so descriptors can live in the uniform datapath.
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\tselp.u32 %0, 1, 0, px;\n\t}\n" : "=r"(pred) : "r"(0xffffffffu));
    return pred;
}
template <int UNROLL>
__global__ void __launch_bounds__(128, 1)
time_kernel2(int N, int a_tmem, int nmma, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 48 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u + e;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tm = tmem_base_s;
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 16384;
        const uint64_t da = make_desc(sA, 2048, 128);
        const uint64_t db = make_desc(sB, (uint32_t)N * 16, 128);
        const long long t0 = clock64();
        for (int i = 0; i < nmma; i += UNROLL) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const uint32_t step = (uint32_t)(u & 3);
                const uint64_t dbi = db + (uint64_t)((step * 2 * (uint32_t)N * 16) >> 4);
                const uint64_t dai = da + (uint64_t)((step * 4096) >> 4);
                if (elect_one()) {
                    if (a_tmem)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tm),
                                     "r"(tm + 480u), "l"(dbi), "r"(idesc), "r"(1u) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm),
                                     "l"(dai), "l"(dbi), "r"(idesc), "r"(1u) : "memory");
                }
            }
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        mbar_wait(&bar, 0, nullptr, 0);
        const long long t1 = clock64();
        if ((tid & 31) == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
}


// Timing v3: G MMAs per elected block (like the real kernel's 3 per stage), descriptors advanced by adds.
template <int G, int SAME>
__global__ void __launch_bounds__(128, 1)
time_kernel3(int N, int a_tmem, int nmma, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 48 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u + e;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tm = tmem_base_s;
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 16384;
        // SAME == 2: 128-byte-swizzled K-major operands (8-row x 128 B atoms, SBO = 1024 B), k-steps 32 B apart
        const uint64_t da = SAME == 2 ? (make_desc(sA, 16, 1024) | (2ull << 61)) : make_desc(sA, 2048, 128);
        const uint64_t db = SAME == 2 ? (make_desc(sB, 16, 1024) | (2ull << 61)) : make_desc(sB, (uint32_t)N * 16, 128);
        const long long t0 = clock64();
        for (int i = 0; i < nmma; i += G) {
            if (elect_one()) {
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    const uint64_t dbi = SAME == 1 ? db : (SAME == 2 ? db + (uint64_t)((u & 3) * 2) : db + (uint64_t)(u * 64));
                    const uint64_t dai = SAME == 1 ? da : (SAME == 2 ? da + (uint64_t)((u & 3) * 2) : da + (uint64_t)(u * 16));
                    if (a_tmem)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tm),
                                     "r"(tm + 480u), "l"(dbi), "r"(idesc), "r"(1u) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm),
                                     "l"(dai), "l"(dbi), "r"(idesc), "r"(1u) : "memory");
                }
            }
            __syncwarp();
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        mbar_wait(&bar, 0, nullptr, 0);
        const long long t1 = clock64();
        if ((tid & 31) == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
}
template <int G, int SAME>
static int run3(int N, int a_tmem, long long* d) {
    const int nmma = 1536;
    cudaFuncSetAttribute(time_kernel3<G, SAME>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) {
        time_kernel3<G, SAME><<<1, 128, 96 * 1024>>>(N, a_tmem, nmma, d);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
    }
    long long c = 0;
    CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
    printf("timing3 N=%d a_tmem=%d G=%d same=%d : %.1f cycles/MMA (floor N/2 = %d)\n", N, a_tmem, G, SAME, (double)c / nmma, N / 2);
    return 0;
}

static int run_timing(int argc, char** argv) {
    const int N = argc > 2 ? atoi(argv[2]) : 176;
    const int a_tmem = argc > 3 ? atoi(argv[3]) : 0;
    const int layout = argc > 4 ? atoi(argv[4]) : 0;
    const int alt_d = argc > 5 ? atoi(argv[5]) : 0;
    const int per_commit = argc > 6 ? atoi(argv[6]) : 0;
    const int nmma = 2048;
    long long* d;
    CK(cudaMalloc(&d, 8));
    cudaFuncSetAttribute(time_kernel2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(time_kernel2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(time_kernel2<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (argv[1][1] == '3') {
        for (int N : {64, 144, 240})
            for (int at = 0; at < 2; ++at) {
                run3<1, 0>(N, at, d); run3<3, 0>(N, at, d); run3<6, 0>(N, at, d); run3<12, 0>(N, at, d); run3<6, 1>(N, at, d); run3<6, 2>(N, at, d);
            }
        return 0;
    }
    if (argv[1][1] == '2') {
        const int unroll = argc > 4 ? atoi(argv[4]) : 1;
        for (int rep = 0; rep < 2; ++rep) {
            if (unroll == 1) time_kernel2<1><<<1, 128, 96 * 1024>>>(N, a_tmem, nmma, d);
            else if (unroll == 4) time_kernel2<4><<<1, 128, 96 * 1024>>>(N, a_tmem, nmma, d);
            else time_kernel2<8><<<1, 128, 96 * 1024>>>(N, a_tmem, nmma, d);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
        }
        long long c2 = 0;
        CK(cudaMemcpy(&c2, d, 8, cudaMemcpyDeviceToHost));
        printf("timing2 N=%d a_tmem=%d unroll=%d : %.1f cycles/MMA (floor N/2 = %d)\n", N, a_tmem, unroll, (double)c2 / nmma, N / 2);
        return 0;
    }
    CK(cudaFuncSetAttribute(time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(time_kernel2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(time_kernel2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(time_kernel2<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int rep = 0; rep < 2; ++rep) {
        time_kernel<<<1, 128, 96 * 1024>>>(N, a_tmem, layout, alt_d, per_commit, nmma, d);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
    }
    long long cyc = 0;
    CK(cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost));
    printf("timing N=%d a_tmem=%d layout=%d alt_d=%d per_commit=%d : %.1f cycles/MMA (floor N/2 = %d)\n", N, a_tmem, layout, alt_d,
           per_commit, (double)cyc / nmma, N / 2);
    return 0;
}


// ---------------------------------------------------------------------------------------------
// Cluster probe: can a 1-D bulk copy issued by CTA 1 into ITS OWN shared memory signal (complete_tx) a
// barrier that lives in CTA 0?  And what does a remote mbarrier arrive cost?
//   umma_probe c
// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1)
cluster_kernel(const uint32_t* __restrict__ src, uint32_t* out, long long* cyc, int mode) {
    __shared__ __align__(128) uint32_t buf[1024];
    __shared__ __align__(8) uint64_t bar;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
    const int tid = threadIdx.x;
    for (int i = tid; i < 1024; i += 64) buf[i] = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (mode == 0) {
        // CTA 0 arms its barrier with the byte count; CTA 1 copies into its own buffer, signalling CTA 0's barrier
        if (rank == 0 && tid == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(4096u) : "memory");
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        if (rank == 1 && tid == 0) {
            uint32_t rbar;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(rbar) : "r"(smem_u32(&bar)), "r"(0u));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(buf)),
                         "l"(src), "r"(4096u), "r"(rbar)
                         : "memory");
        }
        if (rank == 0) {
            if (tid == 0) {
                int flag = 0;
                const bool ok = mbar_wait(&bar, 0, &flag, 1);
                out[2048] = ok ? 1u : 0u;
            }
        }
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        if (rank == 1) for (int i = tid; i < 1024; i += 64) out[i] = buf[i];
    } else {
        // remote arrive cost: CTA 1 arrives N times on CTA 0's barrier (count 1 -> a phase per arrive), CTA 0 waits each
        const int N = 256;
        if (rank == 1 && tid == 0) {
            uint32_t rbar;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(rbar) : "r"(smem_u32(&bar)), "r"(0u));
            const long long t0 = clock64();
            for (int i = 0; i < N; ++i) {
                if (mode == 1) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(rbar) : "memory");
                else asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(rbar) : "memory");
                // wait until CTA 0 consumed it (it bumps a flag in my smem) to avoid phase overrun
                while (atomicAdd(&buf[0], 0) != (uint32_t)(i + 1)) {}
            }
            cyc[0] = clock64() - t0;
        }
        if (rank == 0 && tid == 0) {
            uint32_t rflag;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(rflag) : "r"(smem_u32(&buf[0])), "r"(1u));
            for (int i = 0; i < N; ++i) {
                mbar_wait(&bar, i & 1, nullptr, 0);
                asm volatile("st.shared::cluster.u32 [%0], %1;\n" ::"r"(rflag), "r"((uint32_t)(i + 1)) : "memory");
            }
        }
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    }
}
static int run_cluster() {
    uint32_t *dsrc, *dout;
    long long* dc;
    std::vector<uint32_t> h(1024);
    for (int i = 0; i < 1024; ++i) h[i] = 0xabc00000u + i;
    CK(cudaMalloc(&dsrc, 4096));
    CK(cudaMalloc(&dout, 4 * 4096));
    CK(cudaMalloc(&dc, 8));
    CK(cudaMemcpy(dsrc, h.data(), 4096, cudaMemcpyHostToDevice));
    // mode 0 (bulk copy signalling a barrier of the OTHER CTA) faults with 'unspecified launch failure' on B200:
    // the mbarrier of cp.async.bulk must live in the destination CTA.  Kept for the record, not run by default.
    for (int mode = (getenv("PROBE_REMOTE_TX") ? 0 : 1); mode < 3; ++mode) {
        CK(cudaMemset(dout, 0, 4 * 4096));
        cluster_kernel<<<2, 64>>>(dsrc, dout, dc, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("cluster mode %d: CUDA error %s\n", mode, cudaGetErrorString(e));
            return 2;
        }
        if (mode == 0) {
            std::vector<uint32_t> o(2049);
            CK(cudaMemcpy(o.data(), dout, 2049 * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int i = 0; i < 1024; ++i) bad += (o[i] != h[i]);
            printf("cluster: bulk copy by CTA1 into own smem signalling CTA0's barrier: wait_ok=%u data_mismatch=%d  %s\n", o[2048], bad,
                   (o[2048] == 1 && bad == 0) ? "PASS" : "FAIL");
        } else {
            long long c = 0;
            CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
            printf("cluster: remote arrive (%s) + consume round trip: %.0f cycles\n", mode == 1 ? "release" : "relaxed", (double)c / 256);
        }
    }
    return 0;
}


// ---------------------------------------------------------------------------------------------
// kind::f8f6f4 (e4m3 x e4m3, K = 32 per instruction) probe: validates the un-swizzled K-major operand image for
// 8-bit elements ([k/16][row][16 x 8 bit]), A from TMEM (four elements per 32-bit column, byte 0 = lowest k) and
// -- with mixed = 1 -- accumulating kind::f16 and kind::f8f6f4 MMAs into the SAME fp32 accumulator.
//   umma_probe f <N> <K> <a_tmem> <mixed>
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
probe_f8_kernel(const uint8_t* __restrict__ A8, const uint8_t* __restrict__ B8img, const uint16_t* __restrict__ A16,
                const uint16_t* __restrict__ B16img, float* __restrict__ D, int N, int K, int a_tmem, int mixed, int* flag) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA8 = smem;                                           // [K/16][128][16]
    uint8_t* sB8 = sA8 + 128 * K;                                  // [K/16][N][16]
    uint16_t* sA16 = reinterpret_cast<uint16_t*>(sB8 + N * K);     // [K/8][128][8]
    uint16_t* sB16 = sA16 + 128 * K;                               // [K/8][N][8]
    __shared__ __align__(8) uint64_t bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 128 * K; e += 128) {
        const int m = e / K, k = e - m * K;
        sA8[((k >> 4) * 128 + m) * 16 + (k & 15)] = A8[e];
        sA16[((k >> 3) * 128 + m) * 8 + (k & 7)] = A16[e];
    }
    for (int e = tid; e < N * K; e += 128) {
        sB8[e] = B8img[e];
        sB16[e] = B16img[e];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar_mma)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tm = tmem_base_s;
    const uint32_t tmemA8 = 256, tmemA16 = 256 + 64;  // K <= 256: 64 columns of 8-bit, then K/2 columns of 16-bit
    if (a_tmem) {
        for (int c = 0; c < K / 4; c += 8) {
            uint32_t r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = 4 * (c + i);
                r[i] = (uint32_t)A8[tid * K + k] | ((uint32_t)A8[tid * K + k + 1] << 8) | ((uint32_t)A8[tid * K + k + 2] << 16) |
                       ((uint32_t)A8[tid * K + k + 3] << 24);
            }
            const uint32_t addr = tm + ((uint32_t)(warp * 32) << 16) + tmemA8 + c;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(addr), "r"(r[0]),
                         "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        }
        for (int c = 0; c < K / 2; c += 8) {
            uint32_t r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = 2 * (c + i);
                r[i] = (uint32_t)A16[tid * K + k] | ((uint32_t)A16[tid * K + k + 1] << 16);
            }
            const uint32_t addr = tm + ((uint32_t)(warp * 32) << 16) + tmemA16 + c;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(addr), "r"(r[0]),
                         "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
    }
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t idesc8 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // e4m3 x e4m3 -> f32
        const uint32_t idesc16 = idesc8;                                                               // f16 x f16 -> f32
        uint32_t acc = 0;
        if (mixed) {  // 16-bit pass first, the 8-bit MMAs then accumulate on top
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint64_t db = make_desc(smem_u32(sB16) + ks * 2 * N * 16, (uint32_t)N * 16, 128);
                if (a_tmem) {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tm),
                                 "r"(tm + tmemA16 + ks * 8), "l"(db), "r"(idesc16), "r"(acc) : "memory");
                } else {
                    const uint64_t da = make_desc(smem_u32(sA16) + ks * 2 * 2048, 2048, 128);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm),
                                 "l"(da), "l"(db), "r"(idesc16), "r"(acc) : "memory");
                }
                acc = 1;
            }
        }
        for (int ks = 0; ks < K / 32; ++ks) {
            const uint64_t db = make_desc(smem_u32(sB8) + ks * 2 * N * 16, (uint32_t)N * 16, 128);
            if (a_tmem) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tm),
                             "r"(tm + tmemA8 + ks * 8), "l"(db), "r"(idesc8), "r"(acc) : "memory");
            } else {
                const uint64_t da = make_desc(smem_u32(sA8) + ks * 2 * 2048, 2048, 128);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm),
                             "l"(da), "l"(db), "r"(idesc8), "r"(acc) : "memory");
            }
            acc = 1;
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    if (!mbar_wait(&bar_mma, 0, flag, 12)) return;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    for (int c = 0; c < N; c += 8) {
        uint32_t r[8];
        const uint32_t addr = tm + ((uint32_t)(warp * 32) << 16) + c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) D[tid * N + c + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
}

// e4m3 encoding of a multiple of 1/8 in [-1, 1] (exact): sign | exponent (bias 7) | 3 mantissa bits
static uint8_t to_e4m3_exact(float x) {
    if (x == 0.f) return 0;
    const uint8_t sgn = x < 0 ? 0x80 : 0;
    float a = fabsf(x);
    int e = 0;
    while (a < 1.f) { a *= 2.f; --e; }
    while (a >= 2.f) { a *= 0.5f; ++e; }
    const int mant = (int)((a - 1.f) * 8.f + 0.5f);
    return sgn | (uint8_t)((e + 7) << 3) | (uint8_t)mant;
}
static uint16_t to16(float x, int fp16);
static int run_f8(int argc, char** argv) {
    const int N = argc > 2 ? atoi(argv[2]) : 64, K = argc > 3 ? atoi(argv[3]) : 32, a_tmem = argc > 4 ? atoi(argv[4]) : 0,
              mixed = argc > 5 ? atoi(argv[5]) : 0;
    std::vector<float> A8f(128 * K), B8f(N * K), A16f(128 * K), B16f(N * K);
    uint32_t s = 777u + N + K;
    auto rnd = [&]() {
        s = s * 1664525u + 1013904223u;
        return (float)((int)((s >> 16) % 17) - 8) / 8.0f;
    };
    for (auto& x : A8f) x = rnd();
    for (auto& x : B8f) x = rnd();
    for (auto& x : A16f) x = rnd();
    for (auto& x : B16f) x = rnd();
    std::vector<uint8_t> A8(128 * K), B8(N * K);
    std::vector<uint16_t> A16(128 * K), B16(N * K);
    for (int i = 0; i < 128 * K; ++i) { A8[i] = to_e4m3_exact(A8f[i]); A16[i] = to16(A16f[i], 1); }
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            B8[((k >> 4) * N + n) * 16 + (k & 15)] = to_e4m3_exact(B8f[n * K + k]);
            B16[((k >> 3) * N + n) * 8 + (k & 7)] = to16(B16f[n * K + k], 1);
        }
    std::vector<float> ref(128 * N);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            float acc = 0;
            for (int k = 0; k < K; ++k) acc += A8f[m * K + k] * B8f[n * K + k] + (mixed ? A16f[m * K + k] * B16f[n * K + k] : 0.f);
            ref[m * N + n] = acc;
        }
    uint8_t *dA8, *dB8;
    uint16_t *dA16, *dB16;
    float* dD;
    int* dflag;
    CK(cudaMalloc(&dA8, A8.size())); CK(cudaMalloc(&dB8, B8.size()));
    CK(cudaMalloc(&dA16, A16.size() * 2)); CK(cudaMalloc(&dB16, B16.size() * 2));
    CK(cudaMalloc(&dD, ref.size() * 4)); CK(cudaMalloc(&dflag, 4));
    CK(cudaMemset(dflag, 0, 4)); CK(cudaMemset(dD, 0xff, ref.size() * 4));
    CK(cudaMemcpy(dA8, A8.data(), A8.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB8, B8.data(), B8.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA16, A16.data(), A16.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB16, B16.data(), B16.size() * 2, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)3 * (128 + N) * K + 1024;
    CK(cudaFuncSetAttribute(probe_f8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    probe_f8_kernel<<<1, 128, smem>>>(dA8, dB8, dA16, dB16, dD, N, K, a_tmem, mixed, dflag);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> out(ref.size());
    int flag = 0;
    CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (size_t i = 0; i < ref.size(); ++i) bad += !(out[i] == ref[i]);
    printf("f8 probe N=%d K=%d a_tmem=%d mixed=%d : flag=%d mismatches=%d/%zu  %s\n", N, K, a_tmem, mixed, flag, bad, ref.size(),
           (bad == 0 && flag == 0) ? "PASS" : "FAIL");
    int shown = 0;
    for (size_t i = 0; i < ref.size() && shown < 6 && bad; ++i)
        if (out[i] != ref[i]) { printf("   [m=%zu n=%zu] got %g want %g\n", i / N, i % N, out[i], ref[i]); ++shown; }
    return (bad == 0 && flag == 0) ? 0 : 3;
}

// ---------------------------------------------------------------------------------------------
// Timing v4: cycles per MMA for kind::f16 (K = 16) vs kind::f8f6f4 (K = 32), one CTA or a CTA pair (cta_group::2, M = 256).
//   umma_probe u
// ---------------------------------------------------------------------------------------------
template <int CG, int KIND, int G, int MROWS = 128>  // MROWS: accumulator rows per CTA (128, or 64 = half-height MMAs)
__global__ void __launch_bounds__(128, 1)
time_kernel4(int N, int a_tmem, int nmma, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
    for (int e = tid; e < 64 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = KIND ? 0x38383838u : 0x3c003c00u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tm = tmem_base_s;
    if (warp == 0 && rank == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((CG * MROWS) >> 4) << 24);
        const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 32768;
        const uint32_t b_kg = (uint32_t)(N / CG) * 16;
        const uint64_t da = make_desc(sA, 2048, 128);
        const uint64_t db = make_desc(sB, b_kg, 128);
        const long long t0 = clock64();
        for (int i = 0; i < nmma; i += G) {
            if (elect_one()) {
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    const uint64_t dbi = db + (uint64_t)(((u & 3) * 2 * b_kg) >> 4);
                    const uint64_t dai = da + (uint64_t)(((u & 3) * 4096) >> 4);
                    const uint32_t ta = tm + 448u + (u & 3) * 8;
#define MMA4(cg, kind)                                                                                                                  \
    if (a_tmem)                                                                                                                          \
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::" cg ".kind::" kind " [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tm), \
                     "r"(ta), "l"(dbi), "r"(idesc), "r"(1u) : "memory");                                                               \
    else                                                                                                                                 \
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::" cg ".kind::" kind " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm),   \
                     "l"(dai), "l"(dbi), "r"(idesc), "r"(1u) : "memory");
                    if (CG == 1 && KIND == 0) { MMA4("1", "f16") }
                    if (CG == 1 && KIND == 1) { MMA4("1", "f8f6f4") }
                    if (CG == 2 && KIND == 0) { MMA4("2", "f16") }
                    if (CG == 2 && KIND == 1) { MMA4("2", "f8f6f4") }
                }
            }
            __syncwarp();
        }
        if (elect_one()) {
            if (CG == 2)
                asm volatile("{\n\t.reg .b16 m;\n\tmov.b16 m, 1;\n\ttcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}\n" ::"r"(smem_u32(&bar)) : "memory");
            else
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        mbar_wait(&bar, 0, nullptr, 0);
        const long long t1 = clock64();
        if ((tid & 31) == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (warp == 0) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
    }
}
template <int CG, int KIND, int MROWS = 128>
static int run4(int N, int a_tmem, long long* d) {
    const int nmma = 1536;
    cudaFuncSetAttribute(time_kernel4<CG, KIND, 6, MROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(CG);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = 96 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CG;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, time_kernel4<CG, KIND, 6, MROWS>, N, a_tmem, nmma, d));
        CK(cudaDeviceSynchronize());
    }
    long long c = 0;
    CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
    printf("timing4 cta_group=%d rows/CTA=%d kind=%s N=%d a_tmem=%d : %.1f cycles/MMA (N/2 = %d)\n", CG, MROWS, KIND ? "f8f6f4(K32)" : "f16(K16)", N,
           a_tmem, (double)c / nmma, N / 2);
    return 0;
}
static int run_timing4(int half_height) {
    long long* d;
    CK(cudaMalloc(&d, 8));
    if (half_height) {  // 64 accumulator rows per CTA: does the MMA take half the time?
        for (int N : {64, 128, 176, 256})
            for (int at = 0; at < 2; ++at) {
                run4<1, 0, 64>(N, at, d);
                run4<2, 0, 64>(N, at, d);
                run4<1, 1, 64>(N, at, d);
            }
        return 0;
    }
    for (int N : {112, 144, 176, 192, 224, 256})
        for (int at = 0; at < 2; ++at) {
            run4<1, 0>(N, at, d);
            run4<1, 1>(N, at, d);
            run4<2, 0>(N, at, d);
            run4<2, 1>(N, at, d);
        }
    return 0;
}


// ---------------------------------------------------------------------------------------------
// Bulk-tensor (TMA) store probe: an [n, 451] fp32 array described as 4-row super-rows (dim0 = 1804, pitch 7216 B);
// one box of (W, 8) at element coordinate (x, y) with x NOT a multiple of 4.
//   umma_probe s <x> <W>
// ---------------------------------------------------------------------------------------------
#include <cuda.h>
__global__ void tma_store_kernel(const __grid_constant__ CUtensorMap map, int x, int y, int W) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* st = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < 8 * W; i += blockDim.x) st[i] = 1000.f * (i / W) + (i % W);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(reinterpret_cast<uint64_t>(&map)),
                     "r"(smem_u32(smem)), "r"(x), "r"(y) : "memory");
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    }
}
static int run_tma_store(int argc, char** argv) {
    const int x = argc > 2 ? atoi(argv[2]) : 451 + 16, W = argc > 3 ? atoi(argv[3]) : 112, n = 64, NO = 451, y = 2;
    float* d;
    CK(cudaMalloc(&d, (size_t)n * NO * 4));
    CK(cudaMemset(d, 0, (size_t)n * NO * 4));
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap map;
    const cuuint64_t dims[2] = {4ull * NO, (cuuint64_t)(n / 4)};
    const cuuint64_t strides[1] = {16ull * NO};
    const cuuint32_t box[2] = {(cuuint32_t)W, 8u};
    const cuuint32_t es[2] = {1u, 1u};
    CUresult r = ((Fn)fnp)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode result %d (query %d)\n", (int)r, (int)q);
    if (r != CUDA_SUCCESS) return 3;
    tma_store_kernel<<<1, 128, 8 * W * 4 + 1024>>>(map, x, y, W);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tma store x=%d W=%d: CUDA error %s\n", x, W, cudaGetErrorString(e)); return 2; }
    std::vector<float> h((size_t)n * NO);
    CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0, nz = 0;
    for (int row = 0; row < n; ++row)
        for (int c = 0; c < NO; ++c) {
            // element (x', y') of the super-row view: row = 4 y' + x' / NO, col = x' % NO
            float want = 0.f;
            const int yy = row / 4, xx = (row % 4) * NO + c;
            if (yy >= y && yy < y + 8 && xx >= x && xx < x + W) want = 1000.f * (yy - y) + (xx - x);
            const float got = h[(size_t)row * NO + c];
            bad += (got != want);
            nz += (got != 0.f);
        }
    printf("tma store x=%d W=%d: nonzero=%d mismatches=%d  %s\n", x, W, nz, bad, bad == 0 ? "PASS" : "FAIL");
    return bad ? 3 : 0;
}


// ---------------------------------------------------------------------------------------------
// Hand-off latency probe: warp 0 signals barrier X (mode 0: mbarrier.arrive, 1: tcgen05.commit with no MMA pending,
// 2: tcgen05.commit right after one MMA N=64), warp 1 waits for X and arrives on Y, warp 0 waits for Y.  Cycles per round trip.
//   umma_probe p
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
pingpong_kernel(int mode, int rounds, int fences, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t barx, bary;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < 32 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&barx)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bary)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tm = tmem_base_s;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = make_desc(smem_u32(smem), 2048, 128), db = make_desc(smem_u32(smem) + 16384, 64 * 16, 128);
    if (warp == 0) {
        const long long t0 = clock64();
        for (int r = 0; r < rounds; ++r) {
            if (fences) asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (elect_one()) {
                if (mode == 2)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm),
                                 "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                if (mode == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&barx)) : "memory");
                else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&barx)) : "memory");
            }
            __syncwarp();
            mbar_wait(&bary, r & 1, nullptr, 0);
        }
        if (lane == 0) out[0] = clock64() - t0;
    } else if (warp == 1) {
        for (int r = 0; r < rounds; ++r) {
            mbar_wait(&barx, r & 1, nullptr, 0);
            if (fences) {
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&bary)) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
}
static int run_pingpong() {
    long long* d;
    CK(cudaMalloc(&d, 8));
    CK(cudaFuncSetAttribute(pingpong_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int fences = 0; fences < 2; ++fences)
        for (int mode = 0; mode < 3; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                pingpong_kernel<<<1, 128, 48 * 1024>>>(mode, 1000, fences, d);
                CK(cudaDeviceSynchronize());
            }
            long long c = 0;
            CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
            printf("pingpong mode=%d (%s) fences=%d : %.0f cycles per round trip\n", mode,
                   mode == 0 ? "arrive" : mode == 1 ? "commit, pipe idle" : "MMA N=64 + commit", fences, (double)c / 1000);
        }
    return 0;
}


// ---------------------------------------------------------------------------------------------
// Layout probe for HALF-HEIGHT pair MMAs (cta_group::2, M = 128: 64 accumulator rows per CTA) -- groundwork for running two
// half-height tiles per CTA.  A[m][k] = m + 128 k (exact in fp16), B = identity on the first 16 columns, so D[m][n] = A[m][n] names
// its own (m, n); TMEM is pre-filled with -1.  Prints, per CTA, which accumulator row every TMEM lane holds.
//   umma_probe h <d_lane_offset: 0 or 64>
// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
half_height_kernel(float* __restrict__ out /*[2][128][32]*/, int d_lane) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
    constexpr int K = 16, N = 32;
    uint16_t* sA = reinterpret_cast<uint16_t*>(smem);           // [K/8][64][8]
    uint16_t* sB = reinterpret_cast<uint16_t*>(smem + 4096);    // [K/8][N/2][8]: this CTA's rows n = 16 rank .. 16 rank + 15
    for (int e = tid; e < 64 * K; e += 128) {
        const int m = e / K, k = e % K;
        const float v = (float)(64 * (int)rank + m + 128 * k);
        __half h = __float2half_rn(v);
        sA[((k >> 3) * 64 + m) * 8 + (k & 7)] = *reinterpret_cast<uint16_t*>(&h);
    }
    for (int e = tid; e < (N / 2) * K; e += 128) {
        const int nl = e / K, k = e % K, n = (N / 2) * (int)rank + nl;
        __half h = __float2half_rn((n < 16 && n == k) ? 1.f : 0.f);
        sB[((k >> 3) * (N / 2) + nl) * 8 + (k & 7)] = *reinterpret_cast<uint16_t*>(&h);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tm = tmem_base_s;
    {   // sentinel fill: every lane, columns 0..31
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(-1.f);
        for (int c = 0; c < N; c += 8) {
            const uint32_t addr = tm + ((uint32_t)(warp * 32) << 16) + c;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(addr), "r"(r[0]), "r"(r[1]),
                         "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        }
        // A operand candidate in TMEM columns 64..71 (k pairs): lane L, column c holds fp16 (L + 128 (2c)) | fp16 (L + 128 (2c + 1)) << 16
        uint32_t ar[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            __half lo = __float2half_rn((float)(tid + 128 * (2 * c))), hi = __float2half_rn((float)(tid + 128 * (2 * c + 1)));
            ar[c] = (uint32_t)(*reinterpret_cast<uint16_t*>(&lo)) | ((uint32_t)(*reinterpret_cast<uint16_t*>(&hi)) << 16);
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(tm + ((uint32_t)(warp * 32) << 16) + 64u),
                     "r"(ar[0]), "r"(ar[1]), "r"(ar[2]), "r"(ar[3]), "r"(ar[4]), "r"(ar[5]), "r"(ar[6]), "r"(ar[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (rank == 0 && tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32, M = 128 over the pair
        const uint64_t da = make_desc(smem_u32(sA), 64 * 16, 128);
        const uint64_t db = make_desc(smem_u32(sB), (N / 2) * 16, 128);
        const uint32_t d = tm + ((uint32_t)(d_lane % 1000) << 16);
        if (d_lane >= 1000)  // A from TMEM columns 64..71: every lane holds lane + 128 (2 col + half), so D[m][k] names its source
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                         "r"(tm + 64u), "l"(db), "r"(idesc), "r"(0u) : "memory");
        else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da),
                     "l"(db), "r"(idesc), "r"(0u) : "memory");
        asm volatile("{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\ttcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}\n" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0, nullptr, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    for (int c = 0; c < N; c += 8) {
        uint32_t r[8];
        const uint32_t addr = tm + ((uint32_t)(warp * 32) << 16) + c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) out[((size_t)rank * 128 + tid) * N + c + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
}
static int run_half_height(int argc, char** argv) {
    const int d_lane = argc > 2 ? atoi(argv[2]) : 0;
    float* d;
    CK(cudaMalloc(&d, 2 * 128 * 32 * 4));
    CK(cudaMemset(d, 0, 2 * 128 * 32 * 4));
    half_height_kernel<<<2, 128, 16 * 1024>>>(d, d_lane);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("half-height probe d_lane=%d: CUDA error %s\n", d_lane, cudaGetErrorString(e)); return 2; }
    std::vector<float> h(2 * 128 * 32);
    CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
    for (int r = 0; r < 2; ++r) {
        printf("CTA %d (d_lane %d): TMEM lane -> accumulator row m held in column 0 (x = untouched), then column -> n for the first written lane\n", r, d_lane);
        int first = -1;
        for (int lane = 0; lane < 128; ++lane) {
            const float v = h[((size_t)r * 128 + lane) * 32 + 0];
            if (v == -1.f) printf(" x");
            else { printf(" %d", (int)v % 128); if (first < 0) first = lane; }  // SS: row m; TS: source TMEM lane of A
            if (lane % 32 == 31) printf("\n");
        }
        if (first >= 0) {
            printf("  lane %d columns:", first);
            for (int c = 0; c < 32; ++c) {
                const float v = h[((size_t)r * 128 + first) * 32 + c];
                if (v == -1.f) printf(" x"); else printf(" k%d", (int)v / 128);
            }
            printf("\n");
        }
    }
    return 0;
}

static uint16_t to16(float x, int fp16) {
    if (fp16) {
        __half h = __float2half_rn(x);
        uint16_t u;
        memcpy(&u, &h, 2);
        return u;
    }
    __nv_bfloat16 b = __float2bfloat16_rn(x);
    uint16_t u;
    memcpy(&u, &b, 2);
    return u;
}

int main(int argc, char** argv) {
    const Variant table[] = {
        // N,   K, swap, a_tmem, fp16, bulk_b, bulk_out, passes
        {64, 16, 0, 0, 0, 0, 0, 1},    // 0 smallest: one MMA, expected LBO/SBO
        {64, 16, 1, 0, 0, 0, 0, 1},    // 1 swapped LBO/SBO
        {176, 64, 0, 0, 0, 0, 0, 1},   // 2 four k-steps, N = 176
        {176, 64, 0, 0, 0, 1, 0, 1},   // 3 + B via bulk copy
        {176, 64, 0, 1, 0, 1, 0, 1},   // 4 A from TMEM
        {80, 352, 0, 1, 0, 1, 0, 3},   // 5 TS, K = 352, 3 accumulate passes
        {224, 288, 0, 0, 1, 1, 1, 3},  // 6 SS fp16, N = 224, bulk store of D
        {144, 16, 0, 0, 0, 1, 0, 1},   // 7 N = 144 (layer-0 style)
        {176, 64, 1, 0, 0, 0, 0, 1},   // 8 swapped on a bigger case
        {128, 224, 0, 1, 1, 1, 1, 3},  // 9 TS fp16 + bulk store
    };
    const int nvar = sizeof(table) / sizeof(table[0]);
    if (argc < 2) {
        printf("%d\n", nvar);
        return 0;
    }
    if (argv[1][0] == 'c') return run_cluster();
    if (argv[1][0] == 't') return run_timing(argc, argv);
    if (argv[1][0] == 'f') return run_f8(argc, argv);
    if (argv[1][0] == 'u') return run_timing4(argc > 2 ? atoi(argv[2]) : 0);
    if (argv[1][0] == 's') return run_tma_store(argc, argv);
    if (argv[1][0] == 'p') return run_pingpong();
    if (argv[1][0] == 'h') return run_half_height(argc, argv);
    const int vi = atoi(argv[1]);
    if (vi < 0 || vi >= nvar) return 1;
    const Variant v = table[vi];
    const int N = v.N, K = v.K;
    std::vector<float> Af(128 * K), Bf(N * K);
    std::vector<uint16_t> A(128 * K), Bimg(N * K);
    uint32_t s = 12345u + vi;
    auto rnd = [&]() {
        s = s * 1664525u + 1013904223u;
        return (float)((int)((s >> 16) % 17) - 8) / 8.0f;  // multiples of 1/8 in [-1, 1]
    };
    for (auto& x : Af) x = rnd();
    for (auto& x : Bf) x = rnd();
    for (int i = 0; i < 128 * K; ++i) A[i] = to16(Af[i], v.fp16);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) Bimg[((k >> 3) * N + n) * 8 + (k & 7)] = to16(Bf[n * K + k], v.fp16);
    std::vector<float> ref(128 * N, 0.f);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            float acc = 0;
            for (int k = 0; k < K; ++k) acc += Af[m * K + k] * Bf[n * K + k];
            ref[m * N + n] = acc * v.passes;
        }
    uint16_t *dA, *dB;
    float* dD;
    int* dflag;
    CK(cudaMalloc(&dA, A.size() * 2));
    CK(cudaMalloc(&dB, Bimg.size() * 2));
    CK(cudaMalloc(&dD, ref.size() * 4));
    CK(cudaMalloc(&dflag, 4));
    CK(cudaMemset(dflag, 0, 4));
    CK(cudaMemset(dD, 0xff, ref.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bimg.data(), Bimg.size() * 2, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)128 * K * 2 + (size_t)N * K * 2 + (v.bulk_out ? (size_t)128 * N * 4 : 0) + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    if (smem > 226 * 1024) {
        printf("variant %d: smem %zu too large\n", vi, smem);
        return 1;
    }
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, v, dflag);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> out(ref.size());
    int flag = 0;
    CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    int bad = 0;
    for (size_t i = 0; i < ref.size(); ++i) {
        double e = fabs((double)out[i] - ref[i]);
        if (!(e <= 0)) ++bad;
        if (e > maxerr || e != e) maxerr = e != e ? 1e30 : e;
    }
    printf("variant %d N=%d K=%d swap=%d a_tmem=%d fp16=%d bulk_b=%d bulk_out=%d passes=%d : flag=%d mismatches=%d/%zu maxerr=%g  %s\n", vi, N,
           K, v.swap_lbo_sbo, v.a_tmem, v.fp16, v.bulk_b, v.bulk_out, v.passes, flag, bad, ref.size(), maxerr,
           (bad == 0 && flag == 0) ? "PASS" : "FAIL");
    if (bad) {
        int shown = 0;
        for (size_t i = 0; i < ref.size() && shown < 6; ++i)
            if (out[i] != ref[i]) {
                printf("   [m=%zu n=%zu] got %g want %g\n", i / N, i % N, out[i], ref[i]);
                ++shown;
            }
    }
    return (bad == 0 && flag == 0) ? 0 : 3;
}
