// Round-2 hardware probes for the tensor-core kernel redesign (B200, sm_100a).  Stand-alone binary, not product code.
//
//   tc_probe2 m            cycles per tcgen05.mma for operand layouts {no swizzle, SW128, SW64, SW32} x A source {smem, TMEM}
//                          x cta_group {1, 2} x accumulator rows per CTA {128, 64}
//   tc_probe2 l            L2 -> shared-memory streaming bandwidth of 1-D bulk copies through a ring, all SMs, with and
//                          without cluster multicast (the weight ring of the fused kernel)
//   tc_probe2 o            how many clusters of 2 / 4 / 8 CTAs (1 CTA per SM, ~200 KB of shared memory) the GPU co-schedules
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e = (x);                                                                \
        if (e != cudaSuccess) {                                                             \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            return 2;                                                                       \
        }                                                                                   \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\tselp.u32 %0, 1, 0, px;\n\t}\n" : "=r"(pred) : "r"(0xffffffffu));
    return pred;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t addr, uint32_t parity) {
    for (long long it = 0; it < (1ll << 22); ++it) {
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
    return false;
}

// ---------------------------------------------------------------------------------------------
// m: MMA timing.  layout: 0 none (K-major core matrices, LBO = k-group stride, SBO = 128), 1 SW128, 2 SW64, 3 SW32
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | ((uint64_t)layout_type << 61);
}

template <int CG>
__global__ void __launch_bounds__(128, 1)
mma_time_kernel(int N, int mrows, int a_tmem, int layout, int alt_d, int nmma, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank = 0;
    if (CG == 2) rank = cluster_ctarank();
    for (int e = tid; e < 96 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u;
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
    if (CG == 2) cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tm = tmem_base_s;
    if (warp == 0 && rank == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((CG * mrows) >> 4) << 24);
        const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 32768;
        const uint32_t brows = (uint32_t)(N / CG);  // B rows held by this CTA
        // per-layout descriptor + the byte advance from one k-step (16 elements) to the next, cycling over 4 k-steps
        uint64_t da, db;
        uint32_t a_step[4], b_step[4];
        if (layout == 0) {
            da = make_desc(sA, (uint32_t)mrows * 16, 128, 0);
            db = make_desc(sB, brows * 16, 128, 0);
            for (int u = 0; u < 4; ++u) { a_step[u] = u * (uint32_t)mrows * 32; b_step[u] = u * brows * 32; }
        } else {
            const uint32_t sw = layout == 1 ? 128u : layout == 2 ? 64u : 32u;   // swizzle span in bytes
            const uint32_t lt = layout == 1 ? 2u : layout == 2 ? 4u : 6u;
            da = make_desc(sA, 16, 8 * sw, lt);
            db = make_desc(sB, 16, 8 * sw, lt);
            const uint32_t per = sw / 32;  // k-steps inside one swizzle span
            for (int u = 0; u < 4; ++u) {
                a_step[u] = (u % per) * 32 + (u / per) * (uint32_t)mrows * sw;
                b_step[u] = (u % per) * 32 + (u / per) * brows * sw;
                if (a_step[u] + (uint32_t)mrows * sw > 32768) a_step[u] = (u % per) * 32;
                if (b_step[u] + brows * sw > 60 * 1024) b_step[u] = (u % per) * 32;
            }
        }
        const long long t0 = clock64();
        for (int i = 0; i < nmma; i += 6) {
            if (elect_one()) {
#pragma unroll
                for (int u = 0; u < 6; ++u) {
                    const uint64_t dbi = db + (uint64_t)(b_step[u & 3] >> 4);
                    const uint64_t dai = da + (uint64_t)(a_step[u & 3] >> 4);
                    const uint32_t ta = tm + 448u + (u & 3) * 8;
                    const uint32_t d = tm + (alt_d ? (uint32_t)((u & 1) * (mrows == 64 ? N / 2 : N)) : 0u);
                    if (CG == 1) {
                        if (a_tmem)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                                         "r"(ta), "l"(dbi), "r"(idesc), "r"(1u) : "memory");
                        else
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                                         "l"(dai), "l"(dbi), "r"(idesc), "r"(1u) : "memory");
                    } else {
                        if (a_tmem)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                                         "r"(ta), "l"(dbi), "r"(idesc), "r"(1u) : "memory");
                        else
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                                         "l"(dai), "l"(dbi), "r"(idesc), "r"(1u) : "memory");
                    }
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
        mbar_wait(smem_u32(&bar), 0);
        const long long t1 = clock64();
        if ((tid & 31) == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
    }
}

template <int CG>
static int run_mma(int N, int mrows, int a_tmem, int layout, int alt_d, long long* d) {
    const int nmma = 1536;
    cudaFuncSetAttribute(mma_time_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(CG);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = 98 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CG;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, mma_time_kernel<CG>, N, mrows, a_tmem, layout, alt_d, nmma, d));
        CK(cudaDeviceSynchronize());
    }
    long long c = 0;
    CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
    static const char* lname[] = {"none ", "SW128", "SW64 ", "SW32 "};
    printf("mma cg=%d rows/CTA=%3d N=%3d layout=%s a_tmem=%d alt_d=%d : %6.1f cycles/MMA (ideal %d)\n", CG, mrows, N, lname[layout], a_tmem, alt_d,
           (double)c / nmma, mrows == 64 ? N / 4 : N / 2);
    return 0;
}

static int run_mma_all() {
    long long* d;
    CK(cudaMalloc(&d, 8));
    for (int N : {112, 176, 256})
        for (int layout = 0; layout < 4; ++layout) {
            run_mma<1>(N, 128, 0, layout, 0, d);
            run_mma<2>(N, 128, 0, layout, 0, d);
            run_mma<2>(N, 64, 0, layout, 0, d);
        }
    for (int N : {112, 176, 256}) {
        run_mma<2>(N, 128, 1, 0, 0, d);
        run_mma<2>(N, 128, 1, 1, 0, d);
        run_mma<2>(N, 64, 1, 0, 0, d);
    }
    // two accumulators used alternately (does the A fetch of the next MMA overlap when D differs?)
    for (int layout : {0, 1}) {
        run_mma<2>(176, 128, 0, layout, 1, d);
        run_mma<2>(176, 64, 0, layout, 1, d);
        run_mma<2>(176, 128, 1, layout, 1, d);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// l: L2 -> smem bulk-copy streaming through a ring, every CTA, optional cluster multicast
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(bar), "r"(rank) : "memory");
}

// Each CTA streams `nfill` slots of `slot_bytes` from its half (rank & 1) of the image.  mcast = 1: the CTAs of a cluster that
// share a half (same rank & 1) each fetch 1/G of every slot and multicast it to all G of them.
__global__ void __launch_bounds__(128, 1)
l2_stream_kernel(const uint8_t* __restrict__ img, unsigned half_bytes, int slot_bytes, int nslots, int nfill, int mcast, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank(), csize = cluster_nctarank();
    const uint32_t G = (mcast && csize >= 2) ? csize / 2 : 1;
    const uint32_t base = smem_u32(smem);
    const uint32_t bar_full = base, bar_empty = base + 8 * 16, ring = base + 1024;
    if (tid == 0) {
        for (int s = 0; s < nslots; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar_full + 8 * s), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar_empty + 8 * s), "r"(G));
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();
    const long long t0 = clock64();
    if (warp == 0 && lane == 0) {
        const uint8_t* src0 = img + (size_t)(rank & 1) * half_bytes;
        unsigned off = 0;
        uint16_t mask = 0;
        for (uint32_t g = 0; g < G; ++g) mask |= (uint16_t)(1u << ((rank & 1) + 2 * g));
        const uint32_t part = (uint32_t)slot_bytes / G, my = (rank >> 1) * part;
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < nfill; ++i) {
            if (!mbar_wait(bar_empty + 8 * s, ph ^ 1u)) { out[1] = 1; break; }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_full + 8 * s), "r"((uint32_t)slot_bytes) : "memory");
            if (G > 1) {
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n" ::"r"(
                        ring + s * slot_bytes + my),
                    "l"(src0 + off + my), "r"(part), "r"(bar_full + 8 * s), "h"(mask)
                    : "memory");
            } else {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(ring + s * slot_bytes),
                             "l"(src0 + off), "r"((uint32_t)slot_bytes), "r"(bar_full + 8 * s)
                             : "memory");
            }
            off += slot_bytes;
            if (off + slot_bytes > half_bytes) off = 0;
            if (++s == nslots) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1 && lane == 0) {
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < nfill; ++i) {
            if (!mbar_wait(bar_full + 8 * s, ph)) { out[1] = 2; break; }
            if (G > 1) {
                for (uint32_t g = 0; g < G; ++g) mbar_arrive_remote(bar_empty + 8 * s, (rank & 1) + 2 * g);
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar_empty + 8 * s) : "memory");
            }
            if (++s == nslots) { s = 0; ph ^= 1u; }
        }
    }
    __syncthreads();
    cluster_sync_all();
    if (tid == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

static int run_l2_one(const uint8_t* img, unsigned half_bytes, int csize, int mcast, int slot_bytes, int nslots, int grid, long long* d) {
    const int nfill = 2800;  // ~40 MB per CTA at 14 KB slots, like one 1M-row launch
    const size_t smem = 2048 + (size_t)nslots * slot_bytes;
    CK(cudaFuncSetAttribute(l2_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (csize > 8) CK(cudaFuncSetAttribute(l2_stream_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = csize;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, l2_stream_kernel, img, half_bytes, slot_bytes, nslots, nfill, mcast, d));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    long long flags[2] = {0, 0};
    CK(cudaMemcpy(flags, d, 16, cudaMemcpyDeviceToHost));
    if (flags[1]) printf("  (TIMED OUT in a barrier wait, code %lld)\n", flags[1]);
    CK(cudaMemset(d, 0, 16));
    const double delivered = (double)grid * nfill * slot_bytes;
    const int G = (mcast && csize >= 2) ? csize / 2 : 1;
    printf("l2 cluster=%d mcast=%d grid=%3d slot=%5d B x %2d slots (smem %3zu KB): %.3f ms  delivered %.2f TB/s  L2 reads %.2f TB/s  per-SM %.1f B/ns\n", csize,
           mcast, grid, slot_bytes, nslots, smem / 1024, best, delivered / best / 1e9, delivered / G / best / 1e9, (double)nfill * slot_bytes / best / 1e6);
    return 0;
}

static int run_l2() {
    const unsigned half_bytes = 751616;  // half of the 1,503,232-byte operand image
    uint8_t* img;
    long long* d;
    CK(cudaMalloc(&img, 2 * half_bytes));
    CK(cudaMemset(img, 1, 2 * half_bytes));
    CK(cudaMalloc(&d, 16));
    CK(cudaMemset(d, 0, 16));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    // ring depth / slot size sensitivity without multicast (clusters of 2 as in the product kernel)
    for (int nslots : {2, 4, 8, 12}) run_l2_one(img, half_bytes, 2, 0, 14336, nslots, sms, d);
    for (int slot : {7168, 28672, 57344}) run_l2_one(img, half_bytes, 2, 0, slot, 3, sms, d);
    run_l2_one(img, half_bytes, 2, 0, 14336, 4, sms / 2, d);  // half the SMs: is the limit per SM or chip wide?
    run_l2_one(img, half_bytes, 1, 0, 14336, 4, sms, d);
    // multicast: clusters of 4 and 8 (pairs share each half)
    for (int cs : {4, 8}) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(sms / cs * cs);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = 2048 + 4 * 14336;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int ncl = 0;
        CK(cudaFuncSetAttribute(l2_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        cudaOccupancyMaxActiveClusters(&ncl, l2_stream_kernel, &cfg);
        const int grid = ncl * cs;
        printf("cluster=%d: %d co-resident clusters (small smem) -> grid %d\n", cs, ncl, grid);
        if (grid <= 0) continue;
        for (int nslots : {4, 8}) {
            run_l2_one(img, half_bytes, cs, 1, 14336, nslots, grid, d);
            run_l2_one(img, half_bytes, cs, 0, 14336, nslots, grid, d);
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// o: co-resident clusters with the product kernel's footprint (640 threads, 220 KB of shared memory)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(640, 1) occ_kernel(int* p) {
    extern __shared__ uint8_t sm[];
    if (p && threadIdx.x == 0) p[blockIdx.x] = sm[0];
}
static int run_occ() {
    CK(cudaFuncSetAttribute(occ_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(occ_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(sms / cs * cs);
        cfg.blockDim = dim3(640);
        cfg.dynamicSmemBytes = 220 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int ncl = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, occ_kernel, &cfg);
        printf("occupancy: cluster size %2d, 640 threads, 220 KB smem: %d clusters = %d of %d SMs (%s)\n", cs, ncl, ncl * cs, sms, cudaGetErrorString(e));
    }
    return 0;
}


// ---------------------------------------------------------------------------------------------
// c: does shared-memory traffic from other warps (epilogue LDS/STS, bulk copies landing in the weight ring) slow tcgen05.mma?
//    warp 0 issues MMAs (cg 1, M = 128); warps 1.. generate traffic until warp 0 is done.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(288, 1)
contention_kernel(int N, int a_tmem, int nmma, int traffic, int ntraffic_warps, int bulk, const uint8_t* __restrict__ img, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t rbar[4];
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int done;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < 160 * 1024 / 4; e += 288) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u;
    if (tid == 0) {
        done = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&rbar[i])));
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
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 32768;
        const uint64_t da = make_desc(sA, 2048, 128, 0);
        const uint64_t db = make_desc(sB, (uint32_t)N * 16, 128, 0);
        const long long t0 = clock64();
        for (int i = 0; i < nmma; i += 6) {
            if (elect_one()) {
#pragma unroll
                for (int u = 0; u < 6; ++u) {
                    const uint64_t dbi = db + (uint64_t)(((u & 3) * 2 * (uint32_t)N * 16) >> 4);
                    const uint64_t dai = da + (uint64_t)(((u & 3) * 4096) >> 4);
                    const uint32_t ta = tm + 448u + (u & 3) * 8;
                    if (a_tmem)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tm),
                                     "r"(ta), "l"(dbi), "r"(idesc), "r"(1u) : "memory");
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
        mbar_wait(smem_u32(&bar), 0);
        const long long t1 = clock64();
        done = 1;
        if (lane == 0) out[0] = t1 - t0;
    } else if (warp == 8) {
        // bulk copies landing in a 4 x 14336 B ring at smem + 96 KB
        long long copies = 0;
        if (lane == 0 && bulk) {
            unsigned off = 0;
            uint32_t ph = 0;
            int s = 0;
            while (!done) {
                const uint32_t b = smem_u32(&rbar[s]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(14336u) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                                 smem_u32(smem) + 98304u + s * 14336u),
                             "l"(img + off), "r"(14336u), "r"(b)
                             : "memory");
                if (bulk == 1) mbar_wait(b, ph);  // one copy in flight at a time; bulk == 2: four in flight
                off = (off + 14336u) % (700u * 1024u);
                ++copies;
                if (++s == 4) {
                    s = 0;
                    if (bulk == 2)
                        for (int q = 0; q < 4; ++q) mbar_wait(smem_u32(&rbar[q]), ph);
                    ph ^= 1u;
                }
            }
            out[2] = copies;
        }
    } else if (warp >= 1 && warp <= 8 && warp != 8 && (traffic & 12) && warp - 1 < ntraffic_warps) {
        // TMEM traffic: this warp's sub-partition (warp % 4), columns 256.. (the MMAs accumulate into columns 0..N-1, A operand at 448..)
        const uint32_t taddr = tm + ((uint32_t)((warp & 3) * 32) << 16) + 256u + (uint32_t)((warp - 1) / 4) * 64u;
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = tid + i;
        long long n = 0;
        while (!done) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (traffic & 4) {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr + 16u * j));
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                }
                if (traffic & 8) {
                    asm volatile(
                        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(
                            taddr + 16u * j),
                        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                        : "memory");
                    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
                }
            }
            n += 4;
        }
        if (r[3] == 0x12345678u) out[3] = 1;
        if (lane == 0) atomicAdd((unsigned long long*)&out[1], (unsigned long long)n * 4ull);  // in units of 512 B like the LSU counter (2 KB each)
    } else if (warp - 1 < ntraffic_warps && traffic) {
        // conflict-free 128-bit accesses to a private 16 KB window per warp above the MMA operands
        uint4* w = reinterpret_cast<uint4*>(smem + 160 * 1024 - 16384 * 0) - 0;  // placeholder, replaced below
        w = reinterpret_cast<uint4*>(smem + 65536 + (warp - 1) * 4096);
        uint4 v = make_uint4(tid, 1, 2, 3);
        long long n = 0;
        while (!done) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (traffic & 1) w[j * 32 + lane] = v;
                if (traffic & 2) {
                    const uint4 r = w[((j + 3) & 7) * 32 + lane];
                    v.x ^= r.y;
                }
            }
            n += 8;
        }
        if (v.x == 0x12345678u) out[3] = 1;
        if (lane == 0) atomicAdd((unsigned long long*)&out[1], (unsigned long long)n);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
}

static int run_contention() {
    long long* d;
    uint8_t* img;
    CK(cudaMalloc(&d, 32));
    CK(cudaMalloc(&img, 1 << 20));
    CK(cudaMemset(img, 1, 1 << 20));
    CK(cudaFuncSetAttribute(contention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int nmma = 3072;
    for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
        for (int N : {144, 176})
            for (int cfg = 0; cfg < 13; ++cfg) {
                // traffic: 0 none, 1 STS, 2 LDS, 3 both, 4 tcgen05.ld, 8 tcgen05.st, 12 both; warps; bulk
                int traffic = (cfg == 0 || cfg == 7 || cfg == 8) ? 0 : (cfg <= 2 ? 1 : cfg <= 4 ? 2 : 3);
                int nw = (cfg == 1 || cfg == 3 || cfg == 5) ? 2 : 7;
                const int bulk = cfg == 7 ? 1 : cfg == 8 ? 2 : (cfg == 6 ? 2 : 0);
                if (cfg == 9) { traffic = 4; nw = 4; }
                if (cfg == 10) { traffic = 4; nw = 7; }
                if (cfg == 11) { traffic = 8; nw = 4; }
                if (cfg == 12) { traffic = 12; nw = 7; }
                CK(cudaMemset(d, 0, 32));
                for (int rep = 0; rep < 2; ++rep) {
                    CK(cudaMemset(d, 0, 32));
                    contention_kernel<<<1, 288, 180 * 1024>>>(N, a_tmem, nmma, traffic, nw, bulk, img, d);
                    CK(cudaDeviceSynchronize());
                }
                long long r[4];
                CK(cudaMemcpy(r, d, 32, cudaMemcpyDeviceToHost));
                const double cyc = (double)r[0];
                // each traffic iteration = 8 warp-wide 128-bit accesses = 8 x 4 wavefronts per enabled direction
                const int dirs = traffic >= 4 ? ((traffic >> 2) & 1) + ((traffic >> 3) & 1) : (traffic & 1) + ((traffic >> 1) & 1);
                printf("contention N=%d a_tmem=%d traffic=%2d warps=%d bulk=%d : %6.1f cycles/MMA   other traffic %.1f B/cycle (LSU or TMEM) + %.1f B/cycle (bulk)\n", N,
                       a_tmem, traffic, traffic ? nw : 0, bulk, cyc / nmma, (double)r[1] * 512.0 * dirs / cyc, (double)r[2] * 14336.0 / cyc);
            }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// i: issue-side costs seen by the MMA warp (CTA pair, cta_group::2): cycles the issuing warp spends in
//    (a) a block of G UTCHMMA (N = 176, SS) without waiting for completion, (b) tcgen05.commit multicast to both CTAs,
//    (c) mbarrier.try_wait on an already completed phase, (d) tcgen05.fence after / before, (e) elect.sync, (f) __syncwarp,
//    (g) bar.sync of 64 threads.  Each measured as clock64 deltas around REP repetitions.
// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
issue_cost_kernel(int N, int a_tmem, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar, bar2, bar3;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    for (int e = tid; e < 96 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar2)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar3)));
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
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tm = tmem_base_s;
    if (warp == 0 && rank == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 32768;
        const uint32_t b_kg = (uint32_t)(N / 2) * 16;
        const uint64_t da = make_desc(sA, 2048, 128, 0);
        const uint64_t db = make_desc(sB, b_kg, 128, 0);
        long long t[16];
        const int REP = 32;
        // (a) issue blocks of G = 1, 2, 4, 8 MMAs; after each block wait for completion OUTSIDE the timed region
        int slot = 0;
        for (int G = 1; G <= 8; G *= 2) {
            long long acc = 0;
            uint32_t ph = 0;
            // drain
            for (int r = 0; r < REP; ++r) {
                const long long t0 = clock64();
                if (elect_one()) {
                    for (int u = 0; u < G; ++u) {
                        const uint64_t dbi = db + (uint64_t)(((u & 3) * 2 * b_kg) >> 4);
                        const uint64_t dai = da + (uint64_t)(((u & 3) * 4096) >> 4);
                        const uint32_t ta = tm + 448u + (u & 3) * 8;
                        if (a_tmem)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tm),
                                         "r"(ta), "l"(dbi), "r"(idesc), "r"(1u) : "memory");
                        else
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm),
                                         "l"(dai), "l"(dbi), "r"(idesc), "r"(1u) : "memory");
                    }
                }
                __syncwarp();
                acc += clock64() - t0;
                if (elect_one())
                    asm volatile("{\n\t.reg .b16 m;\n\tmov.b16 m, 1;\n\ttcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}\n" ::"r"(smem_u32(&bar)) : "memory");
                __syncwarp();
                mbar_wait(smem_u32(&bar), ph);
                ph ^= 1u;
            }
            t[slot++] = acc / REP;   // 0..3: G = 1, 2, 4, 8
            if (lane == 0) out[20 + slot] = ph;
        }
        // (a2) 32 MMAs back to back: where does the issuer block (queue depth)?  time after 4, 8, 16, 32
        {
            const long long t0 = clock64();
            long long marks[4] = {0, 0, 0, 0};
            for (int u = 0; u < 32; ++u) {
                if (elect_one()) {
                    const uint64_t dbi = db + (uint64_t)(((u & 3) * 2 * b_kg) >> 4);
                    const uint64_t dai = da + (uint64_t)(((u & 3) * 4096) >> 4);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm),
                                 "l"(dai), "l"(dbi), "r"(idesc), "r"(1u) : "memory");
                }
                __syncwarp();
                if (u == 3) marks[0] = clock64() - t0;
                if (u == 7) marks[1] = clock64() - t0;
                if (u == 15) marks[2] = clock64() - t0;
                if (u == 31) marks[3] = clock64() - t0;
            }
            if (elect_one())
                asm volatile("{\n\t.reg .b16 m;\n\tmov.b16 m, 1;\n\ttcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}\n" ::"r"(smem_u32(&bar2)) : "memory");
            __syncwarp();
            mbar_wait(smem_u32(&bar2), 0);
            t[4] = marks[0]; t[5] = marks[1]; t[6] = marks[2]; t[7] = marks[3];
            t[8] = clock64() - t0;  // all 32 complete
        }
        // (b) commit cost (nothing pending), multicast to both CTAs; consume the arrivals afterwards
        {
            uint32_t ph = 1;  // bar2 has completed phase 0
            long long acc = 0;
            for (int r = 0; r < REP; ++r) {
                const long long t0 = clock64();
                if (elect_one())
                    asm volatile("{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\ttcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}\n" ::"r"(smem_u32(&bar2)) : "memory");
                __syncwarp();
                acc += clock64() - t0;
                mbar_wait(smem_u32(&bar2), ph);
                ph ^= 1u;
            }
            t[9] = acc / REP;
            // (c) try_wait on a completed phase
            acc = 0;
            for (int r = 0; r < REP; ++r) {
                const long long t0 = clock64();
                uint32_t ok;
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar2)), "r"(ph ^ 1u) : "memory");
                if (!ok) out[30] = 1;
                acc += clock64() - t0;
            }
            t[10] = acc / REP;
        }
        {
            long long acc = 0;
            for (int r = 0; r < REP; ++r) {
                const long long t0 = clock64();
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                acc += clock64() - t0;
            }
            t[11] = acc / REP;
            acc = 0;
            for (int r = 0; r < REP; ++r) {
                const long long t0 = clock64();
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                acc += clock64() - t0;
            }
            t[12] = acc / REP;
            acc = 0;
            uint32_t e = 0;
            for (int r = 0; r < REP; ++r) {
                const long long t0 = clock64();
                e += elect_one();
                __syncwarp();
                acc += clock64() - t0;
            }
            t[13] = acc / REP + (e == 12345u);
            acc = 0;
            for (int r = 0; r < REP; ++r) {
                const long long t0 = clock64();
                acc += clock64() - t0;
            }
            t[14] = acc / REP;  // timer overhead
            t[15] = 0;
        }
        if (lane == 0)
            for (int i = 0; i < 16; ++i) out[i] = t[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
}

static int run_issue_cost() {
    long long* d;
    CK(cudaMalloc(&d, 8 * 64));
    CK(cudaFuncSetAttribute(issue_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
        for (int N : {64, 176}) {
            CK(cudaMemset(d, 0, 8 * 64));
            for (int rep = 0; rep < 2; ++rep) {
                issue_cost_kernel<<<2, 128, 98 * 1024>>>(N, a_tmem, d);
                CK(cudaDeviceSynchronize());
            }
            long long t[16];
            CK(cudaMemcpy(t, d, sizeof t, cudaMemcpyDeviceToHost));
            printf("issue N=%d a_tmem=%d: block of 1/2/4/8 MMAs %lld/%lld/%lld/%lld cycles to issue; 32 back to back: after 4/8/16/32 issued %lld/%lld/%lld/%lld, all "
                   "complete %lld; commit %lld; try_wait(done) %lld; fence after %lld before %lld; elect+syncwarp %lld; timer %lld\n",
                   N, a_tmem, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], t[10], t[11], t[12], t[13], t[14]);
        }
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        printf("usage: tc_probe2 m|l|o\n");
        return 1;
    }
    if (argv[1][0] == 'm') return run_mma_all();
    if (argv[1][0] == 'l') return run_l2();
    if (argv[1][0] == 'o') return run_occ();
    if (argv[1][0] == 'c') return run_contention();
    if (argv[1][0] == 'i') return run_issue_cost();
    return 1;
}
