// Issue-rate probes for the FP32 kernel's inner loop (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_probe tools/ffma2_probe.cu && ./tools/ffma2_probe
// (1) `stream`: cycles one scheduler needs per FFMA2 (fma.rn.f32x2) / FFMA when 1..4 warps share it, each running a chain-free
//     stream of independent accumulators with operands that defeat the register reuse cache;
// (2) `tile`: the kernel's real register tile -- 8 rows x TN columns per thread, per k: 8 row values x (TN/2 column pairs + one
//     odd column), column-pair outer / row inner like the compiled kernel -- with the operands either held in registers
//     (`regs`) or loaded from shared memory every k exactly as fp32_pipe_kernel.cuh does (`lds`), no barriers, no hand-off:
//     the ceiling of the inner loop itself.
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC, bool PACKED>
__global__ void __launch_bounds__(512, 1) stream(const float* in, float* out, long long* cycles, int iters) {
    float2 acc[NACC];
    const float2 w0 = make_float2(in[threadIdx.x], in[threadIdx.x + 1]);
    const float2 w1 = make_float2(in[threadIdx.x + 2], in[threadIdx.x + 3]);
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = in[64 + i];
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = make_float2(0.f, 0.f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            const float s = a[j & 7];
            if (PACKED) {
                acc[j] = __ffma2_rn(make_float2(s, s), (j & 1) ? w1 : w0, acc[j]);
            } else {
                acc[j].x = fmaf(s, (j & 1) ? w1.x : w0.x, acc[j].x);
                acc[j].y = fmaf(s, (j & 1) ? w1.y : w0.y, acc[j].y);
            }
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int j = 0; j < NACC; ++j) r += acc[j].x + acc[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

constexpr int LDA = 68;
template <int TN, bool LDS>
__global__ void __launch_bounds__(256, 1) tile(const float* in, float* out, long long* cycles, int kiters) {
    extern __shared__ __align__(16) float sm[];
    constexpr int Npad = 32 * TN, NQ = TN >> 2, NS = TN & 3, NP = 2 * NQ + (NS >= 2 ? 1 : 0);
    constexpr bool ODD = (NS & 1) != 0;
    float* act = sm;                 // [8 k][LDA]
    float* ws = sm + 8 * LDA;        // [8 k][Npad]
    for (int i = threadIdx.x; i < 8 * LDA + 8 * Npad; i += blockDim.x) sm[i] = in[i & 1023];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 acc2[8][NP];
    float acc1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc1[i] = 0.f;
#pragma unroll
        for (int q = 0; q < NP; ++q) acc2[i][q] = make_float2(0.f, 0.f);
    }
    float2 wp[NP];
    float w1 = 0.f, av[8];
    if (!LDS) {
#pragma unroll
        for (int q = 0; q < NP; ++q) wp[q] = make_float2(ws[2 * q + lane], ws[2 * q + 1 + lane]);
        w1 = ws[lane + 77];
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = act[i + 8 * warp];
    }
    const long long t0 = clock64();
    for (int kb = 0; kb < kiters; ++kb) {
        const float* ap = act + 8 * warp;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            if (LDS) {
                const float4 a0 = *reinterpret_cast<const float4*>(ap + kk * LDA);
                const float4 a1 = *reinterpret_cast<const float4*>(ap + kk * LDA + 4);
                const float* wk = ws + kk * Npad;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wk + 128 * q + 4 * lane);
                    wp[2 * q] = make_float2(w4.x, w4.y);
                    wp[2 * q + 1] = make_float2(w4.z, w4.w);
                }
                if (NS >= 2) wp[NP - 1] = *reinterpret_cast<const float2*>(wk + 128 * NQ + 2 * lane);
                if (ODD) w1 = wk[128 * NQ + (NS >= 2 ? 64 : 0) + lane];
                av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
                av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 aa = make_float2(av[i], av[i]);
#pragma unroll
                for (int q = 0; q < NP; ++q) acc2[i][q] = __ffma2_rn(aa, wp[q], acc2[i][q]);
                if (ODD) acc1[i] = fmaf(av[i], w1, acc1[i]);
            }
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r += acc1[i];
#pragma unroll
        for (int q = 0; q < NP; ++q) r += acc2[i][q].x + acc2[i][q].y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NACC, bool PACKED>
void run_stream(const char* name, const float* in, float* out, long long* cyc) {
    const int iters = 4000;
    for (int wps = 1; wps <= 4; ++wps) {
        stream<NACC, PACKED><<<1, 128 * wps>>>(in, out, cyc, iters);
        stream<NACC, PACKED><<<1, 128 * wps>>>(in, out, cyc, iters);
        cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        const double n = static_cast<double>(iters) * NACC * (PACKED ? 1 : 2) * wps;
        printf("stream %-6s NACC=%2d warps/scheduler=%d: %.3f cycles per instruction per scheduler, %.1f of 32 fp32 lanes busy\n", name, NACC, wps,
               c / n, n * (PACKED ? 64 : 32) / static_cast<double>(c));
    }
}

template <int TN, bool LDS>
void run_tile(const float* in, float* out, long long* cyc) {
    const int kiters = 2000;
    constexpr int NP = 2 * (TN >> 2) + ((TN & 3) >= 2 ? 1 : 0);
    constexpr bool ODD = (TN & 1) != 0;
    const size_t smem = (8 * LDA + 8 * 32 * TN) * sizeof(float);
    for (int wps = 1; wps <= 2; ++wps) {
        tile<TN, LDS><<<1, 128 * wps, smem>>>(in, out, cyc, kiters);
        tile<TN, LDS><<<1, 128 * wps, smem>>>(in, out, cyc, kiters);
        cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        const double lane_cycles = static_cast<double>(kiters) * 8 * wps * (8.0 * NP * 64 + (ODD ? 8.0 * 32 : 0.0));
        const double nominal = static_cast<double>(kiters) * 8 * wps * (8.0 * NP * 2 + (ODD ? 8.0 : 0.0));
        printf("tile TN=%2d %-4s warps/scheduler=%d: %9lld cycles, %.1f of 32 fp32 lanes busy (nominal pipe cycles %.0f = %.1f %%)\n", TN,
               LDS ? "lds" : "regs", wps, c, lane_cycles / c, nominal, 100.0 * nominal / c);
    }
}

int main() {
    float *in, *out;
    long long* cyc;
    cudaMalloc(&in, 8192);
    cudaMemset(in, 0, 8192);
    cudaMalloc(&out, 4096 * 4);
    cudaMalloc(&cyc, 64);
    run_stream<40, true>("FFMA2", in, out, cyc);
    run_stream<40, false>("FFMA", in, out, cyc);
    run_tile<11, false>(in, out, cyc);
    run_tile<11, true>(in, out, cyc);
    run_tile<15, false>(in, out, cyc);
    run_tile<15, true>(in, out, cyc);
    run_tile<9, true>(in, out, cyc);
    run_tile<7, true>(in, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
