// FP32-SIMT fused emulator kernel: parameter transform -> Dense chain -> output transform
// (or fused chi^2) in ONE launch.  This is the parity path: every product is an fp32 FFMA
// accumulated in k order, bias/ReLU/de-normalisation follow the reference's rounding order.
//
// Reference semantics: VeryAccurateEmulator/emulator.py:401-403, preprocess.py:49-110, :27-46.
//
// Layout per CTA (256 threads, 64 rows per tile, persistent over tiles):
//   activations live in shared memory k-major: act[k][m], row stride LDA = 68 floats, two
//   ping-pong buffers (even / odd layer inputs);
//   weights stream from global/L2 through a cp.async ring of [KB][Npad] fp32 stages
//   (host-packed, zero padded: Kpad % 8 == 0, Npad % 32 == 0);
//   warp w owns rows 8w..8w+7 of the tile; lane t owns TN columns: 4 consecutive ones in each leading
//   128-column block (one LDS.128 per k), then -- packed build, the default -- 2 consecutive ones in a 64-column block
//   (one LDS.64) and at most one single column in a last 32-column slot.  Adjacent columns of a lane are accumulated in
//   pairs by FFMA2 (`fma.rn.f32x2`, sm_100: two independent IEEE fmas per instruction, bit-identical to two FFMA):
//   per k a thread does 2 broadcast LDS.128 (its 8 rows), TN/4 LDS.128 (+ LDS.64, + LDS.32), 8 register duplications of
//   its row values and 8*(TN/2) FFMA2 (+ 8 FFMA for an odd column) -- about half the issue slots of the scalar build
//   (VAE21_FP32_FFMA2=0: one column per remaining 32-column slot, 8*TN FFMA), 18.53 -> 17.90 ms per 1M rows on the same box.
#pragma once
#include "common.cuh"

#ifndef VAE21_FP32_FFMA2
#define VAE21_FP32_FFMA2 1
#endif
#ifndef VAE21_FP32_UNROLL
#define VAE21_FP32_UNROLL 8
#endif

namespace f32k {

constexpr int MT = 64;        // rows per tile
constexpr int LDA = MT + 4;   // 68: 16B-aligned rows; 17 x 16B chunks per row => conflict-free STS.128
constexpr int KB = 8;         // k rows per weight stage
constexpr int NTHREADS = 256;
constexpr int KK_UNROLL = VAE21_FP32_UNROLL;  // unroll depth of the k loop inside a weight stage
constexpr int MAX_SLOTS = 15; // columns per lane => widest layer 480

struct Layer {
    int K, Kpad, N, Npad, relu;
    long long w_off;  // float offset into the packed weight array ([Kpad][Npad] row-major)
    long long b_off;  // float offset into the packed bias array ([Npad])
};

struct Model {
    int n_layers;
    int buf_rows[2];  // rows (k extent) of the two activation buffers
    int max_npad;
    Layer L[VAE21_MAX_LAYERS];
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// One Dense layer for one 64-row tile.  `in`/`outbuf` are k-major smem activation buffers.
template <int TN, int WST>
__device__ __forceinline__ void run_layer(const Layer& L, bool last, const float* __restrict__ Wg,
                                          const float* __restrict__ Bg, const float* in, float* outbuf, float* wst,
                                          int stage_floats, const NormConsts& nc, const LaunchArgs& a, long long row0,
                                          long long next_w_off, int next_npad, int next_kpad, bool prefetched) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Npad = L.Npad;
    const float* Wl = Wg + L.w_off;
    const int nkb = L.Kpad / KB;
    const int chunks = (KB * Npad) >> 2;  // 16-byte chunks per stage

    // Column ownership of a lane: the first NQ*128 columns in "quads" (4 consecutive columns, one LDS.128 per k),
    // the remaining NS 32-column slots one column each.  Slot j of the accumulator array is column col_of(j).
    constexpr int NQ = TN >> 2, NS = TN & 3;
#if VAE21_FP32_FFMA2
    // packed variant: two of the remaining slots form one 64-column block owned in column PAIRS (one LDS.64 per k), so that all
    // but at most one column of a lane are accumulated by FFMA2 (fma.rn.f32x2: two independent IEEE fmas, same rounding as FFMA)
    constexpr int NP = 2 * NQ + (NS >= 2 ? 1 : 0);  // column pairs per lane
    constexpr bool ODD = (NS & 1) != 0;             // plus one single column
    auto col_of = [&](int j) {
        return j < 4 * NQ ? 128 * (j >> 2) + 4 * lane + (j & 3)
                          : (j < 2 * NP ? 128 * NQ + 2 * lane + (j - 4 * NQ) : 128 * NQ + (NS >= 2 ? 64 : 0) + lane);
    };
#else
    auto col_of = [&](int j) { return j < 4 * NQ ? 128 * (j >> 2) + 4 * lane + (j & 3) : 128 * NQ + 32 * (j - 4 * NQ) + lane; };
#endif

#if VAE21_FP32_FFMA2
    float2 acc2[8][NP > 0 ? NP : 1];
    float acc1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc1[i] = 0.f;
#pragma unroll
        for (int p = 0; p < NP; ++p) acc2[i][p] = make_float2(0.f, 0.f);
    }
#else
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
#endif

    // the lane's biases, fetched before the k loop so that the epilogue does not wait on L2 (bias arrays are padded to Npad)
    float bv[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) bv[j] = __ldg(Bg + L.b_off + col_of(j));

    auto load_stage = [&](int kb, int slot) {
        const float* src = Wl + static_cast<long long>(kb) * KB * Npad;
        float* dst = wst + slot * stage_floats;
        for (int c = tid; c < chunks; c += NTHREADS) cp_async16(dst + 4 * c, src + 4 * c);
    };

    if (!prefetched) {  // otherwise the previous layer issued these stages before its epilogue (see below)
#pragma unroll
        for (int s = 0; s < WST - 1; ++s) {
            if (s < nkb) load_stage(s, s);
            cp_async_commit();
        }
    }
    // After the k loop: start the first stages of the NEXT layer (of this tile, or layer 0 of the CTA's next tile) so that their
    // L2 latency hides behind this layer's epilogue instead of opening the next layer.  One extra barrier frees the ring slots.
#define PREFETCH_NEXT                                                                                        \
    if (next_npad) {                                                                                           \
        __syncthreads();                                                                                     \
        const float* nsrc = Wg + next_w_off;                                                                 \
        const int nchunks = (KB * next_npad) >> 2, nnkb = next_kpad / KB;                                    \
        _Pragma("unroll") for (int s = 0; s < WST - 1; ++s) {                                                \
            if (s < nnkb)                                                                                    \
                for (int c = tid; c < nchunks; c += NTHREADS)                                                \
                    cp_async16(wst + s * stage_floats + 4 * c, nsrc + static_cast<long long>(s) * KB * next_npad + 4 * c);  \
            cp_async_commit();                                                                               \
        }                                                                                                    \
    }
    for (int kb = 0; kb < nkb; ++kb) {
        cp_async_wait<WST - 2>();
        __syncthreads();  // stage kb visible to all; everyone is done with the slot refilled below
        const int nxt = kb + WST - 1;
        if (nxt < nkb) load_stage(nxt, nxt % WST);
        cp_async_commit();
        const float* ws = wst + (kb % WST) * stage_floats;
        const float* ap = in + (kb * KB) * LDA + 8 * warp;
#pragma unroll(KK_UNROLL)
        for (int kk = 0; kk < KB; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(ap + kk * LDA);
            const float4 a1 = *reinterpret_cast<const float4*>(ap + kk * LDA + 4);
            const float* wk = ws + kk * Npad;
#if VAE21_FP32_FFMA2
            float2 wp[NP > 0 ? NP : 1];
            float w1 = 0.f;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float4 w4 = *reinterpret_cast<const float4*>(wk + 128 * q + 4 * lane);
                wp[2 * q] = make_float2(w4.x, w4.y);
                wp[2 * q + 1] = make_float2(w4.z, w4.w);
            }
            if (NS >= 2) wp[NP > 0 ? NP - 1 : 0] = *reinterpret_cast<const float2*>(wk + 128 * NQ + 2 * lane);
            if (ODD) w1 = wk[128 * NQ + (NS >= 2 ? 64 : 0) + lane];
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 aa = make_float2(av[i], av[i]);
#pragma unroll
                for (int p = 0; p < NP; ++p) acc2[i][p] = __ffma2_rn(aa, wp[p], acc2[i][p]);
                if (ODD) acc1[i] = fmaf(av[i], w1, acc1[i]);
            }
        }
    }
    PREFETCH_NEXT
    float acc[8][TN];  // register renaming only: slot j of the epilogues below is column col_of(j)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            acc[i][2 * p] = acc2[i][p].x;
            acc[i][2 * p + 1] = acc2[i][p].y;
        }
        if (ODD) acc[i][TN - 1] = acc1[i];
    }
#else
            float wv[TN];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float4 w4 = *reinterpret_cast<const float4*>(wk + 128 * q + 4 * lane);
                wv[4 * q + 0] = w4.x;
                wv[4 * q + 1] = w4.y;
                wv[4 * q + 2] = w4.z;
                wv[4 * q + 3] = w4.w;
            }
#pragma unroll
            for (int sgl = 0; sgl < NS; ++sgl) wv[4 * NQ + sgl] = wk[128 * NQ + 32 * sgl + lane];
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const float w = wv[j];
                acc[0][j] = fmaf(a0.x, w, acc[0][j]);
                acc[1][j] = fmaf(a0.y, w, acc[1][j]);
                acc[2][j] = fmaf(a0.z, w, acc[2][j]);
                acc[3][j] = fmaf(a0.w, w, acc[3][j]);
                acc[4][j] = fmaf(a1.x, w, acc[4][j]);
                acc[5][j] = fmaf(a1.y, w, acc[5][j]);
                acc[6][j] = fmaf(a1.z, w, acc[6][j]);
                acc[7][j] = fmaf(a1.w, w, acc[7][j]);
            }
        }
    }
    PREFETCH_NEXT
#endif
#undef PREFETCH_NEXT

    if (!last) {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = col_of(j);
            const float b = bv[j];
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[i] = __fadd_rn(acc[i][j], b);
                if (L.relu) v[i] = v[i] < 0.f ? 0.f : v[i];  // NaN propagates like np.maximum / tf.nn.relu
            }
            float4* dst = reinterpret_cast<float4*>(outbuf + n * LDA + 8 * warp);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    } else {
        const int N = L.N;
        const long long rbase = row0 + 8 * warp;
        if (a.out_mode == OUT_CHI2) {
            float part[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) part[i] = 0.f;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = col_of(j);
                if (n < N) {
                    const float b = bv[j], mu = __ldg(a.mu + n), ob = __ldg(a.obs + n), is = __ldg(a.isig + n);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float v = __fadd_rn(acc[i][j], b);
                        if (L.relu) v = v < 0.f ? 0.f : v;
                        v = __fadd_rn(__fmul_rn(v, nc.sd), mu);
                        const float r = (v - ob) * is;
                        part[i] = fmaf(r, r, part[i]);
                    }
                }
            }
            unsigned long long best = ~0ull;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float s = warp_sum(part[i]);
                const long long row = rbase + i;
                if (row < a.n) {
                    if (a.chi2 && lane == 0) a.chi2[row] = s;
                    const unsigned long long key = pack_min_key(s, static_cast<unsigned long long>(a.row_base + row));
                    best = key < best ? key : best;
                }
            }
            if (a.argmin_key && lane == 0 && best != ~0ull) atomicMin(a.argmin_key, best);
        } else if (a.out_mode == OUT_ERROR) {
            // fused emulator.py:129-192: rms difference to the row's true signal over the band, optionally in % of its amplitude
            float part[8], amp[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) part[i] = amp[i] = 0.f;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = col_of(j);
                if (n < N) {
                    const float b = bv[j], mu = __ldg(a.mu + n), in_band = __ldg(a.isig + n);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const long long row = rbase + i;
                        float v = __fadd_rn(acc[i][j], b);
                        if (L.relu) v = v < 0.f ? 0.f : v;
                        v = __fadd_rn(__fmul_rn(v, nc.sd), mu);
                        const float t = row < a.n ? __ldg(a.truth + row * N + n) : 0.f;
                        const float r = (v - t) * in_band;
                        part[i] = fmaf(r, r, part[i]);
                        amp[i] = fmaxf(amp[i], fabsf(t) * in_band);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float s = warp_sum(part[i]);
                float m = amp[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                const long long row = rbase + i;
                if (row < a.n && lane == 0) {
                    float e = sqrtf(s * a.err_inv_count);
                    if (a.err_relative) e = e / m * 100.f;
                    a.chi2[row] = e;
                }
            }
        } else {
            const bool denorm = (a.out_mode == OUT_PREDICT);
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = col_of(j);
                if (n < N) {
                    const float b = bv[j];
                    const float mu = denorm ? __ldg(a.mu + n) : 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const long long row = rbase + i;
                        if (row < a.n) {
                            float v = __fadd_rn(acc[i][j], b);
                            if (L.relu) v = v < 0.f ? 0.f : v;
                            if (denorm) v = __fadd_rn(__fmul_rn(v, nc.sd), mu);  // preprocess.py:44-45
                            __stcs(a.out + row * N + n, v);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();  // outputs visible; weight ring free for the next layer
}

template <int WST>
__global__ void __launch_bounds__(NTHREADS, 1)
vae21_fp32_kernel(const Model m, const NormConsts nc, const LaunchArgs a, const float* __restrict__ Wg,
                  const float* __restrict__ Bg) {
    extern __shared__ __align__(16) float smem[];
    float* const buf0 = smem;
    float* const buf1 = buf0 + m.buf_rows[0] * LDA;
    float* wst = buf1 + m.buf_rows[1] * LDA;
    const int stage_floats = KB * m.max_npad;
    const int tid = threadIdx.x;
    const long long ntiles = (a.n + MT - 1) / MT;
    const int K0 = m.L[0].K, Kpad0 = m.L[0].Kpad;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * MT;
        // ---- prologue: load + transform the tile's parameters into buf[0][k][m] ----
        for (int e = tid; e < MT * Kpad0; e += NTHREADS) {
            int r, c;
            float x = 0.f;
            if (e < MT * K0) {
                r = e / K0;
                c = e - r * K0;
                const long long row = row0 + r;
                if (row < a.n) {
                    const long long g = row * K0 + c;
                    if (a.in_mode == IN_PARAMS_F64)
                        x = transform_param(reinterpret_cast<const double*>(a.in)[g], c, nc, false);
                    else if (a.in_mode == IN_PARAMS_F32)
                        x = transform_param(static_cast<double>(reinterpret_cast<const float*>(a.in)[g]), c, nc, true);
                    else if (a.in_mode == IN_GRID) {
                        float xs[VAE21_MAX_PAR];
                        grid_point(a, static_cast<unsigned long long>(a.row_base + row), K0, xs);
                        x = xs[c];
                    } else
                        x = reinterpret_cast<const float*>(a.in)[g];
                }
            } else {  // zero the k padding rows
                const int e2 = e - MT * K0;
                c = K0 + e2 / MT;
                r = e2 - (c - K0) * MT;
            }
            buf0[c * LDA + r] = x;
        }
        __syncthreads();

        for (int l = 0; l < m.n_layers; ++l) {
            const Layer& L = m.L[l];
            const bool last = (l == m.n_layers - 1);
            const float* in = (l & 1) ? buf1 : buf0;
            float* ob = (l & 1) ? buf0 : buf1;
            const int slots = L.Npad >> 5;
            // weight-stage prefetch chain: every layer but the CTA's very first finds its leading stages already in flight
            const int ln = last ? 0 : l + 1;
            const bool has_next = !last || tile + gridDim.x < ntiles;
            const long long nw = m.L[ln].w_off;
            const int nn = has_next ? m.L[ln].Npad : 0, nk = m.L[ln].Kpad;
            const bool pre = !(l == 0 && tile == static_cast<long long>(blockIdx.x));
#define VAE21_CASE(T)                                                                             \
    case T:                                                                                       \
        run_layer<T, WST>(L, last, Wg, Bg, in, ob, wst, stage_floats, nc, a, row0, nw, nn, nk, pre);\
        break;
            switch (slots) {
                VAE21_CASE(1) VAE21_CASE(2) VAE21_CASE(3) VAE21_CASE(4) VAE21_CASE(5)
                VAE21_CASE(6) VAE21_CASE(7) VAE21_CASE(8) VAE21_CASE(9) VAE21_CASE(10)
                VAE21_CASE(11) VAE21_CASE(12) VAE21_CASE(13) VAE21_CASE(14) VAE21_CASE(15)
                default: break;
            }
#undef VAE21_CASE
        }
    }
}

}  // namespace f32k
