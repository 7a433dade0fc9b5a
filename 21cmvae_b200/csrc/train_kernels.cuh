// Training kernels for the Dense stack (reference: emulator.py:339-381 `train`, :51-83
// `relative_mse_loss`, Keras `fit` with batch_size 256 and Adam).  FP32 throughout, like Keras on CPU.
//
// One training step of a 256-row batch is ~0.6 GFLOP: launch-latency- rather than FLOP-bound, so the step is a short
// fixed sequence of plain CUDA-core kernels (captured once into a CUDA graph by the host side), not tensor-core GEMMs:
//   gather      rows idx[0..B) of the resident training set -> contiguous batch buffers
//   forward     H_{l+1} = act(H_l W_l + b_l)                                  (tiled SGEMM, bias + ReLU fused)
//   loss/delta  per-sample relative MSE, dL/dY of the batch-mean loss          (one warp per row)
//   backward    db_l = colsum(D_{l+1});  dW_l = H_l^T D_{l+1};  D_l = (D_{l+1} W_l^T) * [H_l > 0]
//   adam        Keras Adam: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t),  p -= lr_t m / (sqrt(v) + eps)
// Gradients land in ONE flat buffer in Keras `get_weights()` order (kernel [in,out] row-major, then bias, per layer), so
// data-parallel ranks all-reduce it with a single NCCL call between `backward` and `adam`.
// Every reduction has a fixed order: results are bitwise reproducible for a given batch split.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace trk {

constexpr int BK = 16, NT = 256;  // k-depth of a staged tile, threads per block

// C[M,N] = op(A) op(B) with fused epilogues; BM x BN output tile per block, TM x TN outputs per thread ((BM/TM) (BN/TN) = 256).
// The batch is only 256 rows, so parallelism has to come from many small tiles: 32 x 32 tiles put 88..120 blocks on the 148 SMs
// for the layer shapes of the emulator where 64 x 64 tiles would put 24..30.
//   MODE 0 (forward)      A = H [M,K] row-major, B = W [K,N] row-major;  C = A B + bias[n], ReLU if relu
//   MODE 1 (backward data) A = D [M,K'] row-major (K' = reduction = layer outputs), B = W [N,K'] row-major (used transposed);
//                          C[m,n] = sum_k D[m,k] W[n,k], multiplied by [mask[m,n] > 0] when mask != nullptr
//   MODE 2 (backward weight) A = H [K',M] (used transposed: reduction over the batch K'), B = D [K',N];  C[m,n] = sum_k H[k,m] D[k,n]
template <int MODE, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(NT) sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B,
                                                   int ldb, float* __restrict__ C, int ldc, const float* __restrict__ bias, int relu,
                                                   const float* __restrict__ mask, int ldm) {
    static_assert((BM / TM) * (BN / TN) == NT, "tile / thread shape");
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += BK) {
        // stage the two tiles k-major
#pragma unroll
        for (int e = tid; e < BM * BK; e += NT) {
            if (MODE == 2) {  // A element (m, k) = H[k][m]: consecutive threads walk m (contiguous in memory)
                const int mm = e % BM, kk = e / BM, m = m0 + mm, k = k0 + kk;
                As[kk][mm] = (m < M && k < K) ? A[static_cast<size_t>(k) * lda + m] : 0.f;
            } else {          // A element (m, k) = A[m][k]: consecutive threads walk k
                const int kk = e % BK, mm = e / BK, m = m0 + mm, k = k0 + kk;
                As[kk][mm] = (m < M && k < K) ? A[static_cast<size_t>(m) * lda + k] : 0.f;
            }
        }
#pragma unroll
        for (int e = tid; e < BN * BK; e += NT) {
            if (MODE == 1) {  // B element (k, n) = W[n][k]
                const int kk = e % BK, nn = e / BK, n = n0 + nn, k = k0 + kk;
                Bs[kk][nn] = (n < N && k < K) ? B[static_cast<size_t>(n) * ldb + k] : 0.f;
            } else {
                const int nn = e % BN, kk = e / BN, n = n0 + nn, k = k0 + kk;
                Bs[kk][nn] = (n < N && k < K) ? B[static_cast<size_t>(k) * ldb + n] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (MODE == 0) {
                v += bias[n];
                if (relu) v = fmaxf(v, 0.f);
            }
            if (MODE == 1 && mask) v = mask[static_cast<size_t>(m) * ldm + n] > 0.f ? v : 0.f;
            C[static_cast<size_t>(m) * ldc + n] = v;
        }
    }
}

// 32 x 32 tiles unless the problem is large enough to fill the GPU with 64 x 64 ones
template <int MODE>
inline void sgemm(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias, int relu,
                  const float* mask, int ldm, cudaStream_t st) {
    const long long big_blocks = static_cast<long long>((M + 63) / 64) * ((N + 63) / 64);
    if (big_blocks >= 296) {
        sgemm_kernel<MODE, 64, 64, 4, 4><<<dim3((N + 63) / 64, (M + 63) / 64), NT, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, relu, mask, ldm);
    } else {
        sgemm_kernel<MODE, 32, 32, 2, 2><<<dim3((N + 31) / 32, (M + 31) / 32), NT, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, relu, mask, ldm);
    }
}

// rows idx[i] (or first + i when idx == nullptr) of the resident set -> contiguous batch buffers
// (step != nullptr: the kernel runs inside a replayed CUDA graph and takes positions first .. first + batch of batch number *step of
// the permutation `idx`, batches being `stride` positions apart: stride = batch, first = 0 on one GPU; a data-parallel rank takes
// its share [first, first + batch) of every global batch of `stride` rows)
__global__ void gather_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ W, const int* __restrict__ idx,
                              long long first, int batch, int nx, int ny, float* __restrict__ xb, float* __restrict__ yb, float* __restrict__ wb,
                              const int* __restrict__ step, int stride) {
    const int row = blockIdx.x;
    if (row >= batch) return;
    const long long src = step ? idx[static_cast<long long>(*step) * stride + first + row] : (idx ? idx[row] : first + row);
    for (int j = threadIdx.x; j < ny; j += blockDim.x) yb[static_cast<size_t>(row) * ny + j] = Y[src * ny + j];
    for (int j = threadIdx.x; j < nx; j += blockDim.x) xb[static_cast<size_t>(row) * nx + j] = X[src * nx + j];
    if (threadIdx.x == 0) wb[row] = W[src];
}

// One warp per row: loss_i = mean_j (y - p)^2 * w_i (w_i = 1 / amplitude_i^2, emulator.py:70-80);
// delta[i][j] = dL/dp = -2 (y - p) w_i * gscale with gscale = 1 / (n_out * global_batch) (the batch-mean loss Keras minimises).
__global__ void loss_delta_kernel(const float* __restrict__ pred, const float* __restrict__ y, const float* __restrict__ w, int batch, int n,
                                  float gscale, float* __restrict__ delta, float* __restrict__ loss_rows) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= batch) return;
    const float wi = w[row];
    float s = 0.f;
    for (int j = lane; j < n; j += 32) {
        const float d = y[static_cast<size_t>(row) * n + j] - pred[static_cast<size_t>(row) * n + j];
        s = fmaf(d, d, s);
        if (delta) delta[static_cast<size_t>(row) * n + j] = -2.f * d * wi * gscale;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) loss_rows[row] = s / static_cast<float>(n) * wi;
}

// out[0] += sum_i loss_rows[i] in a fixed order (one warp, strided partial sums then a shuffle tree)
__global__ void loss_sum_kernel(const float* __restrict__ loss_rows, int batch, float* __restrict__ out) {
    float s = 0.f;
    for (int i = threadIdx.x; i < batch; i += 32) s += loss_rows[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[0] += s;
}

// db[n] = sum_m D[m][n]; one thread per column, rows in order
__global__ void colsum_kernel(const float* __restrict__ D, int M, int N, float* __restrict__ db) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += D[static_cast<size_t>(m) * N + n];
    db[n] = s;
}

// Keras-2.x Adam (non-amsgrad): m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2; p -= lr_t m / (sqrt(v) + eps)
// (lr_arr != nullptr: replayed graph, the step's learning rate is lr_arr[*step])
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g, long long n,
                            float lr_t, float b1, float b2, float eps, const float* __restrict__ lr_arr, const int* __restrict__ step) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    if (lr_arr) lr_t = lr_arr[*step];
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

__global__ void step_inc_kernel(int* step) { *step += 1; }
__global__ void add_scalar_kernel(float* dst, const float* src) { dst[0] += src[0]; }

}  // namespace trk
