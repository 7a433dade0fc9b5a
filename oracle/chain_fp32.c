/* CPU ORACLE -- test infrastructure only, never a product path (only tests/ load this; the product never does).
 *
 * Plain-C restatement of the Dense chain behind DirectEmulator.predict
 * (/root/reference/VeryAccurateEmulator/emulator.py:37-47 builds it, :402 evaluates it):
 *     h_{l+1} = act_l(h_l . kernel_l + bias_l),   kernel [in, out] row-major, ReLU on flagged layers only,
 * in float32 with ONE fixed arithmetic order: k ascending, every term a true fused multiply-add (fmaf: one rounding), the
 * accumulator starting at +0, the bias added afterwards with its own rounding, then max(., 0).  This is the order the
 * repository's FP32 CUDA kernel uses, and one member of the family of orders TensorFlow's CPU SGEMM may take (see
 * oracle/refmath.py::dense_chain_fp32_ordered, whose "seq_fma" order emulates fmaf through float64 and is checked against this
 * file in tests/test_oracle.py).  A third implementation of the chain, sharing no code with the numpy and torch ones.
 *
 * Measured on a B200: the FP32 CUDA kernel reproduces this function bit for bit (tests/test_zz_fp32_vs_c_oracle.py).
 *
 * Parity status: as oracle/refmath.py -- pinned to float64 known answers from the reference's shipped weights; TensorFlow's own
 * output is unavailable (absent from the image), so bit-level parity with it is unpinned.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o oracle/_build/libchain_fp32.so oracle/chain_fp32.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* returns 0, or 1 when out of memory / bad arguments */
int oracle_chain_fp32_seq_fma(const float* x, long n, int n_layers, const int* dims, const float* const* kernels,
                              const float* const* biases, const int* relu, float* out) {
    if (!x || !dims || !kernels || !biases || !relu || !out || n < 0 || n_layers < 1) return 1;
    int widest = 0;
    for (int l = 0; l <= n_layers; ++l)
        if (dims[l] > widest) widest = dims[l];
    float* a = (float*)malloc(sizeof(float) * (size_t)widest);
    float* b = (float*)malloc(sizeof(float) * (size_t)widest);
    if (!a || !b) {
        free(a);
        free(b);
        return 1;
    }
    for (long r = 0; r < n; ++r) {
        memcpy(a, x + r * dims[0], sizeof(float) * (size_t)dims[0]);
        for (int l = 0; l < n_layers; ++l) {
            const int K = dims[l], N = dims[l + 1];
            const float* W = kernels[l];
            for (int j = 0; j < N; ++j) {
                float acc = 0.0f;
                for (int k = 0; k < K; ++k) acc = fmaf(a[k], W[(size_t)k * N + j], acc);
                acc = acc + biases[l][j];
                if (relu[l]) acc = acc < 0.0f ? 0.0f : acc; /* NaN (and -0) pass through, like tf.nn.relu / np.maximum */
                b[j] = acc;
            }
            float* t = a;
            a = b;
            b = t;
        }
        memcpy(out + r * dims[n_layers], a, sizeof(float) * (size_t)dims[n_layers]);
    }
    free(a);
    free(b);
    return 0;
}
