// Shared definitions for the vae21 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define VAE21_MAX_LAYERS 12
#define VAE21_MAX_PAR 16

// what the kernel reads
enum : int { IN_PARAMS_F32 = 0, IN_PARAMS_F64 = 1, IN_NORMALISED_F32 = 2, IN_GRID = 3 };
// what the kernel writes
enum : int { OUT_PREDICT = 0, OUT_NORMALISED = 1, OUT_CHI2 = 2, OUT_ERROR = 3 };

// Prologue constants: x_j = ((T(p_j) - pmin_j) / prange_j) * 2 - 1 in fp64,
// T = log10 on masked columns with an exact 0 in `floor_col` replaced by
// `floor_val` (preprocess.py:74-78, :105-108).
struct NormConsts {
    double pmin[VAE21_MAX_PAR];
    double prange[VAE21_MAX_PAR];  // pmax - pmin, computed on the host in fp64 like numpy
    int log_mask[VAE21_MAX_PAR];
    int n_par;
    int floor_col;
    double floor_val;
    float sd;  // np.std(signal_train)
    double pscale[VAE21_MAX_PAR];  // 2 / prange (tensor-core prologue: one fma instead of the reference's divide)
};

struct LaunchArgs {
    const void* in;      // [n, K0] params (f32/f64) or normalised x (f32)
    float* out;          // [n, Nout] or nullptr
    const float* mu;     // [Nout] per-bin mean (device)
    const float* obs;    // [Nout] observed signal (device), chi2 mode
    const float* isig;   // [Nout] 1/sigma (device), chi2 mode
    float* chi2;         // [n] or nullptr
    unsigned long long* argmin_key;  // packed (float bits << 32 | row) running minimum, or nullptr
    unsigned long long* sat;         // tensor-core paths: += number of epilogue threads that converted a hidden activation beyond the
                                     // range of the operand format (fp16 hi: 65504; e4m3 corrections of fp16e4m3: 448); or nullptr
    // OUT_ERROR (emulator.py:129-192): chi2[r] = sqrt(mean_band((pred - truth[r])^2)) [* 100 / max_band |truth[r]| if err_relative];
    // the band is isig[k] in {0, 1}; err_inv_count = 1 / (number of bins in the band)
    const float* truth;  // [n, Nout] true signals (device)
    float err_inv_count;
    int err_relative;
    long long n;         // rows in this launch
    long long row_base;  // global index of row 0 (for argmin)
    int in_mode;
    int out_mode;
    // IN_GRID: no input array -- row r is point (row_base + r) of a regular grid in NORMALISED coordinates, C order (last
    // dimension fastest): x_j = grid_lo[j] + i_j * grid_step[j], i_j in [0, grid_n[j])
    int grid_n[VAE21_MAX_PAR];
    float grid_lo[VAE21_MAX_PAR];
    float grid_step[VAE21_MAX_PAR];
};

// coordinate j of grid point `idx` (see LaunchArgs)
__device__ __forceinline__ void grid_point(const LaunchArgs& a, unsigned long long idx, int n_dim, float* x) {
    for (int j = n_dim - 1; j >= 0; --j) {
        const unsigned long long q = idx / static_cast<unsigned>(a.grid_n[j]);
        const unsigned i = static_cast<unsigned>(idx - q * static_cast<unsigned>(a.grid_n[j]));
        x[j] = fmaf(static_cast<float>(i), a.grid_step[j], a.grid_lo[j]);
        idx = q;
    }
}

// ---- device helpers ---------------------------------------------------------

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ unsigned long long pack_min_key(float v, unsigned long long row) {
    // chi2 >= 0 or NaN: the IEEE bit pattern is monotone for non-negative floats and every NaN
    // pattern compares above +inf, so an unsigned min picks the smallest finite value.
    return (static_cast<unsigned long long>(__float_as_uint(v)) << 32) | (row & 0xffffffffull);
}

// The parameter transform of preprocess.py:74-78,105-108 for one element.
// `f32_in`: the caller's array was float32, where numpy does the floor substitution and the
// log10 in float32 before the float64 affine map -- round the same way.
__device__ __forceinline__ float transform_param(double p, int col, const NormConsts& nc, bool f32_in) {
    if (col == nc.floor_col && p == 0.0) p = f32_in ? static_cast<double>(static_cast<float>(nc.floor_val)) : nc.floor_val;
    double t = p;
    if (nc.log_mask[col]) {
        t = log10(p);
        if (f32_in) t = static_cast<double>(static_cast<float>(t));
    }
    t = t - nc.pmin[col];
    t = t / nc.prange[col];
    t = t * 2.0;
    t = t - 1.0;
    return static_cast<float>(t);  // Keras casts its input to float32
}
