/*
 * vae21.h -- C ABI of the B200-native 21cmVAE emulator-evaluation library.
 *
 * The reference (christianhbye/21cmVAE) has no FFI: its seam is the Python
 * method DirectEmulator.predict.  Everything between
 *   VeryAccurateEmulator/emulator.py:401  pp.par_transform(params, self.par_train)
 *   VeryAccurateEmulator/emulator.py:402  self.emulator.predict(transformed_params)
 *   VeryAccurateEmulator/emulator.py:403  pp.unpreproc(proc_pred, self.signal_train)
 * becomes ONE call into this library (vae21_predict); the squeeze rule at
 * emulator.py:404-407 stays in Python.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns an int status: 0 = OK, non-zero = VAE21_ERR_*;
 *     vae21_last_error() returns a thread-local message for the last failure.
 *   - no exceptions or aborts cross the ABI.
 *   - the caller owns every buffer it passes; the library owns the opaque
 *     handle, its device copies of weights/constants and its scratch buffers.
 *   - one handle per (process, GPU); calls on one handle must be serialised
 *     by the caller; distinct handles are independent.
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with VAE21_ERR_CUDA.
 */
#ifndef VAE21_H_
#define VAE21_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAE21_VERSION 100 /* 0.1.0 */

/* status codes */
#define VAE21_OK 0
#define VAE21_ERR_ARG 1         /* bad argument (null pointer, bad enum, n < 0 ...) */
#define VAE21_ERR_STATE 2       /* model or normalisation constants not set */
#define VAE21_ERR_CUDA 3        /* CUDA runtime error (message has the CUDA string) */
#define VAE21_ERR_UNSUPPORTED 4 /* layer stack does not fit the requested kernel */
#define VAE21_ERR_NOMEM 5

/* params_dtype */
#define VAE21_F32 0
#define VAE21_F64 1

/* precision: which kernel evaluates the Dense chain */
#define VAE21_FP32_SIMT 0   /* FFMA, fp32 accumulate: parity path (<= 1e-5 of amplitude vs TF CPU) */
#define VAE21_TC_BF16X3 1   /* tcgen05 kind::f16, 3-pass bf16 hi/lo split, fp32 accumulate in TMEM */
#define VAE21_TC_FP16X3 2   /* same with fp16 hi/lo split (fp32-class error; inputs must stay < 65504) */
#define VAE21_TC_FP16E4M3 3 /* fp16 hi x hi (kind::f16) + both first-order correction terms as ONE e4m3 kind::f8f6f4 MMA
                               (K = 32): 2 MMAs per k-step instead of 3; error ~2.5x bf16x3, inside 0.01 mK rms / 0.05 mK max */

typedef struct vae21_handle vae21_handle;

int vae21_version(void);
const char* vae21_last_error(void);
int vae21_device_count(int* count);

/* Handle life cycle.  `device` is a CUDA ordinal. */
int vae21_create(int device, vae21_handle** out);
int vae21_destroy(vae21_handle* h);

/*
 * The Dense stack (replaces the tf.keras.Sequential built at emulator.py:37-47
 * and loaded at emulator.py:333-337).
 *   dims[n_layers + 1]     layer widths, dims[0] = number of parameters
 *   kernels[l]             host pointer, fp32 row-major [dims[l], dims[l+1]] (Keras kernel layout)
 *   biases[l]              host pointer, fp32 [dims[l+1]]
 *   relu_flags[l]          1 = ReLU after layer l, 0 = linear
 * Weights are copied, padded and (for the tensor-core path) split/packed on
 * upload; the host arrays may be freed afterwards.
 */
int vae21_set_model(vae21_handle* h, int n_layers, const int* dims, const float* const* kernels,
                    const float* const* biases, const int* relu_flags);

/*
 * Normalisation constants (replace the per-call statistics of
 * preprocess.py:89-101 and :44-45).
 *   par_min/par_max[n_par] min/max of the log-transformed training parameters (fp64)
 *   log_mask[n_par]        1 = take log10 of this column first (columns 0..2 in the reference)
 *   fx_floor               value substituted for an exact 0 in column `floor_col` (1e-6, column 2)
 *   sig_mean[n_out]        per-bin training mean;  sig_std: scalar training std
 */
int vae21_set_norm(vae21_handle* h, int n_par, const double* par_min, const double* par_max, const int* log_mask,
                   int floor_col, double fx_floor, int n_out, const float* sig_mean, float sig_std);

/*
 * DirectEmulator.predict (emulator.py:383-407 minus the squeeze rule).
 *   params   [n, dims[0]] row-major, fp32 or fp64, host or device memory
 *   out      [n, dims[last]] row-major fp32, host or device memory
 *   stream   cudaStream_t used when BOTH buffers are on the device (the call
 *            is then asynchronous); host buffers use the library's own
 *            copy/compute pipeline and the call returns when `out` is complete.
 */
int vae21_predict(vae21_handle* h, const void* params, int params_dtype, int params_on_device, int64_t n, float* out,
                  int out_on_device, int precision, void* stream);

/*
 * emu.emulator.predict(x): the bare Dense stack on already-normalised inputs
 * (emulator.py:402), output in sigma units (no de-normalisation).
 */
int vae21_forward_normalised(vae21_handle* h, const float* x, int x_on_device, int64_t n, float* y, int y_on_device,
                             int precision, void* stream);

/*
 * Fused likelihood: chi2[i] = sum_k ((predict(params_i)[k] - obs[k]) * inv_sigma[k])^2
 * without writing the spectra.  obs / inv_sigma are HOST pointers ([n_out],
 * copied on each call).  chi2 may be NULL when only the argmin is wanted.
 * best_val/best_idx (host pointers, may be NULL) receive the minimum over the
 * n rows and its row index (-1 if every chi2 is NaN).  Not in the reference
 * (callers do this in numpy on the 451-bin output).
 */
int vae21_chi2(vae21_handle* h, const void* params, int params_dtype, int params_on_device, int64_t n,
               const float* obs, const float* inv_sigma, float* chi2, int chi2_on_device, float* best_val,
               int64_t* best_idx, int precision, void* stream);

/*
 * Fused likelihood over a REGULAR GRID generated on the device (BASELINE config 3: 1e8-point grids are never materialised):
 * point i (C order, last dimension fastest) has NORMALISED coordinates x_j = x_lo[j] + i_j (x_hi[j] - x_lo[j]) / (npts[j] - 1)
 * -- the [-1, 1] box is the training range of preprocess.par_transform (preprocess.py:105-108).  Evaluates points
 * [first, first + count) (first + count < 2^32 per call; shard larger grids), writes chi2_dev[i - first] (DEVICE pointer, may be
 * NULL) and returns the minimum and its global grid index.  obs / inv_sigma: host pointers as in vae21_chi2.
 */
int vae21_chi2_grid(vae21_handle* h, int n_dim, const int* npts, const double* x_lo, const double* x_hi, int64_t first, int64_t count,
                    const float* obs, const float* inv_sigma, float* chi2_dev, float* best_val, int64_t* best_idx, int precision,
                    void* stream);

/*
 * Ensemble MCMC (BASELINE config 4; the reference has no sampler -- its users put DirectEmulator.predict, emulator.py:383-407, inside
 * their own likelihood): n_steps stretch-move steps (Goodman & Weare 2010, scale a, red/blue halves) of an ensemble of n_walkers
 * (even) walkers with ln p = -chi^2 / 2 inside the box [lo, hi] and -inf outside, entirely on the device: per half-step one
 * proposal kernel, one fused emulate + chi^2 launch on the proposals, one accept kernel.
 *   x_dev [n_walkers, n_dim] float64 DEVICE: positions in the coordinates of preprocess.par_transform's box (log10 on the masked
 *     columns), updated in place;  logp_dev [n_walkers] float64 DEVICE: ln p of the positions, updated in place (computed first when
 *     init_logp != 0);  lo / hi: host, same coordinates;  obs / inv_sigma: host, as in vae21_chi2.
 *   Random numbers are a stateless hash of (seed, first_step + s, half, walker, draw): a run can be continued by passing the next
 *   first_step, and oracle/mcmc_ref.py restates the generator.  n_accepted (host, may be NULL: then the call does not synchronise)
 *   receives the number of accepted proposals of this call.
 */
int vae21_mcmc_run(vae21_handle* h, double* x_dev, double* logp_dev, int64_t n_walkers, int n_dim, const double* lo, const double* hi,
                   const float* obs, const float* inv_sigma, double a, uint64_t seed, int64_t first_step, int n_steps, int init_logp,
                   int precision, void* stream, int64_t* n_accepted);

/*
 * Host only (no GPU needed): plan the tensor-core schedule of a Dense stack dims[0] -> ... -> dims[n_layers] (ReLU on all but the last
 * layer) and replay its issue table against the epilogue's barrier arrivals for several tiles (phase parities, commit counts, k-step
 * coverage).  Returns 0 when the stack has a consistent schedule, 1 when it does not fit the tensor-core kernel, 2 when the planner
 * produced an inconsistent schedule (a bug); msg receives the reason.  Used by the CPU tests; mirrors nothing in the reference.
 */
int vae21_check_plan(int n_layers, const int* dims, char* msg, int msg_len);

/*
 * Fused figure of merit (emulator.py:129-192 `error`, :409-439 `test_error`): err[i] = sqrt(mean_k (predict(params_i)[k] -
 * truth[i][k])^2) over the bins with band_mask[k] != 0 (NULL = all bins), in mK; with relative != 0 divided by max_k |truth[i][k]|
 * over the same bins and multiplied by 100 (%).  The 451-bin predictions never leave the GPU.  params / truth / err may each be
 * host or device pointers; band_mask is a host pointer.  Synchronous.
 */
int vae21_error(vae21_handle* h, const void* params, int params_dtype, int params_on_device, int64_t n, const float* truth,
                int truth_on_device, const float* band_mask, int relative, float* err, int err_on_device, int precision, void* stream);

/* Pinned host memory from the library's caching pool (for PCIe-rate copies). */
void* vae21_host_alloc(size_t bytes);
void vae21_host_free(void* p);
void vae21_host_trim(void); /* release cached pinned blocks */

/*
 * Introspection for benchmarks and tests.
 *   kernel_launches      number of the library's own kernels launched on this handle so far
 *   last_kernel_ms       device time of the compute kernels of the last call that used the
 *                        internal pipeline (CUDA events); <0 if not measured
 *   tc_supported         1 if the loaded model fits the tensor-core kernel
 */
int vae21_get_info(vae21_handle* h, int64_t* kernel_launches, float* last_kernel_ms, int* tc_supported);

/*
 * Operand-range statistics of the tensor-core paths.  The fp16-based operand formats have a finite range: VAE21_TC_FP16X3 splits
 * hidden activations into fp16 hi/lo (|h| must stay below 65504), VAE21_TC_FP16E4M3 carries its first-order corrections in e4m3
 * (saturating at 448: beyond it the corrections lose accuracy and the result degrades towards a one-pass fp16 product).  Every
 * launch adds the number of epilogue threads that converted a hidden activation beyond the range to a per-handle counter;
 * *saturated receives it (0 = every launch since the last reset stayed inside the range the error budget was pinned for).
 * Synchronises the device.  VAE21_TC_BF16X3 and the FP32 path have no such limit.
 */
int vae21_get_tc_stats(vae21_handle* h, int64_t* saturated, int reset);

/*
 * Benchmark helper: run the predict kernel `iters` times back to back on
 * device-resident buffers and return the mean device time per launch in ms,
 * measured with CUDA events on the launching stream.
 */
int vae21_time_predict(vae21_handle* h, const void* params_dev, int params_dtype, int64_t n, float* out_dev,
                       int precision, int iters, float* ms_per_launch);

/*
 * ---- training (replaces the Keras `fit` behind DirectEmulator.train, emulator.py:339-381) ---------------
 * A trainer owns the fp32 parameters of a Dense stack in Keras `get_weights()` order (per layer: kernel
 * [in,out] row-major, then bias), the Adam moments, and batch workspaces.  One optimisation step is
 *   1. vae21_trainer_forward_backward: gather the batch rows idx[0..batch) (or first..first+batch when idx is NULL) from the
 *      device-resident set x_all [n, n_in], y_all [n, n_out], w_all [n]; forward; loss_i = w_i * mean_k (y_ik - p_ik)^2 (the
 *      relative MSE of emulator.py:51-83 with w_i = 1 / amplitude_i^2); backward.  grad [num_params] receives
 *      d/dparams of grad_scale * n_out * sum_i loss_i: pass grad_scale = 1 / (n_out * rows of the GLOBAL batch) and the sum of
 *      the ranks' gradients is the gradient of the batch-mean loss Keras minimises.  loss_sum[0] += sum_i loss_i.
 *      grad == NULL: forward + loss only (validation).
 *   2. [data parallel: all-reduce (sum) `grad` over the ranks]
 *   3. vae21_trainer_adam: Keras Adam; the caller passes lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) of update number t.
 * All pointers except `out` / `n` / flat_host are DEVICE pointers.
 */
typedef struct vae21_trainer vae21_trainer;
int vae21_trainer_create(int device, int n_layers, const int* dims, const int* relu_flags, int max_batch, vae21_trainer** out);
int vae21_trainer_destroy(vae21_trainer* t);
int vae21_trainer_num_params(vae21_trainer* t, int64_t* n);
int vae21_trainer_set_params(vae21_trainer* t, const float* flat_host, int reset_moments);
int vae21_trainer_get_params(vae21_trainer* t, float* flat_host);
/* Adam slot variables (first / second moment), flat in the order of the parameters: what Keras keeps in a saved model's
 * `optimizer_weights` group (`Adam/<layer>/kernel/m:0` ...), so that a retrained model continues where it stopped. */
int vae21_trainer_set_moments(vae21_trainer* t, const float* m_host, const float* v_host);
int vae21_trainer_get_moments(vae21_trainer* t, float* m_host, float* v_host);
int vae21_trainer_forward_backward(vae21_trainer* t, const float* x_all, const float* y_all, const float* w_all, const int* idx,
                                   int64_t first, int batch, float grad_scale, float* grad, float* loss_sum, void* stream);
int vae21_trainer_adam(vae21_trainer* t, const float* grad, float lr_t, float beta1, float beta2, float eps, void* stream);
/* One epoch on ONE GPU in one call: for every batch of `batch` rows of perm[0..n) (device int32 permutation, NULL = natural
 * order) steps 1 and 3 above with grad_scale = 1 / (n_out * rows) and update numbers iterations_before + 1, + 2, ...  Full batches
 * replay one captured CUDA graph (batch number and learning rate are read from device memory).  Bitwise identical to the calls. */
int vae21_trainer_epoch(vae21_trainer* t, const float* x_all, const float* y_all, const float* w_all, const int* perm, int64_t n, int batch,
                        float lr, float beta1, float beta2, float eps, int64_t iterations_before, float* loss_sum, void* stream);
/* Data-parallel epoch on one rank: the step is TWO replayable CUDA graphs around the caller's all-reduce of `grad` (NCCL lives in the
 * host runtime, not in this library).  vae21_trainer_dp_begin copies the epoch's permutation and learning rates (update numbers
 * iterations_before + 1 ...) to the device, resets the step counter and captures -- once per (buffers, batch, share) -- graph A:
 * rows perm[k*batch + share_first .. + share_rows) of step k -> forward, loss (loss_sum += ...), backward into `grad` scaled by
 * 1 / (n_out * batch), and graph B: Adam from `grad` with the step's learning rate, step counter += 1.  Per FULL batch k the caller
 * runs: vae21_trainer_dp_forward_backward (skip it and zero `grad` when share_rows == 0), all-reduce, vae21_trainer_dp_adam.  The
 * trailing short batch goes through vae21_trainer_forward_backward / vae21_trainer_adam.  Same arithmetic as those calls. */
int vae21_trainer_dp_begin(vae21_trainer* t, const float* x_all, const float* y_all, const float* w_all, const int* perm, int64_t n, int batch,
                           int share_first, int share_rows, float lr, float beta1, float beta2, float eps, int64_t iterations_before,
                           float* grad, float* loss_sum, void* stream);
int vae21_trainer_dp_forward_backward(vae21_trainer* t, void* stream);
int vae21_trainer_dp_adam(vae21_trainer* t, void* stream);
int vae21_trainer_launches(vae21_trainer* t, int64_t* n);

#ifdef __cplusplus
}
#endif
#endif /* VAE21_H_ */
