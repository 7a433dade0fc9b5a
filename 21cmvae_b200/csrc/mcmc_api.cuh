// Ensemble MCMC around the fused chi^2 kernel (BASELINE config 4: emcee-style stretch move, 1e5 walkers).  Included by vae21_api.cu.
//
// The reference has no sampler; its README points MCMC users at `DirectEmulator.predict` inside their own likelihood.  What a
// stretch-move step needs from this library is "emulate + chi^2 for half of the ensemble", twice per step -- so the move itself
// lives here too, as two small element-wise kernels around each fused chi^2 launch, and a whole run of steps is ONE library call
// with nothing but the acceptance count leaving the GPU:
//   propose (one thread per walker of the active half): partner j and stretch z from a counter-based generator,
//            y = x_c[j] + z (x_s[i] - x_c[j]) in the coordinates of the prior box, normalised like preprocess.py:105-108
//   chi^2   the tensor-core (or FP32) kernel on the normalised proposals, one float per walker
//   accept  ln r = (d - 1) ln z + ln p(y) - ln p(x_s[i]); x_s[i] <- y with probability min(1, r)
// Goodman & Weare (2010) red/blue update: the two halves are updated in turn, each against the other's current positions.
// The generator is a stateless hash of (seed, step, half, walker, draw), restated in oracle/mcmc_ref.py.
#pragma once
#include <cmath>

namespace mck {

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z) {  // splitmix64 finaliser
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
// uniform in [0, 1) with 53 random bits
__host__ __device__ __forceinline__ double u01(unsigned long long seed, unsigned long long step, unsigned half, unsigned long long walker,
                                               unsigned draw) {
    const unsigned long long k = mix64(mix64(mix64(seed) ^ (step * 2ull + half)) ^ (walker * 4ull + draw));
    return static_cast<double>(k >> 11) * (1.0 / 9007199254740992.0);
}

struct Box {
    double lo[VAE21_MAX_PAR], hi[VAE21_MAX_PAR];      // prior box, in the coordinates the walkers live in (log10 on the masked columns)
    double pmin[VAE21_MAX_PAR], prange[VAE21_MAX_PAR];  // the emulator's normalisation (training-set range, same coordinates)
};

__global__ void propose_kernel(const double* __restrict__ x, int d, long long s0, long long c0, long long m, double a,
                               unsigned long long seed, unsigned long long step, unsigned half, const __grid_constant__ Box box,
                               double* __restrict__ y, float* __restrict__ xn, double* __restrict__ zterm, unsigned char* __restrict__ inside) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= m) return;
    long long j = static_cast<long long>(u01(seed, step, half, static_cast<unsigned long long>(i), 0) * static_cast<double>(m));
    if (j >= m) j = m - 1;
    const double r = __dadd_rn(__dmul_rn(a - 1.0, u01(seed, step, half, static_cast<unsigned long long>(i), 1)), 1.0);
    const double z = __ddiv_rn(__dmul_rn(r, r), a);
    const double* xs = x + (s0 + i) * d;
    const double* xc = x + (c0 + j) * d;
    bool in = true;
    for (int k = 0; k < d; ++k) {
        const double yk = __dadd_rn(xc[k], __dmul_rn(z, __dadd_rn(xs[k], -xc[k])));  // no fma contraction: numpy-reproducible
        y[i * d + k] = yk;
        in = in && yk >= box.lo[k] && yk <= box.hi[k];
        double t = __dadd_rn(yk, -box.pmin[k]);
        t = __ddiv_rn(t, box.prange[k]);
        t = __dmul_rn(t, 2.0);
        t = __dadd_rn(t, -1.0);
        xn[i * d + k] = static_cast<float>(t);
    }
    zterm[i] = static_cast<double>(d - 1) * log(z);
    inside[i] = in ? 1 : 0;
}

__global__ void accept_kernel(double* __restrict__ x, double* __restrict__ logp, int d, long long s0, long long m, const double* __restrict__ y,
                              const float* __restrict__ chi, const double* __restrict__ zterm, const unsigned char* __restrict__ inside,
                              unsigned long long seed, unsigned long long step, unsigned half, unsigned long long* __restrict__ n_acc) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    bool acc = false;
    if (i < m) {
        const double lpy = inside[i] ? -0.5 * static_cast<double>(chi[i]) : -INFINITY;
        const double lnr = zterm[i] + lpy - logp[s0 + i];
        acc = log(u01(seed, step, half, static_cast<unsigned long long>(i), 2)) < lnr;  // NaN chi^2 never accepts
        if (acc) {
            for (int k = 0; k < d; ++k) x[(s0 + i) * d + k] = y[i * d + k];
            logp[s0 + i] = lpy;
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(n_acc, static_cast<unsigned long long>(__popc(ballot)));
}

// ln p of the current positions (run start): normalise, then the caller launches chi^2 and `init_logp_kernel`
__global__ void normalise_kernel(const double* __restrict__ x, int d, long long n, const __grid_constant__ Box box, float* __restrict__ xn,
                                 unsigned char* __restrict__ inside) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    bool in = true;
    for (int k = 0; k < d; ++k) {
        const double v = x[i * d + k];
        in = in && v >= box.lo[k] && v <= box.hi[k];
        double t = __dadd_rn(v, -box.pmin[k]);
        t = __ddiv_rn(t, box.prange[k]);
        t = __dmul_rn(t, 2.0);
        t = __dadd_rn(t, -1.0);
        xn[i * d + k] = static_cast<float>(t);
    }
    inside[i] = in ? 1 : 0;
}
__global__ void init_logp_kernel(double* __restrict__ logp, long long n, const float* __restrict__ chi, const unsigned char* __restrict__ inside) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i < n) logp[i] = inside[i] ? -0.5 * static_cast<double>(chi[i]) : -INFINITY;
}

}  // namespace mck

extern "C" {

int vae21_mcmc_run(vae21_handle* h, double* x_dev, double* logp_dev, int64_t n_walkers, int n_dim, const double* lo, const double* hi,
                   const float* obs, const float* inv_sigma, double a, uint64_t seed, int64_t first_step, int n_steps, int init_logp,
                   int precision, void* stream, int64_t* n_accepted) {
    if (!h || !x_dev || !logp_dev || !lo || !hi || !obs || !inv_sigma) return fail(VAE21_ERR_ARG, "null argument");
    if (!h->model_set || !h->norm_set) return fail(VAE21_ERR_STATE, "vae21_set_model / vae21_set_norm have not been called");
    if (n_dim != h->dims[0]) return fail(VAE21_ERR_ARG, "walkers have %d coordinates, the model %d inputs", n_dim, h->dims[0]);
    if (n_walkers < 2 || (n_walkers & 1)) return fail(VAE21_ERR_ARG, "the ensemble needs an even number of walkers >= 2 (got %lld)", (long long)n_walkers);
    if (!(a > 1.0) || n_steps < 0 || first_step < 0) return fail(VAE21_ERR_ARG, "bad stretch scale / step range");
    if (int rc = use_device(h)) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int NO = h->dims[h->n_layers];
    const long long m = n_walkers / 2;
    if (int rc = upload_observation(h, obs, inv_sigma, NO, st)) return rc;
    // scratch: proposals, normalised proposals, stretch terms, inside flags, chi^2 (sized for the whole ensemble: the start-up pass)
    const size_t need = static_cast<size_t>(n_walkers) * (n_dim * (sizeof(double) + sizeof(float)) + sizeof(double) + sizeof(float) + 1) + 64;
    if (h->mcmc_cap < need) {
        if (h->d_mcmc) cudaFree(h->d_mcmc);
        h->d_mcmc = nullptr;
        h->mcmc_cap = 0;
        CK(cudaMalloc(&h->d_mcmc, need));
        h->mcmc_cap = need;
    }
    uint8_t* base = static_cast<uint8_t*>(h->d_mcmc);
    double* d_y = reinterpret_cast<double*>(base);
    double* d_z = d_y + static_cast<size_t>(n_walkers) * n_dim;
    unsigned long long* d_acc = reinterpret_cast<unsigned long long*>(d_z + n_walkers);
    float* d_xn = reinterpret_cast<float*>(d_acc + 2);
    float* d_chi = d_xn + static_cast<size_t>(n_walkers) * n_dim;
    unsigned char* d_in = reinterpret_cast<unsigned char*>(d_chi + n_walkers);
    mck::Box box{};
    for (int k = 0; k < n_dim; ++k) {
        box.lo[k] = lo[k];
        box.hi[k] = hi[k];
        box.pmin[k] = h->nc.pmin[k];
        box.prange[k] = h->nc.prange[k];
    }
    LaunchArgs la{};
    la.in = d_xn;
    la.mu = h->d_mu;
    la.obs = h->d_obs;
    la.isig = h->d_isig;
    la.chi2 = d_chi;
    la.in_mode = IN_NORMALISED_F32;
    la.out_mode = OUT_CHI2;
    const int TB = 256;
    if (init_logp) {
        mck::normalise_kernel<<<static_cast<unsigned>((n_walkers + TB - 1) / TB), TB, 0, st>>>(x_dev, n_dim, n_walkers, box, d_xn, d_in);
        la.n = n_walkers;
        if (int rc = launch(h, la, precision, st)) return rc;
        mck::init_logp_kernel<<<static_cast<unsigned>((n_walkers + TB - 1) / TB), TB, 0, st>>>(logp_dev, n_walkers, d_chi, d_in);
    }
    CK(cudaMemsetAsync(d_acc, 0, sizeof(unsigned long long), st));
    la.n = m;
    const unsigned grid = static_cast<unsigned>((m + TB - 1) / TB);
    for (int s = 0; s < n_steps; ++s) {
        const unsigned long long step = static_cast<unsigned long long>(first_step + s);
        for (unsigned half = 0; half < 2; ++half) {
            const long long s0 = half ? m : 0, c0 = half ? 0 : m;
            mck::propose_kernel<<<grid, TB, 0, st>>>(x_dev, n_dim, s0, c0, m, a, seed, step, half, box, d_y, d_xn, d_z, d_in);
            if (int rc = launch(h, la, precision, st)) return rc;
            mck::accept_kernel<<<grid, TB, 0, st>>>(x_dev, logp_dev, n_dim, s0, m, d_y, d_chi, d_z, d_in, seed, step, half, d_acc);
        }
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev_use, st));
    h->use_pending = true;
    if (n_accepted) {
        unsigned long long acc = 0;
        CK(cudaMemcpyAsync(&acc, d_acc, sizeof acc, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        *n_accepted = static_cast<int64_t>(acc);
    }
    return 0;
}

}  // extern "C"
