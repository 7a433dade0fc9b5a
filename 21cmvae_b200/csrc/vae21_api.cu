// C-ABI of the vae21 library (see include/vae21.h): handle management, weight packing,
// the host<->device copy/compute pipeline and kernel dispatch.  sm_100a only.
#include "../../include/vae21.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "fp32_kernel.cuh"
#include "fp32_pipe_kernel.cuh"
#include "tc_kernel.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(VAE21_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ---- pinned host pool ---------------------------------------------------------------------
struct PinnedPool {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks;
    std::map<void*, size_t> live;
    static size_t round_up(size_t b) {
        const size_t g = b < (1u << 20) ? 4096 : (2u << 20);
        return (b + g - 1) / g * g;
    }
    void* alloc(size_t bytes) {
        if (bytes == 0) bytes = 1;
        const size_t sz = round_up(bytes);
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = free_blocks.lower_bound(sz);
            if (it != free_blocks.end() && it->first <= sz + sz / 4) {
                void* p = it->second;
                live[p] = it->first;
                free_blocks.erase(it);
                return p;
            }
        }
        void* p = nullptr;
        if (cudaHostAlloc(&p, sz, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            trim();
            if (cudaHostAlloc(&p, sz, cudaHostAllocPortable) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
        }
        std::lock_guard<std::mutex> lk(mu);
        live[p] = sz;
        return p;
    }
    void release(void* p) {
        if (!p) return;
        std::lock_guard<std::mutex> lk(mu);
        auto it = live.find(p);
        if (it == live.end()) return;
        free_blocks.emplace(it->second, p);
        live.erase(it);
    }
    void trim() {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& kv : free_blocks) cudaFreeHost(kv.second);
        free_blocks.clear();
    }
};
PinnedPool g_pool;

#ifndef VAE21_FP32_PIPE_DEFAULT
#define VAE21_FP32_PIPE_DEFAULT 1  // barrier-free FP32 kernel (bit-identical to the block-barrier one; 17.6 -> 14.9 ms per 1M rows)
#endif
constexpr int NSLOT = 3;              // pipeline depth of the host-buffer path
constexpr long long CHUNK_ROWS = 32768;  // rows per pipeline chunk (59 MB of output)

}  // namespace

struct vae21_handle {
    int device = 0;
    int sm_count = 0;
    bool model_set = false, norm_set = false;
    int n_layers = 0;
    int dims[VAE21_MAX_LAYERS + 1] = {0};
    // fp32 path
    f32k::Model f32{};
    float* d_w32 = nullptr;
    float* d_b32 = nullptr;
    size_t f32_smem = 0;
    int f32_wst = 3;
    // barrier-free variant (fp32_pipe_kernel.cuh): rows of its single in-place activation buffer and its shared-memory bytes
    int f32p_rows = 0, f32p_stages = 0;
    size_t f32p_smem = 0;
    // tensor-core path
    tck::Plan tc{};
    bool tc_ok = false;
    std::string tc_why;
    void* d_wtc[3] = {nullptr, nullptr, nullptr};  // packed operand images: [0] bf16 hi/lo, [1] fp16 hi/lo, [2] fp16 + e4m3
    float* d_btc = nullptr;
    // constants
    NormConsts nc{};
    float* d_mu = nullptr;
    float* d_obs = nullptr;
    float* d_isig = nullptr;
    float* d_mask = nullptr;  // band mask of vae21_error (its own buffer: d_isig keeps the cached observation)
    int alloc_out = 0;        // output width d_mu / d_obs / d_isig / d_mask are allocated for
    unsigned long long* d_key = nullptr;
    // host copies of what d_obs / d_isig hold: an MCMC or grid loop passes the same observation on every call, and two small
    // pageable-memory uploads per call are a visible part of a 45-microsecond launch
    std::vector<float> h_obs, h_isig;
    bool obs_cached = false;
    // pipeline
    cudaStream_t streams[NSLOT] = {nullptr, nullptr, nullptr};
    void* d_in[NSLOT] = {nullptr, nullptr, nullptr};
    float* d_out[NSLOT] = {nullptr, nullptr, nullptr};
    size_t cap_in[NSLOT] = {0, 0, 0}, cap_out[NSLOT] = {0, 0, 0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // ordering against the caller's stream: `ev_in` marks "the caller's device buffers are ready" for the pipeline streams,
    // `ev_use` the last kernel launched on a caller's stream (it reads d_mu / d_obs / d_isig / the weight images)
    cudaEvent_t ev_in = nullptr, ev_use = nullptr;
    bool use_pending = false;
    void* d_mcmc = nullptr;  // scratch of vae21_mcmc_run (proposals, stretch terms, chi^2)
    size_t mcmc_cap = 0;
    unsigned long long* d_sat = nullptr;  // [2] saturation counters of the tensor-core operand conversion (see vae21_get_info)
    long long launches = 0;
    float last_ms = -1.f;
};

namespace {

int use_device(vae21_handle* h) {
    CK(cudaSetDevice(h->device));
    return 0;
}

// A kernel launched asynchronously on a caller's stream may still be reading the handle's constants (d_mu, d_obs, d_isig, d_mask,
// weight images): wait for it before any of them is rewritten.
int wait_last_use(vae21_handle* h) {
    if (h->use_pending) {
        CK(cudaEventSynchronize(h->ev_use));
        h->use_pending = false;
    }
    return 0;
}

int ensure_slot(vae21_handle* h, int s, size_t in_bytes, size_t out_bytes) {
    if (in_bytes > h->cap_in[s]) {
        if (h->d_in[s]) CK(cudaFree(h->d_in[s]));
        h->d_in[s] = nullptr;
        h->cap_in[s] = 0;
        CK(cudaMalloc(&h->d_in[s], in_bytes));
        h->cap_in[s] = in_bytes;
    }
    if (out_bytes > h->cap_out[s]) {
        if (h->d_out[s]) CK(cudaFree(h->d_out[s]));
        h->d_out[s] = nullptr;
        h->cap_out[s] = 0;
        CK(cudaMalloc(&h->d_out[s], out_bytes));
        h->cap_out[s] = out_bytes;
    }
    return 0;
}

// ---- fp32 model packing -----------------------------------------------------------------
int pack_fp32(vae21_handle* h, const float* const* kernels, const float* const* biases, const int* relu) {
    f32k::Model& m = h->f32;
    m = f32k::Model{};
    m.n_layers = h->n_layers;
    long long woff = 0, boff = 0;
    m.buf_rows[0] = m.buf_rows[1] = 0;
    m.max_npad = 0;
    for (int l = 0; l < h->n_layers; ++l) {
        f32k::Layer& L = m.L[l];
        L.K = h->dims[l];
        L.N = h->dims[l + 1];
        L.Npad = (L.N + 31) / 32 * 32;
        // the input of layer l has the k extent of the previous layer's padded output (zeros there)
        L.Kpad = (L.K + f32k::KB - 1) / f32k::KB * f32k::KB;
        L.relu = relu[l] ? 1 : 0;
        L.w_off = woff;
        L.b_off = boff;
        woff += static_cast<long long>(L.Kpad) * L.Npad;
        boff += L.Npad;
        if (L.Npad / 32 > f32k::MAX_SLOTS)
            return fail(VAE21_ERR_UNSUPPORTED, "layer %d width %d exceeds the fp32 kernel's limit of %d", l, L.N,
                        32 * f32k::MAX_SLOTS);
        m.max_npad = std::max(m.max_npad, L.Npad);
        int& in_rows = m.buf_rows[l & 1];
        in_rows = std::max(in_rows, L.Kpad);
        if (l + 1 < h->n_layers) {
            int& out_rows = m.buf_rows[(l + 1) & 1];
            out_rows = std::max(out_rows, L.Npad);
        }
    }
    if (m.buf_rows[1] == 0) m.buf_rows[1] = f32k::KB;
    std::vector<float> W(woff, 0.f), B(boff, 0.f);
    for (int l = 0; l < h->n_layers; ++l) {
        const f32k::Layer& L = m.L[l];
        for (int k = 0; k < L.K; ++k)
            memcpy(&W[L.w_off + static_cast<long long>(k) * L.Npad], kernels[l] + static_cast<long long>(k) * L.N,
                   sizeof(float) * L.N);
        memcpy(&B[L.b_off], biases[l], sizeof(float) * L.N);
    }
    const size_t act = static_cast<size_t>(m.buf_rows[0] + m.buf_rows[1]) * f32k::LDA * sizeof(float);
    const size_t stage = static_cast<size_t>(f32k::KB) * m.max_npad * sizeof(float);
    const size_t limit = 227 * 1024;
    if (act + 3 * stage <= limit)
        h->f32_wst = 3;
    else if (act + 2 * stage <= limit)
        h->f32_wst = 2;
    else
        return fail(VAE21_ERR_UNSUPPORTED, "layer stack needs %zu B of activation shared memory (+%zu B/stage): too wide",
                    act, stage);
    h->f32_smem = act + h->f32_wst * stage;
    h->f32p_rows = std::max(m.buf_rows[0], m.buf_rows[1]);
    h->f32p_stages = 0;
    for (int l = 0; l < h->n_layers; ++l) h->f32p_stages += m.L[l].Kpad / f32p::KB;
    h->f32p_smem = f32p::HEADER_BYTES + static_cast<size_t>(h->f32p_rows) * f32k::LDA * sizeof(float) +
                   static_cast<size_t>(f32p::WST) * f32p::STAGE_FLOATS * sizeof(float);  // <= 224 KB for any width <= 480
    if (h->d_w32) cudaFree(h->d_w32);
    if (h->d_b32) cudaFree(h->d_b32);
    h->d_w32 = h->d_b32 = nullptr;
    CK(cudaMalloc(&h->d_w32, W.size() * sizeof(float)));
    CK(cudaMalloc(&h->d_b32, B.size() * sizeof(float)));
    CK(cudaMemcpy(h->d_w32, W.data(), W.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_b32, B.data(), B.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(f32k::vae21_fp32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
    CK(cudaFuncSetAttribute(f32k::vae21_fp32_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
    CK(cudaFuncSetAttribute(f32p::vae21_fp32_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
    return 0;
}

int launch_fp32(vae21_handle* h, const LaunchArgs& a, cudaStream_t st) {
    const long long ntiles = (a.n + f32k::MT - 1) / f32k::MT;
    if (ntiles == 0) return 0;
    const int grid = (int)std::min<long long>(ntiles, h->sm_count);
    // VAE21_FP32_PIPE=0 selects the block-barrier kernel (same bits; kept for A/B measurements and for stacks the ring does not fit)
    static const int pipe_env = std::getenv("VAE21_FP32_PIPE") ? std::atoi(std::getenv("VAE21_FP32_PIPE")) : VAE21_FP32_PIPE_DEFAULT;
    if (pipe_env && h->f32p_smem <= 227 * 1024 && h->f32p_stages <= f32p::MAX_STAGES)
        f32p::vae21_fp32_pipe_kernel<<<grid, f32p::NTHREADS, h->f32p_smem, st>>>(h->f32, h->nc, a, h->d_w32, h->d_b32, h->f32p_rows,
                                                                                  h->f32p_stages);
    else if (h->f32_wst == 3)
        f32k::vae21_fp32_kernel<3><<<grid, f32k::NTHREADS, h->f32_smem, st>>>(h->f32, h->nc, a, h->d_w32, h->d_b32);
    else
        f32k::vae21_fp32_kernel<2><<<grid, f32k::NTHREADS, h->f32_smem, st>>>(h->f32, h->nc, a, h->d_w32, h->d_b32);
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}

int launch(vae21_handle* h, const LaunchArgs& a, int precision, cudaStream_t st) {
    if (precision == VAE21_FP32_SIMT) return launch_fp32(h, a, st);
    if (precision == VAE21_TC_BF16X3 || precision == VAE21_TC_FP16X3 || precision == VAE21_TC_FP16E4M3) {
        if (!h->tc_ok)
            return fail(VAE21_ERR_UNSUPPORTED, "tensor-core path unavailable for this layer stack: %s", h->tc_why.c_str());
        const int fmt = precision == VAE21_TC_FP16X3 ? 1 : precision == VAE21_TC_FP16E4M3 ? 2 : 0;
        LaunchArgs at = a;
        at.sat = h->d_sat;
        cudaError_t e = tck::launch(h->tc, h->nc, at, h->d_wtc[fmt], h->d_btc, fmt, h->sm_count, st);
        if (e != cudaSuccess) return fail(VAE21_ERR_CUDA, "tensor-core kernel launch failed: %s", cudaGetErrorString(e));
        h->launches++;
        return 0;
    }
    return fail(VAE21_ERR_ARG, "unknown precision %d", precision);
}

// Upload obs / inv_sigma unless the device already holds exactly these values.
int upload_observation(vae21_handle* h, const float* obs, const float* isig, int NO, cudaStream_t st) {
    if (h->obs_cached && static_cast<int>(h->h_obs.size()) == NO && std::memcmp(h->h_obs.data(), obs, sizeof(float) * NO) == 0 &&
        std::memcmp(h->h_isig.data(), isig, sizeof(float) * NO) == 0)
        return 0;
    h->obs_cached = false;
    if (int rc = wait_last_use(h)) return rc;
    h->h_obs.assign(obs, obs + NO);
    h->h_isig.assign(isig, isig + NO);
    CK(cudaMemcpyAsync(h->d_obs, h->h_obs.data(), sizeof(float) * NO, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_isig, h->h_isig.data(), sizeof(float) * NO, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // the values are in place before any stream of the handle may use them
    h->obs_cached = true;
    return 0;
}

// Generic driver for predict / forward_normalised / chi2.
int run(vae21_handle* h, const void* in, int in_mode, bool in_dev, long long n, float* out, int out_mode, bool out_dev,
        const float* obs_host, const float* isig_host, float* best_val, int64_t* best_idx, int precision, void* stream) {
    if (!h) return fail(VAE21_ERR_ARG, "null handle");
    if (!h->model_set) return fail(VAE21_ERR_STATE, "vae21_set_model has not been called");
    if (in_mode != IN_NORMALISED_F32 || out_mode != OUT_NORMALISED)
        if (!h->norm_set) return fail(VAE21_ERR_STATE, "vae21_set_norm has not been called");
    if (n < 0) return fail(VAE21_ERR_ARG, "negative row count");
    if (n >= (1ll << 32)) return fail(VAE21_ERR_ARG, "row count must be below 2^32 per call");
    if (n > 0 && !in) return fail(VAE21_ERR_ARG, "null input pointer");
    if (out_mode != OUT_CHI2 && n > 0 && !out) return fail(VAE21_ERR_ARG, "null output pointer");
    if (int rc = use_device(h)) return rc;

    const int K0 = h->dims[0], NO = h->dims[h->n_layers];
    const size_t in_elt = (in_mode == IN_PARAMS_F64) ? 8 : 4;
    const bool chi = (out_mode == OUT_CHI2);
    const bool want_best = chi && (best_val || best_idx);
    const bool all_dev = in_dev && (out_dev || (chi && !out));
    cudaStream_t ust = reinterpret_cast<cudaStream_t>(stream);
    cudaStream_t st0 = all_dev ? ust : h->streams[0];

    if (chi) {
        if (!obs_host || !isig_host) return fail(VAE21_ERR_ARG, "obs / inv_sigma must not be null");
        if (int rc = upload_observation(h, obs_host, isig_host, NO, st0)) return rc;
        if (want_best) CK(cudaMemsetAsync(h->d_key, 0xff, sizeof(unsigned long long), st0));
        if (!all_dev) CK(cudaStreamSynchronize(st0));  // other pipeline streams read these too
    }

    LaunchArgs a{};
    a.mu = h->d_mu;
    a.obs = h->d_obs;
    a.isig = h->d_isig;
    a.argmin_key = want_best ? h->d_key : nullptr;
    a.in_mode = in_mode;
    a.out_mode = out_mode;

    if (n > 0) {
        if (all_dev) {
            a.in = in;
            a.out = chi ? nullptr : out;
            a.chi2 = chi ? out : nullptr;
            a.n = n;
            a.row_base = 0;
            if (int rc = launch(h, a, precision, ust)) return rc;
            CK(cudaEventRecord(h->ev_use, ust));
            h->use_pending = true;
        } else {
            // A device-resident input (or output) belongs to the caller's stream: the pipeline streams must not touch it before
            // the work already queued there has finished.  (The call returns only after the pipeline streams have drained, so
            // nothing the caller queues afterwards can overtake it.)
            if (in_dev || out_dev) {
                CK(cudaEventRecord(h->ev_in, ust));
                for (int s = 0; s < NSLOT; ++s) CK(cudaStreamWaitEvent(h->streams[s], h->ev_in, 0));
            }
            const size_t out_row = chi ? sizeof(float) : sizeof(float) * NO;
            const long long chunk = chi ? CHUNK_ROWS * 16 : CHUNK_ROWS;
            long long done = 0;
            int c = 0;
            while (done < n) {
                const long long rows = std::min(chunk, n - done);
                const int s = c % NSLOT;
                cudaStream_t st = h->streams[s];
                const size_t ib = rows * K0 * in_elt, ob = rows * out_row;
                if (int rc = ensure_slot(h, s, in_dev ? 0 : ib, (out_dev || (chi && !out)) ? 0 : ob)) return rc;
                const char* src = reinterpret_cast<const char*>(in) + done * K0 * in_elt;
                if (in_dev) {
                    a.in = src;
                } else {
                    CK(cudaMemcpyAsync(h->d_in[s], src, ib, cudaMemcpyHostToDevice, st));
                    a.in = h->d_in[s];
                }
                float* dst_final = out ? reinterpret_cast<float*>(reinterpret_cast<char*>(out) + done * out_row) : nullptr;
                float* kout = out_dev ? dst_final : (out ? h->d_out[s] : nullptr);
                a.out = chi ? nullptr : kout;
                a.chi2 = chi ? kout : nullptr;
                a.n = rows;
                a.row_base = done;
                if (int rc = launch(h, a, precision, st)) return rc;
                if (!out_dev && out) CK(cudaMemcpyAsync(dst_final, h->d_out[s], ob, cudaMemcpyDeviceToHost, st));
                done += rows;
                ++c;
            }
            for (int s = 0; s < NSLOT; ++s) CK(cudaStreamSynchronize(h->streams[s]));
        }
    }
    if (want_best) {
        unsigned long long key = ~0ull;
        CK(cudaMemcpyAsync(&key, h->d_key, sizeof key, cudaMemcpyDeviceToHost, st0));
        CK(cudaStreamSynchronize(st0));
        const uint32_t bits = static_cast<uint32_t>(key >> 32);
        float v;
        memcpy(&v, &bits, 4);
        const bool none = (key == ~0ull) || std::isnan(v);
        if (best_val) *best_val = none ? NAN : v;
        if (best_idx) *best_idx = none ? -1 : static_cast<int64_t>(key & 0xffffffffull);
    }
    return 0;
}

}  // namespace

// ---- exported C ABI -------------------------------------------------------------------------
extern "C" {

int vae21_version(void) { return VAE21_VERSION; }

const char* vae21_last_error(void) { return g_err.c_str(); }

int vae21_device_count(int* count) {
    if (!count) return fail(VAE21_ERR_ARG, "null count");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
        return fail(VAE21_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = c;
    return 0;
}

int vae21_create(int device, vae21_handle** out) {
    if (!out) return fail(VAE21_ERR_ARG, "null out pointer");
    *out = nullptr;
    int cnt = 0;
    if (int rc = vae21_device_count(&cnt)) return rc;
    if (cnt == 0) return fail(VAE21_ERR_CUDA, "no CUDA device present (this library has no CPU fallback)");
    if (device < 0 || device >= cnt) return fail(VAE21_ERR_ARG, "device %d out of range [0,%d)", device, cnt);
    CK(cudaSetDevice(device));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) return fail(VAE21_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
    vae21_handle* h = new (std::nothrow) vae21_handle();
    if (!h) return fail(VAE21_ERR_NOMEM, "out of host memory");
    h->device = device;
    h->sm_count = p.multiProcessorCount;
    for (int s = 0; s < NSLOT; ++s) {
        cudaError_t e = cudaStreamCreateWithFlags(&h->streams[s], cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            vae21_destroy(h);
            return fail(VAE21_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
    }
    if (cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_use, cudaEventDisableTiming) != cudaSuccess ||
        cudaMalloc(&h->d_key, sizeof(unsigned long long)) != cudaSuccess || cudaMalloc(&h->d_sat, 2 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(h->d_sat, 0, 2 * sizeof(unsigned long long)) != cudaSuccess) {
        vae21_destroy(h);
        return fail(VAE21_ERR_CUDA, "handle resource creation failed");
    }
    *out = h;
    return 0;
}

int vae21_destroy(vae21_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    for (int s = 0; s < NSLOT; ++s) {
        if (h->streams[s]) {
            cudaStreamSynchronize(h->streams[s]);
            cudaStreamDestroy(h->streams[s]);
        }
        if (h->d_in[s]) cudaFree(h->d_in[s]);
        if (h->d_out[s]) cudaFree(h->d_out[s]);
    }
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    if (h->ev_use) {
        cudaEventSynchronize(h->ev_use);
        cudaEventDestroy(h->ev_use);
    }
    void* ptrs[] = {h->d_w32, h->d_b32, h->d_wtc[0], h->d_wtc[1], h->d_wtc[2], h->d_btc, h->d_mu, h->d_obs, h->d_isig, h->d_mask, h->d_key, h->d_sat, h->d_mcmc};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete h;
    return 0;
}

int vae21_set_model(vae21_handle* h, int n_layers, const int* dims, const float* const* kernels,
                    const float* const* biases, const int* relu_flags) {
    if (!h || !dims || !kernels || !biases || !relu_flags) return fail(VAE21_ERR_ARG, "null argument");
    if (n_layers < 1 || n_layers > VAE21_MAX_LAYERS) return fail(VAE21_ERR_ARG, "n_layers must be in [1,%d]", VAE21_MAX_LAYERS);
    for (int l = 0; l <= n_layers; ++l)
        if (dims[l] < 1) return fail(VAE21_ERR_ARG, "dims[%d] = %d", l, dims[l]);
    if (dims[0] > VAE21_MAX_PAR) return fail(VAE21_ERR_UNSUPPORTED, "at most %d input parameters", VAE21_MAX_PAR);
    for (int l = 0; l < n_layers; ++l)
        if (!kernels[l] || !biases[l]) return fail(VAE21_ERR_ARG, "null kernel/bias for layer %d", l);
    if (int rc = use_device(h)) return rc;
    for (int s = 0; s < NSLOT; ++s) CK(cudaStreamSynchronize(h->streams[s]));
    if (int rc = wait_last_use(h)) return rc;  // weight images are about to be replaced
    h->model_set = false;
    const int old_in = h->n_layers ? h->dims[0] : -1;
    h->n_layers = n_layers;
    for (int l = 0; l <= n_layers; ++l) h->dims[l] = dims[l];
    // a failure below leaves model_set == false (every compute call refuses) and n_layers == 0, so a retry starts from scratch
    struct Rollback {
        vae21_handle* h;
        bool ok = false;
        ~Rollback() {
            if (!ok) {
                h->n_layers = 0;
                h->tc_ok = false;
                h->norm_set = false;
            }
        }
    } guard{h};
    if (int rc = pack_fp32(h, kernels, biases, relu_flags)) return rc;
    // tensor-core plan (optional: a stack that does not fit leaves the fp32 path usable)
    h->tc_ok = false;
    {
        std::string why;
        std::vector<unsigned short> img[3];
        std::vector<float> bias_img;
        if (tck::build_plan(n_layers, dims, kernels, biases, relu_flags, h->tc, img, bias_img, why)) {
            for (int f = 0; f < 3; ++f) {
                if (h->d_wtc[f]) cudaFree(h->d_wtc[f]);
                h->d_wtc[f] = nullptr;
                CK(cudaMalloc(&h->d_wtc[f], img[f].size() * sizeof(unsigned short)));
                CK(cudaMemcpy(h->d_wtc[f], img[f].data(), img[f].size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
            }
            if (h->d_btc) cudaFree(h->d_btc);
            h->d_btc = nullptr;
            CK(cudaMalloc(&h->d_btc, bias_img.size() * sizeof(float)));
            CK(cudaMemcpy(h->d_btc, bias_img.data(), bias_img.size() * sizeof(float), cudaMemcpyHostToDevice));
            cudaError_t e = tck::prepare();
            if (e != cudaSuccess) return fail(VAE21_ERR_CUDA, "tensor-core kernel attribute setup: %s", cudaGetErrorString(e));
            h->tc_ok = true;
        } else {
            h->tc_why = why;
        }
    }
    const int NO = dims[n_layers];
    if (NO != h->alloc_out) {  // the width the constant buffers are ALLOCATED for, not what an earlier (possibly failed) call recorded
        h->obs_cached = false;
        h->norm_set = false;
        h->alloc_out = 0;
        for (float** p : {&h->d_mu, &h->d_obs, &h->d_isig, &h->d_mask}) {
            if (*p) cudaFree(*p);
            *p = nullptr;
            CK(cudaMalloc(p, sizeof(float) * NO));
            CK(cudaMemset(*p, 0, sizeof(float) * NO));
        }
        h->alloc_out = NO;
    }
    if (dims[0] != old_in) h->norm_set = false;  // the prologue constants were for another input width
    guard.ok = true;
    h->model_set = true;
    return 0;
}

int vae21_set_norm(vae21_handle* h, int n_par, const double* par_min, const double* par_max, const int* log_mask,
                   int floor_col, double fx_floor, int n_out, const float* sig_mean, float sig_std) {
    if (!h || !par_min || !par_max || !log_mask || !sig_mean) return fail(VAE21_ERR_ARG, "null argument");
    if (!h->model_set) return fail(VAE21_ERR_STATE, "call vae21_set_model first");
    if (n_par != h->dims[0]) return fail(VAE21_ERR_ARG, "n_par %d != model input width %d", n_par, h->dims[0]);
    if (n_out != h->dims[h->n_layers]) return fail(VAE21_ERR_ARG, "n_out %d != model output width %d", n_out, h->dims[h->n_layers]);
    if (int rc = use_device(h)) return rc;
    NormConsts& nc = h->nc;
    nc = NormConsts{};
    nc.n_par = n_par;
    for (int j = 0; j < n_par; ++j) {
        nc.pmin[j] = par_min[j];
        nc.prange[j] = par_max[j] - par_min[j];  // same fp64 subtraction numpy performs (preprocess.py:106)
        nc.log_mask[j] = log_mask[j] ? 1 : 0;
        nc.pscale[j] = 2.0 / nc.prange[j];
    }
    nc.floor_col = floor_col;
    nc.floor_val = fx_floor;
    nc.sd = sig_std;
    for (int s = 0; s < NSLOT; ++s) CK(cudaStreamSynchronize(h->streams[s]));
    if (int rc = wait_last_use(h)) return rc;
    CK(cudaMemcpy(h->d_mu, sig_mean, sizeof(float) * n_out, cudaMemcpyHostToDevice));
    h->norm_set = true;
    return 0;
}

int vae21_predict(vae21_handle* h, const void* params, int params_dtype, int params_on_device, int64_t n, float* out,
                  int out_on_device, int precision, void* stream) {
    if (params_dtype != VAE21_F32 && params_dtype != VAE21_F64) return fail(VAE21_ERR_ARG, "bad params_dtype %d", params_dtype);
    return run(h, params, params_dtype == VAE21_F64 ? IN_PARAMS_F64 : IN_PARAMS_F32, params_on_device != 0, n, out,
               OUT_PREDICT, out_on_device != 0, nullptr, nullptr, nullptr, nullptr, precision, stream);
}

int vae21_forward_normalised(vae21_handle* h, const float* x, int x_on_device, int64_t n, float* y, int y_on_device,
                             int precision, void* stream) {
    return run(h, x, IN_NORMALISED_F32, x_on_device != 0, n, y, OUT_NORMALISED, y_on_device != 0, nullptr, nullptr,
               nullptr, nullptr, precision, stream);
}

int vae21_chi2(vae21_handle* h, const void* params, int params_dtype, int params_on_device, int64_t n,
               const float* obs, const float* inv_sigma, float* chi2, int chi2_on_device, float* best_val,
               int64_t* best_idx, int precision, void* stream) {
    if (params_dtype != VAE21_F32 && params_dtype != VAE21_F64) return fail(VAE21_ERR_ARG, "bad params_dtype %d", params_dtype);
    return run(h, params, params_dtype == VAE21_F64 ? IN_PARAMS_F64 : IN_PARAMS_F32, params_on_device != 0, n, chi2,
               OUT_CHI2, chi2_on_device != 0, obs, inv_sigma, best_val, best_idx, precision, stream);
}

void* vae21_host_alloc(size_t bytes) {
    void* p = g_pool.alloc(bytes);
    if (!p) fail(VAE21_ERR_NOMEM, "pinned allocation of %zu bytes failed", bytes);
    return p;
}
void vae21_host_free(void* p) { g_pool.release(p); }
void vae21_host_trim(void) { g_pool.trim(); }

int vae21_chi2_grid(vae21_handle* h, int n_dim, const int* npts, const double* x_lo, const double* x_hi, int64_t first, int64_t count,
                    const float* obs, const float* inv_sigma, float* chi2_dev, float* best_val, int64_t* best_idx, int precision,
                    void* stream) {
    if (!h || !npts || !x_lo || !x_hi) return fail(VAE21_ERR_ARG, "null argument");
    if (!h->model_set || !h->norm_set) return fail(VAE21_ERR_STATE, "vae21_set_model / vae21_set_norm have not been called");
    if (n_dim != h->dims[0]) return fail(VAE21_ERR_ARG, "grid has %d dimensions, the model %d inputs", n_dim, h->dims[0]);
    if (!obs || !inv_sigma) return fail(VAE21_ERR_ARG, "obs / inv_sigma must not be null");
    double total = 1.0;
    for (int j = 0; j < n_dim; ++j) {
        if (npts[j] < 1) return fail(VAE21_ERR_ARG, "npts[%d] = %d", j, npts[j]);
        total *= npts[j];
    }
    if (first < 0 || count < 0 || static_cast<double>(first) + static_cast<double>(count) > total)
        return fail(VAE21_ERR_ARG, "point range [%lld, %lld) outside the grid", (long long)first, (long long)(first + count));
    if (first + count > (1ll << 32)) return fail(VAE21_ERR_ARG, "grid indices of one call must stay below 2^32 (shard the grid)");
    if (int rc = use_device(h)) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int NO = h->dims[h->n_layers];
    const bool want_best = best_val || best_idx;
    if (int rc = upload_observation(h, obs, inv_sigma, NO, st)) return rc;
    if (want_best) CK(cudaMemsetAsync(h->d_key, 0xff, sizeof(unsigned long long), st));
    LaunchArgs a{};
    a.in = h->d_obs;  // unused in IN_GRID mode (must be non-null)
    a.mu = h->d_mu;
    a.obs = h->d_obs;
    a.isig = h->d_isig;
    a.chi2 = chi2_dev;
    a.argmin_key = want_best ? h->d_key : nullptr;
    a.in_mode = IN_GRID;
    a.out_mode = OUT_CHI2;
    a.n = count;
    a.row_base = first;
    for (int j = 0; j < n_dim; ++j) {
        a.grid_n[j] = npts[j];
        a.grid_lo[j] = static_cast<float>(x_lo[j]);
        a.grid_step[j] = npts[j] > 1 ? static_cast<float>((x_hi[j] - x_lo[j]) / (npts[j] - 1)) : 0.f;
    }
    if (count > 0) {
        if (int rc = launch(h, a, precision, st)) return rc;
        CK(cudaEventRecord(h->ev_use, st));
        h->use_pending = true;
    }
    if (want_best) {
        unsigned long long key = ~0ull;
        CK(cudaMemcpyAsync(&key, h->d_key, sizeof key, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const uint32_t bits = static_cast<uint32_t>(key >> 32);
        float v;
        memcpy(&v, &bits, 4);
        const bool none = (key == ~0ull) || std::isnan(v);
        if (best_val) *best_val = none ? NAN : v;
        if (best_idx) *best_idx = none ? -1 : static_cast<int64_t>(key & 0xffffffffull);
    }
    return 0;
}

int vae21_error(vae21_handle* h, const void* params, int params_dtype, int params_on_device, int64_t n, const float* truth,
                int truth_on_device, const float* band_mask, int relative, float* err, int err_on_device, int precision, void* stream) {
    if (!h) return fail(VAE21_ERR_ARG, "null handle");
    if (!h->model_set || !h->norm_set) return fail(VAE21_ERR_STATE, "vae21_set_model / vae21_set_norm have not been called");
    if (n < 0 || n >= (1ll << 32)) return fail(VAE21_ERR_ARG, "row count must be in [0, 2^32)");
    if (n == 0) return 0;
    if (!params || !truth || !err) return fail(VAE21_ERR_ARG, "null pointer");
    if (params_dtype != VAE21_F32 && params_dtype != VAE21_F64) return fail(VAE21_ERR_ARG, "params_dtype %d", params_dtype);
    if (int rc = use_device(h)) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int K0 = h->dims[0], NO = h->dims[h->n_layers];
    const size_t in_elt = params_dtype == VAE21_F64 ? 8 : 4;
    // band mask (all ones when NULL) travels in the inv_sigma slot
    std::vector<float> mask(NO, 1.f);
    int in_band = NO;
    if (band_mask) {
        in_band = 0;
        for (int k = 0; k < NO; ++k) {
            mask[k] = band_mask[k] != 0.f ? 1.f : 0.f;
            in_band += mask[k] != 0.f;
        }
    }
    if (in_band == 0) return fail(VAE21_ERR_ARG, "the frequency band contains no bin");
    if (int rc = wait_last_use(h)) return rc;  // an earlier asynchronous error launch may still read the mask
    CK(cudaMemcpyAsync(h->d_mask, mask.data(), sizeof(float) * NO, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // `mask` goes out of scope
    void *d_par = nullptr, *d_truth = nullptr, *d_err = nullptr;
    auto cleanup = [&]() {
        for (void* p : {d_par, d_truth, d_err})
            if (p) cudaFree(p);
    };
    LaunchArgs a{};
    a.in = params;
    a.truth = truth;
    a.chi2 = err;
    if (!params_on_device) {
        if (cudaMalloc(&d_par, n * K0 * in_elt) != cudaSuccess) { cleanup(); return fail(VAE21_ERR_NOMEM, "device allocation failed"); }
        if (cudaMemcpyAsync(d_par, params, n * K0 * in_elt, cudaMemcpyHostToDevice, st) != cudaSuccess) { cleanup(); return fail(VAE21_ERR_CUDA, "parameter upload failed: %s", cudaGetErrorString(cudaGetLastError())); }
        a.in = d_par;
    }
    if (!truth_on_device) {
        if (cudaMalloc(&d_truth, sizeof(float) * n * NO) != cudaSuccess) { cleanup(); return fail(VAE21_ERR_NOMEM, "device allocation failed"); }
        if (cudaMemcpyAsync(d_truth, truth, sizeof(float) * n * NO, cudaMemcpyHostToDevice, st) != cudaSuccess) { cudaStreamSynchronize(st); cleanup(); return fail(VAE21_ERR_CUDA, "upload of the true signals failed: %s", cudaGetErrorString(cudaGetLastError())); }
        a.truth = static_cast<const float*>(d_truth);
    }
    if (!err_on_device) {
        if (cudaMalloc(&d_err, sizeof(float) * n) != cudaSuccess) { cleanup(); return fail(VAE21_ERR_NOMEM, "device allocation failed"); }
        a.chi2 = static_cast<float*>(d_err);
    }
    a.mu = h->d_mu;
    a.obs = h->d_obs;
    a.isig = h->d_mask;  // the band mask travels in the inv_sigma slot
    a.in_mode = params_dtype == VAE21_F64 ? IN_PARAMS_F64 : IN_PARAMS_F32;
    a.out_mode = OUT_ERROR;
    a.n = n;
    a.err_inv_count = 1.f / static_cast<float>(in_band);
    a.err_relative = relative ? 1 : 0;
    int rc = launch(h, a, precision, st);
    if (rc == 0 && !err_on_device) {
        if (cudaMemcpyAsync(err, d_err, sizeof(float) * n, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = fail(VAE21_ERR_CUDA, "copy back failed");
    }
    if (cudaStreamSynchronize(st) != cudaSuccess && rc == 0) rc = fail(VAE21_ERR_CUDA, "error kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    cleanup();
    return rc;
}

int vae21_get_info(vae21_handle* h, int64_t* kernel_launches, float* last_kernel_ms, int* tc_supported) {
    if (!h) return fail(VAE21_ERR_ARG, "null handle");
    if (kernel_launches) *kernel_launches = h->launches;
    if (last_kernel_ms) *last_kernel_ms = h->last_ms;
    if (tc_supported) *tc_supported = h->tc_ok ? 1 : 0;
    return 0;
}

int vae21_get_tc_stats(vae21_handle* h, int64_t* saturated, int reset) {
    if (!h) return fail(VAE21_ERR_ARG, "null handle");
    if (int rc = use_device(h)) return rc;
    unsigned long long v[2] = {0, 0};
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(v, h->d_sat, sizeof v, cudaMemcpyDeviceToHost));
    if (saturated) *saturated = static_cast<int64_t>(v[0]);
    if (reset) CK(cudaMemset(h->d_sat, 0, sizeof v));
    return 0;
}

int vae21_time_predict(vae21_handle* h, const void* params_dev, int params_dtype, int64_t n, float* out_dev,
                       int precision, int iters, float* ms_per_launch) {
    if (!h || !ms_per_launch) return fail(VAE21_ERR_ARG, "null argument");
    if (iters < 1) return fail(VAE21_ERR_ARG, "iters must be >= 1");
    if (int rc = use_device(h)) return rc;
    cudaStream_t st = h->streams[0];
    // one untimed launch so lazy module loading is outside the timed region
    if (int rc = vae21_predict(h, params_dev, params_dtype, 1, n, out_dev, 1, precision, st)) return rc;
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(h->ev0, st));
    for (int i = 0; i < iters; ++i)
        if (int rc = vae21_predict(h, params_dev, params_dtype, 1, n, out_dev, 1, precision, st)) return rc;
    CK(cudaEventRecord(h->ev1, st));
    CK(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms / iters;
    *ms_per_launch = h->last_ms;
    return 0;
}

int vae21_check_plan(int n_layers, const int* dims, char* msg, int msg_len) {
    // Host only: plan the tensor-core schedule of a Dense stack (zero weights: only shapes matter) and replay it (tck::check_schedule).
    // 0 = the stack has a valid schedule, 1 = it does not fit the tensor-core kernel (msg says why), 2 = the schedule is inconsistent.
    if (!dims || n_layers < 1 || n_layers > VAE21_MAX_LAYERS) return fail(VAE21_ERR_ARG, "bad layer stack");
    std::vector<std::vector<float>> ks(n_layers), bs(n_layers);
    std::vector<const float*> kp(n_layers), bp(n_layers);
    std::vector<int> relu(n_layers, 1);
    relu[n_layers - 1] = 0;
    for (int l = 0; l < n_layers; ++l) {
        if (dims[l] < 1 || dims[l + 1] < 1) return fail(VAE21_ERR_ARG, "bad layer width");
        ks[l].assign(static_cast<size_t>(dims[l]) * dims[l + 1], 0.f);
        bs[l].assign(dims[l + 1], 0.f);
        kp[l] = ks[l].data();
        bp[l] = bs[l].data();
    }
    tck::Plan P;
    std::vector<unsigned short> img[3];
    std::vector<float> bias_img;
    std::string why;
    int rc = 0;
    if (!tck::build_plan(n_layers, dims, kp.data(), bp.data(), relu.data(), P, img, bias_img, why)) rc = 1;
    else if (!tck::check_schedule(P, bias_img, why)) rc = 2;
    if (msg && msg_len > 0) {
        std::snprintf(msg, static_cast<size_t>(msg_len), "%s", why.c_str());
    }
    return rc;
}

#if VAE21_TC_TIMING
// profiling builds only (not declared in vae21.h)
int vae21_debug_tc_timing(long long* out) { return tck::read_timing(out) == cudaSuccess ? 0 : 3; }
int vae21_debug_tc_rec_timing(unsigned long long* out, int reset) { return tck::read_rec_timing(out, reset) == cudaSuccess ? 0 : 3; }

#endif

}  // extern "C"

#include "mcmc_api.cuh"
#include "train_api.cuh"
