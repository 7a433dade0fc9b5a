// C-ABI of the trainer (declared in include/vae21.h).  Included by vae21_api.cu.
#pragma once
#include <vector>

#include "train_kernels.cuh"

struct vae21_trainer {
    int device = 0;
    int n_layers = 0;
    int dims[VAE21_MAX_LAYERS + 1] = {0};
    int relu[VAE21_MAX_LAYERS] = {0};
    int max_batch = 0;
    long long n_params = 0;
    long long w_off[VAE21_MAX_LAYERS] = {0}, b_off[VAE21_MAX_LAYERS] = {0};  // float offsets into the flat parameter vector
    float *p = nullptr, *m = nullptr, *v = nullptr;                           // parameters and Adam moments [n_params]
    float* act[VAE21_MAX_LAYERS + 1] = {nullptr};                             // act[0] = batch inputs, act[l+1] = output of layer l
    float* delta[2] = {nullptr, nullptr};                                     // ping-pong dL/d(pre-activation... post-mask) buffers
    float *yb = nullptr, *wb = nullptr, *loss_rows = nullptr;
    float* grad = nullptr;  // [n_params], used by vae21_trainer_epoch (single-GPU fast path)
    cudaEvent_t throttle[2] = {nullptr, nullptr};
    // one optimisation step as a replayable CUDA graph (vae21_trainer_epoch): the batch number and the per-step learning rate
    // are read from device memory, so the same executable graph serves every full batch of every epoch
    cudaStream_t gstream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    cudaGraphExec_t gexec = nullptr;
    const void *g_x = nullptr, *g_y = nullptr, *g_w = nullptr;
    int g_batch = 0;
    long long g_kernels = 0;  // kernels in the captured step
    float g_b1 = 0, g_b2 = 0, g_eps = 0;
    int* d_perm = nullptr;
    long long perm_cap = 0;
    float* d_lr = nullptr;
    long long lr_cap = 0;
    int* d_step = nullptr;
    float* d_loss = nullptr;
    // data-parallel step as TWO replayable graphs around the caller's gradient all-reduce (vae21_trainer_dp_*): A = gather of this
    // rank's share of the step's batch + forward + loss + backward into the caller's gradient buffer, B = Adam + step counter
    cudaGraphExec_t dpA = nullptr, dpB = nullptr;
    const void *dp_x = nullptr, *dp_y = nullptr, *dp_w = nullptr, *dp_grad = nullptr, *dp_loss = nullptr;
    int dp_batch = 0, dp_first = 0, dp_rows = 0;
    float dp_b1 = 0, dp_b2 = 0, dp_eps = 0;
    long long dpA_kernels = 0, dp_steps = 0;
    long long launches = 0;
};

namespace {
int trainer_use(vae21_trainer* t) {
    if (!t) return fail(VAE21_ERR_ARG, "null trainer");
    CK(cudaSetDevice(t->device));
    return 0;
}
void trainer_forward(vae21_trainer* t, int batch, cudaStream_t st) {
    for (int l = 0; l < t->n_layers; ++l) {
        const int K = t->dims[l], N = t->dims[l + 1];
        trk::sgemm<0>(batch, N, K, t->act[l], K, t->p + t->w_off[l], N, t->act[l + 1], N, t->p + t->b_off[l], t->relu[l], nullptr, 0, st);
        t->launches++;
    }
}
}  // namespace

extern "C" {

int vae21_trainer_create(int device, int n_layers, const int* dims, const int* relu_flags, int max_batch, vae21_trainer** out) {
    if (!dims || !relu_flags || !out) return fail(VAE21_ERR_ARG, "null argument");
    if (n_layers < 1 || n_layers > VAE21_MAX_LAYERS || max_batch < 1) return fail(VAE21_ERR_ARG, "bad n_layers / max_batch");
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(VAE21_ERR_ARG, "device %d out of range (%d visible)", device, count);
    CK(cudaSetDevice(device));
    vae21_trainer* t = new vae21_trainer();
    t->device = device;
    t->n_layers = n_layers;
    t->max_batch = max_batch;
    long long off = 0;
    int widest = 0;
    for (int l = 0; l <= n_layers; ++l) {
        if (dims[l] < 1) { delete t; return fail(VAE21_ERR_ARG, "dims[%d] = %d", l, dims[l]); }
        t->dims[l] = dims[l];
        widest = dims[l] > widest ? dims[l] : widest;
    }
    for (int l = 0; l < n_layers; ++l) {
        t->relu[l] = relu_flags[l] ? 1 : 0;
        t->w_off[l] = off;
        off += static_cast<long long>(dims[l]) * dims[l + 1];
        t->b_off[l] = off;
        off += dims[l + 1];
    }
    t->n_params = off;
    for (float** q : {&t->p, &t->m, &t->v}) {
        CK(cudaMalloc(q, sizeof(float) * off));
        CK(cudaMemset(*q, 0, sizeof(float) * off));
    }
    for (int l = 0; l <= n_layers; ++l) CK(cudaMalloc(&t->act[l], sizeof(float) * static_cast<size_t>(max_batch) * dims[l]));
    for (int i = 0; i < 2; ++i) CK(cudaMalloc(&t->delta[i], sizeof(float) * static_cast<size_t>(max_batch) * widest));
    CK(cudaMalloc(&t->yb, sizeof(float) * static_cast<size_t>(max_batch) * dims[n_layers]));
    CK(cudaMalloc(&t->wb, sizeof(float) * max_batch));
    CK(cudaMalloc(&t->loss_rows, sizeof(float) * max_batch));
    CK(cudaMalloc(&t->grad, sizeof(float) * off));
    for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&t->throttle[i], cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&t->gstream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&t->ev_in, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&t->ev_out, cudaEventDisableTiming));
    CK(cudaMalloc(&t->d_step, sizeof(int)));
    CK(cudaMalloc(&t->d_loss, sizeof(float)));
    *out = t;
    return 0;
}

int vae21_trainer_destroy(vae21_trainer* t) {
    if (!t) return 0;
    cudaSetDevice(t->device);
    cudaDeviceSynchronize();
    for (float* q : {t->p, t->m, t->v, t->delta[0], t->delta[1], t->yb, t->wb, t->loss_rows, t->grad})
        if (q) cudaFree(q);
    for (int l = 0; l <= t->n_layers; ++l)
        if (t->act[l]) cudaFree(t->act[l]);
    for (int i = 0; i < 2; ++i)
        if (t->throttle[i]) cudaEventDestroy(t->throttle[i]);
    if (t->gexec) cudaGraphExecDestroy(t->gexec);
    if (t->dpA) cudaGraphExecDestroy(t->dpA);
    if (t->dpB) cudaGraphExecDestroy(t->dpB);
    if (t->gstream) cudaStreamDestroy(t->gstream);
    if (t->ev_in) cudaEventDestroy(t->ev_in);
    if (t->ev_out) cudaEventDestroy(t->ev_out);
    for (void* q : {static_cast<void*>(t->d_perm), static_cast<void*>(t->d_lr), static_cast<void*>(t->d_step), static_cast<void*>(t->d_loss)})
        if (q) cudaFree(q);
    delete t;
    return 0;
}

int vae21_trainer_num_params(vae21_trainer* t, int64_t* n) {
    if (!t || !n) return fail(VAE21_ERR_ARG, "null argument");
    *n = t->n_params;
    return 0;
}

int vae21_trainer_set_params(vae21_trainer* t, const float* flat_host, int reset_moments) {
    if (int rc = trainer_use(t)) return rc;
    if (!flat_host) return fail(VAE21_ERR_ARG, "null parameters");
    CK(cudaMemcpy(t->p, flat_host, sizeof(float) * t->n_params, cudaMemcpyHostToDevice));
    if (reset_moments) {
        CK(cudaMemset(t->m, 0, sizeof(float) * t->n_params));
        CK(cudaMemset(t->v, 0, sizeof(float) * t->n_params));
    }
    return 0;
}

int vae21_trainer_get_params(vae21_trainer* t, float* flat_host) {
    if (int rc = trainer_use(t)) return rc;
    if (!flat_host) return fail(VAE21_ERR_ARG, "null output");
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(flat_host, t->p, sizeof(float) * t->n_params, cudaMemcpyDeviceToHost));
    return 0;
}

int vae21_trainer_set_moments(vae21_trainer* t, const float* m_host, const float* v_host) {
    if (int rc = trainer_use(t)) return rc;
    if (!m_host || !v_host) return fail(VAE21_ERR_ARG, "null moments");
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(t->m, m_host, sizeof(float) * t->n_params, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(t->v, v_host, sizeof(float) * t->n_params, cudaMemcpyHostToDevice));
    return 0;
}

int vae21_trainer_get_moments(vae21_trainer* t, float* m_host, float* v_host) {
    if (int rc = trainer_use(t)) return rc;
    if (!m_host || !v_host) return fail(VAE21_ERR_ARG, "null output");
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(m_host, t->m, sizeof(float) * t->n_params, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(v_host, t->v, sizeof(float) * t->n_params, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"

namespace {
// forward + loss (+ backward when grad != nullptr) of one batch; step != nullptr: graph form (batch number read on the device)
int enqueue_step(vae21_trainer* t, const float* x_all, const float* y_all, const float* w_all, const int* idx, int64_t first, int batch,
                 float grad_scale, float* grad, float* loss_sum, cudaStream_t st, const int* step, int stride = 0) {
    const int L = t->n_layers, NO = t->dims[L];
    trk::gather_kernel<<<batch, 128, 0, st>>>(x_all, y_all, w_all, idx, first, batch, t->dims[0], NO, t->act[0], t->yb, t->wb, step,
                                              stride ? stride : batch);
    trainer_forward(t, batch, st);
    float* d_cur = t->delta[0];
    trk::loss_delta_kernel<<<(batch + 7) / 8, 256, 0, st>>>(t->act[L], t->yb, t->wb, batch, NO, grad_scale, grad ? d_cur : nullptr, t->loss_rows);
    trk::loss_sum_kernel<<<1, 32, 0, st>>>(t->loss_rows, batch, loss_sum);
    t->launches += 3;
    if (grad) {
        for (int l = L - 1; l >= 0; --l) {
            const int K = t->dims[l], N = t->dims[l + 1];
            trk::colsum_kernel<<<(N + 127) / 128, 128, 0, st>>>(d_cur, batch, N, grad + t->b_off[l]);
            trk::sgemm<2>(K, N, batch, t->act[l], K, d_cur, N, grad + t->w_off[l], N, nullptr, 0, nullptr, 0, st);
            t->launches += 2;
            if (l > 0) {
                float* d_next = (d_cur == t->delta[0]) ? t->delta[1] : t->delta[0];
                // dL/d(pre-activation of layer l-1) = (D W_l^T) masked by the ReLU of layer l-1 (its output is act[l])
                trk::sgemm<1>(batch, K, N, d_cur, N, t->p + t->w_off[l], N, d_next, K, nullptr, 0, t->relu[l - 1] ? t->act[l] : nullptr, K, st);
                t->launches++;
                d_cur = d_next;
            }
        }
    }
    CK(cudaGetLastError());
    return 0;
}
int enqueue_adam(vae21_trainer* t, const float* grad, float lr_t, float beta1, float beta2, float eps, cudaStream_t st, const float* lr_arr,
                 const int* step) {
    const long long n = t->n_params;
    trk::adam_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(t->p, t->m, t->v, grad, n, lr_t, beta1, beta2, eps, lr_arr, step);
    t->launches++;
    CK(cudaGetLastError());
    return 0;
}
}  // namespace

extern "C" {

int vae21_trainer_forward_backward(vae21_trainer* t, const float* x_all, const float* y_all, const float* w_all, const int* idx, int64_t first,
                                   int batch, float grad_scale, float* grad, float* loss_sum, void* stream) {
    if (int rc = trainer_use(t)) return rc;
    if (!x_all || !y_all || !w_all || !loss_sum) return fail(VAE21_ERR_ARG, "null device pointer");
    if (batch < 1 || batch > t->max_batch) return fail(VAE21_ERR_ARG, "batch %d outside [1,%d]", batch, t->max_batch);
    return enqueue_step(t, x_all, y_all, w_all, idx, first, batch, grad_scale, grad, loss_sum, static_cast<cudaStream_t>(stream), nullptr);
}

int vae21_trainer_adam(vae21_trainer* t, const float* grad, float lr_t, float beta1, float beta2, float eps, void* stream) {
    if (int rc = trainer_use(t)) return rc;
    if (!grad) return fail(VAE21_ERR_ARG, "null gradient");
    return enqueue_adam(t, grad, lr_t, beta1, beta2, eps, static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

int vae21_trainer_epoch(vae21_trainer* t, const float* x_all, const float* y_all, const float* w_all, const int* perm, int64_t n, int batch,
                        float lr, float beta1, float beta2, float eps, int64_t iterations_before, float* loss_sum, void* stream) {
    if (int rc = trainer_use(t)) return rc;
    if (!x_all || !y_all || !w_all || !loss_sum) return fail(VAE21_ERR_ARG, "null device pointer");
    if (n < 0 || batch < 1 || batch > t->max_batch) return fail(VAE21_ERR_ARG, "bad n / batch");
    const int NO = t->dims[t->n_layers];
    cudaStream_t ust = static_cast<cudaStream_t>(stream);
    auto lr_of = [&](int64_t it) {
        return static_cast<float>(static_cast<double>(lr) * std::sqrt(1.0 - std::pow(static_cast<double>(beta2), static_cast<double>(it))) /
                                  (1.0 - std::pow(static_cast<double>(beta1), static_cast<double>(it))));
    };
    const int64_t n_full = n / batch, rem = n - n_full * batch;
    static const bool no_graph = std::getenv("VAE21_TRAIN_NO_GRAPH") != nullptr;
    if (no_graph || !perm || n_full < 2) {
        // plain path: one launch sequence per batch.  Keep at most ~2 x 16 steps (< 800 launches) in flight: past the driver's launch-
        // queue depth every further launch blocks in a slow path (measured 3.2 ms per step instead of 0.3 ms for a whole epoch at once).
        int64_t it = iterations_before;
        int blk = 0, recorded[2] = {0, 0};
        for (int64_t lo = 0, step = 0; lo < n; lo += batch, ++step) {
            if (step > 0 && step % 16 == 0) {
                CK(cudaEventRecord(t->throttle[blk], ust));
                recorded[blk] = 1;
                blk ^= 1;
                if (recorded[blk]) CK(cudaEventSynchronize(t->throttle[blk]));
            }
            const int b = static_cast<int>(std::min<int64_t>(batch, n - lo));
            if (int rc = enqueue_step(t, x_all, y_all, w_all, perm ? perm + lo : nullptr, lo, b, static_cast<float>(1.0 / (static_cast<double>(NO) * b)),
                                      t->grad, loss_sum, ust, nullptr))
                return rc;
            if (int rc = enqueue_adam(t, t->grad, lr_of(++it), beta1, beta2, eps, ust, nullptr, nullptr)) return rc;
        }
        return 0;
    }
    // graph path: every full batch replays ONE executable graph on the trainer's own stream; the batch number and the learning rate
    // of the step come from device memory (d_step, d_lr), the permutation from a trainer-owned copy (stable pointers)
    cudaStream_t gs = t->gstream;
    CK(cudaEventRecord(t->ev_in, ust));
    CK(cudaStreamWaitEvent(gs, t->ev_in, 0));
    if (t->perm_cap < n) {
        if (t->d_perm) cudaFree(t->d_perm);
        t->d_perm = nullptr;
        CK(cudaMalloc(&t->d_perm, sizeof(int) * n));
        t->perm_cap = n;
        if (t->gexec) { cudaGraphExecDestroy(t->gexec); t->gexec = nullptr; }
    }
    if (t->lr_cap < n_full) {
        if (t->d_lr) cudaFree(t->d_lr);
        t->d_lr = nullptr;
        CK(cudaMalloc(&t->d_lr, sizeof(float) * n_full));
        t->lr_cap = n_full;
        if (t->gexec) { cudaGraphExecDestroy(t->gexec); t->gexec = nullptr; }
    }
    CK(cudaMemcpyAsync(t->d_perm, perm, sizeof(int) * n, cudaMemcpyDeviceToDevice, gs));
    std::vector<float> lrs(n_full);
    for (int64_t k = 0; k < n_full; ++k) lrs[k] = lr_of(iterations_before + k + 1);
    CK(cudaMemcpyAsync(t->d_lr, lrs.data(), sizeof(float) * n_full, cudaMemcpyHostToDevice, gs));
    CK(cudaMemsetAsync(t->d_step, 0, sizeof(int), gs));
    CK(cudaMemsetAsync(t->d_loss, 0, sizeof(float), gs));
    CK(cudaStreamSynchronize(gs));  // `lrs` is pageable host memory
    if (!t->gexec || t->g_x != x_all || t->g_y != y_all || t->g_w != w_all || t->g_batch != batch || t->g_b1 != beta1 || t->g_b2 != beta2 ||
        t->g_eps != eps) {
        if (t->gexec) { cudaGraphExecDestroy(t->gexec); t->gexec = nullptr; }
        cudaGraph_t graph = nullptr;
        const long long launches_before = t->launches;
        CK(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_step(t, x_all, y_all, w_all, t->d_perm, 0, batch, static_cast<float>(1.0 / (static_cast<double>(NO) * batch)), t->grad, t->d_loss,
                              gs, t->d_step);
        if (rc == 0) rc = enqueue_adam(t, t->grad, 0.f, beta1, beta2, eps, gs, t->d_lr, t->d_step);
        if (rc == 0) trk::step_inc_kernel<<<1, 1, 0, gs>>>(t->d_step);
        const cudaError_t ce = cudaStreamEndCapture(gs, &graph);
        if (rc != 0 || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            return rc ? rc : fail(VAE21_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
        }
        const cudaError_t ie = cudaGraphInstantiate(&t->gexec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) return fail(VAE21_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(ie));
        t->g_kernels = t->launches - launches_before + 1;  // + step_inc
        t->launches = launches_before;                      // capturing launched nothing
        t->g_x = x_all; t->g_y = y_all; t->g_w = w_all; t->g_batch = batch; t->g_b1 = beta1; t->g_b2 = beta2; t->g_eps = eps;
    }
    for (int64_t k = 0; k < n_full; ++k) CK(cudaGraphLaunch(t->gexec, gs));
    t->launches += n_full * t->g_kernels;
    if (rem > 0) {
        if (int rc = enqueue_step(t, x_all, y_all, w_all, t->d_perm + n_full * batch, 0, static_cast<int>(rem),
                                  static_cast<float>(1.0 / (static_cast<double>(NO) * rem)), t->grad, t->d_loss, gs, nullptr))
            return rc;
        if (int rc = enqueue_adam(t, t->grad, lr_of(iterations_before + n_full + 1), beta1, beta2, eps, gs, nullptr, nullptr)) return rc;
    }
    trk::add_scalar_kernel<<<1, 1, 0, gs>>>(loss_sum, t->d_loss);
    CK(cudaGetLastError());
    CK(cudaEventRecord(t->ev_out, gs));
    CK(cudaStreamWaitEvent(ust, t->ev_out, 0));
    return 0;
}

int vae21_trainer_dp_begin(vae21_trainer* t, const float* x_all, const float* y_all, const float* w_all, const int* perm, int64_t n, int batch,
                           int share_first, int share_rows, float lr, float beta1, float beta2, float eps, int64_t iterations_before,
                           float* grad, float* loss_sum, void* stream) {
    if (int rc = trainer_use(t)) return rc;
    if (!x_all || !y_all || !w_all || !perm || !grad || !loss_sum) return fail(VAE21_ERR_ARG, "null device pointer");
    if (n < 0 || batch < 1 || share_first < 0 || share_rows < 0 || share_first + share_rows > batch || share_rows > t->max_batch)
        return fail(VAE21_ERR_ARG, "bad n / batch / share");
    const int NO = t->dims[t->n_layers];
    cudaStream_t ust = static_cast<cudaStream_t>(stream);
    const int64_t n_full = n / batch;
    auto drop = [&]() {
        if (t->dpA) { cudaGraphExecDestroy(t->dpA); t->dpA = nullptr; }
        if (t->dpB) { cudaGraphExecDestroy(t->dpB); t->dpB = nullptr; }
    };
    if (t->perm_cap < n) {
        CK(cudaStreamSynchronize(ust));
        if (t->d_perm) cudaFree(t->d_perm);
        t->d_perm = nullptr;
        CK(cudaMalloc(&t->d_perm, sizeof(int) * std::max<int64_t>(n, 1)));
        t->perm_cap = n;
        drop();
        if (t->gexec) { cudaGraphExecDestroy(t->gexec); t->gexec = nullptr; }
    }
    if (t->lr_cap < n_full) {
        CK(cudaStreamSynchronize(ust));
        if (t->d_lr) cudaFree(t->d_lr);
        t->d_lr = nullptr;
        CK(cudaMalloc(&t->d_lr, sizeof(float) * std::max<int64_t>(n_full, 1)));
        t->lr_cap = n_full;
        drop();
        if (t->gexec) { cudaGraphExecDestroy(t->gexec); t->gexec = nullptr; }
    }
    CK(cudaMemcpyAsync(t->d_perm, perm, sizeof(int) * n, cudaMemcpyDeviceToDevice, ust));
    std::vector<float> lrs(std::max<int64_t>(n_full, 1));
    for (int64_t k = 0; k < n_full; ++k)
        lrs[k] = static_cast<float>(static_cast<double>(lr) * std::sqrt(1.0 - std::pow(static_cast<double>(beta2), static_cast<double>(iterations_before + k + 1))) /
                                    (1.0 - std::pow(static_cast<double>(beta1), static_cast<double>(iterations_before + k + 1))));
    CK(cudaMemcpyAsync(t->d_lr, lrs.data(), sizeof(float) * n_full, cudaMemcpyHostToDevice, ust));
    CK(cudaMemsetAsync(t->d_step, 0, sizeof(int), ust));
    CK(cudaStreamSynchronize(ust));  // `lrs` is pageable host memory
    t->dp_steps = n_full;
    const bool same = t->dpB && t->dp_x == x_all && t->dp_y == y_all && t->dp_w == w_all && t->dp_grad == grad && t->dp_loss == loss_sum &&
                      t->dp_batch == batch && t->dp_first == share_first && t->dp_rows == share_rows && t->dp_b1 == beta1 && t->dp_b2 == beta2 &&
                      t->dp_eps == eps;
    if (same) return 0;
    drop();
    cudaStream_t gs = t->gstream;
    const long long launches_before = t->launches;
    if (share_rows > 0) {
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_step(t, x_all, y_all, w_all, t->d_perm, share_first, share_rows, static_cast<float>(1.0 / (static_cast<double>(NO) * batch)),
                                    grad, loss_sum, gs, t->d_step, batch);
        const cudaError_t ce = cudaStreamEndCapture(gs, &graph);
        if (rc != 0 || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            t->launches = launches_before;
            return rc ? rc : fail(VAE21_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
        }
        const cudaError_t ie = cudaGraphInstantiate(&t->dpA, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) return fail(VAE21_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(ie));
    }
    t->dpA_kernels = t->launches - launches_before;
    t->launches = launches_before;  // capturing launched nothing
    {
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_adam(t, grad, 0.f, beta1, beta2, eps, gs, t->d_lr, t->d_step);
        trk::step_inc_kernel<<<1, 1, 0, gs>>>(t->d_step);
        const cudaError_t ce = cudaStreamEndCapture(gs, &graph);
        t->launches = launches_before;
        if (rc != 0 || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            drop();
            return rc ? rc : fail(VAE21_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
        }
        const cudaError_t ie = cudaGraphInstantiate(&t->dpB, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { drop(); return fail(VAE21_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(ie)); }
    }
    t->dp_x = x_all; t->dp_y = y_all; t->dp_w = w_all; t->dp_grad = grad; t->dp_loss = loss_sum;
    t->dp_batch = batch; t->dp_first = share_first; t->dp_rows = share_rows; t->dp_b1 = beta1; t->dp_b2 = beta2; t->dp_eps = eps;
    return 0;
}

int vae21_trainer_dp_forward_backward(vae21_trainer* t, void* stream) {
    if (int rc = trainer_use(t)) return rc;
    if (!t->dpB) return fail(VAE21_ERR_STATE, "vae21_trainer_dp_begin has not been called");
    if (!t->dpA) return fail(VAE21_ERR_STATE, "this rank has no rows in a batch (zero its gradient instead)");
    CK(cudaGraphLaunch(t->dpA, static_cast<cudaStream_t>(stream)));
    t->launches += t->dpA_kernels;
    return 0;
}

int vae21_trainer_dp_adam(vae21_trainer* t, void* stream) {
    if (int rc = trainer_use(t)) return rc;
    if (!t->dpB) return fail(VAE21_ERR_STATE, "vae21_trainer_dp_begin has not been called");
    CK(cudaGraphLaunch(t->dpB, static_cast<cudaStream_t>(stream)));
    t->launches += 2;
    return 0;
}

int vae21_trainer_launches(vae21_trainer* t, int64_t* n) {
    if (!t || !n) return fail(VAE21_ERR_ARG, "null argument");
    *n = t->launches;
    return 0;
}

}  // extern "C"
