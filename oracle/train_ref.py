"""CPU restatement of one training step of the reference (TEST INFRASTRUCTURE ONLY -- the product never imports oracle/).

Follows VeryAccurateEmulator/emulator.py:
  :51-83   relative_mse_loss: loss_i = mean_j (y_ij - p_ij)^2 / amp_i^2, amp_i = max_j |y_ij + mean/std|
  :339-381 train -> tf.keras Model.fit(batch_size=256): minimises the batch MEAN of loss_i
  notebooks/Training.ipynb cell 4: Adam(learning_rate=0.01) with the Keras-2.x defaults beta1 0.9, beta2 0.999, eps 1e-7
Keras Adam (non-amsgrad): lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);  m, v exponential averages;  p -= lr_t m / (sqrt(v) + eps).
Parity status: the arithmetic lives in TensorFlow (absent here) -> pinned against torch autograd (tests/test_training.py), an
independent implementation of the same derivative, not against TensorFlow itself.
"""
import numpy as np


def sample_weights(y_proc, mean_over_std):
    """w_i = 1 / amp_i^2 with amp_i = max_j |y_ij + mean_j / std| (emulator.py:70-76)."""
    amp = np.max(np.abs(np.asarray(y_proc, np.float64) + np.asarray(mean_over_std, np.float64)), axis=1)
    return 1.0 / amp**2


def flatten(kernels, biases):
    return np.concatenate([np.concatenate([np.asarray(k).ravel(), np.asarray(b).ravel()]) for k, b in zip(kernels, biases)])


def unflatten(flat, dims):
    ks, bs, off = [], [], 0
    for l in range(len(dims) - 1):
        n = dims[l] * dims[l + 1]
        ks.append(np.asarray(flat[off:off + n]).reshape(dims[l], dims[l + 1]))
        off += n
        bs.append(np.asarray(flat[off:off + dims[l + 1]]))
        off += dims[l + 1]
    return ks, bs


def loss_and_grad(x, y, w, kernels, biases, relu, grad_scale, dtype=np.float64):
    """Per-sample losses and the flat gradient of grad_scale * n_out * sum_i loss_i (grad_scale = 1/(n_out * B): batch mean)."""
    h = [np.asarray(x, dtype)]
    for k, b, r in zip(kernels, biases, relu):
        z = h[-1] @ np.asarray(k, dtype) + np.asarray(b, dtype)
        h.append(np.maximum(z, 0) if r else z)
    d = np.asarray(y, dtype) - h[-1]
    n_out = d.shape[1]
    loss_rows = np.mean(d * d, axis=1) * np.asarray(w, dtype)
    delta = -2.0 * d * np.asarray(w, dtype)[:, None] * grad_scale
    gk, gb = [None] * len(kernels), [None] * len(kernels)
    for l in range(len(kernels) - 1, -1, -1):
        gb[l] = delta.sum(axis=0)
        gk[l] = h[l].T @ delta
        if l > 0:
            delta = delta @ np.asarray(kernels[l], dtype).T
            if relu[l - 1]:
                delta = delta * (h[l] > 0)
    return loss_rows, flatten(gk, gb)


def adam_step(p, m, v, g, lr, t, beta1=0.9, beta2=0.999, eps=1e-7):
    """In-place Keras Adam update number t (1-based)."""
    lr_t = lr * np.sqrt(1.0 - beta2**t) / (1.0 - beta1**t)
    m[:] = beta1 * m + (1 - beta1) * g
    v[:] = beta2 * v + (1 - beta2) * g * g
    p[:] = p - lr_t * m / (np.sqrt(v) + eps)
    return lr_t
