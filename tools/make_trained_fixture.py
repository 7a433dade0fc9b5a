"""Train a DirectEmulator-shaped network (7-288-352-288-224-451) to TRAINED-SCALE weights and store it as a test fixture.

The reference's own DirectEmulator weights (models/emulator.h5) and its dataset are absent from the checkout,
so every test of the DirectEmulator architecture used seeded Glorot weights (|W| <= 0.1), on which the tensor-core operand formats
look ~70x more accurate than on trained weights (|W| up to 2.4 in the shipped autoencoder-based emulator).  This script makes
the missing fixture the way BASELINE.json config 5 describes: the reference's shipped, really trained autoencoder-based emulator
(ae_emulator.h5 + decoder.h5, committed as tests/golden/ae_chain.npz) is the TEACHER; 30,000 parameter vectors drawn from the
prior ranges are pushed through it; the student is trained on those signals with THIS repository's CUDA trainer through the
reference's own API (DirectEmulator.train, emulator.py:339-381: batch 256, relative-MSE loss, Adam(0.01), EarlyStopping +
ReduceLROnPlateau as in notebooks/Training.ipynb cells 4-5).  Needs a B200:

    gpurun -- python tools/make_trained_fixture.py 300        # writes gpurun_out/direct_trained.{h5,npz}
    cp gpurun_out/direct_trained.h5 gpurun_out/direct_trained.npz tests/golden/

The trainer is bitwise deterministic for a fixed seed, so the fixture can be regenerated.
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refmath as rm  # noqa: E402  (teacher evaluation and synthetic parameter draws: fixture generation only)

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 300
out_dir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
emu_mod = importlib.import_module("21cmvae_b200.emulator")
tr = importlib.import_module("21cmvae_b200.training")

g = np.load(os.path.join(ROOT, "tests", "golden", "ae_chain.npz"))
nl = sum(1 for k in g.files if k.startswith("k") and k[1:].isdigit())
t_ks, t_bs, t_relu = [g[f"k{i}"] for i in range(nl)], [g[f"b{i}"] for i in range(nl)], [bool(r) for r in g["relu"]]

# 30k draws from the prior, 1 % with fx exactly 0; teacher signals in mK around a 21-cm-like mean profile
par = rm.draw_params(30_000, seed=5, zero_fx_frac=0.01)
pmin, pmax = rm.prior_par_stats()
x = rm.par_transform_cached(par, pmin, pmax).astype(np.float32)
y_sigma = rm.dense_chain(x, t_ks, t_bs, t_relu, dtype=np.float64)
z = np.linspace(5, 50, 451)
mu_t = (-90.0 * np.exp(-(((z - 17.0) / 6.0) ** 2)) + 8.0 * np.exp(-(((z - 9.0) / 2.5) ** 2))).astype(np.float32)
signals = (50.0 * y_sigma + mu_t).astype(np.float32)
n_tr, n_val = 27_000, 1_500
emu = emu_mod.DirectEmulator(par_train=par[:n_tr], par_val=par[n_tr:n_tr + n_val], par_test=par[n_tr + n_val:],
                             signal_train=signals[:n_tr], signal_val=signals[n_tr:n_tr + n_val], signal_test=signals[n_tr + n_val:])
emu.emulator = emu_mod._gen_model(7, emu_mod.hidden_dims, 451, "relu", name="emulator", seed=2022)
emu.emulator.compile(optimizer=tr.Adam(0.01), loss=emu_mod.relative_mse_loss(emu.signal_train))
cbs = [tr.EarlyStopping(monitor="val_loss", patience=25, min_delta=1e-10, restore_best_weights=True),
       tr.ReduceLROnPlateau(monitor="val_loss", patience=5, factor=0.9, min_delta=5e-9, min_lr=1e-4)]
t0 = time.perf_counter()
loss, val_loss = emu.train(epochs, callbacks=cbs, verbose=0, seed=5)
secs = time.perf_counter() - t0
err = emu.test_error(relative=True, precision="fp32")
err_mk = emu.test_error(relative=False, precision="fp32")
w = emu.emulator.weights
wmax = [float(np.abs(k).max()) for k in w.kernels]
os.makedirs(out_dir, exist_ok=True)
# the reference's layer names (notebooks/sample_notebook.ipynb cell 3)
w.layer_names = ["em_hidden_layer_0", "em_hidden_layer_1", "em_hidden_layer_2", "em_hidden_layer_3", "dense_16"]
emu.save_model(os.path.join(out_dir, "direct_trained.h5"))
s = emu.stats
idx = np.arange(0, len(emu.par_test), max(1, len(emu.par_test) // 256))[:256]
np.savez_compressed(os.path.join(out_dir, "direct_trained.npz"), par_min=np.asarray(s.par_min, np.float64), par_max=np.asarray(s.par_max, np.float64),
                    sig_mean=np.asarray(s.sig_mean, np.float32), sig_std=np.float32(s.sig_std), par_test=emu.par_test[idx],
                    signal_test=emu.signal_test[idx])
print(json.dumps({"epochs_run": len(loss), "seconds": secs, "loss_first": loss[0], "loss_last": loss[-1], "val_loss_best": min(val_loss),
                  "test_rel_err_mean_pct": float(np.mean(err)), "test_rel_err_median_pct": float(np.median(err)),
                  "test_abs_err_mean_mK": float(np.mean(err_mk)), "max_abs_weight_per_layer": wmax, "sig_std": float(s.sig_std)}))
