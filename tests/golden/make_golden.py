"""Generate the committed golden fixtures from the REFERENCE checkout (run in the build
container only; /root/reference does not exist on the GPU box).

    python tests/golden/make_golden.py [/root/reference]

Writes, next to this script:
  ae_chain.npz       the only real trained weights the reference ships
                     (models/autoencoder_based_emulator/{ae_emulator,decoder}.h5), float32, plus
                     float64 known-answer outputs of the chain 7->352->352->352->224->9->32->352->451
                     for fixed normalised inputs (Dense semantics of emulator.py:41-47, :789-790)
  preprocess.npz     inputs and outputs of the reference's REAL preprocess.py (imported by file
                     path: preproc / unpreproc / par_transform) on seeded stand-in training arrays
  tiny_keras.h5      a small Keras-2.x-layout file written by our own writer (loader fixture)
"""
import hashlib
import importlib
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"

kh = importlib.import_module("21cmvae_b200.keras_h5")
from oracle import refmath as rm  # noqa: E402


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    base = os.path.join(REF, "VeryAccurateEmulator/models/autoencoder_based_emulator")
    em = kh.load_dense_chain(os.path.join(base, "ae_emulator.h5"))
    de = kh.load_dense_chain(os.path.join(base, "decoder.h5"))
    ch = em.concat(de)
    rng = np.random.default_rng(20211)
    x = np.concatenate([np.zeros((1, 7)), np.linspace(-1, 1, 7)[None, :], rng.uniform(-1, 1, size=(62, 7))])
    x = x.astype(np.float32)  # what Keras would see
    latent = rm.dense_chain(x, em.kernels, em.biases, em.relu, dtype=np.float64)
    y = rm.dense_chain(x, ch.kernels, ch.biases, ch.relu, dtype=np.float64)
    out = {"x": x, "latent64": latent, "y64": y, "relu": np.array(ch.relu, dtype=np.int8),
           "sha_ae_emulator": sha(os.path.join(base, "ae_emulator.h5")),
           "sha_decoder": sha(os.path.join(base, "decoder.h5"))}
    for i, (k, b) in enumerate(zip(ch.kernels, ch.biases)):
        out[f"k{i}"] = k
        out[f"b{i}"] = b
    np.savez_compressed(os.path.join(HERE, "ae_chain.npz"), **out)

    # the reference's real preprocess.py, loaded by path (importing the package would try to
    # download the dataset, VeryAccurateEmulator/__init__.py:8-16)
    spec = importlib.util.spec_from_file_location("ref_preprocess", os.path.join(REF, "VeryAccurateEmulator/preprocess.py"))
    ref_pp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_pp)
    par_train = rm.draw_params(600, seed=11, zero_fx_frac=0.02)
    params64 = rm.draw_params(257, seed=12, zero_fx_frac=0.05)
    params32 = params64.astype(np.float32)
    sig_train = (rng.normal(size=(120, 451)) * 40 - 30).astype(np.float32)
    sig = (rng.normal(size=(9, 451)) * 40 - 30).astype(np.float32)
    np.savez_compressed(
        os.path.join(HERE, "preprocess.npz"),
        par_train=par_train, params64=params64, params32=params32, sig_train=sig_train, sig=sig,
        pt64=ref_pp.par_transform(params64, par_train), pt32=ref_pp.par_transform(params32, par_train),
        pt_single=ref_pp.par_transform(params64[0], par_train),
        pre=ref_pp.preproc(sig, sig_train), unpre=ref_pp.unpreproc(sig, sig_train),
        unpre64=ref_pp.unpreproc(sig.astype(np.float64), sig_train))

    ks, bs, relu = rm.glorot_chain((7, 16, 24, 11), seed=5)
    kh.save_dense_chain(os.path.join(HERE, "tiny_keras.h5"),
                        kh.DenseChainWeights(ks, bs, relu, ["em_hidden_layer_0", "em_hidden_layer_1", "dense_16"], "emulator"))
    np.savez(os.path.join(HERE, "tiny_keras_expected.npz"), **{f"k{i}": k for i, k in enumerate(ks)},
             **{f"b{i}": b for i, b in enumerate(bs)})
    print("golden written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
