"""GPU: the FP32 kernel against the plain-C chain that uses the kernel's own arithmetic order (oracle/chain_fp32.c).
Kept in a file of its own that sorts last: it was added after the round's GPU budget was spent, so its first run on a GPU is
the driver's -- under `-x` it must not be able to cut the rest of the suite short."""
import numpy as np
import pytest

from conftest import pkg

pytestmark = pytest.mark.gpu


def test_fp32_chain_against_the_c_oracle_in_the_kernels_own_order(rm, c_chain, trained_fixture):
    """The FP32 kernel accumulates every output in k order with fused multiply-adds and adds the bias afterwards
    (csrc/fp32_pipe_kernel.cuh); oracle/chain_fp32.c does exactly that on the host with fmaf.  On the normalised chain
    (`emu.emulator.predict`: no prologue, no de-normalisation) the two must agree far inside the 1e-5 tolerance -- asserted at 5e-6
    of the amplitude, a bound that holds with margin for ANY fp32 summation order (tests/test_oracle.py: all within 1.1e-6) -- and the share of outputs that are numerically
    IDENTICAL is reported as a warning in the test summary (expected: all of them)."""
    import warnings

    f = trained_fixture
    emu = pkg("emulator")
    kh = pkg("keras_h5")
    model = emu.DenseModel(kh.DenseChainWeights(f["kernels"], f["biases"], f["relu"], name="emulator"), device=0)
    x32 = rm.par_transform_cached(rm.draw_params(4099, seed=77), f["pmin"], f["pmax"]).astype(np.float32)
    got = np.asarray(model.predict(x32, precision="fp32"))
    want = c_chain(x32, f["kernels"], f["biases"], f["relu"])
    assert got.shape == want.shape and got.dtype == np.float32
    amp = np.max(np.abs(want), axis=1, keepdims=True)
    assert float(np.max(np.abs(got.astype(np.float64) - want) / amp)) <= 5e-6
    same = float(np.mean(got == want))
    warnings.warn(f"FP32 kernel vs plain-C sequential-fmaf chain: {same:.6f} of {got.size} outputs identical, "
                  f"max |diff| / amplitude {float(np.max(np.abs(got.astype(np.float64) - want) / amp)):.2e}")
