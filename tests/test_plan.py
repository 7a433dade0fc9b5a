"""CPU: the tensor-core schedule planner (csrc/tc_kernel.cuh build_plan / build_iters) replayed on the host through the C ABI
(vae21_check_plan): issue-table structure, commit counts and -- above all -- the phase parities of every barrier wait, for the
emulator stacks and a few hundred random ones.  No GPU needed."""
import numpy as np
import pytest

from conftest import pkg


def test_reference_stacks_have_consistent_schedules():
    L = pkg("_lib")
    for dims in [(7, 288, 352, 288, 224, 451),                      # DirectEmulator (emulator.py:37-47, hidden_dims of the README)
                 (7, 352, 352, 352, 224, 9, 32, 352, 451),          # autoencoder-based emulator chain (ae_emulator.h5 -> decoder.h5)
                 (7, 16, 11), (3, 40, 24, 451), (16, 100, 200, 100, 30), (7, 480, 64), (5, 33, 47, 19, 130, 7)]:
        rc, why = L.check_plan(dims)
        assert rc == 0, (dims, rc, why)


def test_stacks_that_do_not_fit_are_refused_not_mis_scheduled():
    L = pkg("_lib")
    rc, why = L.check_plan((17, 64, 8))          # more than 16 inputs
    assert rc == 1 and why
    rc, why = L.check_plan((7, 8))               # a single layer
    assert rc == 1 and why


def test_random_stacks_never_get_an_inconsistent_schedule():
    L = pkg("_lib")
    rng = np.random.default_rng(2024)
    ok = 0
    for _ in range(300):
        n_layers = int(rng.integers(2, 9))
        dims = [int(rng.integers(1, 17))] + [int(rng.integers(1, 481)) for _ in range(n_layers - 1)] + [int(rng.integers(1, 600))]
        rc, why = L.check_plan(dims)
        assert rc in (0, 1), (dims, rc, why)     # 2 = the planner produced a schedule its own replay rejects
        ok += rc == 0
    assert ok > 100                              # most stacks of this size fit
