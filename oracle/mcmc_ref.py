"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the ensemble stretch move of csrc/mcmc_api.cuh (Goodman & Weare 2010; the
reference itself has no sampler: emulator.py:383-407 is what a user's likelihood calls).  Same stateless generator, same float64
arithmetic order as the CUDA kernels, chi^2 from the float64 oracle (oracle/refmath.py).  Imported by tests only."""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(z):
    z = (np.asarray(z, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M64
    return z ^ (z >> np.uint64(31))


def u01(seed, step, half, walker, draw):
    with np.errstate(over="ignore"):
        k = mix64(mix64(mix64(np.uint64(seed)) ^ np.uint64(step * 2 + half)) ^ (np.asarray(walker, dtype=np.uint64) * np.uint64(4) + np.uint64(draw)))
    return (k >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def propose(x, half, seed, step, a=2.0):
    """Proposals of the active half: (y [m, d], z [m], partner index j [m])."""
    n, d = x.shape
    m = n // 2
    s0, c0 = (m, 0) if half else (0, m)
    i = np.arange(m)
    j = np.minimum((u01(seed, step, half, i, 0) * float(m)).astype(np.int64), m - 1)
    r = (a - 1.0) * u01(seed, step, half, i, 1) + 1.0
    z = (r * r) / a
    xs, xc = x[s0:s0 + m], x[c0 + j]
    y = xc + z[:, None] * (xs - xc)
    return y, z, j


def half_step(x, logp, half, seed, step, log_prob, a=2.0):
    """One half-step in place; `log_prob(y) -> [m]` float64.  Returns (accept mask, ln r - ln u margin)."""
    n, d = x.shape
    m = n // 2
    s0 = m if half else 0
    y, z, _ = propose(x, half, seed, step, a)
    lpy = log_prob(y)
    lnr = (d - 1) * np.log(z) + lpy - logp[s0:s0 + m]
    with np.errstate(divide="ignore"):
        lnu = np.log(u01(seed, step, half, np.arange(m), 2))
    acc = lnu < lnr
    x[s0:s0 + m][acc] = y[acc]
    logp[s0:s0 + m][acc] = lpy[acc]
    return acc, lnr - lnu
