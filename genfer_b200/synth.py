"""Synthetic coefficient tensors for the product sweep: documented PRNG, fixed seeds."""
import numpy as np

SEED_X, SEED_Y = 20230517, 20231210


def splitmix64(seed: int, n: int) -> np.ndarray:
    """Documented PRNG for synthetic tensors: uniform [0,1) doubles from splitmix64."""
    out = np.empty(n, dtype=np.uint64)
    state = np.uint64(seed)
    with np.errstate(over="ignore"):
        idx = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) + state
        z = idx
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    out[:] = z
    return (out >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def synth_uniform(shape, seed):
    """Distribution U of SURVEY 8(d): iid uniform [0,1)."""
    return splitmix64(seed, int(np.prod(shape))).reshape(shape)


def synth_pgf(shape, seed):
    """Distribution P of SURVEY 8(d): separable Poisson-like decay times (1 + 0.1 u)."""
    from math import lgamma
    u = synth_uniform(shape, seed)
    t = np.ones(shape)
    for ax, d in enumerate(shape):
        lam = d / 4.0
        k = np.arange(d)
        w = np.exp(-lam + k * np.log(lam) - np.array([lgamma(i + 1.0) for i in k]))
        sh = [1] * len(shape)
        sh[ax] = d
        t = t * w.reshape(sh)
    return t * (1.0 + 0.1 * u)
