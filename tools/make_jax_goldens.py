#!/usr/bin/env python
"""Dump jax.random known answers for the exact call patterns of jQMC's hot path into tests/golden/jax_rng_goldens.npz.

Run this WHERE JAX IS INSTALLED (pinned range of the reference: jax>=0.6,<0.8, setup.cfg:25-26):

    JAX_PLATFORMS=cpu python tools/make_jax_goldens.py

It cannot run in the build image or on the GPU box of this project (no jax / jaxlib wheel, no network: probe recorded in
DESIGN.md §6); until somebody runs it, tests/test_rng.py::test_against_jax_goldens is skipped and the bit-exact claims of the
engine rest on the restatement oracle/jaxrng.py alone ("parity unpinned" for split counters, 64-bit assembly, randint, the fp64
mantissa fill and erf_inv).  Call sites covered (jqmc/jqmc_mcmc.py:4322-4367, 4499-4500, 4232-4233; jqmc/jqmc_gfmc.py:5275-5283,
4813-4815, 5059):  PRNGKey(seed), split(key), split(key, n), randint(key, (), 0, n) [int64 under x64], normal(key, ()),
uniform(key, ()), uniform(key, (3,), minval=-2 pi, maxval=2 pi).
"""

import os

import numpy as np


def main():
    import jax

    jax.config.update("jax_enable_x64", True)
    import jax.numpy as jnp
    from jax import random as jr

    out = {"jax_version": np.array(jax.__version__), "threefry_partitionable": np.array(bool(jax.config.jax_threefry_partitionable))}
    seeds = [0, 42, 34456, 2**33 + 5]
    for s in seeds:
        key = jr.PRNGKey(s)
        out[f"key_{s}"] = np.asarray(key, dtype=np.uint32)
        out[f"split_{s}"] = np.asarray(jr.split(key), dtype=np.uint32)
        out[f"split4_{s}"] = np.asarray(jr.split(key, 4), dtype=np.uint32)
        k = key
        chain = []
        for _ in range(6):  # the split chain of one Metropolis proposal
            k, sub = jr.split(k)
            chain.append(np.asarray(sub, dtype=np.uint32))
        out[f"chain_{s}"] = np.stack(chain)
        sub = jr.split(key)[1]
        for n in (3, 4, 8, 13):
            out[f"randint_{s}_{n}"] = np.asarray(jr.randint(sub, (), 0, n), dtype=np.int64)
        out[f"normal_{s}"] = np.asarray(jr.normal(sub, (), dtype=jnp.float64))
        out[f"uniform_{s}"] = np.asarray(jr.uniform(sub, (), dtype=jnp.float64))
        out[f"angles_{s}"] = np.asarray(jr.uniform(sub, (3,), dtype=jnp.float64, minval=-2 * np.pi, maxval=2 * np.pi))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "jax_rng_goldens.npz")
    np.savez(path, **out)
    print("wrote", path, "jax", jax.__version__)


if __name__ == "__main__":
    main()
