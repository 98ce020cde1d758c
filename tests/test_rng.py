"""jax.random restatement: Threefry-2x32 known-answer vectors (Random123 kat_vectors; the same three
appear in JAX's own testThreefry2x32) and internal consistency of the host/oracle implementations."""

import numpy as np
from scipy.special import erfinv

from jqmc_b200 import rng_host
from oracle import jaxrng as R


def test_threefry_known_answers():
    assert R.threefry2x32(0, 0, 0, 0) == (0x6B200159, 0x99BA4EFE)
    assert R.threefry2x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF) == (0x1CB996FC, 0xBB002BE7)
    assert R.threefry2x32(0x13198A2E, 0x03707344, 0x243F6A88, 0x85A308D3) == (0xC4923A9C, 0x483DF7A0)


def test_legacy_split_known_answer():
    # jax.random.split(PRNGKey(0)) in the pre-partitionable layout: counters (0,1 | 2,3)
    a = R.threefry2x32(0, 0, 0, 2)
    b = R.threefry2x32(0, 0, 1, 3)
    assert [a[0], b[0], a[1], b[1]] == [4146024105, 967050713, 2718843009, 1272950319]


def test_host_split_matches_oracle():
    key = rng_host.PRNGKey(34456 * 3)
    assert tuple(int(x) for x in key) == R.PRNGKey(34456 * 3)
    ks = rng_host.split(key, 17)
    ref = R.split(R.PRNGKey(34456 * 3), 17)
    assert [tuple(int(x) for x in k) for k in ks] == ref


def test_erf_inv_and_ranges():
    xs = np.linspace(-0.9, 0.9, 501)
    err = [abs(R.erf_inv(x) - erfinv(x)) / max(abs(erfinv(x)), 1e-300) for x in xs if x != 0.0]
    assert max(err) < 5e-15
    key = R.PRNGKey(7)
    us = [R.uniform(k) for k in R.split(key, 200)]
    assert 0.0 <= min(us) and max(us) < 1.0
    ints = [R.randint(k, 0, 8) for k in R.split(key, 400)]
    assert set(ints) == set(range(8))
    ns = np.array([R.normal(k) for k in R.split(key, 2000)])
    assert abs(ns.mean()) < 0.1 and abs(ns.std() - 1.0) < 0.1
