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


def _normal_f32(key, fold):
    """jax.random.normal(key, ()) in float32 from the restated pieces: one Threefry block on counter (0, 0), 32 random bits
    = fold(hi, lo), 23-bit mantissa fill, XLA's single-precision erf_inv polynomial (Giles)."""
    f32 = np.float32
    hi, lo = R.threefry2x32(key[0], key[1], 0, 0)
    u = np.uint32((fold(hi, lo) >> 9) | 0x3F800000).view(f32) - f32(1)
    m = np.nextafter(f32(-1), f32(0))
    x = max(m, f32(u * f32(f32(1) - m) + m))
    w = f32(-np.log(f32((f32(1) - x) * (f32(1) + x))))
    if w < f32(5):
        w = f32(w - f32(2.5))
        cs = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941]
    else:
        w = f32(np.sqrt(w) - f32(3))
        cs = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682]
    p = f32(cs[0])
    for c in cs[1:]:
        p = f32(f32(c) + f32(p * w))
    return float(f32(np.sqrt(f32(2))) * f32(p * x))


def test_scalar_draw_known_answers_from_jax_documentation():
    """Known answers printed in the JAX documentation pin the key layout (PRNGKey(seed) = [0, seed]), the counter of a scalar
    draw ((0, 0)) and the partitionable bit layout: `random.normal(random.key(42))` = -0.028304616 (current docs,
    jax_threefry_partitionable=True: 32 bits = hi ^ lo); the pre-0.5 docs print -0.18471177 for PRNGKey(42) and -0.20584226
    for PRNGKey(0) (32 bits = first output word)."""
    np.testing.assert_allclose(_normal_f32(R.PRNGKey(42), lambda h, l: h ^ l), -0.028304616, rtol=2e-7)
    np.testing.assert_allclose(_normal_f32(R.PRNGKey(42), lambda h, l: h), -0.18471177, rtol=2e-7)
    np.testing.assert_allclose(_normal_f32(R.PRNGKey(0), lambda h, l: h), -0.20584226, rtol=2e-7)


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


def test_against_jax_goldens():
    """jax.random itself, when somebody has run tools/make_jax_goldens.py where JAX is installed (it is not in this project's
    image: DESIGN.md §6).  Every call pattern of the hot path -- key layout, split counters, the per-proposal split chain,
    randint's two-draw composition, the fp64 mantissa fill of uniform, erf_inv of normal -- against oracle/jaxrng.py."""
    import os

    import pytest

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jax_rng_goldens.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/jax_rng_goldens.npz not generated yet (needs a JAX install: tools/make_jax_goldens.py)")
    g = np.load(path)
    assert bool(g["threefry_partitionable"]), "goldens were generated with the legacy (non-partitionable) Threefry layout"
    for s in (0, 42, 34456, 2**33 + 5):
        key = R.PRNGKey(s)
        assert key == tuple(int(x) for x in g[f"key_{s}"])
        assert R.split(key) == [tuple(int(x) for x in k) for k in g[f"split_{s}"]]
        assert R.split(key, 4) == [tuple(int(x) for x in k) for k in g[f"split4_{s}"]]
        k = key
        for ref in g[f"chain_{s}"]:
            k, sub = R.split(k)
            assert sub == tuple(int(x) for x in ref)
        sub = R.split(key)[1]
        for n in (3, 4, 8, 13):
            assert R.randint(sub, 0, n) == int(g[f"randint_{s}_{n}"])
        assert R.normal(sub) == float(g[f"normal_{s}"])
        assert R.uniform(sub) == float(g[f"uniform_{s}"])
        np.testing.assert_array_equal(np.array(R.uniform(sub, 3, -2 * np.pi, 2 * np.pi)), g[f"angles_{s}"])
