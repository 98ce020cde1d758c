"""NumPy restatement of the ``jax.random`` stream jQMC consumes (TEST INFRASTRUCTURE).

jQMC pins ``jax>=0.6.0,<0.8.0`` (setup.cfg:25-26) and enables x64, so the semantics restated here
are: default PRNG impl ``threefry2x32`` with ``jax_threefry_partitionable=True``; raw keys
``uint32[2]``; ``split`` = Threefry on the 64-bit counters (0,i); ``random_bits`` (64-bit) =
``hi<<32 | lo`` of Threefry on counter (0,i); ``uniform`` fp64 = 52-bit mantissa fill;
``normal`` = sqrt(2)*erf_inv(uniform(nextafter(-1,0), 1)) with XLA's fp64 ``erf_inv`` (Giles'
polynomial); ``randint`` (int64) = two 64-bit draws combined modulo the span.

JAX is a third-party dependency that is absent from /root/reference and from this image, so this
file follows JAX's published algorithm.  The Threefry block function is pinned against the
Random123 known-answer vectors (tests/test_rng.py); the derived stream is pinned only through the scalar-draw known answers of tests/test_rng.py ("parity unpinned" for the 64-bit draws).

Call sites in the reference: jqmc/jqmc_mcmc.py:4322-4367, 4499-4500, 4232-4233;
jqmc/jqmc_gfmc.py:5275-5283, 4813-4815, 5059.
"""

from __future__ import annotations

import numpy as np

_M32 = 0xFFFFFFFF
_M64 = 0xFFFFFFFFFFFFFFFF
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return ((x << r) | (x >> (32 - r))) & _M32


def threefry2x32(k0: int, k1: int, x0: int, x1: int):
    """Threefry-2x32, 20 rounds (Salmon et al. 2011), the block function behind jax.random."""
    ks = (k0 & _M32, k1 & _M32, (k0 ^ k1 ^ 0x1BD11BDA) & _M32)
    x0 = (x0 + ks[0]) & _M32
    x1 = (x1 + ks[1]) & _M32
    for i in range(5):
        for r in _ROT[i % 2]:
            x0 = (x0 + x1) & _M32
            x1 = _rotl(x1, r)
            x1 ^= x0
        x0 = (x0 + ks[(i + 1) % 3]) & _M32
        x1 = (x1 + ks[(i + 2) % 3] + i + 1) & _M32
    return x0, x1


def PRNGKey(seed: int):
    seed = int(seed) & _M64
    return (seed >> 32) & _M32, seed & _M32


def split(key, num: int = 2):
    """jax.random.split (partitionable / "foldlike"): child i = Threefry(key, (0, i))."""
    return [threefry2x32(key[0], key[1], 0, i) for i in range(num)]


def random_bits64(key, n: int | None = None):
    """64 random bits per element; shape () when n is None, else (n,)."""
    if n is None:
        hi, lo = threefry2x32(key[0], key[1], 0, 0)
        return (hi << 32) | lo
    return [(lambda h, l: (h << 32) | l)(*threefry2x32(key[0], key[1], 0, i)) for i in range(n)]


def _bits_to_unit(bits: int) -> float:
    """[0,1) double from 64 random bits: mantissa fill, minus 1."""
    u = np.array([(bits >> 12) | 0x3FF0000000000000], dtype=np.uint64).view(np.float64)[0]
    return float(u - 1.0)


def uniform(key, n: int | None = None, minval: float = 0.0, maxval: float = 1.0):
    def one(bits):
        f = _bits_to_unit(bits)
        return max(minval, f * (maxval - minval) + minval)

    if n is None:
        return one(random_bits64(key))
    return [one(b) for b in random_bits64(key, n)]


_W_LT_6_25 = (
    -3.6444120640178196996e-21, -1.685059138182016589e-19, 1.2858480715256400167e-18,
    1.115787767802518096e-17, -1.333171662854620906e-16, 2.0972767875968561637e-17,
    6.6376381343583238325e-15, -4.0545662729752068639e-14, -8.1519341976054721522e-14,
    2.6335093153082322977e-12, -1.2975133253453532498e-11, -5.4154120542946279317e-11,
    1.051212273321532285e-09, -4.1126339803469836976e-09, -2.9070369957882005086e-08,
    4.2347877827932403518e-07, -1.3654692000834678645e-06, -1.3882523362786468719e-05,
    0.0001867342080340571352, -0.00074070253416626697512, -0.0060336708714301490533,
    0.24015818242558961693, 1.6536545626831027356,
)  # fmt: skip
_W_LT_16 = (
    2.2137376921775787049e-09, 9.0756561938885390979e-08, -2.7517406297064545428e-07,
    1.8239629214389227755e-08, 1.5027403968909827627e-06, -4.013867526981545969e-06,
    2.9234449089955446044e-06, 1.2475304481671778723e-05, -4.7318229009055733981e-05,
    6.8284851459573175448e-05, 2.4031110387097893999e-05, -0.0003550375203628474796,
    0.00095328937973738049703, -0.0016882755560235047313, 0.0024914420961078508066,
    -0.0037512085075692412107, 0.005370914553590063617, 1.0052589676941592334,
    3.0838856104922207635,
)  # fmt: skip
_W_GE_16 = (
    -2.7109920616438573243e-11, -2.5556418169965252055e-10, 1.5076572693500548083e-09,
    -3.7894654401267369937e-09, 7.6157012080783393804e-09, -1.4960026627149240478e-08,
    2.9147953450901080826e-08, -6.7711997758452339498e-08, 2.2900482228026654717e-07,
    -9.9298272942317002539e-07, 4.5260625972231537039e-06, -1.9681778105531670567e-05,
    7.5995277030017761139e-05, -0.00021503011930044477347, -0.00013871931833623122026,
    1.0103004648645343977, 4.8499064014085844221,
)  # fmt: skip


def erf_inv(x: float) -> float:
    """fp64 erf_inv as XLA evaluates it: Giles, "Approximating the erfinv function" (2010)."""
    if abs(x) == 1.0:
        return float(np.copysign(np.inf, x))
    w = -float(np.log1p(-x * x))
    if w < 6.25:
        w -= 3.125
        coefs = _W_LT_6_25
    elif w < 16.0:
        w = float(np.sqrt(w)) - 3.25
        coefs = _W_LT_16
    else:
        w = float(np.sqrt(w)) - 5.0
        coefs = _W_GE_16
    p = coefs[0]
    for c in coefs[1:]:
        p = c + p * w
    return p * x


_NORMAL_LO = float(np.nextafter(np.float64(-1.0), np.float64(0.0)))
_SQRT2 = float(np.sqrt(2.0))


def normal(key):
    u = uniform(key, None, _NORMAL_LO, 1.0)
    return _SQRT2 * erf_inv(u)


def randint(key, minval: int, maxval: int) -> int:
    """jax.random.randint for int64 (x64 enabled): two 64-bit draws, modulo-span combination."""
    k1, k2 = split(key, 2)
    hi, lo = random_bits64(k1), random_bits64(k2)
    span = (maxval - minval) & _M64
    if maxval <= minval:
        span = 1
    mult = (1 << 32) % span
    mult = (mult * mult) % span
    off = (((hi % span) * mult) & _M64) + (lo % span)
    off = (off & _M64) % span
    return minval + off
