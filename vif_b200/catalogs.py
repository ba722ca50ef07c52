"""Synthetic catalogs of SURVEY.md section 8(d): one counter-based generator so that any row can be
regenerated anywhere (host numpy for <= 1e7 rows, device torch for larger ones) with identical bits.

    u(seed, stream, i) = (mix64(mix64(seed ^ stream*0xD6E8FEB86659FD93) + i*0x9E3779B97F4A7C15) >> 11) * 2^-53

with the splitmix64 finaliser as mix64. Integer work only up to the final scale, so host and device
agree bit for bit on u; coordinates that need transcendentals (C3 offsets, C4/C5 declinations) are
generated ONCE and the same arrays are handed to both the oracle and the GPU path.
"""
import numpy as np

_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB
_GOLD = 0x9E3779B97F4A7C15
_STREAM = 0xD6E8FEB86659FD93
_MASK = (1 << 64) - 1


def _mix64_int(z):
    z &= _MASK
    z = ((z ^ (z >> 30)) * _M1) & _MASK
    z = ((z ^ (z >> 27)) * _M2) & _MASK
    return z ^ (z >> 31)


def _mix64_np(z):
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
        return z ^ (z >> np.uint64(31))


def uniform(seed, stream, n, start=0, index=None):
    """u(seed, stream, i) for i in [start, start+n) (or for the given uint64 index array)."""
    base = np.uint64(_mix64_int(seed ^ ((stream * _STREAM) & _MASK)))
    i = np.arange(start, start + n, dtype=np.uint64) if index is None else index.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = _mix64_np(base + i * np.uint64(_GOLD))
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def _s64(v):
    """python int (mod 2^64) -> signed int64 value with the same bits (torch has no uint64 maths)."""
    v &= _MASK
    return v - (1 << 64) if v >= (1 << 63) else v


def uniform_torch(seed, stream, n, device, start=0, index=None):
    """Same bits as uniform(), computed with wrapping int64 arithmetic on `device`."""
    import torch

    def lsr(z, k):   # logical shift right on int64 bit patterns
        return (z >> k) & ((1 << (64 - k)) - 1)

    def mix(z):
        z = (z ^ lsr(z, 30)) * _s64(_M1)
        z = (z ^ lsr(z, 27)) * _s64(_M2)
        return z ^ lsr(z, 31)

    base = _s64(_mix64_int(seed ^ ((stream * _STREAM) & _MASK)))
    i = torch.arange(start, start + n, dtype=torch.int64, device=device) if index is None else index
    z = mix(i * _s64(_GOLD) + base)
    return lsr(z, 11).to(torch.float64) * 2.0 ** -53


# ---- the benchmark configurations ---------------------------------------------------------------
def box(seed, n, ra0, ra_span, dec0, dec_span, xp="numpy", device=None, start=0):
    if xp == "numpy":
        return ra0 + ra_span * uniform(seed, 0, n, start), dec0 + dec_span * uniform(seed, 1, n, start)
    return (ra0 + ra_span * uniform_torch(seed, 0, n, device, start),
            dec0 + dec_span * uniform_torch(seed, 1, n, device, start))


def c1(seed, n=100_000, **kw):
    """C1: uniform over RA[0,1) x Dec[0,1) (reference CPU-runnable case)."""
    return box(seed, n, 0.0, 1.0, 0.0, 1.0, **kw)


def c2(seed, n=10_000_000, **kw):
    """C2 (headline): uniform over RA[0,10) x Dec[-5,5)."""
    return box(seed, n, 0.0, 10.0, -5.0, 10.0, **kw)


def c3_clustered(seed, levels, xp="numpy", device=None):
    """C3: Soneira-Peebles-like hierarchy, 10 children per level, N = 10**levels rows inside
    RA[0,10] x Dec[-5,5] (modelled on randpos_power_base, astro.hpp:856-887)."""
    n = 10 ** levels
    if xp == "numpy":
        ra = np.full(n, 5.0)
        dec = np.zeros(n)
        idx = np.arange(n, dtype=np.uint64)
        for lvl in range(levels):
            pre = idx // np.uint64(10 ** (levels - 1 - lvl))
            r = (2.5 / 2 ** lvl) * np.sqrt(uniform(seed, 2 * lvl, 0, index=pre))
            th = 2.0 * np.pi * uniform(seed, 2 * lvl + 1, 0, index=pre)
            ra += r * np.cos(th)
            dec += r * np.sin(th)
        return ra, dec
    import torch
    ra = torch.full((n,), 5.0, dtype=torch.float64, device=device)
    dec = torch.zeros(n, dtype=torch.float64, device=device)
    idx = torch.arange(n, dtype=torch.int64, device=device)
    for lvl in range(levels):
        pre = idx // (10 ** (levels - 1 - lvl))
        r = (2.5 / 2 ** lvl) * torch.sqrt(uniform_torch(seed, 2 * lvl, 0, device, index=pre))
        th = 2.0 * np.pi * uniform_torch(seed, 2 * lvl + 1, 0, device, index=pre)
        ra += r * torch.cos(th)
        dec += r * torch.sin(th)
    return ra, dec


def fullsky(seed, n, xp="numpy", device=None, start=0):
    """C4 / C5: uniform on the sphere, ra = 360 u, dec = asin(2 v - 1) in degrees."""
    if xp == "numpy":
        return (360.0 * uniform(seed, 0, n, start),
                np.degrees(np.arcsin(2.0 * uniform(seed, 1, n, start) - 1.0)))
    import torch
    return (360.0 * uniform_torch(seed, 0, n, device, start),
            torch.rad2deg(torch.asin(2.0 * uniform_torch(seed, 1, n, device, start) - 1.0)))
