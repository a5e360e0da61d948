"""Test infrastructure only: numpy restatement of the trit-plane extension.

PARITY UNPINNED: the reference's model/Trit_Plane.py:25-57 crashes and contains neither trit planes nor a
likelihood, so there are no reference vectors for this function.  It follows the definition in include/ldic.h
(ldic_tritplane_likelihood) in float64; the quantiser is the reference's round-half-to-even (np.rint = torch.round,
model/net.py:419) applied to float32 v - mu.
"""
from __future__ import annotations

import math

import numpy as np


def _phi_mass(a, b, s):
    erfc = np.vectorize(math.erfc, otypes=[np.float64])
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    flip = (a + b) < 0                       # mirror to the upper tail (same as the kernel): erfc stays precise
    a, b = np.where(flip, -b, a), np.where(flip, -a, b)
    k = 1.0 / (s * math.sqrt(2.0))
    return np.maximum(0.5 * (erfc(a * k) - erfc(b * k)), 1e-30)


def tritplane(v, sigma, mu=None, planes=4, scale_bound=0.11, lik_bound=1e-9):
    v = np.asarray(v, np.float32)
    mu = np.zeros_like(v) if mu is None else np.asarray(mu, np.float32)
    s = np.maximum(np.asarray(sigma, np.float32), np.float32(scale_bound)).astype(np.float64)
    L = int(planes)
    H = (3 ** L - 1) // 2
    q = np.clip(np.rint((v - mu).astype(np.float32)), -H, H).astype(np.int64)
    u = q + H
    trits = np.zeros((L,) + v.shape, np.int8)
    sums = np.zeros(L, np.float64)
    lo = np.zeros(v.shape, np.int64)
    for l in range(L - 1, -1, -1):
        w = 3 ** l
        t = (u // w) % 3
        trits[l] = t
        a = (lo - H).astype(np.float64) - 0.5
        parent = _phi_mass(a, a + 3.0 * w, s)
        child = _phi_mass(a + t * w, a + (t + 1) * w, s)
        sums[l] = np.log(np.maximum(child / parent, lik_bound)).sum()
        lo = lo + t * w
    return trits, q.astype(np.int32), sums
