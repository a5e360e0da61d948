"""CPU restatement (numpy) of the rANS coder of csrc/rans.cu -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product path
(ldic_b200) never does.

PARITY UNPINNED against the reference: xiaobucc/learning-driven-image-compression-algorithm holds no entropy coder or
bitstream at all ("absent in reference"); it only estimates the rate from the likelihoods
(model/net.py:856-861).  This file therefore restates the builder-defined format of csrc/rans.cu (header comment there)
independently -- the normal-CDF table is recomputed from math.erfc instead of being read from the library -- and the
tests require byte-for-byte equal streams, exact round trips, and a coded size that tracks the reference's estimated
rate sum(-log2 L) (GaussianModel, model/net.py:272-286).

Per segment, symbols k[i] (integers), parameters mu[i], sigma[i] (float32, already broadcast; for quant 2 the caller
passes mu = 0), S streams: stream s codes the run of symbols [s Ls, (s+1) Ls), Ls = ceil(n / S)."""
from __future__ import annotations

import math
import struct

import numpy as np

MAGIC = 0x3141524C
HEADER = 32
PHI_N = 2048
f32 = np.float32


def phi_table() -> np.ndarray:
    """T[i] = round(Phi(-8 + i/128) * 2^24), i = 0..2048."""
    return np.array([int(round(0.5 * math.erfc(-(-8.0 + i / 128.0) / math.sqrt(2.0)) * (1 << 24))) for i in range(PHI_N + 1)],
                    dtype=np.int64)


_T = phi_table()


def make_model(mu: np.ndarray, sigma: np.ndarray):
    """Sanitised (mu, sigma), window centre m = rint(mu) and half width R = min(1023, max(15, 2 + ceil(6 sigma)))."""
    mu = np.asarray(mu, dtype=f32).copy()
    sigma = np.asarray(sigma, dtype=f32).copy()
    with np.errstate(invalid="ignore"):
        bad = ~(np.abs(mu) <= f32(2097152.0))
        mu[bad] = np.where(mu[bad] > 0, f32(2097152.0), np.where(mu[bad] < 0, f32(-2097152.0), f32(0.0)))
        sigma[~(sigma >= f32(1e-6))] = f32(1e-6)
    sigma[sigma > f32(1e6)] = f32(1e6)
    m = np.rint(mu).astype(np.int64)
    r = np.ceil(f32(6.0) * sigma)
    R = np.where(r >= 1021.0, 1023, np.maximum(15, 2 + r.astype(np.int64))).astype(np.int64)
    return mu, sigma, m, R


def phi24(t: np.ndarray) -> np.ndarray:
    t = np.asarray(t, dtype=f32)
    with np.errstate(over="ignore", invalid="ignore"):
        tq = (t * f32(128.0) + f32(1024.0)).astype(f32)
    tq = np.minimum(np.maximum(tq, f32(0.0)), f32(2048.0))
    i = np.minimum(tq.astype(np.int64), PHI_N - 1)
    f = ((tq - i.astype(f32)).astype(f32) * f32(4096.0)).astype(np.int64)
    a, b = _T[i], _T[i + 1]
    return a + (((b - a) * f) >> 12)


def cdf_at(mu, sigma, m, R, j) -> np.ndarray:
    """C(j) of the integer model, j in [0, Nsym] (arrays)."""
    nsym = 2 * R + 1
    j = np.asarray(j, dtype=np.int64)
    k = (m - R + j).astype(f32)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        inv = (f32(1.0) / sigma).astype(f32)
        t = (((k - f32(0.5)).astype(f32) - mu).astype(f32) * inv).astype(f32)
    c = ((phi24(t) * (65536 - nsym)) >> 24) + j
    return np.where(j <= 0, 0, np.where(j >= nsym, 65536, c)).astype(np.int64)


def layout(S: int):
    states_off = HEADER
    counts_off = states_off + 4 * S
    esc_off = counts_off + ((2 * S + 3) & ~3)
    return states_off, counts_off, esc_off


def run_length(n: int, S: int) -> int:
    return (n + S - 1) // S if S else 0


def stream_counts(n: int, S: int) -> np.ndarray:
    s = np.arange(S, dtype=np.int64)
    Ls = run_length(n, S)
    return np.clip(n - s * Ls, 0, Ls)


def group_order(rows: int, cols: int, groups: int) -> np.ndarray:
    """Coding order of a [rows, cols] segment with `groups` column groups: indices into the row-major flattening, all
    rows' columns [0, cols/G) first, then the next cols/G, ... (csrc/rans.cu: LdicRansArgs.col_groups)."""
    idx = np.arange(rows * cols, dtype=np.int64).reshape(rows, groups, cols // groups)
    return idx.transpose(1, 0, 2).reshape(-1)


def encode_segment(k, mu, sigma, S: int, quant: int = 1, groups: int = 0) -> bytes:
    """k, mu, sigma in CODING order (see group_order); `groups` is only recorded in the header (0: row-major)."""
    k = np.asarray(k, dtype=np.int64).ravel()
    n = k.size
    mu, sigma, m, R = make_model(np.asarray(mu, dtype=f32).ravel(), np.asarray(sigma, dtype=f32).ravel())
    nsym = 2 * R + 1
    j = np.clip(k - (m - R), 0, nsym - 1)
    esc = (j == 0) | (j == nsym - 1)
    start = cdf_at(mu, sigma, m, R, j)
    freq = cdf_at(mu, sigma, m, R, j + 1) - start
    assert np.all(freq >= 1) and np.all(freq < 65536)
    cnt = stream_counts(n, S)
    Ls = run_length(n, S)
    maxc = int(cnt.max()) if S else 0
    x = np.full(S, 1 << 16, dtype=np.int64)
    words = np.zeros((S, max(maxc, 1)), dtype=np.int64)          # in emission order
    nw = np.zeros(S, dtype=np.int64)
    sidx = np.arange(S, dtype=np.int64)
    for p in range(maxc - 1, -1, -1):
        act = sidx[cnt > p]
        i = act * Ls + p
        fr, stt = freq[i], start[i]
        xa = x[act]
        emit = xa >= (fr << 16)
        es = act[emit]
        words[es, nw[es]] = xa[emit] & 0xFFFF
        nw[es] += 1
        xa = np.where(emit, xa >> 16, xa)
        x[act] = ((xa // fr) << 16) + (xa % fr) + stt
    states_off, counts_off, esc_off = layout(S)
    E = int(esc.sum())
    W = int(nw.sum())
    out = bytearray(esc_off + 8 * E + 2 * W)
    struct.pack_into("<8I", out, 0, MAGIC, n, S, E, W, quant, groups if groups > 1 else 0, 0)
    out[states_off:states_off + 4 * S] = x.astype("<u4").tobytes()
    out[counts_off:counts_off + 2 * S] = nw.astype("<u2").tobytes()
    ei = np.nonzero(esc)[0]
    pairs = np.empty((E, 2), dtype="<u4")
    pairs[:, 0] = ei
    pairs[:, 1] = k[ei].astype(np.int32).view(np.uint32) if E else 0
    out[esc_off:esc_off + 8 * E] = pairs.tobytes()
    allw = np.concatenate([words[s, :nw[s]][::-1] for s in range(S)]) if W else np.zeros(0, dtype=np.int64)
    out[esc_off + 8 * E:] = allw.astype("<u2").tobytes()
    return bytes(out)


def decode_segment(buf: bytes, mu, sigma, S: int, quant: int = 1, groups: int = 0) -> np.ndarray:
    magic, n, S_h, E, W, q_h, g_h, _ = struct.unpack_from("<8I", buf, 0)
    if magic != MAGIC or S_h != S or q_h != quant or g_h != (groups if groups > 1 else 0):
        raise ValueError("bad header")
    mu, sigma, m, R = make_model(np.asarray(mu, dtype=f32).ravel(), np.asarray(sigma, dtype=f32).ravel())
    if mu.size != n:
        raise ValueError("symbol count mismatch")
    states_off, counts_off, esc_off = layout(S)
    if len(buf) != esc_off + 8 * E + 2 * W:
        raise ValueError("size mismatch")
    x = np.frombuffer(buf, dtype="<u4", count=S, offset=states_off).astype(np.int64)
    nw = np.frombuffer(buf, dtype="<u2", count=S, offset=counts_off).astype(np.int64)
    pairs = np.frombuffer(buf, dtype="<u4", count=2 * E, offset=esc_off).reshape(E, 2)
    words = np.frombuffer(buf, dtype="<u2", count=W, offset=esc_off + 8 * E).astype(np.int64)
    if int(nw.sum()) != W:
        raise ValueError("word counts do not add up")
    wpos = np.concatenate([[0], np.cumsum(nw)[:-1]]).astype(np.int64) if S else np.zeros(0, dtype=np.int64)
    wend = wpos + nw
    cnt = stream_counts(n, S)
    Ls = run_length(n, S)
    out = np.zeros(n, dtype=np.int64)
    sidx = np.arange(S, dtype=np.int64)
    nsym = 2 * R + 1
    for p in range(int(cnt.max()) if S else 0):
        act = sidx[cnt > p]
        i = act * Ls + p
        xa = x[act]
        slot = xa & 0xFFFF
        mi, si, mm, RR = mu[i], sigma[i], m[i], R[i]
        lo = np.zeros(act.size, dtype=np.int64)
        hi = nsym[i] - 1
        c_lo = np.zeros(act.size, dtype=np.int64)
        c_hi = np.full(act.size, 65536, dtype=np.int64)
        while np.any(lo < hi):
            mid = (lo + hi + 1) >> 1
            cm = cdf_at(mi, si, mm, RR, mid)
            go = (cm <= slot) & (lo < hi)
            stay = (~(cm <= slot)) & (lo < hi)
            lo = np.where(go, mid, lo)
            c_lo = np.where(go, cm, c_lo)
            hi = np.where(stay, mid - 1, hi)
            c_hi = np.where(stay, cm, c_hi)
        fr = c_hi - c_lo
        xa = fr * (xa >> 16) + slot - c_lo
        need = xa < (1 << 16)
        ns = act[need]
        if np.any(wpos[ns] >= wend[ns]):
            raise ValueError("stream ran out of words")
        xa[need] = (xa[need] << 16) | words[wpos[ns]]
        wpos[ns] += 1
        x[act] = xa
        out[i] = mm - RR + lo
    if np.any(x != (1 << 16)) or np.any(wpos != wend):
        raise ValueError("corrupt stream")
    if E:
        if np.any(pairs[:, 0] >= n):
            raise ValueError("escape index")
        out[pairs[:, 0].astype(np.int64)] = pairs[:, 1].copy().view(np.int32).astype(np.int64)
    return out


def ideal_bits(k, mu, sigma) -> float:
    """sum(-log2 P_int(k)) of the integer model: what an ideal coder would spend on the in-band symbols."""
    k = np.asarray(k, dtype=np.int64).ravel()
    mu, sigma, m, R = make_model(np.asarray(mu, dtype=f32).ravel(), np.asarray(sigma, dtype=f32).ravel())
    nsym = 2 * R + 1
    j = np.clip(k - (m - R), 0, nsym - 1)
    freq = cdf_at(mu, sigma, m, R, j + 1) - cdf_at(mu, sigma, m, R, j)
    return float(np.sum(16.0 - np.log2(freq.astype(np.float64))))
