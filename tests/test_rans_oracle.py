"""CPU tests of the rANS restatement (oracle/rans_ref.py) that the CUDA coder is compared with on the GPU:
known-answer vectors, exact round trips incl. escapes and degenerate parameters, coded size against the estimated rate,
and the format's normal-CDF table (library copy == committed header == recomputed from math.erfc)."""
import os
import re

import numpy as np
import pytest

from oracle import rans_ref as rr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _kat():
    d = np.load(os.path.join(GOLDEN, "rans_kat.npz"))
    return [{k: d[f"c{i}_{k}"] for k in ("v", "mu", "sigma", "k", "S", "quant", "blob")} for i in range(int(d["cases"]))]


def test_known_answer_vectors():
    for c in _kat():
        S, quant = int(c["S"]), int(c["quant"])
        mu = c["mu"] if quant == 1 else np.zeros_like(c["mu"])
        blob = rr.encode_segment(c["k"], mu, c["sigma"], S, quant)
        assert blob == c["blob"].tobytes()
        assert np.array_equal(rr.decode_segment(blob, mu, c["sigma"], S, quant), c["k"])


def test_phi_table_three_copies_agree():
    import ldic_b200
    T = rr.phi_table()
    assert T[0] == 0 and T[-1] == 1 << 24 and np.all(np.diff(T) >= 0) and T[1024] == 1 << 23
    lib_T = np.array(ldic_b200.ops.rans_phi_table(), dtype=np.int64)          # host function, no GPU involved
    assert np.array_equal(T, lib_T)
    hdr = open(os.path.join(os.path.dirname(GOLDEN), "..", "learning-driven-image-compression-algorithm_b200", "csrc",
                            "rans_phi_table.h")).read()
    vals = np.array([int(x) for x in re.findall(r"(\d+)u", hdr)], dtype=np.int64)
    assert np.array_equal(T, vals)


def test_integer_model_is_a_valid_distribution():
    rng = np.random.default_rng(0)
    mu = np.concatenate([rng.standard_normal(200) * 50, [0.5, -0.5, 1e9, -1e9, np.nan, 0.0]]).astype(np.float32)
    sigma = np.concatenate([np.exp(rng.standard_normal(200) * 3), [0.0, -1.0, np.nan, 1e-30, 1e30, 200.0]]).astype(np.float32)
    mu_s, sg_s, m, R = rr.make_model(mu, sigma)
    assert np.all(R >= 15) and np.all(R <= 1023) and np.all(np.isfinite(mu_s)) and np.all(sg_s > 0)
    for i in range(mu.size):
        j = np.arange(0, 2 * R[i] + 2)
        c = rr.cdf_at(mu_s[i:i + 1], sg_s[i:i + 1], m[i:i + 1], R[i:i + 1], j)
        assert c[0] == 0 and c[-1] == 65536 and np.all(np.diff(c) >= 1), (mu[i], sigma[i])


@pytest.mark.parametrize("n,S", [(0, 1), (1, 1), (1, 5), (31, 4), (1000, 1), (1000, 33), (5000, 5000)])
def test_round_trip_sizes_and_stream_counts(n, S):
    rng = np.random.default_rng(n * 31 + S)
    mu = (rng.standard_normal(n) * 4).astype(np.float32)
    sigma = np.exp(rng.standard_normal(n) * 1.5).astype(np.float32)
    k = np.rint(mu + sigma * rng.standard_normal(n)).astype(np.int64)
    blob = rr.encode_segment(k, mu, sigma, S)
    assert np.array_equal(rr.decode_segment(blob, mu, sigma, S), k)
    states_off, counts_off, esc_off = rr.layout(S)
    assert len(blob) >= esc_off and len(blob) <= esc_off + 10 * n


def test_round_trip_escapes_and_degenerate_parameters():
    rng = np.random.default_rng(7)
    n = 3000
    mu = (rng.standard_normal(n) * 2).astype(np.float32)
    sigma = np.exp(rng.standard_normal(n)).astype(np.float32)
    k = np.rint(mu + sigma * rng.standard_normal(n)).astype(np.int64)
    k[::97] += 12345                      # far outside every window
    k[5::131] -= 1 << 29
    sigma[3::211] = 0.0                   # sanitised to 1e-6
    sigma[4::223] = np.nan
    mu[6::199] = np.nan
    mu[8::251] = 3e9
    sigma[9::241] = 1e20                  # window capped at R = 1023
    blob = rr.encode_segment(k, mu, sigma, 16)
    assert np.array_equal(rr.decode_segment(blob, mu, sigma, 16), k)
    E = np.frombuffer(blob, dtype="<u4", count=8)[3]
    assert E >= len(k[::97])


def test_coded_size_tracks_the_estimated_rate():
    """8 * bytes vs sum(-log2 L) with L = Phi((k-mu+.5)/s) - Phi((k-mu-.5)/s) (GaussianModel, model/net.py:272-286)."""
    from math import erf, sqrt
    rng = np.random.default_rng(11)
    n, S = 40000, 20
    mu = (rng.standard_normal(n) * 3).astype(np.float32)
    sigma = np.exp(rng.standard_normal(n) * 0.8 + 0.3).astype(np.float32)
    k = np.rint(mu + sigma * rng.standard_normal(n)).astype(np.int64)
    Phi = np.vectorize(lambda t: 0.5 * (1.0 + erf(t / sqrt(2.0))))
    L = Phi((k - mu.astype(np.float64) + 0.5) / sigma) - Phi((k - mu.astype(np.float64) - 0.5) / sigma)
    est = float(-np.log2(np.maximum(L, 1e-8)).sum())
    blob = rr.encode_segment(k, mu, sigma, S)
    E = int(np.frombuffer(blob, dtype="<u4", count=8)[3])
    overhead = 8 * (rr.HEADER + 6 * S + 8 * E)
    ideal = rr.ideal_bits(k, mu, sigma)
    assert abs(ideal - est) < 0.005 * est                      # the 16-bit integer model costs < 0.5 % over the true Gaussian
    assert 8 * len(blob) - overhead < ideal + 32 * S + 64      # rANS itself: below one state flush per stream over ideal
    assert 8 * len(blob) < 1.01 * est + overhead


def test_corrupt_streams_are_rejected_or_differ():
    rng = np.random.default_rng(3)
    n, S = 2000, 8
    mu = (rng.standard_normal(n) * 2).astype(np.float32)
    sigma = np.exp(rng.standard_normal(n) * 0.5).astype(np.float32)
    k = np.rint(mu + sigma * rng.standard_normal(n)).astype(np.int64)
    blob = bytearray(rr.encode_segment(k, mu, sigma, S))
    with pytest.raises(ValueError):
        rr.decode_segment(bytes(blob[:-2]), mu, sigma, S)                    # truncated
    bad = bytearray(blob); bad[0] ^= 1
    with pytest.raises(ValueError):
        rr.decode_segment(bytes(bad), mu, sigma, S)                          # magic
    bad = bytearray(blob); bad[-1] ^= 0x10
    try:
        out = rr.decode_segment(bytes(bad), mu, sigma, S)
        assert not np.array_equal(out, k)
    except ValueError:
        pass


def test_container_round_trip_and_rejection():
    import ldic_b200
    ev = ldic_b200.evaluation
    s = {"z": b"\x01\x02\x03", "y": b"", "syntax": bytes(range(40))}
    blob = ev.pack_container(100, 70, 128, 128, s)
    assert ev.unpack_container(blob) == (100, 70, 128, 128, s)
    for bad in (blob[:-1], b"XDIC" + blob[4:], blob[:20]):
        with pytest.raises(ValueError):
            ev.unpack_container(bad)


def test_wavefront_schedule_and_row_aligned_streams():
    """Host logic of Net.decompress: every pixel is decoded exactly once, after the pixels its context reads
    (model/net.py:219-242: rows r-3..r-1 at columns c-2..c+1, and (r, c-2), (r, c-1)) and after its predecessor in its
    streams; content streams never cross a row of the latent nor a column group."""
    import ldic_b200
    for high in (False, True):
        net = ldic_b200.Net((1, 64, 64, 3), (1, 64, 64, 3), high, False)
        Cc, G = net.N - net.M, net.y_groups()
        cg = Cc // G
        assert G == 4 and cg * G == Cc
        for h, w in [(4, 4), (8, 12), (32, 48), (16, 16), (5, 7)]:
            S = net.y_streams(h, w)
            assert S % (h * G) == 0 and (h * w * Cc) % S == 0
            run = h * w * Cc // S                               # symbols per stream
            assert (w * cg) % run == 0 and run <= 65535         # a whole number of streams per (row, column group)
            table = net._wavefront_table(h, w, "cpu").numpy()
            T = w + 2 * (h - 1)
            assert table.shape == (T, h, G, 2)
            when = {}
            for t in range(T):
                for r in range(h):
                    if table[t, r, 0, 1] == 0:
                        assert not table[t, r, :, 1].any()
                        continue
                    pix = set()
                    for g in range(G):
                        first, count = table[t, r, g]
                        assert count == cg and (first - g * h * w * cg) % cg == 0
                        p = (first - g * h * w * cg) // cg
                        assert 0 <= p < h * w and first // run == (first + count - 1) // run      # inside one stream
                        pix.add(int(p))
                    assert len(pix) == 1
                    p = pix.pop()
                    assert p // w == r and (r, p % w) not in when
                    when[(r, p % w)] = t
            assert len(when) == h * w
            for (r, c), t in when.items():
                deps = [(r + i - 3, c + j - 2) for i in range(4) for j in range(4) if not (i == 3 and j >= 2)]
                for rr, cc in deps:
                    if 0 <= rr < h and 0 <= cc < w:
                        assert when[(rr, cc)] < t, ((r, c), (rr, cc))
                if c > 0:                                       # stream predecessor: the pixel to the left (same row)
                    assert when[(r, c - 1)] == t - 1


def test_column_groups_reorder_only():
    """col_groups changes the ORDER in which a segment is coded (and one header word), nothing else."""
    rng = np.random.default_rng(5)
    rows, cols, G, S = 6, 8, 4, 6
    mu = (rng.standard_normal(rows * cols) * 2).astype(np.float32)
    sigma = np.exp(rng.standard_normal(rows * cols) * 0.5).astype(np.float32)
    k = np.rint(mu + sigma * rng.standard_normal(rows * cols)).astype(np.int64)
    order = rr.group_order(rows, cols, G)
    assert sorted(order.tolist()) == list(range(rows * cols)) and order[:3].tolist() == [0, 1, 8]
    blob = rr.encode_segment(k[order], mu[order], sigma[order], S, groups=G)
    back = np.empty_like(k)
    back[order] = rr.decode_segment(blob, mu[order], sigma[order], S, groups=G)
    assert np.array_equal(back, k)
    plain = rr.encode_segment(k[order], mu[order], sigma[order], S)
    assert blob[:24] == plain[:24] and blob[24:28] == (4).to_bytes(4, "little") and blob[28:] == plain[28:]
    with pytest.raises(ValueError):
        rr.decode_segment(blob, mu[order], sigma[order], S)
