"""Known-answer vectors of the rANS bitstream format (csrc/rans.cu), written by the CPU restatement
oracle/rans_ref.py:  python tests/golden/make_golden_rans.py  ->  tests/golden/rans_kat.npz

The reference has no entropy coder, so these vectors pin the builder-defined format against
accidental change: the CPU oracle and the CUDA coder must both reproduce them byte for byte."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import rans_ref as rr  # noqa: E402


def case(seed, n, S, quant, spread=3.0, log_sigma_std=1.0, escapes=0):
    rng = np.random.default_rng(seed)
    mu = (rng.standard_normal(n) * spread).astype(np.float32)
    sigma = np.exp(rng.standard_normal(n) * log_sigma_std - 0.5).astype(np.float32)
    v = (mu + sigma * rng.standard_normal(n)).astype(np.float32)
    if escapes:
        idx = rng.choice(n, escapes, replace=False)
        v[idx] += rng.choice([-1.0, 1.0], escapes).astype(np.float32) * rng.integers(50, 100000, escapes).astype(np.float32)
    if quant == 2:
        k = np.rint((v - mu).astype(np.float32)).astype(np.int64)
        blob = rr.encode_segment(k, np.zeros_like(mu), sigma, S, quant)
    else:
        k = np.rint(v).astype(np.int64)
        blob = rr.encode_segment(k, mu, sigma, S, quant)
    return dict(v=v, mu=mu, sigma=sigma, k=k, S=np.int64(S), quant=np.int64(quant), blob=np.frombuffer(blob, dtype=np.uint8))


if __name__ == "__main__":
    cases = [case(1, 1000, 1, 1), case(2, 4099, 7, 1, escapes=5), case(3, 3000, 64, 2, spread=0.7),
             case(4, 37, 64, 1), case(5, 6000, 3, 1, log_sigma_std=2.5, escapes=3)]
    flat = {}
    for i, c in enumerate(cases):
        for k, v in c.items():
            flat[f"c{i}_{k}"] = v
    flat["cases"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rans_kat.npz"), **flat)
    print({f"c{i}": len(c["blob"]) for i, c in enumerate(cases)})
