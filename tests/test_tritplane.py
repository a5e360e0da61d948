"""a12 trit-plane extension (builder-defined, parity unpinned: see oracle/tritplane_ref.py)."""
import numpy as np
import pytest
import torch

from oracle import tritplane_ref as tr


def _inputs(n, seed):
    g = np.random.default_rng(seed)
    v = (g.standard_normal(n) * 4).astype(np.float32)
    mu = g.standard_normal(n).astype(np.float32)
    sg = np.clip(np.exp(g.standard_normal(n)), 0.05, 20).astype(np.float32)
    edge = np.array([0.5, -0.5, 1.5, -1.5, 2.5, -0.0, 40.0, -40.0, 39.5, 13.0, -13.0], np.float32)   # ties, clamp, range ends
    v[:edge.size] = edge
    mu[:edge.size] = 0
    sg[:4] = [0.0, 0.11, -1.0, 0.05]
    return v, mu, sg


@pytest.mark.parametrize("L", [1, 3, 4])
def test_oracle_properties(L):
    """Exact reconstruction of the symbols from the planes, symbols = round-half-to-even, telescoping product."""
    v, mu, sg = _inputs(4096, 1)
    t, q, sums = tr.tritplane(v, sg, mu, planes=L)
    H = (3 ** L - 1) // 2
    rec = sum(t[l].astype(np.int64) * 3 ** l for l in range(L)) - H
    assert np.array_equal(rec, q)
    assert np.array_equal(q, np.clip(np.rint(v - mu), -H, H).astype(np.int32))
    assert t.min() >= 0 and t.max() <= 2
    # prod_l L_l = P(q) / P(|q| <= H)   (away from the likelihood floor)
    s = np.maximum(sg, 0.11).astype(np.float64)
    pq = tr._phi_mass(q - 0.5, q + 0.5, s) / tr._phi_mass(np.full(q.shape, -H - 0.5), np.full(q.shape, H + 0.5), s)
    ok = pq > 1e-6
    tt, _, per = tr.tritplane(v[ok], sg[ok], mu[ok], planes=L, lik_bound=0.0)
    assert abs(per.sum() - np.log(pq[ok]).sum()) < 1e-6 * abs(np.log(pq[ok]).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("n,L", [(1, 4), (1000, 3), (1 << 20, 4), (777777, 5)])
def test_kernel_vs_oracle(n, L):
    import ldic_b200
    v, mu, sg = _inputs(max(n, 16), 2)
    v, mu, sg = v[:n], mu[:n], sg[:n]
    planes, q, sums = ldic_b200.ops.tritplane_likelihood(torch.from_numpy(v).cuda(), torch.from_numpy(sg).cuda(),
                                                         torch.from_numpy(mu).cuda(), planes=L)
    t_ref, q_ref, s_ref = tr.tritplane(v, sg, mu, planes=L)
    assert torch.equal(q.cpu(), torch.from_numpy(q_ref))                 # bit-exact symbols
    assert torch.equal(planes.cpu(), torch.from_numpy(t_ref))            # bit-exact planes
    H = (3 ** L - 1) // 2
    rec = sum(planes[l].long() * 3 ** l for l in range(L)) - H
    assert torch.equal(rec.int(), q)
    np.testing.assert_allclose(sums.cpu().numpy(), s_ref, rtol=2e-4, atol=1e-3)


@pytest.mark.gpu
def test_kernel_full_size_properties():
    """BASELINE configs[4] shape (2 images per GPU: 2 x 192 x 128 x 128 latents): exact reconstruction and the
    telescoping identity against the a8 likelihood kernel."""
    import ldic_b200
    from ldic_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    shape = (2, 192, 128, 128)
    mu = torch.randn(shape, device="cuda", generator=g)
    sg = torch.exp(torch.randn(shape, device="cuda", generator=g)).clamp_(0.5, 20)
    v = mu + sg * torch.randn(shape, device="cuda", generator=g).clamp_(-4, 4)     # every P(q) stays far above the 1e-9 floors
    L = 5
    planes, q, sums = ops.tritplane_likelihood(v, sg, mu, planes=L)
    H = (3 ** L - 1) // 2
    rec = sum(planes[l].int() * 3 ** l for l in range(L)) - H
    assert torch.equal(rec, q)
    assert torch.equal(q, torch.round(v - mu).clamp(-H, H).int())
    # sum over planes = sum ln P(q) - sum ln P(|q| <= H); with H = 121 and sigma <= 20 the second term is ~0
    _, lik, s_all = ops.gaussian_likelihood(v, sg, mu, quant=ops.QUANT_DEQUANT, form=ops.FORM_GAUSSIAN_CONDITIONAL)
    assert abs(sums.sum().item() - s_all.item()) < 2e-3 * abs(s_all.item())
