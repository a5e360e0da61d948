"""Thin torch-facing wrappers over the C ABI (include/ldic.h).

torch is used for device memory, streams and autograd plumbing only; every op
launches hand-written sm_100a kernels from libldic_b200.so through ctypes with
raw device pointers and the current CUDA stream.  No op has a CPU or
torch-eager fallback: a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ConvDesc, LikelihoodArgs, RansArgs, LdicError, check

_ws_cache = {}
# bench.py sets this to a list to get (layer, (B,H,W), start_event, end_event) per conv launch
PROFILE = None


def _L():
    return _lib.load()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype=None, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise LdicError(f"{name} must be a CUDA tensor (ldic_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise LdicError(f"{name} must be {dtype}, got {t.dtype}")
    if t.device.index != torch.cuda.current_device():
        # kernels launch on the CURRENT device's stream: a tensor of another GPU would be dereferenced on the wrong
        # device.  Net / the module classes enter torch.cuda.device(x.device) themselves; raw op callers must too.
        raise LdicError(f"{name} lives on {t.device} but the current device is cuda:{torch.cuda.current_device()} "
                        "(wrap the call in torch.cuda.device(tensor.device))")
    return t


def set_tuning(key: str, value: int) -> int:
    """ldic_set_tuning: change one of the load-time tuning switches (tests / A-B runs); returns the previous value."""
    r = _L().ldic_set_tuning(key.encode(), int(value))
    if r < 0:
        check(r, "ldic_set_tuning")
    return r


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def launch_count() -> int:
    return int(_L().ldic_launch_count())


def _workspace(device) -> torch.Tensor:
    # one reduction workspace per (device, stream): launches on different streams never share a ticket counter
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    ws = _ws_cache.get(key)
    if ws is None:
        n = int(_L().ldic_likelihood_workspace_bytes())
        ws = torch.zeros((n + 7) // 8, dtype=torch.int64, device=device)
        _ws_cache[key] = ws
    return ws


# --------------------------------------------------------------------------------------
# a3
# --------------------------------------------------------------------------------------
def lower_bound_fwd(x: torch.Tensor, bound: float) -> torch.Tensor:
    x = _req(x, torch.float32, "x").contiguous()
    y = torch.empty_like(x)
    check(_L().ldic_lower_bound(_ptr(x), float(bound), _ptr(y), x.numel(), _stream()), "ldic_lower_bound")
    return y


def lower_bound_bwd(x: torch.Tensor, bound: float, grad_out: torch.Tensor) -> torch.Tensor:
    x = _req(x, torch.float32, "x").contiguous()
    g = _req(grad_out, torch.float32, "grad").contiguous()
    gi = torch.empty_like(g)
    check(_L().ldic_lower_bound_bwd(_ptr(x), float(bound), _ptr(g), _ptr(gi), x.numel(), _stream()), "ldic_lower_bound_bwd")
    return gi


class _LowerBoundFn(torch.autograd.Function):
    """ops/bound_ops.py:30-41 / model/gdn.py:11-26 with both passes on our kernels."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x)
        ctx.bound = float(bound)
        return lower_bound_fwd(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        (x,) = ctx.saved_tensors
        return lower_bound_bwd(x, ctx.bound, grad_output), None


def lower_bound(x: torch.Tensor, bound: float) -> torch.Tensor:
    return _LowerBoundFn.apply(x, float(bound))


def nonneg_reparam(p: torch.Tensor, bound: float, pedestal: float) -> torch.Tensor:
    p = _req(p, torch.float32, "p").contiguous()
    out = torch.empty_like(p)
    check(_L().ldic_nonneg_reparam(_ptr(p), float(bound), float(pedestal), _ptr(out), p.numel(), _stream()),
          "ldic_nonneg_reparam")
    return out


def gdn_prepare(beta_p: torch.Tensor, gamma_p: torch.Tensor, beta_bound: float, gamma_bound: float, pedestal: float,
                tc_groups: int = 0, tc_np: int = 0):
    """-> (beta_eff[C], gamma_eff[C,C]) and, when tc_np>0, (gamma_bf16[Np,Np], beta_tiled[Np])."""
    beta_p = _req(beta_p, torch.float32, "beta").contiguous()
    gamma_p = _req(gamma_p, torch.float32, "gamma").contiguous()
    Cc = beta_p.numel()
    if gamma_p.numel() != Cc * Cc:
        raise LdicError("gamma must be CxC")
    dev = beta_p.device
    beta_eff = torch.empty(Cc, dtype=torch.float32, device=dev)
    gamma_eff = torch.empty(Cc, Cc, dtype=torch.float32, device=dev)
    g16 = bt = None
    if tc_np:
        g16 = torch.empty(tc_np, tc_np, dtype=torch.bfloat16, device=dev)
        bt = torch.empty(tc_np, dtype=torch.float32, device=dev)
    check(_L().ldic_gdn_prepare(_ptr(beta_p), _ptr(gamma_p), Cc, float(beta_bound), float(gamma_bound), float(pedestal),
                                _ptr(beta_eff), _ptr(gamma_eff), _ptr(g16), _ptr(bt), int(tc_groups or 1), int(tc_np),
                                int(tc_np), _stream()), "ldic_gdn_prepare")
    return beta_eff, gamma_eff, g16, bt


def gdn_nchw(x: torch.Tensor, beta_eff: torch.Tensor, gamma_eff: torch.Tensor, inverse: bool, use_rsqrt: bool) -> torch.Tensor:
    x = _req(x, torch.float32, "x").contiguous()
    if x.dim() != 4:
        raise LdicError("GDN expects (B,C,H,W)")
    B, Cc, H, W = x.shape
    y = torch.empty_like(x)
    check(_L().ldic_gdn_nchw_f32(_ptr(x), _ptr(beta_eff), _ptr(gamma_eff), _ptr(y), B, Cc, H, W, int(bool(inverse)),
                                 int(bool(use_rsqrt)), _stream()), "ldic_gdn_nchw_f32")
    return y


# --------------------------------------------------------------------------------------
# a6-a9
# --------------------------------------------------------------------------------------
QUANT_NONE, QUANT_ROUND, QUANT_DEQUANT, QUANT_STE = 0, 1, 2, 3
FORM_GAUSSIAN_MODEL, FORM_GAUSSIAN_CONDITIONAL = 0, 1


def likelihood_rows(v, rows, cols, *, v_rs, v_off=0, mu=None, mu_mode=0, mu_rs=0, mu_off=0,
                    sigma=None, sigma_mode=2, sigma_rs=0, sigma_off=0, sigma_period=1,
                    quant=QUANT_NONE, form=FORM_GAUSSIAN_MODEL, sigma_is_log=False,
                    lik_bound=1e-8, scale_bound=0.11,
                    v_hat=None, v_hat_rs=0, v_hat_off=0, v_hat_bf16=None, vb_rs=0, vb_off=0,
                    lik=None, sum_out=None) -> torch.Tensor:
    """Raw strided form of ldic_round_likelihood_bpp; returns the 1-element sum(ln L) tensor."""
    _req(v, torch.float32, "v")
    _req(sigma, torch.float32, "sigma")
    if sum_out is None:
        sum_out = torch.empty(1, dtype=torch.float32, device=v.device)
    a = LikelihoodArgs()
    a.v, a.v_rs, a.v_off = _ptr(v), v_rs, v_off
    a.mu, a.mu_rs, a.mu_off, a.mu_mode = _ptr(mu), mu_rs, mu_off, mu_mode
    a.sigma, a.sigma_rs, a.sigma_off, a.sigma_mode, a.sigma_period = _ptr(sigma), sigma_rs, sigma_off, sigma_mode, sigma_period
    a.rows, a.cols = rows, cols
    a.quant, a.form, a.sigma_is_log = quant, form, int(bool(sigma_is_log))
    a.lik_bound, a.scale_bound = float(lik_bound), float(scale_bound)
    a.v_hat, a.v_hat_rs, a.v_hat_off = _ptr(v_hat), v_hat_rs, v_hat_off
    a.v_hat_bf16, a.vb_rs, a.vb_off = _ptr(v_hat_bf16), vb_rs, vb_off
    a.lik = _ptr(lik)
    a.sum_ln_out = _ptr(sum_out)
    a.workspace = _ptr(_workspace(v.device))
    check(_L().ldic_round_likelihood_bpp(C.byref(a), _stream()), "ldic_round_likelihood_bpp")
    return sum_out


def gaussian_likelihood(v: torch.Tensor, sigma: torch.Tensor, mu: Optional[torch.Tensor] = None, *,
                        quant: int = QUANT_NONE, form: int = FORM_GAUSSIAN_MODEL, lik_bound: float = 1e-8,
                        scale_bound: float = 0.11, want_lik: bool = True, want_vhat: bool = False,
                        sum_out: Optional[torch.Tensor] = None
                        ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], torch.Tensor]:
    """Module-surface form on (B,C,H,W) tensors with torch broadcasting rules limited to the
    reference's two cases: same-shape sigma/mu, or (1,C,1,1) sigma with mu == 0 / (1,C,1,1).
    Returns (v_hat or None, likelihood or None, sum_ln[1])."""
    _req(v, torch.float32, "v")
    if v.dim() != 4:
        raise LdicError("expected (B,C,H,W)")
    B, Cc, H, W = v.shape
    per_channel = sigma.numel() == Cc and sigma.numel() != v.numel()
    if per_channel:
        if mu is not None and torch.count_nonzero(mu).item() != 0:
            # general per-channel mean: fold it in by expanding (rare; not on the reference path)
            return gaussian_likelihood(v, sigma.expand_as(v).contiguous(), mu.expand_as(v).contiguous(), quant=quant,
                                       form=form, lik_bound=lik_bound, scale_bound=scale_bound, want_lik=want_lik,
                                       want_vhat=want_vhat, sum_out=sum_out)
        vc = v.contiguous()
        sg = sigma.reshape(Cc).contiguous()
        lik = torch.empty_like(vc) if want_lik else None
        vh = torch.empty_like(vc) if want_vhat else None
        s = likelihood_rows(vc, B * Cc, H * W, v_rs=H * W, sigma=sg, sigma_mode=3, sigma_period=Cc, quant=quant,
                            form=form, lik_bound=lik_bound, scale_bound=scale_bound, v_hat=vh, v_hat_rs=H * W, lik=lik,
                            sum_out=sum_out)
        return vh, lik, s
    if sigma.shape != v.shape or (mu is not None and mu.shape != v.shape):
        sigma = sigma.expand_as(v)
        mu = None if mu is None else mu.expand_as(v)
    # the reference hands us NCHW *views* of (b,h,w,c) storage (model/net.py:314-317): keep whatever
    # memory format v has and bring the others to it.
    if v.is_contiguous():
        fmt = torch.contiguous_format
    elif v.is_contiguous(memory_format=torch.channels_last):
        fmt = torch.channels_last
    else:
        v = v.contiguous()
        fmt = torch.contiguous_format
    sg = sigma.contiguous(memory_format=fmt)
    m = None if mu is None else mu.contiguous(memory_format=fmt)
    n = v.numel()
    lik = torch.empty_like(v) if want_lik else None      # preserves v's memory format
    vh = torch.empty_like(v) if want_vhat else None
    s = likelihood_rows(v, 1, n, v_rs=n, mu=m, mu_mode=2 if m is not None else 0, mu_rs=n, sigma=sg, sigma_mode=2,
                        sigma_rs=n, quant=quant, form=form, lik_bound=lik_bound, scale_bound=scale_bound, v_hat=vh,
                        v_hat_rs=n, lik=lik, sum_out=sum_out)
    return vh, lik, s


# --------------------------------------------------------------------------------------
# f4: rANS entropy coder (builder-defined extension: the reference only estimates the rate, see include/ldic.h)
# --------------------------------------------------------------------------------------
RANS_STATUS = {1: "a symbol was NaN or beyond 2^30", 2: "output capacity too small", 4: "bad header", 8: "corrupt stream",
               16: "incremental decode out of order"}


def rans_streams_for(seg_elems: int, symbols_per_stream: int = 2048) -> int:
    """Number of interleaved rANS states per segment: about `symbols_per_stream` symbols each (6 header bytes per
    stream; fewer symbols per stream = more GPU parallelism, more header)."""
    return max(1, -(-int(seg_elems) // max(1, min(int(symbols_per_stream), 65535))))


def _rans_status_check(status: torch.Tensor, what: str):
    st = status.cpu().tolist()
    bad = [(i, v) for i, v in enumerate(st) if v]
    if bad:
        i, v = bad[0]
        why = ", ".join(t for b, t in RANS_STATUS.items() if v & b)
        raise LdicError(f"{what}: segment {i}: {why}")


class RansStreams:
    """Device-resident result of rans_encode_rows: `buf` uint8 [segments, stride], `sizes` int32 [segments]."""

    def __init__(self, buf, sizes, status, streams, seg_elems, quant):
        self.buf, self.sizes, self.status = buf, sizes, status
        self.streams, self.seg_elems, self.quant = streams, seg_elems, quant

    def check(self):
        _rans_status_check(self.status, "rans encode")
        return self

    def nbytes(self):
        """Bytes per segment (synchronises)."""
        self.check()
        return self.sizes.cpu().tolist()

    def tobytes(self):
        """One bytes object per segment (synchronises)."""
        sizes = self.nbytes()
        m = max(sizes) if sizes else 0
        host = self.buf[:, :m].cpu().numpy()
        return [host[i, :n].tobytes() for i, n in enumerate(sizes)]


_pinned_pool = {}


def _pinned_bytes(slot: int, rows: int, cols: int) -> torch.Tensor:
    """(rows, cols) uint8 view of a grow-only pinned staging buffer (cudaHostAlloc per call would stall the device);
    the caller copies the bytes out before the next call with the same slot."""
    need = max(rows * cols, 1)
    buf = _pinned_pool.get(slot)
    if buf is None or buf.numel() < need:
        buf = torch.empty(1 << max(need - 1, 1).bit_length(), dtype=torch.uint8, pin_memory=True)
        _pinned_pool[slot] = buf
    return buf[:rows * cols].view(rows, cols)


def rans_tobytes(streams, copy_stream: Optional["torch.cuda.Stream"] = None) -> list:
    """[RansStreams, ...] -> [[bytes per segment], ...] with two synchronisations in total: one device-to-host copy of
    all sizes and status words, then the coded bytes of every stream (only the bytes that were written).
    With `copy_stream` the copies run there (after everything enqueued so far on the current stream) and only that
    stream is synchronised, so kernels launched afterwards on the current stream overlap with the readback."""
    streams = list(streams)
    if not streams:
        return []
    if copy_stream is not None:
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            return rans_tobytes(streams)
    # copy engines only (no gather / cat kernels: a kernel on a side stream waits behind the persistent conv CTAs that
    # hold every SM, 2 ms per step measured): contiguous device -> pinned copies, one per tensor / bitstream row
    nseg = [e.sizes.numel() for e in streams]
    meta_h = _pinned_bytes(-1, 1, 8 * sum(nseg)).reshape(-1).view(torch.int32)
    pos = 0
    for e, k in zip(streams, nseg):
        meta_h[pos:pos + k].copy_(e.sizes, non_blocking=True)
        meta_h[pos + k:pos + 2 * k].copy_(e.status, non_blocking=True)
        pos += 2 * k
    torch.cuda.current_stream().synchronize()
    meta = meta_h.tolist()
    sizes, pos = [], 0
    for e, k in zip(streams, nseg):
        sz, st = meta[pos:pos + k], meta[pos + k:pos + 2 * k]
        pos += 2 * k
        bad = [(i, v) for i, v in enumerate(st) if v]
        if bad:
            i, v = bad[0]
            raise LdicError(f"rans encode: segment {i}: " + ", ".join(t for b, t in RANS_STATUS.items() if v & b))
        sizes.append(sz)
    hosts = []
    for slot, (e, sz) in enumerate(zip(streams, sizes)):
        m = max(sz) if sz else 0
        h = _pinned_bytes(slot, len(sz), m)
        for i, nb in enumerate(sz):
            if nb:
                h[i, :nb].copy_(e.buf[i, :nb], non_blocking=True)
        hosts.append(h)
    torch.cuda.current_stream().synchronize()
    return [[h[i, :n].numpy().tobytes() for i, n in enumerate(sz)] for h, sz in zip(hosts, sizes)]


def _rans_args(rows, cols, rows_per_segment, v, v_rs, v_off, mu, mu_mode, mu_rs, mu_off, sigma, sigma_mode, sigma_rs, sigma_off,
               sigma_period, quant, sigma_is_log, scale_bound, streams, col_groups=1):
    a = RansArgs()
    a.col_groups = int(col_groups)
    a.v, a.v_rs, a.v_off = _ptr(v), v_rs, v_off
    a.mu, a.mu_rs, a.mu_off, a.mu_mode = _ptr(mu), mu_rs, mu_off, mu_mode
    a.sigma, a.sigma_rs, a.sigma_off, a.sigma_mode, a.sigma_period = _ptr(sigma), sigma_rs, sigma_off, sigma_mode, sigma_period
    a.rows, a.cols, a.rows_per_segment = rows, cols, rows_per_segment
    a.quant, a.sigma_is_log, a.scale_bound, a.streams = quant, int(bool(sigma_is_log)), float(scale_bound), streams
    return a


def _rans_ws(device, segs, seg_elems, streams):
    n = int(_L().ldic_rans_workspace_bytes(segs, seg_elems, streams))
    return torch.empty(max(n, 256), dtype=torch.uint8, device=device)     # torch's allocator hands out >= 512-byte alignment


def rans_encode_rows(v, rows, cols, rows_per_segment, *, v_rs, v_off=0, mu=None, mu_mode=0, mu_rs=0, mu_off=0,
                     sigma=None, sigma_mode=2, sigma_rs=0, sigma_off=0, sigma_period=1, quant=QUANT_ROUND,
                     sigma_is_log=False, scale_bound=0.0, streams: Optional[int] = None,
                     capacity: Optional[int] = None, col_groups: int = 1) -> RansStreams:
    """Raw strided form of ldic_rans_encode (same addressing as likelihood_rows): one bitstream per `rows_per_segment`
    rows.  Stream-ordered, no synchronisation; read the result with RansStreams.tobytes() / .nbytes()."""
    _req(v, torch.float32, "v")
    _req(sigma, torch.float32, "sigma")
    if rows_per_segment <= 0 or rows % rows_per_segment:
        raise LdicError("rows must be a multiple of rows_per_segment")
    segs, seg_elems = rows // rows_per_segment, rows_per_segment * cols
    S = int(streams) if streams else rans_streams_for(seg_elems)
    stride = int(capacity) if capacity else int(_L().ldic_rans_max_bytes(seg_elems, S))
    stride = (stride + 3) & ~3
    buf = torch.empty((max(segs, 1), stride), dtype=torch.uint8, device=v.device)
    sizes = torch.zeros(max(segs, 1), dtype=torch.int32, device=v.device)
    status = torch.zeros(max(segs, 1), dtype=torch.int32, device=v.device)
    if segs == 0:
        return RansStreams(buf[:0], sizes[:0], status[:0], S, seg_elems, quant)
    a = _rans_args(rows, cols, rows_per_segment, v, v_rs, v_off, mu, mu_mode, mu_rs, mu_off, sigma, sigma_mode, sigma_rs,
                   sigma_off, sigma_period, quant, sigma_is_log, scale_bound, S, col_groups)
    ws = _rans_ws(v.device, segs, seg_elems, S)
    check(_L().ldic_rans_encode(C.byref(a), _ptr(buf), stride, _ptr(sizes), _ptr(status), _ptr(ws), _stream()), "ldic_rans_encode")
    return RansStreams(buf[:segs], sizes[:segs], status[:segs], S, seg_elems, quant)


def rans_decode_rows(data, rows, cols, rows_per_segment, v_hat, *, v_hat_rs, v_hat_off=0, mu=None, mu_mode=0, mu_rs=0, mu_off=0,
                     sigma=None, sigma_mode=2, sigma_rs=0, sigma_off=0, sigma_period=1, quant=QUANT_ROUND,
                     sigma_is_log=False, scale_bound=0.0, streams: Optional[int] = None, check_status: bool = True,
                     col_groups: int = 1) -> torch.Tensor:
    """ldic_rans_decode: `data` is a RansStreams or a list of bytes objects (one per segment); the symbols land in
    v_hat[row * v_hat_rs + v_hat_off + col].  Returns the per-segment status tensor (raises on a bad stream unless
    check_status=False)."""
    _req(v_hat, torch.float32, "v_hat")
    _req(sigma, torch.float32, "sigma")
    segs, seg_elems = rows // rows_per_segment, rows_per_segment * cols
    if isinstance(data, RansStreams):
        buf, sizes, S = data.buf, data.sizes, data.streams
        if not buf.is_contiguous():
            buf = buf.contiguous()
    else:
        if len(data) != segs:
            raise LdicError(f"expected {segs} bitstreams, got {len(data)}")
        S = int(streams) if streams else rans_streams_for(seg_elems)
        stride = (max([len(b) for b in data] + [32]) + 3) & ~3
        host = torch.zeros((segs, stride), dtype=torch.uint8)
        for i, b in enumerate(data):
            if len(b):
                host[i, :len(b)] = torch.frombuffer(bytearray(b), dtype=torch.uint8)
        buf = host.to(v_hat.device)
        sizes = torch.tensor([len(b) for b in data], dtype=torch.int32, device=v_hat.device)
    status = torch.zeros(max(segs, 1), dtype=torch.int32, device=v_hat.device)
    if segs == 0:
        return status[:0]
    a = _rans_args(rows, cols, rows_per_segment, None, 0, 0, mu, mu_mode, mu_rs, mu_off, sigma, sigma_mode, sigma_rs, sigma_off,
                   sigma_period, quant, sigma_is_log, scale_bound, S, col_groups)
    ws = _rans_ws(v_hat.device, segs, seg_elems, S)
    check(_L().ldic_rans_decode(C.byref(a), _ptr(buf), buf.stride(0) if segs else 0, _ptr(sizes), _ptr(v_hat), v_hat_rs, v_hat_off,
                                _ptr(status), _ptr(ws), _stream()), "ldic_rans_decode")
    if check_status:
        _rans_status_check(status[:segs], "rans decode")
    return status[:segs]


class RansDecoder:
    """Incremental decoding (ldic_rans_decode_begin / _ranges): for decoders whose (mu, sigma) depend on symbols decoded
    earlier.  `dec = RansDecoder(data, rows, cols, rows_per_segment, streams=S, quant=...)`, then per step
    `dec.decode(ranges, nranges, v_hat, v_hat_rs=..., mu=..., sigma=..., ...)` with `ranges` an int32 device tensor of
    (first symbol, count) pairs (segment-relative); every range must continue the stream it lies in.  `dec.finish()`
    synchronises and raises on a corrupt or mis-ordered decode."""

    def __init__(self, data, rows, cols, rows_per_segment, *, streams: int, quant: int = QUANT_ROUND, device=None,
                 col_groups: int = 1):
        segs, seg_elems = rows // rows_per_segment, rows_per_segment * cols
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        if isinstance(data, RansStreams):
            buf, sizes = data.buf, data.sizes
            if not buf.is_contiguous():
                buf = buf.contiguous()
        else:
            if len(data) != segs:
                raise LdicError(f"expected {segs} bitstreams, got {len(data)}")
            stride = (max([len(b) for b in data] + [32]) + 3) & ~3
            host = torch.zeros((segs, stride), dtype=torch.uint8)
            for i, b in enumerate(data):
                if len(b):
                    host[i, :len(b)] = torch.frombuffer(bytearray(b), dtype=torch.uint8)
            buf = host.to(dev)
            sizes = torch.tensor([len(b) for b in data], dtype=torch.int32, device=dev)
        self.buf, self.sizes, self.S, self.quant, self.col_groups = buf, sizes, int(streams), quant, int(col_groups)
        self.rows, self.cols, self.rps, self.segs = rows, cols, rows_per_segment, segs
        self.status = torch.zeros(max(segs, 1), dtype=torch.int32, device=dev)
        self.state = torch.empty((max(segs, 1), self.S, 4), dtype=torch.int32, device=dev)
        ones = torch.ones(1, dtype=torch.float32, device=dev)             # begin only needs the geometry of the problem
        a = _rans_args(rows, cols, rows_per_segment, None, 0, 0, None, 0, 0, 0, ones, 3, 0, 0, 1, quant, False, 0.0, self.S,
                       self.col_groups)
        ws = _rans_ws(dev, segs, seg_elems, self.S)
        if segs:
            check(_L().ldic_rans_decode_begin(C.byref(a), _ptr(buf), buf.stride(0), _ptr(sizes), _ptr(self.state),
                                              _ptr(self.status), _ptr(ws), _stream()), "ldic_rans_decode_begin")

    def decode(self, ranges: torch.Tensor, nranges: int, v_hat: torch.Tensor, *, v_hat_rs, v_hat_off=0, v_hat_bf16=None,
               vb_rs=0, vb_off=0, param_row_map=None, bf16_row_map=None, mu=None, mu_mode=0, mu_rs=0, mu_off=0, sigma=None, sigma_mode=2, sigma_rs=0, sigma_off=0,
               sigma_period=1, sigma_is_log=False, scale_bound=0.0):
        _req(v_hat, torch.float32, "v_hat")
        _req(sigma, torch.float32, "sigma")
        _req(ranges, torch.int32, "ranges")
        if v_hat_bf16 is not None:
            _req(v_hat_bf16, torch.bfloat16, "v_hat_bf16")
        a = _rans_args(self.rows, self.cols, self.rps, None, 0, 0, mu, mu_mode, mu_rs, mu_off, sigma, sigma_mode, sigma_rs,
                       sigma_off, sigma_period, self.quant, sigma_is_log, scale_bound, self.S, self.col_groups)
        if self.segs and nranges:
            check(_L().ldic_rans_decode_ranges(C.byref(a), _ptr(self.buf), self.buf.stride(0), _ptr(self.state), _ptr(ranges),
                                               int(nranges), _ptr(v_hat), v_hat_rs, v_hat_off, _ptr(v_hat_bf16), vb_rs, vb_off,
                                               _ptr(param_row_map), _ptr(bf16_row_map), _ptr(self.status), _stream()),
                  "ldic_rans_decode_ranges")

    def finish(self):
        _rans_status_check(self.status[:self.segs], "rans incremental decode")
        st = self.state[:self.segs].cpu()
        return st


def _rans_surface(v_shape, sigma, mu):
    """Addressing of the module-surface forms: (B,C,H,W) contiguous, sigma / mu of the same shape or per channel."""
    B, Cc, H, W = v_shape
    per_channel = sigma.numel() == Cc and sigma.numel() != B * Cc * H * W
    if per_channel:
        kw = dict(sigma=sigma.reshape(Cc).contiguous(), sigma_mode=3, sigma_period=Cc)
        if mu is not None:
            kw.update(mu=mu.reshape(Cc).contiguous(), mu_mode=3)
        return B * Cc, H * W, Cc, kw
    n = Cc * H * W
    kw = dict(sigma=sigma.expand(v_shape).contiguous(), sigma_mode=2, sigma_rs=n)
    if mu is not None:
        kw.update(mu=mu.expand(v_shape).contiguous(), mu_mode=2, mu_rs=n)
    return B, n, 1, kw


def rans_encode(v: torch.Tensor, sigma: torch.Tensor, mu: Optional[torch.Tensor] = None, *, quant: int = QUANT_ROUND,
                scale_bound: float = 0.0, streams: Optional[int] = None) -> RansStreams:
    """Entropy-codes a (B,C,H,W) latent, one bitstream per image, with the Gaussian parameters the likelihood ops take
    (same-shape sigma / mu, or (1,C,1,1) per-channel ones)."""
    v = _req(v, torch.float32, "v").contiguous()
    if v.dim() != 4:
        raise LdicError("expected (B,C,H,W)")
    rows, cols, rps, kw = _rans_surface(tuple(v.shape), sigma, mu)
    return rans_encode_rows(v, rows, cols, rps, v_rs=cols, quant=quant, scale_bound=scale_bound, streams=streams, **kw)


def rans_decode(data, shape, sigma: torch.Tensor, mu: Optional[torch.Tensor] = None, *, quant: int = QUANT_ROUND,
                scale_bound: float = 0.0, streams: Optional[int] = None) -> torch.Tensor:
    """Inverse of rans_encode: the (B,C,H,W) fp32 tensor of symbols (quant 1) or symbols + mu (quant 2)."""
    shape = tuple(int(d) for d in shape)
    rows, cols, rps, kw = _rans_surface(shape, sigma, mu)
    out = torch.empty(shape, dtype=torch.float32, device=sigma.device)
    rans_decode_rows(data, rows, cols, rps, out, v_hat_rs=cols, quant=quant, scale_bound=scale_bound, streams=streams, **kw)
    return out


def rans_phi_table():
    """The 2049-entry 24-bit normal-CDF table of the bitstream format (host list)."""
    n = C.c_int(0)
    p = _L().ldic_rans_phi_table(C.byref(n))
    return [int(p[i]) for i in range(n.value)]


# --------------------------------------------------------------------------------------
# a12 (builder-defined extension, no reference oracle: see include/ldic.h)
# --------------------------------------------------------------------------------------
_tp_ws_cache = {}


def tritplane_likelihood(v: torch.Tensor, sigma: torch.Tensor, mu: Optional[torch.Tensor] = None, planes: int = 4, *,
                         scale_bound: float = 0.11, lik_bound: float = 1e-9, want_planes: bool = True, want_q: bool = True):
    """Trit planes (int8 [L, *v.shape]), symbols q = clamp(round(v - mu)) (int32) and sum(ln L_l) per plane (float [L])."""
    v = _req(v, torch.float32, "v").contiguous()
    sigma = _req(sigma, torch.float32, "sigma").contiguous()
    if mu is not None:
        mu = _req(mu, torch.float32, "mu").contiguous()
    if sigma.shape != v.shape or (mu is not None and mu.shape != v.shape):
        raise LdicError("tritplane: v, mu, sigma must have the same shape")
    n = v.numel()
    key = (v.device.index if v.device.index is not None else torch.cuda.current_device(), _stream())
    ws = _tp_ws_cache.get(key)
    if ws is None:
        ws = torch.zeros((int(_L().ldic_tritplane_workspace_bytes()) + 7) // 8, dtype=torch.int64, device=v.device)
        _tp_ws_cache[key] = ws
    pl = torch.empty((planes,) + tuple(v.shape), dtype=torch.int8, device=v.device) if want_planes else None
    q = torch.empty(v.shape, dtype=torch.int32, device=v.device) if want_q else None
    sums = torch.empty(planes, dtype=torch.float32, device=v.device)
    check(_L().ldic_tritplane_likelihood(_ptr(v), _ptr(mu), _ptr(sigma), n, int(planes), float(scale_bound), float(lik_bound),
                                         _ptr(pl), _ptr(q), _ptr(sums), _ptr(ws), _stream()), "ldic_tritplane_likelihood")
    return pl, q, sums


# --------------------------------------------------------------------------------------
# a11
# --------------------------------------------------------------------------------------
def mse_sum(x: torch.Tensor, x_tilde: torch.Tensor, clamp_pm1: bool = False) -> torch.Tensor:
    """Exact per-image sum of squared 8-bit-level errors (int64[B])."""
    x = _req(x, torch.float32, "x").contiguous()
    xt = _req(x_tilde, torch.float32, "x_tilde").contiguous()
    if x.shape != xt.shape:
        raise LdicError("x and x_tilde must have the same shape")
    B = x.shape[0] if x.dim() > 0 else 1
    chw = x.numel() // max(B, 1)
    out = torch.zeros(B, dtype=torch.int64, device=x.device)
    check(_L().ldic_mse_sum(_ptr(x), _ptr(xt), B, chw, int(bool(clamp_pm1)), _ptr(out), _stream()), "ldic_mse_sum")
    return out


def rd_pack_metrics(bits3: torch.Tensor, sq_err: torch.Tensor, chw: int, want_v_mse: bool = True):
    """[sum ln L z, y, syntax] (float[3]) + exact squared-error sums (int64[B]) -> (packed5 double[5], v_mse float[B])."""
    bits3 = _req(bits3, torch.float32, "bits")
    sq_err = _req(sq_err, torch.int64, "sq_err")
    if bits3.numel() != 3 or not bits3.is_contiguous() or not sq_err.is_contiguous():
        raise LdicError("rd_pack_metrics: bits must be 3 contiguous floats, sq_err contiguous int64")
    B = sq_err.numel()
    packed = torch.empty(5, dtype=torch.float64, device=bits3.device)
    v_mse = torch.empty(B, dtype=torch.float32, device=bits3.device) if want_v_mse else None
    check(_L().ldic_rd_pack_metrics(_ptr(bits3), _ptr(sq_err), B, int(chw), _ptr(packed), _ptr(v_mse), _stream()),
          "ldic_rd_pack_metrics")
    return packed, v_mse


def rd_finish_metrics(packed5: torch.Tensor, pixels_per_image: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """packed5 (possibly all-reduced) -> float[2] = (bpp, v_psnr)."""
    packed5 = _req(packed5, torch.float64, "packed5")
    if packed5.numel() != 5 or not packed5.is_contiguous():
        raise LdicError("rd_finish_metrics: packed5 must be 5 contiguous doubles")
    if out is None:
        out = torch.empty(2, dtype=torch.float32, device=packed5.device)
    check(_L().ldic_rd_finish_metrics(_ptr(packed5), float(pixels_per_image), _ptr(out), _stream()), "ldic_rd_finish_metrics")
    return out


def syntax_conv_mse(x_nchw, xt_nhwc, w, want_x_tilde=False, tanh_out=False):
    x = _req(x_nchw, torch.float32, "x").contiguous()
    xt = _req(xt_nhwc, torch.float32, "x_tilde16").contiguous()
    w = _req(w, torch.float32, "w").contiguous()
    B, _, H, W = x.shape
    M = xt.shape[-1]
    out = torch.zeros(B, dtype=torch.int64, device=x.device)
    xo = torch.empty_like(x) if want_x_tilde else None
    check(_L().ldic_syntax_conv_mse(_ptr(x), _ptr(xt), _ptr(w), B, M, H, W, int(bool(tanh_out)), _ptr(xo), _ptr(out), _stream()),
          "ldic_syntax_conv_mse")
    return out, xo


# --------------------------------------------------------------------------------------
# glue
# --------------------------------------------------------------------------------------
def nchw_to_nhwc_bf16(x: torch.Tensor, Cp: Optional[int] = None, apply_abs: bool = False) -> torch.Tensor:
    x = _req(x, torch.float32, "x").contiguous()
    B, Cc, H, W = x.shape
    Cp = Cp or Cc
    y = torch.empty(B, H, W, Cp, dtype=torch.bfloat16, device=x.device)
    check(_L().ldic_nchw_f32_to_nhwc_bf16(_ptr(x), _ptr(y), B, Cc, H, W, Cp, int(apply_abs), _stream()), "nchw_to_nhwc")
    return y


def nhwc_to_nchw_f32(x: torch.Tensor, Cc: Optional[int] = None) -> torch.Tensor:
    if x.dtype not in (torch.bfloat16, torch.float32):
        raise LdicError("nhwc_to_nchw: bf16 or fp32 input")
    _req(x, None, "x")
    x = x.contiguous()
    B, H, W, Cp = x.shape
    Cc = Cc or Cp
    y = torch.empty(B, Cc, H, W, dtype=torch.float32, device=x.device)
    check(_L().ldic_nhwc_to_nchw_f32(_ptr(x), int(x.dtype == torch.bfloat16), _ptr(y), B, Cc, H, W, Cp, _stream()),
          "nhwc_to_nchw")
    return y


def u8_to_f32_pm1(x: torch.Tensor) -> torch.Tensor:
    """uint8 levels -> fp32 (u/255)*2-1, the reference's input map (ToTensor + eval_net.py:84), bit-exact."""
    x = _req(x, torch.uint8, "x").contiguous()
    y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    check(_L().ldic_u8_to_f32_pm1(_ptr(x), _ptr(y), x.numel(), _stream()), "ldic_u8_to_f32_pm1")
    return y


def latent_prep(y: torch.Tensor, want_round_bf16=True, want_abs_bf16=True, want_round_f32=False):
    y = _req(y, torch.float32, "y").contiguous()
    mk = lambda dt: torch.empty(y.shape, dtype=dt, device=y.device)
    yr = mk(torch.bfloat16) if want_round_bf16 else None
    ya = mk(torch.bfloat16) if want_abs_bf16 else None
    yf = mk(torch.float32) if want_round_f32 else None
    check(_L().ldic_latent_prep(_ptr(y), y.numel(), _ptr(yr), _ptr(ya), _ptr(yf), _stream()), "ldic_latent_prep")
    return yr, ya, yf


def im2col_5x5s2(x: torch.Tensor, Kp: int = 128, out_f32: bool = False) -> torch.Tensor:
    x = _req(x, torch.float32, "x").contiguous()
    B, Cin, H, W = x.shape
    a = torch.empty(B, H // 2, W // 2, Kp, dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
    fn = _L().ldic_im2col_5x5s2_f32 if out_f32 else _L().ldic_im2col_5x5s2
    check(fn(_ptr(x), _ptr(a), B, Cin, H, W, Kp, _stream()), "ldic_im2col_5x5s2")
    return a


# --------------------------------------------------------------------------------------
# a1 / a4 / a5 / a10: tensor-core conv layers
# --------------------------------------------------------------------------------------
def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


class ConvTC:
    """One conv / transposed-conv layer of the transforms, weights pre-packed for the
    tcgen05 kernel.  Input and output are NHWC tensors (bf16 in; bf16 or fp32 out)."""

    def __init__(self, kind: int, weight: torch.Tensor, bias: Optional[torch.Tensor], *, act: int = _lib.ACT_NONE,
                 out_f32: bool = False, cin_pad: Optional[int] = None, cin_offset: int = 0,
                 cout_pad: Optional[int] = None, gdn: Optional[tuple] = None, aux=(0, 0), precision: str = "bf16"):
        if precision not in ("bf16", "tf32", "tf32_last"):
            raise LdicError("ConvTC precision must be 'bf16', 'tf32' or 'tf32_last'")
        # TF32 parity mode: fp32 NHWC activations, kind::tf32 MMAs (~half rate).  'tf32' rounds the outputs to tf32 (they
        # feed another tf32 layer), 'tf32_last' leaves them fp32 (they feed the quantiser)
        self.tf32 = 1 if precision == "tf32" else (2 if precision == "tf32_last" else 0)
        if self.tf32:
            out_f32 = True
        w = _req(weight, torch.float32, "weight").contiguous()
        transposed = kind in (_lib.LDIC_DECONV_GS_5x5, _lib.LDIC_DECONV_HS_5x5, _lib.LDIC_DECONV_S1_3x3,
                              _lib.LDIC_DECONV_GS_5x5_MERGED)
        self.aux = tuple(aux)
        probe_hw = (16, 16)
        if kind == _lib.LDIC_CONV_1x1:
            w = w.reshape(w.shape[0], -1).contiguous()
            cout, cin = w.shape
        elif kind == _lib.LDIC_CONV_FIRST_5x5S2:     # (Cout, 3, 5, 5); input is the NCHW fp32 image
            cout, cin = w.shape[0], w.shape[1]
            cin_pad = 128
        elif kind == _lib.LDIC_CTX_CONV1:            # (N, 2N-M, 3, 3); aux = (N, M)
            cout, cin = w.shape[0], w.shape[1]
            cin_pad = 2 * self.aux[0]
        elif kind in (_lib.LDIC_CTX_CONV2, _lib.LDIC_CTX_CONV3):
            cout, cin = w.shape[0], w.shape[1]
            probe_hw = (4, 4) if kind == _lib.LDIC_CTX_CONV2 else (2, 2)
        elif kind == _lib.LDIC_CTX_FC:               # Linear(4N, 2*Cout): rows [mu | log sigma], cols (c,h,w)
            cin = w.shape[1] // 4
            cout = w.shape[0] // 2
            cout_pad = cout_pad or _pad64(cout)
            probe_hw = (2, 2)
        elif transposed:
            cin, cout = w.shape[0], w.shape[1]
        else:
            cout, cin = w.shape[0], w.shape[1]
        self.kind, self.act, self.out_f32 = kind, act, bool(out_f32)
        self.cin, self.cout = cin, cout
        self.cin_pad = cin_pad or _pad64(cin + cin_offset)
        self.cout_pad = cout_pad or cout
        self.device = w.device
        self._probe_hw = probe_hw
        d = self._desc(1, *probe_hw)
        L = _L()
        self.np_cols = L.ldic_conv_n_cols(C.byref(d))
        n = L.ldic_conv_weight_elems(C.byref(d))
        nb = L.ldic_conv_bias_elems(C.byref(d))
        if n < 0 or self.np_cols < 0 or nb < 0:
            check(-1, "ldic_conv_weight_elems")
        self.w_packed = torch.empty(n, dtype=torch.float32 if self.tf32 else torch.bfloat16, device=w.device)
        self.bias_packed = torch.empty(nb, dtype=torch.float32, device=w.device)
        b = None if bias is None else _req(bias, torch.float32, "bias").contiguous()
        check(L.ldic_conv_pack_weights(C.byref(d), _ptr(w), _ptr(b), int(cin_offset), _ptr(self.w_packed),
                                       _ptr(self.bias_packed), _stream()), "ldic_conv_pack_weights")
        self.gamma_bf16 = self.beta_tiled = None
        if act in (_lib.ACT_GDN, _lib.ACT_IGDN):
            if gdn is None:
                raise LdicError("GDN epilogue needs (beta_p, gamma_p, beta_bound, gamma_bound, pedestal)")
            beta_p, gamma_p, bb, gb, ped = gdn
            groups = self.np_cols // self.cout_pad
            if self.tf32:        # fp32 gamma_eff [C, C] rounded to tf32 (the tensor core would truncate), beta_eff as is
                if groups != 1 or self.np_cols != beta_p.numel():
                    raise LdicError("TF32 mode: GDN over exactly the layer's output channels")
                be, ge, _, _ = gdn_prepare(beta_p, gamma_p, bb, gb, ped)
                gi = ge.contiguous().view(torch.int32)
                self.gamma_bf16 = ((gi + 0x1000) & ~0x1FFF).view(torch.float32).contiguous()
                self.beta_tiled = be.contiguous()
            else:
                _, _, self.gamma_bf16, self.beta_tiled = gdn_prepare(beta_p, gamma_p, bb, gb, ped, tc_groups=groups,
                                                                     tc_np=self.np_cols)

    def _desc(self, B, H, W, sm_limit: int = 0) -> ConvDesc:
        return ConvDesc(self.kind, B, H, W, self.cin, self.cout, self.cin_pad, self.cout_pad, self.act,
                        int(self.out_f32), int(self.aux[0]), int(self.aux[1]), int(sm_limit), int(self.tf32))

    def out_dims(self, B, H, W):
        d = self._desc(B, H, W)
        dims = (C.c_int * 4)()
        check(_L().ldic_conv_out_dims(C.byref(d), dims), "ldic_conv_out_dims")
        return tuple(dims)

    def out_hw(self, H, W):
        return self.out_dims(1, H, W)[1:3]

    def flops(self, B: int, H: int, W: int) -> float:
        """Algorithmic FLOPs (2 x MACs on logical channels, no zero-insertion / padding waste):
        conv: k*k*Cin*Cout per OUTPUT pixel; transposed conv: k*k*Cin*Cout per INPUT pixel;
        GDN/IGDN: C*C per output pixel; context layers: the taps that fall inside the patch."""
        K = _lib
        if self.kind == K.LDIC_CTX_CONV1:      # 100 in-patch taps per position, 10 of them without the masked y part
            N, M = self.aux[0], self.aux[1] & 0xff
            return 2.0 * B * H * W * self.cout * (100 * self.cin - 10 * (N - M))
        if self.kind == K.LDIC_CTX_CONV2:
            return 2.0 * B * 25 * self.cin * self.cout
        if self.kind == K.LDIC_CTX_CONV3:
            return 2.0 * B * 16 * self.cin * self.cout
        if self.kind == K.LDIC_CTX_FC:
            return 2.0 * B * 4 * self.cin * 2 * self.cout
        ho, wo = self.out_hw(H, W)
        k = {K.LDIC_CONV_1x1: 1, K.LDIC_CONV_S1_3x3_P1: 3, K.LDIC_DECONV_S1_3x3: 3}.get(self.kind, 5)
        transposed = self.kind in (K.LDIC_DECONV_GS_5x5, K.LDIC_DECONV_HS_5x5, K.LDIC_DECONV_S1_3x3,
                                   K.LDIC_DECONV_GS_5x5_MERGED)
        pix = B * (H * W if transposed else ho * wo)
        f = 2.0 * k * k * self.cin * self.cout * pix
        if self.act in (K.ACT_GDN, K.ACT_IGDN):
            f += 2.0 * self.cout * self.cout * B * ho * wo
        return f

    def fused_tail(self, x: torch.Tensor, image: torch.Tensor, conv_w: torch.Tensor, *, want_x_tilde: bool = False,
                   want_out: bool = False, tanh_out: bool = False):
        """Merged last synthesis deconv + batch_conv + squared level error in one kernel
        (ldic_conv_forward_fused_tail).  x: NHWC bf16 layer input; image: NCHW fp32 (B,3,2H,2W); conv_w: (B,3,M).
        Returns (sq_err int64[B], x_tilde NCHW or None, layer output NHWC fp32 or None)."""
        _req(x, torch.bfloat16, "x")
        if image.dtype not in (torch.float32, torch.uint8):
            raise LdicError("fused_tail: image must be fp32 in [-1,1] or its uint8 levels")
        image = _req(image, None, "image").contiguous()
        conv_w = _req(conv_w, torch.float32, "conv_w").contiguous()
        if x.dim() != 4 or x.shape[-1] != self.cin_pad or not x.is_contiguous():
            raise LdicError(f"conv input must be contiguous NHWC bf16 with {self.cin_pad} channels, got {tuple(x.shape)}")
        B, H, W, _ = x.shape
        if tuple(image.shape) != (B, 3, 2 * H, 2 * W) or conv_w.numel() != B * 3 * self.cout:
            raise LdicError("fused_tail: image must be (B,3,2H,2W) and conv_w (B,3,Cout)")
        sq = torch.zeros(B, dtype=torch.int64, device=x.device)
        xo = torch.empty(image.shape, dtype=torch.float32, device=x.device) if want_x_tilde else None
        out = torch.empty(self.out_dims(B, H, W), dtype=torch.float32, device=x.device) if want_out else None
        t = _lib.ConvTail(_ptr(image), _ptr(conv_w), _ptr(xo), _ptr(sq), 2 * H, 2 * W, int(image.dtype == torch.uint8), int(bool(tanh_out)))
        d = self._desc(B, H, W)
        prof = PROFILE
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(_L().ldic_conv_forward_fused_tail(C.byref(d), _ptr(x), _ptr(self.w_packed), _ptr(self.bias_packed),
                                                _ptr(self.gamma_bf16), _ptr(self.beta_tiled), _ptr(out), C.byref(t), _stream()),
              "ldic_conv_forward_fused_tail")
        if prof is not None:
            e1.record()
            prof.append((self, (B, H, W), e0, e1))
        return sq, xo, out

    def column_of_band(self, x: torch.Tensor, x_org: int) -> torch.Tensor:
        """LDIC_CTX_CONV1 on a (B,h,w_in,2N) band of the (sheared) context input, for the ONE output column whose patches
        start at band column x_org (aux1 = M | shear << 8 | x_org << 12 | w_in << 16): (B*h, 4, 4, N)."""
        if self.kind != _lib.LDIC_CTX_CONV1:
            raise LdicError("column_of_band: context conv 1 only")
        _req(x, torch.bfloat16, "x")
        if x.dim() != 4 or x.shape[-1] != self.cin_pad:
            raise LdicError(f"conv input must be NHWC bf16 with {self.cin_pad} channels, got {tuple(x.shape)}")
        B, H, w_in, Cp = x.shape
        # the band may be a column slice x_full[:, :, t:t+w_in] of a wider contiguous image: only the row pitch differs
        pitch = x.stride(1) // Cp
        if (x.stride(3) != 1 or x.stride(2) != Cp or x.stride(1) != pitch * Cp or x.stride(0) != H * pitch * Cp or pitch < w_in
                or pitch >= (1 << 19)):
            raise LdicError("column_of_band: x must be a contiguous NHWC image or a column slice of one")
        if not (0 <= x_org < 16 and x_org < w_in < 32768 and self.aux[1] < (1 << 12) and self.aux[0] < (1 << 12)):
            raise LdicError("column_of_band: bad band geometry")
        out = torch.empty(self.out_dims(B, H, 1), dtype=torch.float32 if self.out_f32 else torch.bfloat16, device=x.device)
        d = self._desc(B, H, 1, 0)
        d.aux0 = int(self.aux[0]) | ((int(pitch) << 12) if pitch != w_in else 0)
        d.aux1 = int(self.aux[1]) | (int(x_org) << 12) | (int(w_in) << 16)
        check(_L().ldic_conv_forward(C.byref(d), _ptr(x), _ptr(self.w_packed), _ptr(self.bias_packed),
                                     _ptr(self.gamma_bf16), _ptr(self.beta_tiled), _ptr(out), _stream()), "ldic_conv_forward")
        return out

    def __call__(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, sm_limit: int = 0,
                 residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        u8 = False
        if self.kind == _lib.LDIC_CONV_FIRST_5x5S2:
            u8 = x.dtype == torch.uint8          # 8-bit levels: the kernel applies x = (u/255)*2-1 itself (eval_net.py:84)
            _req(x, torch.uint8 if u8 else torch.float32, "x")
            if x.dim() != 4 or x.shape[1] != self.cin or not x.is_contiguous():
                raise LdicError(f"first conv input must be a contiguous NCHW fp32 / uint8 image with {self.cin} channels, got {tuple(x.shape)}")
            B, _, H, W = x.shape
        else:
            _req(x, torch.float32 if self.tf32 else torch.bfloat16, "x")
            if x.dim() != 4 or x.shape[-1] != self.cin_pad or not x.is_contiguous():
                raise LdicError(f"conv input must be contiguous NHWC {'fp32' if self.tf32 else 'bf16'} with {self.cin_pad} "
                                f"channels, got {tuple(x.shape)}")
            B, H, W, _ = x.shape
        if out is None:
            out = torch.empty(self.out_dims(B, H, W), dtype=torch.float32 if self.out_f32 else torch.bfloat16,
                              device=x.device)
        d = self._desc(B, H, W, sm_limit)
        if u8:
            d.aux0 = 1
        prof = PROFILE
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if residual is not None:          # y = act(conv(x) + bias) + residual (NHWC bf16 of the output's shape)
            _req(residual, torch.bfloat16, "residual")
            if tuple(residual.shape) != tuple(out.shape) or not residual.is_contiguous():
                raise LdicError("conv residual must be a contiguous NHWC bf16 tensor of the output's shape")
            check(_L().ldic_conv_forward_residual(C.byref(d), _ptr(x), _ptr(self.w_packed), _ptr(self.bias_packed),
                                                  _ptr(residual), _ptr(out), _stream()), "ldic_conv_forward_residual")
        else:
            check(_L().ldic_conv_forward(C.byref(d), _ptr(x), _ptr(self.w_packed), _ptr(self.bias_packed),
                                         _ptr(self.gamma_bf16), _ptr(self.beta_tiled), _ptr(out), _stream()),
                  "ldic_conv_forward")
        if prof is not None:
            e1.record()
            prof.append((self, (B, H, W), e0, e1))
        return out


# --------------------------------------------------------------------------------------
# f2: window attention (layers/win_attention.py)
# --------------------------------------------------------------------------------------
def window_attention_bias(table: torch.Tensor, index: torch.Tensor, heads: int, ws: int) -> torch.Tensor:
    """relative_position_bias_table[(2ws-1)^2, heads] + relative_position_index[ws^2, ws^2] -> bias [heads, ws^2, ws^2]."""
    table = _req(table, torch.float32, "relative_position_bias_table").contiguous()
    index = _req(index, torch.int64, "relative_position_index").contiguous()
    n = ws * ws
    if tuple(table.shape) != ((2 * ws - 1) ** 2, heads) or index.numel() != n * n:
        raise LdicError("window_attention_bias: table must be ((2ws-1)^2, heads) and index (ws^2, ws^2)")
    bias = torch.empty(heads, n, n, dtype=torch.float32, device=table.device)
    check(_L().ldic_window_attention_bias(_ptr(table), _ptr(index), _ptr(bias), int(heads), int(ws), _stream()),
          "ldic_window_attention_bias")
    return bias


def window_attention_core(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, bias: torch.Tensor, heads: int, ws: int,
                          shift: int) -> torch.Tensor:
    """q (pre-scaled), k, v: contiguous NHWC bf16 [B,H,W,C] -> softmax(q k^T + bias + shift mask) v, same layout."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, torch.bfloat16, n)
        if t.dim() != 4 or not t.is_contiguous() or t.shape != q.shape:
            raise LdicError("window_attention_core: q, k, v must be contiguous NHWC bf16 tensors of one shape")
    _req(bias, torch.float32, "bias")
    B, H, W, Cc = q.shape
    out = torch.empty_like(q)
    if bias.dim() == 2:          # the module's relative_position_bias_table [(2ws-1)^2, heads]: indexed inside the kernel
        if tuple(bias.shape) != ((2 * ws - 1) ** 2, heads) or not bias.is_contiguous():
            raise LdicError("window_attention_core: bias table must be contiguous ((2ws-1)^2, heads)")
        check(_L().ldic_window_attention_core_table(_ptr(q), _ptr(k), _ptr(v), _ptr(bias), _ptr(out), B, H, W, Cc, int(heads),
                                                    int(ws), int(shift), _stream()), "ldic_window_attention_core_table")
        return out
    if tuple(bias.shape) != (heads, ws * ws, ws * ws) or not bias.is_contiguous():
        raise LdicError("window_attention_core: bias must be contiguous (heads, ws^2, ws^2) or the table ((2ws-1)^2, heads)")
    check(_L().ldic_window_attention_core(_ptr(q), _ptr(k), _ptr(v), _ptr(bias), _ptr(out), B, H, W, Cc, int(heads), int(ws),
                                          int(shift), _stream()), "ldic_window_attention_core")
    return out


def residual_nhwc_to_nchw(o_nhwc: torch.Tensor, shortcut_nchw: torch.Tensor) -> torch.Tensor:
    """shortcut (B,C,H,W) fp32 + o (B,H,W,Cp>=C) fp32 -> (B,C,H,W) fp32."""
    o = _req(o_nhwc, torch.float32, "o")
    sc = _req(shortcut_nchw, torch.float32, "shortcut").contiguous()
    B, Cc, H, W = sc.shape
    if o.dim() != 4 or not o.is_contiguous() or tuple(o.shape[:3]) != (B, H, W) or o.shape[3] < Cc:
        raise LdicError("residual_nhwc_to_nchw: o must be contiguous (B,H,W,Cp>=C)")
    y = torch.empty_like(sc)
    check(_L().ldic_residual_nhwc_to_nchw_f32(_ptr(o), _ptr(sc), _ptr(y), B, Cc, H, W, int(o.shape[3]), _stream()),
          "ldic_residual_nhwc_to_nchw_f32")
    return y


def gate_residual_nhwc_to_nchw(a_nhwc: torch.Tensor, b_nhwc: torch.Tensor, x_nchw: torch.Tensor) -> torch.Tensor:
    """x (B,C,H,W) fp32 + a * sigmoid(b), a / b contiguous NHWC bf16 (B,H,W,Cp>=C) -> (B,C,H,W) fp32."""
    _req(a_nhwc, torch.bfloat16, "a"); _req(b_nhwc, torch.bfloat16, "b")
    x = _req(x_nchw, torch.float32, "x").contiguous()
    B, Cc, H, W = x.shape
    if a_nhwc.shape != b_nhwc.shape or not a_nhwc.is_contiguous() or not b_nhwc.is_contiguous() or \
            tuple(a_nhwc.shape[:3]) != (B, H, W) or a_nhwc.shape[3] < Cc:
        raise LdicError("gate_residual_nhwc_to_nchw: a, b must be contiguous (B,H,W,Cp>=C) bf16 tensors")
    y = torch.empty_like(x)
    check(_L().ldic_gate_residual_nhwc_to_nchw_f32(_ptr(a_nhwc), _ptr(b_nhwc), _ptr(x), _ptr(y), B, Cc, H, W,
                                                   int(a_nhwc.shape[3]), _stream()), "ldic_gate_residual_nhwc_to_nchw_f32")
    return y


def syntax_branch(y: torch.Tensor, h2: torch.Tensor, M: int, syntax_model, prediction_model_syntax, conv_weights_gen,
                  z3_round_in: Optional[torch.Tensor] = None):
    """The syntax side branch on libldic_b200 (ldic_syntax_branch): y, h2 are NHWC fp32 [B,h,w,N].
    Returns (z3 [B,M,1,1], z3_round, mu, sigma [B,M,1,1], conv_w [B,3,M,1,1]).  `z3_round_in` ([B,M] fp32, the decoder's
    path): conv_w is generated from these symbols instead of round(Syntax_Model(y))."""
    _req(y, torch.float32, "y"); _req(h2, torch.float32, "h2")
    if y.shape != h2.shape or not y.is_contiguous() or not h2.is_contiguous():
        raise LdicError("syntax_branch: y and h2 must be contiguous NHWC fp32 tensors of the same shape")
    B, h, w, N = y.shape
    h1, w1 = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    h2s, w2s = (h1 - 1) // 2 + 1, (w1 - 1) // 2 + 1
    dev = y.device
    f = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
    a = _lib.SyntaxArgs()
    a.B, a.h, a.w, a.N, a.M = B, h, w, N, M
    a.y, a.h2 = _ptr(y), _ptr(h2)
    keep = []

    def P(t):
        t = _req(t.detach(), torch.float32, "weight").contiguous()
        keep.append(t)
        return t.data_ptr()
    sm, ps, cg = syntax_model, prediction_model_syntax, conv_weights_gen.transform
    a.sm_down0_w, a.sm_down0_b, a.sm_down1_w, a.sm_down1_b = P(sm.down0.weight), P(sm.down0.bias), P(sm.down1.weight), P(sm.down1.bias)
    a.sm_conv_w, a.sm_conv_b = P(sm.conv.weight), P(sm.conv.bias)
    a.ps_down0_w, a.ps_down0_b, a.ps_down1_w, a.ps_down1_b = P(ps.down0.weight), P(ps.down0.bias), P(ps.down1.weight), P(ps.down1.bias)
    a.ps_fc_w, a.ps_fc_b = P(ps.fc.weight), P(ps.fc.bias)
    a.cg_w0, a.cg_b0, a.cg_w1, a.cg_b1, a.cg_w2, a.cg_b2 = (P(cg[0].weight), P(cg[0].bias), P(cg[2].weight), P(cg[2].bias),
                                                            P(cg[4].weight), P(cg[4].bias))
    if (tuple(sm.down0.weight.shape) != (32, M, 3, 3) or tuple(sm.down1.weight.shape) != (64, 32, 3, 3)
            or tuple(ps.down0.weight.shape) != (M, N, 3, 3) or tuple(ps.down1.weight.shape) != (M, M, 3, 3)
            or tuple(ps.fc.weight.shape) != (2 * M, N + 2 * M) or tuple(cg[4].weight.shape) != (3 * M, 256)):
        raise LdicError("syntax_branch: unexpected module shapes")
    ds1, ds2, p0, p1 = f(B, h1, w1, 32), f(B, h2s, w2s, 64), f(B, h1, w1, M), f(B, h2s, w2s, M)
    part = f(int(_L().ldic_syntax_workspace_elems(B, h, w, N, M)))
    a.sm_ds1, a.sm_ds2, a.ps_ds0, a.ps_ds1, a.pool_part = _ptr(ds1), _ptr(ds2), _ptr(p0), _ptr(p1), _ptr(part)
    z3, z3r, mu, sg, cw = f(B, M, 1, 1), f(B, M, 1, 1), f(B, M, 1, 1), f(B, M, 1, 1), f(B, 3, M, 1, 1)
    a.z3, a.z3_round, a.mu, a.sigma, a.conv_w = _ptr(z3), _ptr(z3r), _ptr(mu), _ptr(sg), _ptr(cw)
    if z3_round_in is not None:
        z3_round_in = _req(z3_round_in, torch.float32, "z3_round_in").reshape(B, M).contiguous()
        keep.append(z3_round_in)
    a.z3_round_in = _ptr(z3_round_in)
    check(_L().ldic_syntax_branch(C.byref(a), _stream()), "ldic_syntax_branch")
    return z3, z3r, mu, sg, cw


def ctx_pack_input(y_round_bf16: torch.Tensor, h2: torch.Tensor) -> torch.Tensor:
    """[B,h,w,N] bf16 (rounded latent) and [B,h,w,N] fp32 (h_s output) -> [B,h,w,2N] bf16."""
    _req(y_round_bf16, torch.bfloat16, "y_round")
    _req(h2, torch.float32, "h2")
    if y_round_bf16.shape != h2.shape or not y_round_bf16.is_contiguous() or not h2.is_contiguous():
        raise LdicError("ctx_pack_input: y_round and h2 must be contiguous NHWC tensors of the same shape")
    B, h, w, N = h2.shape
    x = torch.empty(B, h, w, 2 * N, dtype=torch.bfloat16, device=h2.device)
    check(_L().ldic_ctx_pack_input(_ptr(y_round_bf16), _ptr(h2), _ptr(x), B * h * w, N, _stream()), "ldic_ctx_pack_input")
    return x


def conv_reference_f32(kind: int, x_nhwc: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """CUDA-core fp32 direct convolution (validation aid, ldic_conv_forward_f32_reference_kernel)."""
    x = _req(x_nhwc, torch.float32, "x").contiguous()
    w = _req(weight, torch.float32, "w").contiguous()
    transposed = kind in (_lib.LDIC_DECONV_GS_5x5, _lib.LDIC_DECONV_HS_5x5, _lib.LDIC_DECONV_S1_3x3,
                          _lib.LDIC_DECONV_GS_5x5_MERGED)
    if kind == _lib.LDIC_CONV_1x1:
        w = w.reshape(w.shape[0], -1).contiguous()
        cout, cin = w.shape
    elif transposed:
        cin, cout = w.shape[0], w.shape[1]
    else:
        cout, cin = w.shape[0], w.shape[1]
    B, H, W, _ = x.shape
    dq = ConvDesc(kind, B, H, W, cin, 8, _pad64(cin), 64, 0, 1, 0, 0, 0, 0)    # shape query only
    d = ConvDesc(kind, B, H, W, cin, cout, _pad64(cin), 64, 0, 1, 0, 0, 0, 0)
    ho, wo = C.c_int(), C.c_int()
    _L().ldic_conv_out_shape(C.byref(dq), C.byref(ho), C.byref(wo))
    y = torch.empty(B, ho.value, wo.value, cout, dtype=torch.float32, device=x.device)
    check(_L().ldic_conv_forward_f32_reference_kernel(C.byref(d), _ptr(x), _ptr(w), _ptr(bias), _ptr(y), _stream()),
          "conv_reference_f32")
    return y
