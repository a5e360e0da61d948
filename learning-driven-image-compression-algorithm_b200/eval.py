"""Evaluation driver of the reference (eval_net.py:19-128, the branch without pre-processing) on the B200 path.

What the reference does per image and this module reproduces:
  * ToTensor -> (3,h,w) in [0,1]; pad with ONES at the bottom / right to a multiple of 64 (eval_net.py:68-81);
    map to [-1,1] (:84);
  * Net((1,h,w,3),(1,h,w,3), is_high, post_processing) -- the UNPADDED size is the bpp denominator
    (model/net.py:856-859) while v_mse / v_psnr are taken over the padded tensor (:864-869);
  * load_state_dict(torch.load(weight_path), strict=True); forward(data, 'test', num).

Unlike the reference (one image per forward, a new Net per image) images that share a padded size are batched;
per-image bpp is formed from the per-image sums of ln L (the reference only returns the batch mean).
The module never touches a CPU fallback: tensors go to the current CUDA device and through libldic_b200.
"""
from __future__ import annotations

import glob
import math
from typing import Dict, Iterable, List, Sequence, Tuple

import torch

from . import ops
from .layers import psnr_from_sq_err
from .net import Net


def pad_to_multiple(img: torch.Tensor, multiple: int = 64) -> torch.Tensor:
    """eval_net.py:68-84: (3,h,w) in [0,1] -> (1,3,hp,wp) in [-1,1], padded with ONES (white) bottom / right."""
    if img.dim() != 3:
        raise ValueError("expected a (C,h,w) image")
    c, h, w = img.shape
    hp = h if h % multiple == 0 else (h // multiple) * multiple + multiple
    wp = w if w % multiple == 0 else (w // multiple) * multiple + multiple
    out = torch.ones(c, hp, wp, dtype=img.dtype, device=img.device)
    out[:, :h, :w] = img
    return out.unsqueeze(0) * 2.0 - 1.0


def load_image(path: str) -> torch.Tensor:
    """PIL -> (3,h,w) float32 in [0,1] (what torchvision's ToTensor returns for an 8-bit RGB image)."""
    from PIL import Image          # optional dependency, only for reading files
    import numpy as np
    a = np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8)
    return torch.from_numpy(a).permute(2, 0, 1).float().div_(255.0)


CONTAINER_MAGIC = b"LDIC"


def pack_container(h: int, w: int, hp: int, wp: int, streams: Dict[str, bytes]) -> bytes:
    """One image's bitstreams as a file: 'LDIC' | u32 h, w, hp, wp | u32 len(z), len(y), len(syntax) | z | y | syntax
    (each part is a self-describing rANS segment of csrc/rans.cu)."""
    import struct
    parts = [streams[k] for k in ("z", "y", "syntax")]
    return CONTAINER_MAGIC + struct.pack("<7I", h, w, hp, wp, *(len(p) for p in parts)) + b"".join(parts)


def unpack_container(blob: bytes):
    """Inverse of pack_container: (h, w, hp, wp, {"z","y","syntax"})."""
    import struct
    if blob[:4] != CONTAINER_MAGIC or len(blob) < 32:
        raise ValueError("not an LDIC container")
    h, w, hp, wp, nz, ny, ns = struct.unpack_from("<7I", blob, 4)
    if 32 + nz + ny + ns != len(blob):
        raise ValueError("LDIC container: sizes do not add up")
    o = 32
    return h, w, hp, wp, {"z": blob[o:o + nz], "y": blob[o + nz:o + nz + ny], "syntax": blob[o + nz + ny:]}


@torch.no_grad()
def evaluate_images(net: Net, images: Sequence[torch.Tensor], batch_size: int = 16,
                    bitstreams: bool = False) -> List[Dict[str, float]]:
    """Per-image {'bpp', 'psnr', 'mse', 'h', 'w'} exactly as the reference forms them for a batch of one
    (bpp over the unpadded h*w, MSE/PSNR over the padded tensor).  Images are grouped by padded size.
    `bitstreams`: also entropy-code every image (Net.entropy_encode) and add 'bpp_coded' = 8 * bytes / (h*w) and
    'container' (pack_container of its three streams) -- the reference only estimates 'bpp'."""
    dev = next(net.parameters()).device
    groups: Dict[Tuple[int, int], List[int]] = {}
    padded = []
    for i, img in enumerate(images):
        x = pad_to_multiple(img.float())
        padded.append(x)
        groups.setdefault((x.shape[2], x.shape[3]), []).append(i)
    out: List[Dict[str, float]] = [None] * len(images)          # type: ignore
    for (hp, wp), idx in groups.items():
        for k in range(0, len(idx), batch_size):
            chunk = idx[k:k + batch_size]
            xb = torch.cat([padded[i] for i in chunk], 0).to(dev, non_blocking=True)
            r = net.rd_forward(xb, per_image_bits=True)
            v_mse, _ = psnr_from_sq_err(r["sq_err"], 3 * hp * wp)
            bits = r["bits_per_image"].sum(1).cpu()              # sum of ln L over the three streams, per image
            coded = None
            if bitstreams:
                with torch.cuda.device(dev):
                    enc = net.entropy_encode(r)
                    coded = dict(zip(enc.keys(), ops.rans_tobytes(enc.values())))
            for j, i in enumerate(chunk):
                h, w = images[i].shape[1], images[i].shape[2]
                mse = float(v_mse[j])
                out[i] = {"bpp": float(bits[j]) / (-math.log(2) * h * w), "mse": mse,
                          "psnr": 20.0 * math.log10(255.0 / math.sqrt(mse)) if mse > 0 else float("inf"), "h": h, "w": w}
                if coded is not None:
                    blob = pack_container(h, w, hp, wp, {k: v[j] for k, v in coded.items()})
                    out[i]["container"] = blob
                    out[i]["bpp_coded"] = 8.0 * len(blob) / (h * w)
    return out


@torch.no_grad()
def decode_containers(net: Net, blobs: Sequence[bytes], batch_size: int = 16) -> List[torch.Tensor]:
    """The decoder side of `evaluate_images(bitstreams=True)`: LDIC containers -> (3,h,w) images in [0,1] (the padded
    reconstruction of Net.decompress cropped to the unpadded size, mapped back from [-1,1] and clamped).  Containers of
    one padded size are decoded together."""
    metas = [unpack_container(b) for b in blobs]
    groups: Dict[Tuple[int, int], List[int]] = {}
    for i, (_, _, hp, wp, _) in enumerate(metas):
        groups.setdefault((hp, wp), []).append(i)
    out: List[torch.Tensor] = [None] * len(blobs)               # type: ignore
    for (hp, wp), idx in groups.items():
        for k in range(0, len(idx), batch_size):
            chunk = idx[k:k + batch_size]
            x_hat = net.decompress([metas[i][4] for i in chunk], hp, wp)
            for j, i in enumerate(chunk):
                h, w = metas[i][0], metas[i][1]
                out[i] = ((x_hat[j, :, :h, :w] + 1.0) * 0.5).clamp_(0.0, 1.0)
    return out


def val(data_path: str, weight_path: str, is_high: bool = False, post_processing: bool = False, batch_size: int = 16,
        device: str = "cuda", bitstream_dir: str = None):
    """eval_net.py:19 `val` (pre_processing=False branch): prints the per-image line and the averages.
    `bitstream_dir`: also write one `<image name>.ldic` file per image and print the coded bpp next to the estimate."""
    paths = sorted(glob.glob(data_path))
    images = [load_image(p) for p in paths]
    net = Net((1, 64, 64, 3), (1, 64, 64, 3), is_high, post_processing).to(device).eval()
    net.load_state_dict(torch.load(weight_path, map_location=device), strict=True)
    res = evaluate_images(net, images, batch_size, bitstreams=bitstream_dir is not None)
    for p, r in zip(paths, res):
        if bitstream_dir is not None:
            import os
            os.makedirs(bitstream_dir, exist_ok=True)
            with open(os.path.join(bitstream_dir, os.path.splitext(os.path.basename(p))[0] + ".ldic"), "wb") as f:
                f.write(r["container"])
            print(p, r["bpp"], r["psnr"], r["mse"], "coded bpp", r["bpp_coded"])
        else:
            print(p, r["bpp"], r["psnr"], r["mse"])
    n = max(len(res), 1)
    print('[WITHOUT PRE-PROCESSING] bpp: %.4f psnr: %.4f  v_mse: %.4f' % (
        sum(r["bpp"] for r in res) / n, sum(r["psnr"] for r in res) / n, sum(r["mse"] for r in res) / n))
    return res


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="eval_net.py of the reference on libldic_b200")
    ap.add_argument("--data", required=True, help="glob of image files")
    ap.add_argument("--weights", required=True, help="reference checkpoint (state dict)")
    ap.add_argument("--high", action="store_true")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--bitstreams", default=None, metavar="DIR", help="write one .ldic bitstream file per image into DIR")
    ap.add_argument("--decode", default=None, metavar="DIR",
                    help="decoder mode: --data is a glob of .ldic files, reconstructions are written as PNG files into DIR")
    a = ap.parse_args()
    if a.decode is not None:
        import os
        from PIL import Image
        paths = sorted(glob.glob(a.data))
        net = Net((1, 64, 64, 3), (1, 64, 64, 3), a.high, False).to("cuda").eval()
        net.load_state_dict(torch.load(a.weights, map_location="cuda"), strict=True)
        os.makedirs(a.decode, exist_ok=True)
        for p, img in zip(paths, decode_containers(net, [open(p, "rb").read() for p in paths], a.batch)):
            arr = (img.permute(1, 2, 0) * 255.0).round().to(torch.uint8).cpu().numpy()
            Image.fromarray(arr).save(os.path.join(a.decode, os.path.splitext(os.path.basename(p))[0] + ".png"))
            print(p, "->", arr.shape)
    else:
        val(a.data, a.weights, is_high=a.high, batch_size=a.batch, bitstream_dir=a.bitstreams)
