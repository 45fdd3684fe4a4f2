"""Dependency-light visualisation for the training scripts: loss curves as PNG and field animations as GIF
(the reference draws both with matplotlib, ``experiment_navier_stokes.py:198-227``; only Pillow is available
here, so the plots are rasterised directly).  Everything degrades to a warning when Pillow is missing."""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np

__all__ = ["save_curves_png", "save_field_gif"]


def _pil():
    try:
        from PIL import Image, ImageDraw
        return Image, ImageDraw
    except Exception:                                    # noqa: BLE001
        import warnings
        warnings.warn("Pillow is not installed: visualisation skipped")
        return None, None


def _colormap(v: np.ndarray) -> np.ndarray:
    """[0, 1] -> RGB (a compact blue-white-red diverging map)."""
    v = np.clip(v, 0.0, 1.0)[..., None]
    lo, mid, hi = np.array([33, 64, 154.0]), np.array([245, 245, 245.0]), np.array([178, 24, 43.0])
    return np.where(v < 0.5, lo + (mid - lo) * (v * 2), mid + (hi - mid) * (v * 2 - 1)).astype(np.uint8)


def save_curves_png(path: str, curves: Dict[str, Sequence[float]], size=(640, 400), log_y: bool = True) -> bool:
    """Line plot of named series (e.g. train / test loss per epoch)."""
    Image, ImageDraw = _pil()
    if Image is None or not any(len(v) for v in curves.values()):
        return False
    W, H, pad = size[0], size[1], 45
    img = Image.new("RGB", (W, H), "white")
    dr = ImageDraw.Draw(img)
    vals = np.concatenate([np.asarray(v, dtype=np.float64) for v in curves.values() if len(v)])
    vals = vals[np.isfinite(vals) & ((vals > 0) if log_y else True)]
    if vals.size == 0:
        return False
    f = (lambda a: np.log10(np.maximum(a, 1e-30))) if log_y else (lambda a: a)
    lo, hi = float(f(vals).min()), float(f(vals).max())
    hi = hi if hi > lo else lo + 1.0
    n = max(len(v) for v in curves.values())
    dr.rectangle([pad, 10, W - 10, H - pad], outline="black")
    colors = ["#1f77b4", "#d62728", "#2ca02c", "#9467bd"]
    for c, (name, v) in zip(colors, curves.items()):
        if not len(v):
            continue
        y = f(np.asarray(v, dtype=np.float64))
        xs = pad + (np.arange(len(v)) / max(n - 1, 1)) * (W - 10 - pad)
        ys = (H - pad) - (y - lo) / (hi - lo) * (H - pad - 10)
        dr.line(list(zip(xs.tolist(), ys.tolist())), fill=c, width=2)
        dr.text((pad + 8, 14 + 12 * colors.index(c)), f"{name}: {v[-1]:.4g}", fill=c)
    dr.text((5, 10), f"{(10 ** hi if log_y else hi):.2g}", fill="black")
    dr.text((5, H - pad - 10), f"{(10 ** lo if log_y else lo):.2g}", fill="black")
    dr.text((W // 2 - 20, H - 30), "epoch", fill="black")
    img.save(path)
    return True


def save_field_gif(path: str, fields: Dict[str, np.ndarray], scale: int = 4, ms_per_frame: int = 120) -> bool:
    """Side-by-side animation of 2-D fields over time: ``fields[name]`` is ``[X, Y, T]`` (e.g. truth / prediction)."""
    Image, ImageDraw = _pil()
    if Image is None or not fields:
        return False
    arrs = {k: np.asarray(v, dtype=np.float64) for k, v in fields.items()}
    lo = min(float(a.min()) for a in arrs.values())
    hi = max(float(a.max()) for a in arrs.values())
    hi = hi if hi > lo else lo + 1.0
    T = min(a.shape[-1] for a in arrs.values())
    frames = []
    for t in range(T):
        tiles = []
        for name, a in arrs.items():
            rgb = _colormap((a[..., t] - lo) / (hi - lo))
            tile = Image.fromarray(rgb).resize((a.shape[1] * scale, a.shape[0] * scale), Image.NEAREST)
            ImageDraw.Draw(tile).text((3, 3), f"{name} t={t}", fill="black")
            tiles.append(tile)
        w, h = sum(t_.width for t_ in tiles) + 4 * (len(tiles) - 1), max(t_.height for t_ in tiles)
        frame = Image.new("RGB", (w, h), "white")
        x = 0
        for t_ in tiles:
            frame.paste(t_, (x, 0))
            x += t_.width + 4
        frames.append(frame)
    frames[0].save(path, save_all=True, append_images=frames[1:], duration=ms_per_frame, loop=0)
    return True
