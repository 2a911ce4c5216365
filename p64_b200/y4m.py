"""Y4M (YUV4MPEG2, 4:2:0) reading/writing and the seeded synthetic clips used by tests and bench.

The reference ingests Y4M through its vendored Daala reader (y4m_input.c:556-677, io.c:635-644); with
``C420jpeg`` no chroma conversion runs (y4m_input.c:588-596), so a frame is just the three planes
back to back.  Only that layout is produced/consumed here.
"""
from __future__ import annotations

import numpy as np

# image types, same numbering as the C ABI (include/p64_b200.h) -- globals.h IT_NTSC/IT_CIF/IT_QCIF
IT_NTSC, IT_CIF, IT_QCIF = 0, 1, 2
DIMS = {IT_NTSC: (352, 240), IT_CIF: (352, 288), IT_QCIF: (176, 144)}
FLAG = {IT_NTSC: "-NTSC", IT_CIF: "-CIF", IT_QCIF: "-QCIF"}


def frame_bytes(image_type: int) -> int:
    w, h = DIMS[image_type]
    return w * h * 3 // 2


def synth_clip(image_type: int, n_frames: int, seed: int = 1234, pan=(3, 2), noise: int = 20,
               temporal_noise: int = 3) -> np.ndarray:
    """Seeded synthetic clip, uint8 [n_frames, W*H*3/2] (planar Y,U,V per frame).

    Smooth sinusoid texture + static per-pixel noise, global pan, a moving 48x48 patch and small
    temporal noise; chroma are slow gradients that drift.
    """
    w, h = DIMS[image_type]
    rng = np.random.default_rng(seed)
    big_h, big_w = h + 64 + abs(pan[1]) * n_frames, w + 64 + abs(pan[0]) * n_frames
    yy, xx = np.mgrid[0:big_h, 0:big_w].astype(np.float64)
    tex = 128 + 60 * np.sin(xx / 17.0) * np.cos(yy / 23.0) + 30 * np.sin((xx + yy) / 7.0)
    tex = tex + rng.integers(-noise, noise + 1, size=tex.shape)
    patch = rng.integers(0, 256, size=(48, 48)).astype(np.float64)
    out = np.empty((n_frames, w * h * 3 // 2), dtype=np.uint8)
    cy, cx = np.mgrid[0:h // 2, 0:w // 2].astype(np.float64)
    by, bx = max(32, -pan[1] * (n_frames - 1)), max(32, -pan[0] * (n_frames - 1))     # long clips panning up / left start further in
    for f in range(n_frames):
        oy, ox = by + pan[1] * f, bx + pan[0] * f
        y = tex[oy:oy + h, ox:ox + w].copy()
        py, px = (20 + 5 * f) % (h - 48), (30 + 7 * f) % (w - 48)
        y[py:py + 48, px:px + 48] = patch
        if temporal_noise:
            y = y + rng.integers(-temporal_noise, temporal_noise + 1, size=y.shape)
        u = 128 + 40 * np.sin((cx + 2 * f) / 31.0) + 0.1 * cy
        v = 128 + 40 * np.cos((cy + f) / 29.0) - 0.1 * cx
        out[f, :w * h] = np.clip(np.rint(y), 0, 255).astype(np.uint8).ravel()
        out[f, w * h:w * h + w * h // 4] = np.clip(np.rint(u), 0, 255).astype(np.uint8).ravel()
        out[f, w * h + w * h // 4:] = np.clip(np.rint(v), 0, 255).astype(np.uint8).ravel()
    return out


def random_pair(image_type: int, seed: int, shift=(0, 0), noise: int = 4):
    """(reference luma, current luma) uint8 [H,W]: uniform-random base and a shifted, noised copy
    (SURVEY.md 8(d), config 4)."""
    w, h = DIMS[image_type]
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(h + 64, w + 64), dtype=np.int64)
    ref = base[32:32 + h, 32:32 + w]
    cur = base[32 + shift[1]:32 + shift[1] + h, 32 + shift[0]:32 + shift[0] + w]
    cur = np.clip(cur + rng.integers(-noise, noise + 1, size=cur.shape), 0, 255)
    return ref.astype(np.uint8), cur.astype(np.uint8)


def synth_payloads(image_type: int, n_frames: int, seed: int, chroma: str) -> np.ndarray:
    """Seeded synthetic clip as unconverted Y4M frame payloads of chroma type `chroma` (uint8 [n_frames, payload]):
    the luma of synth_clip(); chroma planes of the type's own shape (the 4:2:0 planes stretched, plus seeded noise so
    the re-siting filters have something to do); an alpha ramp for 444alpha."""
    w, h = DIMS[image_type]
    base = synth_clip(image_type, n_frames, seed)
    if chroma in ("420", "420jpeg"):
        return base
    rng = np.random.default_rng(seed + 7919)
    cw, ch = w // 2, h // 2
    shape = {"420mpeg2": (ch, cw), "420paldv": (ch, cw), "422": (h, cw), "411": (h, (w + 3) // 4), "444": (h, w),
             "444alpha": (h, w), "mono": None}[chroma]
    out = []
    for f in range(n_frames):
        parts = [base[f, :w * h]]
        if shape is not None:
            for pl in range(2):
                c = base[f, w * h + pl * cw * ch:w * h + (pl + 1) * cw * ch].reshape(ch, cw).astype(np.int64)
                c = np.repeat(c, shape[0] // ch, axis=0)
                c = np.repeat(c, 2, axis=1) if shape[1] == w else (c[:, ::2] if shape[1] < cw else c)
                c = np.clip(c + rng.integers(-24, 25, size=c.shape), 0, 255)
                parts.append(c.astype(np.uint8).ravel())
            if chroma == "444alpha":
                parts.append((np.arange(w * h) % 251).astype(np.uint8))
        out.append(np.concatenate(parts))
    return np.stack(out)


def write_y4m(path: str, image_type: int, frames: np.ndarray, rate=(30000, 1001), chroma: str = "420jpeg") -> None:
    w, h = DIMS[image_type]
    with open(path, "wb") as f:
        f.write(f"YUV4MPEG2 W{w} H{h} F{rate[0]}:{rate[1]} Ip C{chroma}\n".encode())
        for fr in frames:
            f.write(b"FRAME\n")
            f.write(np.ascontiguousarray(fr, dtype=np.uint8).tobytes())


def read_y4m(path: str):
    """-> (width, height, uint8 [n_frames, W*H*3/2]).  4:2:0 without chroma re-siting only."""
    with open(path, "rb") as f:
        data = f.read()
    nl = data.index(b"\n")
    hdr = data[:nl].split()
    if hdr[0] != b"YUV4MPEG2":
        raise ValueError("not a YUV4MPEG2 file")
    w = h = 0
    for tok in hdr[1:]:
        if tok[:1] == b"W":
            w = int(tok[1:])
        elif tok[:1] == b"H":
            h = int(tok[1:])
        elif tok[:1] == b"C" and not tok.startswith(b"C420"):
            raise ValueError(f"unsupported chroma format {tok!r}")
    fb = w * h * 3 // 2
    frames, pos = [], nl + 1
    while pos < len(data):
        e = data.index(b"\n", pos)
        if not data[pos:e].startswith(b"FRAME"):
            raise ValueError("bad frame header")
        frames.append(np.frombuffer(data, dtype=np.uint8, count=fb, offset=e + 1))
        pos = e + 1 + fb
    return w, h, np.stack(frames) if frames else np.empty((0, fb), np.uint8)
