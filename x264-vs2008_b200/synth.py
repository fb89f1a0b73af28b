"""Seeded synthetic YUV420 content: band-limited noise field sampled with a per-frame global
translation, a few rectangles moving with their own vectors, plus small i.i.d. noise per frame.  Pure integer
numpy, reproducible from (width, height, seed)."""
import numpy as np


def _box_blur(a, n):
    a = a.astype(np.int32)
    for _ in range(n):
        a = (a + np.roll(a, 1, 0) + np.roll(a, -1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 1) + 2) // 5
    return a


class Clip:
    def __init__(self, width, height, seed=1234, blur=6, motion=(5, 3), noise=2, nrect=3):
        self.w, self.h, self.seed, self.motion, self.noise = width, height, seed, motion, noise
        rng = np.random.default_rng(seed)
        fw, fh = width + 256, height + 256
        # low-res noise upsampled x4 then blurred -> band-limited texture with large and small scale structure
        base = rng.integers(0, 256, (fh // 4 + 1, fw // 4 + 1), dtype=np.int32)
        field = np.kron(base, np.ones((4, 4), np.int32))[:fh, :fw]
        field = (field * 3 + rng.integers(0, 256, (fh, fw), dtype=np.int32)) // 4
        field = _box_blur(field, blur)
        lo, hi = int(field.min()), int(field.max())
        self.field = ((field - lo) * 255 // max(1, hi - lo)).astype(np.uint8)
        self.rects = []
        for _ in range(nrect):
            rw, rh = int(rng.integers(width // 16 + 8, width // 6 + 16)), int(rng.integers(height // 16 + 8, height // 6 + 16))
            self.rects.append(dict(x=int(rng.integers(0, max(1, width - rw))), y=int(rng.integers(0, max(1, height - rh))),
                                   w=rw, h=rh, vx=int(rng.integers(-7, 8)), vy=int(rng.integers(-5, 6)),
                                   tex=rng.integers(0, 256, (rh, rw), dtype=np.uint8) // 2 + 64))

    def luma(self, n):
        """frame n luma, (h, w) uint8"""
        ox = 128 + (n * self.motion[0]) % 96
        oy = 128 + (n * self.motion[1]) % 96
        y = self.field[oy:oy + self.h, ox:ox + self.w].astype(np.int16)
        for r in self.rects:
            x0 = (r["x"] + n * r["vx"]) % max(1, self.w - r["w"])
            y0 = (r["y"] + n * r["vy"]) % max(1, self.h - r["h"])
            sub = y[y0:y0 + r["h"], x0:x0 + r["w"]]
            sub[...] = r["tex"][:sub.shape[0], :sub.shape[1]]  # (cropped only for pictures smaller than a rectangle)
        if self.noise:
            rng = np.random.default_rng(self.seed * 1000003 + n)
            y = y + rng.integers(-self.noise, self.noise + 1, y.shape, dtype=np.int16)
        return np.clip(y, 0, 255).astype(np.uint8)

    def yuv420(self, n):
        y = self.luma(n)
        u = (255 - y[::2, ::2] // 2 - 32).astype(np.uint8)
        v = (y[1::2, 1::2] // 2 + 64).astype(np.uint8)
        return y, u, v
