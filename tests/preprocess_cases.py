"""Seeded synthetic frames shared by the preprocessing golden generator and tests (no file I/O, no network)."""
import numpy as np

# name -> (H, W, seed): landscape, portrait, square at / below / above the tower's 336, HD, odd sizes
CASES = {"hd_1080p": (1080, 1920, 1), "vga": (480, 640, 2), "portrait": (640, 360, 3), "square_336": (336, 336, 4),
         "square_500": (500, 500, 5), "tiny_upscale": (30, 50, 6), "odd": (61, 97, 7), "tall_odd": (97, 61, 8)}


def make_frame(h: int, w: int, seed: int) -> np.ndarray:
    """uint8 [h, w, 3]: smooth gradients + a sharp checker + noise, so both the anti-aliasing taps and the clipping matter."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([(x * 255) // max(w - 1, 1), (y * 255) // max(h - 1, 1), ((x + y) * 255) // max(h + w - 2, 1)], axis=-1)
    checker = (((x // 3) + (y // 5)) % 2 * 255)[..., None]
    noise = rng.integers(0, 256, (h, w, 3))
    sel = rng.integers(0, 3, (h, w, 1))
    img = np.where(sel == 0, base, np.where(sel == 1, checker, noise))
    return img.astype(np.uint8)
