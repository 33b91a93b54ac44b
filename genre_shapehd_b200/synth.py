"""Synthetic inputs for the benchmark configs: numpy only, no oracle, no GPU."""
import numpy as np


def sphere_depth(h=256, w=256, fl=418.3, cam_dist=2.2, radius=0.4, background=0.0):
    """Analytic ray depth of a sphere at the origin seen from (-cam_dist, 0, 0); fp32 [h, w].
    Ray of pixel (h~, w~): (fl, -w~, -h~)/norm, the convention of back_projection_kernel.cu:239-242."""
    hh = np.arange(h, dtype=np.float64)[:, None] - (h - 1) / 2.0
    ww = np.arange(w, dtype=np.float64)[None, :] - (w - 1) / 2.0
    norm = np.sqrt(hh * hh + ww * ww + fl * fl)
    dx = fl / norm
    bq = -cam_dist * dx
    disc = bq * bq - (cam_dist * cam_dist - radius * radius)
    t = -bq - np.sqrt(np.maximum(disc, 0.0))
    return np.where(disc > 0, t, background).astype(np.float32)


def uniform_depth(seed, h=256, w=256, lo=1.7, hi=2.7, fg=0.6, background=0.0):
    """d ~ U(lo, hi) on a random `fg` fraction of the pixels, `background` elsewhere."""
    rng = np.random.RandomState(seed)
    d = rng.uniform(lo, hi, size=(h, w))
    m = rng.uniform(size=(h, w)) < fg
    return np.where(m, d, background).astype(np.float32)


def bench_depth_batch(n, h=256, w=256):
    """BASELINE.json configs[1] input: [n,1,h,w]; even samples are spheres, odd ones U(1.7,2.7) on a 60 % mask;
    the generator seed is the sample index."""
    out = np.empty((n, 1, h, w), np.float32)
    for i in range(n):
        out[i, 0] = sphere_depth(h, w, radius=0.3 + 0.01 * (i % 16)) if i % 2 == 0 else uniform_depth(i, h, w)
    return out
