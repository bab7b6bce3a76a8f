"""Synthetic inputs of BASELINE.json's five configs.

Everything is generated from numpy's PCG64 with fixed seeds, once per run, into arrays
that both the engine and the checkers consume; nothing is compared across machines.
Sizes are parameters so tests can use small cuts of the same distributions.
"""
import numpy as np


# ---- config 1: 2-D points of example/visualize_2d/main.cpp:18-43 -------------------------
def config1_points(n=1000, seed=1):
    """r ~ Normal(0, 5), theta ~ U[0, pi], p = (r cos theta, r sin theta); float64."""
    rng = np.random.default_rng(seed)
    r = rng.normal(0.0, 5.0, n)
    theta = rng.uniform(0.0, np.pi, n)
    return np.stack([r * np.cos(theta), r * np.sin(theta)], axis=1).astype(np.float64)


def config1_queries(n=10000, seed=2):
    """boxes: centre ~ U[-10,10]^2, half extents ~ U[0.1,3]; float64 [n][2][2]."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-10.0, 10.0, (n, 2))
    h = rng.uniform(0.1, 3.0, (n, 2))
    return np.stack([c - h, c + h], axis=1).astype(np.float64)


# ---- config 2 / 5: uniform 3-D points, small cubes ------------------------------------------
def uniform_points(n, dim=3, seed=42):
    rng = np.random.default_rng(seed)
    return rng.random((n, dim), dtype=np.float32)


def cube_half_width(n_points, hits_per_query=10.0, dim=3):
    """h with (2h)^dim * n_points = hits_per_query."""
    return np.float32(0.5 * (hits_per_query / float(n_points)) ** (1.0 / dim))


def cube_queries(n, n_points, dim=3, seed=43, hits_per_query=10.0):
    """cubes with centre ~ U[0,1)^dim sized for ~hits_per_query hits; float32 [n][2][dim]."""
    rng = np.random.default_rng(seed)
    c = rng.random((n, dim), dtype=np.float32)
    h = cube_half_width(n_points, hits_per_query, dim)
    q = np.empty((n, 2, dim), np.float32)
    q[:, 0, :] = c - h
    q[:, 1, :] = c + h
    return q


# ---- config 4: 8-D clustered Gaussian mixture -------------------------------------------------
def mixture_points(n, dim=8, clusters=256, sigma=0.02, seed=45):
    rng = np.random.default_rng(seed)
    centres = rng.random((clusters, dim))
    which = rng.integers(0, clusters, n)
    pts = centres[which] + rng.normal(0.0, sigma, (n, dim))
    return pts.astype(np.float32)


def mixture_queries(n, points, half_width, sigma=0.02, seed=46):
    """centre = random data point + N(0, sigma/4) jitter; float32 [n][2][dim]."""
    rng = np.random.default_rng(seed)
    dim = points.shape[1]
    c = points[rng.integers(0, len(points), n)] + rng.normal(0.0, sigma / 4.0, (n, dim)).astype(np.float32)
    c = c.astype(np.float32)
    h = np.float32(half_width)
    q = np.empty((n, 2, dim), np.float32)
    q[:, 0, :] = c - h
    q[:, 1, :] = c + h
    return q


def calibrate_half_width(points, target_hits=10.0, sample=2000, sigma=0.02, seed=47, iters=24):
    """bisection on a query sample (brute force on a point sample) for ~target_hits hits/query."""
    rng = np.random.default_rng(seed)
    dim = points.shape[1]
    sub = points[rng.integers(0, len(points), min(len(points), 200000))]
    scale = len(points) / float(len(sub))
    c = points[rng.integers(0, len(points), sample)] + rng.normal(0.0, sigma / 4.0, (sample, dim)).astype(np.float32)
    lo, hi = 0.0, 0.5
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        hits = 0
        for i in range(0, sample, 64):
            d = np.abs(sub[None, :, :] - c[i:i + 64, None, :]).max(axis=2)
            hits += int((d <= mid).sum())
        if hits * scale / sample > target_hits:
            hi = mid
        else:
            lo = mid
    return np.float32(0.5 * (lo + hi))


# ---- config 3: a teapot-like triangle mesh and random-direction rays -------------------------
def _revolve(profile, n_theta):
    """triangles of the surface of revolution of a (r, z) polyline around the z axis."""
    profile = np.asarray(profile, np.float64)
    th = np.linspace(0.0, 2.0 * np.pi, n_theta + 1)
    r, z = profile[:, 0], profile[:, 1]
    x = r[:, None] * np.cos(th)[None, :]
    y = r[:, None] * np.sin(th)[None, :]
    zz = np.repeat(z[:, None], n_theta + 1, axis=1)
    p = np.stack([x, y, zz], axis=2)  # [n_prof][n_theta+1][3]
    a, b, c, d = p[:-1, :-1], p[1:, :-1], p[1:, 1:], p[:-1, 1:]
    tris = np.concatenate([np.stack([a, b, c], axis=2), np.stack([a, c, d], axis=2)], axis=0)
    return tris.reshape(-1, 3, 3)


def _tube(centres, radii, n_ring):
    """triangles of a tube swept along a polyline (in the x-z plane) with per-point radii."""
    centres = np.asarray(centres, np.float64)
    radii = np.asarray(radii, np.float64)
    tang = np.gradient(centres, axis=0)
    tang /= np.linalg.norm(tang, axis=1, keepdims=True)
    side = np.array([0.0, 1.0, 0.0])
    nrm = np.cross(tang, side)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    th = np.linspace(0.0, 2.0 * np.pi, n_ring + 1)
    ring = (nrm[:, None, :] * np.cos(th)[None, :, None] + side[None, None, :] * np.sin(th)[None, :, None])
    p = centres[:, None, :] + radii[:, None, None] * ring
    a, b, c, d = p[:-1, :-1], p[1:, :-1], p[1:, 1:], p[:-1, 1:]
    tris = np.concatenate([np.stack([a, b, c], axis=2), np.stack([a, c, d], axis=2)], axis=0)
    return tris.reshape(-1, 3, 3)


def _bezier(ctrl, n):
    ctrl = np.asarray(ctrl, np.float64)
    t = np.linspace(0.0, 1.0, n)[:, None]
    return ((1 - t) ** 3 * ctrl[0] + 3 * (1 - t) ** 2 * t * ctrl[1] + 3 * (1 - t) * t ** 2 * ctrl[2]
            + t ** 3 * ctrl[3])


def teapot_triangles(res=1.0):
    """A deterministic teapot-like mesh (lathe body + lid + torus handle + tapered spout).

    The reference ships neither a mesh nor ray code (only a screenshot), so the shape is ours;
    res=1 gives ~6.3 k triangles like the classic tessellation, res~7.1 gives ~320 k.
    Returns float32 [n_tri][3 vertices][3].
    """
    k = max(1, int(round(8 * res)))
    n_theta = 6 * k
    body_prof = np.concatenate([
        _bezier([(0.0, 0.0), (0.75, 0.0), (1.5, 0.05), (1.5, 0.15)], 2 * k)[:-1],
        _bezier([(1.5, 0.15), (2.0, 0.45), (2.0, 1.35), (1.5, 2.25)], 4 * k)[:-1],
        _bezier([(1.5, 2.25), (1.44, 2.38), (1.34, 2.38), (1.4, 2.4)], k + 1)])
    lid_prof = np.concatenate([
        _bezier([(1.4, 2.4), (1.3, 2.55), (0.4, 2.55), (0.2, 2.7)], 2 * k)[:-1],
        _bezier([(0.2, 2.7), (0.0, 2.85), (0.45, 2.95), (0.0, 3.15)], 2 * k)])
    parts = [_revolve(body_prof, n_theta), _revolve(lid_prof, n_theta)]
    handle = _bezier([(-1.6, 0.0, 2.025), (-3.1, 0.0, 2.1), (-3.0, 0.0, 0.9), (-1.9, 0.0, 0.6)], 3 * k + 1)
    parts.append(_tube(handle, np.full(len(handle), 0.15), max(6, 2 * k)))
    spout = _bezier([(1.7, 0.0, 0.9), (2.7, 0.0, 0.9), (2.3, 0.0, 2.1), (3.3, 0.0, 2.4)], 3 * k + 1)
    parts.append(_tube(spout, np.linspace(0.33, 0.12, len(spout)), max(6, 2 * k)))
    return np.concatenate(parts, axis=0).astype(np.float32)


def triangle_aabbs(tris):
    """[n][2][3] float32 boxes (min corner, max corner) of triangles [n][3][3]."""
    return np.stack([tris.min(axis=1), tris.max(axis=1)], axis=1).astype(np.float32)


def random_rays(n, boxes, seed=44, tmax=np.inf, threads=None):
    """Rays [n][7] = origin, 1/direction, tmax.

    Origins are uniform on the sphere of 3x the scene's bounding radius; directions are
    uniformly random unit vectors conditioned on hitting the scene's bounding sphere (drawn
    directly inside the subtended cone, which is the same distribution as re-drawing until hit).
    Generated in 1 Mi-ray chunks, chunk c from its own generator seeded (seed, c), so the result
    does not depend on how many threads filled it.
    """
    lo = boxes[:, 0, :].min(axis=0).astype(np.float64)
    hi = boxes[:, 1, :].max(axis=0).astype(np.float64)
    centre = 0.5 * (lo + hi)
    radius = 0.5 * float(np.linalg.norm(hi - lo))
    out = np.empty((n, 7), np.float32)
    step = 1 << 20

    def fill(c):
        s = c * step
        m = min(step, n - s)
        rng = np.random.default_rng([seed, c])
        v = rng.normal(size=(m, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        origin = centre + 3.0 * radius * v
        axis = -v
        cos_max = np.sqrt(1.0 - (1.0 / 3.0) ** 2)
        cos_t = rng.uniform(cos_max, 1.0, m)
        sin_t = np.sqrt(np.maximum(0.0, 1.0 - cos_t ** 2))
        phi = rng.uniform(0.0, 2.0 * np.pi, m)
        helper = np.where(np.abs(axis[:, :1]) < 0.9, [[1.0, 0.0, 0.0]], [[0.0, 1.0, 0.0]])
        u = np.cross(axis, helper)
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        w = np.cross(axis, u)
        d = (axis * cos_t[:, None] + u * (sin_t * np.cos(phi))[:, None] + w * (sin_t * np.sin(phi))[:, None])
        d32 = d.astype(np.float32)
        out[s:s + m, 0:3] = origin.astype(np.float32)
        with np.errstate(divide="ignore"):
            out[s:s + m, 3:6] = np.float32(1.0) / d32
        out[s:s + m, 6] = np.float32(tmax)

    chunks = range((n + step - 1) // step)
    if threads is None:
        import os
        threads = min(16, os.cpu_count() or 1)
    if threads > 1 and len(chunks) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as pool:
            list(pool.map(fill, chunks))
    else:
        for c in chunks:
            fill(c)
    return out
