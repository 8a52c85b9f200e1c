"""Synthetic scene generators for the BASELINE.json configs.

Pure numpy; every generator returns (verts float32 [nv,3], idx uint32 [nt,3]) plus, for the path-traced
scenes, per-triangle material / light tables.  Vertices are computed in float64 and rounded once to float32.
These are inputs for BOTH the CUDA path and the CPU oracle — they contain no intersection logic.
"""
import numpy as np


def uv_sphere(radius=1.0, center=(0.0, 0.0, 0.0), n_theta=158, n_phi=316):
    """Lat-long sphere: two pole fans + (n_theta-2) quad bands -> 2*n_phi*(n_theta-1) triangles."""
    c = np.asarray(center, dtype=np.float64)
    verts = [c + np.array([0.0, radius, 0.0])]
    for i in range(1, n_theta):
        th = np.pi * i / n_theta
        ph = 2.0 * np.pi * np.arange(n_phi) / n_phi
        ring = np.stack([radius * np.sin(th) * np.cos(ph), np.full(n_phi, radius * np.cos(th)),
                         radius * np.sin(th) * np.sin(ph)], axis=1) + c
        verts.extend(ring)
    verts.append(c + np.array([0.0, -radius, 0.0]))
    verts = np.asarray(verts, dtype=np.float64)
    tris = []
    j = np.arange(n_phi)
    jn = (j + 1) % n_phi
    first = 1
    tris.append(np.stack([np.zeros(n_phi, dtype=np.int64), first + jn, first + j], axis=1))
    for i in range(n_theta - 2):
        a = 1 + i * n_phi
        b = a + n_phi
        tris.append(np.stack([a + j, a + jn, b + j], axis=1))
        tris.append(np.stack([a + jn, b + jn, b + j], axis=1))
    last = 1 + (n_theta - 2) * n_phi
    south = len(verts) - 1
    tris.append(np.stack([last + j, last + jn, np.full(n_phi, south)], axis=1))
    idx = np.concatenate(tris, axis=0)
    return verts.astype(np.float32), idx.astype(np.uint32)


def ground_grid(y=-1.0, lo=-10.0, hi=10.0, n=20):
    xs = np.linspace(lo, hi, n + 1)
    X, Z = np.meshgrid(xs, xs, indexing="xy")
    verts = np.stack([X.ravel(), np.full(X.size, y), Z.ravel()], axis=1)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    v00 = (j * (n + 1) + i).ravel()
    v10 = v00 + 1
    v01 = v00 + (n + 1)
    v11 = v01 + 1
    idx = np.concatenate([np.stack([v00, v01, v10], axis=1), np.stack([v10, v01, v11], axis=1)], axis=0)
    return verts.astype(np.float32), idx.astype(np.uint32)


def merge(*meshes):
    verts, idx, off = [], [], 0
    for v, i in meshes:
        verts.append(v)
        idx.append(i.astype(np.int64) + off)
        off += len(v)
    return np.concatenate(verts).astype(np.float32), np.concatenate(idx).astype(np.uint32)


def scene_c1():
    """C1: 99,224-triangle UV sphere + 800-triangle ground = 100,024 triangles."""
    return merge(uv_sphere(), ground_grid())


C1_CAMERA = dict(pos=(0.0, 1.0, -4.0), look=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=45.0, res=(1024, 1024))
C3_CAMERA = dict(pos=(0.0, 30.0, -80.0), look=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=45.0, res=(1024, 1024))
C3_POINT_LIGHT = (0.0, 200.0, 0.0)


def _hash2(ix, iz, seed):
    """Integer lattice hash -> [0,1) (uint64 arithmetic, wraps)."""
    h = (ix.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ (iz.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F))
    h ^= np.uint64(seed) * np.uint64(0x165667B19E3779F9)
    h ^= h >> np.uint64(29)
    h *= np.uint64(0xBF58476D1CE4E5B9)
    h ^= h >> np.uint64(32)
    h *= np.uint64(0x94D049BB133111EB)
    h ^= h >> np.uint64(29)
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def value_noise(x, z, seed):
    x0 = np.floor(x)
    z0 = np.floor(z)
    fx = x - x0
    fz = z - z0
    ix = x0.astype(np.int64)
    iz = z0.astype(np.int64)
    sx = fx * fx * (3.0 - 2.0 * fx)
    sz = fz * fz * (3.0 - 2.0 * fz)
    with np.errstate(over="ignore"):
        v00 = _hash2(ix, iz, seed)
        v10 = _hash2(ix + 1, iz, seed)
        v01 = _hash2(ix, iz + 1, seed)
        v11 = _hash2(ix + 1, iz + 1, seed)
    a = v00 + (v10 - v00) * sx
    b = v01 + (v11 - v01) * sx
    return a + (b - a) * sz


def fbm(x, z, seed=1234, octaves=5):
    # base frequency 0.35 -> wavelengths 2.9 .. 0.18 units (the finest octave spans ~4 grid quads at n=2237): rugged
    # enough that ~13% of the cosine-bounce rays re-hit the terrain instead of all escaping
    amp, freq, total = 0.5, 0.35, np.zeros_like(x)
    for o in range(octaves):
        total += amp * (2.0 * value_noise(x * freq, z * freq, seed + o) - 1.0)
        amp *= 0.5
        freq *= 2.0
    return total


def displaced_grid(n=2237, lo=-50.0, hi=50.0, height=2.0, seed=1234):
    """C3: n x n quads on [lo,hi]^2, y = height * fbm(x,z).  n=2237 -> 10,008,338 triangles."""
    xs = np.linspace(lo, hi, n + 1)
    X, Z = np.meshgrid(xs, xs, indexing="xy")
    Y = height * fbm(X, Z, seed=seed)
    verts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)
    i, j = np.meshgrid(np.arange(n, dtype=np.int64), np.arange(n, dtype=np.int64), indexing="xy")
    v00 = (j * (n + 1) + i).ravel()
    v10 = v00 + 1
    v01 = v00 + (n + 1)
    v11 = v01 + 1
    idx = np.empty((2 * n * n, 3), dtype=np.uint32)
    idx[0::2] = np.stack([v00, v01, v10], axis=1)
    idx[1::2] = np.stack([v10, v01, v11], axis=1)
    return verts, idx


def scene_c3(n=2237):
    return displaced_grid(n=n)


def random_soup(n_tris, seed=0, extent=10.0, size=1.0):
    """Random triangle soup for property tests."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, size=(n_tris, 1, 3))
    v = (c + rng.uniform(-size, size, size=(n_tris, 3, 3))).astype(np.float32).reshape(-1, 3)
    idx = np.arange(3 * n_tris, dtype=np.uint32).reshape(-1, 3)
    return v, idx
