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


# ---------------------------------------------------------------------------------------------------------------
# Path-traced scenes (BASELINE configs 1, 3, 4).  Materials / lights are plain dicts so that both the CUDA binding and
# the oracle binding can build their own structs from them.
# ---------------------------------------------------------------------------------------------------------------
def _quad(a, b, c, d):
    """Quad a,b,c,d -> two triangles (a,b,c), (a,c,d); geometric normal = normalize((a-c) x (b-c))."""
    return np.array([a, b, c, d], dtype=np.float64), np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int64)


def _append(meshes, mats, quad, mat):
    meshes.append((quad[0].astype(np.float32), quad[1].astype(np.uint32)))
    mats.extend([mat, mat])


WHITE, RED, GREEN = (0.725, 0.71, 0.68), (0.63, 0.065, 0.05), (0.14, 0.45, 0.091)


def cornell_box(light_y=548.3, blocks=True):
    """Classic Cornell box data (555-unit room): 5 walls + ceiling light + short and tall block = 32 triangles."""
    meshes, tm = [], []
    W, R, G = 0, 1, 2
    _append(meshes, tm, _quad((552.8, 0, 0), (0, 0, 0), (0, 0, 559.2), (549.6, 0, 559.2)), W)                    # floor
    _append(meshes, tm, _quad((556, 548.8, 0), (556, 548.8, 559.2), (0, 548.8, 559.2), (0, 548.8, 0)), W)       # ceiling
    _append(meshes, tm, _quad((549.6, 0, 559.2), (0, 0, 559.2), (0, 548.8, 559.2), (556, 548.8, 559.2)), W)     # back
    _append(meshes, tm, _quad((0, 0, 559.2), (0, 0, 0), (0, 548.8, 0), (0, 548.8, 559.2)), G)                    # right
    _append(meshes, tm, _quad((552.8, 0, 0), (549.6, 0, 559.2), (556, 548.8, 559.2), (556, 548.8, 0)), R)       # left
    light_first = 2 * len(meshes)
    _append(meshes, tm, _quad((343, light_y, 227), (343, light_y, 332), (213, light_y, 332), (213, light_y, 227)), W)
    if blocks:
        short = [((130, 165, 65), (82, 165, 225), (240, 165, 272), (290, 165, 114)),
                 ((290, 0, 114), (290, 165, 114), (240, 165, 272), (240, 0, 272)),
                 ((130, 0, 65), (130, 165, 65), (290, 165, 114), (290, 0, 114)),
                 ((82, 0, 225), (82, 165, 225), (130, 165, 65), (130, 0, 65)),
                 ((240, 0, 272), (240, 165, 272), (82, 165, 225), (82, 0, 225))]
        tall = [((423, 330, 247), (265, 330, 296), (314, 330, 456), (472, 330, 406)),
                ((423, 0, 247), (423, 330, 247), (472, 330, 406), (472, 0, 406)),
                ((472, 0, 406), (472, 330, 406), (314, 330, 456), (314, 0, 456)),
                ((314, 0, 456), (314, 330, 456), (265, 330, 296), (265, 0, 296)),
                ((265, 0, 296), (265, 330, 296), (423, 330, 247), (423, 0, 247))]
        for q in short + tall:
            _append(meshes, tm, _quad(*q), W)
    verts, idx = merge(*meshes)
    materials = [dict(type="matte", kd=WHITE), dict(type="matte", kd=RED), dict(type="matte", kd=GREEN)]
    lights = [dict(type="area", prim=light_first, L=(17.0, 12.0, 4.0), two_sided=False),
              dict(type="area", prim=light_first + 1, L=(17.0, 12.0, 4.0), two_sided=False)]
    return dict(verts=verts, idx=idx, tri_material=np.array(tm, dtype=np.uint32), materials=materials, lights=lights)


C2_CAMERA = dict(pos=(278.0, 273.0, -800.0), look=(278.0, 273.0, 0.0), up=(0.0, 1.0, 0.0), fov=39.3, res=(512, 512))
C2_PATH = dict(max_depth=5, rr_threshold=1.0, light_strategy="uniform", spp=64)


def scene_c2():
    """C2: Cornell box, matte walls, diffuse quad light, PathIntegrator maxdepth=5, 512x512 @ 64 spp."""
    return cornell_box()


def scene_c4(n_theta=158, n_phi=316):
    """C4: Cornell room + three tessellated spheres (matte / plastic / glass), ceiling area light + point light."""
    room = cornell_box(blocks=False)
    meshes = [(room["verts"], room["idx"])]
    tm = list(room["tri_material"])
    for k, (cx, mat) in enumerate(((140.0, 3), (278.0, 4), (416.0, 5))):
        v, i = uv_sphere(radius=90.0, center=(cx, 90.0 + 40.0 * (k == 1), 280.0 - 60.0 * (k == 1)), n_theta=n_theta, n_phi=n_phi)
        meshes.append((v, i))
        tm.extend([mat] * len(i))
    verts, idx = merge(*meshes)
    materials = room["materials"] + [dict(type="matte", kd=(0.5, 0.5, 0.8)),
                                     dict(type="plastic", kd=(0.25, 0.25, 0.25), ks=(0.25, 0.25, 0.25), roughness=0.1, remap=True),
                                     dict(type="glass", kr=(1.0, 1.0, 1.0), kt=(1.0, 1.0, 1.0), eta=1.5)]
    lights = room["lights"] + [dict(type="point", p=(278.0, 400.0, 100.0), I=(40000.0, 40000.0, 40000.0))]
    return dict(verts=verts, idx=idx, tri_material=np.array(tm, dtype=np.uint32), materials=materials, lights=lights)


def scene_c4_smooth(n_theta=24, n_phi=48, uvs=True, tangents=False, emissive_normals=True):
    """C4's room as a TriangleMesh WITH per-vertex attributes (src/shapes/triangle.rs:17-26): the three spheres get their analytic
    normals (smooth shading; the glass one too), the room's flat faces their face normals — the ceiling light's tilted and, on one
    vertex, flipped, so Triangle::sample's face_forward matters —, UVs from the vertex positions and, optionally, tangents."""
    sc = scene_c4(n_theta=n_theta, n_phi=n_phi)
    v, idx = sc["verts"].astype(np.float64), sc["idx"]
    nrm = np.zeros_like(v)
    fn = np.cross(v[idx[:, 1]] - v[idx[:, 0]], v[idx[:, 2]] - v[idx[:, 0]])
    for k in range(3):
        np.add.at(nrm, idx[:, k], fn)
    centers = [((140.0, 90.0, 280.0), 3), ((278.0, 130.0, 220.0), 4), ((416.0, 90.0, 280.0), 5)]
    for c, mat in centers:
        used = np.unique(idx[sc["tri_material"] == mat])
        nrm[used] = v[used] - np.array(c)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)
    if emissive_normals:
        for l in sc["lights"]:
            if l["type"] == "area":
                nrm[idx[l["prim"]]] += np.array((0.3, 0.0, 0.2))
        first = idx[[l["prim"] for l in sc["lights"] if l["type"] == "area"][0]][0]
        nrm[first] = -nrm[first]
    sc["normals"] = nrm.astype(np.float32)
    if uvs:
        sc["uvs"] = np.stack([v[:, 0] / 556.0 + 0.37 * v[:, 1] / 556.0, v[:, 2] / 559.2 - 0.21 * v[:, 1] / 556.0], axis=1).astype(np.float32)
    if tangents:
        t = np.cross(nrm, np.array((0.0, 1.0, 0.0)))
        t[np.linalg.norm(t, axis=1) < 1e-6] = (1.0, 0.0, 0.0)
        sc["tangents"] = t.astype(np.float32)
    return sc


def scene_materials(n_theta=32, n_phi=64):
    """C4's room with every material the backend knows: Oren-Nayar matte (sigma = 35), metal (copper-like conductor), mirror on
    the three big spheres, glass and plastic on two small ones in front, Lambertian walls."""
    sc = scene_c4(n_theta=n_theta, n_phi=n_phi)
    meshes = [(sc["verts"], sc["idx"])]
    tm = list(sc["tri_material"])
    for (cx, cz), mat in (((200.0, 120.0), 6), ((350.0, 120.0), 7)):
        v, i = uv_sphere(radius=45.0, center=(cx, 45.0, cz), n_theta=max(8, n_theta // 2), n_phi=max(16, n_phi // 2))
        meshes.append((v, i))
        tm.extend([mat] * len(i))
    verts, idx = merge(*meshes)
    mats = list(sc["materials"])
    mats[3] = dict(type="matte", kd=(0.6, 0.6, 0.7), sigma=35.0)
    mats[4] = dict(type="metal", metal_eta=(0.2, 0.92, 1.1), metal_k=(3.9, 2.45, 2.14), roughness=0.05, remap=True)
    mats[5] = dict(type="mirror", kr=(0.9, 0.9, 0.9))
    mats += [dict(type="glass", kr=(1.0, 1.0, 1.0), kt=(1.0, 1.0, 1.0), eta=1.5, roughness=0.2, remap=True),      # frosted
             dict(type="plastic", kd=(0.25, 0.35, 0.25), ks=(0.3, 0.3, 0.3), roughness=0.1, remap=True)]
    mats[2] = dict(type="substrate", kd=(0.14, 0.45, 0.091), ks=(0.04, 0.04, 0.04), roughness=0.05, remap=True)   # the green wall, lacquered
    mats.append(dict(type="glass", kr=(1.0, 1.0, 1.0), kt=(1.0, 1.0, 1.0), eta=1.5))                                 # smooth
    tm = np.array(tm, dtype=np.uint32)
    tm[tm == 7] = np.where(np.arange((tm == 7).sum()) % 2 == 0, 7, 8)          # half of the plastic ball's triangles are smooth glass
    tm = list(tm)
    return dict(verts=verts, idx=idx, tri_material=np.array(tm, dtype=np.uint32), materials=mats, lights=sc["lights"])


def scene_all_lights(n_theta=24, n_phi=48):
    """C4's room with every light the backend knows: the ceiling area light, a point light, a spot light aimed at the matte
    sphere (src/lights/spot.rs; axis = normalize(to - from), 30 degree cone, falloff from 20 degrees) and a distant light
    shining in through the open front (src/lights/distant.rs)."""
    sc = scene_c4(n_theta=n_theta, n_phi=n_phi)
    frm, to = np.array((450.0, 500.0, 60.0), np.float32), np.array((140.0, 90.0, 280.0), np.float32)
    d = to - frm
    axis = d / np.float32(np.sqrt(np.float32(np.dot(d, d))))
    sc["lights"] = sc["lights"] + [dict(type="spot", p=tuple(frm), axis=tuple(float(a) for a in axis), I=(9e5, 8e5, 6e5), total_width=30.0, falloff_start=20.0),
                                   dict(type="distant", w=(0.2, 0.3, -1.0), L=(1.5, 1.5, 2.0))]
    return sc


C4_CAMERA = dict(pos=(278.0, 273.0, -800.0), look=(278.0, 273.0, 0.0), up=(0.0, 1.0, 0.0), fov=39.3, res=(1920, 1080))
C4_PATH = dict(max_depth=8, rr_threshold=1.0, light_strategy="power", spp=256)
C5_CAMERA = dict(C4_CAMERA, res=(3840, 2160))
C5_PATH = dict(C4_PATH, spp=1024)


def furnace_box(L=0.5, kd=0.5):
    """White furnace: closed matte cube whose six walls all emit L (two-sided) -> radiance L / (1 - kd) everywhere."""
    c = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]
    faces = [(0, 1, 2, 3), (5, 4, 7, 6), (4, 0, 3, 7), (1, 5, 6, 2), (3, 2, 6, 7), (4, 5, 1, 0)]
    meshes, tm = [], []
    for f in faces:
        _append(meshes, tm, _quad(*[c[k] for k in f]), 0)
    verts, idx = merge(*meshes)
    lights = [dict(type="area", prim=k, L=(L, L, L), two_sided=True) for k in range(len(idx))]
    return dict(verts=verts, idx=idx, tri_material=np.array(tm, dtype=np.uint32), materials=[dict(type="matte", kd=(kd, kd, kd))],
                lights=lights)


def _rot_scale(center, axis_angle_deg=(0.0, 0.0, 1.0, 0.0), scale=(1.0, 1.0, 1.0)):
    """Row-major affine object_to_world = translate(center) * rotate(axis, angle) * scale(s) as float32."""
    ax = np.asarray(axis_angle_deg[:3], dtype=np.float64)
    ang = np.radians(axis_angle_deg[3])
    m = np.eye(4)
    if np.linalg.norm(ax) > 0 and ang != 0.0:
        ax = ax / np.linalg.norm(ax)
        k = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        m[:3, :3] = np.eye(3) + np.sin(ang) * k + (1 - np.cos(ang)) * (k @ k)
    m[:3, :3] = m[:3, :3] @ np.diag(scale)
    m[:3, 3] = center
    return m.astype(np.float32)


def scene_spheres(sphere_light=True, partial=True):
    """Cornell room (no blocks) with ANALYTIC spheres (src/shapes/sphere.rs) instead of tessellated ones: matte, plastic and glass
    balls, an ellipsoid (rotated, non-uniformly scaled sphere), a partial sphere (z_min / z_max / phi_max, reversed orientation)
    and — besides the ceiling quad light and a point light — a spherical DiffuseAreaLight.  Sphere primitive ids follow the
    triangles': len(idx) + k."""
    room = cornell_box(blocks=False)
    nt = len(room["idx"])
    materials = room["materials"] + [dict(type="matte", kd=(0.5, 0.5, 0.8)),
                                     dict(type="plastic", kd=(0.25, 0.25, 0.25), ks=(0.25, 0.25, 0.25), roughness=0.1, remap=True),
                                     dict(type="glass", kr=(1.0, 1.0, 1.0), kt=(1.0, 1.0, 1.0), eta=1.5)]
    spheres = [dict(center=(140.0, 90.0, 300.0), radius=90.0, material=3),
               dict(center=(278.0, 130.0, 220.0), radius=90.0, material=4),
               dict(center=(430.0, 90.0, 280.0), radius=90.0, material=5),
               dict(o2w=_rot_scale((300.0, 330.0, 420.0), (0.3, 1.0, 0.2, 35.0), (1.6, 0.7, 1.0)), radius=60.0, material=1)]
    if partial:
        spheres.append(dict(o2w=_rot_scale((120.0, 300.0, 380.0), (1.0, 0.0, 0.0, -70.0)), radius=70.0, z_min=-40.0, z_max=55.0, phi_max=250.0,
                            reverse_orientation=True, material=2))
    lights = room["lights"] + [dict(type="point", p=(278.0, 400.0, 100.0), I=(20000.0, 20000.0, 20000.0))]
    if sphere_light:
        spheres.append(dict(center=(460.0, 420.0, 150.0), radius=25.0, material=0))
        lights.append(dict(type="area", prim=nt + len(spheres) - 1, L=(60.0, 55.0, 40.0), two_sided=False))
    return dict(verts=room["verts"], idx=room["idx"], tri_material=room["tri_material"], materials=materials, lights=lights, spheres=spheres)


NO_MATERIAL = 0xFFFFFFFF


def _box(lo, hi):
    """Axis-aligned box as 12 triangles with outward normals."""
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    quads = [((x0, y0, z0), (x0, y1, z0), (x1, y1, z0), (x1, y0, z0)),      # z = z0, normal -z
             ((x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)),      # z = z1, normal +z
             ((x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0)),      # x = x0, normal -x
             ((x1, y0, z0), (x1, y1, z0), (x1, y1, z1), (x1, y0, z1)),      # x = x1, normal +x
             ((x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1)),      # y = y0, normal -y
             ((x0, y1, z0), (x0, y1, z1), (x1, y1, z1), (x1, y1, z0))]      # y = y1, normal +y
    return merge(*[_quad(*q) for q in quads])


def scene_media(fog_everywhere=True, g=0.3):
    """Cornell room for the VolPathIntegrator (src/integrators/volpath.rs, src/media/homogeneous.rs): the camera sits in a thin
    homogeneous fog that fills the room (medium 0), a material-less box in the middle encloses a dense, forward-scattering
    coloured smoke (medium 1: the box surface only separates the media), beside a matte block, a glass ball (analytic sphere) and
    the quad area light + a point light."""
    room = cornell_box(blocks=False)
    meshes = [(room["verts"], room["idx"])]
    tm = list(room["tri_material"])
    bv, bi = _box((90.0, 60.0, 200.0), (250.0, 260.0, 360.0))
    meshes.append((bv, bi))
    n_room = len(room["idx"])
    tm.extend([NO_MATERIAL] * len(bi))
    sv, si = _box((330.0, 0.0, 300.0), (450.0, 150.0, 420.0))
    meshes.append((sv, si))
    tm.extend([0] * len(si))
    verts, idx = merge(*meshes)
    materials = room["materials"] + [dict(type="glass", kr=(1.0, 1.0, 1.0), kt=(1.0, 1.0, 1.0), eta=1.5)]
    spheres = [dict(center=(400.0, 210.0, 360.0), radius=60.0, material=3)]
    fog = 0 if fog_everywhere else -1
    n_prims = len(idx) + len(spheres)
    inside = np.full(n_prims, fog, dtype=np.int32)
    outside = np.full(n_prims, fog, dtype=np.int32)
    inside[n_room:n_room + len(bi)] = 1                 # the smoke box: medium 1 inside, the fog outside
    media = [dict(sigma_a=(0.0002, 0.0002, 0.0003), sigma_s=(0.0012, 0.0012, 0.0012), g=0.0),
             dict(sigma_a=(0.002, 0.004, 0.008), sigma_s=(0.02, 0.018, 0.012), g=g)]
    lights = room["lights"] + [dict(type="point", p=(278.0, 400.0, 100.0), I=(20000.0, 20000.0, 20000.0))]
    return dict(verts=verts, idx=idx, tri_material=np.array(tm, dtype=np.uint32), materials=materials, lights=lights, spheres=spheres,
                media=media, prim_inside=inside, prim_outside=outside, camera_medium=fog)
