"""Synthetic scenes and ray sets for the configs named in BASELINE.json.

Everything is counter based: value = splitmix64(seed ^ index), so any slice of a workload can be
regenerated anywhere (tests on CPU, bench on the GPU box) without storing it.  These are *inputs*:
the same arrays are handed to the CPU oracle and to the CUDA path, so nothing here needs to match
the reference bit for bit except the teapot triangles (file order of objects/teapot.obj, fixture
tests/golden/teapot_tris.npy) and the benchmark camera (examples/benchmark.rs:74-142).
"""
from __future__ import annotations

import os
import numpy as np

RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("t_min", "<f4"), ("direction", "<f4", 3), ("t", "<f4")])
PACKET_DTYPE = np.dtype(
    [("origin_x", "<f4", 4), ("origin_y", "<f4", 4), ("origin_z", "<f4", 4), ("direction_x", "<f4", 4),
     ("direction_y", "<f4", 4), ("direction_z", "<f4", 4), ("t", "<f4", 4)]
)
T_MIN = np.float32(1e-4)   # Ray::DEFAULT_T_MIN, src/ray.rs:163
T_MAX = np.float32(1e34)   # Ray::DEFAULT_T_MAX, src/ray.rs:164

SEED_TEAPOT_RANDOM = 0x7EA907
SEED_SOUP = 0x50A90002
SEED_MESH = 0x3E510003
SEED_SCENE = 0x5CE40004

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def hash_unit(seed: int, index: np.ndarray, lane: int) -> np.ndarray:
    """Uniform float32 in [0,1) from (seed, index, lane)."""
    with np.errstate(over="ignore"):
        key = np.uint64(seed) ^ (np.asarray(index, dtype=np.uint64) * np.uint64(16) + np.uint64(lane))
    return ((splitmix64(key) >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# geometry
# ------------------------------------------------------------------------------------------------
def teapot(path: str | None = None) -> np.ndarray:
    """[6320,3,3] float32: objects/teapot.obj triangles in file order (config 1)."""
    return np.load(path or os.path.join(_GOLDEN, "teapot_tris.npy"))


def quad() -> np.ndarray:
    """The 2-triangle quad at z=1 of the FFI `intersect` KAT (rtbvh_ffi/src/lib.rs:947-954)."""
    v = np.array([[-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float32)
    return v.reshape(2, 3, 3)


def soup(n: int = 1 << 20, seed: int = SEED_SOUP, aniso=(1.0, 1.0, 1.0)) -> np.ndarray:
    """Config 2: n triangles, centre uniform in [0,1]^3, vertices = centre + uniform offsets in [-s,s]^3,
    s = 0.5 * n^(-1/3).  `aniso` stretches the offsets per axis (config 5 uses (8,1,1))."""
    idx = np.arange(n, dtype=np.uint64)
    s = np.float32(0.5 * n ** (-1.0 / 3.0))
    c = np.stack([hash_unit(seed, idx, k) for k in range(3)], axis=1)
    out = np.empty((n, 3, 3), dtype=np.float32)
    for v in range(3):
        for k in range(3):
            off = (hash_unit(seed, idx, 3 + v * 3 + k) * np.float32(2.0) - np.float32(1.0)) * s * np.float32(aniso[k])
            out[:, v, k] = c[:, k] + off
    return out


def heightfield(nx: int = 2237, ny: int = 2237, seed: int = SEED_MESH) -> np.ndarray:
    """Config 3: (nx x ny) quads x 2 triangles, z = sum of 4 sines + 0.02 * hash noise, shared vertices."""
    gx, gy = np.meshgrid(np.arange(nx + 1, dtype=np.float32), np.arange(ny + 1, dtype=np.float32), indexing="xy")
    x = gx / np.float32(nx)
    y = gy / np.float32(ny)
    z = (0.10 * np.sin(6.1 * x) + 0.07 * np.sin(9.7 * y + 0.5) + 0.04 * np.sin(23.0 * (x + y)) +
         0.02 * np.sin(41.0 * (x - y))).astype(np.float32)
    vid = (gy.astype(np.uint64) * np.uint64(nx + 1) + gx.astype(np.uint64))
    z = z + np.float32(0.02) * (hash_unit(seed, vid, 0) - np.float32(0.5))
    p = np.stack([x, y, z], axis=-1).astype(np.float32)  # [ny+1, nx+1, 3]
    p00, p10, p01, p11 = p[:-1, :-1], p[:-1, 1:], p[1:, :-1], p[1:, 1:]
    t0 = np.stack([p00, p10, p11], axis=2)
    t1 = np.stack([p00, p11, p01], axis=2)
    return np.stack([t0, t1], axis=2).reshape(-1, 3, 3).astype(np.float32)


def displaced_sphere(nu: int = 708, nv: int = 707, seed: int = SEED_SCENE) -> np.ndarray:
    """A UV sphere of 2*nu*nv triangles (1 001 112 by default) with a smooth radial displacement."""
    f = np.float32
    u = (np.arange(nu + 1, dtype=f) / f(nu)) * f(2 * np.pi)
    v = (np.arange(nv + 1, dtype=f) / f(nv)) * f(np.pi)
    uu, vv = np.meshgrid(u, v, indexing="xy")
    r = (f(1.0) + f(0.05) * np.sin(f(9) * uu) * np.sin(f(7) * vv) + f(0.02) * np.sin(f(31) * vv + f(3) * uu)).astype(f)
    p = np.stack([r * np.sin(vv) * np.cos(uu), r * np.cos(vv), r * np.sin(vv) * np.sin(uu)], axis=-1).astype(f)
    p00, p10, p01, p11 = p[:-1, :-1], p[:-1, 1:], p[1:, :-1], p[1:, 1:]
    t0 = np.stack([p00, p10, p11], axis=2)
    t1 = np.stack([p00, p11, p01], axis=2)
    return np.stack([t0, t1], axis=2).reshape(-1, 3, 3).astype(f)


def instanced_scene(instances: int = 30, nu: int = 708, nv: int = 707, seed: int = SEED_SCENE) -> np.ndarray:
    """Config 4: `instances` baked copies (random rigid transforms + uniform scale) of a displaced sphere above a
    2-triangle ground and 4 walls (8 triangles).  30 x 1 001 112 + 10 = 30 033 370 triangles by default."""
    f = np.float32
    base = displaced_sphere(nu, nv, seed)
    out = np.empty((instances * len(base) + 10, 3, 3), dtype=f)
    side = int(np.ceil(instances ** (1.0 / 3.0)))
    for i in range(instances):
        h = [float(hash_unit(seed, np.array([i], dtype=np.uint64), k)[0]) for k in range(8)]
        # rotation from a random unit quaternion
        q = np.array([h[0] - 0.5, h[1] - 0.5, h[2] - 0.5, h[3] - 0.5], dtype=np.float64)
        q /= np.linalg.norm(q) + 1e-12
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=f)
        cell = np.array([i % side, (i // side) % side, i // (side * side)], dtype=f)
        t = (cell * f(2.6) + np.array([h[4], h[5], h[6]], dtype=f) * f(0.3) + f(1.3)).astype(f)
        sc = f(0.8 + 0.4 * h[7])
        out[i * len(base):(i + 1) * len(base)] = (base.reshape(-1, 3) @ R.T * sc + t).reshape(-1, 3, 3)
    ext = f(side * 2.6 + 0.5)
    lo, hi = f(-0.5), ext
    c = [np.array(p, dtype=f) for p in ((lo, lo, lo), (hi, lo, lo), (hi, lo, hi), (lo, lo, hi), (lo, hi, lo), (hi, hi, lo),
                                        (hi, hi, hi), (lo, hi, hi))]
    quads = [(0, 1, 2, 3), (0, 4, 5, 1), (1, 5, 6, 2), (2, 6, 7, 3), (3, 7, 4, 0)]  # ground + 4 walls, open top
    k = instances * len(base)
    for a, b, cc, d in quads:
        out[k] = np.stack([c[a], c[b], c[cc]])
        out[k + 1] = np.stack([c[a], c[cc], c[d]])
        k += 2
    return out


def bounds(verts: np.ndarray):
    v = verts.reshape(-1, 3)
    return v.min(axis=0), v.max(axis=0)


# ------------------------------------------------------------------------------------------------
# rays
# ------------------------------------------------------------------------------------------------
def _normalize(d: np.ndarray) -> np.ndarray:
    # glam Vec3::normalize = v * (1 / sqrt(v.v)), dot = (x*x + y*y) + z*z
    l2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    inv = np.float32(1.0) / np.sqrt(l2, dtype=np.float32)
    return (d * inv[..., None]).astype(np.float32)


def make_rays(origin: np.ndarray, direction: np.ndarray, t_min=T_MIN, t_max=T_MAX) -> np.ndarray:
    n = direction.shape[0]
    r = np.zeros(n, dtype=RAY_DTYPE)
    r["origin"] = origin
    r["direction"] = direction
    r["t_min"] = t_min
    r["t"] = t_max
    return r


def benchmark_camera(width: int = 1000, height: int = 1000):
    """The camera of examples/benchmark.rs:74-98 (note quirk Q12: the fov is converted to radians twice)."""
    f = np.float32
    fov = f(np.radians(f(90.0)))
    screen = f(np.tan(f(fov * f(0.5)) / f(f(180.0) / f(np.pi))))
    up = np.array([0, 1, 0], dtype=f)
    right = np.cross(np.array([0, 0, 1], dtype=f), up).astype(f)  # Z x Y = (-1, 0, 0)
    pos = np.array([0.0, 1.5, -100.0], dtype=f)
    center = pos + np.array([0, 0, 1], dtype=f)
    aspect = f(width) / f(height)
    p1 = center - screen * right * aspect + screen * up
    p2 = center + screen * right * aspect + screen * up
    p3 = center - screen * right * aspect - screen * up
    return dict(pos=pos, p1=p1.astype(f), right=(p2 - p1).astype(f), up=(p3 - p1).astype(f),
                inv_width=f(1.0) / f(width), inv_height=f(1.0) / f(height), width=width, height=height)


def camera_rays(cam: dict, y0: int = 0, y1: int | None = None, jitter_seed: int | None = None, frame: int = 0):
    """CameraView3D::generate_ray over rows [y0,y1) (shared/src/lib.rs:158-165); optional sub-pixel jitter."""
    f = np.float32
    w, h = cam["width"], cam["height"]
    y1 = h if y1 is None else y1
    xs, ys = np.meshgrid(np.arange(w, dtype=f), np.arange(y0, y1, dtype=f), indexing="xy")
    xs, ys = xs.reshape(-1), ys.reshape(-1)
    if jitter_seed is not None:
        pix = (np.uint64(frame) * np.uint64(w * h) + ys.astype(np.uint64) * np.uint64(w) + xs.astype(np.uint64))
        xs = xs + hash_unit(jitter_seed, pix, 0)
        ys = ys + hash_unit(jitter_seed, pix, 1)
    u = (xs * cam["inv_width"]).astype(f)
    v = (ys * cam["inv_height"]).astype(f)
    point = cam["p1"][None, :] + u[:, None] * cam["right"][None, :] + v[:, None] * cam["up"][None, :]
    d = _normalize((point - cam["pos"][None, :]).astype(f))
    return make_rays(np.broadcast_to(cam["pos"], d.shape), d)


def pinhole_camera(pos, look_at, fov_deg: float, width: int, height: int):
    """Generic pinhole with the same (p1, right, up) parameterisation as CameraView3D."""
    f = np.float32
    pos = np.asarray(pos, dtype=f)
    fwd = _normalize((np.asarray(look_at, dtype=f) - pos)[None, :])[0]
    upv = np.array([0, 1, 0], dtype=f)
    right = _normalize(np.cross(fwd, upv).astype(f)[None, :])[0]
    up = np.cross(right, fwd).astype(f)
    screen = f(np.tan(np.radians(fov_deg) * 0.5))
    aspect = f(width) / f(height)
    center = pos + fwd
    p1 = center - screen * right * aspect + screen * up
    p2 = center + screen * right * aspect + screen * up
    p3 = center - screen * right * aspect - screen * up
    return dict(pos=pos, p1=p1.astype(f), right=(p2 - p1).astype(f), up=(p3 - p1).astype(f),
                inv_width=f(1.0) / f(width), inv_height=f(1.0) / f(height), width=width, height=height)


def soup_camera(width: int = 1000, height: int = 1000):
    """Config 2 camera: from (0.5, 0.5, -1.5) at the unit cube, 50 degree fov."""
    return pinhole_camera((0.5, 0.5, -1.5), (0.5, 0.5, 0.5), 50.0, width, height)


def random_rays(n: int, lo, hi, seed: int = SEED_TEAPOT_RANDOM, first: int = 0) -> np.ndarray:
    """Incoherent set: origin on a sphere around the bounds, aimed at a uniform point inside the bounds."""
    f = np.float32
    lo, hi = np.asarray(lo, dtype=f), np.asarray(hi, dtype=f)
    c = (lo + hi) * f(0.5)
    rad = f(np.linalg.norm(hi - lo)) * f(1.5)
    idx = np.arange(first, first + n, dtype=np.uint64)
    z = hash_unit(seed, idx, 0) * f(2) - f(1)
    phi = hash_unit(seed, idx, 1) * f(2 * np.pi)
    rxy = np.sqrt(np.maximum(f(0), f(1) - z * z)).astype(f)
    o = c[None, :] + rad * np.stack([rxy * np.cos(phi), rxy * np.sin(phi), z], axis=1).astype(f)
    tgt = lo[None, :] + (hi - lo)[None, :] * np.stack([hash_unit(seed, idx, 2 + k) for k in range(3)], axis=1)
    d = _normalize((tgt - o).astype(f))
    return make_rays(o.astype(f), d)


def shadow_rays(verts: np.ndarray, n: int, seed: int = SEED_SCENE, first: int = 0) -> np.ndarray:
    """Config 4 style any-hit set: origin = random surface point + 1e-3 * normal, target = random point on a
    unit area light above the scene; t = |target - origin| * (1 - 1e-4).  Order is hash-shuffled (incoherent)."""
    f = np.float32
    lo, hi = bounds(verts)
    idx = np.arange(first, first + n, dtype=np.uint64)
    tri = (splitmix64(np.uint64(seed) ^ (idx * np.uint64(16) + np.uint64(7))) % np.uint64(len(verts))).astype(np.int64)
    a, b = hash_unit(seed, idx, 0), hash_unit(seed, idx, 1)
    flip = (a + b) > 1
    a = np.where(flip, f(1) - a, a).astype(f)
    b = np.where(flip, f(1) - b, b).astype(f)
    v0, v1, v2 = verts[tri, 0], verts[tri, 1], verts[tri, 2]
    p = v0 + a[:, None] * (v1 - v0) + b[:, None] * (v2 - v0)
    nrm = np.cross(v1 - v0, v2 - v0).astype(f)
    ln = np.linalg.norm(nrm, axis=1).astype(f)
    nrm = nrm / np.maximum(ln, f(1e-20))[:, None]
    ext = (hi - lo).astype(f)
    light_c = np.array([(lo[0] + hi[0]) * 0.5, hi[1] + ext[1], (lo[2] + hi[2]) * 0.5], dtype=f)
    tgt = light_c[None, :] + np.stack([(hash_unit(seed, idx, 2) - f(0.5)) * ext[0], np.zeros(n, dtype=f),
                                       (hash_unit(seed, idx, 3) - f(0.5)) * ext[2]], axis=1).astype(f)
    nrm = np.where(((tgt - p) * nrm).sum(axis=1, keepdims=True) < 0, -nrm, nrm).astype(f)
    o = (p + f(1e-3) * nrm).astype(f)
    dvec = (tgt - o).astype(f)
    dist = np.linalg.norm(dvec, axis=1).astype(f)
    d = (dvec / dist[:, None]).astype(f)
    return make_rays(o, d, T_MIN, (dist * f(1.0 - 1e-4)).astype(f))


def pack4(rays: np.ndarray) -> np.ndarray:
    """Four consecutive rays -> one RayPacket4 (examples/benchmark.rs:135-141 packs x, x+1, x+2, x+3)."""
    assert len(rays) % 4 == 0
    r = rays.reshape(-1, 4)
    p = np.zeros(len(r), dtype=PACKET_DTYPE)
    for k, ax in enumerate("xyz"):
        p["origin_" + ax] = r["origin"][:, :, k]
        p["direction_" + ax] = r["direction"][:, :, k]
    p["t"] = r["t"]
    return p


# ------------------------------------------------------------------------------------------------------------------
# tree statistics for the builders' roofline
# ------------------------------------------------------------------------------------------------------------------
def leaf_depth_stats(nodes: np.ndarray) -> dict:
    """Depth statistics of a binary tree in the reference's node format (`count >= 0` leaf, `count == -1` inner with
    children `left_first`, `left_first + 1`; root at depth 0), computed level by level without recursion.
    `mean_leaf_depth_per_prim` is D-bar of the binned-SAH byte model `148 + 60 * D-bar` B per triangle: the number of
    levels whose bin and partition passes touch a primitive = the depth of its leaf."""
    count = nodes["count"].astype(np.int64)
    left = nodes["left_first"].astype(np.int64)
    level = np.array([0], dtype=np.int64)
    depth = 0
    leaves = prims = 0
    sum_leaf = sum_prim = 0
    max_depth = 0
    while len(level):
        valid = level[left[level] >= 0]
        is_leaf = count[valid] >= 0
        lv = valid[is_leaf]
        if len(lv):
            leaves += len(lv)
            p = int(count[lv].sum())
            prims += p
            sum_leaf += depth * len(lv)
            sum_prim += depth * p
            max_depth = depth
        inner = left[valid[~is_leaf]]
        level = np.concatenate([inner, inner + 1])
        depth += 1
    return {"leaves": int(leaves), "prims": int(prims), "max_depth": int(max_depth),
            "mean_leaf_depth": sum_leaf / max(1, leaves), "mean_leaf_depth_per_prim": sum_prim / max(1, prims)}


def binned_sah_bytes_per_tri(mean_leaf_depth_per_prim: float) -> float:
    """36 (vertices in) + 44 (aabb + centroid out) + D-bar * [(4 + 12 + 24) bin pass + (4 + 12 + 4)
    partition] + 2 * 32 (nodes out) + 4 (index out) = 148 + 60 * D-bar."""
    return 148.0 + 60.0 * mean_leaf_depth_per_prim
