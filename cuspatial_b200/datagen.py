"""Seeded synthetic inputs for the quadtree point-in-polygon join.

The NYC taxi-zone file the reference's notebooks use is not available offline, so the
benchmark polygons are "taxi-zone-like": star-shaped simple polygons on a jittered grid with
overlapping bounding boxes, log-normal vertex counts (median 200, clipped to [8, 4096]),
closed rings, 5 % with one interior hole.  A second family reproduces the reference's own
benchmark shape, regular n-gons (cpp/include/cuspatial_test/geometry_generator.cuh:104-115).

Everything here is host-side numpy except `uniform_points_torch` / `clustered_points_torch`,
which generate large point clouds directly in device memory for bench.py.
"""
import math

import numpy as np


def taxi_zone_like_polygons(n_poly=263, extent=(0.0, 1.0, 0.0, 1.0), seed=20251017,
                            dtype=np.float64, median_vertices=200, hole_fraction=0.05,
                            min_vertices=8, max_vertices=4096):
    """Return (poly_offsets u32[n+1], ring_offsets u32[r+1], vx, vy) in GeoArrow layout."""
    rng = np.random.default_rng(seed)
    x0, x1, y0, y1 = extent
    g = int(math.ceil(math.sqrt(n_poly)))
    pitch_x, pitch_y = (x1 - x0) / g, (y1 - y0) / g
    cells = rng.permutation(g * g)[:n_poly]
    cells.sort()
    poly_offsets = [0]
    ring_offsets = [0]
    xs, ys = [], []
    for c in cells:
        cx = x0 + ((c % g) + 0.5 + rng.uniform(-0.2, 0.2)) * pitch_x
        cy = y0 + ((c // g) + 0.5 + rng.uniform(-0.2, 0.2)) * pitch_y
        nv = int(np.clip(round(math.exp(rng.normal(math.log(median_vertices), 0.6))),
                         min_vertices, max_vertices))
        r0 = rng.uniform(0.6, 0.9)
        k = np.arange(1, 6)
        a = rng.uniform(0.0, 0.12, size=5) / k ** 0.5
        phi = rng.uniform(0, 2 * math.pi, size=5)
        theta = np.sort(rng.uniform(0, 2 * math.pi, size=nv))
        r = r0 * (1.0 + (a[None, :] * np.cos(k[None, :] * theta[:, None] + phi[None, :])).sum(1))
        px = cx + r * pitch_x * np.cos(theta)
        py = cy + r * pitch_y * np.sin(theta)
        xs.append(np.concatenate([px, px[:1]]))
        ys.append(np.concatenate([py, py[:1]]))
        ring_offsets.append(ring_offsets[-1] + nv + 1)
        n_rings = 1
        if rng.uniform() < hole_fraction:
            nh = max(min_vertices, nv // 4)
            th = np.sort(rng.uniform(0, 2 * math.pi, size=nh))[::-1]
            rh = 0.3 * r0 * (1.0 + 0.1 * np.cos(3 * th + phi[0]))
            hx = cx + rh * pitch_x * np.cos(th)
            hy = cy + rh * pitch_y * np.sin(th)
            xs.append(np.concatenate([hx, hx[:1]]))
            ys.append(np.concatenate([hy, hy[:1]]))
            ring_offsets.append(ring_offsets[-1] + nh + 1)
            n_rings = 2
        poly_offsets.append(poly_offsets[-1] + n_rings)
    vx = np.concatenate(xs).astype(dtype)
    vy = np.concatenate(ys).astype(dtype)
    return (np.asarray(poly_offsets, dtype=np.uint32), np.asarray(ring_offsets, dtype=np.uint32),
            vx, vy)


def regular_ngons(n_poly, n_sides, radius, centroid=(0.0, 0.0), dtype=np.float64):
    """The reference benchmark's shape: n regular polygons sharing one centroid, closed rings."""
    t = np.arange(n_sides + 1) % n_sides * (2 * math.pi / n_sides)
    vx = np.tile(centroid[0] + radius * np.cos(t), n_poly).astype(dtype)
    vy = np.tile(centroid[1] + radius * np.sin(t), n_poly).astype(dtype)
    poly_offsets = np.arange(n_poly + 1, dtype=np.uint32)
    ring_offsets = (np.arange(n_poly + 1) * (n_sides + 1)).astype(np.uint32)
    return poly_offsets, ring_offsets, vx, vy


def polygon_extent(vx, vy, pad_fraction=1e-3):
    """Bounding box of all polygons padded so the quadtree area strictly contains them."""
    x0, x1 = float(vx.min()), float(vx.max())
    y0, y1 = float(vy.min()), float(vy.max())
    px, py = (x1 - x0) * pad_fraction, (y1 - y0) * pad_fraction
    return x0 - px, x1 + px, y0 - py, y1 + py


def quadtree_params(extent, max_depth=15):
    """scale = max(dx, dy) / 2^max_depth keeps every in-bbox cell index < 2^max_depth
    (the well-defined regime of the reference)."""
    x0, x1, y0, y1 = extent
    return max(x1 - x0, y1 - y0) / float(1 << max_depth)


def uniform_points(n, extent, seed=1, dtype=np.float64):
    rng = np.random.default_rng(seed)
    x0, x1, y0, y1 = extent
    x = (x0 + (x1 - x0) * rng.random(n)).astype(dtype)
    y = (y0 + (y1 - y0) * rng.random(n)).astype(dtype)
    # keep the half-open box after the cast (fp32 rounding can land on the upper edge)
    x = np.minimum(x, np.nextafter(dtype(x1), dtype(x0)))
    y = np.minimum(y, np.nextafter(dtype(y1), dtype(y0)))
    return x, y


def clustered_points(n, extent, seed=2, dtype=np.float64, components=64, clip=True):
    """Gaussian mixture: centres uniform in the box, sigma log-uniform in [0.2 %, 5 %] of the
    extent, Dirichlet(1) weights.  clip=True folds outliers back into the half-open box."""
    rng = np.random.default_rng(seed)
    x0, x1, y0, y1 = extent
    w = rng.dirichlet(np.ones(components))
    cx = rng.uniform(x0, x1, components)
    cy = rng.uniform(y0, y1, components)
    sig = np.exp(rng.uniform(math.log(0.002), math.log(0.05), components))
    comp = rng.choice(components, size=n, p=w)
    x = cx[comp] + rng.normal(size=n) * sig[comp] * (x1 - x0)
    y = cy[comp] + rng.normal(size=n) * sig[comp] * (y1 - y0)
    if clip:
        x = x0 + np.mod(x - x0, x1 - x0)
        y = y0 + np.mod(y - y0, y1 - y0)
    x, y = x.astype(dtype), y.astype(dtype)
    if clip:
        x = np.clip(x, dtype(x0), np.nextafter(dtype(x1), dtype(x0)))
        y = np.clip(y, dtype(y0), np.nextafter(dtype(y1), dtype(y0)))
    return x, y


def uniform_points_torch(n, extent, seed, dtype, device):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x0, x1, y0, y1 = extent
    x = torch.rand(n, generator=g, device=device, dtype=dtype) * (x1 - x0) + x0
    y = torch.rand(n, generator=g, device=device, dtype=dtype) * (y1 - y0) + y0
    hi_x = float(np.nextafter(np.dtype(str(dtype).split(".")[-1]).type(x1), -np.inf))
    hi_y = float(np.nextafter(np.dtype(str(dtype).split(".")[-1]).type(y1), -np.inf))
    x.clamp_(max=hi_x)
    y.clamp_(max=hi_y)
    return x, y


def mixture_parameters(extent, seed, components=64):
    """The Gaussian mixture of BASELINE.json configs[3]: centres uniform in
    the box, sigma log-uniform in [0.2 %, 5 %] of the extent, Dirichlet(1) weights."""
    rng = np.random.default_rng(seed)
    x0, x1, y0, y1 = extent
    w = rng.dirichlet(np.ones(components))
    cx = rng.uniform(x0, x1, components)
    cy = rng.uniform(y0, y1, components)
    sig = np.exp(rng.uniform(math.log(0.002), math.log(0.05), components))
    return w, cx, cy, sig


def clustered_points_torch(n, extent, seed, dtype, device, components=64, mixture_seed=None,
                           out=None, chunk=1 << 26):
    """`n` samples of the mixture, generated on the device in chunks (bounded temporaries: a
    1 G-point cloud needs no more than its own 16 GB).  Samples falling outside the box are folded
    back into it (wrap-around), then clamped to the half-open box of `dtype`.  `mixture_seed`
    fixes the mixture independently of the sample stream (ranks of a sharded run draw different
    samples of the SAME mixture); `out` = (x, y) tensors to fill in place."""
    import torch

    x0, x1, y0, y1 = extent
    w, cx, cy, sig = mixture_parameters(extent, seed if mixture_seed is None else mixture_seed,
                                        components)
    w = torch.tensor(w, device=device, dtype=torch.float32)
    cx = torch.tensor(cx, device=device, dtype=torch.float64)
    cy = torch.tensor(cy, device=device, dtype=torch.float64)
    sig = torch.tensor(sig, device=device, dtype=torch.float64)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    npdt = np.dtype(str(dtype).split(".")[-1]).type
    lo_x, hi_x = float(npdt(x0)), float(np.nextafter(npdt(x1), -np.inf))
    lo_y, hi_y = float(npdt(y0)), float(np.nextafter(npdt(y1), -np.inf))
    if out is None:
        out = (torch.empty(n, dtype=dtype, device=device), torch.empty(n, dtype=dtype, device=device))
    ox, oy = out
    for a in range(0, n, chunk):
        m = min(chunk, n - a)
        comp = torch.multinomial(w, m, replacement=True, generator=g)
        s = sig[comp]
        x = cx[comp] + torch.randn(m, generator=g, device=device, dtype=torch.float64) * s * (x1 - x0)
        y = cy[comp] + torch.randn(m, generator=g, device=device, dtype=torch.float64) * s * (y1 - y0)
        del comp, s
        x = x0 + torch.remainder(x - x0, x1 - x0)
        y = y0 + torch.remainder(y - y0, y1 - y0)
        ox[a: a + m] = x.to(dtype).clamp_(min=lo_x, max=hi_x)
        oy[a: a + m] = y.to(dtype).clamp_(min=lo_y, max=hi_y)
        del x, y
    return ox, oy


def nested_rectangle_points(size, seed=0, dtype=np.float64):
    """The reference's quadtree benchmark cloud (cpp/benchmarks/indexing/quadtree_on_points.cu:
    33-100): golden-ratio nested rectangles inside [0, size]^2, sqrt(area * 1e6) uniform points
    in each -- ever denser towards the spiral's centre.  Returns (x, y)."""
    phi = (1 + math.sqrt(5)) * 0.5
    tl, br = [0.0, 0.0], [float(size), float(size)]
    rng = np.random.default_rng(seed)
    xs, ys, k = [], [], 0
    while True:
        if k % 4 == 0:
            br[0] = tl[0] - (tl[0] - br[0]) / phi
        elif k % 4 == 1:
            br[1] = tl[1] - (tl[1] - br[1]) / phi
        elif k % 4 == 2:
            tl[0] = tl[0] + (br[0] - tl[0]) / phi
        else:
            tl[1] = tl[1] + (br[1] - tl[1]) / phi
        ax, ay = br[0] - tl[0], br[1] - tl[1]
        m = int(math.sqrt(ax * ay * 1_000_000))
        xs.append(tl[0] + ax * rng.random(m))
        ys.append(tl[1] + ay * rng.random(m))
        k += 1
        if not (ax > 1 and ay > 1):
            break
    return np.concatenate(xs).astype(dtype), np.concatenate(ys).astype(dtype)
