"""Synthetic urban tiles: 50 x 50 m tile in RD coordinates with
ground, facades (+ balconies), cars, street furniture / tree crowns, clutter and uniform
noise, a matching AHN raster (ground / building surfaces, float16, NaN holes) and BGT-like
building / road polygons.  Used by the tests, the golden-vector generator and bench.py;
the real demo clouds of the reference are not part of its checkout.

Coordinates are quantised to the LAS grid (multiples of 0.001 m, offset 0), float64.
"""
import numpy as np

from .utils.bgt_utils import RingPolygon

TILE_ORIGIN = (119300.0, 485100.0)      # tile code 2386_9702
TILE_SIZE = 50.0


def tilecode_for(origin=TILE_ORIGIN):
    return f'{int(origin[0] // 50)}_{int(origin[1] // 50)}'


def _ground_z(x, y, origin):
    u = (x - origin[0]) / TILE_SIZE
    v = (y - origin[1]) / TILE_SIZE
    return 0.6 + 0.3 * np.sin(2 * np.pi * u * 1.5) * np.cos(2 * np.pi * v) + 0.15 * np.sin(2 * np.pi * (u + v) * 2.0)


def _quantise(a):
    return np.round(a * 1000.0) * 0.001


def make_buildings(rng, origin, n_buildings=8):
    """Axis-aligned building boxes on a jittered 3 x 3 layout (one slot left for the crossing)."""
    slots = [(i, j) for i in range(3) for j in range(3) if not (i == 1 and j == 1)]
    rng.shuffle(slots)
    boxes = []
    for (i, j) in slots[:n_buildings]:
        w = rng.uniform(8.0, 13.0)
        d = rng.uniform(8.0, 13.0)
        cx = origin[0] + 2.0 + i * 17.0 + rng.uniform(0.5, 14.0 - w + 0.5) + w / 2 - 1.0
        cy = origin[1] + 2.0 + j * 17.0 + rng.uniform(0.5, 14.0 - d + 0.5) + d / 2 - 1.0
        h = rng.uniform(6.0, 18.0)
        boxes.append((cx - w / 2, cy - d / 2, cx + w / 2, cy + d / 2, h))
    return boxes


def make_tile(n_points, seed=20240000, origin=TILE_ORIGIN, facade_fraction=0.33, shuffle=True,
              order='F', n_cars=40):
    """Returns a dict with
         points   (n,3) float64 (F-ordered by default, like pipeline.py:124 delivers)
         kind     uint8[n]  0 ground, 1 facade, 2 car, 3 furniture, 4 clutter, 5 noise
         ahn_tile dict x, y, ground_surface, building_surface (float64 from float16)
         building_polygons / road_polygons  lists of RingPolygon
         tilecode
         car_centres / car_dims, pole_xy / pole_r / pole_h   layout of the cars and poles
    """
    rng = np.random.Generator(np.random.Philox(seed))
    ox, oy = origin
    n = int(n_points)
    f_fac = facade_fraction
    rest = 1.0 - f_fac
    # fractions of the non-facade remainder follow the 45/8/10/2/2 mix of the survey
    base = np.array([0.45, 0.08, 0.10, 0.02, 0.02])
    base = base / base.sum() * rest
    counts = np.floor(np.array([base[0], f_fac, base[1], base[2], base[3], base[4]]) * n).astype(np.int64)
    counts[0] += n - counts.sum()
    n_ground, n_facade, n_car, n_furn, n_clutter, n_noise = counts

    boxes = make_buildings(rng, origin)
    xs, ys, zs, kinds = [], [], [], []

    # ground
    gx = ox + rng.random(n_ground) * TILE_SIZE
    gy = oy + rng.random(n_ground) * TILE_SIZE
    gzv = _ground_z(gx, gy, origin) + rng.normal(0.0, 0.02, n_ground)
    xs.append(gx); ys.append(gy); zs.append(gzv); kinds.append(np.zeros(n_ground, np.uint8))

    # facades: vertical walls of the boxes, sigma_perp = 0.01 m, 6 % balcony points
    walls = []
    for (x0, y0, x1, y1, h) in boxes:
        walls += [(x0, y0, x1, y0, h, 0, -1), (x1, y0, x1, y1, h, 1, 0),
                  (x1, y1, x0, y1, h, 0, 1), (x0, y1, x0, y0, h, -1, 0)]
    warea = np.array([np.hypot(w[2] - w[0], w[3] - w[1]) * w[4] for w in walls])
    wsel = rng.choice(len(walls), size=n_facade, p=warea / warea.sum())
    wa = np.array(walls)
    t = rng.random(n_facade)
    fx = wa[wsel, 0] + t * (wa[wsel, 2] - wa[wsel, 0])
    fy = wa[wsel, 1] + t * (wa[wsel, 3] - wa[wsel, 1])
    base_z = _ground_z(fx, fy, origin)
    fz = base_z + rng.random(n_facade) * wa[wsel, 4]
    perp = rng.normal(0.0, 0.01, n_facade)
    balcony = rng.random(n_facade) < 0.06
    perp = np.where(balcony, rng.uniform(0.0, 0.8, n_facade), perp)
    fx = fx + perp * wa[wsel, 5]
    fy = fy + perp * wa[wsel, 6]
    xs.append(fx); ys.append(fy); zs.append(fz); kinds.append(np.ones(n_facade, np.uint8))

    # roads: two crossing strips; cars are 4.5 x 1.8 x 1.5 m box shells on them
    road_h = (ox, oy + 19.5, ox + TILE_SIZE, oy + 25.0)        # x0, y0, x1, y1
    road_v = (ox + 19.5, oy, ox + 25.0, oy + TILE_SIZE)
    car_c = np.empty((n_cars, 2)); car_dim = np.empty((n_cars, 2))
    for c in range(n_cars):
        if c % 2 == 0:
            car_c[c] = (ox + rng.uniform(3, 47), oy + rng.uniform(20.6, 23.9)); car_dim[c] = (4.5, 1.8)
        else:
            car_c[c] = (ox + rng.uniform(20.6, 23.9), oy + rng.uniform(3, 47)); car_dim[c] = (1.8, 4.5)
    csel = rng.integers(0, n_cars, n_car)
    face = rng.integers(0, 5, n_car)
    u = rng.random(n_car) - 0.5; v = rng.random(n_car) - 0.5; w = rng.random(n_car)
    lx = np.where(face == 0, -0.5, np.where(face == 1, 0.5, u))
    ly = np.where(face == 2, -0.5, np.where(face == 3, 0.5, v))
    lz = np.where(face == 4, 1.0, w)
    cx = car_c[csel, 0] + lx * car_dim[csel, 0] + rng.normal(0, 0.005, n_car)
    cy = car_c[csel, 1] + ly * car_dim[csel, 1] + rng.normal(0, 0.005, n_car)
    cz = _ground_z(cx, cy, origin) + 0.15 + lz * 1.5
    xs.append(cx); ys.append(cy); zs.append(cz); kinds.append(np.full(n_car, 2, np.uint8))

    # street furniture: 60 poles (cylinders) + 25 ellipsoid crowns
    n_pole, n_crown = 60, 25
    pole_xy = np.column_stack((ox + rng.uniform(1, 49, n_pole), oy + rng.uniform(1, 49, n_pole)))
    pole_r = rng.uniform(0.1, 0.3, n_pole); pole_h = rng.uniform(2.0, 8.0, n_pole)
    crown_xy = np.column_stack((ox + rng.uniform(4, 46, n_crown), oy + rng.uniform(4, 46, n_crown)))
    crown_r = rng.uniform(2.0, 3.0, n_crown); crown_z = rng.uniform(5.0, 10.0, n_crown)
    is_crown = rng.random(n_furn) < 0.6
    psel = rng.integers(0, n_pole, n_furn); ksel = rng.integers(0, n_crown, n_furn)
    ang = rng.uniform(0, 2 * np.pi, n_furn)
    hx = pole_xy[psel, 0] + pole_r[psel] * np.cos(ang)
    hy = pole_xy[psel, 1] + pole_r[psel] * np.sin(ang)
    hz = _ground_z(hx, hy, origin) + rng.random(n_furn) * pole_h[psel]
    d = rng.normal(size=(n_furn, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rr = crown_r[ksel] * rng.random(n_furn) ** (1.0 / 3.0)
    kx = crown_xy[ksel, 0] + d[:, 0] * rr
    ky = crown_xy[ksel, 1] + d[:, 1] * rr
    kz = crown_z[ksel] + d[:, 2] * rr * 0.8
    xs.append(np.where(is_crown, kx, hx)); ys.append(np.where(is_crown, ky, hy))
    zs.append(np.where(is_crown, kz, hz)); kinds.append(np.full(n_furn, 3, np.uint8))

    # low-density clutter near the poles
    qsel = rng.integers(0, n_pole, n_clutter)
    qx = pole_xy[qsel, 0] + rng.normal(0, 0.4, n_clutter)
    qy = pole_xy[qsel, 1] + rng.normal(0, 0.4, n_clutter)
    qz = _ground_z(qx, qy, origin) + np.abs(rng.normal(0.3, 0.4, n_clutter))
    xs.append(qx); ys.append(qy); zs.append(qz); kinds.append(np.full(n_clutter, 4, np.uint8))

    # uniform noise
    xs.append(ox + rng.random(n_noise) * TILE_SIZE); ys.append(oy + rng.random(n_noise) * TILE_SIZE)
    zs.append(rng.uniform(-2.0, 25.0, n_noise)); kinds.append(np.full(n_noise, 5, np.uint8))

    x = np.clip(np.concatenate(xs), ox, ox + TILE_SIZE)
    y = np.clip(np.concatenate(ys), oy, oy + TILE_SIZE)
    z = np.concatenate(zs)
    kind = np.concatenate(kinds)
    if shuffle:
        perm = rng.permutation(n)
        x, y, z, kind = x[perm], y[perm], z[perm], kind[perm]
    points = np.empty((n, 3), dtype=np.float64, order=order)
    points[:, 0] = _quantise(x); points[:, 1] = _quantise(y); points[:, 2] = _quantise(z)

    # AHN raster: 500 x 500 cells of 0.1 m, x ascending, y descending, float16 like the npz
    gx_c = ox + 0.05 + 0.1 * np.arange(500)
    gy_c = oy + TILE_SIZE - 0.05 - 0.1 * np.arange(500)
    XX, YY = np.meshgrid(gx_c, gy_c)
    ground = _ground_z(XX, YY, origin)
    building = np.full_like(ground, np.nan)
    for (x0, y0, x1, y1, h) in boxes:
        inside = (XX >= x0) & (XX <= x1) & (YY >= y0) & (YY <= y1)
        building[inside] = (_ground_z(XX, YY, origin) + h)[inside]
        ground[(XX >= x0 + 0.3) & (XX <= x1 - 0.3) & (YY >= y0 + 0.3) & (YY <= y1 - 0.3)] = np.nan
    ahn_tile = {'x': gx_c, 'y': gy_c,
                'ground_surface': ground.astype(np.float16).astype(float),
                'building_surface': building.astype(np.float16).astype(float)}

    # BGT-like polygons: footprints enlarged by 0.25 m (stands in for offset=0.25, which
    # needs shapely), each with a short collinear extra vertex; road strips as rings
    building_polygons = []
    for b, (x0, y0, x1, y1, h) in enumerate(boxes):
        e = 0.25
        ring = [(x0 - e, y0 - e), ((x0 + x1) / 2, y0 - e), (x1 + e, y0 - e), (x1 + e, y1 + e),
                (x0 - e, y1 + e), (x0 - e, y0 - e)]
        holes = []
        if b % 3 == 0:      # a courtyard
            mx, my = (x0 + x1) / 2, (y0 + y1) / 2
            holes = [[(mx - 1.5, my - 1.5), (mx - 1.5, my + 1.5), (mx + 1.5, my + 1.5),
                      (mx + 1.5, my - 1.5), (mx - 1.5, my - 1.5)]]
        building_polygons.append(RingPolygon(_quantise(np.array(ring)),
                                             [_quantise(np.array(hh)) for hh in holes]))
    def strip(r, k):
        x0, y0, x1, y1 = r
        top = [(x0 + (x1 - x0) * i / k, y1 + 0.2 * np.sin(i)) for i in range(k + 1)]
        bot = [(x1 - (x1 - x0) * i / k, y0 + 0.2 * np.cos(i)) for i in range(k + 1)]
        return RingPolygon(_quantise(np.array(bot[::-1][::-1] + top[::-1] + [bot[0]])))
    road_polygons = [strip(road_h, 24), strip(road_v, 24)]

    return {'points': points, 'kind': kind, 'ahn_tile': ahn_tile, 'boxes': boxes,
            'building_polygons': building_polygons, 'road_polygons': road_polygons,
            'tilecode': tilecode_for(origin), 'origin': origin,
            'car_centres': car_c, 'car_dims': car_dim, 'pole_xy': pole_xy, 'pole_r': pole_r, 'pole_h': pole_h}


def make_layout(seed, origin=TILE_ORIGIN, facade_fraction=0.33, n_cars=40):
    """Scene of one synthetic tile WITHOUT its points: buildings / walls, cars, poles, crowns, the matching AHN
    raster and the BGT-like polygons -- everything ``make_device_tile`` needs besides the per-point random
    stream (which lives on the device).  Same scene model as ``make_tile``; own random sequence."""
    rng = np.random.Generator(np.random.Philox(int(seed) + (1 << 40)))
    ox, oy = origin
    boxes = make_buildings(rng, origin)
    walls = []
    for (x0, y0, x1, y1, h) in boxes:
        walls += [(x0, y0, x1, y0, h, 0, -1), (x1, y0, x1, y1, h, 1, 0),
                  (x1, y1, x0, y1, h, 0, 1), (x0, y1, x0, y0, h, -1, 0)]
    wa = np.array(walls, dtype=np.float64)
    area = np.hypot(wa[:, 2] - wa[:, 0], wa[:, 3] - wa[:, 1]) * wa[:, 4]
    cum = np.cumsum(area) / area.sum()
    cum[-1] = 1.0
    walls8 = np.ascontiguousarray(np.column_stack((wa, cum)))
    cars = np.empty((n_cars, 4))
    for c in range(n_cars):
        if c % 2 == 0:
            cars[c] = (ox + rng.uniform(3, 47), oy + rng.uniform(20.6, 23.9), 4.5, 1.8)
        else:
            cars[c] = (ox + rng.uniform(20.6, 23.9), oy + rng.uniform(3, 47), 1.8, 4.5)
    n_pole, n_crown = 60, 25
    poles = np.column_stack((ox + rng.uniform(1, 49, n_pole), oy + rng.uniform(1, 49, n_pole),
                             rng.uniform(0.1, 0.3, n_pole), rng.uniform(2.0, 8.0, n_pole)))
    crowns = np.column_stack((ox + rng.uniform(4, 46, n_crown), oy + rng.uniform(4, 46, n_crown),
                              rng.uniform(2.0, 3.0, n_crown), rng.uniform(5.0, 10.0, n_crown)))
    rest = 1.0 - facade_fraction
    base = np.array([0.45, 0.08, 0.10, 0.02, 0.02])
    base = base / base.sum() * rest
    frac = np.array([base[0], facade_fraction, base[1], base[2], base[3], base[4]])
    cum_kind = np.cumsum(frac)
    cum_kind[-1] = 1.0
    ahn_tile, building_polygons, road_polygons = _raster_and_polygons(boxes, origin)
    return {'origin': origin, 'tilecode': tilecode_for(origin), 'boxes': boxes, 'walls8': walls8, 'cars4': cars,
            'poles4': np.ascontiguousarray(poles), 'crowns4': np.ascontiguousarray(crowns), 'cum_kind': cum_kind,
            'ahn_tile': ahn_tile, 'building_polygons': building_polygons, 'road_polygons': road_polygons,
            'seed': int(seed)}


def _raster_and_polygons(boxes, origin):
    """AHN raster (500 x 500 cells of 0.1 m, float16 values, NaN holes) and BGT-like rings of a scene, exactly
    as ``make_tile`` derives them from its building boxes."""
    ox, oy = origin
    gx_c = ox + 0.05 + 0.1 * np.arange(500)
    gy_c = oy + TILE_SIZE - 0.05 - 0.1 * np.arange(500)
    XX, YY = np.meshgrid(gx_c, gy_c)
    gz = _ground_z(XX, YY, origin)
    ground = gz.copy()
    building = np.full_like(ground, np.nan)
    for (x0, y0, x1, y1, h) in boxes:
        inside = (XX >= x0) & (XX <= x1) & (YY >= y0) & (YY <= y1)
        building[inside] = (gz + h)[inside]
        ground[(XX >= x0 + 0.3) & (XX <= x1 - 0.3) & (YY >= y0 + 0.3) & (YY <= y1 - 0.3)] = np.nan
    ahn_tile = {'x': gx_c, 'y': gy_c,
                'ground_surface': ground.astype(np.float16).astype(float),
                'building_surface': building.astype(np.float16).astype(float)}
    building_polygons = []
    for b, (x0, y0, x1, y1, h) in enumerate(boxes):
        e = 0.25
        ring = [(x0 - e, y0 - e), ((x0 + x1) / 2, y0 - e), (x1 + e, y0 - e), (x1 + e, y1 + e),
                (x0 - e, y1 + e), (x0 - e, y0 - e)]
        holes = []
        if b % 3 == 0:
            mx, my = (x0 + x1) / 2, (y0 + y1) / 2
            holes = [[(mx - 1.5, my - 1.5), (mx - 1.5, my + 1.5), (mx + 1.5, my + 1.5),
                      (mx + 1.5, my - 1.5), (mx - 1.5, my - 1.5)]]
        building_polygons.append(RingPolygon(_quantise(np.array(ring)), [_quantise(np.array(hh)) for hh in holes]))

    def strip(r, k):
        x0, y0, x1, y1 = r
        top = [(x0 + (x1 - x0) * i / k, y1 + 0.2 * np.sin(i)) for i in range(k + 1)]
        bot = [(x1 - (x1 - x0) * i / k, y0 + 0.2 * np.cos(i)) for i in range(k + 1)]
        return RingPolygon(_quantise(np.array(bot[::-1][::-1] + top[::-1] + [bot[0]])))
    road_h = (ox, oy + 19.5, ox + TILE_SIZE, oy + 25.0)
    road_v = (ox + 19.5, oy, ox + 25.0, oy + TILE_SIZE)
    return ahn_tile, building_polygons, [strip(road_h, 24), strip(road_v, 24)]


def tile_origin(index):
    """Origin of synthetic tile ``index`` of a batch: a 64-wide mosaic of 50 m tiles, so every tile has its
    own tile code (own AHN raster, own BGT query) like the tiles of a real run."""
    return (TILE_ORIGIN[0] + TILE_SIZE * (index % 64), TILE_ORIGIN[1] + TILE_SIZE * (index // 64))


def make_device_tile(n_points, seed=20240000, origin=TILE_ORIGIN, facade_fraction=0.33, device=None, layout=None):
    """Synthetic tile generated ON THE DEVICE (``upcp_synth_tile``).  Returns
    (DeviceTile, d_kind uint8 cuda tensor, layout dict as ``make_layout``).  The host never sees the points;
    ``DeviceTile`` coordinates can be downloaded (``tile.x.cpu()`` ...) where the oracle has to label them."""
    import ctypes
    import torch
    from . import _lib
    from . import device as dev
    lib = _lib.require_cuda()
    device = device if device is not None else dev.current_device()
    if layout is None:
        layout = make_layout(seed, origin=origin, facade_fraction=facade_fraction)
    n = int(n_points)
    x = torch.empty(n, dtype=torch.float64, device=device)
    y = torch.empty(n, dtype=torch.float64, device=device)
    z = torch.empty(n, dtype=torch.float64, device=device)
    kind = torch.empty(n, dtype=torch.uint8, device=device)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)      # noqa: E731
    walls, cars, poles, crowns = up(layout['walls8']), up(layout['cars4']), up(layout['poles4']), up(layout['crowns4'])
    o3 = (ctypes.c_double * 3)(float(layout['origin'][0]), float(layout['origin'][1]), TILE_SIZE)
    ck = (ctypes.c_double * 6)(*[float(v) for v in layout['cum_kind']])
    _lib.check(lib.upcp_synth_tile(n, int(seed), o3, ck, dev.ptr(walls), len(layout['walls8']), dev.ptr(cars),
                                   len(layout['cars4']), dev.ptr(poles), len(layout['poles4']), dev.ptr(crowns),
                                   len(layout['crowns4']), dev.ptr(x), dev.ptr(y), dev.ptr(z), dev.ptr(kind),
                                   dev.stream_ptr()), 'upcp_synth_tile')
    torch.cuda.current_stream(device).synchronize()          # the small layout tensors may go out of scope now
    return dev.DeviceTile.from_device(x, y, z), kind, layout


def make_demo_cloud(n_points, ahn_tile, building_rings, seed=23869702, origin=TILE_ORIGIN):
    """Config 1: a synthetic cloud over the REAL demo tile 2386_9702 -- the reference's demo
    clouds are not part of its checkout, its AHN raster (NaN holes included) and BGT footprints are.
    Ground follows the real AHN ground surface, facades stand on the real building rings and rise to
    the real building surface, roofs lie on it; plus clutter, crowns and uniform noise.
    ``ahn_tile``: dict x, y (descending), ground_surface, building_surface; ``building_rings``: list of
    closed (V, 2) rings.  Returns (points (n, 3) float64 F-ordered on the LAS millimetre grid, kind uint8
    0 ground, 1 facade, 2 roof, 3 clutter, 4 crown, 5 noise)."""
    rng = np.random.Generator(np.random.Philox(seed))
    ox, oy = origin
    n = int(n_points)
    frac = np.array([0.42, 0.30, 0.16, 0.05, 0.05, 0.02])
    counts = np.floor(frac * n).astype(np.int64)
    counts[0] += n - counts.sum()
    gx0, gy0 = float(ahn_tile['x'][0]), float(ahn_tile['y'][0])
    step = float(ahn_tile['x'][1] - ahn_tile['x'][0])
    ground, building = ahn_tile['ground_surface'], ahn_tile['building_surface']
    ny, nx = ground.shape

    def cell(x, y):
        ix = np.clip(np.floor((x - gx0) / step + 0.5).astype(np.int64), 0, nx - 1)
        iy = np.clip(np.floor((gy0 - y) / step + 0.5).astype(np.int64), 0, ny - 1)
        return iy, ix

    def ground_at(x, y):
        iy, ix = cell(x, y)
        return np.nan_to_num(ground[iy, ix], nan=0.5)

    xs, ys, zs, kinds = [], [], [], []
    # ground
    x = ox + rng.random(counts[0]) * TILE_SIZE; y = oy + rng.random(counts[0]) * TILE_SIZE
    xs.append(x); ys.append(y); zs.append(ground_at(x, y) + rng.normal(0.0, 0.03, counts[0])); kinds.append(0)
    # facades on the ring edges that lie inside the tile
    seg = np.vstack([np.column_stack((r[:-1], r[1:])) for r in building_rings])
    mid = 0.5 * (seg[:, :2] + seg[:, 2:])
    keep = (mid[:, 0] > ox) & (mid[:, 0] < ox + TILE_SIZE) & (mid[:, 1] > oy) & (mid[:, 1] < oy + TILE_SIZE)
    seg = seg[keep]
    length = np.hypot(seg[:, 2] - seg[:, 0], seg[:, 3] - seg[:, 1])
    sel = rng.choice(len(seg), size=counts[1], p=length / length.sum())
    t = rng.random(counts[1])
    fx = seg[sel, 0] + t * (seg[sel, 2] - seg[sel, 0]); fy = seg[sel, 1] + t * (seg[sel, 3] - seg[sel, 1])
    nrm = np.column_stack((seg[sel, 3] - seg[sel, 1], -(seg[sel, 2] - seg[sel, 0]))) / length[sel, None]
    # eaves height: the building surface a little to either side of the wall, 9 m where AHN has none
    iy, ix = cell(fx + 0.3 * nrm[:, 0], fy + 0.3 * nrm[:, 1]); h1 = building[iy, ix]
    iy, ix = cell(fx - 0.3 * nrm[:, 0], fy - 0.3 * nrm[:, 1]); h2 = building[iy, ix]
    top = np.nan_to_num(np.fmax(h1, h2), nan=9.0)
    base = ground_at(fx, fy)
    perp = rng.normal(0.0, 0.012, counts[1])
    xs.append(fx + perp * nrm[:, 0]); ys.append(fy + perp * nrm[:, 1])
    zs.append(base + rng.random(counts[1]) * np.maximum(top - base, 2.5)); kinds.append(1)
    # roofs: raster cells with a building surface
    by, bx = np.nonzero(np.isfinite(building))
    pick = rng.integers(0, len(by), counts[2])
    rx = gx0 + (bx[pick] + rng.random(counts[2]) - 0.5) * step
    ry = gy0 - (by[pick] + rng.random(counts[2]) - 0.5) * step
    xs.append(rx); ys.append(ry); zs.append(building[by[pick], bx[pick]] + rng.normal(0.0, 0.03, counts[2])); kinds.append(2)
    # clutter: blobs on the ground (cars, bins, ...)
    n_blob = 60
    bxy = np.column_stack((ox + rng.uniform(2, 48, n_blob), oy + rng.uniform(2, 48, n_blob)))
    bsz = rng.uniform(0.3, 1.6, n_blob)
    b = rng.integers(0, n_blob, counts[3])
    cx_ = bxy[b, 0] + rng.uniform(-1, 1, counts[3]) * bsz[b]; cy_ = bxy[b, 1] + rng.uniform(-1, 1, counts[3]) * bsz[b]
    xs.append(cx_); ys.append(cy_); zs.append(ground_at(cx_, cy_) + 0.1 + rng.random(counts[3]) * 1.4); kinds.append(3)
    # crowns
    n_crown = 20
    cxy = np.column_stack((ox + rng.uniform(4, 46, n_crown), oy + rng.uniform(4, 46, n_crown)))
    cr = rng.uniform(1.5, 3.0, n_crown); cz = rng.uniform(5.0, 9.0, n_crown)
    k = rng.integers(0, n_crown, counts[4])
    d = rng.normal(size=(counts[4], 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rr = cr[k] * rng.random(counts[4]) ** (1.0 / 3.0)
    xs.append(cxy[k, 0] + d[:, 0] * rr); ys.append(cxy[k, 1] + d[:, 1] * rr); zs.append(cz[k] + 0.8 * d[:, 2] * rr)
    kinds.append(4)
    # noise
    xs.append(ox + rng.random(counts[5]) * TILE_SIZE); ys.append(oy + rng.random(counts[5]) * TILE_SIZE)
    zs.append(rng.uniform(-2.0, 25.0, counts[5])); kinds.append(5)

    x = np.clip(np.concatenate(xs), ox, ox + TILE_SIZE - 0.001)
    y = np.clip(np.concatenate(ys), oy + 0.001, oy + TILE_SIZE)
    z = np.concatenate(zs)
    kind = np.concatenate([np.full(c, k_, np.uint8) for c, k_ in zip(counts, kinds)])
    perm = rng.permutation(n)
    points = np.empty((n, 3), dtype=np.float64, order='F')
    points[:, 0] = _quantise(x[perm]); points[:, 1] = _quantise(y[perm]); points[:, 2] = _quantise(z[perm])
    return points, kind[perm]


def make_polygon_lattice(n_side=71, seed=5000, origin=TILE_ORIGIN, hole_fraction=0.1):
    """Config 4: jittered n_side x n_side lattice of star-shaped rings (circumradius
    0.25-0.45 m, 5-60 vertices), a fraction with one hole."""
    rng = np.random.Generator(np.random.Philox(seed))
    pitch = TILE_SIZE / n_side
    polys = []
    for i in range(n_side):
        for j in range(n_side):
            cx = origin[0] + (i + 0.5) * pitch + rng.uniform(-0.05, 0.05)
            cy = origin[1] + (j + 0.5) * pitch + rng.uniform(-0.05, 0.05)
            nv = int(min(60, max(5, rng.geometric(0.1))))
            ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
            rad = rng.uniform(0.25, 0.45) * rng.uniform(0.6, 1.0, nv)
            ring = np.column_stack((cx + rad * np.cos(ang), cy + rad * np.sin(ang)))
            ring = np.vstack([ring, ring[:1]])
            holes = []
            if rng.random() < hole_fraction:
                ha = np.linspace(0, 2 * np.pi, 7)
                holes = [np.column_stack((cx + 0.08 * np.cos(ha), cy + 0.08 * np.sin(ha)))]
                holes[0][-1] = holes[0][0]
            polys.append(RingPolygon(_quantise(ring), [_quantise(h) for h in holes]))
    return polys
