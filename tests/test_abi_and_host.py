"""CPU: the C-ABI library loads and exports every symbol include/upcp_cuda.h declares, the
product refuses to run without a GPU (no CPU fallback), and the host-side logic (PIP plan
builder, octree parameters, readers, synthetic generator, device arithmetic compiled for
the host) agrees with the oracle."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, has_cuda
from oracle import natives, reference_py as R
from urban_pointcloud_processing_b200 import _lib, synthetic
from urban_pointcloud_processing_b200 import device as dev
from urban_pointcloud_processing_b200.utils import ahn_utils, bgt_utils, las_utils, math_utils


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'upcp_cuda.h')).read()
    declared = set(re.findall(r'\b(upcp_[a-z0-9_]+)\s*\(', header))
    declared -= {'upcp_pip_plan'}
    assert len(declared) >= 35
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in include/upcp_cuda.h but not exported'
    # and the ctypes prototypes cover exactly the declared entry points
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert _lib.load().upcp_abi_version() == 1


def test_no_reference_or_oracle_imports_in_product():
    pkg = os.path.join(ROOT, 'urban_pointcloud_processing_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f
                assert '/root/reference' not in src, f


@pytest.mark.skipif(has_cuda(), reason='checks the no-GPU behaviour')
def test_compute_fails_loudly_without_gpu():
    pts = np.zeros((4, 3))
    with pytest.raises(_lib.UpcpCudaError):
        dev.DeviceTile(pts)
    lib = _lib.load()
    assert lib.upcp_device_count() == 0
    # a compute entry point called directly returns UPCP_E_CUDA, it does not compute on the CPU
    st = lib.upcp_labels_assign(None, 8, None, 1, None)
    assert st == -2 and b'upcp' not in lib.upcp_last_error()[:0]


def test_pip_plan_builder_host():
    tile = synthetic.make_tile(100, seed=1)
    ring_xy, off, poly, hole = dev.rings_from_polygons(tile['building_polygons'])
    assert off[-1] == len(ring_xy) and hole.sum() == 3 and poly.max() == 7
    lib = _lib.load()
    handle = ctypes.c_void_p()
    st = lib.upcp_pip_plan_create(ring_xy.ctypes.data_as(ctypes.c_void_p), off.ctypes.data_as(ctypes.c_void_p),
                                  poly.ctypes.data_as(ctypes.c_void_p), hole.ctypes.data_as(ctypes.c_void_p),
                                  len(poly), 0, ctypes.byref(handle))
    assert st == 0 and handle.value
    assert lib.upcp_pip_plan_bytes(handle) > 4096
    lib.upcp_pip_plan_destroy(handle)
    # invalid input: a polygon whose first ring is a hole
    bad_hole = hole.copy(); bad_hole[0] = 1
    st = lib.upcp_pip_plan_create(ring_xy.ctypes.data_as(ctypes.c_void_p), off.ctypes.data_as(ctypes.c_void_p),
                                  poly.ctypes.data_as(ctypes.c_void_p), bad_hole.ctypes.data_as(ctypes.c_void_p),
                                  len(poly), 0, ctypes.byref(handle))
    assert st == -1 and b'exterior' in lib.upcp_last_error()


def test_octree_helpers_match_oracle(golden):
    lib = _lib.load()
    g = golden('scalars_golden.npz')
    for e, gs, lvl in zip(g['extent'], g['grid_size'], g['level']):
        bbox = (ctypes.c_double * 6)(0, 0, 0, e, e * 0.3, 0)
        assert lib.upcp_octree_level(bbox, 3, float(gs)) == lvl
    rng = np.random.default_rng(3)
    for scale, org in ((50, (119300, 485100, 0)), (1.0, (119300, 485100, 3)), (1e-4, (0, 0, 0)), (300, (1e6, 2e6, -5))):
        pts = rng.random((500, 3)) * scale + np.array(org)
        f32 = pts.astype(np.float32)
        want = natives.octree_params(f32[:, 0], f32[:, 1], f32[:, 2])
        bbox = (ctypes.c_double * 6)(*pts.min(0), *pts.max(0))
        out = (ctypes.c_float * 4)()
        assert lib.upcp_octree_params(bbox, 21, out) == 0
        assert np.array_equal(np.array(out[:], dtype=np.float32), want)


@pytest.fixture(scope='module')
def emu():
    out = os.path.join(ROOT, 'tests', 'host_emu', '_emu.so')
    src = os.path.join(ROOT, 'tests', 'host_emu', 'emu.cpp')
    hdr = os.path.join(ROOT, 'urban_pointcloud_processing_b200', 'csrc', 'upcp_math.cuh')
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-fno-fast-math', '-fPIC', '-shared', '-x', 'c++',
                               src, '-o', out, '-lm'])
    return ctypes.CDLL(out)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_device_math_pip_matches_oracle(emu, golden):
    g = golden('pip_golden.npz')
    for name in ('bgt_roads', 'handmade', 'lattice'):
        pts = np.ascontiguousarray(g[f'{name}__points'])
        xy, off = g[f'{name}__ring_xy'], g[f'{name}__ring_off']
        px, py = np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, 1])
        for r in range(len(off) - 1):
            ring = np.ascontiguousarray(xy[off[r]:off[r + 1]])
            if len(ring) < 2:
                continue
            out = np.zeros(len(pts), dtype=np.uint8)
            emu.emu_pip_ring(_p(ring), len(ring), _p(px), _p(py), ctypes.c_long(len(pts)), _p(out))
            assert np.array_equal(out.view(bool), natives.is_inside(px, py, ring)), (name, r)


def test_device_math_digitize_matches_numpy(emu, golden):
    g = golden('raster_golden.npz')
    ahn = np.load(os.path.join(GOLDEN, 'demo', 'ahn_2386_9702.npz'))
    gx, gy = ahn['x'], ahn['y']
    bx = np.ascontiguousarray(gx - (gx[1] - gx[0]) / 2)
    by = np.ascontiguousarray(gy + (gy[0] - gy[1]) / 2)
    pts = g['points']
    px, py = np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, 1])
    ix = np.zeros(len(pts), dtype=np.int32); iy = np.zeros(len(pts), dtype=np.int32)
    emu.emu_digitize(_p(bx), len(bx), _p(by), len(by), _p(px), _p(py), ctypes.c_long(len(pts)), _p(ix), _p(iy))
    wx = np.digitize(px, bx) - 1; wy = np.digitize(py, by, right=True) - 1
    assert np.array_equal(ix, np.where(wx < 0, len(bx) - 1, wx))
    assert np.array_equal(iy, np.where(wy < 0, len(by) - 1, wy))
    for guess in (-5, 0, 250, 499, 9999):     # the fix-up must not depend on the first guess
        emu.emu_digitize_badguess(_p(bx), len(bx), _p(by), len(by), _p(px), _p(py), ctypes.c_long(len(pts)),
                                  guess, _p(ix), _p(iy))
        assert np.array_equal(ix, np.where(wx < 0, len(bx) - 1, wx))
        assert np.array_equal(iy, np.where(wy < 0, len(by) - 1, wy))


def test_device_math_classify(emu):
    rng = np.random.default_rng(1)
    z = rng.normal(0, 1, 4000); v = rng.normal(0, 1, 4000)
    v[::7] = np.nan; v[::11] = np.inf
    z[::13] = v[::13] + 0.2          # exactly on the threshold
    a, b = 0.2, 0.7
    with np.errstate(invalid='ignore'):
        want = {0: np.abs(z - v) < a, 1: z < v + a, 2: ~np.isfinite(v) | (z <= v + a),
                3: (z > v + a) & (z <= v + b), 4: (z <= v + b) & (z >= v - a), 5: (z - v) < -a}
    for mode, w in want.items():
        out = np.zeros(len(z), dtype=np.uint8)
        emu.emu_classify(mode, _p(z), _p(v), ctypes.c_long(len(z)), ctypes.c_double(a), ctypes.c_double(b), _p(out))
        assert np.array_equal(out.view(bool), w), mode
    out = np.zeros(len(z), dtype=np.uint8)
    emu.emu_classify(3, _p(z), _p(v), ctypes.c_long(len(z)), ctypes.c_double(-np.inf), ctypes.c_double(np.inf), _p(out))
    with np.errstate(invalid='ignore'):
        assert np.array_equal(out.view(bool), (z > v - np.inf) & (z <= v + np.inf))


def test_device_math_eigen_cov_angle_match_oracle(emu):
    rng = np.random.default_rng(2)
    pts = np.ascontiguousarray(rng.normal(size=(3000, 3)) * np.array([0.05, 0.05, 0.005]) + np.array([119320.0, 485120.0, 2.0]))
    knn = natives.knn(pts, 15)
    cov = natives.cov_knn(pts, knn)
    mine = np.zeros(6)
    for i in (0, 17, 2999):
        emu.emu_cov(_p(pts), _p(np.ascontiguousarray(knn[i])), 15, _p(mine))
        assert np.array_equal(mine, cov[i])
    # FastEigen3x3: general, planar (rank 2), diagonal, zero matrices
    special = np.array([[1, 0, 0, 2, 0, 3], [2, 0, 0, 2, 0, 1], [0, 0, 0, 0, 0, 0], [1, 1, 0, 1, 0, 0],
                        [1e-3, 2e-4, 0, 5e-4, 0, 0], [4, 0, 0, 4, 0, 4]], dtype=np.float64)
    allc = np.ascontiguousarray(np.vstack([cov, special]))
    want = natives.fast_eigen(allc)
    want[np.linalg.norm(want, axis=1) == 0] = (0, 0, 1)
    got = np.zeros_like(want)
    emu.emu_normal_from_cov(_p(allc), ctypes.c_long(len(allc)), _p(got), None)
    assert np.array_equal(got, want)
    u = rng.normal(size=(500, 3)); v = rng.normal(size=(500, 3))
    out = np.zeros(500)
    emu.emu_angle(_p(np.ascontiguousarray(u)), _p(np.ascontiguousarray(v)), ctypes.c_long(500), _p(out))
    want = np.array([R.vector_angle(a, b) for a, b in zip(u, v)])
    np.testing.assert_allclose(out, want, rtol=0, atol=1e-9)


def test_device_math_octree_pos(emu):
    rng = np.random.default_rng(4)
    pts = (rng.random(5000) * 50 + 485100).astype(np.float32)
    p4 = natives.octree_params(pts, pts, pts)
    out = np.zeros(len(pts), dtype=np.int32)
    emu.emu_octree_pos(_p(pts), ctypes.c_long(len(pts)), ctypes.c_float(p4[1]), ctypes.c_float(p4[3]), (1 << 21) - 1, _p(out))
    t = (pts - p4[1]) / p4[3]
    assert t.dtype == np.float32
    assert np.array_equal(out, np.clip(t.astype(np.int64), 0, (1 << 21) - 1))


def test_readers_on_demo_data():
    reader = ahn_utils.NPZReader(os.path.join(GOLDEN, 'demo'))
    t = reader.filter_tile('2386_9702')
    assert t['ground_surface'].shape == (500, 500) and t['ground_surface'].dtype == np.float64
    with pytest.raises(ahn_utils.AHNFileNotFoundError):
        reader.filter_tile('0000_0000')
    bld = bgt_utils.BGTPolyReader(bgt_file=os.path.join(GOLDEN, 'demo', 'bgt_buildings_demo.csv'))
    polys = bld.filter_tile('2386_9702', bgt_types=['pand'], merge=True)
    assert len(polys) == 7 and sum(len(p.exterior.coords) for p in polys) == 179
    road = bgt_utils.BGTPolyReader(bgt_file=os.path.join(GOLDEN, 'demo', 'bgt_roads_demo.csv'))
    assert len(road.filter_tile('2386_9702', bgt_types=['rijbaan_lokale_weg', 'parkeervlak', 'fietspad'])) == 15
    assert len(road.filter_tile('2386_9702', bgt_types=['fietspad'])) < 15
    if not bgt_utils.HAVE_SHAPELY:
        # no GEOS: the buffer is carried as a pending offset and evaluated by the point-in-polygon kernel
        buffered = bld.filter_tile('2386_9702', bgt_types=['pand'], offset=0.25, merge=True)
        assert len(buffered) == 7 and all(p.offset == 0.25 for p in buffered)
        assert [p.exterior.coords for p in buffered] == [p.exterior.coords for p in polys]
        assert buffered[0].buffer(-0.25).offset == 0.0 and polys[0].buffer(0) is polys[0]
    assert las_utils.get_bbox_from_tile_code('2386_9702') == ((119300, 485150), (119350, 485100))
    assert las_utils.get_tilecode_from_filename('filtered_2386_9702.laz') == '2386_9702'
    assert math_utils.get_octree_level(np.array([[0, 0, 0], [50., 20, 3]]), 0.05) == 10


def test_synthetic_generator_is_deterministic():
    a = synthetic.make_tile(5000, seed=42)
    b = synthetic.make_tile(5000, seed=42)
    assert np.array_equal(a['points'], b['points']) and a['points'].flags.f_contiguous
    assert np.all(np.abs(a['points'] * 1000 - np.round(a['points'] * 1000)) < 1e-6)
    c = synthetic.make_tile(5000, seed=43)
    assert not np.array_equal(a['points'], c['points'])
    assert len(synthetic.make_polygon_lattice(n_side=5)) == 25


def test_native_las_io_round_trip(tmp_path):
    """utils.las_io: what las_utils.read_las / label_and_save_las fall back to without laspy."""
    from urban_pointcloud_processing_b200.utils import las_io, las_utils
    rng = np.random.default_rng(3)
    P = (rng.integers(0, 50000, (5000, 3)) + np.array([119300000, 485100000, -2000])).astype(np.int32)
    scales, offsets = [0.001, 0.001, 0.001], [0.0, 0.0, 0.0]
    las = las_io.create(P, scales, offsets, point_format_id=1)
    assert las.point_format.extra_dimension_names == [] and las.header.point_count == 5000
    assert np.array_equal(las.x, P[:, 0] * 0.001 + 0.0) and las.x.dtype == np.float64
    f1 = str(tmp_path / 'filtered_2386_9702.las')
    las.write(f1)
    back = las_utils.read_las(f1)
    assert np.array_equal(back.X, P[:, 0]) and np.array_equal(back.Y, P[:, 1]) and np.array_equal(back.Z, P[:, 2])
    assert np.array_equal(back.header.scales, scales) and back.point_format.extra_dimension_names == []
    labels = rng.integers(0, 99, 5000).astype(np.uint16)
    f2 = str(tmp_path / 'labelled_2386_9702.las')
    las_utils.label_and_save_las(back, labels, f2)
    again = las_utils.read_las(f2)
    assert again.point_format.extra_dimension_names == ['label']
    assert np.array_equal(again.label, labels.astype(np.uint8)) and np.array_equal(again.X, P[:, 0])
    assert again.header.point_size == back.header.point_size       # 28 + 1 extra byte, header rewritten
    # a second labelling overwrites the existing dimension instead of adding one
    las_utils.label_and_save_las(again, labels[::-1].copy(), f2)
    third = las_utils.read_las(f2)
    assert third.point_format.extra_dimension_names == ['label'] and np.array_equal(third.label, labels[::-1].astype(np.uint8))
    assert las_utils.get_tilecode_from_filename(f2) == '2386_9702'


def _las14_bytes(fmt, n, evlr_payloads, rng):
    """A minimal LAS 1.4 file of point format 6-8 built by hand (375-byte header, no VLRs, optional EVLRs)."""
    import struct
    size = {6: 30, 7: 36, 8: 38}[fmt]
    raw = bytearray(375)
    raw[0:4] = b'LASF'; raw[24], raw[25] = 1, 4
    struct.pack_into('<H', raw, 94, 375); struct.pack_into('<I', raw, 96, 375); struct.pack_into('<I', raw, 100, 0)
    raw[104] = fmt; struct.pack_into('<H', raw, 105, size); struct.pack_into('<I', raw, 107, 0)
    struct.pack_into('<3d', raw, 131, 0.001, 0.001, 0.001); struct.pack_into('<3d', raw, 155, 0.0, 0.0, 0.0)
    struct.pack_into('<Q', raw, 247, n)
    rec = rng.integers(0, 255, (n, size), dtype=np.uint8)
    body = rec.tobytes()
    evlr = b''
    for k, payload in enumerate(evlr_payloads):
        evlr += struct.pack('<H16sHQ32s', 0, b'upcp_test', 100 + k, len(payload), b'an extended vlr') + payload
    struct.pack_into('<Q', raw, 235, 375 + len(body) if evlr_payloads else 0)
    struct.pack_into('<I', raw, 243, len(evlr_payloads))
    return bytes(raw) + body + evlr, rec, evlr


@pytest.mark.parametrize('fmt', [6, 7, 8])
@pytest.mark.parametrize('n_evlr', [0, 2])
def test_native_las14_round_trip_keeps_evlrs(tmp_path, fmt, n_evlr):
    """LAS 1.4, point formats 6-8: adding the `label` byte moves everything behind the point records; the extended
    VLRs are carried through and the header points at their new place."""
    import struct
    from urban_pointcloud_processing_b200.utils import las_io, las_utils
    rng = np.random.default_rng(fmt * 10 + n_evlr)
    blob, rec, evlr = _las14_bytes(fmt, 777, [bytes(rng.integers(0, 255, 300 + 50 * k, dtype=np.uint8)) for k in range(n_evlr)], rng)
    f1 = tmp_path / 'filtered_2386_9702.las'
    f1.write_bytes(blob)
    las = las_io.read(str(f1))
    assert las.header.point_count == 777 and las.n_evlrs == n_evlr and las.evlrs == evlr
    assert np.array_equal(las.classification, rec[:, 16])
    labels = rng.integers(0, 99, 777).astype(np.uint16)
    f2 = str(tmp_path / 'labelled_2386_9702.las')
    las_utils.label_and_save_las(las, labels, f2)
    out = open(f2, 'rb').read()
    back = las_io.read(f2)
    assert np.array_equal(back.label, labels.astype(np.uint8)) and np.array_equal(back.records[:, :rec.shape[1]], rec)
    start, count = struct.unpack_from('<Q', out, 235)[0], struct.unpack_from('<I', out, 243)[0]
    assert count == n_evlr and back.n_evlrs == n_evlr and back.evlrs == evlr
    if n_evlr:
        assert start == back.header.offset_to_points + 777 * back.header.point_size and out[start:] == evlr
    else:
        assert start == 0 and len(out) == back.header.offset_to_points + 777 * back.header.point_size


def test_native_laz_decoder_on_the_reference_demo_cloud(tmp_path):
    """utils.las_io reads .laz through the native LASzip decoder (csrc/laz.cu: POINT10 + GPSTIME11 v2, the codecs of
    the reference's demo AHN cloud, LAS 1.2 format 1).  Conformance: the decoded coordinates attain exactly the
    header's min / max, the returns histogram equals the header's, and -- test_gpu_parity -- the rasters computed
    from the decoded points equal the ahn_2386_9702.npz the reference shipped beside the cloud."""
    import struct
    from urban_pointcloud_processing_b200.utils import las_io, las_utils
    f = os.path.join(GOLDEN, 'demo', 'ahn_2386_9702.laz')
    raw = open(f, 'rb').read()
    las = las_utils.read_las(f)
    assert las.header.point_count == 43536 and las.header.point_format_id == 1 and las.records.shape == (43536, 28)
    mx = struct.unpack_from('<6d', raw, 179)
    assert (las.x.max(), las.x.min(), las.y.max(), las.y.min(), las.z.max(), las.z.min()) == mx
    assert tuple(np.bincount(las.records[:, 14] & 7, minlength=6)[1:6]) == struct.unpack_from('<5I', raw, 111)
    assert set(np.unique(las.classification)) == {1, 2, 6}
    gps = np.ascontiguousarray(las.records[:, 20:28]).view('<f8').ravel()
    assert np.isfinite(gps).all() and 5.2e5 < gps.min() and gps.max() < 5.4e5
    # a truncated stream is an error, not garbage
    bad = tmp_path / 'cut_2386_9702.laz'
    bad.write_bytes(raw[:len(raw) // 2])
    with pytest.raises(_lib.UpcpCudaError):
        las_io.read(str(bad))
    # variable-size chunks (chunk size 0xFFFFFFFF in the laszip VLR) are refused, not mis-read as one chunk
    var = bytearray(raw)
    pos = struct.unpack_from('<H', raw, 94)[0]
    for _ in range(struct.unpack_from('<I', raw, 100)[0]):
        _, _, rec, length, _ = struct.unpack_from('<H16sHH32s', raw, pos)
        if rec == 22204:
            struct.pack_into('<i', var, pos + 54 + 12, -1)
        pos += 54 + length
    (tmp_path / 'var_2386_9702.laz').write_bytes(bytes(var))
    with pytest.raises(_lib.UpcpCudaError, match='variable-size'):
        las_io.read(str(tmp_path / 'var_2386_9702.laz'))
    # written back as a plain .las (laszip VLR dropped, compression bit cleared) and read again
    out = str(tmp_path / 'plain_2386_9702.las')
    las.write(out)
    again = las_io.read(out)
    assert np.array_equal(again.records, las.records) and open(out, 'rb').read()[104] == 1


def test_bgt_point_reader_matches_reference_semantics(tmp_path):
    """utils.bgt_utils.BGTPointReader (bgt_utils.py:97-126): inclusive tile bbox with padding, type include /
    exclude lists, tuples with or without the type; and, where /root/reference is present, the reference's own
    reader on the same file."""
    import pandas as pd
    rows = [('lichtmast', 119300.0, 485100.0), ('lichtmast', 119350.0, 485150.0), ('boom', 119325.5, 485120.25),
            ('verkeersbord', 119350.5, 485120.0), ('lichtmast', 119299.5, 485149.0), ('boom', 119400.0, 485100.0)]
    path = str(tmp_path / 'bgt_points.csv')
    pd.DataFrame(rows, columns=['bgt_type', 'x', 'y']).to_csv(path, index=False)
    reader = bgt_utils.BGTPointReader(bgt_file=path)
    assert reader.filter_tile('2386_9702') == [(119300.0, 485100.0), (119350.0, 485150.0), (119325.5, 485120.25)]
    assert reader.filter_tile('2386_9702', bgt_types=['lichtmast']) == [(119300.0, 485100.0), (119350.0, 485150.0)]
    assert reader.filter_tile('2386_9702', exclude_types=['lichtmast']) == [(119325.5, 485120.25)]
    padded = reader.filter_tile('2386_9702', padding=1, return_types=True)
    assert [tuple(r) for r in padded] == [tuple(r) for r in rows[:5]]
    assert bgt_utils.get_points(path, '2386_9702', return_types=False) == reader.filter_tile('2386_9702')
    arr = bgt_utils.ArrayBGTPointReader({'2386_9702': rows})
    assert arr.filter_tile('2386_9702', bgt_types=['boom']) == [(119325.5, 485120.25), (119400.0, 485100.0)]
    assert arr.filter_tile('0000_0000') == []
    from oracle import ref_harness
    if ref_harness.available():
        ref_harness.load_reference()
        from upcp.utils.bgt_utils import BGTPointReader as RefReader
        ref = RefReader(bgt_file=path)
        for kw in ({}, {'bgt_types': ['lichtmast']}, {'exclude_types': ['boom']}, {'padding': 1}):
            assert reader.filter_tile('2386_9702', **kw) == [tuple(p) for p in ref.filter_tile('2386_9702', **kw)]


def test_tile_worker_scheduling_host_logic(monkeypatch):
    """Pipeline._run_tile_workers (the host side of 'several tiles in flight'): every tile is handed out exactly
    once, each worker thread runs inside its own persistent stream, the first exception is re-raised and stops
    the hand-out.  CUDA is stood in for by recording stubs -- the test covers the scheduling, not the kernels."""
    import contextlib
    import threading

    import torch

    from urban_pointcloud_processing_b200 import device as dev
    from urban_pointcloud_processing_b200.pipeline import Pipeline

    class FakeStream:
        def __init__(self, name):
            self.name, self.synced = name, 0

        def synchronize(self):
            self.synced += 1

    streams, entered = {}, []
    tls = threading.local()

    def worker_stream(device, name):
        return streams.setdefault(name, FakeStream(name))

    @contextlib.contextmanager
    def stream_ctx(stream):
        tls.stream = stream
        entered.append(stream.name)
        yield
        tls.stream = None

    monkeypatch.setattr(dev, 'current_device', lambda: 'cuda:0')
    monkeypatch.setattr(dev, 'worker_stream', worker_stream)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    monkeypatch.setattr(torch.cuda, 'stream', stream_ctx)
    pipe = Pipeline(processors=(), caching=False)
    for n_tiles, workers in ((11, 3), (2, 4), (0, 2), (5, 1)):
        seen, lock = [], threading.Lock()
        streams.clear(); entered.clear()

        def body(i):
            with lock:
                seen.append((i, tls.stream.name))
        pipe._run_tile_workers(n_tiles, workers, body)
        assert sorted(i for i, _ in seen) == list(range(n_tiles))
        assert sorted(entered) == list(range(workers)) and all(s.synced == 1 for s in streams.values())
        assert {name for _, name in seen} <= set(range(workers))

    done = []

    def failing(i):
        if i == 3:
            raise RuntimeError('tile 3 is broken')
        done.append(i)
    with pytest.raises(RuntimeError, match='tile 3'):
        pipe._run_tile_workers(200, 2, failing)
    assert len(done) < 199                                    # the hand-out stopped early
    # argument checks happen before any device work
    with pytest.raises(ValueError):
        pipe.process_clouds(['a', 'b'], [None], [None, None])
    with pytest.raises(ValueError):
        pipe.process_tiles_device(['a'], [], [None])


def test_ahn_reader_cache_is_safe_under_concurrent_tiles(tmp_path, monkeypatch):
    """One NPZReader shared by worker threads that label DIFFERENT tiles (process_clouds /
    process_folder(workers=W)): every filter_tile / interpolation-cache access must deliver the tile
    that was asked for, however the (slow) file loads interleave."""
    import threading
    import time
    codes = ['1000_2000', '1001_2000', '1002_2000']
    for k, code in enumerate(codes):
        np.savez(tmp_path / f'ahn_{code}.npz', x=np.arange(4.0), y=np.arange(4.0)[::-1].copy(),
                 ground=np.full((4, 4), float(k), dtype=np.float16), building=np.full((4, 4), 10.0 + k, dtype=np.float16))
    real_load = ahn_utils.load_ahn_tile

    def slow_load(fname):
        tile = real_load(fname)
        time.sleep(0.002 * (1 + (hash(fname) & 3)))         # uneven load times: the interleavings that used to mix tiles
        return tile
    monkeypatch.setattr(ahn_utils, 'load_ahn_tile', slow_load)
    reader = ahn_utils.NPZReader(tmp_path, caching=True)
    bad = []

    def worker(k):
        for it in range(60):
            code = codes[(k + it) % len(codes)] if it % 7 == 0 else codes[k]
            want = float(codes.index(code))
            tile = reader.filter_tile(code)
            if tile['ground_surface'][0, 0] != want or tile['building_surface'][0, 0] != want + 10.0:
                bad.append((k, it, code))
            if it % 5 == 0:
                reader._clear_cache()                        # what cache_interpolator of another tile does
    threads = [threading.Thread(target=worker, args=(k,)) for k in range(3)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not bad, bad[:5]
    # caching off: still the right tile, nothing kept
    reader.set_caching(False)
    assert reader.filter_tile(codes[1])['ground_surface'][0, 0] == 1.0 and len(reader._tiles) == 0


def test_per_thread_diagnostics():
    """one processor object is driven by several worker threads; what a call leaves behind for
    inspection belongs to the calling thread (abstract_processor.PerThread)."""
    import threading
    from urban_pointcloud_processing_b200.abstract_processor import PerThread

    class P:
        last = PerThread(list)
        level = PerThread(lambda: None)

    p, q = P(), P()
    p.last = ['main']
    seen = {}

    def work(k):
        assert p.last == [] and p.level is None            # a fresh thread reads the default, not main's value
        p.last = [k]
        p.level = k
        barrier.wait()
        seen[k] = (list(p.last), p.level)

    barrier = threading.Barrier(2)
    ts = [threading.Thread(target=work, args=(k,)) for k in (1, 2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert seen == {1: ([1], 1), 2: ([2], 2)}
    assert p.last == ['main'] and q.last == [] and p.level is None
    # objects that carry such diagnostics still pickle / deep-copy (the per-thread store is dropped)
    import copy
    import pickle
    from urban_pointcloud_processing_b200.fusion import CarFuser
    from urban_pointcloud_processing_b200.pipeline import Pipeline
    car = CarFuser(40, ahn_reader=None, bgt_reader=None, params={'min_height': 1.0})
    car.last_objects = [('ring', 0.0, 1.0)]
    pipe = Pipeline(processors=(car,), caching=False)
    pipe.last_timings = [('CarFuser', 0.1)]
    for clone in (pickle.loads(pickle.dumps(pipe)), copy.deepcopy(pipe)):
        assert clone.last_timings == [] and clone.processors[0].last_objects == []
        assert clone.processors[0].params == {'min_height': 1.0} and clone.processors[0].label == 40
    assert pipe.last_timings == [('CarFuser', 0.1)] and car.last_objects == [('ring', 0.0, 1.0)]


def test_rings_fast_path_equals_generic_path():
    """device.rings_from_polygons flattens the readers' own RingPolygon objects without per-ring conversion; the
    ring arrays must be exactly those of the generic (shapely-like .exterior.coords / .interiors) path."""
    from urban_pointcloud_processing_b200 import device as dev
    polys = synthetic.make_polygon_lattice(n_side=9, seed=11)
    assert any(len(p.interiors) for p in polys)

    class Generic:                      # hides the type: forces the generic path
        def __init__(self, p):
            self.exterior, self.interiors = p.exterior, p.interiors
    fast = dev.rings_from_polygons(polys)
    slow = dev.rings_from_polygons([Generic(p) for p in polys])
    for a, b in zip(fast, slow):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    assert fast[1][0] == 0 and fast[1][-1] == fast[0].shape[0] and len(fast[1]) == len(fast[2]) + 1
    empty = dev.rings_from_polygons([])
    assert empty[0].shape == (0, 2) and list(empty[1]) == [0]


def test_laz_encoder_reproduces_the_reference_demo_cloud_byte_for_byte(tmp_path):
    """`.laz` write-back (las_utils.label_and_save_las_as_laz, reference utils/las_utils.py:133-164; laspy writes through
    lazrs).  The native LASzip encoder is pinned on the reference's own data: re-encoding the decoded demo cloud gives
    the file's compressed point data section -- chunk-table position, chunk, arithmetic-coded chunk table -- byte for
    byte, and a rewritten file of the same size that decodes to the same records."""
    import struct
    from urban_pointcloud_processing_b200.utils import las_io
    src = os.path.join(GOLDEN, 'demo', 'ahn_2386_9702.laz')
    raw = open(src, 'rb').read()
    las = las_io.read(src)
    off = struct.unpack_from('<I', raw, 96)[0]
    payload, body = las_io._encode_laz(las.records, las.header.point_format_id, off)
    assert body == raw[off:]
    orig_payload = next(v[3] for v in _raw_vlrs(raw) if v[1] == 22204)
    assert payload[:4] == orig_payload[:4] and payload[12:16] == orig_payload[12:16] and payload[32:] == orig_payload[32:]
    out = tmp_path / 'again.laz'
    las.write(str(out))
    assert os.path.getsize(out) == len(raw)
    back = las_io.read(str(out))
    assert np.array_equal(back.records, las.records)


def _raw_vlrs(raw):
    import struct
    pos = struct.unpack_from('<H', raw, 94)[0]
    out = []
    for _ in range(struct.unpack_from('<I', raw, 100)[0]):
        _, user, rec, length, _ = struct.unpack_from('<H16sHH32s', raw, pos)
        out.append((user.rstrip(b'\0'), rec, length, raw[pos + 54:pos + 54 + length]))
        pos += 54 + length
    return out


@pytest.mark.parametrize('fmt', [0, 1, 2, 3])
def test_laz_round_trip_all_item_codecs(tmp_path, fmt):
    """Encode -> decode is the identity for every version-2 item codec (POINT10, GPSTIME11, RGB12, extra BYTEs) over
    several chunks: return-number contexts, intensity / classification / scan-angle / user-data / point-source changes,
    GPS times that repeat, jump, run backwards and interleave four sequences, grey and coloured RGB, and the `label`
    byte that label_and_save_las adds; through the public route las_utils.label_and_save_las -> read_las."""
    from urban_pointcloud_processing_b200.utils import las_io, las_utils
    rng = np.random.default_rng(100 + fmt)
    n = 123457                                                       # three chunks of 50 000, the last one ragged
    walk = np.cumsum(rng.integers(-40, 41, size=(n, 3)), axis=0) + rng.integers(-10**6, 10**6, size=3)
    walk[::997] += rng.integers(-10**8, 10**8, size=(len(walk[::997]), 3))         # jumps
    las = las_io.create(walk.astype(np.int32), [0.001, 0.001, 0.001], [1e5, 4e5, 0.0], point_format_id=fmt)
    rec = las.records
    rec[:, 12:14] = rng.integers(0, 65536, n).astype('<u2').reshape(-1, 1).view(np.uint8) * (rng.random(n) < 0.3).reshape(-1, 1)
    nret = rng.integers(1, 6, n); ret = np.minimum(rng.integers(1, 6, n), nret)
    rec[:, 14] = (ret | (nret << 3) | (rng.integers(0, 2, n) << 6) | (rng.integers(0, 2, n) << 7)).astype(np.uint8)
    rec[:, 15] = rng.choice(np.array([1, 2, 6, 9, 26], dtype=np.uint8), n)
    rec[:, 16] = rng.integers(-30, 31, n).astype(np.int8).view(np.uint8)
    rec[:, 17] = (rng.random(n) < 0.05) * rng.integers(0, 256, n)
    rec[:, 18:20] = rng.choice(np.array([7, 7, 7, 300, 65535], dtype='<u2'), n).reshape(-1, 1).view(np.uint8)
    col = 20
    if fmt in (1, 3):
        seq = rng.integers(0, 4, n) * (rng.random(n) < 0.02)          # four interleaved time sequences
        t = np.cumsum(rng.choice([0.0, 1e-5, 1e-5, 1e-5, 3e-5, -2e-5, 7.5e-3], n)) + seq * 1e6
        t[rng.random(n) < 0.001] += 1e9
        rec[:, col:col + 8] = t.astype('<f8').reshape(-1, 1).view(np.uint8)
        col += 8
    if fmt in (2, 3):
        rgb = rng.integers(0, 65536, (n, 3)).astype('<u2')
        grey = rng.random(n) < 0.5
        rgb[grey] = rgb[grey, :1]
        rgb[rng.random(n) < 0.3] &= 0xFF00
        rec[:, col:col + 6] = rgb.view(np.uint8)
    labels = rng.choice(np.array([0, 1, 9, 10, 40], dtype=np.uint8), n)
    out = tmp_path / f'fmt{fmt}.laz'
    las_utils.label_and_save_las(las, labels, str(out))
    raw = open(out, 'rb').read()
    assert raw[104] == (fmt | 0x80) and any(v[1] == 22204 for v in _raw_vlrs(raw))
    assert os.path.getsize(out) < 0.8 * las.records.size             # it IS compressed
    back = las_utils.read_las(str(out))
    assert np.array_equal(back.records, las.records) and np.array_equal(back.label, labels)
    assert back.header.point_format_id == fmt and 'label' in back.point_format.extra_dimension_names
    # a plain .las of the same cloud still round-trips too
    las.write(str(tmp_path / 'plain.las'))
    assert np.array_equal(las_io.read(str(tmp_path / 'plain.las')).records, las.records)
    # tiny clouds: 0 and 1 point
    for k in (0, 1):
        small = las_io.create(walk[:k].astype(np.int32), [0.001] * 3, [0.0] * 3, point_format_id=fmt)
        small.write(str(tmp_path / f'small{k}.laz'))
        assert np.array_equal(las_io.read(str(tmp_path / f'small{k}.laz')).records, small.records)
