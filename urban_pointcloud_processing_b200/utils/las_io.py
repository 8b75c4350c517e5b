"""Minimal native LAS reader / writer (uncompressed ``.las``, point formats 0-3 and 6-8).

The reference reads and writes its clouds with ``laspy`` (``las_utils.read_las`` /
``label_and_save_las``, reference src/upcp/utils/las_utils.py:118-130; ``Pipeline.process_file``,
pipeline.py:99-137).  ``laspy`` is used when it is installed (it is required for ``.laz``); this
module covers plain ``.las`` without it and -- the point of the exercise -- exposes the RAW
integer coordinates, so that the pipeline uploads 12 B per point and rebuilds
``x = X * scale + offset`` in fp64 on the device (``upcp_points_from_scaled_int32``), bit for bit
what ``laspy``'s scaled view computes on the host.

Only what the labelling pipeline touches is modelled: header, VLRs (kept verbatim), point
records (kept verbatim), the scaled coordinates and the ``label`` extra byte.
"""
import struct

import numpy as np

_STD_SIZE = {0: 20, 1: 28, 2: 26, 3: 34, 4: 57, 5: 63, 6: 30, 7: 36, 8: 38, 9: 59, 10: 67}
_LASZIP_USER, _LASZIP_RECORD = b'laszip encoded', 22204
_EXTRA_BYTES_USER = b'LASF_Spec'
_EXTRA_BYTES_RECORD = 4
_EB_SIZES = {1: 1, 2: 1, 3: 2, 4: 2, 5: 4, 6: 4, 7: 8, 8: 8, 9: 4, 10: 8}


class LasHeader:
    def __init__(self):
        self.version = (1, 2)
        self.point_format_id = 0
        self.point_size = 20
        self.point_count = 0
        self.scales = np.array([0.001, 0.001, 0.001])
        self.offsets = np.zeros(3)
        self.raw = b''                 # header bytes as read (re-used on write)
        self.header_size = 227
        self.offset_to_points = 227


class _PointFormat:
    def __init__(self, names):
        self.extra_dimension_names = list(names)


class LasData:
    """The subset of ``laspy.LasData`` the pipeline uses: ``x y z`` (scaled fp64), ``X Y Z`` (raw
    int32), ``header``, ``point_format.extra_dimension_names``, ``label``, ``write``."""

    def __init__(self, header, vlrs, records, extra_dims, evlrs=b'', n_evlrs=0):
        self.header = header
        self.evlrs = evlrs                    # LAS 1.4 extended VLRs, kept verbatim (they follow the point records)
        self.n_evlrs = n_evlrs
        self.vlrs = vlrs                      # list of (user_id, record_id, description, payload)
        self.records = records                # uint8 [n, point_size]
        self._extra = extra_dims              # list of (name, numpy dtype, byte offset in record)
        self.point_format = _PointFormat([e[0] for e in extra_dims])

    def _ints(self, k):
        return np.ascontiguousarray(self.records[:, 4 * k:4 * k + 4]).view('<i4').ravel()

    X = property(lambda self: self._ints(0))
    Y = property(lambda self: self._ints(1))
    Z = property(lambda self: self._ints(2))
    # laspy's ScaledArrayView: array * scale + offset in float64
    x = property(lambda self: self.X * self.header.scales[0] + self.header.offsets[0])
    y = property(lambda self: self.Y * self.header.scales[1] + self.header.offsets[1])
    z = property(lambda self: self.Z * self.header.scales[2] + self.header.offsets[2])

    @property
    def classification(self):
        """ASPRS class of every point: 5 bits of byte 15 (formats 0-5), byte 16 (formats 6-10)."""
        if self.header.point_format_id >= 6:
            return self.records[:, 16].copy()
        return self.records[:, 15] & 0x1f

    def _extra_view(self, name):
        for nm, dt, off in self._extra:
            if nm == name:
                return np.ascontiguousarray(self.records[:, off:off + dt.itemsize]).view(dt).ravel()
        raise AttributeError(name)

    @property
    def label(self):
        return self._extra_view('label')

    @label.setter
    def label(self, values):
        values = np.asarray(values)
        if len(values) != self.header.point_count:
            raise ValueError('label: wrong length')
        if 'label' not in self.point_format.extra_dimension_names:
            self.add_extra_dim('label', np.uint8, 'Labels')
        for nm, dt, off in self._extra:
            if nm == 'label':
                self.records[:, off:off + dt.itemsize] = values.astype(dt).reshape(-1, 1).view(np.uint8)

    def add_extra_dim(self, name, dtype=np.uint8, description=''):
        dt = np.dtype(dtype).newbyteorder('<')
        n = self.records.shape[0]
        off = self.records.shape[1]
        self.records = np.concatenate([self.records, np.zeros((n, dt.itemsize), dtype=np.uint8)], axis=1)
        self._extra.append((name, dt, off))
        self.point_format.extra_dimension_names.append(name)
        self.header.point_size = self.records.shape[1]
        code = {np.dtype('u1'): 1, np.dtype('i1'): 2, np.dtype('<u2'): 3, np.dtype('<i2'): 4, np.dtype('<u4'): 5,
                np.dtype('<i4'): 6, np.dtype('<u8'): 7, np.dtype('<i8'): 8, np.dtype('<f4'): 9, np.dtype('<f8'): 10}[dt]
        desc = bytearray(192)
        desc[2] = code
        desc[4:4 + min(len(name), 32)] = name.encode()[:32]
        desc[160:160 + min(len(description), 32)] = description.encode()[:32]
        for i, v in enumerate(self.vlrs):
            if v[0] == _EXTRA_BYTES_USER and v[1] == _EXTRA_BYTES_RECORD:
                self.vlrs[i] = (v[0], v[1], v[2], v[3] + bytes(desc))
                return
        self.vlrs.append((_EXTRA_BYTES_USER, _EXTRA_BYTES_RECORD, b'Extra bytes', bytes(desc)))

    def write(self, path, compress=None):
        """Write the cloud; ``compress`` (default: by the ``.laz`` extension, as laspy does) LASzip-compresses the
        point records through the native encoder (``upcp_laz_encode``: point formats 0-3; other formats are
        written uncompressed)."""
        h = self.header
        if compress is None:
            compress = str(path).lower().endswith('.laz')
        compress = bool(compress) and h.point_format_id <= 3
        raw = bytearray(h.raw[:h.header_size])
        vlrs = [v for v in self.vlrs if not (v[0] == _LASZIP_USER and v[1] == _LASZIP_RECORD)]
        if compress:
            vlrs = vlrs + [(_LASZIP_USER, _LASZIP_RECORD, b'upcp-b200 laszip', bytes(34 + 6 * 4))]       # sized below
        def blob(vs):
            out = b''
            for user, rec, desc, payload in vs:
                out += struct.pack('<H16sHH32s', 0, user.ljust(16, b'\0'), rec, len(payload), desc.ljust(32, b'\0')) + payload
            return out
        if compress:
            # the VLR's length depends on the number of items only: size it, then encode against the final offset
            n_items = 1 + (h.point_format_id in (1, 3)) + (h.point_format_id in (2, 3)) + \
                (self.records.shape[1] > _STD_SIZE[h.point_format_id])
            vlrs[-1] = vlrs[-1][:3] + (bytes(34 + 6 * n_items),)
            offset_to_points = h.header_size + len(blob(vlrs))
            payload, body = _encode_laz(self.records, h.point_format_id, offset_to_points)
            assert len(payload) == 34 + 6 * n_items
            vlrs[-1] = vlrs[-1][:3] + (payload,)
        vlr_blob = blob(vlrs)
        offset_to_points = h.header_size + len(vlr_blob)
        struct.pack_into('<I', raw, 96, offset_to_points)
        struct.pack_into('<I', raw, 100, len(vlrs))
        raw[104] = (h.point_format_id & 0x3f) | (0x80 if compress else 0)
        struct.pack_into('<H', raw, 105, self.records.shape[1])
        if not compress:
            body = np.ascontiguousarray(self.records).tobytes()
        if h.version >= (1, 4) and h.header_size >= 375:
            # LAS 1.4: the extended VLRs follow the point records, whose size changes with the `label` byte --
            # carry them through and point the header at where they now start
            struct.pack_into('<Q', raw, 235, offset_to_points + len(body) if self.n_evlrs else 0)
            struct.pack_into('<I', raw, 243, self.n_evlrs)
        with open(path, 'wb') as f:
            f.write(bytes(raw))
            f.write(vlr_blob)
            f.write(body)
            if self.n_evlrs:
                f.write(self.evlrs)


def read(path):
    """Read an uncompressed LAS file."""
    with open(path, 'rb') as f:
        data = f.read()
    if data[:4] != b'LASF':
        raise ValueError(f'{path}: not a LAS file')
    h = LasHeader()
    h.version = (data[24], data[25])
    h.header_size = struct.unpack_from('<H', data, 94)[0]
    h.offset_to_points = struct.unpack_from('<I', data, 96)[0]
    n_vlrs = struct.unpack_from('<I', data, 100)[0]
    fmt = data[104]
    compressed = bool(fmt & 0x80) or bool(fmt & 0x40)
    h.point_format_id = fmt & 0x3f
    h.point_size = struct.unpack_from('<H', data, 105)[0]
    legacy_count = struct.unpack_from('<I', data, 107)[0]
    h.scales = np.array(struct.unpack_from('<3d', data, 131))
    h.offsets = np.array(struct.unpack_from('<3d', data, 155))
    h.point_count = legacy_count
    if h.version >= (1, 4) and h.header_size >= 375:
        count64 = struct.unpack_from('<Q', data, 247)[0]
        if count64:
            h.point_count = count64
    h.raw = data[:h.header_size]
    vlrs = []
    pos = h.header_size
    for _ in range(n_vlrs):
        _, user, rec, length, desc = struct.unpack_from('<H16sHH32s', data, pos)
        payload = data[pos + 54:pos + 54 + length]
        vlrs.append((user.rstrip(b'\0'), rec, desc.rstrip(b'\0'), payload))
        pos += 54 + length
    if h.point_format_id not in _STD_SIZE:
        raise ValueError(f'{path}: unsupported point format {h.point_format_id}')
    std = _STD_SIZE[h.point_format_id]
    extra = []
    off = std
    for user, rec, _, payload in vlrs:
        if user == _EXTRA_BYTES_USER and rec == _EXTRA_BYTES_RECORD:
            for k in range(len(payload) // 192):
                d = payload[192 * k:192 * (k + 1)]
                code = d[2]
                name = d[4:36].split(b'\0')[0].decode(errors='replace')
                if code == 0:
                    size = d[3]
                    dt = np.dtype(('V', size))
                else:
                    dt = {1: 'u1', 2: 'i1', 3: '<u2', 4: '<i2', 5: '<u4', 6: '<i4', 7: '<u8', 8: '<i8', 9: '<f4',
                          10: '<f8'}.get(code)
                    if dt is None:
                        break
                    dt = np.dtype(dt)
                extra.append((name, dt, off))
                off += dt.itemsize
    n = int(h.point_count)
    if compressed:
        records, vlrs = _decode_laz(path, data, h, vlrs, n)
    else:
        body = np.frombuffer(data, dtype=np.uint8, count=n * h.point_size, offset=h.offset_to_points)
        records = body.reshape(n, h.point_size).copy()
    extra = [e for e in extra if e[2] + e[1].itemsize <= h.point_size and e[1].kind != 'V']
    evlrs, n_evlrs = b'', 0
    if h.version >= (1, 4) and h.header_size >= 375:
        start = struct.unpack_from('<Q', data, 235)[0]
        n_evlrs = struct.unpack_from('<I', data, 243)[0]
        if n_evlrs and h.offset_to_points + n * h.point_size <= start <= len(data):
            evlrs = data[start:]
        else:
            n_evlrs = 0
    return LasData(h, vlrs, records, extra, evlrs, n_evlrs)



def _encode_laz(records, point_format_id, offset_to_points, chunk_size=50000):
    """Uncompressed records -> (payload of the laszip VLR, point data section) through ``upcp_laz_encode``."""
    import ctypes

    from .. import _lib
    lib = _lib.load()
    rec = np.ascontiguousarray(records, dtype=np.uint8)
    n, size = rec.shape
    vlr = np.zeros(58, dtype=np.uint8)
    vlr_bytes = ctypes.c_int32(0)
    out_bytes = ctypes.c_int64(0)
    cap = int(n * size * 0.6) + 4096
    for _ in range(2):
        out = np.empty(cap, dtype=np.uint8)
        rc = lib.upcp_laz_encode(rec.ctypes.data_as(ctypes.c_void_p), n, size, int(point_format_id), int(chunk_size),
                                 int(offset_to_points), vlr.ctypes.data_as(ctypes.c_void_p), ctypes.byref(vlr_bytes),
                                 out.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(out_bytes))
        if rc == -3 and out_bytes.value > cap:          # UPCP_E_WORKSPACE: the size needed is in out_bytes
            cap = int(out_bytes.value)
            continue
        _lib.check(rc, 'upcp_laz_encode')
        return vlr[:vlr_bytes.value].tobytes(), out[:out_bytes.value].tobytes()
    raise RuntimeError('upcp_laz_encode: output size changed between two calls')


def _decode_laz(path, data, h, vlrs, n):
    """LASzip-compressed point data -> uncompressed records through the native decoder (``upcp_laz_decode``,
    csrc/laz.cu).  The result is an ordinary in-memory LAS: the laszip VLR is dropped and the compression bits
    of the header's point format byte are cleared, so ``write`` produces a plain .las."""
    import ctypes

    from .. import _lib
    lib = _lib.load()
    payload = next((v[3] for v in vlrs if v[0] == _LASZIP_USER and v[1] == _LASZIP_RECORD), None)
    if payload is None:
        raise ValueError(f'{path}: compressed point data without a laszip VLR')
    records = np.empty((n, h.point_size), dtype=np.uint8)
    body = np.frombuffer(data, dtype=np.uint8, offset=h.offset_to_points)
    vlr = np.frombuffer(payload, dtype=np.uint8)
    _lib.check(lib.upcp_laz_decode(vlr.ctypes.data_as(ctypes.c_void_p), len(payload), body.ctypes.data_as(ctypes.c_void_p),
                                   len(body), int(h.offset_to_points), n, int(h.point_size),
                                   records.ctypes.data_as(ctypes.c_void_p)), f'upcp_laz_decode({path})')
    raw = bytearray(h.raw)
    raw[104] = h.point_format_id
    h.raw = bytes(raw)
    return records, [v for v in vlrs if not (v[0] == _LASZIP_USER and v[1] == _LASZIP_RECORD)]


def create(points_int, scales, offsets, point_format_id=1):
    """A new LAS 1.2 cloud from raw integer coordinates (tests / synthetic data)."""
    points_int = np.asarray(points_int, dtype='<i4')
    n = len(points_int)
    h = LasHeader()
    h.point_format_id = point_format_id
    h.point_size = _STD_SIZE[point_format_id]
    h.point_count = n
    h.scales = np.asarray(scales, dtype=np.float64)
    h.offsets = np.asarray(offsets, dtype=np.float64)
    raw = bytearray(227)
    raw[0:4] = b'LASF'
    raw[24], raw[25] = 1, 2
    raw[26:26 + 9] = b'upcp-b200'
    struct.pack_into('<H', raw, 94, 227)
    struct.pack_into('<I', raw, 96, 227)
    raw[104] = point_format_id
    struct.pack_into('<H', raw, 105, h.point_size)
    struct.pack_into('<I', raw, 107, n)
    struct.pack_into('<3d', raw, 131, *h.scales)
    struct.pack_into('<3d', raw, 155, *h.offsets)
    if n:
        xyz = points_int * h.scales + h.offsets
        mx, mn = xyz.max(axis=0), xyz.min(axis=0)
        struct.pack_into('<6d', raw, 179, mx[0], mn[0], mx[1], mn[1], mx[2], mn[2])
    h.raw = bytes(raw)
    records = np.zeros((n, h.point_size), dtype=np.uint8)
    records[:, 0:12] = np.ascontiguousarray(points_int).view(np.uint8).reshape(n, 12)
    return LasData(h, [], records, [])
