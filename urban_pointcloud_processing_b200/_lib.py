"""ctypes binding of ``libupcp_cuda.so`` (the C ABI declared in ``include/upcp_cuda.h``).

The product path has NO CPU fallback: if the shared library has not been built
(``python __graft_entry__.py`` / ``__graft_entry__.build()``) importing any compute
entry point raises, and if no CUDA device is present every compute call raises.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_uint8, c_void_p)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libupcp_cuda.so')

UPCP_OK = 0
ERROR_NAMES = {-1: 'UPCP_E_INVALID', -2: 'UPCP_E_CUDA', -3: 'UPCP_E_WORKSPACE',
               -4: 'UPCP_E_CAPACITY', -5: 'UPCP_E_INTERNAL'}

CLS_ABS_LT, CLS_LT, CLS_CAP_LE, CLS_BAND_OC, CLS_BAND_CC, CLS_BELOW_NEG = range(6)


class UpcpCudaError(RuntimeError):
    """Raised when a upcp_cuda entry point returns a non-zero status."""


class UpcpCudaMissing(ImportError):
    """Raised when libupcp_cuda.so has not been built."""


class Tuning(Structure):
    """``upcp_tuning`` of include/upcp_cuda.h (explicit, process-wide kernel tuning knobs)."""
    _fields_ = [('knn_tiers', c_int32), ('knn_tier1_max_block', c_int32), ('knn_tier2_max_block', c_int32),
                ('knn_coarse', c_int32), ('grow_local_chain', c_int32), ('grow_ctas_per_sm', c_int32),
                ('knn_guess_tenths', c_int32), ('knn_guess_max_pct', c_int32), ('reserved', c_int32 * 8)]


_PROTOTYPES = {
    # name: (restype, [argtypes])
    'upcp_abi_version': (c_int, []),
    'upcp_last_error': (c_char_p, []),
    'upcp_device_count': (c_int, []),
    'upcp_launch_count': (c_int64, []),
    'upcp_prof_enable': (c_int, [c_int]),
    'upcp_prof_reset': (c_int, []),
    'upcp_prof_read': (c_int, [c_int, POINTER(c_double), POINTER(c_int64), POINTER(c_int64)]),
    'upcp_prof_slot_name': (c_char_p, [c_int]),
    'upcp_tuning_default': (None, [POINTER(Tuning)]),
    'upcp_tuning_get': (c_int, [POINTER(Tuning)]),
    'upcp_tuning_set': (c_int, [POINTER(Tuning)]),
    'upcp_points_from_scaled_int32': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_double),
                                              POINTER(c_double), c_void_p, c_void_p, c_void_p, c_void_p]),
    'upcp_points_aos_to_soa': (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'upcp_bbox_ws_bytes': (c_size_t, [c_int64]),
    'upcp_bbox': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'upcp_labels_widen_u8': (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    'upcp_labels_widen_u16': (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    'upcp_labels_narrow_u8': (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    'upcp_labels_narrow_u16': (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    'upcp_mask_build': (c_int, [c_void_p, c_int64, c_void_p, c_int, POINTER(c_int32), c_int,
                                POINTER(c_int32), c_int, c_void_p, c_void_p]),
    'upcp_mask_combine': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    'upcp_labels_assign': (c_int, [c_void_p, c_int64, c_void_p, c_int32, c_void_p]),
    'upcp_mask_count': (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    'upcp_label_hist': (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    'upcp_raster_lookup': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p,
                                   c_int, c_void_p, c_void_p, c_void_p]),
    'upcp_raster_classify': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int,
                                     c_void_p, c_int, c_void_p, c_int, c_double, c_double, c_void_p,
                                     c_void_p, c_int32, c_void_p]),
    'upcp_pip_plan_create': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                     POINTER(c_void_p)]),
    'upcp_pip_plan_create_buffered': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_double,
                                              POINTER(c_void_p)]),
    'upcp_pip_plan_bytes': (c_size_t, [c_void_p]),
    'upcp_pip_plan_upload': (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    'upcp_pip_plan_destroy': (None, [c_void_p]),
    'upcp_pip_query': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    'upcp_pip_query_buffered': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    'upcp_octree_params': (c_int, [POINTER(c_double), c_int, POINTER(c_float)]),
    'upcp_octree_level': (c_int, [POINTER(c_double), c_int, c_double]),
    'upcp_lcc_ws_bytes': (c_size_t, [c_int64, c_int64, c_int]),
    'upcp_lcc': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, POINTER(c_float), c_int, c_int,
                         c_int32, c_void_p, c_int32, c_double, c_void_p, c_void_p, c_void_p,
                         c_void_p, c_size_t, c_void_p]),
    'upcp_knn_ws_bytes': (c_size_t, [c_int64, c_int64, c_int]),
    'upcp_knn_normals': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                 POINTER(c_double), c_int, c_int, c_double, c_double, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    'upcp_knn_normals_sorted': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                        POINTER(c_double), c_int, c_int, c_double, c_double, c_void_p,
                                        c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_void_p,
                                        c_size_t, c_void_p]),
    'upcp_mask_take': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    'upcp_mask_put': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    'upcp_region_grow_ws_bytes': (c_size_t, [c_int64]),
    'upcp_region_grow': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_double,
                                 c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'upcp_region_grow_begin': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_double,
                                       c_void_p, c_size_t, c_void_p]),
    'upcp_region_grow_resume': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                        c_int64, c_double, c_void_p, c_size_t, c_void_p]),
    'upcp_region_grow_end': (c_int, [c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'upcp_radius_ws_bytes': (c_size_t, [c_int64, c_int64]),
    'upcp_radius_prepare': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, POINTER(c_double), c_double,
                                    c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'upcp_radius_grow': (c_int, [c_int64, c_int64, POINTER(c_double), c_double, c_void_p, c_double, c_void_p, c_int64,
                                 c_void_p, c_size_t, c_void_p]),
    'upcp_radius_region': (c_int, [c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'upcp_radius_neighbours': (c_int, [c_int64, c_int64, POINTER(c_double), c_double, c_void_p, c_int64, c_void_p, c_void_p,
                                       c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    'upcp_compact_ws_bytes': (c_size_t, [c_int64]),
    'upcp_mask_scatter': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'upcp_mask_gather': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                 c_void_p]),
    'upcp_component_stats_ws_bytes': (c_size_t, [c_int64]),
    'upcp_component_stats': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_size_t, c_void_p]),
    'upcp_gather_xy': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    'upcp_component_mark': (c_int, [c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_void_p]),
    'upcp_prism_clip': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int32, c_void_p, c_void_p]),
    'upcp_laz_decode': (c_int, [c_void_p, c_int32, c_void_p, c_int64, c_int64, c_int64, c_int32, c_void_p]),
    'upcp_laz_encode': (c_int, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_int64, c_void_p, POINTER(c_int32), c_void_p, c_int64,
                              POINTER(c_int64)]),
    'upcp_interpolate_ws_bytes': (c_size_t, [c_int64]),
    'upcp_interpolate': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_double), c_void_p, c_void_p, c_int64,
                                 c_int, c_int, c_double, c_double, c_double, c_double, c_double, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    'upcp_synth_tile': (c_int, [c_int64, ctypes.c_uint64, POINTER(c_double), POINTER(c_double), c_void_p, c_int32,
                                c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    'upcp_sort_ws_bytes': (c_size_t, [c_int64]),
    'upcp_sort_pairs_u32': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p,
                                    c_size_t, c_void_p]),
    'upcp_sort_pairs_u64': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p,
                                    c_size_t, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None


def load():
    """Load the shared library (once) and attach prototypes. Raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UpcpCudaMissing(
            f'{LIB_PATH} not found: build the CUDA extension first '
            '(python -c "import __graft_entry__ as g; g.build()"). There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError if a declared symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status, what=''):
    if status != UPCP_OK:
        msg = load().upcp_last_error()
        msg = msg.decode() if msg else ''
        raise UpcpCudaError(f'{what or "upcp_cuda"} failed: {ERROR_NAMES.get(status, status)}: {msg}')


def require_cuda():
    """Fail loudly when there is no CUDA device -- the product has no CPU path."""
    lib = load()
    if lib.upcp_device_count() < 1:
        raise UpcpCudaError('no CUDA device visible: upcp_cuda has no CPU fallback')
    return lib


PROF_SLOTS = 9


def tuning(**changes):
    """Read the process-wide tuning knobs; with keyword arguments, change those and install the result
    (``upcp_tuning_set``).  Returns the knobs in force as a dict."""
    lib = load()
    t = Tuning()
    check(lib.upcp_tuning_get(ctypes.byref(t)), 'upcp_tuning_get')
    if changes:
        for k, v in changes.items():
            if k not in dict(Tuning._fields_) or k == 'reserved':
                raise KeyError(k)
            setattr(t, k, int(v))
        check(lib.upcp_tuning_set(ctypes.byref(t)), 'upcp_tuning_set')
    return {k: getattr(t, k) for k, _ in Tuning._fields_ if k != 'reserved'}


def tuning_reset():
    lib = load()
    t = Tuning()
    lib.upcp_tuning_default(ctypes.byref(t))
    check(lib.upcp_tuning_set(ctypes.byref(t)), 'upcp_tuning_set')


def prof_enable(on=True):
    load().upcp_prof_enable(int(bool(on)))
    load().upcp_prof_reset()


def prof_read():
    """{slot name: (total_ms, launches, units)} since the last reset (synchronises)."""
    lib = load()
    out = {}
    for slot in range(PROF_SLOTS):
        ms, cnt, units = c_double(), c_int64(), c_int64()
        check(lib.upcp_prof_read(slot, ctypes.byref(ms), ctypes.byref(cnt), ctypes.byref(units)), 'upcp_prof_read')
        out[lib.upcp_prof_slot_name(slot).decode()] = (ms.value, cnt.value, units.value)
    return out


def launch_count():
    return int(load().upcp_launch_count())
