"""Clustering-based region growing (drop-in for
``upcp.region_growing.LabelConnectedComp``, reference
src/upcp/region_growing/label_connected_comp.py:18-231).

The reference hands a float32 copy of the masked cloud to CloudCompare
(``pycc.ccPointCloud`` + ``cccorelib...labelConnectedComponents(cloud, level)``) and
then filters / fills components with Python loops.  Here the whole chain -- float32
octree cell assignment, voxel-key radix sort, 26-connectivity union-find over occupied
cells, component-size and seed histograms, per-point scatter -- is one call into the
CUDA extension (``upcp_lcc``)."""
import ctypes
import logging

import numpy as np
import torch

from .. import _lib
from .. import device as dev
from ..abstract_processor import AbstractProcessor, PerThread
from ..labels import Labels

logger = logging.getLogger(__name__)

MAX_OCTREE_LEVEL = 21     # CCCoreLib DgmOctree::MAX_OCTREE_LEVEL with 64-bit cell codes


def lcc_device(tile, d_sel, grid_size, min_component_size, d_labels=None, seed_label=0,
               threshold=0.0, level_from='all', want_comp=True, want_fill=False, bbox=None):
    """Run the device LCC on the points of ``tile`` selected by ``d_sel``.

    level_from : 'all' -> octree level from the extent of all tile points
                 (label_connected_comp.py:75 passes the unmasked ``points``);
                 'sel' -> from the extent of the selected points (what a nested
                 LabelConnectedComp on ``points[ids]`` computes, layer_lcc.py:86-91).
    Returns (d_comp int32[N] or None, d_fill uint8[N] or None, info dict).
    """
    lib = _lib.load()
    n = tile.n
    if bbox is None:
        bbox = tile.bbox(d_sel)
    bbox_all, bbox_sel, m = bbox
    d_comp = torch.full((n,), -1, dtype=torch.int32, device=tile.device) if want_comp else None
    d_fill = torch.zeros(n, dtype=torch.uint8, device=tile.device) if want_fill else None
    info = {'n_selected': m, 'n_cells': 0, 'n_components': 0, 'n_kept': 0, 'level': 0}
    if m == 0:
        return d_comp, d_fill, info
    ext = bbox_all if level_from == 'all' else bbox_sel
    ext_arr = (ctypes.c_double * 6)(*[float(v) for v in ext])
    level = lib.upcp_octree_level(ext_arr, tile.ncols, float(grid_size))
    info['level'] = level
    sel_arr = (ctypes.c_double * 6)(*[float(v) for v in bbox_sel])
    oct4 = (ctypes.c_float * 4)()
    _lib.check(lib.upcp_octree_params(sel_arr, MAX_OCTREE_LEVEL, oct4), 'upcp_octree_params')
    ws = dev.workspace(tile.device).get(lib.upcp_lcc_ws_bytes(n, m, level))
    d_info = torch.zeros(4, dtype=torch.int64, device=tile.device)
    do_fill = want_fill and d_labels is not None
    _lib.check(lib.upcp_lcc(dev.ptr(tile.x), dev.ptr(tile.y), dev.ptr(tile.z) if tile.ncols == 3 else None,
                            n, dev.ptr(d_sel), oct4, level, MAX_OCTREE_LEVEL, int(min_component_size),
                            dev.ptr(d_labels) if do_fill else None, int(seed_label), float(threshold),
                            dev.ptr(d_comp), dev.ptr(d_fill) if do_fill else None, dev.ptr(d_info),
                            dev.ptr(ws), ws.numel(), dev.stream_ptr()), 'upcp_lcc')
    info['_d_info'] = d_info
    return d_comp, d_fill, info


class LabelConnectedComp(AbstractProcessor):
    """
    Parameters
    ----------
    label : int (default: -1)
        Label to use when labelling the grown region.
    exclude_labels : list-like (default: [])
        Which labels to exclude (optional).
    grid_size : float (default: 0.1)
        Octree grid size for LCC method, in meters.
    min_component_size : int (default: 100)
        Minimum size of connected components to consider.
    threshold : float (default: 0.1)
        When labelling a cluster, at least this proportion of points should already be
        labelled.
    """

    octree_level = PerThread(lambda: None)      # level of the calling thread's last call

    def __init__(self, label=-1, exclude_labels=[], set_debug=False, grid_size=0.1,
                 min_component_size=100, threshold=0.1):
        super().__init__(label)
        self.grid_size = grid_size
        self.min_component_size = min_component_size
        self.threshold = threshold
        self.exclude_labels = exclude_labels
        self.debug = set_debug
        self.octree_level = None

    def _selection_device(self, tile, d_labels, d_mask):
        """Mask assembly of label_connected_comp.py:167-175."""
        if self.exclude_labels:
            return dev.mask_build(d_labels, None, base=True, ne=tuple(self.exclude_labels))
        if d_mask is None:
            return None
        # We need to un-mask all points of the desired class label.
        return dev.mask_build(d_labels, d_mask, eq=(self.label,))

    def _label_mask_device(self, tile, d_labels, d_mask, tilecode=None):
        log = logger.debug if self.debug else logger.info
        log('Clustering based Region Growing ' + f'(label={self.label}, {Labels.get_str(self.label)}).'
            if self.label in Labels.STR_DICT else f'Clustering based Region Growing (label={self.label}).')
        if self.label == -1:
            logger.warning('Label not set, defaulting to -1.')
        d_sel = self._selection_device(tile, d_labels, d_mask)
        _, d_fill, info = lcc_device(tile, d_sel, self.grid_size, self.min_component_size,
                                     d_labels=d_labels, seed_label=self.label, threshold=self.threshold,
                                     level_from='all', want_comp=False, want_fill=True)
        self.octree_level = info['level']
        return d_fill

    # -- reference API ---------------------------------------------------------------
    def get_label_mask(self, points, labels, mask=None, tilecode=None):
        """Label mask for the given pointcloud (bool[N]).  As in the reference
        (label_connected_comp.py:128-129) the accepted points are also written into the
        caller's ``labels`` array."""
        tile = dev.get_tile(points)
        d_labels = dev.upload_labels(labels, tile.device)
        d_mask = dev.upload_mask(mask, tile.n, tile.device)
        label_mask = dev.download_mask(self._label_mask_device(tile, d_labels, d_mask, tilecode))
        labels[label_mask] = self.label
        return label_mask

    def get_labels(self, points, labels, mask=None, tilecode=None):
        """Returns the labels for the given pointcloud (updated in place)."""
        label_mask = self.get_label_mask(points, labels, mask)
        labels[label_mask] = self.label
        return labels

    def get_components_device(self, tile, d_sel=None):
        """int32 cuda tensor [N]: component id per point, -1 for small components and for
        unselected points."""
        d_comp, _, info = lcc_device(tile, d_sel, self.grid_size, self.min_component_size,
                                     level_from='all', want_comp=True, want_fill=False)
        self.octree_level = info['level']
        return d_comp

    def get_components(self, points, labels=None):
        """Cluster id for each (non-excluded) point; clusters smaller than
        ``min_component_size`` are -1.  Ids are arbitrary (compare after relabelling); the
        returned array is float like the CloudCompare scalar field and, as in the reference,
        has one entry per MASKED point (label_connected_comp.py:207-231)."""
        points = np.asarray(points)
        tile = dev.get_tile(points)
        d_sel = None
        if labels is not None and self.exclude_labels:
            d_labels = dev.upload_labels(labels, tile.device)
            d_sel = dev.mask_build(d_labels, None, base=True, ne=tuple(self.exclude_labels))
        comp = self.get_components_device(tile, d_sel).cpu().numpy()
        if d_sel is not None:
            comp = comp[dev.download_mask(d_sel)]
        return comp.astype(np.float64)
