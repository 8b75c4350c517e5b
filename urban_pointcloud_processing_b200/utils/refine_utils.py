"""Ground refinement of AHNFuser (reference fusion/ahn_fuser.py:76-126): tentative ground
points that lie under unlabelled clusters are removed again.

The clustering (LabelConnectedComp, grid 0.4 / min 50) and the clipping of ground points
against the per-cluster hull polygons run on the GPU; the hull construction itself
(``alpha_shape`` = scipy Delaunay + shapely polygonize, then ``poly.buffer(0.05)`` = GEOS)
is tiny per-component host work that needs shapely.  shapely/GEOS is not installed in this
image, and polygon offsetting cannot be restated bit-exactly without it, so without shapely
this raises instead of silently computing something else.  ("next" row 3.)"""
import logging

import numpy as np
import torch

from .. import _lib
from .. import device as dev
from . import bgt_utils

logger = logging.getLogger(__name__)


def refine_ground(fuser, lcc_cls, tile, d_labels, d_mask, d_ground, raster, tilecode):
    """Returns a uint8 cuda mask [N] of ground points to REMOVE (or None)."""
    if not bgt_utils.HAVE_SHAPELY:
        raise NotImplementedError(
            'AHNFuser(refine_ground=True) builds per-cluster alpha shapes with shapely/GEOS '
            '(ahn_fuser.py:92-108), which is not installed; construct the fuser with '
            'refine_ground=False')
    from scipy.spatial import Delaunay            # noqa: F401  (host-side hull construction)
    from shapely.geometry import MultiPoint
    params = fuser.params
    logger.info('Refining ground surface...')
    # band: gz - bottom <= z <= gz + top among the masked points (ahn_fuser.py:115-116)
    d_band = raster.classify(tile, _lib.CLS_BAND_CC, a=params['bottom'], b=params['top'], d_sel=d_mask)
    # unlabelled points of the band: labels (with the tentative ground applied) == UNKNOWN
    d_unl = dev.mask_build(d_labels, None, base=False, eq=(0,))
    d_unl = dev.mask_combine(d_unl, d_ground, 'andnot', out=d_unl)
    d_unl = dev.mask_combine(d_unl, d_band, 'and', out=d_unl)
    if dev.mask_count(d_unl) == 0:
        return None
    lcc = lcc_cls(0, grid_size=params['grid_size'], min_component_size=params['min_comp_size'])
    # the nested LabelConnectedComp sees only points[label_ids]: level from that subset
    from ..region_growing.label_connected_comp import lcc_device
    d_comp, _, _ = lcc_device(tile, d_unl, params['grid_size'], params['min_comp_size'], level_from='sel',
                              want_comp=True, want_fill=False)
    comp = d_comp.cpu().numpy()
    ids = np.where(comp >= 0)[0]
    if len(ids) == 0:
        return None
    xy = np.column_stack((tile.x[torch.as_tensor(ids, device=tile.device)].cpu().numpy(),
                          tile.y[torch.as_tensor(ids, device=tile.device)].cpu().numpy()))
    polys = []
    for cc in np.unique(comp[ids]):
        pts = xy[comp[ids] == cc]
        if params['use_concave']:
            from .alpha_shape_host import alpha_shape
            hulls = alpha_shape(pts, alpha=params['concave_alpha'])
        else:
            hulls = [MultiPoint(pts).convex_hull]
        polys += [h.buffer(params['buffer']) for h in hulls]
    polys = [p for p in polys if hasattr(p, 'exterior') and len(p.exterior.coords) > 2]
    if not polys:
        return None
    d_band_ground = dev.mask_combine(d_ground, d_band, 'and')
    return dev.PipPlan(polys, device=tile.device).query(tile, d_band_ground)
