"""Ground refinement of AHNFuser (reference fusion/ahn_fuser.py:76-126): tentative ground points that lie
under unlabelled clusters are removed again.

On the device: the band test, the clustering of the unlabelled band points (LabelConnectedComp, grid 0.4 /
min 50, on their 3-D coordinates as ``lcc.get_components(points[label_ids])`` does) and the clipping of the
ground points against ALL cluster hulls at once.  On the host, per cluster: the Delaunay triangulation of a
few hundred to a few thousand xy pairs (scipy, as the reference) and the outer edges of its alpha complex
(``alpha_shape_utils``).  ``poly.buffer(0.05)`` is evaluated by the point-in-polygon kernel as a distance
(``upcp_pip_plan_create_buffered``; GEOS is in neither container -- unpinned)."""
import logging

import numpy as np
import torch

from .. import _lib
from .. import device as dev
from . import alpha_shape_utils

logger = logging.getLogger(__name__)


def refine_ground(fuser, tile, d_labels, d_mask, d_ground, raster, tilecode):
    """Returns a uint8 cuda mask [N] of tentative ground points to REMOVE (or None)."""
    from ..region_growing.label_connected_comp import lcc_device
    params = fuser.params
    logger.info('Refining ground surface...')
    # band: gz - bottom <= z <= gz + top among the masked points (ahn_fuser.py:115-116)
    d_band = raster.classify(tile, _lib.CLS_BAND_CC, a=params['bottom'], b=params['top'], d_sel=d_mask)
    # tmp_labels == UNKNOWN inside the band (ahn_fuser.py:162-166,119): label 0 and not tentative ground
    d_unl = dev.mask_build(d_labels, None, base=False, eq=(0,))
    d_unl = dev.mask_combine(d_unl, d_ground, 'andnot', out=d_unl)
    d_unl = dev.mask_combine(d_unl, d_band, 'and', out=d_unl)
    if dev.mask_count(d_unl) == 0:
        return None
    # the nested LabelConnectedComp sees only points[label_ids] (3 columns): level from that subset
    d_comp, _, _ = lcc_device(tile, d_unl, params['grid_size'], params['min_comp_size'], level_from='sel',
                              want_comp=True, want_fill=False)
    d_ids = torch.nonzero(d_comp >= 0).flatten()
    if d_ids.numel() == 0:
        return None
    comp = d_comp[d_ids].cpu().numpy()
    xy = torch.stack((tile.x[d_ids], tile.y[d_ids]), dim=1).cpu().numpy()
    order = np.argsort(comp, kind='stable')
    comp, xy = comp[order], xy[order]
    starts = np.r_[0, np.nonzero(np.diff(comp))[0] + 1, len(comp)]
    polys = []
    for a, b in zip(starts[:-1], starts[1:]):
        pts = xy[a:b]
        try:
            if params['use_concave']:
                hulls = alpha_shape_utils.alpha_shape(pts, alpha=params['concave_alpha'])
            else:
                hulls = [alpha_shape_utils.convex_hull_poly(pts)]
        except Exception as exc:                  # qhull on a degenerate cluster (the reference would raise here)
            logger.warning(f'Skipping a cluster of {len(pts)} points: {exc}')
            continue
        polys += [h.buffer(params['buffer']) for h in hulls]
    polys = [p for p in polys if len(p.exterior.coords) > 2]
    if not polys:
        return None
    d_band_ground = dev.mask_combine(d_ground, d_band, 'and')
    d_ref = dev.PipPlan(polys, device=tile.device).query(tile, d_band_ground)
    logger.info(f'{dev.mask_count(d_ref)} points removed.')
    return d_ref
