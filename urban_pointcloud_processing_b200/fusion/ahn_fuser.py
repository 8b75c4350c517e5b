"""AHN data fuser (drop-in for ``upcp.fusion.AHNFuser``, reference
src/upcp/fusion/ahn_fuser.py:18-173).

Ground / building labelling by comparing z with the AHN raster value of the cell the
point falls in.  The per-point work (raster lookup + threshold) is ONE fused kernel
(``upcp_raster_classify``).  The optional ground refinement (``refine_ground``,
ahn_fuser.py:76-126) needs per-component alpha shapes (scipy Delaunay) and polygon
buffering (shapely/GEOS); it runs its clustering (LabelConnectedComp) and
point-in-polygon on the GPU and keeps the tiny per-component hull construction on the
host, and is only available when shapely is installed."""
import logging

import numpy as np

from .. import _lib
from .. import device as dev
from ..abstract_processor import AbstractProcessor
from ..labels import Labels
from ..utils.ahn_utils import device_raster_for

logger = logging.getLogger(__name__)


class AHNFuser(AbstractProcessor):
    """
    Parameters
    ----------
    label : int
        Class label to use for this fuser.
    ahn_reader : AHNReader object
        Used to read the AHN data.
    target : str (default: 'ground')
        Label either 'ground' or 'building' points.
    epsilon : float (default: 0.2)
        Precision of the fuser.
    refine_ground : bool (default: True)
        Remove tentative ground points that lie under unlabelled clusters.
    params : dict
        Refinement parameters (bottom, top, grid_size, min_comp_size, buffer,
        use_concave, concave_alpha).
    """

    TARGETS = ('ground', 'building')

    def __init__(self, label, ahn_reader, target='ground', epsilon=0.2,
                 refine_ground=True, params={}):
        super().__init__(label)
        if target not in self.TARGETS:
            logger.error(f'Target should be one of {self.TARGETS}.')
            return None
        if getattr(ahn_reader, 'NAME', None) == 'geotiff' and target == 'building':
            logger.error(f'The {ahn_reader.NAME} reader cannot supply {target} data.')
            return None
        self.ahn_reader = ahn_reader
        self.method = getattr(ahn_reader, 'NAME', 'npz')
        self.target = target
        self.epsilon = epsilon
        self.refine_ground = refine_ground
        self.params = self._set_defaults(params)

    def _set_defaults(self, params):
        """Set defaults for parameters if not provided."""
        defaults = {'bottom': 0.02, 'top': 0.5, 'grid_size': 0.4, 'min_comp_size': 50,
                    'buffer': 0.05, 'use_concave': True, 'concave_alpha': 2.0}
        for key, value in defaults.items():
            if key not in params:
                params[key] = value
        return params

    def _refine_ground_device(self, tile, d_labels, d_mask, d_ground, raster, tilecode):
        """ahn_fuser.py:111-126 / :76-109 with the clustering and the clipping on the GPU."""
        from ..utils import refine_utils
        return refine_utils.refine_ground(self, tile, d_labels, d_mask, d_ground, raster, tilecode)

    def _label_mask_device(self, tile, d_labels, d_mask, tilecode):
        logger.info(f'AHN [{self.method}/{self.target}] fuser ' +
                    f'(label={self.label}, {Labels.get_str(self.label)}).')
        if self.target == 'ground':
            raster = device_raster_for(self.ahn_reader, tilecode, 'ground_surface')
            d_ground = raster.classify(tile, _lib.CLS_ABS_LT, a=self.epsilon, d_sel=d_mask)
            if self.refine_ground:
                n_ground = dev.mask_count(d_ground)
                if n_ground > 0:
                    logger.info(f'{n_ground} points added.')
                    d_ref = self._refine_ground_device(tile, d_labels, d_mask, d_ground, raster, tilecode)
                    if d_ref is not None:
                        d_ground = dev.mask_combine(d_ground, d_ref, 'andnot', out=d_ground)
            return d_ground
        raster = device_raster_for(self.ahn_reader, tilecode, 'building_surface')
        return raster.classify(tile, _lib.CLS_LT, a=self.epsilon, d_sel=d_mask)
