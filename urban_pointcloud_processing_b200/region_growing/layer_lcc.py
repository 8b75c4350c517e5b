"""Layered clustering-based region growing (drop-in for
``upcp.region_growing.LayerLCC``, reference src/upcp/region_growing/layer_lcc.py:13-145).
Per height-above-ground layer: fused AHN raster lookup + band test selects the layer,
then a fresh connected-component labelling (octree level from the LAYER's extent) with
seed-fraction fill; labels are fed forward to the next layer."""
import logging

import numpy as np
import torch

from .. import _lib
from .. import device as dev
from ..abstract_processor import AbstractProcessor
from ..labels import Labels
from ..utils.ahn_utils import device_raster_for
from .label_connected_comp import lcc_device

logger = logging.getLogger(__name__)


class LayerLCC(AbstractProcessor):
    """
    Parameters
    ----------
    label : int
        Class label to use.
    ahn_reader : AHNReader object
        Used to get the correct elevation data.
    reset_noise : bool (default: False)
        If set to true, points previously labelled as noise will be considered.
    params : dict or list of dicts
        Per layer: 'bottom' (-inf), 'top' (inf), 'grid_size' (0.1), 'min_comp_size' (100),
        'threshold' (0.5).
    """

    def __init__(self, label, ahn_reader, reset_noise=False, params=[]):
        super().__init__(label)
        self.ahn_reader = ahn_reader
        self.reset_noise = reset_noise
        if isinstance(params, dict):
            params = [params]
        self.params = self._set_defaults(params)

    def _set_defaults(self, params):
        """Set defaults for parameters if not provided."""
        defaults = {'bottom': -np.inf, 'top': np.inf, 'grid_size': 0.1, 'min_comp_size': 100,
                    'threshold': 0.5}
        for layer in params:
            for key, value in defaults.items():
                if key not in layer:
                    layer[key] = value
        return params

    def _filter_layer_device(self, tile, d_mask_copy, d_labels_copy, raster, params):
        """layer_lcc.py:70-93 on the device; returns a uint8 mask [N] or None."""
        # height_mask: gz + bottom < z <= gz + top, among the un-masked points
        d_layer = raster.classify(tile, _lib.CLS_BAND_OC, a=params['bottom'], b=params['top'],
                                  d_sel=d_mask_copy)
        bbox = tile.bbox(d_layer)
        if bbox[2] == 0:
            logger.debug(f"Empty layer: {params['bottom']} -> {params['top']}.")
            return None
        d_seed = dev.mask_build(d_labels_copy, None, base=False, eq=(self.label,))
        n_valid = dev.mask_count(dev.mask_combine(d_seed, d_layer, 'and', out=d_seed))
        if n_valid == 0:
            logger.debug(f"No marked points in layer: {params['bottom']} -> {params['top']}.")
            return None
        _, d_fill, _ = lcc_device(tile, d_layer, params['grid_size'], params['min_comp_size'],
                                  d_labels=d_labels_copy, seed_label=self.label,
                                  threshold=params['threshold'], level_from='sel',
                                  want_comp=False, want_fill=True, bbox=bbox)
        return d_fill

    def _label_mask_device(self, tile, d_labels, d_mask, tilecode):
        logger.info('LayerLCC ' + f'(label={self.label}, {Labels.get_str(self.label)}))')
        if d_mask is None:
            d_mask = torch.ones(tile.n, dtype=torch.uint8, device=tile.device)
        # We need to un-mask all points of the desired class label.
        eq = (self.label, Labels.NOISE) if self.reset_noise else (self.label,)
        d_mask_copy = dev.mask_build(d_labels, d_mask, eq=eq)
        raster = device_raster_for(self.ahn_reader, tilecode, 'ground_surface')
        d_layer_mask = torch.zeros(tile.n, dtype=torch.uint8, device=tile.device)
        d_labels_copy = d_labels.clone()
        for i, layer in enumerate(self.params):
            logger.debug(f'Layer {i}: {layer}')
            d_fill = self._filter_layer_device(tile, d_mask_copy, d_labels_copy, raster, layer)
            if d_fill is not None:
                dev.mask_combine(d_layer_mask, d_fill, 'or', out=d_layer_mask)
            dev.labels_assign(d_labels_copy, d_layer_mask, self.label)
        if self.reset_noise:
            d_allowed = dev.mask_build(d_labels, d_mask, eq=(Labels.NOISE,))
        else:
            d_allowed = d_mask
        return dev.mask_combine(d_layer_mask, d_allowed, 'and', out=d_layer_mask)
