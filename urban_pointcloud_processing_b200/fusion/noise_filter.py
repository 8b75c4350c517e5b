"""Noise filter (drop-in for ``upcp.fusion.NoiseFilter``, reference
src/upcp/fusion/noise_filter.py:15-84): points in small connected components and points
below the AHN ground surface are labelled as noise.  A thin client of
LabelConnectedComp.get_components + the fused raster compare."""
import logging

from .. import _lib
from .. import device as dev
from ..abstract_processor import AbstractProcessor
from ..labels import Labels
from ..region_growing.label_connected_comp import lcc_device
from ..utils.ahn_utils import device_raster_for

logger = logging.getLogger(__name__)


class NoiseFilter(AbstractProcessor):
    """
    Parameters
    ----------
    label : int
        Class label to use for this fuser.
    ahn_reader : AHNReader object
        Elevation data reader.
    epsilon : float (default: 0.2)
        Precision of the fuser.
    grid_size : float (default: 0.1)
        Octree grid size for LCC method, in meters.
    min_component_size : int (default: 100)
        Minimum size of a cluster below which points are classified as noise.
    """

    def __init__(self, label, ahn_reader, epsilon=0.2, grid_size=0.1, min_component_size=100):
        super().__init__(label)
        self.ahn_reader = ahn_reader
        self.epsilon = epsilon
        self.grid_size = grid_size
        self.min_component_size = min_component_size

    def _label_mask_device(self, tile, d_labels, d_mask, tilecode):
        logger.info('Noise filter ' + f'(label={self.label}, {Labels.get_str(self.label)}).')
        import torch
        if d_mask is None:
            d_mask = torch.ones(tile.n, dtype=torch.uint8, device=tile.device)
        # components of points[mask]: the nested LabelConnectedComp sees only the masked
        # cloud, so the octree level comes from ITS extent (noise_filter.py:64-66)
        d_comp, _, _ = lcc_device(tile, d_mask, self.grid_size, self.min_component_size,
                                  level_from='sel', want_comp=True, want_fill=False)
        d_small = ((d_comp < 0).to(torch.uint8))
        d_small = dev.mask_combine(d_small, d_mask, 'and', out=d_small)
        raster = device_raster_for(self.ahn_reader, tilecode, 'ground_surface')
        d_below = raster.classify(tile, _lib.CLS_BELOW_NEG, a=self.epsilon, d_sel=d_mask)
        return dev.mask_combine(d_small, d_below, 'or', out=d_small)
