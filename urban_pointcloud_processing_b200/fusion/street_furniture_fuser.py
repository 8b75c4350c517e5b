"""Street-furniture fuser (drop-in for ``upcp.fusion.BGTStreetFurnitureFuser``, reference
src/upcp/fusion/street_furniture_fuser.py:17-137): clusters of the right height and footprint
whose bounding-rectangle centre lies within ``max_dist`` of a BGT point object of ``bgt_type``.

A client of LabelConnectedComp.get_components; same split as CarFuser: the
device clusters and reduces, the host judges the few height-qualified clusters, the device marks
the accepted clusters' points."""
import logging

import numpy as np

from ..abstract_processor import AbstractProcessor
from ..labels import Labels
from ..region_growing.label_connected_comp import lcc_device
from ..utils import object_utils
from ..utils.ahn_utils import device_raster_for
from ..utils.math_utils import minimum_bounding_rectangle

logger = logging.getLogger(__name__)


class BGTStreetFurnitureFuser(AbstractProcessor):
    """
    Parameters
    ----------
    label : int
        Class label to use for this fuser.
    bgt_type : str
        Specify the 'type' of point object: 'boom', 'lichtmast', 'verkeersbord'
    bgt_reader : BGTPointReader object
        Used to load street furniture points.
    ahn_reader : AHNReader object
        Elevation data reader.
    """

    def __init__(self, label, bgt_type, bgt_reader, ahn_reader, grid_size=0.05, min_component_size=1500,
                 padding=0, max_dist=1., params={}):
        super().__init__(label)
        self.bgt_type = bgt_type
        self.bgt_reader = bgt_reader
        self.ahn_reader = ahn_reader
        self.grid_size = grid_size
        self.min_component_size = min_component_size
        self.padding = padding
        self.max_dist = max_dist
        self.params = params
        self.last_accepted = 0

    def _furniture_like(self, stats, bgt_points, min_height, max_height, min_width, max_width,
                        min_length, max_length):
        """bool[K]: clusters accepted as objects of ``bgt_type`` (street_furniture_fuser.py:62-83)."""
        accepted = np.zeros(stats.n_components, dtype=bool)
        ids, _, _ = object_utils.height_candidates(stats, min_height, max_height)
        for k in ids:
            _, _, mbr_width, mbr_length, center_point = minimum_bounding_rectangle(stats.xy_of(k))
            if min_width < mbr_width < max_width and min_length < mbr_length < max_length:
                for bgt_point in bgt_points:
                    if np.linalg.norm(bgt_point - center_point) <= self.max_dist:
                        accepted[k] = True
                        break
        return accepted

    def _label_mask_device(self, tile, d_labels, d_mask, tilecode):
        logger.info('Street furniture fuser ' + f'(label={self.label}, {Labels.get_str(self.label)}).')
        import torch
        bgt_points = self.bgt_reader.filter_tile(tilecode, bgt_types=[self.bgt_type], padding=self.padding,
                                                 return_types=False)
        if len(bgt_points) == 0:
            logger.debug(f'No {self.bgt_type} objects found in tile, ' + ' skipping.')
            return None
        if d_mask is None:
            d_mask = torch.ones(tile.n, dtype=torch.uint8, device=tile.device)
        d_gz = device_raster_for(self.ahn_reader, tilecode, 'ground_surface').lookup(tile, d_sel=d_mask)
        d_comp, _, _ = lcc_device(tile, d_mask, self.grid_size, self.min_component_size, level_from='sel',
                                  want_comp=True, want_fill=False)
        stats = object_utils.ComponentStats(tile, d_comp, d_gz)
        accepted = self._furniture_like(stats, bgt_points, **self.params)
        self.last_accepted = int(accepted.sum())
        logger.debug(f'{self.last_accepted} {self.bgt_type} objects labelled.')
        return stats.mark(accepted)
