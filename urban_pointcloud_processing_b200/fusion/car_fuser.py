"""Car fuser (drop-in for ``upcp.fusion.CarFuser``, reference src/upcp/fusion/car_fuser.py:18-137):
clusters of car-like height and footprint that overlap a road part are labelled, together with
everything inside the cluster's minimum bounding rectangle between the ground and the car's top.

A client of LabelConnectedComp.get_components.  Device: ground lookup,
clustering, per-cluster reductions (``upcp_component_stats``), the final 3-D rectangle clip of the
whole masked cloud (``upcp_prism_clip``).  Host: the reference's per-object geometry on the few
clusters whose height qualifies (qhull convex hull + minimum bounding rectangle, road overlap)."""
import logging

import numpy as np

from ..abstract_processor import AbstractProcessor, PerThread
from ..labels import Labels
from ..region_growing.label_connected_comp import lcc_device
from ..utils import object_utils
from ..utils.ahn_utils import device_raster_for
from ..utils.math_utils import minimum_bounding_rectangle

logger = logging.getLogger(__name__)


class CarFuser(AbstractProcessor):
    """
    Parameters
    ----------
    label : int
        Class label to use for this fuser.
    ahn_reader : AHNReader object
        Elevation data reader.
    bgt_reader : BGTPolyReader object
        Used to load road part polygons.
    """

    last_objects = PerThread(list)      # accepted (ring, bottom, top) boxes of the calling thread's last tile

    def __init__(self, label, ahn_reader, bgt_reader, grid_size=0.05, min_component_size=5000,
                 overlap_perc=20, params={}, exclude_types=["fietspad", "voetpad"]):
        super().__init__(label)
        self.ahn_reader = ahn_reader
        self.bgt_reader = bgt_reader
        self.grid_size = grid_size
        self.min_component_size = min_component_size
        self.overlap_perc = overlap_perc
        self.params = params
        self.exclude_types = exclude_types
        self.last_objects = []

    def _car_like_objects(self, stats, road_polygons, min_height, max_height, min_width, max_width,
                          min_length, max_length):
        """[(closed rectangle ring (5, 2), bottom, top)] of the clusters accepted as cars
        (car_fuser.py:57-84)."""
        objects = []
        ids, cc_z, max_z = object_utils.height_candidates(stats, min_height, max_height)
        for k in ids:
            mbrect, _, mbr_width, mbr_length, _ = minimum_bounding_rectangle(stats.xy_of(k))
            if not (min_width < mbr_width < max_width and min_length < mbr_length < max_length):
                continue
            ring = np.vstack((mbrect, mbrect[0]))
            for road in road_polygons:
                if object_utils.overlap_percentage(ring, road) > self.overlap_perc:
                    objects.append((ring, float(cc_z[k]), float(max_z[k])))
                    break
        return objects

    def _label_mask_device(self, tile, d_labels, d_mask, tilecode):
        logger.info('Car fuser ' + f'(label={self.label}, {Labels.get_str(self.label)}).')
        import torch
        road_polygons = self.bgt_reader.filter_tile(tilecode, exclude_types=self.exclude_types, merge=False)
        if len(road_polygons) == 0:
            logger.debug('No road parts found for tile, skipping.')
            return None
        if d_mask is None:
            d_mask = torch.ones(tile.n, dtype=torch.uint8, device=tile.device)
        # interpolated ground elevation of the masked points (NaN elsewhere / where AHN is empty)
        d_gz = device_raster_for(self.ahn_reader, tilecode, 'ground_surface').lookup(tile, d_sel=d_mask)
        # clusters of points[mask]: the nested LabelConnectedComp only sees the masked cloud
        d_comp, _, _ = lcc_device(tile, d_mask, self.grid_size, self.min_component_size, level_from='sel',
                                  want_comp=True, want_fill=False)
        stats = object_utils.ComponentStats(tile, d_comp, d_gz)
        objects = self._car_like_objects(stats, road_polygons, **self.params)
        self.last_objects = objects
        logger.debug(f'{len(objects)} cars labelled.')
        return object_utils.prism_clip_device(tile, d_mask, [o[0] for o in objects], [(o[1], o[2]) for o in objects])
