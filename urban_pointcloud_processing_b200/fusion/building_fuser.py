"""BGT building fuser (drop-in for ``upcp.fusion.BGTBuildingFuser``, reference
src/upcp/fusion/building_fuser.py:15-102): footprint point-in-polygon OR-ed over the
tile's building polygons, optionally capped by the AHN building surface."""
import logging

from .. import _lib
from .. import device as dev
from ..abstract_processor import AbstractProcessor
from ..labels import Labels
from ..utils.ahn_utils import device_raster_for

logger = logging.getLogger(__name__)


class BGTBuildingFuser(AbstractProcessor):
    """
    Parameters
    ----------
    label : int
        Class label to use for this fuser.
    bgt_reader : BGTPolyReader object
        Used to load building polygons.
    offset : int (default: 0)
        The footprint polygon will be extended by this amount (in meters).
    padding : float (default: 0)
        Optional padding (in m) around the tile when searching for objects.
    ahn_reader : AHNReader object
        Optional, if provided AHN data will be used to set a maximum height for each
        building polygon.
    ahn_eps : float (default: 0.2)
        Precision for the AHN elevation cut-off for buildings.
    """

    def __init__(self, label, bgt_reader, offset=0, padding=0, ahn_reader=None, ahn_eps=0.2):
        super().__init__(label)
        self.bgt_reader = bgt_reader
        self.offset = offset
        self.padding = padding
        self.ahn_reader = ahn_reader
        self.ahn_eps = ahn_eps

    def _label_mask_device(self, tile, d_labels, d_mask, tilecode):
        logger.info('BGT building fuser ' + f'(label={self.label}, {Labels.get_str(self.label)}).')
        building_polygons = self.bgt_reader.filter_tile(
            tilecode, bgt_types=['pand'], padding=self.padding, offset=self.offset, merge=True)
        if len(building_polygons) == 0:
            logger.debug('No buildings found for tile, skipping.')
            return None
        plan = dev.PipPlan(building_polygons, device=tile.device)
        d_building = plan.query(tile, d_mask)
        if self.ahn_reader is not None:
            # building_fuser.py:89-95: where the AHN building surface is finite the point must
            # satisfy z <= bld_z + ahn_eps
            raster = device_raster_for(self.ahn_reader, tilecode, 'building_surface')
            d_building = raster.classify(tile, _lib.CLS_CAP_LE, a=self.ahn_eps, d_sel=d_building)
        logger.debug(f'{len(building_polygons)} building polygons labelled.')
        return d_building
