"""BGT road fuser (drop-in for ``upcp.fusion.BGTRoadFuser``, reference
src/upcp/fusion/road_fuser.py:15-95): re-labels GROUND points that fall inside BGT road
part polygons.  As in the reference the incoming mask is ignored: the domain is
``labels == Labels.GROUND`` (road_fuser.py:81)."""
import logging

from .. import device as dev
from ..abstract_processor import AbstractProcessor
from ..labels import Labels

logger = logging.getLogger(__name__)


class BGTRoadFuser(AbstractProcessor):
    """
    Parameters
    ----------
    label : int
        Class label to use for this fuser.
    bgt_reader : BGTPolyReader object
        Used to load road part polygons.
    bgt_types : str or list of [str] (default: all available)
        Which road part types (BGT layers) to label as road.
    offset : int (default: 0)
        The road polygon will be extended by this amount (in meters).
    padding : float (default: 0)
        Optional padding (in m) around the tile when searching for objects.
    """

    ALL_TYPES = ['rijbaan_lokale_weg', 'parkeervlak', 'rijbaan_autoweg',
                 'rijbaan_autosnelweg', 'rijbaan_regionale_weg', 'ov-baan', 'fietspad']

    def __init__(self, label, bgt_reader, bgt_types=ALL_TYPES, offset=0, padding=0):
        super().__init__(label)
        self.bgt_reader = bgt_reader
        if isinstance(bgt_types, str):
            bgt_types = [bgt_types]
        self.bgt_types = bgt_types
        self.offset = offset
        self.padding = padding

    def _label_mask_device(self, tile, d_labels, d_mask, tilecode):
        logger.info('BGT road fuser ' + f'(label={self.label}, {Labels.get_str(self.label)}).')
        road_polygons = self.bgt_reader.filter_tile(
            tilecode, bgt_types=self.bgt_types, padding=self.padding, offset=self.offset, merge=True)
        if len(road_polygons) == 0:
            logger.debug('No road parts found in tile, skipping.')
            return None
        # Already labelled ground points can be labelled as road.
        d_ground = dev.mask_build(d_labels, None, base=False, eq=(Labels.GROUND,))
        plan = dev.PipPlan(road_polygons, device=tile.device)
        d_road = plan.query(tile, d_ground)
        logger.debug(f'{len(road_polygons)} road polygons labelled.')
        return d_road
