from .ahn_fuser import AHNFuser
from .building_fuser import BGTBuildingFuser
from .road_fuser import BGTRoadFuser
from .noise_filter import NoiseFilter

__all__ = ['AHNFuser', 'BGTBuildingFuser', 'BGTRoadFuser', 'NoiseFilter']
