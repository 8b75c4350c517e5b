from .ahn_fuser import AHNFuser
from .building_fuser import BGTBuildingFuser
from .road_fuser import BGTRoadFuser
from .noise_filter import NoiseFilter
from .car_fuser import CarFuser
from .street_furniture_fuser import BGTStreetFurnitureFuser

__all__ = ['AHNFuser', 'BGTBuildingFuser', 'BGTRoadFuser', 'NoiseFilter', 'CarFuser', 'BGTStreetFurnitureFuser']
