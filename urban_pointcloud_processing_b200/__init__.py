"""urban_pointcloud_processing_b200 -- B200-native per-tile labelling hot path of
Amsterdam-AI-Team/Urban_PointCloud_Processing (``upcp``).

The module layout mirrors ``upcp`` so that switching is an import change:

    upcp.pipeline.Pipeline                      -> .pipeline.Pipeline
    upcp.fusion.{AHNFuser, BGTBuildingFuser,
                 BGTRoadFuser, NoiseFilter}     -> .fusion.*
    upcp.region_growing.{LabelConnectedComp,
                 LayerLCC}                      -> .region_growing.*
    upcp.region_growing.region_growing.RegionGrowing -> .region_growing.region_growing.RegionGrowing
    upcp.utils.{ahn_utils, bgt_utils, clip_utils,
                interpolation, math_utils, las_utils} -> .utils.*
    upcp.labels.Labels                          -> .labels.Labels

All per-point arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI in
``include/upcp_cuda.h`` (``libupcp_cuda.so``, loaded with ctypes by ``_lib``).  There is
no CPU fallback: without the built library or without a GPU every compute call raises.
"""
__version__ = '0.1.0'

from .labels import Labels  # noqa: F401
