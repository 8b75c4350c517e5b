"""oracle -- CPU restatement of the reference's per-tile labelling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``urban_pointcloud_processing_b200`` imports
this package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs do, and there only as the checker / the reported CPU baseline.

Layout
  native/upcp_oracle.c   plain-C restatement of the third-party natives the reference
                         calls (numba PIP, cccorelib connected components, open3d
                         kNN / normals / covariance, the region-growing queue loop)
  natives.py             ctypes wrappers around the compiled C restatement
  reference_py.py        numpy restatement of the reference's Python-level control flow
                         (FastGridInterpolator, the fusers, LabelConnectedComp, LayerLCC,
                         RegionGrowing, Pipeline.process_cloud), citing file:line
  ref_harness.py         imports the REAL reference from /root/reference with stand-in
                         modules for the missing natives (only usable in the build
                         container; used by make_golden.py to pin the restatement)
  make_golden.py         generates tests/golden/*.npz from the real reference code

Pinning status:
  PINNED against the real reference code run in the build container (golden vectors):
      point-in-polygon (numba clip_utils), raster lookup (FastGridInterpolator),
      get_octree_level, vector_angle, the fuser / LCC / LayerLCC / RegionGrowing /
      Pipeline CONTROL FLOW (reference Python run over the restated natives).
  PARITY UNPINNED (third-party source absent: cccorelib/pycc git master, open3d 0.16):
      octree cell assignment + connected components, kNN ordering, hybrid normals,
      covariance.  The restatement follows the published algorithms; no golden vector of
      the real libraries exists in the reference tree.
"""
