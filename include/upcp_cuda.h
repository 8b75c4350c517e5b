/*
 * upcp_cuda.h -- C ABI of the B200-native per-tile labelling hot path.
 *
 * This is the boundary a maintainer of Amsterdam-AI-Team/Urban_PointCloud_Processing
 * (`upcp`) binds to replace the third-party natives the reference calls on its hot
 * path (numba-JIT point-in-polygon, numpy digitize, cccorelib/pycc connected
 * components, open3d KD-tree / normals).  Each entry point cites the reference
 * interface it replaces as `file:line` relative to the reference checkout.
 *
 * Conventions
 *   - plain C, no torch / C++ types in any signature;
 *   - every pointer named d_* is a DEVICE pointer (cudaMalloc / torch CUDA tensor
 *     data_ptr), every pointer named h_* is a HOST pointer;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - scratch memory is CALLER-owned: `*_ws_bytes()` returns the size, the caller
 *     passes a device buffer of at least that size (256-byte aligned);
 *   - all functions return 0 (UPCP_OK) or a negative UPCP_E_* code, never throw;
 *     `upcp_last_error()` returns a thread-local message for the last failure;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns UPCP_E_CUDA.
 *   - point coordinates are fp64 structure-of-arrays (x[n], y[n], z[n]); labels are
 *     int32; masks are one byte per point (0 / 1).
 */
#ifndef UPCP_CUDA_H
#define UPCP_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UPCP_OK            0
#define UPCP_E_INVALID    -1   /* bad argument                                  */
#define UPCP_E_CUDA       -2   /* CUDA runtime error / no device                */
#define UPCP_E_WORKSPACE  -3   /* caller workspace too small                    */
#define UPCP_E_CAPACITY   -4   /* an internal capacity (k, rings, ...) exceeded */
#define UPCP_E_INTERNAL   -5   /* device-side invariant violated (flag raised)  */

#define UPCP_ABI_VERSION   1

/* ------------------------------------------------------------------ misc -- */
int         upcp_abi_version(void);
const char* upcp_last_error(void);
/* Number of visible CUDA devices (0 when there is none); never fails. */
int         upcp_device_count(void);
/* Total number of kernel launches issued by this library in this process
 * (used by bench.py for the `gpu_launches` figure). */
int64_t     upcp_launch_count(void);

/* Per-kernel device timing for bench.py: when enabled, the dominant kernels are bracketed
 * with cudaEvents on their launching stream.  upcp_prof_read() synchronises the device,
 * adds up the elapsed time of all completed brackets of `slot` since the last reset and
 * returns the number of launches and the units processed (points; for the sort: algorithmic
 * bytes = elements x passes x (3 key bytes + 8)).  Slots: 0 raster, 1 pip, 2 radix sort (all passes
 * of one sort), 3 cc link (union-find), 4 neighbour search + normals (the WHOLE stage S4),
 * 5 region grow (persistent kernel). */
#define UPCP_PROF_RASTER 0
#define UPCP_PROF_PIP    1
#define UPCP_PROF_SORT   2
#define UPCP_PROF_LINK   3
#define UPCP_PROF_KNN    4
#define UPCP_PROF_GROW   5
/* sub-brackets of slot 4 (which spans the whole neighbour-search stage S4): 6 grid build (cell keys,
 * prefix, placement of the sorted records), 7 search (both block tiers, compactions, second grid,
 * per-query tier), 8 epilogue (exact order, covariances, normals, kNN rows) */
#define UPCP_PROF_KNN_BUILD    6
#define UPCP_PROF_KNN_SEARCH   7
#define UPCP_PROF_KNN_EPILOGUE 8
#define UPCP_PROF_SLOTS  9
int         upcp_prof_enable(int on);
int         upcp_prof_reset(void);
int         upcp_prof_read(int slot, double* total_ms, int64_t* launches, int64_t* units);
const char* upcp_prof_slot_name(int slot);

/* Tuning knobs of the kernels, process-wide, explicit (no environment variables are read).
 * upcp_tuning_default() fills in what bench.py was measured with; upcp_tuning_set() validates
 * and installs a copy that every later call reads once at entry. */
typedef struct upcp_tuning {
    int32_t knn_tiers;            /* block tiers before the per-query tier: 1 or 2                       */
    int32_t knn_tier1_max_block;  /* 3^3 blocks with more candidates leave tier 1 to the next tier       */
    int32_t knn_tier2_max_block;  /* the same for the 5^3 blocks of tier 2                               */
    int32_t knn_coarse;           /* cell factor of the per-query tier's second grid (1 = none)          */
    int32_t grow_local_chain;     /* region growth: warp-local continuation cap (0..32)                  */
    int32_t grow_ctas_per_sm;     /* region growth: resident CTAs per SM of the persistent kernel (>= 1) */
    int32_t knn_guess_tenths;     /* block tiers: first histogram range = this / 10 x the k-th squared distance a
                                     planar cloud of the block's point count would have (30 = 3.0)       */
    int32_t knn_guess_max_pct;    /* ... tried only when below this percentage of the certifiable range  */
    int32_t reserved[8];
} upcp_tuning;
void upcp_tuning_default(upcp_tuning* out);
int  upcp_tuning_get(upcp_tuning* out);
int  upcp_tuning_set(const upcp_tuning* in);

/* ------------------------------------------------------- tile / layout ---- */
/* Replaces: the implicit numpy layout of `points` (pipeline.py:124 delivers an
 * F-ordered (N,3) float64 array; callers may also pass C-ordered).  Converts a
 * C-ordered AoS (n, ncols) fp64 array into SoA columns.  ncols is 2 or 3; with
 * ncols == 2, d_z is filled with 0.0 (label_connected_comp.py:61-62). */
int upcp_points_aos_to_soa(const double* d_xyz, int64_t n, int ncols,
                           double* d_x, double* d_y, double* d_z, void* stream);

/* LAS ingest: coordinates from the raw integer fields of the point records,
 *   x = X * scale + offset  (fp64, product and sum rounded separately)
 * which is what laspy's scaled `las.x / las.y / las.z` deliver to pipeline.py:123-124
 * (las_utils.read_las, las_utils.py:118-120) -- computed on the device from 12 B per point
 * instead of uploading 24 B per point of host-side doubles. */
int upcp_points_from_scaled_int32(const int32_t* d_X, const int32_t* d_Y, const int32_t* d_Z, int64_t n,
                                  const double* h_scale3, const double* h_offset3,
                                  double* d_x, double* d_y, double* d_z, void* stream);

/* Bounding boxes in fp64.  d_out receives 13 doubles:
 *   [0..5]  min x,y,z / max x,y,z over ALL n points
 *   [6..11] min x,y,z / max x,y,z over points with d_sel[i] != 0 (all if d_sel NULL)
 *   [12]    number of selected points (as double, exact below 2^53)
 * Replaces: np.min/np.max in math_utils.py:21-33 (get_octree_level) and the
 * cloud bounding box CCCoreLib computes in DgmOctree::updateMinAndMaxTables. */
size_t upcp_bbox_ws_bytes(int64_t n);
int upcp_bbox(const double* d_x, const double* d_y, const double* d_z, int64_t n,
              const uint8_t* d_sel, double* d_out13,
              void* d_ws, size_t ws_bytes, void* stream);

/* Widen / narrow label arrays between the host dtype and the device int32. */
int upcp_labels_widen_u8(const uint8_t* d_in, int64_t n, int32_t* d_out, void* stream);
int upcp_labels_widen_u16(const uint16_t* d_in, int64_t n, int32_t* d_out, void* stream);
int upcp_labels_narrow_u8(const int32_t* d_in, int64_t n, uint8_t* d_out, void* stream);
int upcp_labels_narrow_u16(const int32_t* d_in, int64_t n, uint16_t* d_out, void* stream);

/* out[i] = ((d_mask_in ? d_mask_in[i] != 0 : base) || labels[i] in eq[0..n_eq))
 *          && labels[i] not in ne[0..n_ne)
 * Replaces the numpy mask assembly of pipeline.py:90 (`labels == 0`),
 * road_fuser.py:81, label_connected_comp.py:47-51,167-175, layer_lcc.py:118-122,
 * region_growing.py:34-49.  n_eq, n_ne <= 8. */
int upcp_mask_build(const int32_t* d_labels, int64_t n, const uint8_t* d_mask_in,
                    int base, const int32_t* h_eq, int n_eq,
                    const int32_t* h_ne, int n_ne, uint8_t* d_out, void* stream);

/* op: 0 = a & b, 1 = a | b, 2 = a & ~b.  d_out may alias d_a. */
int upcp_mask_combine(const uint8_t* d_a, const uint8_t* d_b, int64_t n, int op,
                      uint8_t* d_out, void* stream);

/* labels[i] = value where d_mask[i] != 0  (e.g. ahn_fuser.py:172). */
int upcp_labels_assign(int32_t* d_labels, int64_t n, const uint8_t* d_mask,
                       int32_t value, void* stream);

/* *d_count (int64) = number of non-zero mask bytes. */
int upcp_mask_count(const uint8_t* d_mask, int64_t n, int64_t* d_count, void* stream);

/* Label histogram: d_hist[256] (int64) += count of each label value in [0,255];
 * values outside go to bin 255.  Caller zeroes d_hist.
 * Replaces np.unique(labels, return_counts=True) in analysis_tools.py:8-18. */
int upcp_label_hist(const int32_t* d_labels, int64_t n, int64_t* d_hist256, void* stream);

/* ----------------------------------------------- AHN raster (S1) ---------- */
/* Raster = FastGridInterpolator state (interpolation.py:329-334):
 *   d_bin_x[nx] ascending  (= grid_x - step_x/2)
 *   d_bin_y[ny] descending (= grid_y + step_y/2)
 *   d_values[ny*nx] fp64 row-major (NaN where AHN is empty).
 * Cell lookup reproduces np.digitize(x, bin_x) - 1 and
 * np.digitize(y, bin_y, right=True) - 1 including numpy's negative-index wrap
 * (interpolation.py:346-348).
 *
 * upcp_raster_lookup: d_out[i] = values[y_idx, x_idx] for every i (or only where
 * d_sel[i] != 0; other entries are left untouched). */
int upcp_raster_lookup(const double* d_x, const double* d_y, int64_t n,
                       const uint8_t* d_sel,
                       const double* d_bin_x, int nx, const double* d_bin_y, int ny,
                       const double* d_values, double* d_out, void* stream);

/* Fused lookup + height-threshold classification.  v = raster value at the point.
 *   UPCP_CLS_ABS_LT      : |z - v| < a                 (ahn_fuser.py:159, ground)
 *   UPCP_CLS_LT          : z < v + a                   (ahn_fuser.py:170, building)
 *   UPCP_CLS_CAP_LE      : !isfinite(v) || z <= v + a  (building_fuser.py:89-95)
 *   UPCP_CLS_BAND_OC     : z > v + a && z <= v + b     (layer_lcc.py:72-74)
 *   UPCP_CLS_BAND_CC     : z <= v + b && z >= v - a    (ahn_fuser.py:115-116)
 *   UPCP_CLS_BELOW_NEG   : (z - v) < -a                (noise_filter.py:72)
 * d_out[i] = (d_sel ? d_sel[i] != 0 : 1) && cond.  If d_labels != NULL the label
 * `assign` is written where d_out[i] is set (fused `labels[label_mask] = label`).
 * d_out may be NULL when d_labels is given. */
#define UPCP_CLS_ABS_LT    0
#define UPCP_CLS_LT        1
#define UPCP_CLS_CAP_LE    2
#define UPCP_CLS_BAND_OC   3
#define UPCP_CLS_BAND_CC   4
#define UPCP_CLS_BELOW_NEG 5
int upcp_raster_classify(const double* d_x, const double* d_y, const double* d_z,
                         int64_t n, const uint8_t* d_sel,
                         const double* d_bin_x, int nx, const double* d_bin_y, int ny,
                         const double* d_values, int mode, double a, double b,
                         uint8_t* d_out, int32_t* d_labels, int32_t assign,
                         void* stream);

/* ------------------------------------------ BGT point-in-polygon (S2) ----- */
/* Replaces clip_utils.poly_clip / is_inside / _point_inside_poly
 * (clip_utils.py:119-238) OR-ed over polygons as building_fuser.py:83-87 and
 * road_fuser.py:84-87 do.  Rings are closed (first vertex == last vertex),
 * h_ring_xy is (total_vertices, 2) fp64 C-ordered, ring r uses vertices
 * h_ring_off[r] .. h_ring_off[r+1]-1.  h_ring_poly[r] is the polygon id
 * (non-decreasing), h_ring_hole[r] is 0 for the exterior ring (which must come
 * first within its polygon), 1 for an interior ring, 2 for a loop of an EVEN-ODD
 * polygon (the polygon is the XOR of its loops: the outer edges of an alpha complex,
 * alpha_shape_utils.py:40-92, stitched in any order).
 * Semantics per polygon: bbox-inclusive prefilter (clip_utils.py:22-40,219-236),
 * crossing number with boundary == inside; interiors clear what the exterior set.
 * The plan bins ring bounding boxes per uniform grid cell and ring edges per
 * per-ring y-slab; it is built on the host (O(vertices)) and uploaded once. */
typedef struct upcp_pip_plan upcp_pip_plan;
int    upcp_pip_plan_create(const double* h_ring_xy, const int64_t* h_ring_off,
                            const int32_t* h_ring_poly, const uint8_t* h_ring_hole,
                            int32_t n_rings, int32_t grid_dim /* 0 = auto */,
                            upcp_pip_plan** out);
/* The same for BUFFERED polygons: every polygon of the plan is taken as polygon.buffer(offset)
 * (BGTPolyReader.filter_tile(offset=...), bgt_utils.py:154-163; poly.buffer(0.05) in
 * AHNFuser._refine_layer, ahn_fuser.py:104-107).  The buffer is evaluated as a distance -- inside the
 * polygon or within `offset` of its boundary (offset > 0); inside and farther than |offset| from it
 * (offset < 0) -- i.e. the exact Minkowski sum where GEOS cuts the round joins into 8 segments per
 * quarter circle [ext, unverified: parity unpinned against GEOS].  Query such a plan with
 * upcp_pip_query_buffered. */
int    upcp_pip_plan_create_buffered(const double* h_ring_xy, const int64_t* h_ring_off,
                                     const int32_t* h_ring_poly, const uint8_t* h_ring_hole,
                                     int32_t n_rings, int32_t grid_dim, double offset, upcp_pip_plan** out);
/* Size of the DEVICE buffer the plan needs (the host blob plus the tag grid the device fills in). */
size_t upcp_pip_plan_bytes(const upcp_pip_plan* plan);
/* Copies the plan blob into d_plan (cudaMemcpyAsync from the plan's host blob) and classifies the cells of
 * its tag grid (three small kernels on `stream`). */
int    upcp_pip_plan_upload(const upcp_pip_plan* plan, void* d_plan, size_t bytes,
                            void* stream);
void   upcp_pip_plan_destroy(upcp_pip_plan* plan);
/* d_out[i] = (d_sel ? d_sel[i] != 0 : 1) && point i inside the polygon set. */
int    upcp_pip_query(const void* d_plan, const double* d_x, const double* d_y,
                      int64_t n, const uint8_t* d_sel, uint8_t* d_out, void* stream);
int    upcp_pip_query_buffered(const void* d_plan, const double* d_x, const double* d_y,
                               int64_t n, const uint8_t* d_sel, uint8_t* d_out, void* stream);

/* ------------------------------------ connected components (S3) ----------- */
/* Octree parameters exactly as CCCoreLib derives them (float32 arithmetic):
 * bbox of the float32 cloud -> CCMiscTools::MakeMinAndMaxCubical(enlarge 0.01)
 * -> cell size at MAX_OCTREE_LEVEL.  Host-only helper, no device work.
 *   h_bbox_sel6: fp64 min x,y,z / max x,y,z of the selected points
 *   h_out4     : float32 dimMin x,y,z and the level-`max_level` cell size
 * Replaces: the octree set-up inside
 * cccorelib.AutoSegmentationTools.labelConnectedComponents
 * (call site label_connected_comp.py:83-85). */
int upcp_octree_params(const double* h_bbox_sel6, int max_level, float* h_out4);

/* get_octree_level(points, grid_size) (math_utils.py:21-33), host helper working
 * on the fp64 bbox of ALL points passed to the processor. */
int upcp_octree_level(const double* h_bbox_all6, int ncols, double grid_size);

/* Connected components of occupied octree cells at `level` under 26-connectivity
 * for the selected points; per-point component ids.
 *   d_comp[n] (int32): component id >= 0 for selected points whose component has
 *       at least min_component_size points, -1 otherwise (also for unselected).
 *       Ids are arbitrary but deterministic; compare after canonical relabelling.
 *   d_info (int64[4]): [0] n selected, [1] n occupied cells, [2] n components
 *       (before size filter), [3] n components kept.
 * If d_labels != NULL the seed-fraction fill of _fill_components
 * (label_connected_comp.py:99-135) is fused: d_fill[n] = 1 for points of kept
 * components whose (double)seed_count / size > threshold, seeds being selected
 * points with d_labels[i] == seed_label.
 * Replaces: pycc.ccPointCloud + labelConnectedComponents + np.unique/np.in1d
 * filter + _fill_components (label_connected_comp.py:53-135). */
size_t upcp_lcc_ws_bytes(int64_t n, int64_t n_selected_upper, int level);
int upcp_lcc(const double* d_x, const double* d_y, const double* d_z, int64_t n,
             const uint8_t* d_sel, const float* h_octree4, int level, int max_level,
             int32_t min_component_size,
             const int32_t* d_labels, int32_t seed_label, double threshold,
             int32_t* d_comp, uint8_t* d_fill, int64_t* d_info,
             void* d_ws, size_t ws_bytes, void* stream);

/* ------------------------- neighbour search + normals + growth (S4/S5) ---- */
/* Exact k-nearest-neighbour graph + hybrid-search normals + curvature seed test
 * over the selected points, in one pass over a uniform-grid cell list.
 *   m            : number of selected points (ids below are "compact ids" in
 *                  ascending original-index order, i.e. position within
 *                  np.where(sel)[0] -- the index space region_growing.py uses);
 *   knn          : grow_region_knn (neighbours returned per point, query first);
 *   max_nn, radius : KDTreeSearchParamHybrid of estimate_normals;
 *   ordering     : ascending (squared distance, compact id) -- squared distance
 *                  is ((dx*dx + dy*dy) + dz*dz) in fp64 without FMA;
 *   d_knn[m*knn] : int32 compact ids (-1 padding when m < knn);
 *   d_normals[m*3] fp64, Open3D FastEigen3x3 of the cumulant covariance,
 *                  (0,0,1) when fewer than 3 neighbours / zero normal;
 *   d_cov[m*6]   : optional (may be NULL) covariance xx,xy,xz,yy,yz,zz of the knn
 *                  neighbourhood (compute_mean_and_covariance, region_growing.py:69-71);
 *   d_curv[m*3]  : optional analytic eigenvalue ratios eval_i / sum (ascending eval);
 *   d_seed_flag[m] : optional; the curvature seed test `eig_val[0] / sum < threshold_curve`
 *                  (region_growing.py:59-73,132) decided on the device where it does not
 *                  depend on LAPACK's eigenvalue order: 1 = all three analytic ratios are below
 *                  the threshold (minus 1e-9), 0 = all are at or above it (plus 1e-9),
 *                  2 = undecided (mixed sides / non-finite): the caller resolves these few on
 *                  the host with dgeev exactly as the reference does.
 * Replaces: o3d.geometry.KDTreeFlann + search_knn_vector_3d + estimate_normals +
 * compute_mean_and_covariance (region_growing.py:59-73,86-92,105-109). */
size_t upcp_knn_ws_bytes(int64_t n, int64_t m, int kmax);
int upcp_knn_normals(const double* d_x, const double* d_y, const double* d_z, int64_t n,
                     const uint8_t* d_sel, int64_t m, const double* h_bbox_sel6,
                     int knn, int max_nn, double radius, double cell_size,
                     int32_t* d_knn, double* d_knn_d2 /* may be NULL */,
                     double* d_normals, double* d_cov, double* d_curv,
                     double threshold_curve, uint8_t* d_seed_flag /* may be NULL */,
                     void* d_ws, size_t ws_bytes, void* stream);

/* The same search with every output in SORTED space: row s belongs to the s-th point of the
 * library's spatial (grid-cell) order, neighbours in d_knn are sorted positions, and
 * d_order[s] is the ORIGINAL point index (0..n-1) of sorted position s.  Spatially close points
 * get close row numbers, so the frontier expansion of upcp_region_grow gathers from cache
 * instead of HBM, and the epilogue writes its rows contiguously.  Membership results do not
 * depend on the numbering (the region is a least fixpoint); use upcp_mask_take / upcp_mask_put
 * to move masks between the two index spaces.
 * Replaces the same reference calls as upcp_knn_normals (region_growing.py:59-73,86-92,105-109). */
int upcp_knn_normals_sorted(const double* d_x, const double* d_y, const double* d_z, int64_t n,
                            const uint8_t* d_sel, int64_t m, const double* h_bbox_sel6,
                            int knn, int max_nn, double radius, double cell_size,
                            int32_t* d_knn, double* d_normals,
                            double* d_cov /* may be NULL; with d_seed_flag given only the rows with flag 2 are written */,
                            double threshold_curve, uint8_t* d_seed_flag /* may be NULL */,
                            uint32_t* d_order, void* d_ws, size_t ws_bytes, void* stream);
/* d_out[s] = d_full[d_order[s]], s < m          (region_growing.py:80-84: seeds of the masked cloud) */
int upcp_mask_take(const uint8_t* d_full, const uint32_t* d_order, int64_t m, uint8_t* d_out, void* stream);
/* d_full[0..n) = 0; d_full[d_order[s]] = 1 where d_sorted[s] != 0
 * (region_growing.py:139, `label_mask[mask_indices[region]] = True`) */
int upcp_mask_put(const uint8_t* d_sorted, const uint32_t* d_order, int64_t m, int64_t n, uint8_t* d_full,
                  void* stream);

/* Region growing to the least fixpoint (region_growing.py:75-141, method='knn').
 *   d_seed[m]     : 1 for initial seeds (points already carrying the label);
 *   d_can_seed[m] : 1 if an accepted point becomes a seed (curvature <
 *                   threshold_curve, region_growing.py:128-134);
 *   accept p from seed s iff p in knn(s)[1:] and
 *       degrees(acos(clip(dot(n_s/|n_s|, n_p/|n_p|), -1, 1))) < threshold_angle
 *       (math_utils.py:9-18);
 *   d_region[m]   : output, 1 for every point of the grown region (seeds included).
 *   d_info (int64[2]) : [0] region size, [1] number of points that went through the work queue
 *                   (the initial seeds are expanded data-parallel and are not counted).
 * d_can_seed may hold the value 2 = "undecided": such a point is treated as not (yet) a seed. */
size_t upcp_region_grow_ws_bytes(int64_t m);
int upcp_region_grow(const int32_t* d_knn, int knn, const double* d_normals,
                     const uint8_t* d_seed, const uint8_t* d_can_seed, int64_t m,
                     double threshold_angle_deg, uint8_t* d_region, int64_t* d_info,
                     void* d_ws, size_t ws_bytes, void* stream);

/* The same growth in three asynchronous steps, so that the host can decide the seed flags the
 * device left undecided (upcp_knn_normals: d_seed_flag == 2, the curvature test that depends on
 * LAPACK's eigenvalue order, region_growing.py:69-73) WHILE the device grows:
 *   begin   grows from d_seed with the flags as they are (2 = not a seed); returns without
 *           synchronising; the state lives in the caller's workspace (same buffer for all three);
 *   resume  after the caller rewrote d_can_seed for the n_extra points listed in d_extra (int64
 *           indices): continues from those that are in the region, were no seeds and are seeds now.
 *           Growth is monotone, so the result is the least fixpoint for the final flags;
 *   end     writes d_region / d_info and synchronises (work-queue error check). */
int upcp_region_grow_begin(const int32_t* d_knn, int knn, const double* d_normals,
                           const uint8_t* d_seed, const uint8_t* d_can_seed, int64_t m,
                           double threshold_angle_deg, void* d_ws, size_t ws_bytes, void* stream);
int upcp_region_grow_resume(const int32_t* d_knn, int knn, const double* d_normals,
                            const uint8_t* d_seed, const uint8_t* d_can_seed, int64_t m,
                            const int64_t* d_extra, int64_t n_extra,
                            double threshold_angle_deg, void* d_ws, size_t ws_bytes, void* stream);
int upcp_region_grow_end(int64_t m, uint8_t* d_region, int64_t* d_info, void* d_ws, size_t ws_bytes,
                         void* stream);

/* Region growing over FIXED-RADIUS neighbourhoods: RegionGrowing._region_growing(method='radius')
 * (region_growing.py:59-73,104-109; KDTreeFlann.search_radius_vector_3d: squared distance < radius^2,
 * ascending, first returned index dropped).  Neighbourhoods are walked on a uniform grid, never stored.
 * All per-point arrays are in COMPACT-id space (position within np.where(sel)[0]); the four calls share one
 * workspace and must be given the same n, m, bounding box and radius.
 *   prepare     builds the grid; d_flag[m] receives the curvature seed flag of every point (1 seed, 0 not,
 *               2 undecided: the reference's unshifted single-pass covariance could fall on either side --
 *               resolve from upcp_radius_neighbours with LAPACK); d_nb_count[m] (optional) the neighbourhood
 *               sizes; the initial seeds (d_seed) are queued;
 *   grow        d_extra == NULL: grows from the queued initial seeds with d_can_seed (2 = not a seed);
 *               d_extra != NULL: resumes from those of the n_extra listed points that are in the region and
 *               are seeds now (growth is monotone: same least fixpoint);
 *   region      d_region[m] = 1 for the grown region, d_info[2] = region size, seeds expanded; synchronises;
 *   neighbours  for nq query points the neighbours within radius as (x, y, z, compact id) quadruples:
 *               query i owns d_out4[4 * d_offsets2[2i] ...] with d_offsets2[2i + 1] entries (unordered;
 *               *d_total > cap means the buffer was too small: call again with a larger one). */
size_t upcp_radius_ws_bytes(int64_t n, int64_t m);
int upcp_radius_prepare(const double* d_x, const double* d_y, const double* d_z, int64_t n, const uint8_t* d_sel, int64_t m,
                        const double* h_bbox_sel6, double radius, const double* d_normals, const uint8_t* d_seed,
                        double threshold_curve, uint8_t* d_flag, uint32_t* d_nb_count, void* d_ws, size_t ws_bytes, void* stream);
int upcp_radius_grow(int64_t n, int64_t m, const double* h_bbox_sel6, double radius, const uint8_t* d_can_seed,
                     double threshold_angle_deg, const int64_t* d_extra, int64_t n_extra,
                     void* d_ws, size_t ws_bytes, void* stream);
int upcp_radius_region(int64_t n, int64_t m, uint8_t* d_region, int64_t* d_info, void* d_ws, size_t ws_bytes, void* stream);
int upcp_radius_neighbours(int64_t n, int64_t m, const double* h_bbox_sel6, double radius, const int64_t* d_queries, int64_t nq,
                           int64_t* d_offsets2, double* d_out4, int64_t cap, int64_t* d_total, void* d_ws, size_t ws_bytes,
                           void* stream);

/* Scatter a compact per-selected-point mask back to the full tile:
 * d_out[i] = d_compact[rank(i)] for selected i, 0 otherwise
 * (region_growing.py:139, `label_mask[mask_indices[region]] = True`). */
size_t upcp_compact_ws_bytes(int64_t n);
int upcp_mask_scatter(const uint8_t* d_sel, int64_t n, const uint8_t* d_compact,
                      uint8_t* d_out, void* d_ws, size_t ws_bytes, void* stream);
/* Gather: d_compact[rank(i)] = d_full[i] for selected i; *d_count = m. */
int upcp_mask_gather(const uint8_t* d_sel, int64_t n, const uint8_t* d_full,
                     uint8_t* d_compact, int64_t* d_count,
                     void* d_ws, size_t ws_bytes, void* stream);

/* ------------------------------------------- per-object reductions ("next" row 2) ---- */
/* Replaces the per-cluster Python loops of the clients of LabelConnectedComp.get_components:
 * CarFuser._label_car_like_components (reference fusion/car_fuser.py:45-89) and
 * BGTStreetFurnitureFuser._label_street_furniture_like_components
 * (fusion/street_furniture_fuser.py:46-89), which rescan the masked cloud once per cluster
 * (`cc_mask = point_components == cc`, `ground_z[cc_mask]`, `np.amax(points[cc_mask][:, 2])`).
 *
 * d_comp [n]: cluster id per point as upcp_lcc writes it (any value in [0, n), -1 = none).
 * One pass groups the points by cluster and reduces every cluster:
 *   d_dense [n]          dense cluster id 0..K-1 (ascending d_comp value) or -1
 *   d_order [n]          original point indices grouped by dense id, ascending inside a group
 *                        (the order of points[mask][cc_mask], i.e. the qhull input order)
 *   d_seg_first [cap+1]  group k is d_order[d_seg_first[k] .. d_seg_first[k+1])
 *   d_count_finite [cap] number of points with a finite d_gz (np.isfinite, car_fuser.py:62)
 *   d_sum_fixed [cap]    sum of those d_gz in units of 2^-24 m (integer: order-independent;
 *                        exact for float16-derived AHN values)   mean = sum / 2^24 / count
 *   d_max_z [cap]        np.amax of z over the cluster (car_fuser.py:68)
 *   d_bbox_xy [cap][4]   x_min, y_min, x_max, y_max of the cluster
 *   d_info [4]           K, points in clusters, 0, points whose |d_gz| >= 2^20 (ignored)
 * d_gz may be NULL (no elevation statistics).  Synchronises the stream twice.
 * UPCP_E_CAPACITY when K > cap (d_info[0] still holds K). */
size_t upcp_component_stats_ws_bytes(int64_t n);
int upcp_component_stats(const int32_t* d_comp, int64_t n, const double* d_x, const double* d_y,
                         const double* d_z, const double* d_gz, int32_t cap, int32_t* d_dense,
                         uint32_t* d_order, int64_t* d_seg_first, int64_t* d_count_finite,
                         int64_t* d_sum_fixed, double* d_max_z, double* d_bbox_xy, int64_t* d_info,
                         void* d_ws, size_t ws_bytes, void* stream);
/* d_out_xy[s] = (d_x[d_idx[s]], d_y[d_idx[s]]): the xy of one cluster for the host-side
 * convex hull / minimum bounding rectangle (math_utils.py:62-125). */
int upcp_gather_xy(const double* d_x, const double* d_y, const uint32_t* d_idx, int64_t cnt,
                   double* d_out_xy, void* stream);
/* d_out[i] = d_dense[i] >= 0 && d_accepted[d_dense[i]]
 * (`street_furniture_mask[cc_mask] = True`, street_furniture_fuser.py:80). */
int upcp_component_mark(const int32_t* d_dense, int64_t n, const uint8_t* d_accepted,
                        int32_t n_comp, uint8_t* d_out, void* stream);
/* clip_utils.poly_box_clip (clip_utils.py:240-262) for a list of hole-free rings at once:
 * d_out[i] = OR over prisms p of (inclusive bbox prefilter && crossing-number test of ring p,
 * boundary == inside && d_zrange[p][0] <= z <= d_zrange[p][1]), restricted to d_sel.
 * Rings are closed (first == last vertex): ring p = d_ring_xy[d_ring_off[p] .. d_ring_off[p+1]). */
int upcp_prism_clip(const double* d_x, const double* d_y, const double* d_z, int64_t n,
                    const uint8_t* d_sel, const double* d_ring_xy, const int32_t* d_ring_off,
                    const double* d_zrange, int32_t n_prisms, uint8_t* d_out, void* stream);

/* ---------------------------------------- LAZ ingest ("next" row 1), host ---- */
/* Decode LASzip-compressed point records (what laspy[lazrs] does for las_utils.read_las,
 * las_utils.py:118-120).  Host code, no device work.  Covers the "pointwise chunked" compressor with the
 * version-2 item codecs of point formats 0-3 (POINT10, GPSTIME11, RGB12, extra BYTEs); LAS 1.4 layered
 * items return UPCP_E_CAPACITY.
 *   h_vlr / vlr_bytes : payload of the "laszip encoded" VLR (record id 22204);
 *   h_data            : the file from offset-to-point-data to its end (data_bytes), which sits at
 *                       data_file_offset in the file (the chunk table is addressed by file position);
 *   h_out             : n_points * point_size bytes, the records as an uncompressed .las holds them. */
int upcp_laz_decode(const uint8_t* h_vlr, int32_t vlr_bytes, const uint8_t* h_data, int64_t data_bytes,
                    int64_t data_file_offset, int64_t n_points, int32_t point_size, uint8_t* h_out);

/* The inverse: LASzip-compress point records for las_utils.label_and_save_las_as_laz (las_utils.py:133-164;
 * laspy writes `.laz` through lazrs).  Host code.  Point formats 0-3 (+ extra bytes), "pointwise chunked"
 * compressor, version-2 item codecs -- what upcp_laz_decode reads.
 *   h_points        : n_points * point_size bytes, uncompressed records;
 *   data_file_offset: where the returned bytes will sit in the file (= offset to point data of the header);
 *   h_vlr58         : receives the payload of the "laszip encoded" VLR (record id 22204), *vlr_bytes of it;
 *   h_out / out_cap : receives the point data section (chunk-table position, chunks, chunk table);
 *                     *out_bytes is always set, UPCP_E_WORKSPACE means "call again with at least that much". */
int upcp_laz_encode(const uint8_t* h_points, int64_t n_points, int32_t point_size, int32_t point_format, int32_t chunk_size,
                    int64_t data_file_offset, uint8_t* h_vlr58, int32_t* vlr_bytes, uint8_t* h_out, int64_t out_cap,
                    int64_t* out_bytes);

/* ------------------------------------ AHN preprocessing ("next" row 4) ---- */
/* SpatialInterpolator.__call__ (reference utils/interpolation.py:156-308) for 2-D data, as
 * preprocessing/ahn_preprocessing.py:129-236 uses it to rasterise an AHN cloud: for every position the
 * n_neighbors (<= 32) nearest data points with distance < max_dist (scipy cKDTree.query semantics),
 *   method 0 `idw`: value of the nearest point if it lies within conf_dist, else
 *                   sum(w v) / sum(w), w = 1 / (d ** power + reg), in ascending (distance, index) order;
 *   method 1 `max`: maximum of their values;
 * fill_value where no point lies within max_dist.  h_bbox4 = x_min, y_min, x_max, y_max of the data points.
 * Replaces: scipy cKDTree + the per-position Python loop of interpolation.py:276-303. */
size_t upcp_interpolate_ws_bytes(int64_t n);
int upcp_interpolate(const double* d_px, const double* d_py, const double* d_pv, int64_t n, const double* h_bbox4,
                     const double* d_qx, const double* d_qy, int64_t nq, int method, int n_neighbors,
                     double max_dist, double power, double reg, double conf_dist, double fill_value,
                     double* d_out, void* d_ws, size_t ws_bytes, void* stream);

/* ------------------------------------------ synthetic tiles (measurement) ---- */
/* synthetic urban tile i generated ON THE DEVICE from seed 20240000 + i (the
 * reference ships no clouds and 512 x 20 M points cannot be fed from the host).  Scene model and value
 * distributions of synthetic.make_tile: ground, facades (+ balconies), car shells, poles / crowns,
 * clutter, uniform noise, millimetre grid; one counter-based Philox4x32-10 stream per point.  Used by
 * bench.py and the full-size tests, not by the labelling path.
 *   h_origin_size3 : tile origin x, y and edge length;  h_cum_kind6 : cumulative point fractions of the
 *   six kinds (0 ground .. 5 noise);  d_walls8 [n][8] x0, y0, x1, y1, height, outward normal x, y,
 *   cumulative area share;  d_cars4 centre x, y, size x, y;  d_poles4 x, y, radius, height;
 *   d_crowns4 x, y, radius, centre z.  d_kind receives the kind of every point. */
int upcp_synth_tile(int64_t n, uint64_t seed, const double* h_origin_size3, const double* h_cum_kind6,
                    const double* d_walls8, int32_t n_walls, const double* d_cars4, int32_t n_cars,
                    const double* d_poles4, int32_t n_poles, const double* d_crowns4, int32_t n_crowns,
                    double* d_x, double* d_y, double* d_z, uint8_t* d_kind, void* stream);

/* ------------------------------------------------ primitives (exposed) ---- */
/* LSD radix sort of (key, value) pairs, the "Morton/voxel-key radix sort" of the
 * hot path; exposed for tests and benchmarks.  Sorts bits [0, end_bit).
 * Result is written to d_keys_out / d_vals_out (inputs are clobbered). */
size_t upcp_sort_ws_bytes(int64_t n);
int upcp_sort_pairs_u32(uint32_t* d_keys_in, uint32_t* d_vals_in,
                        uint32_t* d_keys_out, uint32_t* d_vals_out, int64_t n,
                        int end_bit, void* d_ws, size_t ws_bytes, void* stream);
int upcp_sort_pairs_u64(uint64_t* d_keys_in, uint32_t* d_vals_in,
                        uint64_t* d_keys_out, uint32_t* d_vals_out, int64_t n,
                        int end_bit, void* d_ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UPCP_CUDA_H */
