/*
 * upcp_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product package).
 *
 * Plain-C CPU restatement of the native code the reference calls on its per-tile
 * labelling hot path.  Written independently of the CUDA kernels (it shares no source
 * with urban_pointcloud_processing_b200/csrc) so that the two can check each other.
 * Compile with:  gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math  (see oracle/build.py)
 *
 *   orc_is_inside        numba clip_utils.is_inside / _point_inside_poly
 *                        (reference src/upcp/utils/clip_utils.py:119-190)      PINNED by
 *                        golden vectors generated from the real numba code.
 *   orc_label_cc         cccorelib.AutoSegmentationTools.labelConnectedComponents
 *                        (call site label_connected_comp.py:83-85).  The algorithm lives
 *                        in CCCoreLib (un-vendored, unpinned git master): DgmOctree
 *                        build (float32 cubical bbox, 2^21 cells per axis), extractCCs at
 *                        `level` with 26-connectivity.  PARITY UNPINNED (source absent).
 *   orc_knn / orc_normals_hybrid / orc_cov_knn
 *                        open3d 0.16 KDTreeFlann.search_knn_vector_3d,
 *                        estimate_normals(KDTreeSearchParamHybrid), compute_mean_and_
 *                        covariance (call sites region_growing.py:59-73,86-92,105-109).
 *                        Restates the published algorithms (nanoflann L2 ordering,
 *                        utility::ComputeCovariance cumulants, FastEigen3x3).  Ties are
 *                        ordered by (squared distance, point index).  PARITY UNPINNED.
 *   orc_region_grow      RegionGrowing._region_growing(method='knn') queue loop
 *                        (region_growing.py:75-141) over precomputed kNN rows / normals.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ PIP ------ */
/* clip_utils._point_inside_poly, literal (clip_utils.py:119-164). */
static int point_inside_poly(const double* poly, int n_vert, double px, double py) {
    int length = n_vert - 1;
    double dy2 = py - poly[1];
    int intersections = 0;
    int ii = 0, jj = 1;
    while (ii < length) {
        double dy = dy2;
        dy2 = py - poly[2 * jj + 1];
        if (dy * dy2 <= 0.0 && (px >= poly[2 * ii] || px >= poly[2 * jj])) {
            if (dy < 0 || dy2 < 0) {
                double F = dy * (poly[2 * jj] - poly[2 * ii]) / (dy - dy2) + poly[2 * ii];
                if (px > F) intersections += 1;
                else if (px == F) return 2;
            } else if (dy2 == 0 && (px == poly[2 * jj] ||
                                    (dy == 0 && (px - poly[2 * ii]) * (px - poly[2 * jj]) <= 0))) {
                return 2;
            }
        }
        ii = jj;
        jj += 1;
    }
    return intersections & 1;
}

/* clip_utils.is_inside (clip_utils.py:167-190): the int result is stored into a boolean. */
void orc_is_inside(const double* x, const double* y, int64_t n, const double* poly, int n_vert,
                   uint8_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = point_inside_poly(poly, n_vert, x[i], y[i]) != 0;
}

/* ------------------------------------------------------- connected components -- */
typedef struct { uint64_t key; int64_t idx; } KeyIdx;
static int cmp_keyidx(const void* a, const void* b) {
    const KeyIdx* ka = (const KeyIdx*)a; const KeyIdx* kb = (const KeyIdx*)b;
    if (ka->key < kb->key) return -1;
    if (ka->key > kb->key) return 1;
    return (ka->idx > kb->idx) - (ka->idx < kb->idx);
}
static int64_t uf_root(int64_t* parent, int64_t a) {
    while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; }
    return a;
}
static int64_t find_cell(const uint64_t* cells, int64_t n_cells, uint64_t key) {
    int64_t lo = 0, hi = n_cells - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) / 2;
        if (cells[mid] == key) return mid;
        if (cells[mid] < key) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

/* Octree set-up as CCCoreLib does it, float32 arithmetic [ext, unverified]:
 *   bbox of the cloud -> CCMiscTools::MakeMinAndMaxCubical(dimMin, dimMax, 0.01)
 *   -> cellSize[0] = dimMax.x - dimMin.x, halved per level.  out4 = dimMin xyz, cs(max_level) */
void orc_octree_params(const float* xs, const float* ys, const float* zs, int64_t m, int max_level,
                       float* out4) {
    float mn[3] = {xs[0], ys[0], zs[0]}, mx[3] = {xs[0], ys[0], zs[0]};
    for (int64_t i = 1; i < m; ++i) {
        if (xs[i] < mn[0]) mn[0] = xs[i]; if (xs[i] > mx[0]) mx[0] = xs[i];
        if (ys[i] < mn[1]) mn[1] = ys[i]; if (ys[i] > mx[1]) mx[1] = ys[i];
        if (zs[i] < mn[2]) mn[2] = zs[i]; if (zs[i] > mx[2]) mx[2] = zs[i];
    }
    float diag[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
    float max_dd = diag[0] > diag[1] ? diag[0] : diag[1];
    max_dd = max_dd > diag[2] ? max_dd : diag[2];
    max_dd = (float)((double)max_dd * (1.0 + 0.01));
    float dim_min[3], dim_max[3];
    for (int d = 0; d < 3; ++d) {
        float md = mx[d] + mn[d];
        dim_min[d] = (md - max_dd) * 0.5f;
        dim_max[d] = dim_min[d] + max_dd;
    }
    float cs = dim_max[0] - dim_min[0];
    for (int k = 1; k <= max_level; ++k) cs = cs / 2.0f;
    out4[0] = dim_min[0]; out4[1] = dim_min[1]; out4[2] = dim_min[2]; out4[3] = cs;
}

/* labels_out[i] in 1..K (component of the cell of point i).  Returns K, or -1 on OOM. */
int64_t orc_label_cc(const float* xs, const float* ys, const float* zs, int64_t m, int level,
                     int max_level, int32_t* labels_out) {
    if (m <= 0) return 0;
    float p4[4];
    orc_octree_params(xs, ys, zs, m, max_level, p4);
    const float cs = p4[3];
    const int max_pos = (1 << max_level) - 1;
    const int shift = max_level - level;
    KeyIdx* ki = (KeyIdx*)malloc((size_t)m * sizeof(KeyIdx));
    uint64_t* cells = (uint64_t*)malloc((size_t)m * sizeof(uint64_t));
    int64_t* cell_of = (int64_t*)malloc((size_t)m * sizeof(int64_t));
    if (!ki || !cells || !cell_of) { free(ki); free(cells); free(cell_of); return -1; }
    for (int64_t i = 0; i < m; ++i) {
        const float p[3] = {xs[i], ys[i], zs[i]};
        int pos[3];
        for (int d = 0; d < 3; ++d) {
            /* getTheCellPosWhichIncludesThePoint: (int)((P - dimMin) / cs), then clipped */
            float t = (p[d] - p4[d]) / cs;
            int c;
            if (!(t > 0.0f)) c = 0;
            else if (t >= (float)max_pos) c = max_pos;
            else c = (int)t;
            pos[d] = c >> shift;
        }
        ki[i].key = ((uint64_t)pos[2] << (2 * level)) | ((uint64_t)pos[1] << level) | (uint64_t)pos[0];
        ki[i].idx = i;
    }
    qsort(ki, (size_t)m, sizeof(KeyIdx), cmp_keyidx);
    int64_t n_cells = 0;
    for (int64_t s = 0; s < m; ++s) {
        if (s == 0 || ki[s].key != ki[s - 1].key) cells[n_cells++] = ki[s].key;
        cell_of[ki[s].idx] = n_cells - 1;
    }
    int64_t* parent = (int64_t*)malloc((size_t)n_cells * sizeof(int64_t));
    int32_t* comp = (int32_t*)malloc((size_t)n_cells * sizeof(int32_t));
    if (!parent || !comp) { free(ki); free(cells); free(cell_of); free(parent); free(comp); return -1; }
    for (int64_t c = 0; c < n_cells; ++c) parent[c] = c;
    const int64_t dim = (int64_t)1 << level;
    const uint64_t amask = ((uint64_t)1 << level) - 1;
    /* extractCCs(level, sixConnexity=false): occupied cells are connected when they touch
     * by a face, an edge or a corner (26-neighbourhood) */
    for (int64_t c = 0; c < n_cells; ++c) {
        int64_t cx = (int64_t)(cells[c] & amask), cy = (int64_t)((cells[c] >> level) & amask),
                cz = (int64_t)((cells[c] >> (2 * level)) & amask);
        for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
            if (!dx && !dy && !dz) continue;
            int64_t nx = cx + dx, ny = cy + dy, nz = cz + dz;
            if (nx < 0 || ny < 0 || nz < 0 || nx >= dim || ny >= dim || nz >= dim) continue;
            uint64_t nk = ((uint64_t)nz << (2 * level)) | ((uint64_t)ny << level) | (uint64_t)nx;
            int64_t nc = find_cell(cells, n_cells, nk);
            if (nc < 0) continue;
            int64_t ra = uf_root(parent, c), rb = uf_root(parent, nc);
            if (ra != rb) { if (ra < rb) parent[rb] = ra; else parent[ra] = rb; }
        }
    }
    int64_t k = 0;
    for (int64_t c = 0; c < n_cells; ++c) comp[c] = 0;
    for (int64_t c = 0; c < n_cells; ++c) {
        int64_t r = uf_root(parent, c);
        if (comp[r] == 0) comp[r] = (int32_t)(++k);
        comp[c] = comp[r];
    }
    for (int64_t i = 0; i < m; ++i) labels_out[i] = comp[cell_of[i]];
    free(ki); free(cells); free(cell_of); free(parent); free(comp);
    return k;
}

/* ------------------------------------------------------------- kNN (exact) ----- */
typedef struct {
    double ox, oy, oz, cell;
    int nx, ny, nz;
    int64_t* start;   /* [ncells + 1] */
    int32_t* order;   /* point ids sorted by cell (ascending id inside a cell) */
} Grid;

static int grid_coord(double v, double o, double cell, int n) {
    double t = (v - o) / cell;
    if (!(t > 0.0)) return 0;
    if (t >= (double)n) return n - 1;
    return (int)floor(t);
}

static int grid_build(Grid* g, const double* pts, int64_t m) {
    double mn[3] = {pts[0], pts[1], pts[2]}, mx[3] = {pts[0], pts[1], pts[2]};
    for (int64_t i = 1; i < m; ++i)
        for (int d = 0; d < 3; ++d) {
            if (pts[3 * i + d] < mn[d]) mn[d] = pts[3 * i + d];
            if (pts[3 * i + d] > mx[d]) mx[d] = pts[3 * i + d];
        }
    double ext = fmax(mx[0] - mn[0], fmax(mx[1] - mn[1], mx[2] - mn[2]));
    /* aim at a handful of points per cell, bounded table size */
    double cell = ext / 200.0;
    double vol = fmax(mx[0] - mn[0], 1e-9) * fmax(mx[1] - mn[1], 1e-9) * fmax(mx[2] - mn[2], 1e-9);
    double target = cbrt(vol / ((double)m / 4.0 + 1.0));
    if (target > cell) cell = target;
    if (!(cell > 0)) cell = 1.0;
    for (;;) {
        double nc = (floor((mx[0] - mn[0]) / cell) + 1) * (floor((mx[1] - mn[1]) / cell) + 1) *
                    (floor((mx[2] - mn[2]) / cell) + 1);
        if (nc <= 6.4e7) break;
        cell *= 1.5;
    }
    g->ox = mn[0]; g->oy = mn[1]; g->oz = mn[2]; g->cell = cell;
    g->nx = (int)floor((mx[0] - mn[0]) / cell) + 1;
    g->ny = (int)floor((mx[1] - mn[1]) / cell) + 1;
    g->nz = (int)floor((mx[2] - mn[2]) / cell) + 1;
    int64_t ncell = (int64_t)g->nx * g->ny * g->nz;
    g->start = (int64_t*)calloc((size_t)ncell + 1, sizeof(int64_t));
    g->order = (int32_t*)malloc((size_t)m * sizeof(int32_t));
    int64_t* cid = (int64_t*)malloc((size_t)m * sizeof(int64_t));
    if (!g->start || !g->order || !cid) { free(cid); return -1; }
    for (int64_t i = 0; i < m; ++i) {
        int cx = grid_coord(pts[3 * i], g->ox, cell, g->nx), cy = grid_coord(pts[3 * i + 1], g->oy, cell, g->ny),
            cz = grid_coord(pts[3 * i + 2], g->oz, cell, g->nz);
        cid[i] = ((int64_t)cz * g->ny + cy) * g->nx + cx;
        g->start[cid[i] + 1]++;
    }
    for (int64_t c = 0; c < ncell; ++c) g->start[c + 1] += g->start[c];
    int64_t* cur = (int64_t*)malloc((size_t)ncell * sizeof(int64_t));
    if (!cur) { free(cid); return -1; }
    memcpy(cur, g->start, (size_t)ncell * sizeof(int64_t));
    for (int64_t i = 0; i < m; ++i) g->order[cur[cid[i]]++] = (int32_t)i;
    free(cur); free(cid);
    return 0;
}
static void grid_free(Grid* g) { free(g->start); free(g->order); }

/* squared distance as the ordering defines it: ((dx*dx + dy*dy) + dz*dz), no FMA */
static double sqdist(const double* a, const double* b) {
    double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return (dx * dx + dy * dy) + dz * dz;
}

/* k nearest of query q in ascending (d2, id) order; returns the count (min(k, m)). */
static int grid_knn(const Grid* g, const double* pts, const double* q, int k, int32_t* best_id,
                    double* best_d2) {
    int cnt = 0;
    int qx = grid_coord(q[0], g->ox, g->cell, g->nx), qy = grid_coord(q[1], g->oy, g->cell, g->ny),
        qz = grid_coord(q[2], g->oz, g->cell, g->nz);
    for (int r = 0;; ++r) {
        int x0 = qx - r, x1 = qx + r, y0 = qy - r, y1 = qy + r, z0 = qz - r, z1 = qz + r;
        for (int cz = z0 < 0 ? 0 : z0; cz <= (z1 >= g->nz ? g->nz - 1 : z1); ++cz)
            for (int cy = y0 < 0 ? 0 : y0; cy <= (y1 >= g->ny ? g->ny - 1 : y1); ++cy)
                for (int cx = x0 < 0 ? 0 : x0; cx <= (x1 >= g->nx ? g->nx - 1 : x1); ++cx) {
                    if (r > 0 && cx > x0 && cx < x1 && cy > y0 && cy < y1 && cz > z0 && cz < z1) continue;
                    int64_t c = ((int64_t)cz * g->ny + cy) * g->nx + cx;
                    for (int64_t s = g->start[c]; s < g->start[c + 1]; ++s) {
                        int32_t id = g->order[s];
                        double d2 = sqdist(pts + 3 * (int64_t)id, q);
                        if (cnt == k && !(d2 < best_d2[k - 1] || (d2 == best_d2[k - 1] && id < best_id[k - 1])))
                            continue;
                        int j = cnt < k ? cnt++ : k - 1;
                        while (j > 0 && (best_d2[j - 1] > d2 || (best_d2[j - 1] == d2 && best_id[j - 1] > id))) {
                            best_d2[j] = best_d2[j - 1]; best_id[j] = best_id[j - 1]; --j;
                        }
                        best_d2[j] = d2; best_id[j] = id;
                    }
                }
        int open = x0 > 0 || y0 > 0 || z0 > 0 || x1 < g->nx - 1 || y1 < g->ny - 1 || z1 < g->nz - 1;
        if (!open) break;
        if (cnt == k) {
            double reach = INFINITY;
            if (x0 > 0) reach = fmin(reach, q[0] - (g->ox + x0 * g->cell));
            if (x1 < g->nx - 1) reach = fmin(reach, (g->ox + (x1 + 1) * g->cell) - q[0]);
            if (y0 > 0) reach = fmin(reach, q[1] - (g->oy + y0 * g->cell));
            if (y1 < g->ny - 1) reach = fmin(reach, (g->oy + (y1 + 1) * g->cell) - q[1]);
            if (z0 > 0) reach = fmin(reach, q[2] - (g->oz + z0 * g->cell));
            if (z1 < g->nz - 1) reach = fmin(reach, (g->oz + (z1 + 1) * g->cell) - q[2]);
            reach -= 1e-6 * g->cell + 1e-9;
            if (reach > 0 && best_d2[k - 1] < reach * reach * (1.0 - 1e-9)) break;
        }
    }
    return cnt;
}

/* pts: (m,3) C-ordered; idx_out (m,k) padded with -1, d2_out (m,k) padded with inf (may be NULL) */
int orc_knn(const double* pts, int64_t m, int k, int32_t* idx_out, double* d2_out) {
    if (m <= 0) return 0;
    Grid g;
    if (grid_build(&g, pts, m)) return -1;
#pragma omp parallel
    {
        int32_t* bid = (int32_t*)malloc((size_t)k * sizeof(int32_t));
        double* bd = (double*)malloc((size_t)k * sizeof(double));
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < m; ++i) {
            int c = grid_knn(&g, pts, pts + 3 * i, k, bid, bd);
            for (int j = 0; j < k; ++j) {
                idx_out[i * k + j] = j < c ? bid[j] : -1;
                if (d2_out) d2_out[i * k + j] = j < c ? bd[j] : INFINITY;
            }
        }
        free(bid); free(bd);
    }
    grid_free(&g);
    return 0;
}

/* ---------------------------------------------- covariance / normals (Open3D) -- */
/* utility::ComputeCovariance, cumulant form [ext, unverified]; cov6 = xx xy xz yy yz zz */
static void covariance_of(const double* pts, const int32_t* ids, int n, double* cov6) {
    double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < n; ++j) {
        const double* p = pts + 3 * (int64_t)ids[j];
        c[0] += p[0]; c[1] += p[1]; c[2] += p[2];
        c[3] += p[0] * p[0]; c[4] += p[0] * p[1]; c[5] += p[0] * p[2];
        c[6] += p[1] * p[1]; c[7] += p[1] * p[2]; c[8] += p[2] * p[2];
    }
    for (int i = 0; i < 9; ++i) c[i] /= (double)n;
    cov6[0] = c[3] - c[0] * c[0];
    cov6[1] = c[4] - c[0] * c[1];
    cov6[2] = c[5] - c[0] * c[2];
    cov6[3] = c[6] - c[1] * c[1];
    cov6[4] = c[7] - c[1] * c[2];
    cov6[5] = c[8] - c[2] * c[2];
}

typedef struct { double v[3]; } V3;
static V3 v3cross(V3 a, V3 b) {
    V3 r = {{a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]}};
    return r;
}
static double v3dot(V3 a, V3 b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }
#define A_(i, j) A[(i) * 3 + (j)]

/* EstimateNormals.cpp ComputeEigenvector0 [ext, unverified] */
static V3 eigvec0(const double* A, double eval0) {
    V3 row0 = {{A_(0, 0) - eval0, A_(0, 1), A_(0, 2)}};
    V3 row1 = {{A_(0, 1), A_(1, 1) - eval0, A_(1, 2)}};
    V3 row2 = {{A_(0, 2), A_(1, 2), A_(2, 2) - eval0}};
    V3 r0xr1 = v3cross(row0, row1), r0xr2 = v3cross(row0, row2), r1xr2 = v3cross(row1, row2);
    double d0 = v3dot(r0xr1, r0xr1), d1 = v3dot(r0xr2, r0xr2), d2 = v3dot(r1xr2, r1xr2);
    double dmax = d0; int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    V3 src = imax == 0 ? r0xr1 : (imax == 1 ? r0xr2 : r1xr2);
    double s = sqrt(imax == 0 ? d0 : (imax == 1 ? d1 : d2));
    V3 r = {{src.v[0] / s, src.v[1] / s, src.v[2] / s}};
    return r;
}

/* EstimateNormals.cpp ComputeEigenvector1 [ext, unverified] */
static V3 eigvec1(const double* A, V3 e0, double eval1) {
    V3 U, V;
    if (fabs(e0.v[0]) > fabs(e0.v[1])) {
        double inv_length = 1 / sqrt(e0.v[0] * e0.v[0] + e0.v[2] * e0.v[2]);
        U.v[0] = -e0.v[2] * inv_length; U.v[1] = 0; U.v[2] = e0.v[0] * inv_length;
    } else {
        double inv_length = 1 / sqrt(e0.v[1] * e0.v[1] + e0.v[2] * e0.v[2]);
        U.v[0] = 0; U.v[1] = e0.v[2] * inv_length; U.v[2] = -e0.v[1] * inv_length;
    }
    V = v3cross(e0, U);
    V3 AU = {{A_(0, 0) * U.v[0] + A_(0, 1) * U.v[1] + A_(0, 2) * U.v[2],
              A_(0, 1) * U.v[0] + A_(1, 1) * U.v[1] + A_(1, 2) * U.v[2],
              A_(0, 2) * U.v[0] + A_(1, 2) * U.v[1] + A_(2, 2) * U.v[2]}};
    V3 AV = {{A_(0, 0) * V.v[0] + A_(0, 1) * V.v[1] + A_(0, 2) * V.v[2],
              A_(0, 1) * V.v[0] + A_(1, 1) * V.v[1] + A_(1, 2) * V.v[2],
              A_(0, 2) * V.v[0] + A_(1, 2) * V.v[1] + A_(2, 2) * V.v[2]}};
    double m00 = U.v[0] * AU.v[0] + U.v[1] * AU.v[1] + U.v[2] * AU.v[2] - eval1;
    double m01 = U.v[0] * AV.v[0] + U.v[1] * AV.v[1] + U.v[2] * AV.v[2];
    double m11 = V.v[0] * AV.v[0] + V.v[1] * AV.v[1] + V.v[2] * AV.v[2] - eval1;
    double absM00 = fabs(m00), absM01 = fabs(m01), absM11 = fabs(m11);
    V3 r;
    if (absM00 >= absM11) {
        double max_abs_comp = absM00 > absM01 ? absM00 : absM01;
        if (max_abs_comp > 0) {
            if (absM00 >= absM01) { m01 /= m00; m00 = 1 / sqrt(1 + m01 * m01); m01 *= m00; }
            else { m00 /= m01; m01 = 1 / sqrt(1 + m00 * m00); m00 *= m01; }
            for (int d = 0; d < 3; ++d) r.v[d] = m01 * U.v[d] - m00 * V.v[d];
            return r;
        }
        return U;
    } else {
        double max_abs_comp = absM11 > absM01 ? absM11 : absM01;
        if (max_abs_comp > 0) {
            if (absM11 >= absM01) { m01 /= m11; m11 = 1 / sqrt(1 + m01 * m01); m01 *= m11; }
            else { m11 /= m01; m01 = 1 / sqrt(1 + m11 * m11); m11 *= m01; }
            for (int d = 0; d < 3; ++d) r.v[d] = m11 * U.v[d] - m01 * V.v[d];
            return r;
        }
        return U;
    }
}

/* EstimateNormals.cpp FastEigen3x3 [ext, unverified]: eigenvector of the smallest
 * eigenvalue of the symmetric matrix given as cov6. */
static V3 fast_eigen3x3(const double* cov6) {
    double A[9] = {cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]};
    double max_coeff = A[0];
    for (int i = 1; i < 9; ++i) if (A[i] > max_coeff) max_coeff = A[i];
    V3 zero = {{0, 0, 0}};
    if (max_coeff == 0) return zero;
    for (int i = 0; i < 9; ++i) A[i] /= max_coeff;
    double norm = A_(0, 1) * A_(0, 1) + A_(0, 2) * A_(0, 2) + A_(1, 2) * A_(1, 2);
    if (norm > 0) {
        double q = (A_(0, 0) + A_(1, 1) + A_(2, 2)) / 3;
        double b00 = A_(0, 0) - q, b11 = A_(1, 1) - q, b22 = A_(2, 2) - q;
        double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2) / 6);
        double c00 = b11 * b22 - A_(1, 2) * A_(1, 2);
        double c01 = A_(0, 1) * b22 - A_(1, 2) * A_(0, 2);
        double c02 = A_(0, 1) * A_(1, 2) - b11 * A_(0, 2);
        double det = (b00 * c00 - A_(0, 1) * c01 + A_(0, 2) * c02) / (p * p * p);
        double half_det = det * 0.5;
        half_det = fmin(fmax(half_det, -1.0), 1.0);
        double angle = acos(half_det) / (double)3;
        double const two_thirds_pi = 2.09439510239319549;
        double beta2 = cos(angle) * 2;
        double beta0 = cos(angle + two_thirds_pi) * 2;
        double beta1 = -(beta0 + beta2);
        double eval0 = q + p * beta0, eval1 = q + p * beta1, eval2 = q + p * beta2;
        if (half_det >= 0) {
            V3 evec2 = eigvec0(A, eval2);
            if (eval2 < eval0 && eval2 < eval1) return evec2;
            V3 evec1 = eigvec1(A, evec2, eval1);
            if (eval1 < eval0 && eval1 < eval2) return evec1;
            return v3cross(evec1, evec2);
        } else {
            V3 evec0 = eigvec0(A, eval0);
            if (eval0 < eval1 && eval0 < eval2) return evec0;
            V3 evec1 = eigvec1(A, evec0, eval1);
            if (eval1 < eval0 && eval1 < eval2) return evec1;
            return v3cross(evec0, evec1);
        }
    } else {
        for (int i = 0; i < 9; ++i) A[i] *= max_coeff;
        V3 r = {{0, 0, 1}};
        if (A_(0, 0) < A_(1, 1) && A_(0, 0) < A_(2, 2)) { r.v[0] = 1; r.v[2] = 0; }
        else if (A_(1, 1) < A_(0, 0) && A_(1, 1) < A_(2, 2)) { r.v[1] = 1; r.v[2] = 0; }
        return r;
    }
}

/* PointCloud::EstimateNormals with KDTreeSearchParamHybrid(radius, max_nn), no prior
 * normals [ext, unverified].  normals_out (m,3). */
int orc_normals_hybrid(const double* pts, int64_t m, double radius, int max_nn, double* normals_out) {
    if (m <= 0) return 0;
    Grid g;
    if (grid_build(&g, pts, m)) return -1;
    const double r2 = radius * radius;
#pragma omp parallel
    {
        int32_t* bid = (int32_t*)malloc((size_t)max_nn * sizeof(int32_t));
        double* bd = (double*)malloc((size_t)max_nn * sizeof(double));
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < m; ++i) {
            int c = grid_knn(&g, pts, pts + 3 * i, max_nn, bid, bd);
            int nn = 0;
            while (nn < c && bd[nn] < r2) ++nn;          /* strict: lower_bound on r*r */
            double nx = 0, ny = 0, nz = 1;
            if (nn >= 3) {
                double cov6[6];
                covariance_of(pts, bid, nn, cov6);
                V3 n = fast_eigen3x3(cov6);
                double len = sqrt(n.v[0] * n.v[0] + n.v[1] * n.v[1] + n.v[2] * n.v[2]);
                if (len != 0.0) { nx = n.v[0]; ny = n.v[1]; nz = n.v[2]; }
            }
            normals_out[3 * i] = nx; normals_out[3 * i + 1] = ny; normals_out[3 * i + 2] = nz;
        }
        free(bid); free(bd);
    }
    grid_free(&g);
    return 0;
}

/* covariance of each point's kNN row (ids < 0 are padding); cov_out (m,6) */
void orc_cov_knn(const double* pts, int64_t m, const int32_t* knn, int k, double* cov_out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < m; ++i) {
        int n = 0;
        while (n < k && knn[i * k + n] >= 0) ++n;
        if (n == 0) { for (int q = 0; q < 6; ++q) cov_out[6 * i + q] = NAN; continue; }
        covariance_of(pts, knn + i * k, n, cov_out + 6 * i);
    }
}

/* covariance -> normal, exposed for unit tests of FastEigen3x3 */
void orc_fast_eigen(const double* cov6, int64_t n, double* out3) {
    for (int64_t i = 0; i < n; ++i) {
        V3 v = fast_eigen3x3(cov6 + 6 * i);
        out3[3 * i] = v.v[0]; out3[3 * i + 1] = v.v[1]; out3[3 * i + 2] = v.v[2];
    }
}

/* math_utils.vector_angle (math_utils.py:9-18) */
static double vector_angle(const double* u, const double* v) {
    double nu = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    double nv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double c = (u[0] / nu) * (v[0] / nv) + (u[1] / nu) * (v[1] / nv) + (u[2] / nu) * (v[2] / nv);
    double cl = fmin(1.0, fmax(c, -1.0));
    return acos(cl) * (180.0 / 3.14159265358979323846);
}

/* RegionGrowing._region_growing(method='knn') (region_growing.py:94-139): FIFO seed list,
 * `processed` set on acceptance only, neighbour_idx[1:k].  can_seed[p] = curvature(p) <
 * threshold_curve (precomputed by the caller with np.linalg.eig).  region_out (m) uint8.
 * near_out (optional, m): 1 where some tested angle was within `angle_tol` of the threshold. */
int64_t orc_region_grow(const int32_t* knn, int k, const double* normals, const uint8_t* seed,
                        const uint8_t* can_seed, int64_t m, double threshold_angle,
                        double angle_tol, uint8_t* region_out, uint8_t* near_out) {
    int32_t* list = (int32_t*)malloc((size_t)(m > 0 ? m : 1) * sizeof(int32_t));
    uint8_t* processed = (uint8_t*)calloc((size_t)(m > 0 ? m : 1), 1);
    if (!list || !processed) { free(list); free(processed); return -1; }
    int64_t n_list = 0;
    memset(region_out, 0, (size_t)m);
    if (near_out) memset(near_out, 0, (size_t)m);
    for (int64_t i = 0; i < m; ++i) if (seed[i]) { list[n_list++] = (int32_t)i; processed[i] = 1; region_out[i] = 1; }
    for (int64_t idx = 0; idx < n_list; ++idx) {
        int32_t s = list[idx];
        for (int j = 1; j < k; ++j) {
            int32_t nb = knn[(int64_t)s * k + j];
            if (nb < 0) break;
            if (processed[nb]) continue;
            double ang = vector_angle(normals + 3 * (int64_t)s, normals + 3 * (int64_t)nb);
            if (near_out && fabs(ang - threshold_angle) <= angle_tol) near_out[nb] = 1;
            if (ang < threshold_angle) {
                region_out[nb] = 1;
                processed[nb] = 1;
                if (can_seed[nb]) list[n_list++] = nb;
            }
        }
    }
    free(list); free(processed);
    return n_list;
}


/* Host threads of the OpenMP loops above (bench.py: torchrun exports OMP_NUM_THREADS=1, the CPU arm wants
 * all cores).  Returns the number of threads in force. */
int orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}
