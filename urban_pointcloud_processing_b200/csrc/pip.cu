// S2 -- BGT point-in-polygon (building / road footprints).
//
// Restates clip_utils.poly_clip + is_inside + _point_inside_poly
// (clip_utils.py:119-238) OR-ed over polygons (building_fuser.py:83-87,
// road_fuser.py:84-87): per ring an inclusive bounding-box prefilter, then the
// crossing-number test with "on boundary == inside"; interior rings clear what
// the exterior ring set.
//
// Acceleration structure ("plan"), built on the host in O(vertices + cells) and
// uploaded once per tile / polygon set:
//   * a uniform grid over the union of the ring bounding boxes; every cell lists
//     the rings whose bbox overlaps it (ring order preserved => exterior before
//     its holes, polygons in input order);
//   * per ring, uniform y-slabs over its bbox; every slab lists the edges whose
//     y-range touches it.  An edge that is not listed for the slab of a point
//     cannot satisfy `dy*dy2 <= 0` (clip_utils.py:139), so skipping it leaves the
//     crossing parity and the boundary test unchanged.  Lists are supersets (one
//     extra slab / cell on each side absorbs host/device rounding of the index
//     arithmetic); every listed edge goes through the exact reference test in fp64.
// The plan is a few MB at most and stays L2-resident; the point stream is read
// once, coalesced (16 B/pt in, 1 B/pt out).
//
// Buffered polygons (`offset` != 0: BGTPolyReader.filter_tile(offset=...), bgt_utils.py:154-163;
// `poly.buffer(0.05)` in AHNFuser._refine_layer, ahn_fuser.py:104-107).  The reference hands the
// polygon to GEOS, which is in neither container; the plan carries the offset instead and the
// query evaluates the buffer as what it is, a distance: a point is in buffer(P, d > 0) iff it is in
// P or within d of its boundary; in buffer(P, d < 0) iff it is in P and farther than |d| from the
// boundary.  That is the exact Minkowski sum where GEOS approximates the round joins at convex
// corners by 8 segments per quarter circle (quad_segs = 8: the arcs are cut by up to
// d (1 - cos(pi / 32)) = 0.0048 d, 1.2 mm for the notebook's d = 0.25 m); along straight edges
// the two agree to rounding.  [ext, unverified -- parity unpinned against GEOS, see DESIGN.md]
#include <math.h>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "upcp_math.cuh"

struct PipHeader {
    int32_t magic, n_rings, gx, gy;
    double x0, y0, inv_cw, inv_ch;
    double offset;           // buffer distance of every polygon of the plan (0 = the reference's exact test)
    int64_t off_ring_bbox;   // double[R][4]  xmin ymin xmax ymax
    int64_t off_ring_meta;   // int32 [R][4]  poly, is_hole, slab_base, n_slabs
    int64_t off_ring_slab;   // double[R][2]  slab y0, inv_h
    int64_t off_slab_start;  // int32 [total_slabs+1]
    int64_t off_slab_edges;  // int32 [n_slab_edges]  first-vertex index of edge
    int64_t off_verts;       // double[V][2]
    int64_t off_cell_start;  // int32 [gx*gy+1]
    int64_t off_cell_rings;  // int32 [n_cell_rings]
    int64_t total_bytes;
};

struct upcp_pip_plan {
    std::vector<char> blob;
};

namespace upcp {

constexpr int32_t kPipMagic = 0x50495031;  // "PIP1"
constexpr int kBlock = 256;

static inline int clampi(long long v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : (int)v); }

static inline int host_cell(double v, double origin, double inv, int n) {
    double t = (v - origin) * inv;
    if (!(t > 0.0)) return 0;
    if (t >= (double)n) return n - 1;
    return (int)floor(t);
}

__device__ __forceinline__ int dev_cell(double v, double origin, double inv, int n) {
    double t = (v - origin) * inv;
    if (!(t > 0.0)) return 0;
    if (t >= (double)n) return n - 1;
    return (int)floor(t);
}

// squared distance from p to the segment a-b (fp64, no FMA: the oracle evaluates the same expressions)
__device__ __forceinline__ double seg_dist2(double ax, double ay, double bx, double by, double px, double py) {
    const double ex = bx - ax, ey = by - ay;
    const double wx = px - ax, wy = py - ay;
    const double len2 = ex * ex + ey * ey;
    double t = len2 > 0.0 ? (wx * ex + wy * ey) / len2 : 0.0;
    t = fmin(fmax(t, 0.0), 1.0);
    const double dx = wx - t * ex, dy = wy - t * ey;
    return dx * dx + dy * dy;
}

// Buffered polygons: see the header comment.  Same plan layout; ring bounding boxes, cell lists and
// slabs were built with the buffer distance added, so one slab holds every edge that the crossing test
// or the distance test of a point may need.
__global__ void __launch_bounds__(kBlock)
pip_query_offset_kernel(const char* __restrict__ plan, const double* __restrict__ x,
                        const double* __restrict__ y, int64_t n, const uint8_t* __restrict__ sel,
                        uint8_t* __restrict__ out) {
    const PipHeader* h = (const PipHeader*)plan;
    const double2* ring_bbox = (const double2*)(plan + h->off_ring_bbox);
    const int4* ring_meta = (const int4*)(plan + h->off_ring_meta);
    const double2* ring_slab = (const double2*)(plan + h->off_ring_slab);
    const int32_t* slab_start = (const int32_t*)(plan + h->off_slab_start);
    const int32_t* slab_edges = (const int32_t*)(plan + h->off_slab_edges);
    const double2* verts = (const double2*)(plan + h->off_verts);
    const int32_t* cell_start = (const int32_t*)(plan + h->off_cell_start);
    const int32_t* cell_rings = (const int32_t*)(plan + h->off_cell_rings);
    const int gx = h->gx, gy = h->gy;
    const double x0 = h->x0, y0 = h->y0, inv_cw = h->inv_cw, inv_ch = h->inv_ch;
    const double d = h->offset, D = fabs(d), D2 = d * d;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);

    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        bool s = sel ? (sel[i] != 0) : true;
        bool inside_any = false;
        if (s) {
            double px = x[i], py = y[i];
            int cx = dev_cell(px, x0, inv_cw, gx);
            int cy = dev_cell(py, y0, inv_ch, gy);
            int c = cy * gx + cx;
            int k0 = __ldg(cell_start + c), k1 = __ldg(cell_start + c + 1);
            int cur_poly = -1;
            bool in_ext = false, in_hole = false;
            double dmin2 = inf;
            auto close_polygon = [&]() {
                const bool in = in_ext && !in_hole;
                return d > 0.0 ? (in || dmin2 <= D2) : (in && dmin2 > D2);
            };
            for (int k = k0; k < k1; ++k) {
                int r = __ldg(cell_rings + k);
                int4 meta = __ldg(ring_meta + r);
                if (meta.x != cur_poly) {
                    if (cur_poly >= 0) inside_any = inside_any || close_polygon();
                    if (inside_any) break;
                    in_ext = false; in_hole = false; dmin2 = inf;
                    cur_poly = meta.x;
                }
                double2 bmin = __ldg(ring_bbox + 2 * r), bmax = __ldg(ring_bbox + 2 * r + 1);
                if (!(px >= bmin.x - D && px <= bmax.x + D && py >= bmin.y - D && py <= bmax.y + D)) continue;
                double2 sl = __ldg(ring_slab + r);
                int si = dev_cell(py, sl.x, sl.y, meta.w);
                int e0 = __ldg(slab_start + meta.z + si), e1 = __ldg(slab_start + meta.z + si + 1);
                int parity = 0;
                bool boundary = false;
                for (int e = e0; e < e1; ++e) {
                    int vi = __ldg(slab_edges + e);
                    double2 a = __ldg(verts + vi), b = __ldg(verts + vi + 1);
                    int res = pip_edge(a.x, a.y, b.x, b.y, px, py);
                    if (res == 2) boundary = true; else parity ^= res;
                    dmin2 = fmin(dmin2, seg_dist2(a.x, a.y, b.x, b.y, px, py));
                }
                const bool in = boundary || (parity != 0);
                if (meta.y == 0) in_ext = in;
                else if (meta.y == 1) { if (in) in_hole = true; }
                else in_ext = in_ext != in;                    // even-odd loop: the polygon is the XOR of its loops
            }
            if (cur_poly >= 0 && !inside_any) inside_any = close_polygon();
        }
        out[i] = inside_any ? 1 : 0;
    }
}

__global__ void __launch_bounds__(kBlock)
pip_query_kernel(const char* __restrict__ plan, const double* __restrict__ x,
                 const double* __restrict__ y, int64_t n, const uint8_t* __restrict__ sel,
                 uint8_t* __restrict__ out) {
    const PipHeader* h = (const PipHeader*)plan;
    const double2* ring_bbox = (const double2*)(plan + h->off_ring_bbox);
    const int4* ring_meta = (const int4*)(plan + h->off_ring_meta);
    const double2* ring_slab = (const double2*)(plan + h->off_ring_slab);
    const int32_t* slab_start = (const int32_t*)(plan + h->off_slab_start);
    const int32_t* slab_edges = (const int32_t*)(plan + h->off_slab_edges);
    const double2* verts = (const double2*)(plan + h->off_verts);
    const int32_t* cell_start = (const int32_t*)(plan + h->off_cell_start);
    const int32_t* cell_rings = (const int32_t*)(plan + h->off_cell_rings);
    const int gx = h->gx, gy = h->gy;
    const double x0 = h->x0, y0 = h->y0, inv_cw = h->inv_cw, inv_ch = h->inv_ch;

    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        bool s = sel ? (sel[i] != 0) : true;
        bool inside_any = false;
        if (s) {
            double px = x[i], py = y[i];
            int cx = dev_cell(px, x0, inv_cw, gx);
            int cy = dev_cell(py, y0, inv_ch, gy);
            int c = cy * gx + cx;
            int k0 = __ldg(cell_start + c), k1 = __ldg(cell_start + c + 1);
            int cur_poly = -1;
            bool cur_in = false;
            for (int k = k0; k < k1; ++k) {
                int r = __ldg(cell_rings + k);
                int4 meta = __ldg(ring_meta + r);
                if (meta.x != cur_poly) {
                    inside_any = inside_any || cur_in;
                    if (inside_any) break;      // OR over polygons: nothing can undo it
                    cur_in = false;
                    cur_poly = meta.x;
                }
                bool is_hole = meta.y == 1;
                if (is_hole && !cur_in) continue;
                double2 bmin = __ldg(ring_bbox + 2 * r), bmax = __ldg(ring_bbox + 2 * r + 1);
                // rectangle_clip (clip_utils.py:38-39): inclusive on all four sides
                if (!(px >= bmin.x && px <= bmax.x && py >= bmin.y && py <= bmax.y)) continue;
                double2 sl = __ldg(ring_slab + r);
                int si = dev_cell(py, sl.x, sl.y, meta.w);
                int e0 = __ldg(slab_start + meta.z + si), e1 = __ldg(slab_start + meta.z + si + 1);
                int parity = 0;
                bool boundary = false;
                for (int e = e0; e < e1; ++e) {
                    int vi = __ldg(slab_edges + e);
                    double2 a = __ldg(verts + vi), b = __ldg(verts + vi + 1);
                    int res = pip_edge(a.x, a.y, b.x, b.y, px, py);
                    if (res == 2) { boundary = true; break; }
                    parity ^= res;
                }
                bool in = boundary || (parity != 0);
                if (meta.y == 0) cur_in = in;
                else if (is_hole) { if (in) cur_in = false; }
                else cur_in = cur_in != in;                    // even-odd loop
            }
            inside_any = inside_any || cur_in;
        }
        out[i] = inside_any ? 1 : 0;
    }
}

}  // namespace upcp

using namespace upcp;

extern "C" {

int upcp_pip_plan_create(const double* h_ring_xy, const int64_t* h_ring_off,
                         const int32_t* h_ring_poly, const uint8_t* h_ring_hole,
                         int32_t n_rings, int32_t grid_dim, upcp_pip_plan** out) {
    return upcp_pip_plan_create_buffered(h_ring_xy, h_ring_off, h_ring_poly, h_ring_hole, n_rings, grid_dim, 0.0, out);
}

int upcp_pip_plan_create_buffered(const double* h_ring_xy, const int64_t* h_ring_off,
                                  const int32_t* h_ring_poly, const uint8_t* h_ring_hole,
                                  int32_t n_rings, int32_t grid_dim, double offset, upcp_pip_plan** out) {
    if (!out) UPCP_FAIL(UPCP_E_INVALID, "pip_plan_create: out is NULL");
    *out = nullptr;
    if (!(offset == offset) || fabs(offset) > 1e12) UPCP_FAIL(UPCP_E_INVALID, "pip_plan_create: bad offset");
    const double D = fabs(offset);
    if (n_rings < 0 || (n_rings > 0 && (!h_ring_xy || !h_ring_off || !h_ring_poly || !h_ring_hole)))
        UPCP_FAIL(UPCP_E_INVALID, "pip_plan_create: NULL input");
    const int R = n_rings;
    const int64_t V = R > 0 ? h_ring_off[R] : 0;
    if (V > 0x7ffffff0LL) UPCP_FAIL(UPCP_E_CAPACITY, "pip_plan_create: too many vertices");
    for (int r = 0; r < R; ++r) {
        if (h_ring_off[r + 1] < h_ring_off[r]) UPCP_FAIL(UPCP_E_INVALID, "pip_plan_create: ring offsets not monotone");
        if (r > 0 && h_ring_poly[r] < h_ring_poly[r - 1]) UPCP_FAIL(UPCP_E_INVALID, "pip_plan_create: polygon ids must be non-decreasing");
    }

    // --- per-ring bbox (math_utils.compute_bounding_box over all ring coords)
    std::vector<double> bbox((size_t)R * 4);
    std::vector<char> active(R, 0);
    double ux0 = INFINITY, uy0 = INFINITY, ux1 = -INFINITY, uy1 = -INFINITY;
    {
        int prev_poly = -1;
        bool poly_alive = false;
        for (int r = 0; r < R; ++r) {
            int64_t a = h_ring_off[r], b = h_ring_off[r + 1];
            int nv = (int)(b - a);
            const int kind = h_ring_hole[r];                  // 0 exterior, 1 interior, 2 loop of an even-odd polygon
            if (kind > 2) UPCP_FAIL(UPCP_E_INVALID, "pip_plan_create: ring kind must be 0, 1 or 2");
            bool hole = kind == 1;
            if (h_ring_poly[r] != prev_poly) {
                prev_poly = h_ring_poly[r];
                // clip_utils.py:214-216: exterior with < 3 coords => polygon clips nothing
                poly_alive = kind == 2 || (!hole && nv >= 3);
                if (hole) UPCP_FAIL(UPCP_E_INVALID, "pip_plan_create: first ring of a polygon must be its exterior");
            } else if (kind == 0) {
                UPCP_FAIL(UPCP_E_INVALID, "pip_plan_create: a polygon has one exterior ring");
            }
            double xmin = INFINITY, ymin = INFINITY, xmax = -INFINITY, ymax = -INFINITY;
            for (int64_t v = a; v < b; ++v) {
                double vx = h_ring_xy[2 * v], vy = h_ring_xy[2 * v + 1];
                xmin = fmin(xmin, vx); xmax = fmax(xmax, vx);
                ymin = fmin(ymin, vy); ymax = fmax(ymax, vy);
            }
            bbox[4 * r + 0] = xmin; bbox[4 * r + 1] = ymin; bbox[4 * r + 2] = xmax; bbox[4 * r + 3] = ymax;
            // clip_utils.py:229-231: interiors with < 3 coords are skipped
            active[r] = poly_alive && nv >= 3 && isfinite(xmin) && isfinite(ymin) && isfinite(xmax) && isfinite(ymax);
            if (active[r]) {
                ux0 = fmin(ux0, xmin - D); uy0 = fmin(uy0, ymin - D); ux1 = fmax(ux1, xmax + D); uy1 = fmax(uy1, ymax + D);
            }
        }
    }
    bool any_active = isfinite(ux0);
    if (!any_active) { ux0 = uy0 = 0.0; ux1 = uy1 = 1.0; }

    // --- grid over the union bbox
    int G = grid_dim;
    if (G <= 0) {
        G = (int)ceil(4.0 * sqrt((double)(R > 0 ? R : 1)));
        if (G < 32) G = 32;
        if (G > 1024) G = 1024;
    }
    const int gx = G, gy = G;
    double w = ux1 - ux0, hgt = uy1 - uy0;
    double inv_cw = w > 0 ? (double)gx / w : 0.0;
    double inv_ch = hgt > 0 ? (double)gy / hgt : 0.0;

    std::vector<int32_t> cell_start((size_t)gx * gy + 1, 0);
    auto ring_cells = [&](int r, int& cx0, int& cx1, int& cy0, int& cy1) {
        cx0 = clampi((long long)host_cell(bbox[4 * r + 0] - D, ux0, inv_cw, gx) - 1, 0, gx - 1);
        cx1 = clampi((long long)host_cell(bbox[4 * r + 2] + D, ux0, inv_cw, gx) + 1, 0, gx - 1);
        cy0 = clampi((long long)host_cell(bbox[4 * r + 1] - D, uy0, inv_ch, gy) - 1, 0, gy - 1);
        cy1 = clampi((long long)host_cell(bbox[4 * r + 3] + D, uy0, inv_ch, gy) + 1, 0, gy - 1);
    };
    for (int r = 0; r < R; ++r) {
        if (!active[r]) continue;
        int cx0, cx1, cy0, cy1;
        ring_cells(r, cx0, cx1, cy0, cy1);
        for (int cy = cy0; cy <= cy1; ++cy)
            for (int cx = cx0; cx <= cx1; ++cx) cell_start[(size_t)cy * gx + cx + 1]++;
    }
    for (size_t c = 0; c < (size_t)gx * gy; ++c) {
        if ((int64_t)cell_start[c] + cell_start[c + 1] > 0x7fffffffLL)
            UPCP_FAIL(UPCP_E_CAPACITY, "pip_plan_create: cell ring lists overflow int32");
        cell_start[c + 1] += cell_start[c];
    }
    std::vector<int32_t> cell_rings((size_t)cell_start[(size_t)gx * gy]);
    {
        std::vector<int32_t> cursor(cell_start.begin(), cell_start.end() - 1);
        for (int r = 0; r < R; ++r) {   // ring order => lists sorted by ring index
            if (!active[r]) continue;
            int cx0, cx1, cy0, cy1;
            ring_cells(r, cx0, cx1, cy0, cy1);
            for (int cy = cy0; cy <= cy1; ++cy)
                for (int cx = cx0; cx <= cx1; ++cx) cell_rings[cursor[(size_t)cy * gx + cx]++] = r;
        }
    }

    // --- per-ring y-slabs
    std::vector<int32_t> meta((size_t)R * 4, 0);
    std::vector<double> slab_geo((size_t)R * 2, 0.0);
    int64_t total_slabs = 0;
    for (int r = 0; r < R; ++r) {
        int nv = (int)(h_ring_off[r + 1] - h_ring_off[r]);
        int ns = active[r] ? (nv < 1 ? 1 : (nv > 4096 ? 4096 : nv)) : 1;
        double ymin = bbox[4 * r + 1] - D, ymax = bbox[4 * r + 3] + D;      // slabs cover the buffered extent
        double hh = active[r] ? (ymax - ymin) : 0.0;
        meta[4 * r + 0] = h_ring_poly[r];
        meta[4 * r + 1] = h_ring_hole[r];
        meta[4 * r + 2] = (int32_t)total_slabs;
        meta[4 * r + 3] = ns;
        slab_geo[2 * r + 0] = active[r] ? ymin : 0.0;
        slab_geo[2 * r + 1] = hh > 0 ? (double)ns / hh : 0.0;
        total_slabs += ns;
    }
    if (total_slabs > 0x7ffffff0LL) UPCP_FAIL(UPCP_E_CAPACITY, "pip_plan_create: too many slabs");
    std::vector<int32_t> slab_start((size_t)total_slabs + 1, 0);
    auto edge_slabs = [&](int r, int64_t v, int& s0, int& s1) {
        double ya = h_ring_xy[2 * v + 1], yb = h_ring_xy[2 * (v + 1) + 1];
        double lo = fmin(ya, yb) - D, hi = fmax(ya, yb) + D;               // every slab a point within D of the edge can lie in
        int ns = meta[4 * r + 3];
        s0 = clampi((long long)host_cell(lo, slab_geo[2 * r], slab_geo[2 * r + 1], ns) - 1, 0, ns - 1);
        s1 = clampi((long long)host_cell(hi, slab_geo[2 * r], slab_geo[2 * r + 1], ns) + 1, 0, ns - 1);
        if (!(lo == lo) || !(hi == hi)) { s0 = 0; s1 = ns - 1; }  // NaN vertex: list everywhere
    };
    for (int r = 0; r < R; ++r) {
        if (!active[r]) continue;
        int base = meta[4 * r + 2];
        for (int64_t v = h_ring_off[r]; v + 1 < h_ring_off[r + 1]; ++v) {
            int s0, s1;
            edge_slabs(r, v, s0, s1);
            for (int s = s0; s <= s1; ++s) slab_start[(size_t)base + s + 1]++;
        }
    }
    for (int64_t s = 0; s < total_slabs; ++s) {
        if ((int64_t)slab_start[s] + slab_start[s + 1] > 0x7fffffffLL)
            UPCP_FAIL(UPCP_E_CAPACITY, "pip_plan_create: slab edge lists overflow int32");
        slab_start[s + 1] += slab_start[s];
    }
    std::vector<int32_t> slab_edges((size_t)slab_start[(size_t)total_slabs]);
    {
        std::vector<int32_t> cursor(slab_start.begin(), slab_start.end() - 1);
        for (int r = 0; r < R; ++r) {
            if (!active[r]) continue;
            int base = meta[4 * r + 2];
            for (int64_t v = h_ring_off[r]; v + 1 < h_ring_off[r + 1]; ++v) {
                int s0, s1;
                edge_slabs(r, v, s0, s1);
                for (int s = s0; s <= s1; ++s) slab_edges[cursor[(size_t)base + s]++] = (int32_t)v;
            }
        }
    }

    // --- serialise
    PipHeader hd;
    memset(&hd, 0, sizeof(hd));
    hd.magic = kPipMagic; hd.n_rings = R; hd.gx = gx; hd.gy = gy;
    hd.x0 = ux0; hd.y0 = uy0; hd.inv_cw = inv_cw; hd.inv_ch = inv_ch; hd.offset = offset;
    size_t off = align_up(sizeof(PipHeader), 256);
    auto place = [&](size_t bytes) { size_t o = off; off = align_up(off + (bytes ? bytes : 1), 256); return (int64_t)o; };
    hd.off_ring_bbox = place((size_t)R * 4 * sizeof(double));
    hd.off_ring_meta = place((size_t)R * 4 * sizeof(int32_t));
    hd.off_ring_slab = place((size_t)R * 2 * sizeof(double));
    hd.off_slab_start = place(slab_start.size() * sizeof(int32_t));
    hd.off_slab_edges = place(slab_edges.size() * sizeof(int32_t));
    hd.off_verts = place((size_t)V * 2 * sizeof(double));
    hd.off_cell_start = place(cell_start.size() * sizeof(int32_t));
    hd.off_cell_rings = place(cell_rings.size() * sizeof(int32_t));
    hd.total_bytes = (int64_t)off;

    upcp_pip_plan* plan = new (std::nothrow) upcp_pip_plan();
    if (!plan) UPCP_FAIL(UPCP_E_INTERNAL, "pip_plan_create: out of host memory");
    plan->blob.assign(off, 0);
    char* p = plan->blob.data();
    memcpy(p, &hd, sizeof(hd));
    if (R) {
        memcpy(p + hd.off_ring_bbox, bbox.data(), bbox.size() * sizeof(double));
        memcpy(p + hd.off_ring_meta, meta.data(), meta.size() * sizeof(int32_t));
        memcpy(p + hd.off_ring_slab, slab_geo.data(), slab_geo.size() * sizeof(double));
    }
    memcpy(p + hd.off_slab_start, slab_start.data(), slab_start.size() * sizeof(int32_t));
    if (!slab_edges.empty()) memcpy(p + hd.off_slab_edges, slab_edges.data(), slab_edges.size() * sizeof(int32_t));
    if (V) memcpy(p + hd.off_verts, h_ring_xy, (size_t)V * 2 * sizeof(double));
    memcpy(p + hd.off_cell_start, cell_start.data(), cell_start.size() * sizeof(int32_t));
    if (!cell_rings.empty()) memcpy(p + hd.off_cell_rings, cell_rings.data(), cell_rings.size() * sizeof(int32_t));
    *out = plan;
    return UPCP_OK;
}

size_t upcp_pip_plan_bytes(const upcp_pip_plan* plan) { return plan ? plan->blob.size() : 0; }

int upcp_pip_plan_upload(const upcp_pip_plan* plan, void* d_plan, size_t bytes, void* stream) {
    if (!plan || !d_plan) UPCP_FAIL(UPCP_E_INVALID, "pip_plan_upload: NULL argument");
    if (bytes < plan->blob.size()) UPCP_FAIL(UPCP_E_WORKSPACE, "pip_plan_upload: device buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    UPCP_CUDA_OK(cudaMemcpyAsync(d_plan, plan->blob.data(), plan->blob.size(), cudaMemcpyHostToDevice, st));
    // the blob is pageable host memory owned by the plan: make the copy complete
    // before the caller may destroy the plan
    UPCP_CUDA_OK(cudaStreamSynchronize(st));
    return UPCP_OK;
}

void upcp_pip_plan_destroy(upcp_pip_plan* plan) { delete plan; }

int upcp_pip_query(const void* d_plan, const double* d_x, const double* d_y, int64_t n,
                   const uint8_t* d_sel, uint8_t* d_out, void* stream) {
    if (n < 0 || !d_plan) UPCP_FAIL(UPCP_E_INVALID, "pip_query: bad argument");
    if (n == 0) return UPCP_OK;
    prof_begin(UPCP_PROF_PIP, (cudaStream_t)stream);
    pip_query_kernel<<<grid_for(n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        (const char*)d_plan, d_x, d_y, n, d_sel, d_out);
    prof_end(UPCP_PROF_PIP, (cudaStream_t)stream, n);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

int upcp_pip_query_buffered(const void* d_plan, const double* d_x, const double* d_y, int64_t n,
                            const uint8_t* d_sel, uint8_t* d_out, void* stream) {
    if (n < 0 || !d_plan) UPCP_FAIL(UPCP_E_INVALID, "pip_query_buffered: bad argument");
    if (n == 0) return UPCP_OK;
    pip_query_offset_kernel<<<grid_for(n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        (const char*)d_plan, d_x, d_y, n, d_sel, d_out);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

}  // extern "C"
