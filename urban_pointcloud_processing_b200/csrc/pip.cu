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
//   * a fine TAG grid (<= 512 x 512 one-byte cells over the same bbox): 0 = no point of the cell is inside
//     the set, 1 = every point is, 2 = an edge passes through or next to the cell.  The crossing rule of
//     the reference is the consistent half-open one (an edge is counted iff its other end lies strictly
//     above the ray, clip_utils.py:139-160), so the result is constant over any cell no edge comes near;
//     such a cell is classified ONCE with the exact test at its centre (three small kernels right after
//     the upload) and its points need one byte lookup instead of the ring / slab / edge walk.  Only the
//     thin band of cells along the edges (inflated by a cell: rounding of the index arithmetic, points
//     within rounding of an edge) runs the exact test.
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
    int32_t gt, n_verts;     // tag grid size (gt x gt over the same bbox), 0 = none; number of vertices
    double inv_tw, inv_th;
    int64_t off_tags;        // uint8 [gt*gt]: 0 outside, 1 inside, 2 mixed (filled on the device after the upload)
    int64_t off_edge_ok;     // uint8 [n_verts]: vertex v -> v + 1 is an edge of an active ring
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
    std::vector<char> blob;        // header + lists + edge flags (host side, uploaded)
    size_t device_bytes = 0;       // blob + the tag grid the device fills in after the upload
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

#if defined(__CUDA_ARCH__)
#define PIP_LD(p) __ldg(p)
#else
#define PIP_LD(p) (*(p))
#endif

__host__ __device__ __forceinline__ int pip_cell(double v, double origin, double inv, int n) {
    double t = (v - origin) * inv;
    if (!(t > 0.0)) return 0;
    if (t >= (double)n) return n - 1;
    return (int)floor(t);
}

// The exact test of one point against the polygon set of a plan (reference semantics: inclusive bounding-box
// prefilter, crossing number with boundary == inside, interiors clear what the exterior set, OR over
// polygons).  Device: the query kernel; host: classification of the uniform cells of the tag grid.
__host__ __device__ __forceinline__ bool pip_exact(const char* __restrict__ plan, double px, double py) {
    const PipHeader* h = (const PipHeader*)plan;
    const double2* ring_bbox = (const double2*)(plan + h->off_ring_bbox);
    const int4* ring_meta = (const int4*)(plan + h->off_ring_meta);
    const double2* ring_slab = (const double2*)(plan + h->off_ring_slab);
    const int32_t* slab_start = (const int32_t*)(plan + h->off_slab_start);
    const int32_t* slab_edges = (const int32_t*)(plan + h->off_slab_edges);
    const double2* verts = (const double2*)(plan + h->off_verts);
    const int32_t* cell_start = (const int32_t*)(plan + h->off_cell_start);
    const int32_t* cell_rings = (const int32_t*)(plan + h->off_cell_rings);
    const int cx = pip_cell(px, h->x0, h->inv_cw, h->gx);
    const int cy = pip_cell(py, h->y0, h->inv_ch, h->gy);
    const int c = cy * h->gx + cx;
    const int k0 = PIP_LD(cell_start + c), k1 = PIP_LD(cell_start + c + 1);
    bool inside_any = false;
    int cur_poly = -1;
    bool cur_in = false;
    for (int k = k0; k < k1; ++k) {
        const int r = PIP_LD(cell_rings + k);
        const int4 meta = PIP_LD(ring_meta + r);
        if (meta.x != cur_poly) {
            inside_any = inside_any || cur_in;
            if (inside_any) break;      // OR over polygons: nothing can undo it
            cur_in = false;
            cur_poly = meta.x;
        }
        const bool is_hole = meta.y == 1;
        if (is_hole && !cur_in) continue;
        const double2 bmin = PIP_LD(ring_bbox + 2 * r), bmax = PIP_LD(ring_bbox + 2 * r + 1);
        // rectangle_clip (clip_utils.py:38-39): inclusive on all four sides
        if (!(px >= bmin.x && px <= bmax.x && py >= bmin.y && py <= bmax.y)) continue;
        const double2 sl = PIP_LD(ring_slab + r);
        const int si = pip_cell(py, sl.x, sl.y, meta.w);
        const int e0 = PIP_LD(slab_start + meta.z + si), e1 = PIP_LD(slab_start + meta.z + si + 1);
        int parity = 0;
        bool boundary = false;
        for (int e = e0; e < e1; ++e) {
            const int vi = PIP_LD(slab_edges + e);
            const double2 a = PIP_LD(verts + vi), b = PIP_LD(verts + vi + 1);
            const int res = pip_edge(a.x, a.y, b.x, b.y, px, py);
            if (res == 2) { boundary = true; break; }
            parity ^= res;
        }
        const bool in = boundary || (parity != 0);
        if (meta.y == 0) cur_in = in;
        else if (is_hole) { if (in) cur_in = false; }
        else cur_in = cur_in != in;                    // even-odd loop
    }
    return inside_any || cur_in;
}

// ---- tag grid, built on the device from the uploaded plan (three tiny launches) ----
__global__ void __launch_bounds__(kBlock)
pip_tag_init_kernel(char* __restrict__ plan) {
    const PipHeader* h = (const PipHeader*)plan;
    const int gt = h->gt;
    uint8_t* tags = (uint8_t*)(plan + h->off_tags);
    for (int c = blockIdx.x * kBlock + threadIdx.x; c < gt * gt; c += gridDim.x * kBlock) {
        const int cx = c % gt, cy = c / gt;
        // the border cells also receive the points outside the grid (clamped index): always exact
        tags[c] = (cx == 0 || cy == 0 || cx == gt - 1 || cy == gt - 1) ? 2 : 0;
    }
}
// every edge marks the cells it passes through or next to: pieces of at most ~2 cells, the cells of every
// piece's bounding box plus one more cell on every side
__global__ void __launch_bounds__(kBlock)
pip_tag_edges_kernel(char* __restrict__ plan) {
    const PipHeader* h = (const PipHeader*)plan;
    const int gt = h->gt;
    uint8_t* tags = (uint8_t*)(plan + h->off_tags);
    const uint8_t* edge_ok = (const uint8_t*)(plan + h->off_edge_ok);
    const double2* verts = (const double2*)(plan + h->off_verts);
    for (int v = blockIdx.x * kBlock + threadIdx.x; v + 1 < h->n_verts; v += gridDim.x * kBlock) {
        if (!edge_ok[v]) continue;
        const double2 a = verts[v], b = verts[v + 1];
        const double len_cells = fmax(fabs(b.x - a.x) * h->inv_tw, fabs(b.y - a.y) * h->inv_th);
        const int pieces = len_cells > 2.0 ? (int)fmin(ceil(len_cells / 2.0), 1e6) : 1;
        for (int k = 0; k < pieces; ++k) {
            const double t0 = (double)k / pieces, t1 = (double)(k + 1) / pieces;
            const double x_0 = a.x + t0 * (b.x - a.x), x_1 = a.x + t1 * (b.x - a.x);
            const double y_0 = a.y + t0 * (b.y - a.y), y_1 = a.y + t1 * (b.y - a.y);
            const int cx0 = max(pip_cell(fmin(x_0, x_1), h->x0, h->inv_tw, gt) - 1, 0), cx1 = min(pip_cell(fmax(x_0, x_1), h->x0, h->inv_tw, gt) + 1, gt - 1);
            const int cy0 = max(pip_cell(fmin(y_0, y_1), h->y0, h->inv_th, gt) - 1, 0), cy1 = min(pip_cell(fmax(y_0, y_1), h->y0, h->inv_th, gt) + 1, gt - 1);
            for (int cy = cy0; cy <= cy1; ++cy)
                for (int cx = cx0; cx <= cx1; ++cx) tags[cy * gt + cx] = 2;
        }
    }
}
// every cell no edge comes near: ONE exact evaluation at its centre decides all its points
__global__ void __launch_bounds__(kBlock)
pip_tag_classify_kernel(char* __restrict__ plan) {
    const PipHeader* h = (const PipHeader*)plan;
    const int gt = h->gt;
    uint8_t* tags = (uint8_t*)(plan + h->off_tags);
    for (int c = blockIdx.x * kBlock + threadIdx.x; c < gt * gt; c += gridDim.x * kBlock) {
        if (tags[c] == 2) continue;
        const double px = h->x0 + ((double)(c % gt) + 0.5) / h->inv_tw, py = h->y0 + ((double)(c / gt) + 0.5) / h->inv_th;
        tags[c] = pip_exact(plan, px, py) ? 1 : 0;
    }
}

// Query kernel.  The points of a tile arrive in no spatial order, so nearly every warp holds SOME point of
// the thin band of mixed cells; a per-point "tag, else exact" would make every warp walk the exact path
// (a chain of ~8 dependent loads) with one or two lanes.  A CTA therefore takes 4 x kBlock consecutive
// points per round: every thread loads four points with all loads independent (selection byte and
// coordinates together, then the four tag bytes), writes the result of the decided ones and appends the
// undecided ones -- coordinates and index -- to a CTA-local list in shared memory.  The list is walked by
// the exact test only when it holds at least kPipFlush points (and once at the end), so that every lane of
// every warp of the CTA has a point while the other CTAs of the SM keep streaming.
constexpr int kPipPerThread = 4;
constexpr int kPipTile = kPipPerThread * kBlock;
constexpr int kPipFlush = 512;
constexpr int kPipCap = kPipFlush + kPipTile;

__global__ void __launch_bounds__(kBlock, 4)
pip_query_kernel(const char* __restrict__ plan, const double* __restrict__ x,
                 const double* __restrict__ y, int64_t n, const uint8_t* __restrict__ sel,
                 uint8_t* __restrict__ out) {
    __shared__ double s_x[kPipCap], s_y[kPipCap];
    __shared__ uint32_t s_i[kPipCap];
    __shared__ int s_cnt;
    const PipHeader* h = (const PipHeader*)plan;
    const int gt = h->gt;
    const uint8_t* tags = (const uint8_t*)(plan + h->off_tags);
    const double x0 = h->x0, y0 = h->y0, inv_tw = h->inv_tw, inv_th = h->inv_th;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int64_t base = (int64_t)blockIdx.x * kPipTile; ; base += (int64_t)gridDim.x * kPipTile) {
        const bool last = base >= n;        // uniform over the CTA: one more pass to drain the list
        if (!last) {
            double px[kPipPerThread], py[kPipPerThread];
            bool s[kPipPerThread];
            int tag[kPipPerThread];
#pragma unroll
            for (int u = 0; u < kPipPerThread; ++u) {
                const int64_t i = base + u * kBlock + threadIdx.x;
                const bool in = i < n;
                s[u] = in && (sel ? (sel[i] != 0) : true);
                px[u] = in ? x[i] : 0.0;
                py[u] = in ? y[i] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kPipPerThread; ++u) {
                tag[u] = 0;
                if (s[u]) tag[u] = gt > 0 ? (int)__ldg(tags + pip_cell(py[u], y0, inv_th, gt) * gt + pip_cell(px[u], x0, inv_tw, gt)) : 2;
            }
#pragma unroll
            for (int u = 0; u < kPipPerThread; ++u) {
                const int64_t i = base + u * kBlock + threadIdx.x;
                if (i < n) out[i] = tag[u] == 1 ? 1 : 0;
                const uint32_t bal = __ballot_sync(0xffffffffu, tag[u] == 2);
                if (bal) {
                    int at = 0;
                    const int leader = __ffs(bal) - 1;
                    if (lane == leader) at = atomicAdd(&s_cnt, __popc(bal));
                    at = __shfl_sync(0xffffffffu, at, leader) + __popc(bal & ((1u << lane) - 1u));
                    if (tag[u] == 2) { s_x[at] = px[u]; s_y[at] = py[u]; s_i[at] = (uint32_t)i; }
                }
            }
        }
        __syncthreads();
        const int m = s_cnt;
        __syncthreads();                    // nobody appends for the next round before everyone has read the count
        if (m >= kPipFlush || (last && m > 0)) {
            for (int k = threadIdx.x; k < m; k += kBlock) out[s_i[k]] = pip_exact(plan, s_x[k], s_y[k]) ? 1 : 0;
            __syncthreads();
            if (threadIdx.x == 0) s_cnt = 0;
            __syncthreads();
        }
        if (last) break;
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
    // tag grid (see the header comment): space and geometry here, the classification itself runs on the
    // device right after the upload (upcp_pip_plan_upload).  Not for buffered plans or NaN / infinite vertices.
    bool want_tags = offset == 0.0 && any_active;
    for (int r = 0; r < R && want_tags; ++r) {
        if (!active[r]) continue;
        for (int64_t v = h_ring_off[r]; v < h_ring_off[r + 1]; ++v)
            want_tags = want_tags && isfinite(h_ring_xy[2 * v]) && isfinite(h_ring_xy[2 * v + 1]);
    }
    size_t host_bytes = off;                 // what the host blob holds and the upload copies
    if (want_tags) {
        // resolution: the band of cells along the edges (about four cells wide) should stay a small part of the
        // area -- 512 x 512 for an ordinary BGT tile, finer where thousands of footprints leave no cell far
        // from an edge (the table stays L2-resident at up to 4 MB)
        double edge_len = 0.0;
        for (int r = 0; r < R; ++r) {
            if (!active[r]) continue;
            for (int64_t v = h_ring_off[r]; v + 1 < h_ring_off[r + 1]; ++v)
                edge_len += hypot(h_ring_xy[2 * (v + 1)] - h_ring_xy[2 * v], h_ring_xy[2 * (v + 1) + 1] - h_ring_xy[2 * v + 1]);
        }
        int gt = 512;
        const double area = w * hgt, side = fmax(w, hgt);
        while (gt < 2048 && area > 0.0 && edge_len * (4.0 * side / gt) > 0.25 * area) gt *= 2;
        hd.gt = gt; hd.n_verts = (int32_t)V;
        hd.inv_tw = w > 0 ? (double)gt / w : 0.0; hd.inv_th = hgt > 0 ? (double)gt / hgt : 0.0;
        hd.off_edge_ok = place((size_t)(V > 0 ? V : 1));
        host_bytes = off;
        hd.off_tags = place((size_t)gt * gt);             // device only: filled by the tag kernels after the upload
    }
    hd.total_bytes = (int64_t)off;

    upcp_pip_plan* plan = new (std::nothrow) upcp_pip_plan();
    if (!plan) UPCP_FAIL(UPCP_E_INTERNAL, "pip_plan_create: out of host memory");
    plan->blob.assign(host_bytes, 0);
    plan->device_bytes = off;
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
    if (want_tags) {
        uint8_t* edge_ok = (uint8_t*)(p + hd.off_edge_ok);               // 1: vertex v and v + 1 form an edge of an active ring
        for (int r = 0; r < R; ++r) {
            if (!active[r]) continue;
            for (int64_t v = h_ring_off[r]; v + 1 < h_ring_off[r + 1]; ++v) edge_ok[(size_t)v] = 1;
        }
    }
    *out = plan;
    return UPCP_OK;
}

size_t upcp_pip_plan_bytes(const upcp_pip_plan* plan) { return plan ? plan->device_bytes : 0; }

int upcp_pip_plan_upload(const upcp_pip_plan* plan, void* d_plan, size_t bytes, void* stream) {
    if (!plan || !d_plan) UPCP_FAIL(UPCP_E_INVALID, "pip_plan_upload: NULL argument");
    if (bytes < plan->device_bytes) UPCP_FAIL(UPCP_E_WORKSPACE, "pip_plan_upload: device buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    UPCP_CUDA_OK(cudaMemcpyAsync(d_plan, plan->blob.data(), plan->blob.size(), cudaMemcpyHostToDevice, st));
    const PipHeader* hp = (const PipHeader*)plan->blob.data();
    if (hp->gt > 0) {
        const int cells = hp->gt * hp->gt;
        pip_tag_init_kernel<<<grid_for(cells, kBlock), kBlock, 0, st>>>((char*)d_plan);
        UPCP_LAUNCH_OK();
        if (hp->n_verts > 1) {
            pip_tag_edges_kernel<<<grid_for(hp->n_verts, kBlock), kBlock, 0, st>>>((char*)d_plan);
            UPCP_LAUNCH_OK();
        }
        pip_tag_classify_kernel<<<grid_for(cells, kBlock), kBlock, 0, st>>>((char*)d_plan);
        UPCP_LAUNCH_OK();
    }
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
    if (n >= ((int64_t)1 << 32)) UPCP_FAIL(UPCP_E_CAPACITY, "pip_query: n must be < 2^32");
    prof_begin(UPCP_PROF_PIP, (cudaStream_t)stream);
    const int64_t tiles = (n + kPipTile - 1) / kPipTile;
    const int64_t resident = (int64_t)num_sms() * 4;
    pip_query_kernel<<<(unsigned)(tiles < resident ? tiles : resident), kBlock, 0, (cudaStream_t)stream>>>(
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
