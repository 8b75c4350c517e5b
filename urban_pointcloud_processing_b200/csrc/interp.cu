// "Next" row 4 -- AHN preprocessing: SpatialInterpolator (inverse-distance weighting / maximum over
// the k nearest AHN points within max_dist) evaluated at the cells of the 0.1 m raster the labelling path later
// reads (reference utils/interpolation.py:156-308, called from preprocessing/ahn_preprocessing.py:129-236).
//
// The reference asks a scipy cKDTree for the n_neighbors nearest points with distance < max_dist of every
// position and then runs a Python loop over the positions.  Here: the data points are counting-sorted into
// a 2-D uniform grid with cells of max_dist (so the 3 x 3 block around a position holds every point within
// max_dist), one thread per position keeps its k nearest in a small sorted list and folds them exactly as
// the loop body does: dk = sqrt(d2), the confusion test dk <= conf_dist, w = 1 / (dk ** power + reg)
// (power == 2 is dk * dk, as numpy squares), wtot = sum(w), value = dot(w, values) / wtot; `max` takes the
// maximum of the k values.  Distances are ((dx*dx) + (dy*dy)) in fp64 without FMA, as cKDTree accumulates
// them; ties at equal distance are ordered by point index (cKDTree: heap order) [ext, unverified: only
// the summation order of tied weights can differ].
#include <math.h>
#include "common.cuh"
#include "primitives.cuh"

namespace upcp {

constexpr int kIBlock = 256;
constexpr int kMaxNb = 32;

struct IGrid { double ox, oy, inv_cell; int nx, ny; };

__device__ __forceinline__ int icell(double v, double o, double inv, int n) {
    double t = (v - o) * inv;
    if (!(t > 0.0)) return 0;
    if (t >= (double)n) return n - 1;
    return (int)floor(t);
}

__global__ void __launch_bounds__(kIBlock)
interp_keys_kernel(const double* __restrict__ px, const double* __restrict__ py, int64_t n, IGrid g,
                   uint32_t* __restrict__ keys, uint32_t* __restrict__ rank, uint32_t* __restrict__ cell_count) {
    int64_t stride = (int64_t)gridDim.x * kIBlock;
    for (int64_t i = (int64_t)blockIdx.x * kIBlock + threadIdx.x; i < n; i += stride) {
        const uint32_t key = (uint32_t)icell(py[i], g.oy, g.inv_cell, g.ny) * (uint32_t)g.nx + (uint32_t)icell(px[i], g.ox, g.inv_cell, g.nx);
        keys[i] = key;
        rank[i] = atomicAdd(cell_count + key, 1u);
    }
}

__global__ void __launch_bounds__(kIBlock)
interp_place_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rank,
                    const uint32_t* __restrict__ cell_first, int64_t n, uint32_t* __restrict__ sorted_idx) {
    int64_t stride = (int64_t)gridDim.x * kIBlock;
    for (int64_t i = (int64_t)blockIdx.x * kIBlock + threadIdx.x; i < n; i += stride)
        sorted_idx[cell_first[keys[i]] + rank[i]] = (uint32_t)i;
}

struct InterpArgs {
    const double* px; const double* py; const double* pv;     // data points and their values
    const uint32_t* cell_first; const uint32_t* sorted_idx;
    IGrid g;
    const double* qx; const double* qy; int64_t nq;            // positions
    int k, method;                                             // 0 idw, 1 max
    double max_dist, power, reg, conf_dist, fill;
    double* out;
};

__global__ void __launch_bounds__(kIBlock)
interp_query_kernel(InterpArgs a) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    int64_t stride = (int64_t)gridDim.x * kIBlock;
    for (int64_t q = (int64_t)blockIdx.x * kIBlock + threadIdx.x; q < a.nq; q += stride) {
        const double x = a.qx[q], y = a.qy[q];
        double bd[kMaxNb]; uint32_t bi[kMaxNb];
        int cnt = 0;
        const int cx = icell(x, a.g.ox, a.g.inv_cell, a.g.nx), cy = icell(y, a.g.oy, a.g.inv_cell, a.g.ny);
        // positions outside the grid by more than a cell cannot have a neighbour within max_dist; the clamped
        // cell block is still a superset of what lies within max_dist of positions inside or next to the grid
        for (int yy = max(cy - 1, 0); yy <= min(cy + 1, a.g.ny - 1); ++yy) {
            const uint32_t row = (uint32_t)yy * (uint32_t)a.g.nx;
            const uint32_t p0 = a.cell_first[row + max(cx - 1, 0)], p1 = a.cell_first[row + min(cx + 1, a.g.nx - 1) + 1];
            for (uint32_t p = p0; p < p1; ++p) {
                const uint32_t i = a.sorted_idx[p];
                const double dx = a.px[i] - x, dy = a.py[i] - y;
                const double d = sqrt(dx * dx + dy * dy);
                if (!(d < a.max_dist)) continue;                               // cKDTree: distance_upper_bound is exclusive
                if (cnt == a.k && !(d < bd[cnt - 1] || (d == bd[cnt - 1] && i < bi[cnt - 1]))) continue;
                int j = cnt < a.k ? cnt++ : cnt - 1;                             // insertion, ascending (distance, index)
                while (j > 0 && (bd[j - 1] > d || (bd[j - 1] == d && bi[j - 1] > i))) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
                bd[j] = d; bi[j] = i;
            }
        }
        double res = a.fill;
        if (cnt > 0) {
            if (a.method == 1) {
                double m = -inf;
                for (int j = 0; j < cnt; ++j) m = fmax(m, a.pv[bi[j]]);
                res = m;
            } else {
                bool confused = false;
                if (a.conf_dist > 0.0)
                    for (int j = 0; j < cnt && !confused; ++j)
                        if (bd[j] <= a.conf_dist) { res = a.pv[bi[j]]; confused = true; }
                if (!confused) {
                    double wtot = 0.0, dot = 0.0;
                    for (int j = 0; j < cnt; ++j) {
                        const double pw = a.power == 2.0 ? bd[j] * bd[j] : (a.power == 1.0 ? bd[j] : pow(bd[j], a.power));
                        const double w = 1.0 / (pw + a.reg);
                        wtot += w;
                        dot += w * a.pv[bi[j]];
                    }
                    res = wtot > 0.0 ? dot / wtot : a.fill;
                }
            }
        }
        a.out[q] = res;
    }
}

}  // namespace upcp

using namespace upcp;

extern "C" {

static void interp_grid(const double* h_bbox4, double max_dist, int64_t n, IGrid* g) {
    double cell = max_dist;
    const double ex = h_bbox4[2] - h_bbox4[0], ey = h_bbox4[3] - h_bbox4[1];
    // at most ~4 cells per data point (and 2^24 in all): coarser cells only make the 3 x 3 block a larger superset
    const double cap = fmin(16777216.0, 4.0 * (double)(n < 1024 ? 1024 : n));
    while ((floor(ex / cell) + 1) * (floor(ey / cell) + 1) > cap) cell *= 1.25;
    g->ox = h_bbox4[0]; g->oy = h_bbox4[1]; g->inv_cell = 1.0 / cell;
    g->nx = (int)floor(ex / cell) + 1; g->ny = (int)floor(ey / cell) + 1;
}

size_t upcp_interpolate_ws_bytes(int64_t n) {
    if (n < 1) n = 1;
    const size_t cells = (size_t)fmin(16777216.0, 4.0 * (double)(n < 1024 ? 1024 : n)) + 2;
    return 3 * align_up((size_t)n * 4) + align_up(cells * 4) + align_up(scan_ws_words((int64_t)cells) * 4) + 4096;
}

int upcp_interpolate(const double* d_px, const double* d_py, const double* d_pv, int64_t n, const double* h_bbox4,
                     const double* d_qx, const double* d_qy, int64_t nq, int method, int n_neighbors,
                     double max_dist, double power, double reg, double conf_dist, double fill_value,
                     double* d_out, void* d_ws, size_t ws_bytes, void* stream) {
    if (n < 1 || nq < 0 || !h_bbox4 || n_neighbors < 1 || (method != 0 && method != 1))
        UPCP_FAIL(UPCP_E_INVALID, "interpolate: bad argument");
    if (n_neighbors > kMaxNb) UPCP_FAIL(UPCP_E_CAPACITY, "interpolate: n_neighbors is limited to 32");
    if (!(max_dist > 0.0) || !isfinite(max_dist)) UPCP_FAIL(UPCP_E_INVALID, "interpolate: a finite max_dist > 0 is required");
    if (n >= ((int64_t)1 << 31)) UPCP_FAIL(UPCP_E_CAPACITY, "interpolate: n must be < 2^31");
    if (nq == 0) return UPCP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    IGrid g;
    interp_grid(h_bbox4, max_dist, n, &g);
    const size_t cells = (size_t)g.nx * (size_t)g.ny;
    Workspace ws(d_ws, ws_bytes);
    uint32_t* keys = ws.take<uint32_t>((size_t)n);
    uint32_t* rank = ws.take<uint32_t>((size_t)n);
    uint32_t* sorted_idx = ws.take<uint32_t>((size_t)n);
    uint32_t* cell_first = ws.take<uint32_t>(cells + 1);
    uint32_t* scan_ws = ws.take<uint32_t>(scan_ws_words((int64_t)cells + 1));
    if (!ws.ok) UPCP_FAIL(UPCP_E_WORKSPACE, "interpolate: workspace too small");
    UPCP_CUDA_OK(cudaMemsetAsync(cell_first, 0, (cells + 1) * 4, st));
    interp_keys_kernel<<<grid_for(n, kIBlock), kIBlock, 0, st>>>(d_px, d_py, n, g, keys, rank, cell_first);
    UPCP_LAUNCH_OK();
    UPCP_TRY(exclusive_scan_u32(cell_first, cell_first, (int64_t)cells + 1, nullptr, scan_ws, st));
    interp_place_kernel<<<grid_for(n, kIBlock), kIBlock, 0, st>>>(keys, rank, cell_first, n, sorted_idx);
    UPCP_LAUNCH_OK();
    InterpArgs a;
    a.px = d_px; a.py = d_py; a.pv = d_pv; a.cell_first = cell_first; a.sorted_idx = sorted_idx; a.g = g;
    a.qx = d_qx; a.qy = d_qy; a.nq = nq; a.k = n_neighbors; a.method = method;
    a.max_dist = max_dist; a.power = power; a.reg = reg; a.conf_dist = conf_dist; a.fill = fill_value; a.out = d_out;
    interp_query_kernel<<<grid_for(nq, kIBlock), kIBlock, 0, st>>>(a);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

}  // extern "C"
