// S1 -- AHN raster lookup fused with the height-threshold compare.
//
// Restates FastGridInterpolator.__call__ (interpolation.py:336-348) and the
// threshold tests of AHNFuser / BGTBuildingFuser / LayerLCC / NoiseFilter (see
// classify_height in upcp_math.cuh).  HBM-bound: 24 B/pt of coordinates in,
// 1 B/pt mask (+4 B/pt label RMW when fused) out; the bin edges (<= 8 KB) are
// staged in shared memory once per CTA and the raster (2 MB fp64 for a 500x500
// AHN tile) stays L2-resident and is read through the read-only path.
#include "common.cuh"
#include "upcp_math.cuh"

namespace upcp {

constexpr int kBlock = 256;
constexpr int kMaxBinsSmem = 2048;  // per axis; larger rasters read bins from global

struct RasterParams {
    const double* bin_x; int nx;
    const double* bin_y; int ny;
    const double* values;
};

// One axis of the cell index.  `t` = (coordinate - first edge) / step in the direction the edges run.
// When the edges of the axis were verified (once per CTA, while they are staged) to lie within 1e-7 of a
// step of first + i x step, a coordinate whose `t` is farther than 1e-6 from an integer is on the same
// side of every edge as its arithmetic position says -- the rounding of `t` is twelve orders smaller --
// and the index is floor(t) without reading an edge.  The others (4 in a million, NaN, non-uniform edges)
// take the exact search against the edges numpy compares with.
template <bool ASC>
__device__ __forceinline__ int raster_axis(const double* bins, int n, bool uniform, double t, double v) {
    const double f = floor(t);
    if (uniform && t - f >= 1e-6 && t - f <= 1.0 - 1e-6)
        return (t < 0.0 || t >= (double)n) ? n - 1 : (int)f;       // "-1" wraps to the last cell like the numpy index
    const int g = (t == t) ? (!(t > -1.0) ? -1 : (t >= (double)n ? n - 1 : (int)f)) : -1;
    return ASC ? digitize_asc_wrap(bins, n, v, g) : digitize_desc_wrap(bins, n, v, g);
}

// kPerThread points per thread and round, every load of a round independent of the others of its kind:
// the selection bytes, then the coordinates of the selected points, then the raster values -- three
// round trips per FOUR points instead of per point (the kernel was latency- not bandwidth-limited with
// one point in flight per thread: ncu r2, issue slots 56-62 % busy at 3.3 TB/s).
constexpr int kPerThread = 4;

template <bool SMEM_BINS, bool CLASSIFY>
__global__ void __launch_bounds__(kBlock)
raster_kernel(const double* __restrict__ x, const double* __restrict__ y,
              const double* __restrict__ z, int64_t n, const uint8_t* __restrict__ sel,
              RasterParams r, int mode, double a, double b,
              double* __restrict__ out_val, uint8_t* __restrict__ out_mask,
              int32_t* __restrict__ labels, int32_t assign) {
    extern __shared__ double s_bins[];
    double* sbx = s_bins;
    double* sby = s_bins + (SMEM_BINS ? r.nx : 0);
    const double b0x = r.bin_x[0], b0y = r.bin_y[0];
    // the mean step over the whole axis: the difference of two neighbouring edges carries their rounding
    const double step_x = r.nx >= 2 ? (r.bin_x[r.nx - 1] - b0x) / (double)(r.nx - 1) : 0.0;
    const double step_y = r.ny >= 2 ? (b0y - r.bin_y[r.ny - 1]) / (double)(r.ny - 1) : 0.0;
    bool uni_x = false, uni_y = false;
    if (SMEM_BINS) {
        bool okx = step_x > 0.0, oky = step_y > 0.0;
        for (int i = threadIdx.x; i < r.nx; i += kBlock) {
            const double e = r.bin_x[i];
            sbx[i] = e;
            okx = okx && fabs(e - (b0x + (double)i * step_x)) <= 1e-7 * step_x;
        }
        for (int i = threadIdx.x; i < r.ny; i += kBlock) {
            const double e = r.bin_y[i];
            sby[i] = e;
            oky = oky && fabs(e - (b0y - (double)i * step_y)) <= 1e-7 * step_y;
        }
        uni_x = __syncthreads_and(okx) != 0;
        uni_y = __syncthreads_and(oky) != 0;
    }
    const double* bx = SMEM_BINS ? sbx : r.bin_x;
    const double* by = SMEM_BINS ? sby : r.bin_y;
    const double inv_step_x = step_x != 0.0 ? 1.0 / step_x : 0.0;
    const double inv_step_y = step_y != 0.0 ? 1.0 / step_y : 0.0;
    constexpr int kTile = kPerThread * kBlock;
    for (int64_t base = (int64_t)blockIdx.x * kTile; base < n; base += (int64_t)gridDim.x * kTile) {
        bool s[kPerThread];
        double px[kPerThread], py[kPerThread], pz[kPerThread], v[kPerThread];
#pragma unroll
        for (int u = 0; u < kPerThread; ++u) {
            const int64_t i = base + u * kBlock + threadIdx.x;
            s[u] = i < n && (sel ? (sel[i] != 0) : true);
        }
#pragma unroll
        for (int u = 0; u < kPerThread; ++u) {
            const int64_t i = base + u * kBlock + threadIdx.x;
            px[u] = py[u] = pz[u] = 0.0;
            if (s[u]) {
                px[u] = x[i]; py[u] = y[i];
                if (CLASSIFY) pz[u] = z[i];
            }
        }
#pragma unroll
        for (int u = 0; u < kPerThread; ++u) {
            v[u] = 0.0;
            if (s[u]) {
                const int ix = raster_axis<true>(bx, r.nx, uni_x, (px[u] - b0x) * inv_step_x, px[u]);
                const int iy = raster_axis<false>(by, r.ny, uni_y, (b0y - py[u]) * inv_step_y, py[u]);
                v[u] = __ldg(r.values + (size_t)iy * r.nx + ix);
            }
        }
#pragma unroll
        for (int u = 0; u < kPerThread; ++u) {
            const int64_t i = base + u * kBlock + threadIdx.x;
            if (CLASSIFY) {
                const bool res = s[u] && classify_height(mode, pz[u], v[u], a, b);
                if (i < n && out_mask) out_mask[i] = res ? 1 : 0;
                if (labels && res) labels[i] = assign;
            } else {
                if (s[u]) out_val[i] = v[u];
            }
        }
    }
}

static int launch_raster(bool classify, const double* d_x, const double* d_y, const double* d_z,
                         int64_t n, const uint8_t* d_sel, const double* d_bin_x, int nx,
                         const double* d_bin_y, int ny, const double* d_values,
                         int mode, double a, double b,
                         double* d_out_val, uint8_t* d_out_mask, int32_t* d_labels,
                         int32_t assign, cudaStream_t st) {
    RasterParams r;
    r.bin_x = d_bin_x; r.nx = nx; r.bin_y = d_bin_y; r.ny = ny; r.values = d_values;
    bool smem_bins = nx <= kMaxBinsSmem && ny <= kMaxBinsSmem;
    size_t smem = smem_bins ? (size_t)(nx + ny) * sizeof(double) : 0;
    prof_begin(UPCP_PROF_RASTER, st);
    // one wave of resident CTAs (the kernel strides over tiles of kPerThread x kBlock points): a grid
    // larger than what is resident would run its tail at partial occupancy
#define UPCP_RASTER_LAUNCH(SB, CL)                                                                       \
    do {                                                                                                 \
        int per_sm = 0;                                                                                  \
        UPCP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, raster_kernel<SB, CL>, kBlock, smem)); \
        int grid = grid_for((n + kPerThread - 1) / kPerThread, kBlock, per_sm < 1 ? 1 : per_sm);         \
        raster_kernel<SB, CL><<<grid, kBlock, smem, st>>>(d_x, d_y, d_z, n, d_sel, r, mode, a, b,        \
                                                         d_out_val, d_out_mask, d_labels, assign);       \
    } while (0)
    if (smem_bins) { if (classify) UPCP_RASTER_LAUNCH(true, true); else UPCP_RASTER_LAUNCH(true, false); }
    else           { if (classify) UPCP_RASTER_LAUNCH(false, true); else UPCP_RASTER_LAUNCH(false, false); }
#undef UPCP_RASTER_LAUNCH
    prof_end(UPCP_PROF_RASTER, st, n);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

}  // namespace upcp

using namespace upcp;

extern "C" {

int upcp_raster_lookup(const double* d_x, const double* d_y, int64_t n, const uint8_t* d_sel,
                       const double* d_bin_x, int nx, const double* d_bin_y, int ny,
                       const double* d_values, double* d_out, void* stream) {
    if (n < 0 || nx < 1 || ny < 1) UPCP_FAIL(UPCP_E_INVALID, "raster_lookup: bad sizes");
    if (n == 0) return UPCP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return launch_raster(false, d_x, d_y, nullptr, n, d_sel, d_bin_x, nx, d_bin_y, ny, d_values,
                         0, 0.0, 0.0, d_out, nullptr, nullptr, 0, st);
}

int upcp_raster_classify(const double* d_x, const double* d_y, const double* d_z, int64_t n,
                         const uint8_t* d_sel, const double* d_bin_x, int nx,
                         const double* d_bin_y, int ny, const double* d_values,
                         int mode, double a, double b, uint8_t* d_out, int32_t* d_labels,
                         int32_t assign, void* stream) {
    if (n < 0 || nx < 1 || ny < 1) UPCP_FAIL(UPCP_E_INVALID, "raster_classify: bad sizes");
    if (mode < 0 || mode > 5) UPCP_FAIL(UPCP_E_INVALID, "raster_classify: bad mode");
    if (!d_out && !d_labels) UPCP_FAIL(UPCP_E_INVALID, "raster_classify: no output given");
    if (n == 0) return UPCP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return launch_raster(true, d_x, d_y, d_z, n, d_sel, d_bin_x, nx, d_bin_y, ny, d_values,
                         mode, a, b, nullptr, d_out, d_labels, assign, st);
}

}  // extern "C"
