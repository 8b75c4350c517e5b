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

template <bool SMEM_BINS>
__device__ __forceinline__ double raster_value(const RasterParams& r, const double* sbx,
                                               const double* sby, double inv_step_x, double inv_step_y,
                                               double px, double py) {
    const double* bx = SMEM_BINS ? sbx : r.bin_x;
    const double* by = SMEM_BINS ? sby : r.bin_y;
    // arithmetic guess (uniform-step estimate, steps inverted once per thread), then exact fix-up
    // against the same fp64 bin edges numpy compares with
    int gx = guess_index(px, bx[0], inv_step_x, r.nx);
    int gy = (py == py) ? guess_index(-py, -by[0], inv_step_y, r.ny) : -1;
    int ix = digitize_asc_wrap(bx, r.nx, px, gx);
    int iy = digitize_desc_wrap(by, r.ny, py, gy);
    return __ldg(r.values + (size_t)iy * r.nx + ix);
}

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
    if (SMEM_BINS) {
        for (int i = threadIdx.x; i < r.nx; i += kBlock) sbx[i] = r.bin_x[i];
        for (int i = threadIdx.x; i < r.ny; i += kBlock) sby[i] = r.bin_y[i];
        __syncthreads();
    }
    const double* bx = SMEM_BINS ? sbx : r.bin_x;
    const double* by = SMEM_BINS ? sby : r.bin_y;
    const double inv_step_x = r.nx >= 2 ? 1.0 / (bx[1] - bx[0]) : 0.0;
    const double inv_step_y = r.ny >= 2 ? 1.0 / (by[0] - by[1]) : 0.0;
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        bool s = sel ? (sel[i] != 0) : true;
        if (CLASSIFY) {
            bool res = false;
            if (s) {
                double v = raster_value<SMEM_BINS>(r, sbx, sby, inv_step_x, inv_step_y, x[i], y[i]);
                res = classify_height(mode, z[i], v, a, b);
            }
            if (out_mask) out_mask[i] = res ? 1 : 0;
            if (labels && res) labels[i] = assign;
        } else {
            if (s) out_val[i] = raster_value<SMEM_BINS>(r, sbx, sby, inv_step_x, inv_step_y, x[i], y[i]);
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
    int grid = grid_for(n, kBlock);
    bool smem_bins = nx <= kMaxBinsSmem && ny <= kMaxBinsSmem;
    size_t smem = smem_bins ? (size_t)(nx + ny) * sizeof(double) : 0;
    prof_begin(UPCP_PROF_RASTER, st);
#define UPCP_RASTER_LAUNCH(SB, CL)                                                   \
    raster_kernel<SB, CL><<<grid, kBlock, smem, st>>>(d_x, d_y, d_z, n, d_sel, r, mode, a, b, \
                                                     d_out_val, d_out_mask, d_labels, assign)
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
