// Tile layout, mask and label kernels (the glue passes between processors).
//
// All of these are single-pass, HBM-bound element-wise / reduction kernels:
// coalesced 16-byte loads where the element type allows, grids sized in
// multiples of the SM count, no shared-memory staging except for reductions
// and the AoS->SoA transposition.
#include "common.cuh"

namespace upcp {

constexpr int kBlock = 256;

// ------------------------------------------------------------ AoS -> SoA ----
template <int NCOLS>
__global__ void __launch_bounds__(kBlock)
aos_to_soa_kernel(const double* __restrict__ xyz, int64_t n,
                  double* __restrict__ x, double* __restrict__ y, double* __restrict__ z) {
    __shared__ double tile[kBlock * NCOLS];
    int64_t n_tiles = (n + kBlock - 1) / kBlock;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int64_t p0 = t * kBlock;
        int64_t rem = n - p0;
        int cnt = rem < kBlock ? (int)rem : kBlock;
        const double* src = xyz + p0 * NCOLS;
        for (int i = threadIdx.x; i < cnt * NCOLS; i += kBlock) tile[i] = src[i];
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            int i = threadIdx.x;
            x[p0 + i] = tile[i * NCOLS + 0];
            y[p0 + i] = tile[i * NCOLS + 1];
            z[p0 + i] = (NCOLS == 3) ? tile[i * NCOLS + (NCOLS - 1)] : 0.0;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ bbox ----
struct BBoxAcc {
    double mn[3], mx[3];   // all points
    double smn[3], smx[3]; // selected points
    unsigned long long cnt;
};

__device__ __forceinline__ double warp_min(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-level reduction of 12 extrema + count; result valid in thread 0.
__device__ void block_reduce_bbox(BBoxAcc& a) {
    __shared__ double s_v[12][kBlock / 32];
    __shared__ unsigned long long s_c[kBlock / 32];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = 0; d < 3; ++d) {
        a.mn[d] = warp_min(a.mn[d]);   a.mx[d] = warp_max(a.mx[d]);
        a.smn[d] = warp_min(a.smn[d]); a.smx[d] = warp_max(a.smx[d]);
    }
    a.cnt = warp_sum(a.cnt);
    if (lane == 0) {
        for (int d = 0; d < 3; ++d) {
            s_v[d][warp] = a.mn[d];      s_v[3 + d][warp] = a.mx[d];
            s_v[6 + d][warp] = a.smn[d]; s_v[9 + d][warp] = a.smx[d];
        }
        s_c[warp] = a.cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kBlock / 32; ++w) {
            for (int d = 0; d < 3; ++d) {
                s_v[d][0] = fmin(s_v[d][0], s_v[d][w]);
                s_v[3 + d][0] = fmax(s_v[3 + d][0], s_v[3 + d][w]);
                s_v[6 + d][0] = fmin(s_v[6 + d][0], s_v[6 + d][w]);
                s_v[9 + d][0] = fmax(s_v[9 + d][0], s_v[9 + d][w]);
            }
            s_c[0] += s_c[w];
        }
        for (int d = 0; d < 3; ++d) {
            a.mn[d] = s_v[d][0];      a.mx[d] = s_v[3 + d][0];
            a.smn[d] = s_v[6 + d][0]; a.smx[d] = s_v[9 + d][0];
        }
        a.cnt = s_c[0];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kBlock)
bbox_partial_kernel(const double* __restrict__ x, const double* __restrict__ y,
                    const double* __restrict__ z, int64_t n,
                    const uint8_t* __restrict__ sel, double* __restrict__ partial) {
    BBoxAcc a;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int d = 0; d < 3; ++d) { a.mn[d] = inf; a.mx[d] = -inf; a.smn[d] = inf; a.smx[d] = -inf; }
    a.cnt = 0;
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        double px = x[i], py = y[i], pz = z ? z[i] : 0.0;
        a.mn[0] = fmin(a.mn[0], px); a.mx[0] = fmax(a.mx[0], px);
        a.mn[1] = fmin(a.mn[1], py); a.mx[1] = fmax(a.mx[1], py);
        a.mn[2] = fmin(a.mn[2], pz); a.mx[2] = fmax(a.mx[2], pz);
        bool s = sel ? (sel[i] != 0) : true;
        if (s) {
            a.smn[0] = fmin(a.smn[0], px); a.smx[0] = fmax(a.smx[0], px);
            a.smn[1] = fmin(a.smn[1], py); a.smx[1] = fmax(a.smx[1], py);
            a.smn[2] = fmin(a.smn[2], pz); a.smx[2] = fmax(a.smx[2], pz);
            a.cnt++;
        }
    }
    block_reduce_bbox(a);
    if (threadIdx.x == 0) {
        double* o = partial + (size_t)blockIdx.x * 13;
        for (int d = 0; d < 3; ++d) {
            o[d] = a.mn[d]; o[3 + d] = a.mx[d]; o[6 + d] = a.smn[d]; o[9 + d] = a.smx[d];
        }
        o[12] = (double)a.cnt;
    }
}

__global__ void __launch_bounds__(kBlock)
bbox_final_kernel(const double* __restrict__ partial, int n_partial, double* __restrict__ out) {
    BBoxAcc a;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int d = 0; d < 3; ++d) { a.mn[d] = inf; a.mx[d] = -inf; a.smn[d] = inf; a.smx[d] = -inf; }
    a.cnt = 0;
    for (int i = threadIdx.x; i < n_partial; i += kBlock) {
        const double* p = partial + (size_t)i * 13;
        for (int d = 0; d < 3; ++d) {
            a.mn[d] = fmin(a.mn[d], p[d]);       a.mx[d] = fmax(a.mx[d], p[3 + d]);
            a.smn[d] = fmin(a.smn[d], p[6 + d]); a.smx[d] = fmax(a.smx[d], p[9 + d]);
        }
        a.cnt += (unsigned long long)p[12];
    }
    block_reduce_bbox(a);
    if (threadIdx.x == 0) {
        for (int d = 0; d < 3; ++d) {
            out[d] = a.mn[d]; out[3 + d] = a.mx[d]; out[6 + d] = a.smn[d]; out[9 + d] = a.smx[d];
        }
        out[12] = (double)a.cnt;
    }
}

// ------------------------------------------------------ label conversions ---
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kBlock)
convert_kernel(const TIn* __restrict__ in, int64_t n, TOut* __restrict__ out) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride)
        out[i] = (TOut)in[i];
}

// ------------------------------------------------------------ mask kernels --
struct LabelSets {
    int32_t eq[8];
    int32_t ne[8];
    int n_eq, n_ne;
};

__device__ __forceinline__ bool mask_rule(int32_t l, bool m, const LabelSets& sets) {
#pragma unroll
    for (int k = 0; k < 8; ++k) if (k < sets.n_eq) m = m || (l == sets.eq[k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) if (k < sets.n_ne) m = m && (l != sets.ne[k]);
    return m;
}

// The glue kernels are pure streams: four elements per thread through 16-byte label and 4-byte mask accesses
// (the entry points check the alignment; the tail is finished element-wise by the same kernel).
__global__ void __launch_bounds__(kBlock)
mask_build_kernel(const int32_t* __restrict__ labels, int64_t n,
                  const uint8_t* __restrict__ mask_in, int base, LabelSets sets,
                  uint8_t* __restrict__ out, int vec) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    const int64_t tid = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t n4 = vec ? n >> 2 : 0;
    for (int64_t i = tid; i < n4; i += stride) {
        const int4 l = reinterpret_cast<const int4*>(labels)[i];
        uchar4 mi = make_uchar4(base != 0, base != 0, base != 0, base != 0);
        if (mask_in) mi = reinterpret_cast<const uchar4*>(mask_in)[i];
        uchar4 o;
        o.x = mask_rule(l.x, mi.x != 0, sets); o.y = mask_rule(l.y, mi.y != 0, sets);
        o.z = mask_rule(l.z, mi.z != 0, sets); o.w = mask_rule(l.w, mi.w != 0, sets);
        reinterpret_cast<uchar4*>(out)[i] = o;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride)
        out[i] = mask_rule(labels[i], mask_in ? (mask_in[i] != 0) : (base != 0), sets) ? 1 : 0;
}

__global__ void __launch_bounds__(kBlock)
mask_combine_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                    int64_t n, int op, uint8_t* __restrict__ out, int vec) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    const int64_t tid = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    auto rule = [op](uint8_t x, uint8_t y) -> uint8_t {
        const bool va = x != 0, vb = y != 0;
        return ((op == 0) ? (va && vb) : (op == 1) ? (va || vb) : (va && !vb)) ? 1 : 0;
    };
    const int64_t n4 = vec ? n >> 2 : 0;
    for (int64_t i = tid; i < n4; i += stride) {
        const uchar4 x = reinterpret_cast<const uchar4*>(a)[i], y = reinterpret_cast<const uchar4*>(b)[i];
        reinterpret_cast<uchar4*>(out)[i] = make_uchar4(rule(x.x, y.x), rule(x.y, y.y), rule(x.z, y.z), rule(x.w, y.w));
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) out[i] = rule(a[i], b[i]);
}

__global__ void __launch_bounds__(kBlock)
labels_assign_kernel(int32_t* __restrict__ labels, int64_t n,
                     const uint8_t* __restrict__ mask, int32_t value, int vec) {
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    const int64_t tid = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t n4 = vec ? n >> 2 : 0;
    for (int64_t i = tid; i < n4; i += stride) {
        const uchar4 m = reinterpret_cast<const uchar4*>(mask)[i];
        int32_t* l = labels + (i << 2);
        if (m.x) l[0] = value;
        if (m.y) l[1] = value;
        if (m.z) l[2] = value;
        if (m.w) l[3] = value;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride)
        if (mask[i]) labels[i] = value;
}

__global__ void __launch_bounds__(kBlock)
mask_count_kernel(const uint8_t* __restrict__ mask, int64_t n, unsigned long long* __restrict__ out, int vec) {
    unsigned long long c = 0;
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    const int64_t tid = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t n16 = vec ? n >> 4 : 0;
    for (int64_t i = tid; i < n16; i += stride) {
        const uint4 w = reinterpret_cast<const uint4*>(mask)[i];
        // (0xff in every non-zero byte) -> 8 bits per counted element
        c += (__popc(__vcmpne4(w.x, 0u)) + __popc(__vcmpne4(w.y, 0u)) + __popc(__vcmpne4(w.z, 0u)) + __popc(__vcmpne4(w.w, 0u))) >> 3;
    }
    for (int64_t i = (n16 << 4) + tid; i < n; i += stride)
        c += mask[i] != 0;
    c = warp_sum(c);
    __shared__ unsigned long long s[kBlock / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < kBlock / 32; ++w) t += s[w];
        if (t) atomicAdd(out, t);
    }
}

__global__ void __launch_bounds__(kBlock)
label_hist_kernel(const int32_t* __restrict__ labels, int64_t n, unsigned long long* __restrict__ hist) {
    __shared__ unsigned int s_hist[256];
    for (int i = threadIdx.x; i < 256; i += kBlock) s_hist[i] = 0;
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        int32_t l = labels[i];
        unsigned int b = (l < 0 || l > 255) ? 255u : (unsigned int)l;
        atomicAdd(&s_hist[b], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += kBlock)
        if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
}

}  // namespace upcp

using namespace upcp;

static inline int aligned_to(const void* p, size_t a) { return ((uintptr_t)p % a) == 0; }   // NULL counts as aligned

namespace upcp {
// x = X * scale + offset in fp64, two roundings (this translation unit is compiled without FMA
// contraction): what laspy's scaled view computes on the host, from 12 B instead of 24 B per point
__global__ void __launch_bounds__(kBlock)
scaled_int_kernel(const int32_t* __restrict__ X, const int32_t* __restrict__ Y, const int32_t* __restrict__ Z,
                  int64_t n, double sx, double sy, double sz, double ox, double oy, double oz,
                  double* __restrict__ x, double* __restrict__ y, double* __restrict__ z) {
    int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        x[i] = __dadd_rn(__dmul_rn((double)X[i], sx), ox);
        y[i] = __dadd_rn(__dmul_rn((double)Y[i], sy), oy);
        z[i] = __dadd_rn(__dmul_rn((double)Z[i], sz), oz);
    }
}
}  // namespace upcp

extern "C" {

int upcp_points_from_scaled_int32(const int32_t* d_X, const int32_t* d_Y, const int32_t* d_Z, int64_t n,
                                  const double* h_scale3, const double* h_offset3,
                                  double* d_x, double* d_y, double* d_z, void* stream) {
    if (n < 0 || !h_scale3 || !h_offset3) UPCP_FAIL(UPCP_E_INVALID, "points_from_scaled_int32: bad argument");
    if (n == 0) return UPCP_OK;
    upcp::scaled_int_kernel<<<grid_for(n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        d_X, d_Y, d_Z, n, h_scale3[0], h_scale3[1], h_scale3[2], h_offset3[0], h_offset3[1], h_offset3[2], d_x, d_y, d_z);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

int upcp_points_aos_to_soa(const double* d_xyz, int64_t n, int ncols,
                           double* d_x, double* d_y, double* d_z, void* stream) {
    if (n < 0 || (ncols != 2 && ncols != 3)) UPCP_FAIL(UPCP_E_INVALID, "aos_to_soa: ncols must be 2 or 3");
    if (n == 0) return UPCP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = grid_for(n, kBlock);
    if (ncols == 3) aos_to_soa_kernel<3><<<grid, kBlock, 0, st>>>(d_xyz, n, d_x, d_y, d_z);
    else aos_to_soa_kernel<2><<<grid, kBlock, 0, st>>>(d_xyz, n, d_x, d_y, d_z);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

size_t upcp_bbox_ws_bytes(int64_t n) {
    (void)n;
    return align_up((size_t)4096 * 13 * sizeof(double));
}

int upcp_bbox(const double* d_x, const double* d_y, const double* d_z, int64_t n,
              const uint8_t* d_sel, double* d_out13, void* d_ws, size_t ws_bytes, void* stream) {
    if (n < 0) UPCP_FAIL(UPCP_E_INVALID, "bbox: n < 0");
    if (ws_bytes < upcp_bbox_ws_bytes(n) || !d_ws) UPCP_FAIL(UPCP_E_WORKSPACE, "bbox: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int grid = grid_for(n, kBlock, 4);
    if (grid > 4096) grid = 4096;
    double* partial = (double*)d_ws;
    bbox_partial_kernel<<<grid, kBlock, 0, st>>>(d_x, d_y, d_z, n, d_sel, partial);
    UPCP_LAUNCH_OK();
    bbox_final_kernel<<<1, kBlock, 0, st>>>(partial, grid, d_out13);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

#define UPCP_CONVERT(NAME, TIN, TOUT)                                              \
    int NAME(const TIN* d_in, int64_t n, TOUT* d_out, void* stream) {              \
        if (n < 0) UPCP_FAIL(UPCP_E_INVALID, #NAME ": n < 0");                     \
        if (n == 0) return UPCP_OK;                                                \
        convert_kernel<TIN, TOUT><<<grid_for(n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(d_in, n, d_out); \
        UPCP_LAUNCH_OK();                                                          \
        return UPCP_OK;                                                            \
    }
UPCP_CONVERT(upcp_labels_widen_u8, uint8_t, int32_t)
UPCP_CONVERT(upcp_labels_widen_u16, uint16_t, int32_t)
UPCP_CONVERT(upcp_labels_narrow_u8, int32_t, uint8_t)
UPCP_CONVERT(upcp_labels_narrow_u16, int32_t, uint16_t)

int upcp_mask_build(const int32_t* d_labels, int64_t n, const uint8_t* d_mask_in,
                    int base, const int32_t* h_eq, int n_eq, const int32_t* h_ne, int n_ne,
                    uint8_t* d_out, void* stream) {
    if (n < 0 || n_eq < 0 || n_eq > 8 || n_ne < 0 || n_ne > 8)
        UPCP_FAIL(UPCP_E_INVALID, "mask_build: at most 8 eq / 8 ne labels");
    if (n == 0) return UPCP_OK;
    LabelSets sets;
    sets.n_eq = n_eq; sets.n_ne = n_ne;
    for (int k = 0; k < 8; ++k) { sets.eq[k] = k < n_eq ? h_eq[k] : 0; sets.ne[k] = k < n_ne ? h_ne[k] : 0; }
    mask_build_kernel<<<grid_for((n + 3) / 4, kBlock), kBlock, 0, (cudaStream_t)stream>>>(d_labels, n, d_mask_in, base, sets, d_out,
        aligned_to(d_labels, 16) && aligned_to(d_mask_in, 4) && aligned_to(d_out, 4));
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

int upcp_mask_combine(const uint8_t* d_a, const uint8_t* d_b, int64_t n, int op,
                      uint8_t* d_out, void* stream) {
    if (n < 0 || op < 0 || op > 2) UPCP_FAIL(UPCP_E_INVALID, "mask_combine: bad op");
    if (n == 0) return UPCP_OK;
    mask_combine_kernel<<<grid_for((n + 3) / 4, kBlock), kBlock, 0, (cudaStream_t)stream>>>(d_a, d_b, n, op, d_out,
        aligned_to(d_a, 4) && aligned_to(d_b, 4) && aligned_to(d_out, 4));
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

int upcp_labels_assign(int32_t* d_labels, int64_t n, const uint8_t* d_mask, int32_t value, void* stream) {
    if (n < 0) UPCP_FAIL(UPCP_E_INVALID, "labels_assign: n < 0");
    if (n == 0) return UPCP_OK;
    labels_assign_kernel<<<grid_for((n + 3) / 4, kBlock), kBlock, 0, (cudaStream_t)stream>>>(d_labels, n, d_mask, value,
        aligned_to(d_labels, 16) && aligned_to(d_mask, 4));
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

int upcp_mask_count(const uint8_t* d_mask, int64_t n, int64_t* d_count, void* stream) {
    if (n < 0) UPCP_FAIL(UPCP_E_INVALID, "mask_count: n < 0");
    cudaStream_t st = (cudaStream_t)stream;
    UPCP_CUDA_OK(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    if (n == 0) return UPCP_OK;
    mask_count_kernel<<<grid_for((n + 15) / 16, kBlock, 4), kBlock, 0, st>>>(d_mask, n, (unsigned long long*)d_count, aligned_to(d_mask, 16));
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

int upcp_label_hist(const int32_t* d_labels, int64_t n, int64_t* d_hist256, void* stream) {
    if (n < 0) UPCP_FAIL(UPCP_E_INVALID, "label_hist: n < 0");
    if (n == 0) return UPCP_OK;
    label_hist_kernel<<<grid_for(n, kBlock, 4), kBlock, 0, (cudaStream_t)stream>>>(d_labels, n, (unsigned long long*)d_hist256);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}

}  // extern "C"
