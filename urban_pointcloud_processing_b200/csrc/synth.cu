// Synthetic urban tiles generated ON THE DEVICE ("tile i generated on device
// from seed 20240000 + i"): measurement infrastructure for bench.py and the full-size tests, not part of
// the labelling path.  Same scene model and value distributions as synthetic.make_tile (ground, facades
// with balconies, car shells, poles / crowns, clutter, uniform noise; coordinates on the LAS millimetre
// grid), one counter-based Philox4x32-10 stream per point, so a tile is a pure function of
// (seed, layout, n) whatever the launch geometry.  The random stream is NOT numpy's: a device tile is
// its own data set (downloaded when the oracle has to label the same points).
#include <math.h>
#include "common.cuh"

namespace upcp {

struct Philox {
    uint32_t c[4], k[2];
    __device__ __forceinline__ static void round_(uint32_t* c, const uint32_t* k) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    }
    // 4 x 32 random bits for (point index, block number) under the tile's seed
    __device__ __forceinline__ static void block(uint64_t index, uint32_t blk, uint64_t seed, uint32_t* out) {
        uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), blk, 0x55504350u};
        uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            round_(c, k);
            k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};

// Per-point generator state: uniform doubles in [0, 1) with 53 random bits (two words each).
struct Rng {
    uint64_t index, seed;
    uint32_t blk, have;
    uint32_t w[4];
    __device__ __forceinline__ uint32_t next32() {
        if (have == 0) { Philox::block(index, blk++, seed, w); have = 4; }
        return w[4 - have--];
    }
    __device__ __forceinline__ double uniform() {
        const uint64_t a = next32(), b = next32();
        return (double)(((a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
    }
    __device__ __forceinline__ double normal() {           // Box-Muller (one of the pair)
        const double u1 = 1.0 - uniform(), u2 = uniform();
        return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
    __device__ __forceinline__ int below(int n) { return (int)(uniform() * (double)n) % n; }
};

struct SynthLayout {
    double ox, oy, size;
    double cum_kind[6];            // cumulative point fractions: ground, facade, car, furniture, clutter, noise
    int n_walls, n_cars, n_poles, n_crowns;
    const double* walls;           // [n_walls][8]: x0, y0, x1, y1, height, normal x, normal y, cumulative area share
    const double* cars;            // [n_cars][4]:  centre x, y, dimension x, y
    const double* poles;           // [n_poles][4]: x, y, radius, height
    const double* crowns;          // [n_crowns][4]: x, y, radius, centre z
};

__device__ __forceinline__ double ground_z(double x, double y, const SynthLayout& L) {
    const double u = (x - L.ox) / L.size, v = (y - L.oy) / L.size;
    const double tp = 6.283185307179586;
    return 0.6 + 0.3 * sin(tp * u * 1.5) * cos(tp * v) + 0.15 * sin(tp * (u + v) * 2.0);
}

__global__ void __launch_bounds__(256)
synth_tile_kernel(int64_t n, uint64_t seed, SynthLayout L, double* __restrict__ x, double* __restrict__ y,
                  double* __restrict__ z, uint8_t* __restrict__ kind) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Rng r; r.index = (uint64_t)i; r.seed = seed; r.blk = 0; r.have = 0;
        const double uk = r.uniform();
        int k = 0;
        while (k < 5 && uk >= L.cum_kind[k]) ++k;
        double px, py, pz;
        if (k == 0) {                                             // ground
            px = L.ox + r.uniform() * L.size; py = L.oy + r.uniform() * L.size;
            pz = ground_z(px, py, L) + 0.02 * r.normal();
        } else if (k == 1) {                                      // facade (6 % balcony points)
            const double ua = r.uniform();
            int w = 0;
            while (w < L.n_walls - 1 && ua >= L.walls[w * 8 + 7]) ++w;
            const double* W = L.walls + w * 8;
            const double t = r.uniform();
            const double fx = W[0] + t * (W[2] - W[0]), fy = W[1] + t * (W[3] - W[1]);
            pz = ground_z(fx, fy, L) + r.uniform() * W[4];
            double perp = 0.01 * r.normal();
            if (r.uniform() < 0.06) perp = 0.8 * r.uniform();
            px = fx + perp * W[5]; py = fy + perp * W[6];
        } else if (k == 2) {                                      // car shells
            const double* C = L.cars + r.below(L.n_cars) * 4;
            const int face = r.below(5);
            const double u = r.uniform() - 0.5, v = r.uniform() - 0.5, w = r.uniform();
            const double lx = face == 0 ? -0.5 : (face == 1 ? 0.5 : u);
            const double ly = face == 2 ? -0.5 : (face == 3 ? 0.5 : v);
            const double lz = face == 4 ? 1.0 : w;
            px = C[0] + lx * C[2] + 0.005 * r.normal();
            py = C[1] + ly * C[3] + 0.005 * r.normal();
            pz = ground_z(px, py, L) + 0.15 + lz * 1.5;
        } else if (k == 3) {                                      // street furniture: crowns (60 %) and poles
            if (r.uniform() < 0.6) {
                const double* K = L.crowns + r.below(L.n_crowns) * 4;
                double dx = r.normal(), dy = r.normal(), dz = r.normal();
                const double nrm = sqrt(dx * dx + dy * dy + dz * dz) + 1e-300;
                const double rr = K[2] * cbrt(r.uniform());
                px = K[0] + dx / nrm * rr; py = K[1] + dy / nrm * rr; pz = K[3] + dz / nrm * rr * 0.8;
            } else {
                const double* P = L.poles + r.below(L.n_poles) * 4;
                const double ang = 6.283185307179586 * r.uniform();
                px = P[0] + P[2] * cos(ang); py = P[1] + P[2] * sin(ang);
                pz = ground_z(px, py, L) + r.uniform() * P[3];
            }
        } else if (k == 4) {                                      // low clutter near the poles
            const double* P = L.poles + r.below(L.n_poles) * 4;
            px = P[0] + 0.4 * r.normal(); py = P[1] + 0.4 * r.normal();
            pz = ground_z(px, py, L) + fabs(0.3 + 0.4 * r.normal());
        } else {                                                  // uniform noise
            px = L.ox + r.uniform() * L.size; py = L.oy + r.uniform() * L.size;
            pz = -2.0 + 27.0 * r.uniform();
        }
        px = fmin(fmax(px, L.ox), L.ox + L.size);
        py = fmin(fmax(py, L.oy), L.oy + L.size);
        x[i] = rint(px * 1000.0) * 0.001; y[i] = rint(py * 1000.0) * 0.001; z[i] = rint(pz * 1000.0) * 0.001;
        kind[i] = (uint8_t)k;
    }
}

}  // namespace upcp

using namespace upcp;

extern "C" int upcp_synth_tile(int64_t n, uint64_t seed, const double* h_origin_size3, const double* h_cum_kind6,
                               const double* d_walls8, int32_t n_walls, const double* d_cars4, int32_t n_cars,
                               const double* d_poles4, int32_t n_poles, const double* d_crowns4, int32_t n_crowns,
                               double* d_x, double* d_y, double* d_z, uint8_t* d_kind, void* stream) {
    if (n < 0 || !h_origin_size3 || !h_cum_kind6 || n_walls < 1 || n_cars < 1 || n_poles < 1 || n_crowns < 1)
        UPCP_FAIL(UPCP_E_INVALID, "synth_tile: bad argument");
    if (n == 0) return UPCP_OK;
    SynthLayout L;
    L.ox = h_origin_size3[0]; L.oy = h_origin_size3[1]; L.size = h_origin_size3[2];
    for (int i = 0; i < 6; ++i) L.cum_kind[i] = h_cum_kind6[i];
    L.n_walls = n_walls; L.n_cars = n_cars; L.n_poles = n_poles; L.n_crowns = n_crowns;
    L.walls = d_walls8; L.cars = d_cars4; L.poles = d_poles4; L.crowns = d_crowns4;
    synth_tile_kernel<<<grid_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(n, seed, L, d_x, d_y, d_z, d_kind);
    UPCP_LAUNCH_OK();
    return UPCP_OK;
}
