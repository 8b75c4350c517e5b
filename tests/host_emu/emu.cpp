// Host build of the device arithmetic header (tests only): lets the CPU test-suite check
// urban_pointcloud_processing_b200/csrc/upcp_math.cuh against the oracle without a GPU.
// Never part of the product library.
#include "../../urban_pointcloud_processing_b200/csrc/upcp_math.cuh"
using namespace upcp;
extern "C" {
void emu_pip_ring(const double* xy, int nv, const double* px, const double* py, long n, unsigned char* out) {
    for (long i = 0; i < n; ++i) out[i] = pip_ring(xy, nv, px[i], py[i]) != 0;
}
void emu_digitize(const double* bx, int nx, const double* by, int ny, const double* px, const double* py, long n,
                  int* ix, int* iy) {
    double isx = nx >= 2 ? 1.0 / (bx[1] - bx[0]) : 0.0, isy = ny >= 2 ? 1.0 / (by[0] - by[1]) : 0.0;
    for (long i = 0; i < n; ++i) {
        int gx = guess_index(px[i], bx[0], isx, nx);
        int gy = (py[i] == py[i]) ? guess_index(-py[i], -by[0], isy, ny) : -1;
        ix[i] = digitize_asc_wrap(bx, nx, px[i], gx);
        iy[i] = digitize_desc_wrap(by, ny, py[i], gy);
    }
}
void emu_digitize_badguess(const double* bx, int nx, const double* by, int ny, const double* px, const double* py,
                           long n, int guess, int* ix, int* iy) {
    for (long i = 0; i < n; ++i) {
        ix[i] = digitize_asc_wrap(bx, nx, px[i], guess);
        iy[i] = digitize_desc_wrap(by, ny, py[i], guess);
    }
}
void emu_classify(int mode, const double* z, const double* v, long n, double a, double b, unsigned char* out) {
    for (long i = 0; i < n; ++i) out[i] = classify_height(mode, z[i], v[i], a, b);
}
void emu_normal_from_cov(const double* cov6, long n, double* out3, double* evals3) {
    for (long i = 0; i < n; ++i) {
        Vec3 v = normal_from_cov(cov6 + 6 * i, evals3 ? evals3 + 3 * i : nullptr);
        out3[3 * i] = v.x; out3[3 * i + 1] = v.y; out3[3 * i + 2] = v.z;
    }
}
void emu_cov(const double* pts, const int* ids, int n, double* cov6) {
    Cumulants c; c.clear();
    for (int j = 0; j < n; ++j) c.add(pts[3 * ids[j]], pts[3 * ids[j] + 1], pts[3 * ids[j] + 2]);
    c.finish(n, cov6);
}
void emu_octree_pos(const float* p, long n, float dim_min, float cs, int maxc, int* out) {
    for (long i = 0; i < n; ++i) out[i] = octree_cell_pos(p[i], dim_min, cs, maxc);
}
void emu_angle(const double* u, const double* v, long n, double* out) {
    for (long i = 0; i < n; ++i) {
        Vec3 a = {u[3 * i], u[3 * i + 1], u[3 * i + 2]}, b = {v[3 * i], v[3 * i + 1], v[3 * i + 2]};
        out[i] = vector_angle_deg(a, b);
    }
}
}
