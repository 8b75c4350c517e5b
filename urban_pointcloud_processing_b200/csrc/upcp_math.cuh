// Exact per-element arithmetic of the hot path, written once for device and host.
//
// Everything here is compiled with FMA contraction OFF (nvcc -fmad=false; the
// host-emulation test build uses -ffp-contract=off) so that +,-,*,/ and sqrt are
// the IEEE-754 operations numpy / numba / the C++ natives of the reference
// perform.  Each function cites the reference lines it restates.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define UPCP_HD __host__ __device__ __forceinline__
#else
#define UPCP_HD inline
#endif

namespace upcp {

// ---------------------------------------------------------------- raster ----
// np.digitize(x, bins) - 1 for ASCENDING bins (interpolation.py:346):
// digitize returns #(bins <= x); the result -1 is then used as a numpy index and
// wraps to the last cell.  `guess` is any starting index (arithmetic estimate).
UPCP_HD int digitize_asc_wrap(const double* bins, int n, double x, int guess) {
    int i = guess;
    if (i < -1) i = -1;
    if (i > n - 1) i = n - 1;
    // i is the candidate for "#(bins <= x) - 1": need bins[i] <= x < bins[i+1].
    int steps = 0;
    while (i + 1 < n && bins[i + 1] <= x) {
        ++i;
        if (++steps > 4) {  // bad guess: finish by bisection on [i, n-1]
            int lo = i, hi = n - 1;  // invariant: bins[lo] <= x
            while (lo < hi) {
                int mid = (lo + hi + 1) >> 1;
                if (bins[mid] <= x) lo = mid; else hi = mid - 1;
            }
            i = lo;
            break;
        }
    }
    steps = 0;
    while (i >= 0 && !(bins[i] <= x)) {
        --i;
        if (++steps > 4) {  // bisection on [-1, i]
            int lo = -1, hi = i;  // invariant: !(bins[hi+1] <= x)
            while (lo < hi) {
                int mid = (lo + hi + 1) >> 1;
                if (bins[mid] <= x) lo = mid; else hi = mid - 1;
            }
            i = lo;
            break;
        }
    }
    return i < 0 ? n - 1 : i;
}

// np.digitize(y, bins, right=True) - 1 for DESCENDING bins (interpolation.py:347):
// digitize returns #(bins >= y); -1 wraps to the last row.
UPCP_HD int digitize_desc_wrap(const double* bins, int n, double y, int guess) {
    int i = guess;
    if (i < -1) i = -1;
    if (i > n - 1) i = n - 1;
    int steps = 0;
    while (i + 1 < n && bins[i + 1] >= y) {
        ++i;
        if (++steps > 4) {
            int lo = i, hi = n - 1;
            while (lo < hi) {
                int mid = (lo + hi + 1) >> 1;
                if (bins[mid] >= y) lo = mid; else hi = mid - 1;
            }
            i = lo;
            break;
        }
    }
    steps = 0;
    while (i >= 0 && !(bins[i] >= y)) {
        --i;
        if (++steps > 4) {
            int lo = -1, hi = i;
            while (lo < hi) {
                int mid = (lo + hi + 1) >> 1;
                if (bins[mid] >= y) lo = mid; else hi = mid - 1;
            }
            i = lo;
            break;
        }
    }
    return i < 0 ? n - 1 : i;
}

UPCP_HD int guess_index(double v, double origin, double inv_step, int n) {
    double t = (v - origin) * inv_step;
    if (!(t > -1.0)) return -1;          // also catches NaN
    if (t >= (double)n) return n - 1;
    return (int)floor(t);
}

// Height-threshold predicates (mode constants mirror include/upcp_cuda.h).
UPCP_HD bool classify_height(int mode, double z, double v, double a, double b) {
    switch (mode) {
        case 0: return fabs(z - v) < a;                    // ahn_fuser.py:159
        case 1: return z < v + a;                          // ahn_fuser.py:170
        case 2: return !isfinite(v) || z <= v + a;         // building_fuser.py:89-95
        case 3: return (z > v + a) && (z <= v + b);        // layer_lcc.py:72-74
        case 4: return (z <= v + b) && (z >= v - a);       // ahn_fuser.py:115-116
        case 5: return (z - v) < -a;                       // noise_filter.py:72
    }
    return false;
}

// ------------------------------------------------------------------- PIP ----
// One edge of clip_utils._point_inside_poly (clip_utils.py:130-160).
// Returns 0 = no crossing, 1 = crossing counted, 2 = point on the boundary.
UPCP_HD int pip_edge(double xi, double yi, double xj, double yj, double px, double py) {
    double dy = py - yi;
    double dy2 = py - yj;
    if (dy * dy2 <= 0.0 && (px >= xi || px >= xj)) {
        if (dy < 0 || dy2 < 0) {                      // non-horizontal line
            double F = dy * (xj - xi) / (dy - dy2) + xi;
            if (px > F) return 1;
            else if (px == F) return 2;
        } else if (dy2 == 0 &&
                   (px == xj || (dy == 0 && (px - xi) * (px - xj) <= 0))) {
            return 2;
        }
    }
    return 0;
}

// Whole ring, literal order (used by the host-emulation tests).
UPCP_HD int pip_ring(const double* xy, int n_vert, double px, double py) {
    int crossings = 0;
    for (int i = 0; i + 1 < n_vert; ++i) {
        int r = pip_edge(xy[2 * i], xy[2 * i + 1], xy[2 * i + 2], xy[2 * i + 3], px, py);
        if (r == 2) return 2;
        crossings += r;
    }
    return crossings & 1;
}

// ---------------------------------------------------------------- octree ----
// DgmOctree::getTheCellPosWhichIncludesThePoint at MAX_OCTREE_LEVEL, float32:
// (int)((p - dimMin) / cs), clamped to [0, 2^max_level - 1]   [ext, unverified].
UPCP_HD int octree_cell_pos(float p, float dim_min, float cs, int max_cells_minus1) {
    float t = (p - dim_min) / cs;
    if (!(t > 0.0f)) return 0;
    if (t >= (float)max_cells_minus1) return max_cells_minus1;
    return (int)t;
}

// ------------------------------------------------------ covariance / eigen --
// Open3D utility::ComputeCovariance / ComputeMeanAndCovariance (v0.16, cumulant
// form) [ext, unverified]: cumulants accumulated in neighbour order, divided by
// n, cov = E[ab] - E[a]E[b].
struct Cumulants {
    double c[9];
    UPCP_HD void clear() { for (int i = 0; i < 9; ++i) c[i] = 0.0; }
    UPCP_HD void add(double x, double y, double z) {
        c[0] += x; c[1] += y; c[2] += z;
        c[3] += x * x; c[4] += x * y; c[5] += x * z;
        c[6] += y * y; c[7] += y * z; c[8] += z * z;
    }
    // cov6 = xx, xy, xz, yy, yz, zz
    UPCP_HD void finish(int n, double* cov6) {
        double dn = (double)n;
        double m[9];
        for (int i = 0; i < 9; ++i) m[i] = c[i] / dn;
        cov6[0] = m[3] - m[0] * m[0];
        cov6[1] = m[4] - m[0] * m[1];
        cov6[2] = m[5] - m[0] * m[2];
        cov6[3] = m[6] - m[1] * m[1];
        cov6[4] = m[7] - m[1] * m[2];
        cov6[5] = m[8] - m[2] * m[2];
    }
};

struct Vec3 { double x, y, z; };
UPCP_HD Vec3 cross3(Vec3 a, Vec3 b) {
    Vec3 r;
    r.x = a.y * b.z - a.z * b.y;
    r.y = a.z * b.x - a.x * b.z;
    r.z = a.x * b.y - a.y * b.x;
    return r;
}
UPCP_HD double dot3(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// Open3D EstimateNormals.cpp ComputeEigenvector0 [ext, unverified].
UPCP_HD Vec3 eigenvector0(const double* A /*00 01 02 11 12 22*/, double eval0) {
    Vec3 row0 = {A[0] - eval0, A[1], A[2]};
    Vec3 row1 = {A[1], A[3] - eval0, A[4]};
    Vec3 row2 = {A[2], A[4], A[5] - eval0};
    Vec3 r0xr1 = cross3(row0, row1);
    Vec3 r0xr2 = cross3(row0, row2);
    Vec3 r1xr2 = cross3(row1, row2);
    double d0 = dot3(r0xr1, r0xr1);
    double d1 = dot3(r0xr2, r0xr2);
    double d2 = dot3(r1xr2, r1xr2);
    double dmax = d0;
    int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    Vec3 r;
    if (imax == 0) { double s = sqrt(d0); r.x = r0xr1.x / s; r.y = r0xr1.y / s; r.z = r0xr1.z / s; }
    else if (imax == 1) { double s = sqrt(d1); r.x = r0xr2.x / s; r.y = r0xr2.y / s; r.z = r0xr2.z / s; }
    else { double s = sqrt(d2); r.x = r1xr2.x / s; r.y = r1xr2.y / s; r.z = r1xr2.z / s; }
    return r;
}

// Open3D EstimateNormals.cpp ComputeEigenvector1 [ext, unverified].
UPCP_HD Vec3 eigenvector1(const double* A, Vec3 evec0, double eval1) {
    Vec3 U, V;
    if (fabs(evec0.x) > fabs(evec0.y)) {
        double inv_length = 1 / sqrt(evec0.x * evec0.x + evec0.z * evec0.z);
        U.x = -evec0.z * inv_length; U.y = 0; U.z = evec0.x * inv_length;
    } else {
        double inv_length = 1 / sqrt(evec0.y * evec0.y + evec0.z * evec0.z);
        U.x = 0; U.y = evec0.z * inv_length; U.z = -evec0.y * inv_length;
    }
    V = cross3(evec0, U);
    Vec3 AU = {A[0] * U.x + A[1] * U.y + A[2] * U.z,
               A[1] * U.x + A[3] * U.y + A[4] * U.z,
               A[2] * U.x + A[4] * U.y + A[5] * U.z};
    Vec3 AV = {A[0] * V.x + A[1] * V.y + A[2] * V.z,
               A[1] * V.x + A[3] * V.y + A[4] * V.z,
               A[2] * V.x + A[4] * V.y + A[5] * V.z};
    double m00 = U.x * AU.x + U.y * AU.y + U.z * AU.z - eval1;
    double m01 = U.x * AV.x + U.y * AV.y + U.z * AV.z;
    double m11 = V.x * AV.x + V.y * AV.y + V.z * AV.z - eval1;
    double absM00 = fabs(m00), absM01 = fabs(m01), absM11 = fabs(m11);
    double max_abs_comp;
    Vec3 r;
    if (absM00 >= absM11) {
        max_abs_comp = absM00 > absM01 ? absM00 : absM01;
        if (max_abs_comp > 0) {
            if (absM00 >= absM01) {
                m01 /= m00;
                m00 = 1 / sqrt(1 + m01 * m01);
                m01 *= m00;
            } else {
                m00 /= m01;
                m01 = 1 / sqrt(1 + m00 * m00);
                m00 *= m01;
            }
            r.x = m01 * U.x - m00 * V.x; r.y = m01 * U.y - m00 * V.y; r.z = m01 * U.z - m00 * V.z;
            return r;
        }
        return U;
    } else {
        max_abs_comp = absM11 > absM01 ? absM11 : absM01;
        if (max_abs_comp > 0) {
            if (absM11 >= absM01) {
                m01 /= m11;
                m11 = 1 / sqrt(1 + m01 * m01);
                m01 *= m11;
            } else {
                m11 /= m01;
                m01 = 1 / sqrt(1 + m11 * m11);
                m11 *= m01;
            }
            r.x = m11 * U.x - m01 * V.x; r.y = m11 * U.y - m01 * V.y; r.z = m11 * U.z - m01 * V.z;
            return r;
        }
        return U;
    }
}

// Open3D EstimateNormals.cpp FastEigen3x3 [ext, unverified]: eigenvector of the
// smallest eigenvalue of a symmetric 3x3 (cov6 = xx xy xz yy yz zz).  Also returns
// the three analytic eigenvalues (ascending-ish: eval[0..2] as the solver orders
// them, scaled back by max_coeff) through `evals` when non-null.
UPCP_HD Vec3 fast_eigen3x3(const double* cov6, double* evals) {
    double A[6];
    double max_coeff = cov6[0];
    for (int i = 1; i < 6; ++i) if (cov6[i] > max_coeff) max_coeff = cov6[i];
    Vec3 zero = {0.0, 0.0, 0.0};
    if (evals) { evals[0] = evals[1] = evals[2] = 0.0; }
    if (max_coeff == 0) return zero;
    for (int i = 0; i < 6; ++i) A[i] = cov6[i] / max_coeff;
    double norm = A[1] * A[1] + A[2] * A[2] + A[4] * A[4];
    if (norm > 0) {
        double q = (A[0] + A[3] + A[5]) / 3;
        double b00 = A[0] - q, b11 = A[3] - q, b22 = A[5] - q;
        double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2) / 6);
        double c00 = b11 * b22 - A[4] * A[4];
        double c01 = A[1] * b22 - A[4] * A[2];
        double c02 = A[1] * A[4] - b11 * A[2];
        double det = (b00 * c00 - A[1] * c01 + A[2] * c02) / (p * p * p);
        double half_det = det * 0.5;
        half_det = fmin(fmax(half_det, -1.0), 1.0);
        double angle = acos(half_det) / (double)3;
        const double two_thirds_pi = 2.09439510239319549;
        double beta2 = cos(angle) * 2;
        double beta0 = cos(angle + two_thirds_pi) * 2;
        double beta1 = -(beta0 + beta2);
        double e0 = q + p * beta0, e1 = q + p * beta1, e2 = q + p * beta2;
        if (evals) { evals[0] = e0 * max_coeff; evals[1] = e1 * max_coeff; evals[2] = e2 * max_coeff; }
        if (half_det >= 0) {
            Vec3 evec2 = eigenvector0(A, e2);
            if (e2 < e0 && e2 < e1) return evec2;
            Vec3 evec1 = eigenvector1(A, evec2, e1);
            if (e1 < e0 && e1 < e2) return evec1;
            return cross3(evec1, evec2);
        } else {
            Vec3 evec0 = eigenvector0(A, e0);
            if (e0 < e1 && e0 < e2) return evec0;
            Vec3 evec1 = eigenvector1(A, evec0, e1);
            if (e1 < e0 && e1 < e2) return evec1;
            return cross3(evec0, evec1);
        }
    } else {
        // diagonal matrix: A *= max_coeff restores cov6's diagonal
        double a00 = A[0] * max_coeff, a11 = A[3] * max_coeff, a22 = A[5] * max_coeff;
        if (evals) { evals[0] = a00; evals[1] = a11; evals[2] = a22; }
        Vec3 r;
        if (a00 < a11 && a00 < a22) { r.x = 1; r.y = 0; r.z = 0; }
        else if (a11 < a00 && a11 < a22) { r.x = 0; r.y = 1; r.z = 0; }
        else { r.x = 0; r.y = 0; r.z = 1; }
        return r;
    }
}

// Open3D EstimateNormals: zero normal -> (0,0,1)  [ext, unverified].
UPCP_HD Vec3 normal_from_cov(const double* cov6, double* evals) {
    Vec3 nrm = fast_eigen3x3(cov6, evals);
    double nn = sqrt(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z);
    if (nn == 0.0) { nrm.x = 0.0; nrm.y = 0.0; nrm.z = 1.0; }
    return nrm;
}

// math_utils.vector_angle (math_utils.py:9-18): degrees between u and v.
UPCP_HD double vector_angle_deg(Vec3 u, Vec3 v) {
    double nu = sqrt(u.x * u.x + u.y * u.y + u.z * u.z);
    double nv = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    double c = (u.x / nu) * (v.x / nv) + (u.y / nu) * (v.y / nv) + (u.z / nu) * (v.z / nv);
    double cl = fmin(1.0, fmax(c, -1.0));
    return acos(cl) * (180.0 / 3.14159265358979323846);
}

// Squared distance exactly as the neighbour-search ordering defines it.
UPCP_HD double dist2(double ax, double ay, double az, double bx, double by, double bz) {
    double dx = ax - bx, dy = ay - by, dz = az - bz;
    return (dx * dx + dy * dy) + dz * dz;
}

}  // namespace upcp
