// solve.cu — K2: one persistent CTA per frame pair does everything after the distance tiles:
//   A. merge K1's partials, ratio test, mutual filter        (src/matching.cpp:50-61, :76-86)
//   B. build matched_pt / matched_ls in ascending prev index  (src/stereoFrameHandler.cpp:144-152, :167-179)
//      as structure-of-arrays in shared memory (coalesced, conflict-free)
//   C. optimizePose                                            (src/stereoFrameHandler.cpp:307-392):
//      Gauss-Newton (:394-431) / robust Gauss-Newton (:433-480) with the per-feature residual, 1x6 Jacobian row
//      and Cauchy weight evaluated by all threads (:549-694, :696-962), warp-shuffle + shared-memory reduction
//      into the 21 + 6 + 1 normal-equation sums, 6x6 solve / SE(3) update / stop tests on-chip,
//      removeOutliers with median / MAD by bitonic sort (:988-1067, src/auxiliar.cpp:387-430),
//      isGoodSolution (:292-305), pose finalisation (:372-391).
// Features are read from HBM exactly once per solve; every GN evaluation runs out of shared memory.
// All arithmetic is double precision like the reference (B200 has a full-rate FP64 pipe); the summation order
// differs from the reference's sequential lists (fixed tree order -> run-to-run deterministic).
#include <math.h>

#include "common.cuh"
#include "match_finalize.cuh"

namespace plstvo {

// ---- SoA views of the matched lists --------------------------------------------------------------
struct Feat {
    double *Px, *Py, *Pz, *pu, *pv, *ps2;                                               // points
    double *sX, *sY, *sZ, *eX, *eY, *eZ, *l0, *l1, *l2, *su, *sv, *eu, *ev, *ls2;       // lines
    uint8_t *inl_p, *inl_l;
    int np, nl;
};

struct State {
    double red[K2_WARPS][ACC_N + 1];
    double acc[ACC_N + 1];   // reduced sums: H upper triangle (21), g (6), e (1), count (1)
    double DT[16];           // pose being optimised
    double DT0[16];          // initial pose of optimizePose
    double H[36];
    double cov[36];
    double err;
    double scal[4];          // block-wide scalars (median, stdv, mean, ...)
    int    ctrl;             // loop control broadcast
    int    n_inl_p, n_inl_l;
    int    evals;
    int    scan[K2_WARPS];
    long long tc[8];         // debug phase timers: 0 match-finalize 1 gather 2 GN-eval 3 GN-serial 4 gates/eig 5 outliers 6 final
    PlPoseResult out;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t k2_smem_bytes(int cap_pt, int cap_ls, int sort_cap, bool feat_in_smem) {
    size_t b = align_up(sizeof(State), 16);
    b += (size_t)sort_cap * sizeof(double);
    b += align_up((size_t)cap_pt, 16) + align_up((size_t)cap_ls, 16);                      // inlier flags
    b += align_up((size_t)cap_pt * 2, 16) + align_up((size_t)cap_ls * 2, 16);              // prev index of entry k
    if (feat_in_smem) b += ((size_t)6 * cap_pt + (size_t)14 * cap_ls) * sizeof(double);
    return b;
}

// ---- tiny dense algebra (thread 0) -------------------------------------------------------------------
__device__ void mat4_identity(double* T) {
    for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
}
__device__ void mat4_mul(const double* A, const double* B, double* C) {
    double R[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0.0;
            for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j];
            R[i * 4 + j] = s;
        }
    for (int i = 0; i < 16; i++) C[i] = R[i];
}
__device__ bool mat4_is_identity(const double* T) {
    for (int i = 0; i < 16; i++)
        if (T[i] != ((i % 5 == 0) ? 1.0 : 0.0)) return false;
    return true;
}
__device__ void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ void skew3(double x, double y, double z, double* S) {  // src/auxiliar.cpp:29-44
    S[0] = 0;  S[1] = -z; S[2] = y;
    S[3] = z;  S[4] = 0;  S[5] = -x;
    S[6] = -y; S[7] = x;  S[8] = 0;
}
__device__ void inverse_se3(const double* T, double* Ti) {  // src/auxiliar.cpp:113-122
    double R[16];
    mat4_identity(R);
    for (int i = 0; i < 3; i++) {
        double s = 0.0;
        for (int j = 0; j < 3; j++) {
            R[i * 4 + j] = T[j * 4 + i];
            s += T[j * 4 + i] * T[j * 4 + 3];
        }
        R[i * 4 + 3] = -s;
    }
    for (int i = 0; i < 16; i++) Ti[i] = R[i];
}
__device__ void expmap_se3(const double* x, double* T) {  // src/auxiliar.cpp:124-141, x = [t; w]
    double t[3] = {x[0], x[1], x[2]};
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double theta = sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5]);
    if (!(theta < 0.000001)) {
        double s[9], ss[9], V[9];
        skew3(x[3] / theta, x[4] / theta, x[5] / theta, s);
        mat3_mul(s, s, ss);
        const double sn = sin(theta), cs = cos(theta);
        for (int i = 0; i < 9; i++) {
            const double I = (i % 4 == 0) ? 1.0 : 0.0;
            R[i] = I + s[i] * sn + ss[i] * (1.0 - cs);
            V[i] = I + s[i] * (1.0 - cs) / theta + ss[i] * (theta - sn) / theta;
        }
        double tv[3];
        for (int i = 0; i < 3; i++) tv[i] = V[i * 3] * t[0] + V[i * 3 + 1] * t[1] + V[i * 3 + 2] * t[2];
        t[0] = tv[0]; t[1] = tv[1]; t[2] = tv[2];
    }
    mat4_identity(T);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) T[i * 4 + j] = R[i * 3 + j];
        T[i * 4 + 3] = t[i];
    }
}
__device__ void mat3_inverse(const double* A, double* Ai) {
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double id = 1.0 / (A[0] * c00 + A[1] * c01 + A[2] * c02);
    Ai[0] = c00 * id; Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id; Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    Ai[3] = c01 * id; Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id; Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    Ai[6] = c02 * id; Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id; Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}
__device__ void logmap_se3(const double* T, double* x) {  // src/auxiliar.cpp:143-173
    double R[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, w[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i * 3 + j] = T[i * 4 + j];
    double cosine = (R[0] + R[4] + R[8] - 1.0) / 2.0;
    if (cosine > 1.0) cosine = 1.0;
    else if (cosine < -1.0) cosine = -1.0;
    double sine = sqrt(1.0 - cosine * cosine);
    if (sine > 1.0) sine = 1.0;
    const double theta = acos(cosine);
    if (theta > 0.000001) {
        w[0] = theta * (R[7] - R[5]) / (2.0 * sine);
        w[1] = theta * (R[2] - R[6]) / (2.0 * sine);
        w[2] = theta * (R[3] - R[1]) / (2.0 * sine);
        double s[9], ss[9];
        skew3(w[0] / theta, w[1] / theta, w[2] / theta, s);
        mat3_mul(s, s, ss);
        for (int i = 0; i < 9; i++) {
            const double I = (i % 4 == 0) ? 1.0 : 0.0;
            V[i] = I + s[i] * (1.0 - cosine) / theta + ss[i] * (theta - sine) / theta;
        }
    }
    double Vi[9];
    mat3_inverse(V, Vi);
    for (int i = 0; i < 3; i++) x[i] = Vi[i * 3] * T[3] + Vi[i * 3 + 1] * T[7] + Vi[i * 3 + 2] * T[11];
    x[3] = w[0]; x[4] = w[1]; x[5] = w[2];
}
__device__ void unccomp_se3(const double* T1, const double* c1, const double* cinc, double* out) {
    // src/auxiliar.cpp:175-197: cov1 + Ad(T1) covinc Ad(T1)^T, Ad = [R, skew(t) R; 0, R]
    double Ad[36], S[9], R[9], SR[9], tmp[36];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i * 3 + j] = T1[i * 4 + j];
    skew3(T1[3], T1[7], T1[11], S);
    mat3_mul(S, R, SR);
    for (int i = 0; i < 36; i++) Ad[i] = 0.0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            Ad[i * 6 + j] = R[i * 3 + j];
            Ad[i * 6 + 3 + j] = SR[i * 3 + j];
            Ad[(i + 3) * 6 + 3 + j] = R[i * 3 + j];
        }
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double s = 0.0;
            for (int k = 0; k < 6; k++) s += Ad[i * 6 + k] * cinc[k * 6 + j];
            tmp[i * 6 + j] = s;
        }
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double s = 0.0;
            for (int k = 0; k < 6; k++) s += tmp[i * 6 + k] * Ad[j * 6 + k];
            out[i * 6 + j] = c1[i * 6 + j] + s;
        }
}

// ColPivHouseholderQR<Matrix6d>(H).solve(g) + logAbsDeterminant (src/stereoFrameHandler.cpp:417-418, :453-455)
__device__ void qr6_solve(const double* H, const double* g, double* x, double* log_abs_det) {
    double A[36], c[6], v[6];
    int perm[6];
    for (int i = 0; i < 36; i++) A[i] = H[i];
    for (int i = 0; i < 6; i++) { c[i] = g[i]; perm[i] = i; }
    double maxpivot = 0.0;
    for (int k = 0; k < 6; k++) {
        int best = k;
        double bestn = -1.0;
        for (int j = k; j < 6; j++) {
            double s = 0.0;
            for (int i = k; i < 6; i++) s += A[i * 6 + j] * A[i * 6 + j];
            if (s > bestn) { bestn = s; best = j; }
        }
        if (best != k) {
            for (int i = 0; i < 6; i++) { const double t = A[i * 6 + k]; A[i * 6 + k] = A[i * 6 + best]; A[i * 6 + best] = t; }
            const int t = perm[k]; perm[k] = perm[best]; perm[best] = t;
        }
        double tail = 0.0;
        for (int i = k + 1; i < 6; i++) tail += A[i * 6 + k] * A[i * 6 + k];
        const double c0 = A[k * 6 + k];
        double beta, tau;
        for (int i = 0; i < 6; i++) v[i] = 0.0;
        if (tail <= 2.2250738585072014e-308) {
            tau = 0.0;
            beta = c0;
        } else {
            beta = sqrt(c0 * c0 + tail);
            if (c0 >= 0.0) beta = -beta;
            for (int i = k + 1; i < 6; i++) v[i] = A[i * 6 + k] / (c0 - beta);
            tau = (beta - c0) / beta;
        }
        v[k] = 1.0;
        for (int j = k + 1; j < 6; j++) {
            double s = 0.0;
            for (int i = k; i < 6; i++) s += v[i] * A[i * 6 + j];
            s *= tau;
            for (int i = k; i < 6; i++) A[i * 6 + j] -= s * v[i];
        }
        double s = 0.0;
        for (int i = k; i < 6; i++) s += v[i] * c[i];
        s *= tau;
        for (int i = k; i < 6; i++) c[i] -= s * v[i];
        A[k * 6 + k] = beta;
        if (fabs(beta) > maxpivot) maxpivot = fabs(beta);
    }
    int rank = 0;
    const double thr = maxpivot * (2.220446049250313e-16 * 6.0);
    double lad = 0.0;
    for (int k = 0; k < 6; k++) {
        if (fabs(A[k * 6 + k]) > thr) rank++;
        lad += log(fabs(A[k * 6 + k]));
    }
    *log_abs_det = lad;
    double y[6] = {0, 0, 0, 0, 0, 0};
    for (int k = rank - 1; k >= 0; k--) {
        double s2 = c[k];
        for (int j = k + 1; j < rank; j++) s2 -= A[k * 6 + j] * y[j];
        y[k] = s2 / A[k * 6 + k];
    }
    for (int k = 0; k < 6; k++) x[perm[k]] = y[k];
}

// Matrix6d::inverse() (partial-pivot LU), src/stereoFrameHandler.cpp:429, :470
__device__ void inv6(const double* Ain, double* Ainv) {
    double A[36], B[36];
    for (int i = 0; i < 36; i++) { A[i] = Ain[i]; B[i] = (i % 7 == 0) ? 1.0 : 0.0; }
    for (int k = 0; k < 6; k++) {
        int piv = k;
        double big = fabs(A[k * 6 + k]);
        for (int i = k + 1; i < 6; i++)
            if (fabs(A[i * 6 + k]) > big) { big = fabs(A[i * 6 + k]); piv = i; }
        if (piv != k)
            for (int j = 0; j < 6; j++) {
                double t = A[k * 6 + j]; A[k * 6 + j] = A[piv * 6 + j]; A[piv * 6 + j] = t;
                t = B[k * 6 + j]; B[k * 6 + j] = B[piv * 6 + j]; B[piv * 6 + j] = t;
            }
        const double d = A[k * 6 + k];
        for (int i = k + 1; i < 6; i++) {
            const double f = A[i * 6 + k] / d;
            for (int j = k + 1; j < 6; j++) A[i * 6 + j] -= f * A[k * 6 + j];
            for (int j = 0; j < 6; j++) B[i * 6 + j] -= f * B[k * 6 + j];
        }
    }
    for (int j = 0; j < 6; j++)
        for (int i = 5; i >= 0; i--) {
            double s = B[i * 6 + j];
            for (int k = i + 1; k < 6; k++) s -= A[i * 6 + k] * Ainv[k * 6 + j];
            Ainv[i * 6 + j] = s / A[i * 6 + i];
        }
}

// SelfAdjointEigenSolver<Matrix6d>::eigenvalues(): lower triangle, ascending (cyclic Jacobi)
__device__ void eig6_sym(const double* Ain, double* w) {
    double A[36];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j <= i; j++) A[i * 6 + j] = A[j * 6 + i] = Ain[i * 6 + j];
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < 6; i++) {
            diag += A[i * 6 + i] * A[i * 6 + i];
            for (int j = 0; j < i; j++) off += 2.0 * A[i * 6 + j] * A[i * 6 + j];
        }
        if (!(off > 1e-30 * diag) || off == 0.0) break;   /* off-diagonal mass at rounding level: eigenvalues settled to ~1e-15 */
        for (int p = 0; p < 5; p++)
            for (int q = p + 1; q < 6; q++) {
                const double apq = A[p * 6 + q];
                if (apq == 0.0) continue;
                const double tau = (A[q * 6 + q] - A[p * 6 + p]) / (2.0 * apq);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = t * cs;
                for (int k = 0; k < 6; k++) {
                    const double akp = A[k * 6 + p], akq = A[k * 6 + q];
                    A[k * 6 + p] = cs * akp - sn * akq;
                    A[k * 6 + q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 6; k++) {
                    const double apk = A[p * 6 + k], aqk = A[q * 6 + k];
                    A[p * 6 + k] = cs * apk - sn * aqk;
                    A[q * 6 + k] = sn * apk + cs * aqk;
                }
            }
    }
    for (int i = 0; i < 6; i++) w[i] = A[i * 6 + i];
    for (int i = 1; i < 6; i++) {  // insertion sort, ascending
        const double v = w[i];
        int j = i - 1;
        while (j >= 0 && w[j] > v) { w[j + 1] = w[j]; j--; }
        w[j + 1] = v;
    }
}

// isGoodSolution (src/stereoFrameHandler.cpp:292-305)
__device__ bool is_good_solution(const double* DT, const double* cov, double err, double* eig_out) {
    double w[6];
    eig6_sym(cov, w);
    if (eig_out)
        for (int i = 0; i < 6; i++) eig_out[i] = w[i];
    bool finite = true;
    for (int i = 0; i < 16; i++) {
        const double d = DT[i] - DT[i];
        if (!(d == d)) finite = false;
    }
    return !(w[0] < 0.0 || w[5] > 1.0 || err < 0.0 || err > 1.0 || !finite);
}

// ---- per-feature arithmetic --------------------------------------------------------------------------
struct Cam {
    double fx, fy, cx, cy;
};

__device__ __forceinline__ void transform(const double* DT, double x, double y, double z, double& X, double& Y, double& Z) {
    X = (DT[0] * x + DT[1] * y + DT[2] * z) + DT[3];   // DT.block(0,0,3,3) * P + DT.col(3).head(3)  (:567)
    Y = (DT[4] * x + DT[5] * y + DT[6] * z) + DT[7];
    Z = (DT[8] * x + DT[9] * y + DT[10] * z) + DT[11];
}

__device__ __forceinline__ double point_residual(const Feat& f, int i, const double* DT, const Cam& c, double& X,
                                                 double& Y, double& Z, double& dx, double& dy) {
    transform(DT, f.Px[i], f.Py[i], f.Pz[i], X, Y, Z);
    dx = (c.cx + c.fx * X / Z) - f.pu[i];   // PinholeStereoCamera::projection (src/pinholeStereoCamera.cpp:231-237)
    dy = (c.cy + c.fy * Y / Z) - f.pv[i];
    return sqrt(dx * dx + dy * dy);
}

struct LineRes {
    double sX, sY, sZ, eX, eY, eZ, spu, spv, epu, epv, ds, de;
};

__device__ __forceinline__ double line_residual(const Feat& f, int i, const double* DT, const Cam& c, LineRes& r) {
    transform(DT, f.sX[i], f.sY[i], f.sZ[i], r.sX, r.sY, r.sZ);
    transform(DT, f.eX[i], f.eY[i], f.eZ[i], r.eX, r.eY, r.eZ);
    r.spu = c.cx + c.fx * r.sX / r.sZ;
    r.spv = c.cy + c.fy * r.sY / r.sZ;
    r.epu = c.cx + c.fx * r.eX / r.eZ;
    r.epv = c.cy + c.fy * r.eY / r.eZ;
    const double l0 = f.l0[i], l1 = f.l1[i], l2 = f.l2[i];
    r.ds = l0 * r.spu + l1 * r.spv + l2;   // :621-622
    r.de = l0 * r.epu + l1 * r.epv + l2;
    return sqrt(r.ds * r.ds + r.de * r.de);
}

__device__ __forceinline__ void jac_aux(double fgz2, double gx, double gy, double gz, double dx, double dy, double* J) {
    J[0] = +fgz2 * dx * gz;                                   // :582-587 / :636-641
    J[1] = +fgz2 * dy * gz;
    J[2] = -fgz2 * (gx * dx + gy * dy);
    J[3] = -fgz2 * (gx * gy * dx + gy * gy * dy + gz * gz * dy);
    J[4] = +fgz2 * (gx * gx * dx + gz * gz * dx + gx * gy * dy);
    J[5] = +fgz2 * (gx * gz * dy - gy * gz * dx);
}

__device__ __forceinline__ double overlap_from_lambdas(double ls, double le) {
    const double lo = (le < ls) ? le : ls, hi = (ls < le) ? le : ls;
    if (lo < 0.0 && hi > 1.0) return 1.0;
    if (hi < 0.0 || lo > 1.0) return 0.0;
    if (lo < 0.0) return hi;
    if (hi > 1.0) return 1.0 - lo;
    return hi - lo;
}

// StereoFrame::lineSegmentOverlap (src/stereoFrame.cpp:510-616) with the PREVIOUS frame's endpoints
__device__ __forceinline__ double line_overlap(double su, double sv, double eu, double ev, double pu, double pv,
                                               double qu, double qv) {
    const double lx = eu - su, ly = ev - sv;
    if (fabs(su - eu) < 1.0) return overlap_from_lambdas((pv - sv) / ly, (qv - sv) / ly);
    if (fabs(sv - ev) < 1.0) return overlap_from_lambdas((pu - su) / lx, (qu - su) / lx);
    const double a = sv - ev, b = eu - su, c = su * ev - eu * sv;
    const double lxy = 1.0 / (a * a + b * b);
    const double sx = (b * (b * pu - a * pv) - a * c) * lxy;
    const double ex = (b * (b * qu - a * qv) - a * c) * lxy;
    return overlap_from_lambdas((sx - su) / lx, (ex - su) / lx);
}

__device__ __forceinline__ void accumulate(double* acc, const double* J, double r, double w) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const double Jw = J[i] * w;
#pragma unroll
        for (int j = i; j < 6; j++) acc[k++] += Jw * J[j];
        acc[21 + i] += Jw * r;
    }
    acc[27] += r * r * w;
    acc[28] += 1.0;
}

// block-wide sum of ACC_N + 1 doubles per thread -> st.acc (fixed order: deterministic)
__device__ void block_reduce_acc(State& st, double* acc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k <= ACC_N; k++) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (lane == 0) st.red[warp][k] = v;
    }
    __syncthreads();
    if (tid <= ACC_N) {
        double s = 0.0;
        for (int w = 0; w < K2_WARPS; w++) s += st.red[w][tid];
        st.acc[tid] = s;
    }
    __syncthreads();
}

__device__ double block_sum(State& st, double v) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    __syncthreads();
    if (lane == 0) st.red[warp][0] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < K2_WARPS; w++) s += st.red[w][0];
    __syncthreads();
    return s;
}

// ---- bitonic sort in shared memory: a[0..m), m a power of two -----------------------------------------
__device__ void bitonic_sort(double* a, int m) {
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int k = 2; k <= m; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < m; i += nth) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const double x = a[i], y = a[ixj];
                    const bool up = ((i & k) == 0);
                    if ((x > y) == up) {
                        a[i] = y;
                        a[ixj] = x;
                    }
                }
            }
            __syncthreads();
        }
}

__device__ __forceinline__ int pow2_ceil(int n) {
    int m = 1;
    while (m < n) m <<= 1;
    return m;
}

// median / MAD of the n finite values placed (unordered, padded with +inf up to m) in `buf`:
// median = sorted[n/2]; stdv = 1.4826 * sorted(|x - median| rounded to float)[n/2]  (src/auxiliar.cpp:396-403)
__device__ void median_mad(double* buf, int n, int m, State& st) {
    const int tid = threadIdx.x, nth = blockDim.x;
    bitonic_sort(buf, m);
    const double median = buf[n / 2];
    __syncthreads();
    for (int i = tid; i < n; i += nth) buf[i] = (double)fabsf((float)(buf[i] - median));
    __syncthreads();
    bitonic_sort(buf, m);
    if (tid == 0) {
        st.scal[0] = median;
        st.scal[1] = 1.4826 * buf[n / 2];
    }
    __syncthreads();
}

// ---- optimizeFunctions / optimizeFunctionsRobust -------------------------------------------------------
__device__ void robust_scales(const Feat& f, State& st, double* sortbuf, const double* DT, const Cam& cam,
                              double& s_p, double& s_l) {
    // pre-weight pass + MAD scales (src/stereoFrameHandler.cpp:707-781)
    const int tid = threadIdx.x, nth = blockDim.x;
    const double th_min = 0.0001, th_max = sqrt(7.815);
    {   // points: residual norms of the inliers, order irrelevant for a median
        const int m = pow2_ceil(max(f.np, 1));
        double cnt = 0.0;
        for (int i = tid; i < m; i += nth) {
            double v = INFINITY;
            if (i < f.np && f.inl_p[i]) {
                double X, Y, Z, dx, dy;
                v = point_residual(f, i, DT, cam, X, Y, Z, dx, dy);
                cnt += 1.0;
            }
            sortbuf[i] = v;
        }
        const int n = (int)block_sum(st, cnt);   // res_p.size(): the inliers only (:710-720)
        if (n > 0) {
            median_mad(sortbuf, n, m, st);
            s_p = st.scal[1];
        } else
            s_p = 0.0;
        __syncthreads();
    }
    {
        const int m = pow2_ceil(max(f.nl, 1));
        double cnt = 0.0;
        for (int i = tid; i < m; i += nth) {
            double v = INFINITY;
            if (i < f.nl && f.inl_l[i]) {
                LineRes r;
                v = line_residual(f, i, DT, cam, r);
                cnt += 1.0;
            }
            sortbuf[i] = v;
        }
        const int n = (int)block_sum(st, cnt);
        if (n > 0) {
            median_mad(sortbuf, n, m, st);
            s_l = st.scal[1];
        } else
            s_l = 0.0;
        __syncthreads();
    }
    if (s_p < th_min) s_p = th_min;
    if (s_p > th_max) s_p = th_max;
    if (s_l < th_min) s_l = th_min;
    if (s_l > th_max) s_l = th_max;
}

// leaves the reduced sums in st.acc (H upper triangle, g, e, N)
__device__ void evaluate(const Feat& f, State& st, double* sortbuf, const double* DT, const Cam& cam,
                         double homog_th, bool robust) {
    const int tid = threadIdx.x, nth = blockDim.x;
    double s_p = 1.0, s_l = 1.0;
    if (robust) robust_scales(f, st, sortbuf, DT, cam, s_p, s_l);

    double acc[ACC_N + 1];
#pragma unroll
    for (int k = 0; k <= ACC_N; k++) acc[k] = 0.0;

    for (int i = tid; i < f.np; i += nth) {   // point block (:563-606 / :785-870)
        if (!f.inl_p[i]) continue;
        double X, Y, Z, dx, dy, J[6];
        const double n = point_residual(f, i, DT, cam, X, Y, Z, dx, dy);
        const double fgz2 = cam.fx / fmax(homog_th, Z * Z);
        jac_aux(fgz2, X, Y, Z, dx, dy, J);
        const double den = fmax(homog_th, n);
#pragma unroll
        for (int k = 0; k < 6; k++) J[k] = J[k] / den;
        double r, w;
        if (!robust) {
            r = n * sqrt(f.ps2[i]);
            w = 1.0 / (1.0 + r * r);             // robustWeightCauchy (src/auxiliar.cpp:556-559)
        } else {
            r = n;
            const double x = r / s_p;
            w = 1.0 / (1.0 + x * x);
        }
        accumulate(acc, J, r, w);
    }
    for (int i = tid; i < f.nl; i += nth) {   // line block (:610-684 / :874-952)
        if (!f.inl_l[i]) continue;
        LineRes lr;
        double Js[6], Je[6], J[6];
        const double n = line_residual(f, i, DT, cam, lr);
        const double lx = f.l0[i], ly = f.l1[i];
        jac_aux(cam.fx / fmax(homog_th, lr.sZ * lr.sZ), lr.sX, lr.sY, lr.sZ, lx, ly, Js);
        jac_aux(cam.fx / fmax(homog_th, lr.eZ * lr.eZ), lr.eX, lr.eY, lr.eZ, lx, ly, Je);
        const double den = fmax(homog_th, n);
#pragma unroll
        for (int k = 0; k < 6; k++) J[k] = (Js[k] * lr.ds + Je[k] * lr.de) / den;
        double r, w;
        if (!robust) {
            r = n * sqrt(f.ls2[i]);
            w = 1.0 / (1.0 + r * r);
        } else {
            r = n;
            const double x = r / s_l;
            w = 1.0 / (1.0 + x * x);
        }
        w *= line_overlap(f.su[i], f.sv[i], f.eu[i], f.ev[i], lr.spu, lr.spv, lr.epu, lr.epv);   // :668, :930
        accumulate(acc, J, r, w);
    }
    block_reduce_acc(st, acc);
}

// thread 0: unpack st.acc into H (full symmetric), g, err = e / N
__device__ void unpack_normal_equations(State& st, double* g, double& err) {
    int k = 0;
    for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++) {
            st.H[i * 6 + j] = st.acc[k];
            st.H[j * 6 + i] = st.acc[k];
            k++;
        }
    for (int i = 0; i < 6; i++) g[i] = st.acc[21 + i];
    err = st.acc[27] / st.acc[28];   // e /= (N_l + N_p)  (:692)
}

__device__ void apply_increment(double* DT, const double* inc) {   // DT << DT * inverse_se3(expmap_se3(inc))  (:419)
    double E[16], Ei[16];
    expmap_se3(inc, E);
    inverse_se3(E, Ei);
    mat4_mul(DT, Ei, DT);
}

// gaussNewtonOptimization (:394-431) and gaussNewtonOptimizationRobust (:433-480).
// Pose in st.DT (in/out), covariance to st.cov, error to st.err.  Block-wide; thread 0 runs the 6x6 part.
__device__ void gauss_newton(const Feat& f, State& st, double* sortbuf, const Cam& cam, const PlConfig& cfg,
                             int max_iters, bool robust) {
    const int tid = threadIdx.x;
    double err_prev = 999999999.9, err = 0.0;   // thread 0's copies are the authoritative ones
    double DTstart[16];
    bool good = true, fail_first = false;
    if (tid == 0) {
        for (int i = 0; i < 16; i++) DTstart[i] = st.DT[i];
        for (int i = 0; i < 36; i++) st.H[i] = 0.0;
        st.evals = 0;
    }
    for (int it = 0; it < max_iters; it++) {
        const long long t_a = clock64();
        evaluate(f, st, sortbuf, st.DT, cam, cfg.homog_th, robust);
        const long long t_b = clock64();
        if (tid == 0) {
            double g[6], inc[6], lad;
            int ctrl = 0;   // 0 continue, 1 stop
            st.evals++;
            unpack_normal_equations(st, g, err);
            if (!robust) {
                if (err > err_prev) {
                    ctrl = 1;
                    if (it == 0) fail_first = true;
                } else if ((err < cfg.min_error) || fabs(err - err_prev) < cfg.min_error_change) {
                    ctrl = 1;
                } else {
                    qr6_solve(st.H, g, inc, &lad);
                    apply_increment(st.DT, inc);
                    if (sqrt(inc[0] * inc[0] + inc[1] * inc[1] + inc[2] * inc[2]) < cfg.min_error_change &&
                        sqrt(inc[3] * inc[3] + inc[4] * inc[4] + inc[5] * inc[5]) < cfg.min_error_change)
                        ctrl = 1;
                    err_prev = err;
                }
            } else {
                if ((fabs(err - err_prev) < cfg.min_error_change) || (err < cfg.min_error)) {
                    ctrl = 1;
                } else {
                    qr6_solve(st.H, g, inc, &lad);
                    if (lad < 0.0) {
                        good = false;
                        ctrl = 1;
                    } else {
                        apply_increment(st.DT, inc);
                        double n2 = 0.0;
                        for (int i = 0; i < 6; i++) n2 += inc[i] * inc[i];
                        if (sqrt(n2) < cfg.min_error_change) ctrl = 1;
                        err_prev = err;
                    }
                }
            }
            st.ctrl = ctrl;
            st.tc[2] += t_b - t_a;
            st.tc[3] += clock64() - t_b;
        }
        __syncthreads();
        const int ctrl = st.ctrl;
        __syncthreads();
        if (ctrl) break;
    }
    if (tid == 0) {
        const long long t_c = clock64();
        if (fail_first) {
            st.err = -1.0;   // :408-409: DT_cov left untouched
        } else if (good) {
            inv6(st.H, st.cov);
            st.err = err;
        } else {   // :473-478
            for (int i = 0; i < 16; i++) st.DT[i] = DTstart[i];
            st.err = -1.0;
            for (int i = 0; i < 36; i++) st.cov[i] = (i % 7 == 0) ? 1.0 : 0.0;
        }
        st.tc[3] += clock64() - t_c;
    }
    __syncthreads();
}

// removeOutliers (:988-1067) at pose DT (shared)
__device__ void remove_outliers(const Feat& f, State& st, double* sortbuf, const double* DT, const Cam& cam,
                                const PlConfig& cfg) {
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int type = 0; type < 2; type++) {
        const int n = type ? f.nl : f.np;
        if (type == 0 ? !cfg.has_points : !cfg.has_lines) continue;
        if (n == 0) continue;   // vector_mean_stdv_mad of an empty vector: nothing to flag
        const int m = pow2_ceil(n);
        // residuals of ALL matched features (inliers or not)
        auto residual = [&](int i) -> double {
            if (type == 0) {
                double X, Y, Z, dx, dy;
                return point_residual(f, i, DT, cam, X, Y, Z, dx, dy) * sqrt(f.ps2[i]);
            }
            LineRes r;
            return line_residual(f, i, DT, cam, r) * sqrt(f.ls2[i]);
        };
        for (int i = tid; i < m; i += nth) sortbuf[i] = (i < n) ? residual(i) : INFINITY;
        __syncthreads();
        median_mad(sortbuf, n, m, st);
        const double stdv = st.scal[1];
        __syncthreads();
        // mean of the residuals below 2 stdv if there are enough of them, else plain mean (auxiliar.cpp:406-427)
        double s_sel = 0.0, c_sel = 0.0, s_all = 0.0;
        for (int i = tid; i < n; i += nth) {
            const double r = residual(i);
            s_all += r;
            if (r < 2.0 * stdv) {
                s_sel += r;
                c_sel += 1.0;
            }
        }
        s_sel = block_sum(st, s_sel);
        c_sel = block_sum(st, c_sel);
        s_all = block_sum(st, s_all);
        const int k = (int)c_sel;
        const double mean = (k >= (int)(0.2 * (double)n)) ? s_sel / (double)k : s_all / (double)n;
        const double th = cfg.inlier_k * stdv;
        double removed = 0.0;
        uint8_t* inl = type ? f.inl_l : f.inl_p;
        for (int i = tid; i < n; i += nth)
            if (inl[i] && fabs(residual(i) - mean) > th) {
                inl[i] = 0;
                removed += 1.0;
            }
        removed = block_sum(st, removed);
        if (tid == 0) {
            if (type == 0) st.n_inl_p -= (int)removed;
            else st.n_inl_l -= (int)removed;
        }
        __syncthreads();
    }
}

// block-wide exclusive scan of one int per thread; returns the thread's offset, total in *total
__device__ int block_exclusive_scan(State& st, int v, int* total) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) st.scan[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
    for (int w = 0; w < K2_WARPS; w++) {
        if (w < warp) base += st.scan[w];
        tot += st.scan[w];
    }
    *total = tot;
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(K2_THREADS, 1) track_solve_kernel(const SolveParams prm) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, nth = blockDim.x;
    const int pair = prm.first_pair + blockIdx.x;

    // ---- carve shared memory ----
    State& st = *reinterpret_cast<State*>(smem);
    size_t off = align_up(sizeof(State), 16);
    double* sortbuf = reinterpret_cast<double*>(smem + off);
    off += (size_t)prm.sort_cap * sizeof(double);
    Feat f;
    f.inl_p = smem + off;  off += align_up((size_t)prm.cap_pt, 16);
    f.inl_l = smem + off;  off += align_up((size_t)prm.cap_ls, 16);
    uint16_t* midx_p = reinterpret_cast<uint16_t*>(smem + off);  off += align_up((size_t)prm.cap_pt * 2, 16);
    uint16_t* midx_l = reinterpret_cast<uint16_t*>(smem + off);  off += align_up((size_t)prm.cap_ls * 2, 16);
    double* fb = prm.feat_in_smem ? reinterpret_cast<double*>(smem + off)
                                  : prm.feat_scratch + (size_t)blockIdx.x * prm.feat_scratch_stride;
    {
        const int cp = prm.cap_pt, cl = prm.cap_ls;
        f.Px = fb; f.Py = fb + cp; f.Pz = fb + 2 * cp; f.pu = fb + 3 * cp; f.pv = fb + 4 * cp; f.ps2 = fb + 5 * cp;
        double* lb = fb + 6 * (size_t)cp;
        f.sX = lb; f.sY = lb + cl; f.sZ = lb + 2 * cl; f.eX = lb + 3 * cl; f.eY = lb + 4 * cl; f.eZ = lb + 5 * cl;
        f.l0 = lb + 6 * cl; f.l1 = lb + 7 * cl; f.l2 = lb + 8 * cl; f.su = lb + 9 * cl; f.sv = lb + 10 * cl;
        f.eu = lb + 11 * cl; f.ev = lb + 12 * cl; f.ls2 = lb + 13 * cl;
    }
    const PlConfig& cfg = prm.cfg;
    const Cam cam = {prm.cam.fx, prm.cam.fy, prm.cam.cx, prm.cam.cy};

    if (tid == 0)
        for (int i = 0; i < 8; i++) st.tc[i] = 0;
    long long t_ph = clock64();
    int n1p = 0, n1l = 0;        // prev-frame feature counts (mode 0) / list lengths (mode 1)
    size_t out_p0 = 0, out_l0 = 0;   // where this pair's inlier flags start

    if (prm.mode == 0) {
        // ---- A. finish the matching: merge partials, ratio test, mutual filter -> m12 (global) ----
        const MatchProblem pp = prm.problems[2 * pair], pl = prm.problems[2 * pair + 1];
        match_finalize_block(pp, reinterpret_cast<int32_t*>(sortbuf));
        __syncthreads();
        match_finalize_block(pl, reinterpret_cast<int32_t*>(sortbuf));
        __syncthreads();
        if (tid == 0) { st.tc[0] += clock64() - t_ph; }
        t_ph = clock64();

        // ---- B. f2fTracking glue: ordered compaction of the matched features into SoA ----
        const FrameDev& P = prm.prev;
        const FrameDev& C = prm.curr;
        const int a0 = P.pt_off[pair], b0 = C.pt_off[pair];
        n1p = P.pt_off[pair + 1] - a0;
        out_p0 = (size_t)a0;
        {
            const int per = (n1p + nth - 1) / nth, lo = min(n1p, tid * per), hi = min(n1p, lo + per);
            int cnt = 0;
            for (int i = lo; i < hi; i++) cnt += (pp.m12[i] >= 0);
            int total;
            int k = block_exclusive_scan(st, cnt, &total);
            for (int i = lo; i < hi; i++) {
                const int i2 = pp.m12[i];
                if (i2 < 0) continue;
                const double* p3 = P.pt_P + 3 * (size_t)(a0 + i);
                f.Px[k] = p3[0]; f.Py[k] = p3[1]; f.Pz[k] = p3[2];
                const double* o2 = C.pt_pl + 2 * (size_t)(b0 + i2);    // pl_obs = curr pl (:148)
                f.pu[k] = o2[0]; f.pv[k] = o2[1];
                f.ps2[k] = P.pt_sigma2[a0 + i];                        // PointFeature::safeCopy keeps sigma2
                f.inl_p[k] = 1;
                midx_p[k] = (uint16_t)i;
                k++;
            }
            f.np = total;
        }
        const int c0 = P.ls_off[pair], d0 = C.ls_off[pair];
        n1l = P.ls_off[pair + 1] - c0;
        out_l0 = (size_t)c0;
        {
            const int per = (n1l + nth - 1) / nth, lo = min(n1l, tid * per), hi = min(n1l, lo + per);
            int cnt = 0;
            for (int i = lo; i < hi; i++) cnt += (pl.m12[i] >= 0);
            int total;
            int k = block_exclusive_scan(st, cnt, &total);
            for (int i = lo; i < hi; i++) {
                const int i2 = pl.m12[i];
                if (i2 < 0) continue;
                const size_t a = (size_t)(c0 + i);
                f.sX[k] = P.ls_sP[3 * a]; f.sY[k] = P.ls_sP[3 * a + 1]; f.sZ[k] = P.ls_sP[3 * a + 2];
                f.eX[k] = P.ls_eP[3 * a]; f.eY[k] = P.ls_eP[3 * a + 1]; f.eZ[k] = P.ls_eP[3 * a + 2];
                const double* le = C.ls_le + 3 * (size_t)(d0 + i2);    // le_obs = curr le (:175)
                f.l0[k] = le[0]; f.l1[k] = le[1]; f.l2[k] = le[2];
                f.su[k] = P.ls_spl[2 * a]; f.sv[k] = P.ls_spl[2 * a + 1];
                f.eu[k] = P.ls_epl[2 * a]; f.ev[k] = P.ls_epl[2 * a + 1];
                // LineFeature::safeCopy -> ctor re-applies the level rule (src/stereoFeatures.cpp:117-135)
                double s2 = P.ls_sigma2[a];
                const int level = P.ls_level ? P.ls_level[a] : 0;
                for (int l = 0; l < level; l++) s2 *= cfg.lsd_scale;
                f.ls2[k] = 1.0 / (s2 * s2);
                f.inl_l[k] = 1;
                midx_l[k] = (uint16_t)i;
                k++;
            }
            f.nl = total;
        }
        if (tid == 0) {
            st.n_inl_p = f.np;   // f2fTracking: n_inliers_* = list sizes (:126-128)
            st.n_inl_l = f.nl;
        }
    } else {
        // ---- explicit matched lists ----
        const MatchedDev& M = prm.matched;
        const int a0 = M.pt_off[pair], c0 = M.ls_off[pair];
        n1p = f.np = M.pt_off[pair + 1] - a0;
        n1l = f.nl = M.ls_off[pair + 1] - c0;
        out_p0 = (size_t)a0;
        out_l0 = (size_t)c0;
        int cp = 0, cl = 0;
        for (int i = tid; i < f.np; i += nth) {
            const size_t a = (size_t)(a0 + i);
            f.Px[i] = M.pt_P[3 * a]; f.Py[i] = M.pt_P[3 * a + 1]; f.Pz[i] = M.pt_P[3 * a + 2];
            f.pu[i] = M.pt_pl_obs[2 * a]; f.pv[i] = M.pt_pl_obs[2 * a + 1];
            f.ps2[i] = M.pt_sigma2[a];
            const uint8_t in = M.pt_inlier ? (M.pt_inlier[a] != 0) : 1;
            f.inl_p[i] = in;
            cp += in;
        }
        for (int i = tid; i < f.nl; i += nth) {
            const size_t a = (size_t)(c0 + i);
            f.sX[i] = M.ls_sP[3 * a]; f.sY[i] = M.ls_sP[3 * a + 1]; f.sZ[i] = M.ls_sP[3 * a + 2];
            f.eX[i] = M.ls_eP[3 * a]; f.eY[i] = M.ls_eP[3 * a + 1]; f.eZ[i] = M.ls_eP[3 * a + 2];
            f.l0[i] = M.ls_le_obs[3 * a]; f.l1[i] = M.ls_le_obs[3 * a + 1]; f.l2[i] = M.ls_le_obs[3 * a + 2];
            f.su[i] = M.ls_spl[2 * a]; f.sv[i] = M.ls_spl[2 * a + 1];
            f.eu[i] = M.ls_epl[2 * a]; f.ev[i] = M.ls_epl[2 * a + 1];
            f.ls2[i] = M.ls_sigma2[a];
            const uint8_t in = M.ls_inlier ? (M.ls_inlier[a] != 0) : 1;
            f.inl_l[i] = in;
            cl += in;
        }
        // the reference sets n_inliers from the list sizes; explicit flags only matter to the evaluator
        (void)cp; (void)cl;
        if (tid == 0) {
            st.n_inl_p = f.np;
            st.n_inl_l = f.nl;
        }
    }
    __syncthreads();
    if (tid == 0) { st.tc[1] += clock64() - t_ph; }

    // ---- C. optimizePose (:307-392) ----
    const PlPrior* prior = prm.priors ? &prm.priors[pair] : nullptr;
    if (tid == 0) {
        st.out.status = PLSTVO_ST_REFINED;
        st.out.iters_stage1 = st.out.iters_stage2 = 0;
        for (int i = 0; i < 36; i++) st.cov[i] = 0.0;
        st.err = -1.0;
        mat4_identity(st.DT0);
        if (cfg.use_motion_model && prior) {   // :317-324
            for (int i = 0; i < 16; i++) st.DT0[i] = prior->DT[i];
            if (!is_good_solution(st.DT0, prior->DT_cov, prior->err_norm, nullptr)) mat4_identity(st.DT0);
        }
        for (int i = 0; i < 16; i++) st.DT[i] = st.DT0[i];
        st.ctrl = (st.n_inl_p + st.n_inl_l >= cfg.min_features) ? 1 : 0;
    }
    __syncthreads();
    const bool robust_mode = (cfg.solver_mode != 0);
    if (st.ctrl) {   // block-uniform
        __syncthreads();
        gauss_newton(f, st, sortbuf, cam, cfg, cfg.max_iters, robust_mode);     // stage 1 on DT_ = DT (:335-338)
        if (tid == 0) {
            st.out.iters_stage1 = st.evals;
            const long long t_g = clock64();
            st.ctrl = is_good_solution(st.DT, st.cov, st.err, nullptr) ? 1 : 0;   // :341
            st.tc[4] += clock64() - t_g;
        }
        __syncthreads();
        if (st.ctrl) {
            __syncthreads();
            const long long t_o = clock64();
            remove_outliers(f, st, sortbuf, st.DT, cam, cfg);                     // at the stage-1 pose (:343)
            if (tid == 0) {
                st.tc[5] += clock64() - t_o;
                st.ctrl = (st.n_inl_p + st.n_inl_l >= cfg.min_features) ? 1 : 0;
                for (int i = 0; i < 16; i++) st.DT[i] = st.DT0[i];                // stage 2 restarts from DT (:347)
            }
            __syncthreads();
            if (st.ctrl) {
                __syncthreads();
                gauss_newton(f, st, sortbuf, cam, cfg, cfg.max_iters_ref, robust_mode);
                if (tid == 0) st.out.iters_stage2 = st.evals;
            } else {
                if (tid == 0) {
                    mat4_identity(st.DT);                                         // :351-355
                    st.out.status = PLSTVO_ST_FEW_AFTER;
                }
            }
        } else {
            __syncthreads();
            if (tid == 0)
                for (int i = 0; i < 16; i++) st.DT[i] = st.DT0[i];
            __syncthreads();
            gauss_newton(f, st, sortbuf, cam, cfg, cfg.max_iters_ref, true);      // fallback (:357-359)
            if (tid == 0) {
                st.out.iters_stage2 = st.evals;
                st.out.status = PLSTVO_ST_ROBUST_FALLBACK;
            }
        }
    } else {
        if (tid == 0) {
            mat4_identity(st.DT);                                                 // :364-368
            st.out.status = PLSTVO_ST_FEW_BEFORE;
        }
    }
    __syncthreads();

    // ---- pose finalisation (:372-391) ----
    if (tid == 0) {
        const long long t_f = clock64();
        PlPoseResult& o = st.out;
        double Tfw_prev[16], Tfw_cov_prev[36];
        if (prior) {
            for (int i = 0; i < 16; i++) Tfw_prev[i] = prior->Tfw[i];
            for (int i = 0; i < 36; i++) Tfw_cov_prev[i] = prior->Tfw_cov[i];
        } else {   // initialize(): Tfw = I, Tfw_cov = I (:43-44)
            mat4_identity(Tfw_prev);
            for (int i = 0; i < 36; i++) Tfw_cov_prev[i] = (i % 7 == 0) ? 1.0 : 0.0;
        }
        for (int i = 0; i < 16; i++) o.DT_opt[i] = st.DT[i];
        double eig[6];
        if (is_good_solution(st.DT, st.cov, st.err, eig) && !mat4_is_identity(st.DT)) {
            double Ti[16], x[6], T2[16];
            inverse_se3(st.DT, Ti);
            logmap_se3(Ti, x);
            expmap_se3(x, o.DT);                                                  // :374
            for (int i = 0; i < 36; i++) o.DT_cov[i] = st.cov[i];
            o.err_norm = st.err;
            mat4_mul(Tfw_prev, o.DT, T2);
            logmap_se3(T2, x);
            expmap_se3(x, o.Tfw);                                                 // :377
            unccomp_se3(Tfw_prev, Tfw_cov_prev, st.cov, o.Tfw_cov);               // :378
            for (int i = 0; i < 6; i++) o.DT_cov_eig[i] = eig[i];
            o.good = 1;
        } else {
            mat4_identity(o.DT);
            for (int i = 0; i < 36; i++) o.DT_cov[i] = 0.0;
            o.err_norm = -1.0;
            for (int i = 0; i < 16; i++) o.Tfw[i] = Tfw_prev[i];
            for (int i = 0; i < 36; i++) o.Tfw_cov[i] = Tfw_cov_prev[i];
            for (int i = 0; i < 6; i++) o.DT_cov_eig[i] = 0.0;
            o.good = 0;
        }
        o.n_matched_pt = f.np;
        o.n_matched_ls = f.nl;
        o.n_inliers_pt = st.n_inl_p;
        o.n_inliers_ls = st.n_inl_l;
        o.n_inliers = st.n_inl_p + st.n_inl_l;
        o.reserved = 0;
        st.tc[6] += clock64() - t_f;
        if (prm.phase_cycles)
            for (int i = 0; i < 8; i++) prm.phase_cycles[(size_t)pair * 8 + i] = st.tc[i];
    }
    __syncthreads();
    {   // result struct -> HBM, cooperatively (sizeof(PlPoseResult) is a multiple of 8)
        const uint64_t* src = reinterpret_cast<const uint64_t*>(&st.out);
        uint64_t* dst = reinterpret_cast<uint64_t*>(&prm.results[pair]);
        for (int i = tid; i < (int)(sizeof(PlPoseResult) / 8); i += nth) dst[i] = src[i];
    }
    // ---- inlier flags back to the caller's indexing ----
    if (prm.mode == 0) {
        if (prm.inlier_pt) {
            for (int i = tid; i < n1p; i += nth) prm.inlier_pt[out_p0 + i] = 0;
        }
        if (prm.inlier_ls) {
            for (int i = tid; i < n1l; i += nth) prm.inlier_ls[out_l0 + i] = 0;
        }
        __syncthreads();
        if (prm.inlier_pt)
            for (int k = tid; k < f.np; k += nth) prm.inlier_pt[out_p0 + midx_p[k]] = f.inl_p[k];
        if (prm.inlier_ls)
            for (int k = tid; k < f.nl; k += nth) prm.inlier_ls[out_l0 + midx_l[k]] = f.inl_l[k];
    } else {
        if (prm.inlier_pt)
            for (int k = tid; k < f.np; k += nth) prm.inlier_pt[out_p0 + k] = f.inl_p[k];
        if (prm.inlier_ls)
            for (int k = tid; k < f.nl; k += nth) prm.inlier_ls[out_l0 + k] = f.inl_l[k];
    }
}

cudaError_t launch_track_solve(const SolveParams& prm, int n_pairs, cudaStream_t stream) {
    if (n_pairs <= 0) return cudaSuccess;
    const size_t smem = k2_smem_bytes(prm.cap_pt, prm.cap_ls, prm.sort_cap, prm.feat_in_smem != 0);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(track_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    track_solve_kernel<<<n_pairs, K2_THREADS, smem, stream>>>(prm);
    return cudaGetLastError();
}

}  // namespace plstvo
