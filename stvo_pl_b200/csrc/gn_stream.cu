// gn_stream.cu — optimizeFunctions (src/stereoFrameHandler.cpp:549-694) evaluated for a batch of problems with the
// matched lists STREAMED from HBM (no shared-memory residency): the HBM-roofline kernel of config C5
// (1920x1080, 8000 points + 2000 lines, B >= 1024 problems resident so that one sweep is far larger than L2).
// grid = B x blocks_per_problem; every CTA reduces its slice of features to 28 + 1 partial sums, a second tiny
// kernel adds the partials of a problem in a fixed order (deterministic) and unpacks H (6x6), g (6), e.
#include <math.h>

#include "common.cuh"

namespace plstvo {

namespace {

constexpr int GS_THREADS = 256;
constexpr int GS_WARPS = GS_THREADS / 32;

__device__ __forceinline__ void gs_jac(double fgz2, double gx, double gy, double gz, double dx, double dy, double* J) {
    J[0] = +fgz2 * dx * gz;
    J[1] = +fgz2 * dy * gz;
    J[2] = -fgz2 * (gx * dx + gy * dy);
    J[3] = -fgz2 * (gx * gy * dx + gy * gy * dy + gz * gz * dy);
    J[4] = +fgz2 * (gx * gx * dx + gz * gz * dx + gx * gy * dy);
    J[5] = +fgz2 * (gx * gz * dy - gy * gz * dx);
}

__device__ __forceinline__ void gs_acc(double* acc, const double* J, double r, double w) {
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

__device__ __forceinline__ double gs_overlap_l(double ls, double le) {
    const double lo = (le < ls) ? le : ls, hi = (ls < le) ? le : ls;
    if (lo < 0.0 && hi > 1.0) return 1.0;
    if (hi < 0.0 || lo > 1.0) return 0.0;
    if (lo < 0.0) return hi;
    if (hi > 1.0) return 1.0 - lo;
    return hi - lo;
}

__device__ __forceinline__ double gs_overlap(double su, double sv, double eu, double ev, double pu, double pv,
                                             double qu, double qv) {   // src/stereoFrame.cpp:510-616
    const double lx = eu - su, ly = ev - sv;
    if (fabs(su - eu) < 1.0) return gs_overlap_l((pv - sv) / ly, (qv - sv) / ly);
    if (fabs(sv - ev) < 1.0) return gs_overlap_l((pu - su) / lx, (qu - su) / lx);
    const double a = sv - ev, b = eu - su, c = su * ev - eu * sv;
    const double lxy = 1.0 / (a * a + b * b);
    const double sx = (b * (b * pu - a * pv) - a * c) * lxy;
    const double ex = (b * (b * qu - a * qv) - a * c) * lxy;
    return gs_overlap_l((sx - su) / lx, (ex - su) / lx);
}

__global__ void __launch_bounds__(GS_THREADS) gn_eval_stream_kernel(PlCamera cam, double homog_th, MatchedDev m,
                                                                    const double* __restrict__ DTs,
                                                                    double* __restrict__ partial, int bpp) {
    __shared__ double red[GS_WARPS][ACC_N + 1];
    __shared__ double sDT[12];
    const int prob = blockIdx.x / bpp, blk = blockIdx.x % bpp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 12) sDT[tid] = DTs[(size_t)prob * 16 + tid];
    __syncthreads();
    double DT[12];
#pragma unroll
    for (int i = 0; i < 12; i++) DT[i] = sDT[i];

    double acc[ACC_N + 1];
#pragma unroll
    for (int k = 0; k <= ACC_N; k++) acc[k] = 0.0;

    const int p0 = m.pt_off[prob], np = m.pt_off[prob + 1] - p0;
    const int l0 = m.ls_off[prob], nl = m.ls_off[prob + 1] - l0;
    const int pchunk = (np + bpp - 1) / bpp, lchunk = (nl + bpp - 1) / bpp;
    const int pbeg = min(np, blk * pchunk), pend = min(np, pbeg + pchunk);
    const int lbeg = min(nl, blk * lchunk), lend = min(nl, lbeg + lchunk);

    for (int i = pbeg + tid; i < pend; i += GS_THREADS) {
        const size_t a = (size_t)(p0 + i);
        if (m.pt_inlier && !m.pt_inlier[a]) continue;
        const double x = __ldg(m.pt_P + 3 * a), y = __ldg(m.pt_P + 3 * a + 1), z = __ldg(m.pt_P + 3 * a + 2);
        const double X = (DT[0] * x + DT[1] * y + DT[2] * z) + DT[3];
        const double Y = (DT[4] * x + DT[5] * y + DT[6] * z) + DT[7];
        const double Z = (DT[8] * x + DT[9] * y + DT[10] * z) + DT[11];
        const double dx = (cam.cx + cam.fx * X / Z) - __ldg(m.pt_pl_obs + 2 * a);
        const double dy = (cam.cy + cam.fy * Y / Z) - __ldg(m.pt_pl_obs + 2 * a + 1);
        const double n = sqrt(dx * dx + dy * dy);
        double J[6];
        gs_jac(cam.fx / fmax(homog_th, Z * Z), X, Y, Z, dx, dy, J);
        const double den = fmax(homog_th, n);
#pragma unroll
        for (int k = 0; k < 6; k++) J[k] = J[k] / den;
        const double r = n * sqrt(__ldg(m.pt_sigma2 + a));
        gs_acc(acc, J, r, 1.0 / (1.0 + r * r));
    }
    for (int i = lbeg + tid; i < lend; i += GS_THREADS) {
        const size_t a = (size_t)(l0 + i);
        if (m.ls_inlier && !m.ls_inlier[a]) continue;
        double P[2][3], uv[2][2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const double* src = (e ? m.ls_eP : m.ls_sP) + 3 * a;
            const double x = __ldg(src), y = __ldg(src + 1), z = __ldg(src + 2);
            P[e][0] = (DT[0] * x + DT[1] * y + DT[2] * z) + DT[3];
            P[e][1] = (DT[4] * x + DT[5] * y + DT[6] * z) + DT[7];
            P[e][2] = (DT[8] * x + DT[9] * y + DT[10] * z) + DT[11];
            uv[e][0] = cam.cx + cam.fx * P[e][0] / P[e][2];
            uv[e][1] = cam.cy + cam.fy * P[e][1] / P[e][2];
        }
        const double lx = __ldg(m.ls_le_obs + 3 * a), ly = __ldg(m.ls_le_obs + 3 * a + 1), lc = __ldg(m.ls_le_obs + 3 * a + 2);
        const double ds = lx * uv[0][0] + ly * uv[0][1] + lc, de = lx * uv[1][0] + ly * uv[1][1] + lc;
        const double n = sqrt(ds * ds + de * de);
        double Js[6], Je[6], J[6];
        gs_jac(cam.fx / fmax(homog_th, P[0][2] * P[0][2]), P[0][0], P[0][1], P[0][2], lx, ly, Js);
        gs_jac(cam.fx / fmax(homog_th, P[1][2] * P[1][2]), P[1][0], P[1][1], P[1][2], lx, ly, Je);
        const double den = fmax(homog_th, n);
#pragma unroll
        for (int k = 0; k < 6; k++) J[k] = (Js[k] * ds + Je[k] * de) / den;
        const double r = n * sqrt(__ldg(m.ls_sigma2 + a));
        double w = 1.0 / (1.0 + r * r);
        w *= gs_overlap(__ldg(m.ls_spl + 2 * a), __ldg(m.ls_spl + 2 * a + 1), __ldg(m.ls_epl + 2 * a),
                        __ldg(m.ls_epl + 2 * a + 1), uv[0][0], uv[0][1], uv[1][0], uv[1][1]);
        gs_acc(acc, J, r, w);
    }
#pragma unroll
    for (int k = 0; k <= ACC_N; k++) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (tid <= ACC_N) {
        double s = 0.0;
        for (int w = 0; w < GS_WARPS; w++) s += red[w][tid];
        partial[((size_t)prob * bpp + blk) * (ACC_N + 1) + tid] = s;
    }
}

__global__ void gn_eval_reduce_kernel(const double* __restrict__ partial, int bpp, double* __restrict__ H,
                                      double* __restrict__ g, double* __restrict__ e) {
    __shared__ double s[ACC_N + 1];
    const int prob = blockIdx.x, tid = threadIdx.x;
    if (tid <= ACC_N) {
        double v = 0.0;
        for (int b = 0; b < bpp; b++) v += partial[((size_t)prob * bpp + b) * (ACC_N + 1) + tid];
        s[tid] = v;
    }
    __syncthreads();
    if (tid == 0) {
        int k = 0;
        for (int i = 0; i < 6; i++)
            for (int j = i; j < 6; j++) {
                H[(size_t)prob * 36 + i * 6 + j] = s[k];
                H[(size_t)prob * 36 + j * 6 + i] = s[k];
                k++;
            }
        for (int i = 0; i < 6; i++) g[(size_t)prob * 6 + i] = s[21 + i];
        e[prob] = s[27] / s[28];
    }
}

}  // namespace

cudaError_t launch_gn_eval_stream(const PlCamera& cam, const PlConfig& cfg, const MatchedDev& m, int B,
                                  const double* DT, double* partial, int bpp, double* H, double* g, double* e,
                                  cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    gn_eval_stream_kernel<<<B * bpp, GS_THREADS, 0, stream>>>(cam, cfg.homog_th, m, DT, partial, bpp);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    gn_eval_reduce_kernel<<<B, 32, 0, stream>>>(partial, bpp, H, g, e);
    return cudaGetLastError();
}

}  // namespace plstvo
