// gn_stream.cu — optimizeFunctions (src/stereoFrameHandler.cpp:549-694) evaluated for a batch of problems with the
// matched lists STREAMED from HBM: the HBM-roofline kernel of config C5 (1920x1080, 8000 points + 2000 lines,
// B >= 1024 problems resident so that one sweep is far larger than L2).
//
// Layout: the fp32-packed records SURVEY 8(a) A4/A5 define — the algorithmic bytes of one evaluation are exactly what
// the kernel reads:
//   point, 32 B : {Px, Py, Pz, sigma2} {u_obs, v_obs, inlier, -}
//   line,  64 B : {sPx, sPy, sPz, sigma2} {ePx, ePy, ePz, inlier} {l0, l1, l2, -} {spl_u, spl_v, epl_u, epl_v}
// (packed once on the device from the double arrays of PlMatchedBatch when the batch is made resident).
// Tiles of 256 records are staged into shared memory by the TMA engine (cp.async.bulk + mbarrier, 4 stages in
// flight per CTA), so HBM latency is decoupled from warp occupancy.  Per-feature arithmetic is fp32 (the survey's
// precision probe: pose deviation 3e-8 rad / 8e-7 m); a thread keeps its <= 64 features' partial sums in fp32, the
// reduction across threads, CTAs and the 10 000 features of a problem runs in fp64 in a fixed order (deterministic).
// The solver proper (solve.cu) stays fp64 end to end; this kernel is the streamed evaluator of the roofline run.
#include <math.h>

#include "common.cuh"

namespace plstvo {

namespace {

constexpr int GS_THREADS = 256;
constexpr int GS_WARPS = GS_THREADS / 32;
constexpr int GS_TILE = 256;                 // records per stage (one per thread)
constexpr int GS_STAGES = 4;
constexpr int GS_STAGE_BYTES = GS_TILE * 64;  // sized for line records

__device__ __forceinline__ void gs_jac(float sc, float gx, float gy, float gz, float dx, float dy, float* J) {
    J[0] = +sc * dx * gz;                                   // :582-587 / :636-641
    J[1] = +sc * dy * gz;
    J[2] = -sc * (gx * dx + gy * dy);
    J[3] = -sc * (gx * gy * dx + gy * gy * dy + gz * gz * dy);
    J[4] = +sc * (gx * gx * dx + gz * gz * dx + gx * gy * dy);
    J[5] = +sc * (gx * gz * dy - gy * gz * dx);
}

__device__ __forceinline__ void gs_acc(float* acc, const float* J, float r, float w) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const float Jw = J[i] * w;
#pragma unroll
        for (int j = i; j < 6; j++) acc[k++] += Jw * J[j];
        acc[21 + i] += Jw * r;
    }
    acc[27] += r * r * w;
    acc[28] += 1.0f;
}

__device__ __forceinline__ float gs_overlap_l(float ls, float le) {
    const float lo = fminf(ls, le), hi = fmaxf(ls, le);
    if (lo < 0.f && hi > 1.f) return 1.f;
    if (hi < 0.f || lo > 1.f) return 0.f;
    if (lo < 0.f) return hi;
    if (hi > 1.f) return 1.f - lo;
    return hi - lo;
}

// StereoFrame::lineSegmentOverlap (src/stereoFrame.cpp:510-616)
__device__ __forceinline__ float gs_overlap(float su, float sv, float eu, float ev, float pu, float pv, float qu,
                                            float qv) {
    const float lx = eu - su, ly = ev - sv;
    if (fabsf(su - eu) < 1.f) {
        const float il = 1.f / ly;
        return gs_overlap_l((pv - sv) * il, (qv - sv) * il);
    }
    const float il = 1.f / lx;
    if (fabsf(sv - ev) < 1.f) return gs_overlap_l((pu - su) * il, (qu - su) * il);
    const float a = sv - ev, b = eu - su, c = su * ev - eu * sv;
    const float lxy = 1.f / (a * a + b * b);
    const float sx = (b * (b * pu - a * pv) - a * c) * lxy;
    const float ex = (b * (b * qu - a * qv) - a * c) * lxy;
    return gs_overlap_l((sx - su) * il, (ex - su) * il);
}

__device__ __forceinline__ double gs_shfl_xor(double v, int m) { return __shfl_xor_sync(0xFFFFFFFFu, v, m); }

// device-side packing of the double arrays into the fp32 records
__global__ void pack_records_kernel(MatchedDev m, int n_pt, int n_ls, float4* __restrict__ pt, float4* __restrict__ ls) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pt) {
        const size_t a = (size_t)i;
        pt[2 * a] = make_float4((float)m.pt_P[3 * a], (float)m.pt_P[3 * a + 1], (float)m.pt_P[3 * a + 2], (float)m.pt_sigma2[a]);
        pt[2 * a + 1] = make_float4((float)m.pt_pl_obs[2 * a], (float)m.pt_pl_obs[2 * a + 1],
                                    (m.pt_inlier && !m.pt_inlier[a]) ? 0.f : 1.f, 0.f);
    }
    if (i < n_ls) {
        const size_t a = (size_t)i;
        ls[4 * a] = make_float4((float)m.ls_sP[3 * a], (float)m.ls_sP[3 * a + 1], (float)m.ls_sP[3 * a + 2], (float)m.ls_sigma2[a]);
        ls[4 * a + 1] = make_float4((float)m.ls_eP[3 * a], (float)m.ls_eP[3 * a + 1], (float)m.ls_eP[3 * a + 2],
                                    (m.ls_inlier && !m.ls_inlier[a]) ? 0.f : 1.f);
        ls[4 * a + 2] = make_float4((float)m.ls_le_obs[3 * a], (float)m.ls_le_obs[3 * a + 1], (float)m.ls_le_obs[3 * a + 2], 0.f);
        ls[4 * a + 3] = make_float4((float)m.ls_spl[2 * a], (float)m.ls_spl[2 * a + 1], (float)m.ls_epl[2 * a], (float)m.ls_epl[2 * a + 1]);
    }
}

__global__ void __launch_bounds__(GS_THREADS, 3)
gn_eval_stream_kernel(PlCamera cam, float homog_th, const int32_t* __restrict__ pt_off, const int32_t* __restrict__ ls_off,
                      const float4* __restrict__ pt, const float4* __restrict__ ls, const double* __restrict__ DTs,
                      double* __restrict__ partial, int bpp) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ double red[GS_WARPS][32];
    __shared__ float sDT[12];
    __shared__ __align__(8) uint64_t bars[GS_STAGES];
    const int prob = blockIdx.x / bpp, blk = blockIdx.x % bpp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 12) sDT[tid] = (float)DTs[(size_t)prob * 16 + tid];
    if (tid == 0) {
        for (int s = 0; s < GS_STAGES; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    float DT[12];
#pragma unroll
    for (int i = 0; i < 12; i++) DT[i] = sDT[i];
    const float fx = (float)cam.fx, fy = (float)cam.fy, cx = (float)cam.cx, cy = (float)cam.cy;

    // this CTA's slice of the problem: whole tiles, points first, then lines
    const int p0 = pt_off[prob], np = pt_off[prob + 1] - p0;
    const int l0 = ls_off[prob], nl = ls_off[prob + 1] - l0;
    const int ptiles = (np + GS_TILE - 1) / GS_TILE, ltiles = (nl + GS_TILE - 1) / GS_TILE;
    const int pt_per = (ptiles + bpp - 1) / bpp, lt_per = (ltiles + bpp - 1) / bpp;
    const int pt_lo = min(ptiles, blk * pt_per), pt_hi = min(ptiles, pt_lo + pt_per);
    const int lt_lo = min(ltiles, blk * lt_per), lt_hi = min(ltiles, lt_lo + lt_per);
    const int n_ptile = pt_hi - pt_lo, n_tiles = n_ptile + (lt_hi - lt_lo);

    auto issue = [&](int t) {   // thread 0: bulk copy of tile t of this CTA into stage t % GS_STAGES
        const int st = t % GS_STAGES;
        const void* src;
        uint32_t bytes;
        if (t < n_ptile) {
            const int f0 = (pt_lo + t) * GS_TILE, cnt = min(GS_TILE, np - f0);
            src = pt + 2 * (size_t)(p0 + f0);
            bytes = (uint32_t)cnt * 32u;
        } else {
            const int f0 = (lt_lo + (t - n_ptile)) * GS_TILE, cnt = min(GS_TILE, nl - f0);
            src = ls + 4 * (size_t)(l0 + f0);
            bytes = (uint32_t)cnt * 64u;
        }
        mbar_arrive_expect_tx(&bars[st], bytes);
        bulk_g2s(smem + (size_t)st * GS_STAGE_BYTES, src, bytes, &bars[st]);
    };
    if (tid == 0)
        for (int t = 0; t < min(GS_STAGES, n_tiles); ++t) issue(t);

    float acc[29];
#pragma unroll
    for (int k = 0; k < 29; k++) acc[k] = 0.f;

    for (int t = 0; t < n_tiles; ++t) {
        const int st = t % GS_STAGES;
        mbar_wait(&bars[st], (uint32_t)((t / GS_STAGES) & 1));
        const float4* s = reinterpret_cast<const float4*>(smem + (size_t)st * GS_STAGE_BYTES);
        if (t < n_ptile) {   // ---- point block (:563-606) ----
            const int f0 = (pt_lo + t) * GS_TILE;
            if (f0 + tid < np) {
                const float4 a = s[2 * tid], b = s[2 * tid + 1];
                if (b.z != 0.f) {
                    const float X = (DT[0] * a.x + DT[1] * a.y + DT[2] * a.z) + DT[3];
                    const float Y = (DT[4] * a.x + DT[5] * a.y + DT[6] * a.z) + DT[7];
                    const float Z = (DT[8] * a.x + DT[9] * a.y + DT[10] * a.z) + DT[11];
                    const float iz = __frcp_rn(Z);
                    const float dx = (cx + fx * X * iz) - b.x, dy = (cy + fy * Y * iz) - b.y;
                    const float n = __fsqrt_rn(dx * dx + dy * dy);
                    const float fgz2 = (Z * Z > homog_th) ? fx * iz * iz : fx / homog_th;
                    float J[6];
                    gs_jac(fgz2 * __frcp_rn(fmaxf(homog_th, n)), X, Y, Z, dx, dy, J);
                    const float r = n * __fsqrt_rn(a.w);
                    gs_acc(acc, J, r, __frcp_rn(1.f + r * r));
                }
            }
        } else {             // ---- line block (:610-684) ----
            const int f0 = (lt_lo + (t - n_ptile)) * GS_TILE;
            if (f0 + tid < nl) {
                const float4 a = s[4 * tid], b = s[4 * tid + 1], c = s[4 * tid + 2], d = s[4 * tid + 3];
                if (b.w != 0.f) {
                    const float sX = (DT[0] * a.x + DT[1] * a.y + DT[2] * a.z) + DT[3];
                    const float sY = (DT[4] * a.x + DT[5] * a.y + DT[6] * a.z) + DT[7];
                    const float sZ = (DT[8] * a.x + DT[9] * a.y + DT[10] * a.z) + DT[11];
                    const float eX = (DT[0] * b.x + DT[1] * b.y + DT[2] * b.z) + DT[3];
                    const float eY = (DT[4] * b.x + DT[5] * b.y + DT[6] * b.z) + DT[7];
                    const float eZ = (DT[8] * b.x + DT[9] * b.y + DT[10] * b.z) + DT[11];
                    const float isz = __frcp_rn(sZ), iez = __frcp_rn(eZ);
                    const float spu = cx + fx * sX * isz, spv = cy + fy * sY * isz;
                    const float epu = cx + fx * eX * iez, epv = cy + fy * eY * iez;
                    const float ds = c.x * spu + c.y * spv + c.z, de = c.x * epu + c.y * epv + c.z;
                    const float n = __fsqrt_rn(ds * ds + de * de);
                    const float iden = __frcp_rn(fmaxf(homog_th, n));
                    float Js[6], Je[6], J[6];
                    gs_jac(((sZ * sZ > homog_th) ? fx * isz * isz : fx / homog_th) * ds * iden, sX, sY, sZ, c.x, c.y, Js);
                    gs_jac(((eZ * eZ > homog_th) ? fx * iez * iez : fx / homog_th) * de * iden, eX, eY, eZ, c.x, c.y, Je);
#pragma unroll
                    for (int k = 0; k < 6; k++) J[k] = Js[k] + Je[k];
                    const float r = n * __fsqrt_rn(a.w);
                    float w = __frcp_rn(1.f + r * r);
                    w *= gs_overlap(d.x, d.y, d.z, d.w, spu, spv, epu, epv);
                    gs_acc(acc, J, r, w);
                }
            }
        }
        __syncthreads();                                   // stage consumed by every warp
        if (tid == 0 && t + GS_STAGES < n_tiles) issue(t + GS_STAGES);
    }

    // fp64 from here on: transposed warp reduction (31 shuffles), then the 8 warps in a fixed order
    double v[32];
#pragma unroll
    for (int k = 0; k < 32; k++) v[k] = (k < 29) ? (double)acc[k] : 0.0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; i++) {
            const double mine = up ? v[i + off] : v[i];
            const double send = up ? v[i] : v[i + off];
            v[i] = mine + gs_shfl_xor(send, off);
        }
    }
    red[warp][lane] = v[0];
    __syncthreads();
    if (tid < 29) {
        double s = 0.0;
#pragma unroll
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

cudaError_t launch_pack_records(const MatchedDev& m, int n_pt, int n_ls, float4* pt, float4* ls, cudaStream_t stream) {
    const int n = n_pt > n_ls ? n_pt : n_ls;
    if (n <= 0) return cudaSuccess;
    pack_records_kernel<<<(n + 255) / 256, 256, 0, stream>>>(m, n_pt, n_ls, pt, ls);
    return cudaGetLastError();
}

cudaError_t launch_gn_eval_stream(const PlCamera& cam, const PlConfig& cfg, const int32_t* pt_off, const int32_t* ls_off,
                                  const float4* pt, const float4* ls, int B, const double* DT, double* partial, int bpp,
                                  double* H, double* g, double* e, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    const size_t smem = (size_t)GS_STAGES * GS_STAGE_BYTES;
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(gn_eval_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        configured = true;
    }
    gn_eval_stream_kernel<<<B * bpp, GS_THREADS, smem, stream>>>(cam, (float)cfg.homog_th, pt_off, ls_off, pt, ls, DT,
                                                                 partial, bpp);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    gn_eval_reduce_kernel<<<B, 32, 0, stream>>>(partial, bpp, H, g, e);
    return cudaGetLastError();
}

}  // namespace plstvo
