// gn_stream.cuh — device-side pieces of the streamed GN evaluation (optimizeFunctions, src/stereoFrameHandler.cpp:549-694, on the
// fp32-packed records): shared by the stand-alone sweep kernel (gn_stream.cu) and the per-problem GN loop of the streamed solver
// (solve.cu).  Internal; see gn_stream.cu for the record layout.
#pragma once
#include <math.h>

#include "common.cuh"

namespace plstvo {

namespace {

constexpr int GS_CONSUMERS = 256;
constexpr int GS_CWARPS = GS_CONSUMERS / 32;
constexpr int GS_THREADS = GS_CONSUMERS + 32;          // + producer warp
constexpr int GS_PT_TILE = 512;                        // points per stage (two per consumer thread)
constexpr int GS_LS_TILE = 256;                        // lines per stage (one per consumer thread)
constexpr int GS_STAGE_BYTES = 16384;
constexpr int GS_STAGES = 6;
constexpr int GS_NACC = ACC_N + 1;                     // 21 H + 6 g + e + count

#ifdef GS_EXACT_MATH   // A/B build: IEEE division / square root instead of the approximate SFU forms
__device__ __forceinline__ float gs_rcp(float x) { return __frcp_rn(x); }
__device__ __forceinline__ float gs_sqrt(float x) { return __fsqrt_rn(x); }
__device__ __forceinline__ float gs_rsqrt(float x) { return __frcp_rn(__fsqrt_rn(x)); }
#else
__device__ __forceinline__ float gs_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float gs_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float gs_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// J (+)= the 1x6 Jacobian row of :582-587 / :636-641 with the scale folded into the direction (dxs, dys)
template <bool ADD>
__device__ __forceinline__ void gs_jac(float X, float Y, float Z, float dxs, float dys, float* J) {
    const float t = fmaf(X, dxs, Y * dys), zz = Z * Z;
    const float j0 = dxs * Z, j1 = dys * Z, j3 = fmaf(Y, t, zz * dys), j4 = fmaf(X, t, zz * dxs);
    const float j5 = Z * fmaf(X, dys, -(Y * dxs));
    if (ADD) { J[0] += j0; J[1] += j1; J[2] -= t; J[3] -= j3; J[4] += j4; J[5] += j5; }
    else     { J[0] = j0;  J[1] = j1;  J[2] = -t; J[3] = -j3; J[4] = j4;  J[5] = j5; }
}

// Per-thread accumulators: 21 unique entries of H = sum w J J^T (row-major upper triangle), 6 of g = sum w r J, e, count.
struct GsAccScalar {
    float a[GS_NACC];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < GS_NACC; i++) a[i] = 0.f;
    }
    __device__ __forceinline__ void add(const float* J, float r, float w, float one) {
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const float Jw = J[i] * w;
#pragma unroll
            for (int j = i; j < 6; j++) a[k] = fmaf(Jw, J[j], a[k]), k++;
            a[21 + i] = fmaf(Jw, r, a[21 + i]);
        }
        a[27] = fmaf(r * w, r, a[27]);
        a[28] += one;
    }
    __device__ __forceinline__ void unpack(float* out) const {
#pragma unroll
        for (int i = 0; i < GS_NACC; i++) out[i] = a[i];
    }
};

// The same sums with Blackwell's packed fp32 pipe (fma.rn.f32x2 -> FFMA2, one operand broadcast): 12 FFMA2 + 3 FMUL2 + 5
// scalar ops per feature instead of 36.  Pairs: (00,01)(02,03)(04,05) (12,13)(14,15) (22,23)(24,25) (34,35) (44,45)
// (g0,g1)(g2,g3)(g4,g5); scalars: 11, 33, 55, e, count.
typedef unsigned long long gs_u64;
__device__ __forceinline__ gs_u64 gs_pk(float lo, float hi) { gs_u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void gs_upk(gs_u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ gs_u64 gs_fma2(gs_u64 a, gs_u64 b, gs_u64 c) { gs_u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ gs_u64 gs_mul2(gs_u64 a, gs_u64 b) { gs_u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
struct GsAccPacked {
    gs_u64 p[12];
    float s[5];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < 12; i++) p[i] = 0ull;
#pragma unroll
        for (int i = 0; i < 5; i++) s[i] = 0.f;
    }
    __device__ __forceinline__ void add(const float* J, float r, float w, float one) {
        const gs_u64 J01 = gs_pk(J[0], J[1]), J23 = gs_pk(J[2], J[3]), J45 = gs_pk(J[4], J[5]), ww = gs_pk(w, w), rr = gs_pk(r, r);
        const gs_u64 W01 = gs_mul2(J01, ww), W23 = gs_mul2(J23, ww), W45 = gs_mul2(J45, ww);
        float w0, w1, w2, w3, w4, w5;
        gs_upk(W01, w0, w1); gs_upk(W23, w2, w3); gs_upk(W45, w4, w5);
        const gs_u64 b0 = gs_pk(w0, w0), b1 = gs_pk(w1, w1), b2 = gs_pk(w2, w2), b3 = gs_pk(w3, w3), b4 = gs_pk(w4, w4);
        p[0] = gs_fma2(b0, J01, p[0]); p[1] = gs_fma2(b0, J23, p[1]); p[2] = gs_fma2(b0, J45, p[2]);
        p[3] = gs_fma2(b1, J23, p[3]); p[4] = gs_fma2(b1, J45, p[4]);
        p[5] = gs_fma2(b2, J23, p[5]); p[6] = gs_fma2(b2, J45, p[6]);
        p[7] = gs_fma2(b3, J45, p[7]);
        p[8] = gs_fma2(b4, J45, p[8]);
        p[9] = gs_fma2(W01, rr, p[9]); p[10] = gs_fma2(W23, rr, p[10]); p[11] = gs_fma2(W45, rr, p[11]);
        s[0] = fmaf(w1, J[1], s[0]); s[1] = fmaf(w3, J[3], s[1]); s[2] = fmaf(w5, J[5], s[2]);
        s[3] = fmaf(r * w, r, s[3]);
        s[4] += one;
    }
    __device__ __forceinline__ void unpack(float* o) const {
        gs_upk(p[0], o[0], o[1]); gs_upk(p[1], o[2], o[3]); gs_upk(p[2], o[4], o[5]);
        o[6] = s[0];
        gs_upk(p[3], o[7], o[8]); gs_upk(p[4], o[9], o[10]);
        gs_upk(p[5], o[11], o[12]); gs_upk(p[6], o[13], o[14]);
        o[15] = s[1];
        gs_upk(p[7], o[16], o[17]);
        gs_upk(p[8], o[18], o[19]);
        o[20] = s[2];
        gs_upk(p[9], o[21], o[22]); gs_upk(p[10], o[23], o[24]); gs_upk(p[11], o[25], o[26]);
        o[27] = s[3];
        o[28] = s[4];
    }
};

__device__ __forceinline__ float gs_overlap_l(float ls, float le) {   // :601-610 on the two parameters
    const float lo = fminf(ls, le), hi = fmaxf(ls, le);
    float ov = fminf(hi, 1.f) - fmaxf(lo, 0.f);                       // covered part of [0, 1]
    ov = (hi < 0.f || lo > 1.f) ? 0.f : ov;
    return ov;
}

// The pose as the evaluator holds it: r[] = the top three rows of DT with the identity taken off the rotation, (R - I | t),
// rounded to fp32 AFTER the subtraction (a frame-to-frame rotation is close to I: its difference from I keeps ~1e-9 of absolute
// accuracy in fp32, R itself only 6e-8).
struct GsPose {
    float r[12];
    float fx, fy, cx, cy, h, inv_h, fx_h;
};
__device__ __forceinline__ float gs_pose_entry(double v, int i) { return (float)((i == 0 || i == 5 || i == 10) ? v - 1.0 : v); }

// ---- records (formed once per list, in double) ------------------------------------------------------------------------------
// Every pose-independent part of a residual is evaluated in fp64 when the record is packed, at the identity pose; the kernel
// adds the pose-dependent CHANGE of the projection, which is a small quantity without cancellation:
//     u(DT) - u_obs = [fx x/z + cx - u_obs]  +  fx (dx - (x/z) dz) / (z + dz),      (dx, dy, dz) = (R - I) P + t
// (the naive fp32 form subtracts two ~1000 px numbers and carries ~2e-4 px of rounding noise into every residual; this form
// carries ~1e-5 px, which matters for the stop tests of the Gauss-Newton loop: they compare errors that differ in the 6th digit).
//   point, 32 B : {x/z, y/z, z, sqrt(sigma2)} {du0, dv0, inlier, -}                du0 = fx x/z + cx - u_obs
//   line,  64 B : {sx/sz, sy/sz, sz, sqrt(sigma2)} {ex/ez, ey/ez, ez, inlier} {l0, l1, ds0, de0} {oa, ob, lam_s0, lam_e0}
//                 ds0 = l . (projection of sP at identity, 1), lam_s0 = (oa, ob, oc) . (that projection, 1)
struct GsCamD {
    double fx, fy, cx, cy;
};
__device__ __forceinline__ void gs_pack_point(const GsCamD& c, double x, double y, double z, double u, double v, double pss,
                                              bool inl, float4& A, float4& B) {
    const double xn = x / z, yn = y / z;
    A = make_float4((float)xn, (float)yn, (float)z, (float)pss);
    B = make_float4((float)(c.fx * xn + c.cx - u), (float)(c.fy * yn + c.cy - v), inl ? 1.f : 0.f, 0.f);
}
__device__ __forceinline__ void gs_pack_line(const GsCamD& c, double sx, double sy, double sz, double ex, double ey, double ez,
                                             double l0, double l1, double l2, double oa, double ob, double oc, double lss,
                                             bool inl, float4& A, float4& B, float4& C, float4& D) {
    const double sxn = sx / sz, syn = sy / sz, exn = ex / ez, eyn = ey / ez;
    const double spu = c.fx * sxn + c.cx, spv = c.fy * syn + c.cy, epu = c.fx * exn + c.cx, epv = c.fy * eyn + c.cy;
    A = make_float4((float)sxn, (float)syn, (float)sz, (float)lss);
    B = make_float4((float)exn, (float)eyn, (float)ez, inl ? 1.f : 0.f);
    C = make_float4((float)l0, (float)l1, (float)(l0 * spu + l1 * spv + l2), (float)(l0 * epu + l1 * epv + l2));
    D = make_float4((float)oa, (float)ob, (float)(oa * spu + ob * spv + oc), (float)(oa * epu + ob * epv + oc));
}

// one 3D point through the pose: camera-frame coordinates and the change of its projection against the identity pose
struct GsProj {
    float X, Y, Z, iz, du, dv;
};
__device__ __forceinline__ GsProj gs_project(const GsPose& P, float xn, float yn, float z, bool use) {
    const float x = xn * z, y = yn * z;
    const float ddx = fmaf(P.r[0], x, fmaf(P.r[1], y, fmaf(P.r[2], z, P.r[3])));
    const float ddy = fmaf(P.r[4], x, fmaf(P.r[5], y, fmaf(P.r[6], z, P.r[7])));
    const float ddz = fmaf(P.r[8], x, fmaf(P.r[9], y, fmaf(P.r[10], z, P.r[11])));
    GsProj q;
    q.X = x + ddx; q.Y = y + ddy; q.Z = z + ddz;
    q.iz = use ? gs_rcp(q.Z) : 0.f;                      // a dead record contributes exact zeros, never a NaN
    q.du = P.fx * fmaf(-xn, ddz, ddx) * q.iz;
    q.dv = P.fy * fmaf(-yn, ddz, ddy) * q.iz;
    return q;
}

// (Measured and dropped: the same per-point arithmetic with TWO points, or the two end points of a segment, in the halves of packed
// f32x2 registers — fewer instructions, but FFMA2 occupies the FMA pipe for two cycles and the pack / unpack moves come on top: the
// sweep kernel fell from 0.82 to 0.71 of the HBM roofline.  Packed registers pay only where the operands are already pairs: the
// accumulation of J J^T above.)
// point block (:563-606); `use` = a live record flagged inlier
template <class Acc>
__device__ __forceinline__ void gs_point(const GsPose& P, const float4 a, const float4 b, bool use, Acc& acc) {
    const GsProj q = gs_project(P, a.x, a.y, a.z, use);
    const float dx = b.x + q.du, dy = b.y + q.dv;
    const float ss = fmaf(dx, dx, dy * dy);
    const float n = gs_sqrt(ss), inv = fminf(P.inv_h, gs_rsqrt(ss));          // 1 / max(homogTh, n)
    const float fg = (q.Z * q.Z > P.h) ? P.fx * q.iz * q.iz : P.fx_h;          // fx / max(homogTh, Z^2)  (:577)
    const float sc = fg * inv;
    float J[6];
    gs_jac<false>(q.X, q.Y, q.Z, sc * dx, sc * dy, J);
    const float r = n * a.w;
    const float w = use ? gs_rcp(fmaf(r, r, 1.f)) : 0.f;
    acc.add(J, r, w, use ? 1.f : 0.f);
}

// line block (:610-684)
template <class Acc>
__device__ __forceinline__ void gs_line(const GsPose& P, const float4 a, const float4 b, const float4 c, const float4 d,
                                        bool use, Acc& acc) {
    const GsProj s = gs_project(P, a.x, a.y, a.z, use), e = gs_project(P, b.x, b.y, b.z, use);
    const float ds = fmaf(c.x, s.du, fmaf(c.y, s.dv, c.z)), de = fmaf(c.x, e.du, fmaf(c.y, e.dv, c.w));
    const float ss = fmaf(ds, ds, de * de);
    const float n = gs_sqrt(ss), iden = fminf(P.inv_h, gs_rsqrt(ss));
    const float ks = ((s.Z * s.Z > P.h) ? P.fx * s.iz * s.iz : P.fx_h) * (ds * iden);
    const float ke = ((e.Z * e.Z > P.h) ? P.fx * e.iz * e.iz : P.fx_h) * (de * iden);
    float J[6];
    gs_jac<false>(s.X, s.Y, s.Z, ks * c.x, ks * c.y, J);
    gs_jac<true>(e.X, e.Y, e.Z, ke * c.x, ke * c.y, J);
    const float r = n * a.w;
    float w = gs_rcp(fmaf(r, r, 1.f));
    w *= gs_overlap_l(fmaf(d.x, s.du, fmaf(d.y, s.dv, d.z)), fmaf(d.x, e.du, fmaf(d.y, e.dv, d.w)));   // :664-670
    acc.add(J, r, use ? w : 0.f, use ? 1.f : 0.f);
}

}  // namespace

}  // namespace plstvo
