// plstvo.hpp — header-only C++17 host layer over the C-ABI (plstvo.h): RAII context plus a StereoFrameHandler with
// the reference's method names (include/stereoFrameHandler.h:41-54) for frames of pre-extracted features.  It has no
// Eigen / OpenCV dependency; the Eigen/OpenCV shim that makes it a drop-in for libstvo.so is in INTEGRATION.md.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "plstvo.h"

namespace plstvo {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

class Context {
public:
    explicit Context(int device = -1) {
        const int rc = plstvo_create(device, &ctx_);
        if (rc != 0) throw Error(rc, "plstvo_create failed: no sm_100 CUDA device (there is no CPU fallback)");
    }
    ~Context() { plstvo_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    PlContext* get() const { return ctx_; }
    int check(int rc) const {   // the reference throws std::runtime_error where the C-ABI returns a negative code
        if (rc < 0) throw Error(rc, plstvo_last_error(ctx_));
        return rc;
    }
private:
    PlContext* ctx_ = nullptr;
};

// One stereo frame's pre-extracted features (StereoFrame::stereo_pt / stereo_ls / pdesc_l / ldesc_l), SoA.
struct StereoFrame {
    std::vector<uint8_t> pdesc, ldesc;                                   // [n][32], [m][32]
    std::vector<double> pt_P, pt_pl, pt_sigma2;                          // [n][3], [n][2], [n]
    std::vector<double> ls_sP, ls_eP, ls_le, ls_spl, ls_epl, ls_sigma2;  // [m][3] x3, [m][2] x2, [m]
    std::vector<int32_t> ls_level;                                       // [m]
    int frame_idx = 0;
    // results published by optimizePose (src/stereoFrameHandler.cpp:372-391), row-major
    std::array<double, 16> Tfw{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, DT = Tfw;
    std::array<double, 36> Tfw_cov{}, DT_cov{};
    std::array<double, 6> DT_cov_eig{};
    double err_norm = -1.0;

    int n_pt() const { return (int)(pdesc.size() / 32); }
    int n_ls() const { return (int)(ldesc.size() / 32); }
    PlFrameBatch view(std::array<int32_t, 2>& pt_off, std::array<int32_t, 2>& ls_off) const {
        pt_off = {0, n_pt()};
        ls_off = {0, n_ls()};
        return PlFrameBatch{1, pt_off.data(), ls_off.data(), pdesc.data(), ldesc.data(), pt_P.data(), pt_pl.data(),
                            pt_sigma2.data(), ls_sP.data(), ls_eP.data(), ls_le.data(), ls_spl.data(), ls_epl.data(),
                            ls_sigma2.data(), ls_level.empty() ? nullptr : ls_level.data()};
    }
};

// matching.h surface (src/matching.cpp:41-91)
inline int matchNNR(Context& c, const std::vector<uint8_t>& d1, const std::vector<uint8_t>& d2, float nnr,
                    std::vector<int>& matches_12) {
    matches_12.assign(d1.size() / 32, -1);
    return c.check(plstvo_match_nnr(c.get(), d1.data(), (int)(d1.size() / 32), d2.data(), (int)(d2.size() / 32), 32, nnr,
                                    matches_12.data()));
}
inline int match(Context& c, const std::vector<uint8_t>& d1, const std::vector<uint8_t>& d2, float nnr,
                 bool best_lr_matches, std::vector<int>& matches_12) {
    matches_12.assign(d1.size() / 32, -1);
    return c.check(plstvo_match(c.get(), d1.data(), (int)(d1.size() / 32), d2.data(), (int)(d2.size() / 32), 32, nnr,
                                best_lr_matches ? 1 : 0, matches_12.data()));
}

// StereoFrame::matchStereoPoints / matchStereoLines (src/stereoFrame.cpp:120-173, :309-398) for one frame of raw stereo
// features: fills the frame's point / line records (and the compacted descriptor rows) exactly like the reference's loops.
struct KeyPoints {                       // cv::KeyPoint fields the step reads
    std::vector<float> pt;               // [n][2]
    std::vector<int32_t> octave;         // [n]
    std::vector<uint8_t> desc;           // [n][32]
    int size() const { return (int)(desc.size() / 32); }
};
struct KeyLines {                        // KeyLine fields the step reads
    std::vector<float> seg;              // [m][4] start / end point
    std::vector<float> angle;            // [m]
    std::vector<int32_t> octave;         // [m]
    std::vector<uint8_t> desc;           // [m][32]
    int size() const { return (int)(desc.size() / 32); }
};
inline int matchStereoPoints(Context& c, const PlCamera& cam, const PlStereoMatchConfig& mc, const PlStereoConfig& sc,
                             const KeyPoints& l, const KeyPoints& r, StereoFrame& out, std::vector<double>* disp = nullptr) {
    const int n = l.size();
    const int32_t lo[2] = {0, n}, ro[2] = {0, r.size()};
    std::vector<double> pl(2 * (size_t)n), d((size_t)n), P(3 * (size_t)n), s2((size_t)n);
    std::vector<int32_t> level((size_t)n), src((size_t)n);
    std::vector<uint8_t> desc(32 * (size_t)n);
    int32_t k = 0;
    c.check(plstvo_match_stereo_points(c.get(), &cam, &mc, &sc, 1, lo, l.pt.data(), l.octave.data(), l.desc.data(), ro, r.pt.data(),
                                       r.desc.data(), nullptr, pl.data(), d.data(), P.data(), s2.data(), level.data(), desc.data(),
                                       src.data(), &k));
    out.pt_pl.assign(pl.begin(), pl.begin() + 2 * (size_t)k);
    out.pt_P.assign(P.begin(), P.begin() + 3 * (size_t)k);
    out.pt_sigma2.assign(s2.begin(), s2.begin() + k);
    out.pdesc.assign(desc.begin(), desc.begin() + 32 * (size_t)k);
    if (disp) disp->assign(d.begin(), d.begin() + k);
    return k;
}
inline int matchStereoLines(Context& c, const PlCamera& cam, const PlStereoMatchConfig& mc, const PlStereoConfig& sc,
                            const KeyLines& l, const KeyLines& r, StereoFrame& out) {
    const int n = l.size();
    const int32_t lo[2] = {0, n}, ro[2] = {0, r.size()};
    const size_t N = (size_t)n;
    std::vector<double> spl(2 * N), epl(2 * N), sd(N), ed(N), sP(3 * N), eP(3 * N), le(3 * N), ang(N), s2(N);
    std::vector<int32_t> level(N), src(N);
    std::vector<uint8_t> desc(32 * N);
    int32_t k = 0;
    c.check(plstvo_match_stereo_lines(c.get(), &cam, &mc, &sc, 1, lo, l.seg.data(), l.angle.data(), l.octave.data(), l.desc.data(), ro,
                                      r.seg.data(), r.desc.data(), nullptr, spl.data(), epl.data(), sd.data(), ed.data(), sP.data(),
                                      eP.data(), le.data(), ang.data(), s2.data(), level.data(), desc.data(), src.data(), &k));
    const size_t K = (size_t)k;
    out.ls_spl.assign(spl.begin(), spl.begin() + 2 * K);
    out.ls_epl.assign(epl.begin(), epl.begin() + 2 * K);
    out.ls_sP.assign(sP.begin(), sP.begin() + 3 * K);
    out.ls_eP.assign(eP.begin(), eP.begin() + 3 * K);
    out.ls_le.assign(le.begin(), le.begin() + 3 * K);
    out.ls_sigma2.assign(s2.begin(), s2.begin() + K);
    out.ls_level.assign(level.begin(), level.begin() + K);
    out.ldesc.assign(desc.begin(), desc.begin() + 32 * K);
    return k;
}

// Host-side SE(3) helpers of the key-frame test (src/auxiliar.cpp:113-122, :143-173, :175-190), row-major arrays.
namespace se3 {
using Mat4 = std::array<double, 16>;
using Mat6 = std::array<double, 36>;
using Vec6 = std::array<double, 6>;

inline Mat4 identity4() { return Mat4{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }
inline Mat4 mul(const Mat4& A, const Mat4& B) {
    Mat4 C{};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += A[4 * i + k] * B[4 * k + j];
            C[4 * i + j] = acc;
        }
    return C;
}
inline Mat4 inverse(const Mat4& T) {                     // inverse_se3: [R^T, -R^T t]
    Mat4 Ti = identity4();
    for (int i = 0; i < 3; ++i) {
        double acc = 0.0;
        for (int k = 0; k < 3; ++k) {
            Ti[4 * i + k] = T[4 * k + i];
            acc += -T[4 * k + i] * T[4 * k + 3];
        }
        Ti[4 * i + 3] = acc;
    }
    return Ti;
}
inline Vec6 logmap(const Mat4& T) {                      // logmap_se3: x = [V^-1 t ; w]
    double cosine = (T[0] + T[5] + T[10] - 1.0) / 2.0;
    cosine = std::min(1.0, std::max(-1.0, cosine));
    const double sine = std::min(1.0, std::sqrt(1.0 - cosine * cosine)), theta = std::acos(cosine);
    double w[3] = {0, 0, 0}, V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (theta > 0.000001) {
        const double k = theta / (2.0 * sine);
        w[0] = k * (T[9] - T[6]);                        // skewcoords(theta (R - R^T) / (2 sine))
        w[1] = k * (T[2] - T[8]);
        w[2] = k * (T[4] - T[1]);
        const double s[9] = {0, -w[2] / theta, w[1] / theta, w[2] / theta, 0, -w[0] / theta, -w[1] / theta, w[0] / theta, 0};
        const double a = (1.0 - cosine) / theta, b = (theta - sine) / theta;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double ss = 0.0;
                for (int m = 0; m < 3; ++m) ss += s[3 * i + m] * s[3 * m + j];
                V[3 * i + j] = (i == j ? 1.0 : 0.0) + s[3 * i + j] * a + ss * b;
            }
    }
    // V^-1 t by cofactors (Eigen's fixed 3x3 inverse)
    const double c00 = V[4] * V[8] - V[5] * V[7], c01 = V[5] * V[6] - V[3] * V[8], c02 = V[3] * V[7] - V[4] * V[6];
    const double det = V[0] * c00 + V[1] * c01 + V[2] * c02;
    const double inv[9] = {c00 / det, (V[2] * V[7] - V[1] * V[8]) / det, (V[1] * V[5] - V[2] * V[4]) / det,
                           c01 / det, (V[0] * V[8] - V[2] * V[6]) / det, (V[2] * V[3] - V[0] * V[5]) / det,
                           c02 / det, (V[1] * V[6] - V[0] * V[7]) / det, (V[0] * V[4] - V[1] * V[3]) / det};
    Vec6 x{};
    for (int i = 0; i < 3; ++i) {
        x[i] = inv[3 * i] * T[3] + inv[3 * i + 1] * T[7] + inv[3 * i + 2] * T[11];
        x[3 + i] = w[i];
    }
    return x;
}
inline Mat6 adjoint(const Mat4& T) {                     // adjoint_se3: [R, skew(t) R; 0, R]
    Mat6 A{};
    const double t[3] = {T[3], T[7], T[11]};
    const double sk[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            A[6 * i + j] = T[4 * i + j];
            A[6 * (i + 3) + j + 3] = T[4 * i + j];
            double acc = 0.0;
            for (int k = 0; k < 3; ++k) acc += sk[3 * i + k] * T[4 * k + j];
            A[6 * i + j + 3] = acc;
        }
    return A;
}
inline Mat6 sandwich(const Mat6& A, const Mat6& C) {     // A C A^T
    Mat6 AC{}, out{};
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 6; ++k) acc += A[6 * i + k] * C[6 * k + j];
            AC[6 * i + j] = acc;
        }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 6; ++k) acc += AC[6 * i + k] * A[6 * j + k];
            out[6 * i + j] = acc;
        }
    return out;
}
inline Mat4 expmap(const Vec6& x) {                      // expmap_se3 (src/auxiliar.cpp:124-141): x = [t ; w]
    const double w[3] = {x[3], x[4], x[5]};
    const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {x[0], x[1], x[2]};
    if (!(theta < 0.000001)) {
        const double s[9] = {0, -w[2] / theta, w[1] / theta, w[2] / theta, 0, -w[0] / theta, -w[1] / theta, w[0] / theta, 0};
        double ss[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double acc = 0.0;
                for (int m = 0; m < 3; ++m) acc += s[3 * i + m] * s[3 * m + j];
                ss[3 * i + j] = acc;
            }
        const double sn = std::sin(theta), cs = std::cos(theta);
        double V[9];
        for (int k = 0; k < 9; ++k) {
            const double id = (k % 4 == 0) ? 1.0 : 0.0;
            R[k] = id + s[k] * sn + ss[k] * (1.0 - cs);
            V[k] = id + s[k] * (1.0 - cs) / theta + ss[k] * (theta - sn) / theta;
        }
        const double t0[3] = {t[0], t[1], t[2]};
        for (int i = 0; i < 3; ++i) t[i] = V[3 * i] * t0[0] + V[3 * i + 1] * t0[1] + V[3 * i + 2] * t0[2];
    }
    return Mat4{R[0], R[1], R[2], t[0], R[3], R[4], R[5], t[1], R[6], R[7], R[8], t[2], 0, 0, 0, 1};
}
inline Mat6 uncTinv(const Mat4& T, const Mat6& cov) { return sandwich(adjoint(inverse(T)), cov); }   // uncTinv_se3
inline double det6(Mat6 A) {                             // Matrix6d::determinant (partial-pivot LU)
    double det = 1.0;
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        for (int r = c + 1; r < 6; ++r)
            if (std::fabs(A[6 * r + c]) > std::fabs(A[6 * piv + c])) piv = r;
        if (A[6 * piv + c] == 0.0) return 0.0;
        if (piv != c) {
            for (int j = 0; j < 6; ++j) std::swap(A[6 * c + j], A[6 * piv + j]);
            det = -det;
        }
        det *= A[6 * c + c];
        for (int r = c + 1; r < 6; ++r) {
            const double f = A[6 * r + c] / A[6 * c + c];
            for (int j = c + 1; j < 6; ++j) A[6 * r + j] -= f * A[6 * c + j];
        }
    }
    return det;
}
}  // namespace se3

// Chains the per-pair results of plstvo_track_stereo_sequence (or of any batch of consecutive pairs solved with identity
// priors) into world poses exactly like optimizePose's tail does frame after frame (src/stereoFrameHandler.cpp:377-378, :388-389):
//   good pair: Tfw_k = expmap(logmap(Tfw_{k-1} DT_k)),  Tfw_cov_k = Tfw_cov_{k-1} + Ad(Tfw_{k-1}) DT_cov_k Ad(Tfw_{k-1})^T
//   failed   : Tfw_k = Tfw_{k-1},                       Tfw_cov_k = Tfw_cov_{k-1}
// Tfw0 / cov0: pose and covariance of the first frame (initialize(): identity, identity).  Writes results[k].Tfw / Tfw_cov.
inline void chainPoses(PlPoseResult* results, int n, se3::Mat4 Tfw0 = se3::identity4(), se3::Mat6 cov0 = [] {
                           se3::Mat6 c{};
                           for (int i = 0; i < 6; ++i) c[7 * i] = 1.0;
                           return c;
                       }()) {
    se3::Mat4 T = Tfw0;
    se3::Mat6 cov = cov0;
    for (int k = 0; k < n; ++k) {
        PlPoseResult& r = results[k];
        if (r.good) {
            se3::Mat4 DT;
            se3::Mat6 dc;
            std::copy(r.DT, r.DT + 16, DT.begin());
            std::copy(r.DT_cov, r.DT_cov + 36, dc.begin());
            const se3::Mat6 add = se3::sandwich(se3::adjoint(T), dc);
            for (int i = 0; i < 36; ++i) cov[i] += add[i];
            T = se3::expmap(se3::logmap(se3::mul(T, DT)));
        }
        std::copy(T.begin(), T.end(), r.Tfw);
        std::copy(cov.begin(), cov.end(), r.Tfw_cov);
    }
}

// The Config values of the handler's host-side state machine (adaptive FAST threshold, key-frame test):
// include/config.h:42-44, :56, :76-80, :100; defaults src/config.cpp:40-42, :52, :72-76, :102.
struct HandlerConfig {
    bool adaptative_fast = true;
    int fast_min_th = 5, fast_max_th = 50, fast_inc_th = 5, fast_feat_th = 50;
    float fast_err_th = 0.5f;
    int orb_fast_th = 20;
    double min_entropy_ratio = 0.85, max_kf_t_dist = 5.0, max_kf_r_dist = 15.0;
};

// StereoFrameHandler::updateFrame's adaptive FAST threshold (src/stereoFrameHandler.cpp:66-86): the threshold the caller's
// keypoint detector should use for the NEXT frame.
inline int updateFastThreshold(const HandlerConfig& c, int orb_fast_th, const StereoFrame& curr, int n_inliers_pt) {
    if (!c.adaptative_fast) return orb_fast_th;
    const int inc = c.fast_inc_th, feat = c.fast_feat_th;
    if (curr.DT == se3::identity4() || curr.err_norm > c.fast_err_th) return std::max(c.fast_min_th, orb_fast_th - 2 * inc);
    if (n_inliers_pt < feat) return std::max(c.fast_min_th, orb_fast_th - 2 * inc);
    if (n_inliers_pt < feat * 2) return std::max(c.fast_min_th, orb_fast_th - inc);
    if (n_inliers_pt > feat * 3) return std::min(c.fast_max_th, orb_fast_th + inc);
    return orb_fast_th;   // the reference's last branch (> 4 feat: + 2 inc, :84-85) is unreachable behind "> 3 feat"
}

// Key-frame test on the accumulated pose entropy (src/stereoFrameHandler.cpp:1136-1218; state include/stereoFrameHandler.h:81-86)
struct KeyframeTest {
    bool prev_f_iskf = true;
    double entropy_first_prevKF = 0.0;
    se3::Mat4 T_prevKF = se3::identity4();
    se3::Mat6 cov_prevKF_currF{};
    int N_prevKF_currF = 0;
    double entropy_curr = 0.0, entropy_ratio = 0.0, kf_t = 0.0, kf_r = 0.0;   // diagnostics of the last test

    bool needNewKF(const HandlerConfig& hcfg_, const StereoFrame& curr_frame) {   // :1136-1187
        const double k_entropy = 3.0 * (1.0 + std::log(2.0 * std::acos(-1.0)));
        if (prev_f_iskf) {                               // :1140-1153
            const double d = se3::det6(curr_frame.DT_cov);
            entropy_first_prevKF = (d != 0.0) ? k_entropy + 0.5 * std::log(d) : -999999999.99;
            prev_f_iskf = false;
        }
        const se3::Vec6 dX = se3::logmap(se3::mul(se3::inverse(curr_frame.Tfw), T_prevKF));   // :1156-1159
        kf_t = std::sqrt(dX[0] * dX[0] + dX[1] * dX[1] + dX[2] * dX[2]);
        kf_r = std::sqrt(dX[3] * dX[3] + dX[4] * dX[4] + dX[5] * dX[5]) * 180.f / 3.1415926535897932384626433832795;
        const se3::Mat6 add = se3::sandwich(se3::adjoint(T_prevKF), se3::uncTinv(curr_frame.DT, curr_frame.DT_cov));   // :1162-1164
        for (int i = 0; i < 36; ++i) cov_prevKF_currF[i] += add[i];
        entropy_curr = k_entropy + 0.5 * std::log(se3::det6(cov_prevKF_currF));
        entropy_ratio = entropy_curr / entropy_first_prevKF;
        bool zero_cov = true;
        for (double v : curr_frame.DT_cov) zero_cov = zero_cov && (v == 0.0);
        if (entropy_ratio < hcfg_.min_entropy_ratio || std::isnan(entropy_ratio) || std::isinf(entropy_ratio) ||
            (zero_cov && curr_frame.DT == se3::identity4()) || kf_t > hcfg_.max_kf_t_dist || kf_r > hcfg_.max_kf_r_dist ||
            N_prevKF_currF > 10)                         // :1173-1175
            return true;
        N_prevKF_currF++;
        return false;
    }
    void currFrameIsKF(StereoFrame& curr_frame) {        // :1189-1218 (feature idx renumbering is the caller's: dense order already)
        curr_frame.Tfw = se3::identity4();
        for (int i = 0; i < 36; ++i) curr_frame.Tfw_cov[i] = (i % 7 == 0) ? 1.0 : 0.0;
        T_prevKF = curr_frame.Tfw;
        cov_prevKF_currF.fill(0.0);
        prev_f_iskf = true;
        N_prevKF_currF = 0;
    }
};

// StereoFrameHandler surface: initialize / insertStereoPair / optimizePose / updateFrame (app/imagesStVO.cpp:88-124)
class StereoFrameHandler {
public:
    StereoFrameHandler(Context& ctx, const PlCamera& cam, const PlConfig& cfg, const HandlerConfig& hcfg = HandlerConfig{})
        : ctx_(ctx), cam_(cam), cfg_(cfg), hcfg_(hcfg) {}

    void initialize(StereoFrame frame) {                 // src/stereoFrameHandler.cpp:35-52
        orb_fast_th = hcfg_.orb_fast_th;                 // :38 (the caller extracts features with this threshold)
        prev_frame = std::move(frame);
        for (int i = 0; i < 36; ++i) prev_frame.Tfw_cov[i] = (i % 7 == 0) ? 1.0 : 0.0;
        kf = KeyframeTest{};                             // :48-51
        has_prev_ = true;
    }
    void insertStereoPair(StereoFrame frame) {           // :54-60 (feature extraction is the caller's)
        if (!has_prev_) throw Error(PLSTVO_E_INVALID, "initialize() first");
        curr_frame = std::move(frame);
        f2fTracking();
    }
    void f2fTracking() {                                 // :106-129, fused with optimizePose on the device
        std::array<int32_t, 2> po1, lo1, po2, lo2;
        const PlFrameBatch prev = prev_frame.view(po1, lo1), curr = curr_frame.view(po2, lo2);
        PlPrior prior;
        std::copy(prev_frame.Tfw.begin(), prev_frame.Tfw.end(), prior.Tfw);
        std::copy(prev_frame.Tfw_cov.begin(), prev_frame.Tfw_cov.end(), prior.Tfw_cov);
        std::copy(prev_frame.DT.begin(), prev_frame.DT.end(), prior.DT);
        std::copy(prev_frame.DT_cov.begin(), prev_frame.DT_cov.end(), prior.DT_cov);
        prior.err_norm = prev_frame.err_norm;
        matches_pt.assign(prev_frame.n_pt(), -1);
        matches_ls.assign(prev_frame.n_ls(), -1);
        inlier_pt.assign(prev_frame.n_pt(), 0);
        inlier_ls.assign(prev_frame.n_ls(), 0);
        ctx_.check(plstvo_track_batch(ctx_.get(), &cam_, &cfg_, &prev, &curr, &prior, &result_, matches_pt.data(),
                                      matches_ls.data(), inlier_pt.data(), inlier_ls.data()));
        n_inliers_pt = result_.n_matched_pt;             // :126-128
        n_inliers_ls = result_.n_matched_ls;
        n_inliers = n_inliers_pt + n_inliers_ls;
    }
    void optimizePose() {                                // :307-392: publish what the device computed
        std::copy(result_.DT, result_.DT + 16, curr_frame.DT.begin());
        std::copy(result_.DT_cov, result_.DT_cov + 36, curr_frame.DT_cov.begin());
        std::copy(result_.DT_cov_eig, result_.DT_cov_eig + 6, curr_frame.DT_cov_eig.begin());
        std::copy(result_.Tfw, result_.Tfw + 16, curr_frame.Tfw.begin());
        std::copy(result_.Tfw_cov, result_.Tfw_cov + 36, curr_frame.Tfw_cov.begin());
        curr_frame.err_norm = result_.err_norm;
        n_inliers_pt = result_.n_inliers_pt;
        n_inliers_ls = result_.n_inliers_ls;
        n_inliers = result_.n_inliers;
    }
    void updateFrame() {                                 // :62-102
        orb_fast_th = updateFastThreshold(hcfg_, orb_fast_th, curr_frame, n_inliers_pt);   // :66-86
        prev_frame = std::move(curr_frame);
        curr_frame = StereoFrame{};
    }

    // key-frame test (:1136-1218), PL-SLAM hooks: state in `kf`
    bool needNewKF() { return kf.needNewKF(hcfg_, curr_frame); }
    void currFrameIsKF() { kf.currFrameIsKF(curr_frame); }
    const PlPoseResult& result() const { return result_; }

    StereoFrame prev_frame, curr_frame;
    std::vector<int32_t> matches_pt, matches_ls;         // prev index -> curr index or -1
    std::vector<uint8_t> inlier_pt, inlier_ls;           // per prev feature: matched and still an inlier
    int n_inliers = 0, n_inliers_pt = 0, n_inliers_ls = 0;
    int orb_fast_th = 20;                                // include/stereoFrameHandler.h:62
    KeyframeTest kf;

private:
    Context& ctx_;
    PlCamera cam_;
    PlConfig cfg_;
    HandlerConfig hcfg_;
    PlPoseResult result_{};
    bool has_prev_ = false;
};

}  // namespace plstvo
