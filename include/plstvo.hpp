// plstvo.hpp — header-only C++17 host layer over the C-ABI (plstvo.h): RAII context plus a StereoFrameHandler with
// the reference's method names (include/stereoFrameHandler.h:41-54) for frames of pre-extracted features.  It has no
// Eigen / OpenCV dependency; the Eigen/OpenCV shim that makes it a drop-in for libstvo.so is in INTEGRATION.md.
#pragma once
#include <array>
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

// StereoFrameHandler surface: initialize / insertStereoPair / optimizePose / updateFrame (app/imagesStVO.cpp:88-124)
class StereoFrameHandler {
public:
    StereoFrameHandler(Context& ctx, const PlCamera& cam, const PlConfig& cfg) : ctx_(ctx), cam_(cam), cfg_(cfg) {}

    void initialize(StereoFrame frame) {                 // src/stereoFrameHandler.cpp:35-52
        prev_frame = std::move(frame);
        for (int i = 0; i < 36; ++i) prev_frame.Tfw_cov[i] = (i % 7 == 0) ? 1.0 : 0.0;
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
        prev_frame = std::move(curr_frame);
        curr_frame = StereoFrame{};
    }
    const PlPoseResult& result() const { return result_; }

    StereoFrame prev_frame, curr_frame;
    std::vector<int32_t> matches_pt, matches_ls;         // prev index -> curr index or -1
    std::vector<uint8_t> inlier_pt, inlier_ls;           // per prev feature: matched and still an inlier
    int n_inliers = 0, n_inliers_pt = 0, n_inliers_ls = 0;

private:
    Context& ctx_;
    PlCamera cam_;
    PlConfig cfg_;
    PlPoseResult result_{};
    bool has_prev_ = false;
};

}  // namespace plstvo
