// Stand-in for include/stereoFrameHandler.h + stereoFrame.h + stereoFeatures.h + pinholeStereoCamera.h: the members the
// shim reads and writes, with the reference's names and types; safeCopy() restates src/stereoFeatures.cpp:66-68, :117-135
// (the line constructor RE-APPLIES the level -> sigma2 rule).  TEST INFRASTRUCTURE, like oracle/ref_shim/.
#pragma once
#include <list>
#include <vector>
#include <opencv2/core.hpp>
#include "../../oracle/ref_shim/eigen_standin.h"
#include "../../include/plstvo.h"
#include "config.h"
using namespace Eigen;
typedef Matrix<double, 6, 1> Vector6d;
typedef Matrix<double, 6, 6> Matrix6d;

class PinholeStereoCamera {
public:
    int width, height;
    double fx, fy, cx, cy, b;
    inline const int getWidth() const { return width; }
    inline const int getHeight() const { return height; }
    inline const double getB() const { return b; }
    inline const double getFx() const { return fx; }
    inline const double getFy() const { return fy; }
    inline const double getCx() const { return cx; }
    inline const double getCy() const { return cy; }
};

namespace StVO {

class PointFeature {
public:
    PointFeature() {}
    PointFeature(Vector2d pl_, double disp_, Vector3d P_, Vector2d pl_obs_, int idx_, int level_, double sigma2_, Matrix3d covP_an_,
                 bool inlier_)
        : idx(idx_), pl(pl_), pl_obs(pl_obs_), disp(disp_), P(P_), inlier(inlier_), level(level_), sigma2(sigma2_), covP_an(covP_an_) {}
    PointFeature* safeCopy() { return new PointFeature(pl, disp, P, pl_obs, idx, level, sigma2, covP_an, inlier); }   // :66-68
    int idx = -1;
    Vector2d pl, pl_obs;
    double disp = 0.0;
    Vector3d P;
    bool inlier = true;
    int level = 0;
    double sigma2 = 1.0;
    Matrix3d covP_an;
};

class LineFeature {
public:
    LineFeature() {}
    LineFeature(Vector2d spl_, double sdisp_, Vector3d sP_, Vector2d spl_obs_, double sdisp_obs_, Vector2d epl_, double edisp_,
                Vector3d eP_, Vector2d epl_obs_, double edisp_obs_, Vector3d le_, Vector3d le_obs_, double angle_, int idx_, int level_,
                bool inlier_, double sigma2_, Matrix3d covE_an_, Matrix3d covS_an_)
        : idx(idx_), spl(spl_), epl(epl_), spl_obs(spl_obs_), epl_obs(epl_obs_), sdisp(sdisp_), edisp(edisp_), angle(angle_),
          sdisp_obs(sdisp_obs_), edisp_obs(edisp_obs_), sP(sP_), eP(eP_), le(le_), le_obs(le_obs_), inlier(inlier_), level(level_),
          sigma2(sigma2_), covE_an(covE_an_), covS_an(covS_an_) {
        for (int i = 0; i < level; i++) sigma2 *= Config::lsdScale();   // :126-128: re-applied on every copy
        sigma2 = 1.f / (sigma2 * sigma2);
    }
    LineFeature* safeCopy() {                                            // :131-135
        return new LineFeature(spl, sdisp, sP, spl_obs, sdisp_obs, epl, edisp, eP, epl_obs, edisp_obs, le, le_obs, angle, idx, level,
                               inlier, sigma2, covE_an, covS_an);
    }
    int idx = -1;
    Vector2d spl, epl, spl_obs, epl_obs;
    double sdisp = 0.0, edisp = 0.0, angle = 0.0, sdisp_obs = 0.0, edisp_obs = 0.0;
    Vector3d sP, eP;
    Vector3d le, le_obs;
    bool inlier = true;
    int level = 0;
    double sigma2 = 1.0;
    Matrix3d covE_an, covS_an;
};

class StereoFrame {
public:
    std::vector<PointFeature*> stereo_pt;
    std::vector<LineFeature*> stereo_ls;
    cv::Mat pdesc_l, pdesc_r, ldesc_l, ldesc_r;
    Matrix4d Tfw, DT;
    Matrix6d Tfw_cov, DT_cov;
    Vector6d DT_cov_eig;
    double err_norm = -1.0;
};

class StereoFrameHandler {
public:
    void f2fTracking();
    void optimizePose();
    std::list<PointFeature*> matched_pt;
    std::list<LineFeature*> matched_ls;
    StereoFrame* prev_frame = nullptr;
    StereoFrame* curr_frame = nullptr;
    PinholeStereoCamera* cam = nullptr;
    int n_inliers = 0, n_inliers_pt = 0, n_inliers_ls = 0;
    // the four members the shim adds (include/stereoFrameHandler.h, after line 85)
    PlPoseResult b200_result;
    std::vector<int32_t> b200_m12_pt, b200_m12_ls;
    std::vector<uint8_t> b200_inl_pt, b200_inl_ls;
};

}  // namespace StVO
