// shim_driver.cpp — TEST INFRASTRUCTURE: builds the reference's objects (stand-in classes) from plain arrays, drives the shim
// exactly like app/imagesStVO.cpp:96-97 does (f2fTracking() inside insertStereoPair, then optimizePose()), and hands the
// handler's public state back.  tests/test_shim.py compares it with plstvo_track_batch called directly.
#include <cstring>
#include "matching.h"
#include "stereoFrameHandler.h"
using namespace StVO;

static StereoFrame* make_frame(const PlFrameBatch* f) {
    StereoFrame* s = new StereoFrame;
    const int n = f->pt_off[1], m = f->ls_off[1];
    s->pdesc_l = cv::Mat(n, 32, CV_8UC1);
    s->ldesc_l = cv::Mat(m, 32, CV_8UC1);
    if (n) memcpy(s->pdesc_l.ptr<uint8_t>(), f->pdesc, (size_t)n * 32);
    if (m) memcpy(s->ldesc_l.ptr<uint8_t>(), f->ldesc, (size_t)m * 32);
    for (int i = 0; i < n; ++i) {
        PointFeature* p = new PointFeature;
        if (f->pt_P) for (int k = 0; k < 3; ++k) p->P(k) = f->pt_P[3 * i + k];
        if (f->pt_pl) for (int k = 0; k < 2; ++k) p->pl(k) = f->pt_pl[2 * i + k];
        if (f->pt_sigma2) p->sigma2 = f->pt_sigma2[i];
        p->idx = i;
        p->covP_an = Matrix3d::Zero();
        s->stereo_pt.push_back(p);
    }
    for (int i = 0; i < m; ++i) {
        LineFeature* l = new LineFeature;
        for (int k = 0; k < 3; ++k) {
            if (f->ls_sP) l->sP(k) = f->ls_sP[3 * i + k];
            if (f->ls_eP) l->eP(k) = f->ls_eP[3 * i + k];
            if (f->ls_le) l->le(k) = f->ls_le[3 * i + k];
        }
        for (int k = 0; k < 2; ++k) {
            if (f->ls_spl) l->spl(k) = f->ls_spl[2 * i + k];
            if (f->ls_epl) l->epl(k) = f->ls_epl[2 * i + k];
        }
        if (f->ls_sigma2) l->sigma2 = f->ls_sigma2[i];
        if (f->ls_level) l->level = f->ls_level[i];
        l->idx = i;
        s->stereo_ls.push_back(l);
    }
    s->Tfw = Matrix4d::Identity();
    s->Tfw_cov = Matrix6d::Identity();
    s->DT = Matrix4d::Identity();
    s->DT_cov = Matrix6d::Zero();
    return s;
}

extern "C" int shim_track_pair(const PlCamera* c, const PlConfig* cfg, const PlFrameBatch* prev, const PlFrameBatch* curr,
                               double DT[16], double DT_cov[36], double* err_norm, double Tfw[16], int32_t n_inliers[3],
                               int32_t* n_matched_pt, int32_t* n_matched_ls, uint8_t* inl_pt /* matched_pt order */,
                               uint8_t* inl_ls) {
    Config::hasPoints() = cfg->has_points; Config::hasLines() = cfg->has_lines; Config::bestLRMatches() = cfg->best_lr_matches;
    Config::useMotionModel() = cfg->use_motion_model; Config::minFeatures() = cfg->min_features; Config::maxIters() = cfg->max_iters;
    Config::maxItersRef() = cfg->max_iters_ref; Config::minRatio12P() = cfg->min_ratio_12_p; Config::minRatio12L() = cfg->min_ratio_12_l;
    Config::homogTh() = cfg->homog_th; Config::minError() = cfg->min_error; Config::minErrorChange() = cfg->min_error_change;
    Config::inlierK() = cfg->inlier_k; Config::lsdScale() = cfg->lsd_scale;
    PinholeStereoCamera cam;
    cam.fx = c->fx; cam.fy = c->fy; cam.cx = c->cx; cam.cy = c->cy; cam.b = c->b; cam.width = c->width; cam.height = c->height;
    StereoFrameHandler h;
    h.cam = &cam;
    h.prev_frame = make_frame(prev);
    h.curr_frame = make_frame(curr);
    try {
        h.f2fTracking();      // src/stereoFrameHandler.cpp:59 (inside insertStereoPair)
        h.optimizePose();     // app/imagesStVO.cpp:97
    } catch (const std::exception&) {
        return -1;
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { DT[4 * i + j] = h.curr_frame->DT(i, j); Tfw[4 * i + j] = h.curr_frame->Tfw(i, j); }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) DT_cov[6 * i + j] = h.curr_frame->DT_cov(i, j);
    *err_norm = h.curr_frame->err_norm;
    n_inliers[0] = h.n_inliers_pt; n_inliers[1] = h.n_inliers_ls; n_inliers[2] = h.n_inliers;
    *n_matched_pt = (int)h.matched_pt.size();
    *n_matched_ls = (int)h.matched_ls.size();
    int k = 0;
    for (auto* p : h.matched_pt) inl_pt[k++] = p->inlier;
    k = 0;
    for (auto* l : h.matched_ls) inl_ls[k++] = l->inlier;
    return 0;
}

extern "C" int shim_match(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int mutual, int32_t* m12) {
    cv::Mat a(n1, 32, CV_8UC1), b(n2, 32, CV_8UC1);
    if (n1) memcpy(a.ptr<uint8_t>(), d1, (size_t)n1 * 32);
    if (n2) memcpy(b.ptr<uint8_t>(), d2, (size_t)n2 * 32);
    std::vector<int> m;
    int n;
    try {
        n = mutual ? StVO::match(a, b, nnr, m) : StVO::matchNNR(a, b, nnr, m);
    } catch (const std::exception&) {
        return -1;
    }
    for (int i = 0; i < n1; ++i) m12[i] = m[i];
    return n;
}
