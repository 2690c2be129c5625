// shim/stereoFrameHandler_b200.cpp — drop-in bodies for StereoFrameHandler::f2fTracking and ::optimizePose
// (src/stereoFrameHandler.cpp:106-180, :307-392 of rubengooj/stvo-pl).
//
// The fused call does the matching, the matched_pt / matched_ls construction and the whole of optimizePose on the GPU and
// returns match indices, per-feature inlier flags and the pose record; the shim then repopulates the handler's public state
// exactly as :144-152, :167-179 and :372-391 do.  Matrices cross the C-ABI row-major; element loops do the conversion (Eigen
// is column-major), so the file needs nothing from Eigen beyond operator()(i, j).
//
// Four members are added to StereoFrameHandler (include/stereoFrameHandler.h, after line 85):
//     PlPoseResult b200_result; std::vector<int32_t> b200_m12_pt, b200_m12_ls; std::vector<uint8_t> b200_inl_pt, b200_inl_ls;
// Reference side: compile instead of the two bodies when STVO_WITH_B200 is set.  Here: compiled and run against shim/standin/.
#include <iostream>
#include <iterator>
#include <stdexcept>
#include <vector>

#include "stereoFrameHandler.h"
#include "plstvo.h"

namespace StVO {

PlContext* b200_context();   // shim/matching_b200.cpp

namespace {

// SoA view of StereoFrame::stereo_pt / stereo_ls (include/stereoFrame.h:105-110); lives as long as the call that uses it
struct FramePack {
    std::vector<int32_t> pt_off, ls_off, ls_level;
    std::vector<double> pt_P, pt_pl, pt_s2, ls_sP, ls_eP, ls_le, ls_spl, ls_epl, ls_s2;
    PlFrameBatch view(const StereoFrame* f) {
        const int n = (int)f->stereo_pt.size(), m = (int)f->stereo_ls.size();
        pt_off.assign(2, 0); ls_off.assign(2, 0);
        pt_off[1] = n; ls_off[1] = m;
        pt_P.resize(3 * (size_t)n); pt_pl.resize(2 * (size_t)n); pt_s2.resize(n);
        for (int i = 0; i < n; ++i) {
            const PointFeature* p = f->stereo_pt[i];
            for (int k = 0; k < 3; ++k) pt_P[3 * i + k] = p->P(k);
            pt_pl[2 * i] = p->pl(0); pt_pl[2 * i + 1] = p->pl(1);
            pt_s2[i] = p->sigma2;
        }
        ls_sP.resize(3 * (size_t)m); ls_eP.resize(3 * (size_t)m); ls_le.resize(3 * (size_t)m);
        ls_spl.resize(2 * (size_t)m); ls_epl.resize(2 * (size_t)m); ls_s2.resize(m); ls_level.resize(m);
        for (int i = 0; i < m; ++i) {
            const LineFeature* l = f->stereo_ls[i];
            for (int k = 0; k < 3; ++k) { ls_sP[3 * i + k] = l->sP(k); ls_eP[3 * i + k] = l->eP(k); ls_le[3 * i + k] = l->le(k); }
            for (int k = 0; k < 2; ++k) { ls_spl[2 * i + k] = l->spl(k); ls_epl[2 * i + k] = l->epl(k); }
            ls_s2[i] = l->sigma2;
            ls_level[i] = l->level;
        }
        PlFrameBatch b;
        b.B = 1;
        b.pt_off = pt_off.data(); b.ls_off = ls_off.data();
        b.pdesc = f->pdesc_l.ptr<uint8_t>(); b.ldesc = f->ldesc_l.ptr<uint8_t>();
        b.pt_P = pt_P.data(); b.pt_pl = pt_pl.data(); b.pt_sigma2 = pt_s2.data();
        b.ls_sP = ls_sP.data(); b.ls_eP = ls_eP.data(); b.ls_le = ls_le.data();
        b.ls_spl = ls_spl.data(); b.ls_epl = ls_epl.data(); b.ls_sigma2 = ls_s2.data(); b.ls_level = ls_level.data();
        return b;
    }
};

PlConfig b200_config() {                           // the ~15 Config values the path reads (include/config.h:39-105)
    PlConfig c;
    c.has_points = Config::hasPoints(); c.has_lines = Config::hasLines();
    c.best_lr_matches = Config::bestLRMatches(); c.use_motion_model = Config::useMotionModel();
    c.min_features = Config::minFeatures(); c.max_iters = Config::maxIters(); c.max_iters_ref = Config::maxItersRef();
    c.solver_mode = 0;                             // `mode` is hard-wired to 0, src/stereoFrameHandler.cpp:329
    c.min_ratio_12_p = Config::minRatio12P(); c.min_ratio_12_l = Config::minRatio12L();
    c.homog_th = Config::homogTh(); c.min_error = Config::minError(); c.min_error_change = Config::minErrorChange();
    c.inlier_k = Config::inlierK(); c.lsd_scale = Config::lsdScale();
    return c;
}

template <class M> void to_rows(const M& m, int n, double* out) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) out[n * i + j] = m(i, j);
}
template <class M> void from_rows(const double* in, int n, M& m) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) m(i, j) = in[n * i + j];
}

}  // namespace

// f2fTracking() keeps its signature; the GPU result is cached for the optimizePose() that follows it
void StereoFrameHandler::f2fTracking() {
    matched_pt.clear();
    matched_ls.clear();
    FramePack pack_prev, pack_curr;                // local: nothing outlives the call
    const PlFrameBatch prev = pack_prev.view(prev_frame), curr = pack_curr.view(curr_frame);
    const PlCamera pcam{cam->getFx(), cam->getFy(), cam->getCx(), cam->getCy(), cam->getB(), cam->getWidth(), cam->getHeight()};
    const PlConfig cfg = b200_config();
    PlPrior prior;                                 // prev_frame state read by optimizePose (:317-326, :377-378)
    to_rows(prev_frame->Tfw, 4, prior.Tfw);
    to_rows(prev_frame->Tfw_cov, 6, prior.Tfw_cov);
    to_rows(prev_frame->DT, 4, prior.DT);
    to_rows(prev_frame->DT_cov, 6, prior.DT_cov);
    prior.err_norm = prev_frame->err_norm;
    b200_m12_pt.assign(prev.pt_off[1], -1); b200_m12_ls.assign(prev.ls_off[1], -1);
    b200_inl_pt.assign(prev.pt_off[1], 0);  b200_inl_ls.assign(prev.ls_off[1], 0);
    if (plstvo_track_batch(b200_context(), &pcam, &cfg, &prev, &curr, &prior, &b200_result, b200_m12_pt.data(),
                           b200_m12_ls.data(), b200_inl_pt.data(), b200_inl_ls.data()) < 0)
        throw std::runtime_error(plstvo_last_error(b200_context()));
    // rebuild matched_pt / matched_ls exactly as :144-152 and :167-179 do (ascending i1)
    for (size_t i1 = 0; i1 < b200_m12_pt.size(); ++i1) {
        const int i2 = b200_m12_pt[i1];
        if (i2 < 0) continue;
        prev_frame->stereo_pt[i1]->pl_obs = curr_frame->stereo_pt[i2]->pl;
        prev_frame->stereo_pt[i1]->inlier = true;
        matched_pt.push_back(prev_frame->stereo_pt[i1]->safeCopy());
        curr_frame->stereo_pt[i2]->idx = prev_frame->stereo_pt[i1]->idx;
    }
    for (size_t i1 = 0; i1 < b200_m12_ls.size(); ++i1) {
        const int i2 = b200_m12_ls[i1];
        if (i2 < 0) continue;
        LineFeature* a = prev_frame->stereo_ls[i1];
        const LineFeature* b = curr_frame->stereo_ls[i2];
        a->sdisp_obs = b->sdisp; a->edisp_obs = b->edisp; a->spl_obs = b->spl; a->epl_obs = b->epl; a->le_obs = b->le;
        a->inlier = true;
        matched_ls.push_back(a->safeCopy());
        curr_frame->stereo_ls[i2]->idx = a->idx;
    }
    n_inliers_pt = (int)matched_pt.size();
    n_inliers_ls = (int)matched_ls.size();
    n_inliers = n_inliers_pt + n_inliers_ls;
}

void StereoFrameHandler::optimizePose() {          // :307-392 — everything was computed on the device by f2fTracking()
    const PlPoseResult& r = b200_result;
    {   // per-feature inlier flags back into the lists (the viewer and PL-SLAM read them)
        std::list<PointFeature*>::iterator it = matched_pt.begin();
        for (size_t i1 = 0; i1 < b200_m12_pt.size() && it != matched_pt.end(); ++i1)
            if (b200_m12_pt[i1] >= 0) { (*it)->inlier = b200_inl_pt[i1] != 0; ++it; }
        std::list<LineFeature*>::iterator jt = matched_ls.begin();
        for (size_t i1 = 0; i1 < b200_m12_ls.size() && jt != matched_ls.end(); ++i1)
            if (b200_m12_ls[i1] >= 0) { (*jt)->inlier = b200_inl_ls[i1] != 0; ++jt; }
    }
    n_inliers_pt = r.n_inliers_pt; n_inliers_ls = r.n_inliers_ls; n_inliers = r.n_inliers;
    from_rows(r.DT, 4, curr_frame->DT);                                            // :374 / :385
    from_rows(r.DT_cov, 6, curr_frame->DT_cov);                                    // :375 / :386
    curr_frame->err_norm = r.err_norm;                                             // :376 / :387
    from_rows(r.Tfw, 4, curr_frame->Tfw);                                          // :377 / :388
    from_rows(r.Tfw_cov, 6, curr_frame->Tfw_cov);                                  // :378 / :389
    for (int i = 0; i < 6; ++i) curr_frame->DT_cov_eig(i) = r.DT_cov_eig[i];       // :380 / :390
    if (r.status == PLSTVO_ST_FEW_BEFORE) std::cout << "[StVO] not enough inliers (before optimization)" << std::endl;   // :367
    if (r.status == PLSTVO_ST_FEW_AFTER) std::cout << "[StVO] not enough inliers (after removal)" << std::endl;          // :354
}

}  // namespace StVO
