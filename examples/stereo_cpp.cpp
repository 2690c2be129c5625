// C++ host example: raw stereo features of two consecutive frames -> pose, written against include/plstvo.hpp in the order
// the reference's main loop runs (app/imagesStVO.cpp:88-124 with feature detection replaced by a file):
//   frame = matchStereoPoints + matchStereoLines (StereoFrame::extractStereoFeatures' matching half)
//   handler.initialize(frame0); handler.insertStereoPair(frame1); handler.optimizePose(); handler.needNewKF(); handler.updateFrame()
//   g++ -std=c++17 -Iinclude examples/stereo_cpp.cpp -Lstvo_pl_b200/lib -lplstvo_b200 -o stereo_cpp
#include <cstdio>
#include <fstream>

#include "plstvo.hpp"

template <typename T>
static void read_vec(std::ifstream& f, std::vector<T>& v) {
    int64_t n = 0;
    f.read(reinterpret_cast<char*>(&n), 8);
    v.resize((size_t)n);
    f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * sizeof(T)));
}

struct RawFrame {
    plstvo::KeyPoints pl, pr;
    plstvo::KeyLines ll, lr;
};
static RawFrame read_raw(std::ifstream& f) {
    RawFrame r;
    read_vec(f, r.pl.pt); read_vec(f, r.pl.octave); read_vec(f, r.pl.desc); read_vec(f, r.pr.pt); read_vec(f, r.pr.desc);
    read_vec(f, r.ll.seg); read_vec(f, r.ll.angle); read_vec(f, r.ll.octave); read_vec(f, r.ll.desc); read_vec(f, r.lr.seg);
    read_vec(f, r.lr.desc);
    return r;
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: stereo_cpp <file>\n"); return 2; }
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) return 2;
    PlCamera cam;
    f.read(reinterpret_cast<char*>(&cam), sizeof(cam));
    PlConfig cfg;
    plstvo_kitti_config(&cfg);
    PlStereoMatchConfig mc;
    plstvo_default_stereo_match_config(&mc);
    PlStereoConfig sc;
    plstvo_default_stereo_config(&sc);
    try {
        plstvo::Context ctx;
        plstvo::StereoFrameHandler h(ctx, cam, cfg);
        int counts[4];
        plstvo::StereoFrame frames[2];
        for (int i = 0; i < 2; ++i) {
            const RawFrame raw = read_raw(f);
            counts[2 * i] = plstvo::matchStereoPoints(ctx, cam, mc, sc, raw.pl, raw.pr, frames[i]);
            counts[2 * i + 1] = plstvo::matchStereoLines(ctx, cam, mc, sc, raw.ll, raw.lr, frames[i]);
        }
        h.initialize(std::move(frames[0]));
        h.insertStereoPair(std::move(frames[1]));
        h.optimizePose();
        const bool kf = h.needNewKF();
        const PlPoseResult& r = h.result();
        std::printf("{\"stereo\": [%d, %d, %d, %d], \"good\": %d, \"status\": %d, \"n_matched_pt\": %d, \"n_matched_ls\": %d, "
                    "\"n_inliers\": %d, \"new_kf\": %d, \"DT\": [", counts[0], counts[1], counts[2], counts[3], r.good, r.status,
                    r.n_matched_pt, r.n_matched_ls, h.n_inliers, kf ? 1 : 0);
        for (int i = 0; i < 16; ++i) std::printf("%s%.17g", i ? ", " : "", h.curr_frame.DT[i]);
        std::printf("]}\n");
        h.updateFrame();
        std::fprintf(stderr, "next FAST threshold %d\n", h.orb_fast_th);
    } catch (const plstvo::Error& e) {
        std::fprintf(stderr, "plstvo error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
