// C++ host example / link check: a consumer of libplstvo_b200.so written against include/plstvo.hpp.
// Reads one frame pair from a flat binary file written by tests (or makes a trivial one), runs the reference's call
// sequence initialize -> insertStereoPair -> optimizePose -> updateFrame, prints the pose record as JSON.
//   g++ -std=c++17 -Iinclude examples/track_cpp.cpp -Lstvo_pl_b200/lib -lplstvo_b200 -o track_cpp
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "plstvo.hpp"

template <typename T>
static void read_vec(std::ifstream& f, std::vector<T>& v) {
    int64_t n = 0;
    f.read(reinterpret_cast<char*>(&n), 8);
    v.resize((size_t)n);
    f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * sizeof(T)));
}

static plstvo::StereoFrame read_frame(std::ifstream& f) {
    plstvo::StereoFrame fr;
    read_vec(f, fr.pdesc); read_vec(f, fr.ldesc); read_vec(f, fr.pt_P); read_vec(f, fr.pt_pl); read_vec(f, fr.pt_sigma2);
    read_vec(f, fr.ls_sP); read_vec(f, fr.ls_eP); read_vec(f, fr.ls_le); read_vec(f, fr.ls_spl); read_vec(f, fr.ls_epl);
    read_vec(f, fr.ls_sigma2); read_vec(f, fr.ls_level);
    return fr;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        std::printf("{\"version\": %d}\n", plstvo_version());   // link check only: no device needed
        return 0;
    }
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    PlCamera cam;
    f.read(reinterpret_cast<char*>(&cam), sizeof(cam));
    PlConfig cfg;
    plstvo_kitti_config(&cfg);
    try {
        plstvo::Context ctx;
        plstvo::StereoFrameHandler h(ctx, cam, cfg);
        h.initialize(read_frame(f));
        h.insertStereoPair(read_frame(f));
        h.optimizePose();
        const PlPoseResult& r = h.result();
        std::printf("{\"good\": %d, \"status\": %d, \"n_matched_pt\": %d, \"n_matched_ls\": %d, \"n_inliers\": %d, \"err_norm\": %.17g, \"DT\": [",
                    r.good, r.status, r.n_matched_pt, r.n_matched_ls, h.n_inliers, h.curr_frame.err_norm);
        for (int i = 0; i < 16; ++i) std::printf("%s%.17g", i ? ", " : "", h.curr_frame.DT[i]);
        std::printf("]}\n");
        h.updateFrame();
    } catch (const plstvo::Error& e) {
        std::fprintf(stderr, "plstvo error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
