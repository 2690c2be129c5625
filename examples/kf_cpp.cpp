// CPU-only driver of the host-side state machine in include/plstvo.hpp (adaptive FAST threshold + key-frame test,
// src/stereoFrameHandler.cpp:62-86, :1136-1218).  Reads a flat binary sequence of per-frame records
// {Tfw[16], DT[16], DT_cov[36], err_norm, n_inliers_pt(as double)}, prints one line per frame:
//   new_kf orb_fast_th entropy_curr entropy_ratio t r N_prevKF_currF
// No device and no library call is involved (the header's inline C-ABI wrappers are not referenced).
//   g++ -std=c++17 -Iinclude examples/kf_cpp.cpp -o kf_cpp
#include <cstdio>
#include <fstream>

#include "plstvo.hpp"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) return 2;
    plstvo::HandlerConfig hc;
    plstvo::KeyframeTest kf;
    int orb_fast_th = hc.orb_fast_th;
    double rec[16 + 16 + 36 + 2];
    while (f.read(reinterpret_cast<char*>(rec), sizeof(rec))) {
        plstvo::StereoFrame fr;
        std::copy(rec, rec + 16, fr.Tfw.begin());
        std::copy(rec + 16, rec + 32, fr.DT.begin());
        std::copy(rec + 32, rec + 68, fr.DT_cov.begin());
        fr.err_norm = rec[68];
        const int n_inliers_pt = (int)rec[69];
        const bool new_kf = kf.needNewKF(hc, fr);
        std::printf("%d %d %.17g %.17g %.17g %.17g %d\n", new_kf ? 1 : 0,
                    orb_fast_th = plstvo::updateFastThreshold(hc, orb_fast_th, fr, n_inliers_pt), kf.entropy_curr,
                    kf.entropy_ratio, kf.kf_t, kf.kf_r, kf.N_prevKF_currF);
        if (new_kf) kf.currFrameIsKF(fr);
    }
    return 0;
}
