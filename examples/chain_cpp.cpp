// CPU-only driver of plstvo::chainPoses (include/plstvo.hpp): reads n records {good, DT[16], DT_cov[36]} (doubles), chains them
// like optimizePose's tail (src/stereoFrameHandler.cpp:377-378, :388-389), prints Tfw (16) and Tfw_cov (36) per frame.
//   g++ -std=c++17 -Iinclude examples/chain_cpp.cpp -o chain_cpp
#include <cstdio>
#include <fstream>

#include "plstvo.hpp"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) return 2;
    std::vector<PlPoseResult> res;
    double rec[1 + 16 + 36];
    while (f.read(reinterpret_cast<char*>(rec), sizeof(rec))) {
        PlPoseResult r{};
        r.good = rec[0] != 0.0;
        std::copy(rec + 1, rec + 17, r.DT);
        std::copy(rec + 17, rec + 53, r.DT_cov);
        res.push_back(r);
    }
    plstvo::chainPoses(res.data(), (int)res.size());
    for (const PlPoseResult& r : res) {
        for (int i = 0; i < 16; ++i) std::printf("%.17g ", r.Tfw[i]);
        for (int i = 0; i < 36; ++i) std::printf("%.17g ", r.Tfw_cov[i]);
        std::printf("\n");
    }
    return 0;
}
