// Stand-in for the sliver of OpenCV the shim touches (cv::Mat as an N x 32 CV_8UC1 descriptor matrix).  TEST INFRASTRUCTURE:
// lets shim/*.cpp be compiled and run in an image without OpenCV headers; on the reference side the real <opencv2/core.hpp> is used.
#pragma once
#include <cstdint>
#include <vector>
#define CV_8UC1 0
namespace cv {
class Mat {
public:
    int rows = 0, cols = 0;
    std::vector<uint8_t> buf;
    Mat() {}
    Mat(int r, int c, int /*type*/) : rows(r), cols(c), buf((size_t)r * c) {}
    bool empty() const { return rows == 0 || cols == 0; }
    bool isContinuous() const { return true; }
    int type() const { return CV_8UC1; }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(buf.data() + (size_t)r * cols); }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(buf.data() + (size_t)r * cols); }
};
}  // namespace cv
