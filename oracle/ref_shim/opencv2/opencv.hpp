// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <opencv2/opencv.hpp>.
//
// The reference's third-order edge detector (/root/reference/src/toed/cpu_toed.cpp)
// touches OpenCV only through cv::Mat::at<uchar>(i,j) (cpu_toed.cpp:93) and
// cv::Point2d (cpu_toed.hpp:28, cpu_toed.cpp:526). This header provides exactly
// those two types so that the UNMODIFIED reference source compiles in place into
// oracle/_ref/libtoed_ref.so (see oracle/Makefile). Nothing here is product code.
#pragma once
#include <cstdint>
#include <cstddef>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;  // OpenCV's global typedef (core/hal/interface.h)

namespace cv {

struct Point2d {
    double x, y;
    Point2d() : x(0), y(0) {}
    Point2d(double x_, double y_) : x(x_), y(y_) {}
};

// Non-owning 8-bit single-channel view.
class Mat {
public:
    int rows, cols;
    const unsigned char* data;
    size_t step;
    Mat() : rows(0), cols(0), data(nullptr), step(0) {}
    Mat(int r, int c, const unsigned char* d, size_t s) : rows(r), cols(c), data(d), step(s) {}
    template <typename T> const T& at(int i, int j) const {
        return *reinterpret_cast<const T*>(data + (size_t)i * step + (size_t)j * sizeof(T));
    }
};

}  // namespace cv
