// Drop-in replacements for the patch / NCC members of the reference class Utility
//   std::pair<cv::Mat, cv::Mat> Utility::get_edge_patches(const Edge, const cv::Mat img, bool)   (src/utility.cpp:182-212)
//   double Utility::get_patch_similarity(const cv::Mat, const cv::Mat)                           (src/utility.cpp:163-180)
// and for MatlabNCCComputer::computeNCC(patch1, patch2) (include/MatlabNCCComputer.h:40, src/MatlabNCCComputer.cpp:58-90;
// dead code in the reference: USE_MATLAB_NCC is defined by no build file and the class has no call site), all over
// the C ABI (ebvo_edge_patches, ebvo_ncc_patch_pair).  Compiled against the reference's own headers.
//
// Linking: compile src/utility.cpp with -Dget_edge_patches=get_edge_patches_cpu -Dget_patch_similarity=get_patch_similarity_cpu
// and add this file (dropin/Makefile does exactly that).  One call handles ONE edge / ONE patch pair and uploads the
// image, so these adapters exist for API compatibility (Temporal_Matches.cpp:440-451 and the reference's tests);
// the batched C-ABI entry points - and the fused kernels inside ebvo_stereo_match - are the fast path.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <limits>
#include <vector>
#include <opencv2/opencv.hpp>
#include <Eigen/Dense>

#include "utility.h"             // the reference header
#include "ebvo_b200.h"
#include "ebvo_dropin_common.hpp"

namespace {
// the reference passes the image converted to CV_64F (Stereo_Matches.cpp:562-563); the library samples 8-bit data
bool to_u8(const cv::Mat& img, std::vector<unsigned char>& out)
{
    out.resize((size_t)img.rows * img.cols);
    for (int r = 0; r < img.rows; ++r)
        for (int c = 0; c < img.cols; ++c) {
            double v;
            switch (img.depth()) {
            case CV_8U: v = img.at<unsigned char>(r, c); break;
            case CV_32F: v = img.at<float>(r, c); break;
            case CV_64F: v = img.at<double>(r, c); break;
            default: return false;
            }
            if (!(v >= 0.0 && v <= 255.0) || v != std::floor(v)) return false;
            out[(size_t)r * img.cols + c] = (unsigned char)v;
        }
    return true;
}
double ncc_of(const cv::Mat& a, const cv::Mat& b)
{
    const int n = a.rows * a.cols;
    if (n != b.rows * b.cols || n != 49 || a.depth() != CV_32F || b.depth() != CV_32F) {
        std::printf("\033[1;31m[ERROR] patch similarity expects two 7x7 CV_32F patches\033[0m\n");
        return std::nan("");
    }
    float pa[49], pb[49];
    for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) { pa[i * 7 + j] = a.at<float>(i, j); pb[i * 7 + j] = b.at<float>(i, j); }
    ebvo_dropin::Lease lease(64, 64, 1024);
    ebvo_ctx* ctx = lease.ctx;
    if (!ctx) return std::nan("");
    double out = std::nan("");
    const int rc = ebvo_ncc_patch_pair(ctx, pa, pb, 1, &out);
    if (rc != EBVO_OK) std::printf("\033[1;31m[ERROR] ebvo_ncc_patch_pair failed (%d): %s\033[0m\n", rc, ebvo_last_error(ctx));
    return out;
}
}  // namespace

std::pair<cv::Mat, cv::Mat> Utility::get_edge_patches(const Edge edge, const cv::Mat img, bool b_debug)
{
    (void)b_debug;
    cv::Mat plus(PATCH_SIZE, PATCH_SIZE, CV_32F), minus(PATCH_SIZE, PATCH_SIZE, CV_32F);      // utility.cpp:190-191
    std::vector<unsigned char> u8;
    if (!to_u8(img, u8)) {
        std::printf("\033[1;31m[ERROR] get_edge_patches: the image must hold 8-bit integer values\033[0m\n");
        return {plus, minus};
    }
    ebvo_dropin::Lease lease(img.cols, img.rows, 1024);
    ebvo_ctx* ctx = lease.ctx;
    if (!ctx) return {plus, minus};
    const ebvo_edge e{edge.location.x, edge.location.y, edge.orientation, edge.index, edge.frame_source};
    float pp[49], pm[49];
    const int rc = ebvo_edge_patches(ctx, u8.data(), img.cols, img.rows, img.cols, &e, 1, pp, pm);
    if (rc != EBVO_OK) { std::printf("\033[1;31m[ERROR] ebvo_edge_patches failed (%d): %s\033[0m\n", rc, ebvo_last_error(ctx)); return {plus, minus}; }
    for (int i = 0; i < PATCH_SIZE; ++i)
        for (int j = 0; j < PATCH_SIZE; ++j) { plus.at<float>(i, j) = pp[i * PATCH_SIZE + j]; minus.at<float>(i, j) = pm[i * PATCH_SIZE + j]; }
    return {plus, minus};
}

double Utility::get_patch_similarity(const cv::Mat patch_one, const cv::Mat patch_two) { return ncc_of(patch_one, patch_two); }

// ---- MatlabNCCComputer (include/MatlabNCCComputer.h, guarded by USE_MATLAB_NCC in the reference) -------------------
#define USE_MATLAB_NCC
namespace matlab { namespace engine { class MATLABEngine {}; } namespace data { class ArrayFactory {}; } }   // the header only forward-declares them
#include "MatlabNCCComputer.h"

MatlabNCCComputer::MatlabNCCComputer() : initialized(false) {}
MatlabNCCComputer::~MatlabNCCComputer() {}
bool MatlabNCCComputer::initialize()
{
    initialized = ebvo_dropin::Lease(64, 64, 1024).ctx != nullptr;     // no MATLAB engine is started: the GPU computes the NCC
    return initialized;
}
double MatlabNCCComputer::computeNCC(const cv::Mat& patch1, const cv::Mat& patch2)
{
    if (!initialized) {                                              // MatlabNCCComputer.cpp:60-64
        std::cerr << "MATLAB engine not initialized!" << std::endl;
        return std::numeric_limits<double>::quiet_NaN();
    }
    return ncc_of(patch1, patch2);
}
MatlabNCCComputer& getMatlabNCCComputer()
{
    static MatlabNCCComputer instance;                               // MatlabNCCComputer.cpp:118-127
    if (!instance.isInitialized()) instance.initialize();
    return instance;
}
