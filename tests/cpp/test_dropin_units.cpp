// Drives the single-purpose drop-ins (dropin/edge_clusterer_b200.cpp, dropin/utility_ncc_b200.cpp) through the
// reference's own class interfaces: EdgeClusterer(std::vector<Edge>, std::vector<int>, bool).performClustering(),
// Utility::get_edge_patches / get_patch_similarity, MatlabNCCComputer::computeNCC.  TEST INFRASTRUCTURE.
//
// usage: test_dropin_units <in.bin> <out.bin>
//   in : int32 W, H, nE, nSets; u8 img[H*W]; double xyt[3*nE]; then nSets x { int32 n; double xyt[3*n] } (cluster inputs)
//   out: nE x { float plus[49], minus[49] }; (nE-1) x double ncc(plus[k], plus[k+1]); (nE-1) x double matlab ncc(minus[k], plus[k+1]);
//        nSets x { int32 ncl; ncl x double[4] {x, y, theta, members}; int32 labels[n] }
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <random>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <opencv2/opencv.hpp>
#include <Eigen/Dense>

#include "utility.h"
#include "EdgeClusterer.h"
#define USE_MATLAB_NCC
#include "MatlabNCCComputer.h"

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int hdr[4];
    if (std::fread(hdr, 4, 4, f) != 4) return 3;
    const int W = hdr[0], H = hdr[1], nE = hdr[2], nSets = hdr[3];
    std::vector<unsigned char> pix((size_t)W * H);   // (indices.hpp leaks a function-like macro named img)
    std::vector<double> xyt((size_t)3 * nE);
    if (std::fread(pix.data(), 1, pix.size(), f) != pix.size() || std::fread(xyt.data(), 8, xyt.size(), f) != xyt.size()) return 3;
    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 6;

    Utility util;
    cv::Mat I8(H, W, CV_8UC1, pix.data(), (size_t)W), I64;
    I8.convertTo(I64, CV_64F);                                       // as Stereo_Matches.cpp:562-563
    std::vector<std::pair<cv::Mat, cv::Mat>> patches;
    for (int k = 0; k < nE; ++k) {
        Edge e; e.location = cv::Point2d(xyt[3 * k], xyt[3 * k + 1]); e.orientation = xyt[3 * k + 2]; e.index = k;
        patches.push_back(util.get_edge_patches(e, I64));
        float buf[98];
        for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) { buf[i * 7 + j] = patches[k].first.at<float>(i, j); buf[49 + i * 7 + j] = patches[k].second.at<float>(i, j); }
        std::fwrite(buf, 4, 98, o);
    }
    for (int k = 0; k + 1 < nE; ++k) { const double s = util.get_patch_similarity(patches[k].first, patches[k + 1].first); std::fwrite(&s, 8, 1, o); }
    MatlabNCCComputer& mc = getMatlabNCCComputer();
    for (int k = 0; k + 1 < nE; ++k) { const double s = mc.computeNCC(patches[k].second, patches[k + 1].first); std::fwrite(&s, 8, 1, o); }

    for (int sidx = 0; sidx < nSets; ++sidx) {
        int n = 0;
        if (std::fread(&n, 4, 1, f) != 1) return 3;
        std::vector<double> c((size_t)3 * n);
        if (std::fread(c.data(), 8, c.size(), f) != c.size()) return 3;
        std::vector<Edge> v((size_t)n); std::vector<int> idx((size_t)n);
        for (int k = 0; k < n; ++k) { v[k].location = cv::Point2d(c[3 * k], c[3 * k + 1]); v[k].orientation = c[3 * k + 2]; idx[k] = k; }
        EdgeClusterer cl(v, idx, true);                              // Stereo_Matches.cpp:1031-1033
        cl.performClustering();
        const int ncl = (int)cl.returned_clusters.size();
        std::fwrite(&ncl, 4, 1, o);
        for (int k = 0; k < ncl; ++k) {
            const double row[4] = {cl.returned_clusters[k].center_edge.location.x, cl.returned_clusters[k].center_edge.location.y,
                                   cl.returned_clusters[k].center_edge.orientation, (double)cl.returned_clusters[k].contributing_edges.size()};
            std::fwrite(row, 8, 4, o);
        }
        std::fwrite(cl.cluster_labels.data(), 4, (size_t)n, o);
    }
    std::fclose(f);
    std::fclose(o);
    return 0;
}
