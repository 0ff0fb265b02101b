// Drives the drop-in members of Stereo_Matches (dropin/stereo_matches_b200.cpp) exactly like
// Pipeline::get_Stereo_Edge_Correspondences (reference src/Pipeline.cpp:109-131): frames come from a StereoIterator
// (the in-memory dropin/synthetic_stereo_iterator.hpp), Find_Stereo_GT_Locations and get_Stereo_Edge_GT_Pairs are
// the REFERENCE'S OWN code (src/Stereo_Matches.cpp compiled in place, see dropin/Makefile), then the two GPU-backed
// members run.  TEST INFRASTRUCTURE: Dataset.cpp (yaml-cpp, file I/O) is not built, its calibration part is restated
// below as in oracle/ref_stereo_harness.cpp.
//
// usage: test_dropin_stereo <in.bin> <out.bin> [output dir for the reference's own text writer]
//   in : int32 W, H, nL, nR; double Kl[9], Kr[9], R21[9], T21[3]; u8 L[H*W], R[H*W]; double Lxyt[3*nL], Rxyt[3*nR]
//   out: int32 n; n x 16 doubles {left index, lx, ly, lth, rx, ry, rth, score, sum(L+), sum(L-), sum(R+), sum(R-), b_is_TP, line c,
//        sum(left descriptor 1) or -1 when empty, sum(right descriptor 2) or -1}
#include <algorithm>
#include <chrono>
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
#include <omp.h>
#include <opencv2/opencv.hpp>
#include <Eigen/Dense>
#include <yaml-cpp/yaml.h>

#define private public
#include "Stereo_Matches.h"
#undef private
#include "../../dropin/synthetic_stereo_iterator.hpp"

cv::Mat merged_visualization_global;   // declared extern in Dataset.h:361 (defined in Dataset.cpp, which is not built)

Dataset::Dataset(YAML::Node n)          // calibration part of Dataset.cpp:99-113
{
    utility_tool = std::make_shared<Utility>();
    omp_threads = omp_get_num_procs();
    file_info.dataset_type = "KITTI";
    file_info.has_gt = false;
    file_info.output_path = "/tmp";
    auto M = [](const double* p) { Eigen::Matrix3d m; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m(i, j) = p[3 * i + j]; return m; };
    camera_info.left.K = M(n.Kl);
    camera_info.right.K = M(n.Kr);
    camera_info.left.R = M(n.R21);
    camera_info.left.T = Eigen::Vector3d(n.T21[0], n.T21[1], n.T21[2]);
    camera_info.left.F = camera_info.right.K.inverse().transpose() * (utility_tool->get_Skew_Symmetric_Matrix(camera_info.left.T) * camera_info.left.R) * camera_info.left.K.inverse();
    camera_info.right.R = camera_info.left.R.transpose();
    camera_info.right.T = -(camera_info.left.R.transpose() * camera_info.left.T);
    camera_info.right.F = camera_info.left.K.inverse().transpose() * (utility_tool->get_Skew_Symmetric_Matrix(camera_info.right.T) * camera_info.right.R) * camera_info.right.K.inverse();
    Total_Num_Of_Imgs = 0;
    left_img_height = left_img_width = right_img_height = right_img_width = 0;
}

static double patch_sum(const cv::Mat& m)
{
    double s = 0;
    for (int i = 0; i < m.rows; ++i) for (int j = 0; j < m.cols; ++j) s += m.at<float>(i, j);
    return s;
}

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int hdr[4];
    double cal[30];
    if (std::fread(hdr, 4, 4, f) != 4 || std::fread(cal, 8, 30, f) != 30) return 3;
    const int W = hdr[0], H = hdr[1], nL = hdr[2], nR = hdr[3];
    std::vector<unsigned char> Lb((size_t)W * H), Rb((size_t)W * H);
    std::vector<double> Le((size_t)3 * nL), Re((size_t)3 * nR);
    if (std::fread(Lb.data(), 1, Lb.size(), f) != Lb.size() || std::fread(Rb.data(), 1, Rb.size(), f) != Rb.size()) return 3;
    if (std::fread(Le.data(), 8, Le.size(), f) != Le.size() || std::fread(Re.data(), 8, Re.size(), f) != Re.size()) return 3;
    std::fclose(f);

    YAML::Node node; node.Kl = cal; node.Kr = cal + 9; node.R21 = cal + 18; node.T21 = cal + 27;
    Dataset::Ptr dataset = std::make_shared<Dataset>(node);
    Stereo_Matches::Ptr engine = std::make_shared<Stereo_Matches>();                       // Pipeline.cpp:18

    // cmd/main_VO.cpp:99-113: frames come from a StereoIterator
    SyntheticStereoIterator it({cv::Mat(H, W, CV_8UC1, Lb.data(), (size_t)W)}, {cv::Mat(H, W, CV_8UC1, Rb.data(), (size_t)W)});
    StereoFrame frame;
    if (!it.hasNext() || !it.getNext(frame)) return 4;
    // Pipeline::prepare_Stereo_Images (Pipeline.cpp:64-107): zero distortion => undistort is the identity; edges supplied
    frame.left_image_undistorted = frame.left_image.clone();
    frame.right_image_undistorted = frame.right_image.clone();
    auto to_edges = [](const std::vector<double>& xyt, int n) {
        std::vector<Edge> v((size_t)n);
        for (int k = 0; k < n; ++k) { v[k].location = cv::Point2d(xyt[3 * k], xyt[3 * k + 1]); v[k].orientation = xyt[3 * k + 2]; v[k].index = k; }
        return v;
    };
    frame.left_edges = to_edges(Le, nL);
    frame.right_edges = to_edges(Re, nR);

    // Pipeline::get_Stereo_Edge_Correspondences (Pipeline.cpp:109-131)
    Stereo_Edge_Pairs pairs;
    pairs.stereo_frame = &frame;
    engine->Find_Stereo_GT_Locations(dataset, cv::Mat(), frame, pairs, true);              // reference code
    engine->get_Stereo_Edge_GT_Pairs(dataset, frame, pairs, true);                         // reference code
    Timing_Statistics timing;
    if (argc > 4 && std::string(argv[3]) == "time") {   // timing mode: wall clock of the two GPU-backed members, argv[4] repetitions
        const int reps = std::atoi(argv[4]);
        const Stereo_Edge_Pairs pairs0 = pairs;
        double ms = 0.0;
        size_t nm = 0;
        for (int r = -1; r < reps; ++r) {               // one untimed pass first (context creation, first-touch)
            Stereo_Edge_Pairs p = pairs0;
            std::vector<final_stereo_edge_pair> mm;
            const auto t0 = std::chrono::steady_clock::now();
            engine->get_Stereo_Edge_Pairs(dataset, p, 0, timing);
            engine->finalize_stereo_edge_mates(p, mm);
            if (r >= 0) ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            nm = mm.size();
        }
        FILE* o = std::fopen(argv[2], "w");
        if (!o) return 6;
        std::fprintf(o, "{\"what\": \"Stereo_Matches::get_Stereo_Edge_Pairs + finalize_stereo_edge_mates (GPU drop-in), one frame\", \"H\": %d, \"W\": %d, "
                        "\"left_edges\": %d, \"right_edges\": %d, \"mates\": %zu, \"wall_ms\": %.3f, \"timing_statistics_total_ms\": %.3f, "
                        "\"time_EP\": %.3f, \"time_SIFT\": %.3f, \"time_NCC\": %.3f, \"time_Refinement\": %.3f, \"time_Clustering\": %.3f, \"time_Post_NCC\": %.3f}\n",
                     H, W, nL, nR, nm, ms / reps, timing.total_time, timing.time_EP, timing.time_SIFT, timing.time_NCC, timing.time_Refinement,
                     timing.time_Clustering, timing.time_Post_NCC);
        std::fclose(o);
        return 0;
    }
    Frame_Evaluation_Metrics metrics = engine->get_Stereo_Edge_Pairs(dataset, pairs, 0, timing);   // GPU drop-in
    std::vector<final_stereo_edge_pair> mates;
    engine->finalize_stereo_edge_mates(pairs, mates);                                      // GPU drop-in
    if (!(timing.total_time > 0.0 && timing.time_Refinement > 0.0)) return 7;              // Timing_Statistics is filled from the kernel times
    if (!metrics.stages.empty() || mates.size() != pairs.focused_edge_indices.size()) return 5;
    if (argc > 3) {   // the on-disk format: the REFERENCE'S OWN writer (Stereo_Matches.cpp:1656-1699) consumes the GPU mates unchanged
        dataset->file_info.output_path = argv[3];
        engine->write_finalized_stereo_edge_pairs_to_file(dataset, mates, 0);              // Pipeline.cpp:131
    }

    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 6;
    const int n = (int)mates.size();
    std::fwrite(&n, 4, 1, o);
    for (int i = 0; i < n; ++i) {
        const final_stereo_edge_pair& m = mates[i];
        const double dsl = m.left_edge_descriptors.first.empty() ? -1.0 : patch_sum(m.left_edge_descriptors.first);
        const double dsr = m.right_edge_descriptors.second.empty() ? -1.0 : patch_sum(m.right_edge_descriptors.second);
        const double row[16] = {(double)pairs.focused_edge_indices[i], m.left_edge.location.x, m.left_edge.location.y, m.left_edge.orientation,
                                m.right_edge.location.x, m.right_edge.location.y, m.right_edge.orientation,
                                pairs.matching_edge_clusters[i].refine_final_scores[0],
                                patch_sum(m.left_edge_patches.first), patch_sum(m.left_edge_patches.second),
                                patch_sum(m.right_edge_patches.first), patch_sum(m.right_edge_patches.second),
                                m.b_is_TP ? 1.0 : 0.0, pairs.epip_line_coeffs_of_left_edges[i](2), dsl, dsr};
        std::fwrite(row, 8, 16, o);
    }
    std::fclose(o);
    return 0;
}
