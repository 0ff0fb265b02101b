// Drives the drop-in member of Temporal_Matches (dropin/temporal_matches_b200.cpp) the way
// Pipeline::get_Temporal_Edge_Correspondences does (reference src/Pipeline.cpp:147-167): add_edges_to_spatial_grid is the
// REFERENCE'S OWN code (src/Temporal_Matches.cpp compiled in place, see dropin/Makefile), then the GPU-backed member runs
// on the reference's own KF_Temporal_Edge_Quads / final_stereo_edge_pair containers.  TEST INFRASTRUCTURE: the groups
// are built from a caller mask instead of build_Veridical_Quads (ground-truth poses), as oracle/ref_temporal_harness.cpp does.
//
// usage: test_dropin_temporal <in.bin> <out.bin>
//   in : int32 W, H, n_kf, n_cf; u8 kfL[H*W], kfR[H*W], cfL[H*W], cfR[H*W]; double kf[6*n_kf], cf[6*n_cf]; u8 mask[n_kf]
//   out: int32 n; n x 13 doubles {kf index, cf index, lx, ly, lth, rx, ry, rth, ncc_l, ncc_r, score_l, score_r, valid}
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

#include "Temporal_Matches.h"

cv::Mat merged_visualization_global;   // declared extern in Dataset.h (defined in Dataset.cpp, which is not built)

Dataset::Dataset(YAML::Node n)
{
    (void)n;
    utility_tool = std::make_shared<Utility>();
    omp_threads = omp_get_num_procs();
    file_info.dataset_type = "ETH3D_slam";
    file_info.has_gt = false;
    file_info.output_path = "/tmp";
    Total_Num_Of_Imgs = 0;
    left_img_height = left_img_width = right_img_height = right_img_width = 0;
}

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    int hdr[4];
    in.read((char*)hdr, sizeof hdr);
    const int W = hdr[0], H = hdr[1], n_kf = hdr[2], n_cf = hdr[3];
    std::vector<unsigned char> img[4];
    for (auto& v : img) { v.resize((size_t)W * H); in.read((char*)v.data(), v.size()); }
    std::vector<double> kf(6 * (size_t)n_kf), cf(6 * (size_t)n_cf);
    in.read((char*)kf.data(), kf.size() * 8); in.read((char*)cf.data(), cf.size() * 8);
    std::vector<unsigned char> mask((size_t)n_kf);
    in.read((char*)mask.data(), mask.size());
    if (!in) { std::fprintf(stderr, "short input\n"); return 2; }

    YAML::Node node;
    Dataset::Ptr dataset = std::make_shared<Dataset>(node);
    Temporal_Matches engine(dataset);
    auto mat = [&](std::vector<unsigned char>& v) { return cv::Mat(H, W, CV_8UC1, (void*)v.data(), (size_t)W).clone(); };
    StereoFrame keyframe, current;
    keyframe.left_image = mat(img[0]); keyframe.left_image_undistorted = mat(img[0]); keyframe.right_image = mat(img[1]); keyframe.right_image_undistorted = mat(img[1]);
    current.left_image = mat(img[2]); current.left_image_undistorted = mat(img[2]); current.right_image = mat(img[3]); current.right_image_undistorted = mat(img[3]);
    auto mates = [](const std::vector<double>& m, int n) {
        std::vector<final_stereo_edge_pair> v((size_t)n);
        for (int i = 0; i < n; ++i) {
            v[i].left_edge.location = cv::Point2d(m[6 * i], m[6 * i + 1]); v[i].left_edge.orientation = m[6 * i + 2]; v[i].left_edge.index = i;
            v[i].right_edge.location = cv::Point2d(m[6 * i + 3], m[6 * i + 4]); v[i].right_edge.orientation = m[6 * i + 5]; v[i].right_edge.index = i;
        }
        return v;
    };
    const std::vector<final_stereo_edge_pair> KF = mates(kf, n_kf), CF = mates(cf, n_cf);
    SpatialGrid gl(W, H, GRID_SIZE), gr(W, H, GRID_SIZE);                    // Pipeline.h:99-100
    engine.add_edges_to_spatial_grid(CF, gl, gr);                            // Pipeline.cpp:153 (reference code)
    std::vector<KF_Temporal_Edge_Quads> quads;
    for (int i = 0; i < n_kf; ++i) {
        KF_Temporal_Edge_Quads k;
        k.KF_stereo_mate = &KF[i];
        k.projected_orientation_left = k.projected_orientation_right = 0.0;
        if (mask[i]) k.veridical_quads.resize(1);
        quads.push_back(k);
    }
    Stereo_Edge_Pairs kfPairs, cfPairs;
    if (argc > 4 && std::string(argv[3]) == "time") {   // timing mode: wall clock of the GPU-backed member, argv[4] repetitions after one untimed call
        const int reps = std::atoi(argv[4]);
        double ms = 0.0;
        size_t nq = 0;
        for (int r = -1; r < reps; ++r) {
            const auto t0 = std::chrono::steady_clock::now();
            engine.get_Temporal_Edge_Pairs_from_Quads(quads, KF, CF, gl, gr, kfPairs, cfPairs, keyframe, current, 0, 1);
            if (r >= 0) ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            nq = 0;
            for (const auto& q : quads) nq += q.candidate_quads.size();
        }
        FILE* o = std::fopen(argv[2], "w");
        if (!o) return 6;
        std::fprintf(o, "{\"what\": \"Temporal_Matches::get_Temporal_Edge_Pairs_from_Quads (GPU drop-in), one keyframe -> current-frame pair\", \"H\": %d, \"W\": %d, "
                        "\"kf_mates\": %d, \"cf_mates\": %d, \"quads\": %zu, \"wall_ms\": %.3f}\n", H, W, n_kf, n_cf, nq, ms / reps);
        std::fclose(o);
        return 0;
    }
    engine.get_Temporal_Edge_Pairs_from_Quads(quads, KF, CF, gl, gr, kfPairs, cfPairs, keyframe, current, 0, 1);   // Pipeline.cpp:159-167 (GPU-backed)

    std::vector<double> rows;
    for (int i = 0; i < n_kf; ++i)
        for (const auto& cq : quads[i].candidate_quads) {
            const double r[13] = {(double)i, (double)cq.CF_left->cf_stereo_edge_mate_index,
                                  cq.CF_left->center_edge.location.x, cq.CF_left->center_edge.location.y, cq.CF_left->center_edge.orientation,
                                  cq.CF_right->center_edge.location.x, cq.CF_right->center_edge.location.y, cq.CF_right->center_edge.orientation,
                                  cq.CF_left->matching_scores.ncc_score, cq.CF_right->matching_scores.ncc_score,
                                  cq.CF_left->refine_final_score, cq.CF_right->refine_final_score, cq.CF_left->refine_validity ? 1.0 : 0.0};
            rows.insert(rows.end(), r, r + 13);
        }
    std::ofstream out(argv[2], std::ios::binary);
    const int n = (int)(rows.size() / 13);
    out.write((const char*)&n, 4);
    out.write((const char*)rows.data(), rows.size() * 8);
    return 0;
}
