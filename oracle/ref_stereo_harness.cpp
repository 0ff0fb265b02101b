// TEST INFRASTRUCTURE ONLY (oracle/): drives the UNMODIFIED reference stereo sources
//   /root/reference/src/Stereo_Matches.cpp, src/utility.cpp, src/EdgeClusterer.cpp
// compiled in place against third_party_shim (mini OpenCV / Eigen / yaml-cpp stand-ins) -> oracle/_ref/libstereo_ref.so.
// Purpose: pin oracle/stereo_oracle.cpp (the restatement) against the reference's own control flow and arithmetic,
// stage by stage (tests/test_oracle_stereo.py, golden fixture tests/golden/stereo_ref_small.npz).
//
// The harness follows Pipeline::get_Stereo_Edge_Correspondences (src/Pipeline.cpp:109-131) and the stage order of
// Stereo_Matches::get_Stereo_Edge_Pairs (src/Stereo_Matches.cpp:1360-1540) by calling the reference's own stage
// methods one by one, with exactly two documented differences:
//   (1) SIFT-off: augment_Edge_Data / apply_SIFT_filtering / apply_Best_Nearly_Best_Test(SIFT) are not called
//       (cv::SIFT is OpenCV code, not reference code);
//   (2) before the second consolidate_redundant_edge_hypothesis call a placeholder index is pushed into every
//       cluster's contributing_edges_toed_indices, because Stereo_Matches.cpp:991 reads element [0] of vectors the
//       first call emptied (:993-997) - undefined behaviour at HEAD (segfault under -O2).  The value read is unused.
// `#define private public` is applied to the reference headers only, to reach remove_empty_clusters() and to
// define Dataset's constructor from numbers instead of a YAML file.
#include <algorithm>
#include <cmath>
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

cv::Mat merged_visualization_global;   // declared extern in Dataset.h:361 (defined in Dataset.cpp, which is not built)

// Dataset.cpp is not compiled (yaml-cpp / filesystem I/O).  Calibration part of its constructor, Dataset.cpp:99-113:
Dataset::Dataset(YAML::Node n)
{
    utility_tool = std::make_shared<Utility>();
    omp_threads = omp_get_num_procs();
    file_info.dataset_type = "KITTI";
    file_info.has_gt = false;                               // Dataset.cpp:120-148 for KITTI / EuRoC / ETH3D_slam
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

namespace {
enum { ST_EPI = 0, ST_DISP, ST_ORIENT, ST_SIFT, ST_NCC, ST_BNB_NCC, ST_BNB_SIFT, ST_SHIFT, ST_GN, ST_CLUSTER, ST_NCC2, ST_BEST, ST_COUNT };
struct StageDump { std::vector<int> off, ridx; std::vector<double> x, y, th, score; };
struct Result {
    int nL = 0;
    std::vector<StageDump> stages;
    std::vector<int> mate_left;
    std::vector<double> mate_rx, mate_ry, mate_rth, mate_score;
    std::vector<double> lines, F21;
};
void dump(Result& r, int st, const Stereo_Edge_Pairs& p)
{
    StageDump& d = r.stages[st];
    const int n = (int)p.matching_edge_clusters.size();
    d.off.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) d.off[i + 1] = d.off[i] + (int)p.matching_edge_clusters[i].edge_clusters.size();
    const size_t tot = d.off[n];
    d.ridx.resize(tot); d.x.resize(tot); d.y.resize(tot); d.th.resize(tot); d.score.resize(tot);
    for (int i = 0; i < n; ++i) {
        const auto& c = p.matching_edge_clusters[i];
        for (size_t j = 0; j < c.edge_clusters.size(); ++j) {
            const size_t k = d.off[i] + j;
            d.ridx[k] = c.edge_clusters[j].contributing_edges_toed_indices.empty() ? -1 : c.edge_clusters[j].contributing_edges_toed_indices[0];
            d.x[k] = c.edge_clusters[j].center_edge.location.x; d.y[k] = c.edge_clusters[j].center_edge.location.y;
            d.th[k] = c.edge_clusters[j].center_edge.orientation;
            d.score[k] = j < c.refine_final_scores.size() ? c.refine_final_scores[j] : std::nan("");
        }
    }
}
}  // namespace

extern "C" {

void* rs_run(const unsigned char* Lraw, const unsigned char* Rraw, const unsigned char* Lund, const unsigned char* Rund, int H, int W,
             const double* Lxyt, int nL, const double* Rxyt, int nR, const double* Kl, const double* Kr, const double* R21, const double* T21)
{
    YAML::Node node; node.Kl = Kl; node.Kr = Kr; node.R21 = R21; node.T21 = T21;
    Dataset::Ptr dataset = std::make_shared<Dataset>(node);
    Stereo_Matches::Ptr engine = std::make_shared<Stereo_Matches>();

    // StereoIterator::getNext + Pipeline::prepare_Stereo_Images (Pipeline.cpp:64-107), images and edges supplied by the caller
    StereoFrame frame;
    frame.left_image = cv::Mat(H, W, CV_8UC1, (void*)Lraw, (size_t)W).clone();
    frame.right_image = cv::Mat(H, W, CV_8UC1, (void*)Rraw, (size_t)W).clone();
    frame.left_image_undistorted = cv::Mat(H, W, CV_8UC1, (void*)Lund, (size_t)W).clone();
    frame.right_image_undistorted = cv::Mat(H, W, CV_8UC1, (void*)Rund, (size_t)W).clone();
    util_compute_Img_Gradients(frame.left_image_undistorted, frame.left_image_gradients_x, frame.left_image_gradients_y);
    util_compute_Img_Gradients(frame.right_image_undistorted, frame.right_image_gradients_x, frame.right_image_gradients_y);
    auto to_edges = [](const double* xyt, int n) {
        std::vector<Edge> v((size_t)n);
        for (int k = 0; k < n; ++k) { v[k].location = cv::Point2d(xyt[3 * k], xyt[3 * k + 1]); v[k].orientation = xyt[3 * k + 2]; v[k].index = k; }
        return v;
    };
    frame.left_edges = to_edges(Lxyt, nL);
    frame.right_edges = to_edges(Rxyt, nR);

    // Pipeline::get_Stereo_Edge_Correspondences (Pipeline.cpp:109-131)
    Stereo_Edge_Pairs pairs;
    pairs.stereo_frame = &frame;
    engine->Find_Stereo_GT_Locations(dataset, cv::Mat(), frame, pairs, true);
    engine->get_Stereo_Edge_GT_Pairs(dataset, frame, pairs, true);

    Result* res = new Result;
    res->nL = nL;
    res->stages.resize(ST_COUNT);
    // Stereo_Matches::get_Stereo_Edge_Pairs stage order (Stereo_Matches.cpp:1374-1526), SIFT stages left out
    engine->apply_Epipolar_Line_Distance_Filtering(pairs, dataset, frame.right_edges, "", true, 0, 0);
    dump(*res, ST_EPI, pairs);
    engine->apply_Disparity_Filtering(pairs, "", 0);
    dump(*res, ST_DISP, pairs);
    engine->apply_orientation_filter(pairs, 10.0, "", 0);
    dump(*res, ST_ORIENT, pairs);
    dump(*res, ST_SIFT, pairs);
    pairs.left_edge_descriptors.resize(pairs.focused_edge_indices.size());   // what augment_Edge_Data would size (:657-658)
    engine->apply_NCC_Filtering(pairs, "", 0);
    dump(*res, ST_NCC, pairs);
    engine->apply_Best_Nearly_Best_Test(pairs, BNB_NCC, "", 0, true);
    dump(*res, ST_BNB_NCC, pairs);
    dump(*res, ST_BNB_SIFT, pairs);
    engine->consolidate_redundant_edge_hypothesis(pairs, 0, true, false);
    dump(*res, ST_SHIFT, pairs);
    engine->refine_edge_disparity(pairs, 0, true);
    dump(*res, ST_GN, pairs);
    for (auto& c : pairs.matching_edge_clusters)                              // difference (2): make the dead read at :991 defined
        for (auto& ec : c.edge_clusters) if (ec.contributing_edges_toed_indices.empty()) ec.contributing_edges_toed_indices.push_back(-1);
    engine->consolidate_redundant_edge_hypothesis(pairs, false, true);        // the call as written at :1483
    dump(*res, ST_CLUSTER, pairs);
    engine->apply_NCC_Filtering(pairs, "", 0);
    dump(*res, ST_NCC2, pairs);
    engine->apply_Lowe_Ratio_Test(pairs, LOWES_RATIO, "", 0);
    dump(*res, ST_BEST, pairs);
    res->lines.resize((size_t)nL * 3);
    for (int i = 0; i < nL && i < (int)pairs.epip_line_coeffs_of_left_edges.size(); ++i)
        for (int k = 0; k < 3; ++k) res->lines[3 * (size_t)i + k] = pairs.epip_line_coeffs_of_left_edges[i](k);
    Eigen::Matrix3d F = dataset->get_fund_mat_21();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) res->F21.push_back(F(i, j));
    engine->remove_empty_clusters(pairs);
    // finalize_stereo_edge_mates (:1578-1653) computes SIFT descriptors of the mates; its edge part is just this copy (:1615-1618)
    for (size_t i = 0; i < pairs.focused_edge_indices.size(); ++i) {
        const Edge& re = pairs.matching_edge_clusters[i].edge_clusters[0].center_edge;
        res->mate_left.push_back(pairs.focused_edge_indices[i]);
        res->mate_rx.push_back(re.location.x); res->mate_ry.push_back(re.location.y); res->mate_rth.push_back(re.orientation);
        res->mate_score.push_back(pairs.matching_edge_clusters[i].refine_final_scores[0]);
    }
    return res;
}

int rs_num_mates(void* h) { return (int)((Result*)h)->mate_left.size(); }
void rs_get_mates(void* h, int* left, double* rx, double* ry, double* rth, double* score)
{
    Result* r = (Result*)h;
    const size_t n = r->mate_left.size();
    std::memcpy(left, r->mate_left.data(), n * 4);
    std::memcpy(rx, r->mate_rx.data(), n * 8); std::memcpy(ry, r->mate_ry.data(), n * 8);
    std::memcpy(rth, r->mate_rth.data(), n * 8); std::memcpy(score, r->mate_score.data(), n * 8);
}
int rs_stage_total(void* h, int st) { Result* r = (Result*)h; return r->stages[st].off.empty() ? -1 : r->stages[st].off.back(); }
void rs_get_stage(void* h, int st, int* off, int* ridx, double* x, double* y, double* th, double* score)
{
    StageDump& d = ((Result*)h)->stages[st];
    std::memcpy(off, d.off.data(), d.off.size() * 4);
    const size_t n = d.ridx.size();
    std::memcpy(ridx, d.ridx.data(), n * 4);
    std::memcpy(x, d.x.data(), n * 8); std::memcpy(y, d.y.data(), n * 8);
    std::memcpy(th, d.th.data(), n * 8); std::memcpy(score, d.score.data(), n * 8);
}
void rs_get_lines(void* h, double* lines, double* F21) { Result* r = (Result*)h; std::memcpy(lines, r->lines.data(), r->lines.size() * 8); std::memcpy(F21, r->F21.data(), 72); }
void rs_free(void* h) { delete (Result*)h; }

// single-function probes of reference code
void rs_edge_patches(const unsigned char* img, int H, int W, double x, double y, double th, float* plus49, float* minus49)
{
    Utility u;
    cv::Mat I8(H, W, CV_8UC1, (void*)img, (size_t)W), I64;
    I8.convertTo(I64, CV_64F);
    Edge e; e.location = cv::Point2d(x, y); e.orientation = th;
    std::pair<cv::Mat, cv::Mat> p = u.get_edge_patches(e, I64);
    for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) { plus49[i * 7 + j] = p.first.at<float>(i, j); minus49[i * 7 + j] = p.second.at<float>(i, j); }
}
double rs_patch_similarity(const float* a, const float* b)
{
    Utility u;
    cv::Mat A(7, 7, CV_32FC1, (void*)a, 28), B(7, 7, CV_32FC1, (void*)b, 28);
    return u.get_patch_similarity(A, B);
}
int rs_cluster(const double* xyt, int n, int by_orientation, double* centers_xyt, int* n_contrib)
{
    std::vector<Edge> v((size_t)n); std::vector<int> idx((size_t)n);
    for (int k = 0; k < n; ++k) { v[k].location = cv::Point2d(xyt[3 * k], xyt[3 * k + 1]); v[k].orientation = xyt[3 * k + 2]; idx[k] = k; }
    EdgeClusterer c(v, idx, by_orientation != 0);
    c.performClustering();
    for (size_t k = 0; k < c.returned_clusters.size(); ++k) {
        centers_xyt[3 * k] = c.returned_clusters[k].center_edge.location.x; centers_xyt[3 * k + 1] = c.returned_clusters[k].center_edge.location.y;
        centers_xyt[3 * k + 2] = c.returned_clusters[k].center_edge.orientation;
        n_contrib[k] = (int)c.returned_clusters[k].contributing_edges.size();
    }
    return (int)c.returned_clusters.size();
}

}  // extern "C"
