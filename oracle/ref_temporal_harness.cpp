// TEST INFRASTRUCTURE ONLY (oracle/): drives the UNMODIFIED reference quad-tracking source
//   /root/reference/src/Temporal_Matches.cpp (+ src/utility.cpp, src/EdgeClusterer.cpp)
// compiled in place against third_party_shim -> oracle/_ref/libtemporal_ref.so.
// Purpose: pin oracle/temporal_oracle.inl (the restatement) against the reference's own control flow and arithmetic,
// stage by stage (tests/test_oracle_temporal.py, golden fixture tests/golden/temporal_ref_small.npz).
//
// The harness follows Pipeline::get_Temporal_Edge_Correspondences (src/Pipeline.cpp:147-165) and the stage order of
// Temporal_Matches::get_Temporal_Edge_Pairs_from_Quads (src/Temporal_Matches.cpp:168-218) by calling the reference's
// own stage methods one by one, with these documented differences:
//   (1) descriptors are not computed (cv::SIFT is OpenCV code): apply_SIFT_filtering_quads and
//       apply_best_nearly_best_filtering_quads("SIFT") run on descriptor pairs supplied by the caller and are not called
//       without them ("SIFT-off": every sift_score stays at its initial 900);
//   (2) build_Veridical_Quads (:57-166) needs ground-truth poses and 3-D points; the harness builds quads_by_kf directly,
//       one KF_Temporal_Edge_Quads per KF mate selected by the caller's mask, with a one-entry veridical_quads list
//       (apply_spatial_grid_filtering_quads only tests it for emptiness, :345);
//   (3) final_stereo_edge_pair patches are filled the way the stereo stage fills them: left patches from the raw left
//       image (Stereo_Matches.cpp:562,578), right patches from the undistorted right image (:1580,1622).
// `#define private public` reaches candidate_cluster_pairs_ (the per-stage state) for the dumps.
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
#include "Temporal_Matches.h"
#undef private

cv::Mat merged_visualization_global;   // declared extern in Dataset.h (defined in Dataset.cpp, which is not built)

// Dataset.cpp is not compiled (yaml-cpp / filesystem I/O); the quad stages only call has_gt() on it
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

namespace {
enum { TQ_GRID = 0, TQ_ORIENT, TQ_NCC, TQ_SIFT, TQ_BNB, TQ_BNB_SIFT, TQ_GN, TQ_CLUSTER, TQ_COUNT };
struct QDump { std::vector<int> off, cf, valid; std::vector<double> l, r, ncc, sc, sift; };
struct TRes { int n_kf = 0; std::vector<int> sel; QDump st[TQ_COUNT]; };

void dump(TRes& R, int s, const Temporal_Matches& eng, const std::vector<KF_Temporal_Edge_Quads>& q)
{
    QDump& d = R.st[s];
    d.off.assign(R.n_kf + 1, 0);
    std::vector<int> cnt(R.n_kf, 0);
    for (size_t g = 0; g < q.size(); ++g) cnt[R.sel[g]] = (int)q[g].candidate_quads.size();
    for (int i = 0; i < R.n_kf; ++i) d.off[i + 1] = d.off[i] + cnt[i];
    const size_t tot = d.off[R.n_kf];
    d.cf.resize(tot); d.valid.resize(tot); d.l.resize(3 * tot); d.r.resize(3 * tot); d.ncc.resize(2 * tot); d.sc.resize(2 * tot); d.sift.resize(2 * tot);
    for (size_t g = 0; g < q.size(); ++g) {
        size_t o = d.off[R.sel[g]];
        for (const auto& cq : q[g].candidate_quads) {
            d.cf[o] = cq.CF_left->cf_stereo_edge_mate_index;
            d.l[3 * o] = cq.CF_left->center_edge.location.x; d.l[3 * o + 1] = cq.CF_left->center_edge.location.y; d.l[3 * o + 2] = cq.CF_left->center_edge.orientation;
            d.r[3 * o] = cq.CF_right->center_edge.location.x; d.r[3 * o + 1] = cq.CF_right->center_edge.location.y; d.r[3 * o + 2] = cq.CF_right->center_edge.orientation;
            d.ncc[2 * o] = cq.CF_left->matching_scores.ncc_score; d.ncc[2 * o + 1] = cq.CF_right->matching_scores.ncc_score;
            d.sift[2 * o] = cq.CF_left->matching_scores.sift_score; d.sift[2 * o + 1] = cq.CF_right->matching_scores.sift_score;
            d.sc[2 * o] = cq.CF_left->refine_final_score; d.sc[2 * o + 1] = cq.CF_right->refine_final_score;
            d.valid[o] = cq.CF_left->refine_validity ? 1 : 0;
            ++o;
        }
    }
}
}  // namespace

extern "C" {

void* rt_run(const unsigned char* kfLraw, const unsigned char* kfLund, const unsigned char* kfRund, const unsigned char* cfLraw,
             const unsigned char* cfLund, const unsigned char* cfRund, int H, int W, const double* kf, int n_kf,
             const unsigned char* kf_mask, const double* cf, int n_cf, const float* descKfL, const float* descKfR,
             const float* descCfL, const float* descCfR)
{
    const bool sift_on = descKfL && descKfR && descCfL && descCfR;
    YAML::Node node;
    Dataset::Ptr dataset = std::make_shared<Dataset>(node);
    Temporal_Matches engine(dataset);
    Utility util{};

    auto mat = [&](const unsigned char* p) { return cv::Mat(H, W, CV_8UC1, (void*)p, (size_t)W).clone(); };
    StereoFrame keyframe, current;
    keyframe.left_image = mat(kfLraw); keyframe.left_image_undistorted = mat(kfLund);
    keyframe.right_image = mat(kfRund); keyframe.right_image_undistorted = mat(kfRund);
    current.left_image = mat(cfLraw); current.left_image_undistorted = mat(cfLund);
    current.right_image = mat(cfRund); current.right_image_undistorted = mat(cfRund);
    // Pipeline::prepare_Stereo_Images (Pipeline.cpp:82-84)
    util_compute_Img_Gradients(current.left_image_undistorted, current.left_image_gradients_x, current.left_image_gradients_y);
    util_compute_Img_Gradients(current.right_image_undistorted, current.right_image_gradients_x, current.right_image_gradients_y);

    auto desc_pair = [](const float* d256) {     // two 1 x 128 CV_32F rows, as cv::SIFT::compute returns them (Stereo_Matches.cpp:1634)
        cv::Mat a(1, 128, CV_32F), b(1, 128, CV_32F);
        for (int k = 0; k < 128; ++k) { a.at<float>(0, k) = d256[k]; b.at<float>(0, k) = d256[128 + k]; }
        return std::make_pair(a, b);
    };
    auto mates = [&](const double* m, int n, const StereoFrame& f, const float* dL, const float* dR) {
        cv::Mat L64, R64;
        f.left_image.convertTo(L64, CV_64F);                  // Stereo_Matches.cpp:562
        f.right_image_undistorted.convertTo(R64, CV_64F);     // Stereo_Matches.cpp:1582
        std::vector<final_stereo_edge_pair> v((size_t)n);
        for (int i = 0; i < n; ++i) {
            v[i].left_edge.location = cv::Point2d(m[6 * i], m[6 * i + 1]); v[i].left_edge.orientation = m[6 * i + 2]; v[i].left_edge.index = i;
            v[i].right_edge.location = cv::Point2d(m[6 * i + 3], m[6 * i + 4]); v[i].right_edge.orientation = m[6 * i + 5]; v[i].right_edge.index = i;
            v[i].left_edge_patches = util.get_edge_patches(v[i].left_edge, L64);
            v[i].right_edge_patches = util.get_edge_patches(v[i].right_edge, R64);
            if (sift_on) { v[i].left_edge_descriptors = desc_pair(dL + (size_t)i * 256); v[i].right_edge_descriptors = desc_pair(dR + (size_t)i * 256); }
        }
        return v;
    };
    const std::vector<final_stereo_edge_pair> KF = mates(kf, n_kf, keyframe, descKfL, descKfR), CF = mates(cf, n_cf, current, descCfL, descCfR);

    // Pipeline.cpp:29-31: SpatialGrid(width, height, 15); Pipeline.cpp:153
    SpatialGrid gl(W, H, 15), gr(W, H, 15);
    engine.add_edges_to_spatial_grid(CF, gl, gr);

    TRes* R = new TRes; R->n_kf = n_kf;
    std::vector<KF_Temporal_Edge_Quads> quads;
    for (int i = 0; i < n_kf; ++i) {
        if (kf_mask && !kf_mask[i]) continue;
        KF_Temporal_Edge_Quads k;
        k.KF_stereo_mate = &KF[i];
        k.projected_orientation_left = k.projected_orientation_right = 0.0;
        k.veridical_quads.resize(1);
        quads.push_back(k);
        R->sel.push_back(i);
    }
    // Temporal_Matches.cpp:185-213, thresholds as written there
    engine.apply_spatial_grid_filtering_quads(quads, CF, gl, gr, 30.0);
    dump(*R, TQ_GRID, engine, quads);
    engine.apply_orientation_filtering_quads(quads, CF, 10.0);
    dump(*R, TQ_ORIENT, engine, quads);
    engine.apply_NCC_filtering_quads(quads, CF, 0.8, keyframe.left_image, keyframe.right_image, current.left_image, current.right_image);
    dump(*R, TQ_NCC, engine, quads);
    if (sift_on) engine.apply_SIFT_filtering_quads(quads, CF, 200.0);
    dump(*R, TQ_SIFT, engine, quads);
    engine.apply_best_nearly_best_filtering_quads(quads, 0.8, "NCC");
    dump(*R, TQ_BNB, engine, quads);
    if (sift_on) engine.apply_best_nearly_best_filtering_quads(quads, 0.8, "SIFT");
    dump(*R, TQ_BNB_SIFT, engine, quads);
    engine.apply_photometric_refinement_quads(quads, CF, keyframe, current);
    dump(*R, TQ_GN, engine, quads);
    engine.apply_temporal_edge_clustering_quads(quads, true);
    dump(*R, TQ_CLUSTER, engine, quads);
    return R;
}

int rt_stage_total(void* h, int st) { return ((TRes*)h)->st[st].off.back(); }
void rt_get_stage(void* h, int st, int* off, int* cf, double* l, double* r, double* ncc, double* sc, int* valid, double* sift)
{
    const QDump& d = ((TRes*)h)->st[st];
    std::memcpy(off, d.off.data(), d.off.size() * 4);
    std::memcpy(cf, d.cf.data(), d.cf.size() * 4); std::memcpy(valid, d.valid.data(), d.valid.size() * 4);
    std::memcpy(l, d.l.data(), d.l.size() * 8); std::memcpy(r, d.r.data(), d.r.size() * 8);
    std::memcpy(ncc, d.ncc.data(), d.ncc.size() * 8); std::memcpy(sc, d.sc.data(), d.sc.size() * 8);
    std::memcpy(sift, d.sift.data(), d.sift.size() * 8);
}
void rt_free(void* h) { delete (TRes*)h; }

}  // extern "C"
