// Drop-in replacement for the two heavy members of the reference class Stereo_Matches
//   Frame_Evaluation_Metrics Stereo_Matches::get_Stereo_Edge_Pairs(Dataset::Ptr, Stereo_Edge_Pairs&, size_t, Timing_Statistics&)
//   void Stereo_Matches::finalize_stereo_edge_mates(Stereo_Edge_Pairs&, std::vector<final_stereo_edge_pair>&)
// (reference src/Stereo_Matches.cpp:1360-1540 and :1578-1653), compiled against the reference's OWN headers
// (include/Stereo_Matches.h, Dataset.h, Stereo_Iterator.h, toed/cpu_toed.hpp - unmodified), so the call sites in
// Pipeline::get_Stereo_Edge_Correspondences (src/Pipeline.cpp:116-131) compile and behave unchanged.  The work is
// done on the GPU through the C ABI (include/ebvo_b200.h: ebvo_stereo_match_full - ONE call per frame).  No CPU fallback.
//
// How a maintainer links it (no reference source is edited):
//   * add this file and libebvo_b200.so to the library;
//   * compile src/Stereo_Matches.cpp with
//       -Dget_Stereo_Edge_Pairs=get_Stereo_Edge_Pairs_cpu -Dfinalize_stereo_edge_mates=finalize_stereo_edge_mates_cpu
//     (CMake: set_source_files_properties(Stereo_Matches.cpp PROPERTIES COMPILE_DEFINITIONS "...")), which keeps the
//     CPU bodies under other names and leaves every other member (Find_Stereo_GT_Locations, get_Stereo_Edge_GT_Pairs,
//     the writers, the individual apply_* filters) exactly as it is.
// dropin/Makefile does precisely this with the reference sources in place and tests/test_gpu_dropin.py runs the result.
//
// Scope: the no-GT branch (KITTI, EuRoC, ETH3D-SLAM: Dataset.cpp:120-148).  The SIFT gate, BNB-SIFT and the
// descriptors of the finalised mates are computed on the device (cv::SIFT::compute at the reference's keypoints,
// restated in csrc/sift.cu) as in the reference's default flow; EBVO_DROPIN_SIFT=0 selects the SIFT-off parity
// configuration (descriptor pairs stay empty).  With has_gt() the per-stage Evaluate_Stereo_Edge_Correspondences
// metrics (diagnostics) are not produced: the returned Frame_Evaluation_Metrics is empty, the mates are the same.
//
// State left in Stereo_Edge_Pairs, as after the reference's remove_empty_clusters (:1543-1576): only matched left
// edges remain; per remaining edge one EdgeCluster whose center_edge is the mate, refine_final_scores = {NCC}.
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
#include <opencv2/opencv.hpp>
#include <Eigen/Dense>

#include "Stereo_Matches.h"      // the reference header
#include "ebvo_b200.h"
#include "ebvo_dropin_common.hpp"

namespace {

void log_error(const char* what, ebvo_ctx* c, int rc)
{
    std::printf("\033[1;31m[ERROR] %s failed (%d): %s\033[0m\n", what, rc, c ? ebvo_last_error(c) : "");
}

ebvo_calib calib_of(Dataset& d)
{
    ebvo_calib c;
    const Eigen::Matrix3d Kl = d.get_left_calib_matrix(), Kr = d.get_right_calib_matrix(), R = d.get_relative_rot_left_to_right();
    const Eigen::Vector3d T = d.get_relative_transl_left_to_right();
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) { c.Kl[3 * i + j] = Kl(i, j); c.Kr[3 * i + j] = Kr(i, j); c.R21[3 * i + j] = R(i, j); }
        c.T21[i] = T(i);
    }
    return c;    // F21 is derived inside the library exactly as Dataset.cpp:102-112 does
}

// two 1 x 128 CV_32F rows, as cv::SIFT::compute returns them (descriptors.row(0), descriptors.row(1))
std::pair<cv::Mat, cv::Mat> descriptor_pair(const float* d256)
{
    cv::Mat a(1, 128, CV_32F), b(1, 128, CV_32F);
    for (int k = 0; k < 128; ++k) { a.at<float>(0, k) = d256[k]; b.at<float>(0, k) = d256[128 + k]; }
    return {a, b};
}

std::pair<cv::Mat, cv::Mat> patch_pair(const float* plus49, const float* minus49)
{
    cv::Mat p(PATCH_SIZE, PATCH_SIZE, CV_32F), m(PATCH_SIZE, PATCH_SIZE, CV_32F);       // utility.cpp:190-191
    for (int i = 0; i < PATCH_SIZE; ++i)
        for (int j = 0; j < PATCH_SIZE; ++j) { p.at<float>(i, j) = plus49[i * PATCH_SIZE + j]; m.at<float>(i, j) = minus49[i * PATCH_SIZE + j]; }
    return {p, m};
}

}  // namespace

Frame_Evaluation_Metrics Stereo_Matches::get_Stereo_Edge_Pairs(Dataset::Ptr dataset, Stereo_Edge_Pairs& pairs, size_t frame_idx, Timing_Statistics& timing_statistics)
{
    (void)frame_idx;
    Frame_Evaluation_Metrics frame_metrics;
    const StereoFrame& f = *pairs.stereo_frame;
    const int W = f.left_image.cols, H = f.left_image.rows;
    const size_t nF = pairs.focused_edge_indices.size();
    const int nR = (int)f.right_edges.size();
    const int cap = (int)std::max<size_t>(std::max<size_t>(nF, (size_t)nR), 1);
    ebvo_dropin::Trace tr("get_Stereo_Edge_Pairs");
    ebvo_dropin::Lease lease(W, H, cap);
    ebvo_ctx* ctx = lease.ctx;
    tr.mark("lease");

    // what the reference's stages leave behind even when nothing survives
    auto fail_empty = [&]() {
        pairs.focused_edge_indices.clear(); pairs.GT_locations_from_left_edges.clear(); pairs.veridical_right_edges_indices.clear();
        pairs.Gamma_in_left_cam_coord.clear(); pairs.Gamma_in_right_cam_coord.clear(); pairs.left_edge_descriptors.clear();
        pairs.epip_line_coeffs_of_left_edges.clear(); pairs.left_edge_patches.clear(); pairs.matching_edge_clusters.clear();
        return frame_metrics;
    };
    if (!ctx) return fail_empty();

    // inputs: raw images for the NCC stages (:562-563), undistorted ones for Gauss-Newton and finalisation (:1293, :1580).
    // The four cv::Mat are handed over as they are when they share one row step (the usual case: continuous matrices);
    // otherwise tightly packed copies are made.
    const cv::Mat* im[4] = {&f.left_image, &f.right_image, &f.left_image_undistorted, &f.right_image_undistorted};
    const unsigned char* ptr[4];
    std::vector<unsigned char> packed[4];
    size_t step = f.left_image.step;
    bool same = true;
    for (int k = 0; k < 4; ++k) same = same && (size_t)im[k]->step == step && im[k]->cols == W && im[k]->rows == H;
    for (int k = 0; k < 4; ++k) {
        if (same) ptr[k] = im[k]->data;
        else { packed[k] = ebvo_dropin::packed_u8(im[k]->data, H, W, im[k]->step); ptr[k] = packed[k].data(); }
    }
    if (!same) step = (size_t)W;
    std::vector<ebvo_edge> L(nF), R((size_t)nR);
    for (size_t i = 0; i < nF; ++i) {                      // the focused left edges, in Stereo_Edge_Pairs order (:192-198)
        const Edge& e = f.left_edges[pairs.focused_edge_indices[i]];
        L[i] = ebvo_edge{e.location.x, e.location.y, e.orientation, pairs.focused_edge_indices[i], e.frame_source};
    }
    for (int k = 0; k < nR; ++k) {
        const Edge& e = f.right_edges[k];
        R[k] = ebvo_edge{e.location.x, e.location.y, e.orientation, k, e.frame_source};
    }
    const ebvo_calib calib = calib_of(*dataset);
    // ONE call: matching, then - on the images already on the device - the left patches of the matched edges from the RAW left
    // image (apply_NCC_Filtering, :570-576), their descriptor pairs (augment_Edge_Data, :655-689), and what
    // finalize_stereo_edge_mates will ask for: the mates' patches from the UNDISTORTED right image (:1580-1582, :1622) and
    // their descriptor pairs (:1627-1635)
    const bool sift = ebvo_dropin::sift_enabled();
    const size_t capM = std::max<size_t>(nF, 1);
    ebvo_dropin::Shared& S = ebvo_dropin::shared();
    ebvo_dropin::FinalizeCache& fin = S.fin;
    fin = ebvo_dropin::FinalizeCache();
    ebvo_mate* mates = static_cast<ebvo_mate*>(S.mates.ensure(capM * sizeof(ebvo_mate)));
    float *pp = S.l_plus.floats(capM * 49), *pm = S.l_minus.floats(capM * 49), *dl = sift ? S.l_desc.floats(capM * 256) : nullptr;
    float *rp = S.r_plus.floats(capM * 49), *rm = S.r_minus.floats(capM * 49), *rd = sift ? S.r_desc.floats(capM * 256) : nullptr;
    if (!mates || !pp || !pm || !rp || !rm || (sift && (!dl || !rd))) { std::printf("\033[1;31m[ERROR] out of host memory for the result buffers\033[0m\n"); return fail_empty(); }
    int n = 0;
    tr.mark("inputs + host buffers");
    int rc = ebvo_stereo_match_full(ctx, &calib, ptr[0], ptr[1], ptr[2], ptr[3], W, H, (int)step, L.data(), (int)nF, R.data(), nR,
                                    mates, (int)capM, &n, pp, pm, rp, rm, dl, rd);
    tr.mark("ebvo_stereo_match_full");
    if (rc != EBVO_OK) { log_error("ebvo_stereo_match_full", ctx, rc); fin = ebvo_dropin::FinalizeCache(); return fail_empty(); }
    if (n > 0) {
        fin.frame = pairs.stereo_frame; fin.n = (size_t)n; fin.has_desc = sift;
        fin.x0 = mates[0].rx; fin.y0 = mates[0].ry; fin.x1 = mates[n - 1].rx; fin.y1 = mates[n - 1].ry;
    }
    {   // Timing_Statistics (Stereo_Matches.h:32-47; the reference's own assignments are commented out at :1376-1538): kernel
        // milliseconds of this call.  The epipolar, disparity and orientation gates are ONE fused kernel: its time is under time_EP.
        auto ms = [&](std::initializer_list<const char*> p) { return ebvo_dropin::kernel_ms(ctx, p.begin(), (int)p.size()); };
        timing_statistics.time_EP = ms({"sobel", "bounds", "gate"});
        timing_statistics.time_DP = 0.0; timing_statistics.time_OR = 0.0;
        timing_statistics.time_SIFT = ms({"sift_"});
        timing_statistics.time_NCC = ms({"patch", "ncc_bnb"});
        timing_statistics.time_BNB_NCC = 0.0; timing_statistics.time_BNB_SIFT = 0.0;       // inside ncc_bnb
        timing_statistics.time_Refinement = ms({"shift", "gn"});
        timing_statistics.time_Clustering = ms({"cluster"});
        timing_statistics.time_Post_NCC = ms({"ncc2_best"});
        timing_statistics.time_Best = 0.0;                                                  // inside ncc2_best
        timing_statistics.time_Finalize = ms({"compact"});
        timing_statistics.total_time = timing_statistics.time_EP + timing_statistics.time_SIFT + timing_statistics.time_NCC +
                                       timing_statistics.time_Refinement + timing_statistics.time_Clustering +
                                       timing_statistics.time_Post_NCC + timing_statistics.time_Finalize;
    }

    // rebuild the per-left-edge containers for the survivors, in left-edge order (the order remove_empty_clusters keeps)
    const Eigen::Matrix3d F21 = dataset->get_fund_mat_21();
    std::vector<int> focused((size_t)n);
    std::vector<cv::Point2d> gt_loc((size_t)n);
    std::vector<std::vector<int>> veridical((size_t)n);
    std::vector<Eigen::Vector3d> g_left((size_t)n), g_right((size_t)n), lines((size_t)n);
    std::vector<std::pair<cv::Mat, cv::Mat>> desc((size_t)n), patches((size_t)n);
    std::vector<Stereo_Matching_Edge_Clusters> clusters((size_t)n);
    // (the per-mate containers - cv::Mat pairs, EdgeCluster - are what the reference's data model costs on the host: built in parallel)
#pragma omp parallel for schedule(static)
    for (int k = 0; k < n; ++k) {
        const int i = mates[k].left_index;                 // position in the focused list handed to the matcher
        focused[k] = pairs.focused_edge_indices[i];
        if ((size_t)i < pairs.GT_locations_from_left_edges.size()) gt_loc[k] = pairs.GT_locations_from_left_edges[i];
        if ((size_t)i < pairs.veridical_right_edges_indices.size()) veridical[k] = pairs.veridical_right_edges_indices[i];
        if ((size_t)i < pairs.Gamma_in_left_cam_coord.size()) g_left[k] = pairs.Gamma_in_left_cam_coord[i];
        if ((size_t)i < pairs.Gamma_in_right_cam_coord.size()) g_right[k] = pairs.Gamma_in_right_cam_coord[i];
        const Eigen::Vector3d x(L[i].x, L[i].y, 1.0);     // CalculateEpipolarLine (:10-20)
        lines[k] = F21 * x;
        patches[k] = patch_pair(&pp[(size_t)k * 49], &pm[(size_t)k * 49]);
        if (sift) desc[k] = descriptor_pair(&dl[(size_t)k * 256]);
        EdgeCluster ec;
        ec.center_edge = Edge(cv::Point2d(mates[k].rx, mates[k].ry), mates[k].rtheta, false, 0);   // :39 / EdgeClusterer.cpp:243
        ec.center_edge.index = -1;                         // uninitialised in the reference (cpu_toed.hpp:35)
        ec.contributing_edges.push_back(ec.center_edge);
        ec.paired_left_edge_index = focused[k];
        clusters[k].edge_clusters.push_back(ec);
        clusters[k].refine_final_scores.push_back(mates[k].score);
        clusters[k].refine_confidences.push_back(0.0);
        clusters[k].refine_validities.push_back(true);
    }
    pairs.focused_edge_indices.swap(focused);
    pairs.GT_locations_from_left_edges.swap(gt_loc);
    pairs.veridical_right_edges_indices.swap(veridical);
    pairs.Gamma_in_left_cam_coord.swap(g_left);
    pairs.Gamma_in_right_cam_coord.swap(g_right);
    pairs.left_edge_descriptors.swap(desc);                // sized like the reference (:657-658); empty pairs when SIFT is off
    pairs.epip_line_coeffs_of_left_edges.swap(lines);
    pairs.left_edge_patches.swap(patches);
    pairs.matching_edge_clusters.swap(clusters);
    tr.mark("containers");
    return frame_metrics;
}

void Stereo_Matches::finalize_stereo_edge_mates(Stereo_Edge_Pairs& pairs, std::vector<final_stereo_edge_pair>& final_stereo_edge_pairs)
{
    const size_t n = pairs.focused_edge_indices.size();
    // the reference's consistency check (:1585-1603)
    if (n != pairs.matching_edge_clusters.size() || n != pairs.Gamma_in_left_cam_coord.size() || n != pairs.Gamma_in_right_cam_coord.size() ||
        n != pairs.left_edge_patches.size() || n != pairs.left_edge_descriptors.size() || n != pairs.GT_locations_from_left_edges.size()) {
        std::printf("\033[1;31m[ERROR] Vector sizes are not consistent in finalize_stereo_edge_mates\033[0m\n");
        return;
    }
    ebvo_dropin::Trace tr("finalize_stereo_edge_mates");
    final_stereo_edge_pairs.clear();
    final_stereo_edge_pairs.resize(n);
    tr.mark("resize");
    if (n == 0) { std::cout << "Size of finalized stereo edge pairs = 0" << std::endl; return; }

    // right patches (UNDISTORTED right image, :1580-1582, :1622) and right descriptor pairs (:1627-1635) of every mate: the
    // matching call computed them while the images were on the device; they are recomputed here only when this is not the
    // result that call produced (another frame, or clusters edited in between)
    const cv::Mat& Rimg = pairs.stereo_frame->right_image_undistorted;
    const int W = Rimg.cols, H = Rimg.rows;
    std::vector<ebvo_edge> Rm(n);
    for (size_t i = 0; i < n; ++i) {
        const Edge& e = pairs.matching_edge_clusters[i].edge_clusters[0].center_edge;
        Rm[i] = ebvo_edge{e.location.x, e.location.y, e.orientation, (int)i, 0};
    }
    ebvo_dropin::Lease lease(W, H, (int)n);
    ebvo_ctx* ctx = lease.ctx;
    if (!ctx) { final_stereo_edge_pairs.clear(); return; }
    const bool sift = ebvo_dropin::sift_enabled();
    ebvo_dropin::Shared& S = ebvo_dropin::shared();
    const ebvo_dropin::FinalizeCache& fin = S.fin;
    const bool cached = fin.frame == pairs.stereo_frame && fin.n == n && fin.x0 == Rm[0].x && fin.y0 == Rm[0].y && fin.x1 == Rm[n - 1].x &&
                        fin.y1 == Rm[n - 1].y && S.r_plus.cap >= n * 49 * sizeof(float) && (!sift || (fin.has_desc && S.r_desc.cap >= n * 256 * sizeof(float)));
    std::vector<float> pp_own, pm_own, dr_own;
    const float *pp = static_cast<const float*>(S.r_plus.p), *pm = static_cast<const float*>(S.r_minus.p), *dr = sift ? static_cast<const float*>(S.r_desc.p) : nullptr;
    if (!cached) {
        pp_own.resize(n * 49); pm_own.resize(n * 49);
        const std::vector<unsigned char> Rund = ebvo_dropin::packed_u8(Rimg.data, H, W, Rimg.step);
        int rc = ebvo_edge_patches(ctx, Rund.data(), W, H, W, Rm.data(), (int)n, pp_own.data(), pm_own.data());
        if (rc != EBVO_OK) { log_error("ebvo_edge_patches", ctx, rc); final_stereo_edge_pairs.clear(); return; }
        if (sift) {
            dr_own.resize(n * 256);
            rc = ebvo_sift_descriptors(ctx, Rund.data(), W, H, W, Rm.data(), (int)n, dr_own.data());
            if (rc != EBVO_OK) { log_error("ebvo_sift_descriptors", ctx, rc); final_stereo_edge_pairs.clear(); return; }
        }
        pp = pp_own.data(); pm = pm_own.data(); dr = sift ? dr_own.data() : nullptr;
    }

    tr.mark("right patches + descriptors");
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        final_stereo_edge_pair mate;
        mate.left_edge = pairs.get_focused_edge_by_Stereo_Edge_Pairs_index(i);
        mate.right_edge = pairs.matching_edge_clusters[i].edge_clusters[0].center_edge;
        mate.left_edge_patches = pairs.left_edge_patches[i];
        mate.right_edge_patches = patch_pair(&pp[i * 49], &pm[i * 49]);
        mate.left_edge_descriptors = pairs.left_edge_descriptors[i];
        if (dr) mate.right_edge_descriptors = descriptor_pair(&dr[i * 256]);
        mate.Gamma_in_left_cam_coord = pairs.Gamma_in_left_cam_coord[i];
        mate.Gamma_in_right_cam_coord = pairs.Gamma_in_right_cam_coord[i];
        mate.gt_right_location = pairs.GT_locations_from_left_edges[i];
        mate.b_is_TP = cv::norm(mate.right_edge.location - pairs.GT_locations_from_left_edges[i]) <= DIST_TO_GT_THRESH;   // :1645
        final_stereo_edge_pairs[i] = mate;
    }
    tr.mark("final_stereo_edge_pair objects");
    std::cout << "Size of finalized stereo edge pairs = " << final_stereo_edge_pairs.size() << std::endl;
}
