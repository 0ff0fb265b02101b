// Drop-in replacement for the heavy member of the reference class Temporal_Matches
//   Frame_Evaluation_Metrics Temporal_Matches::get_Temporal_Edge_Pairs_from_Quads(std::vector<KF_Temporal_Edge_Quads>&,
//       const std::vector<final_stereo_edge_pair>& KF, const std::vector<final_stereo_edge_pair>& CF, const SpatialGrid&,
//       const SpatialGrid&, Stereo_Edge_Pairs&, Stereo_Edge_Pairs&, const StereoFrame& keyframe, const StereoFrame& current, size_t, size_t)
// (reference src/Temporal_Matches.cpp:168-218), compiled against the reference's OWN headers (include/Temporal_Matches.h,
// Dataset.h - unmodified), so the call site in Pipeline::get_Temporal_Edge_Correspondences (src/Pipeline.cpp:159-167)
// compiles and behaves unchanged.  The filter chain (spatial grid, orientation, NCC, best-nearly-best, 2-D Gauss-Newton,
// clustering) runs on the GPU through the C ABI (include/ebvo_b200.h: ebvo_temporal_quads).  No CPU fallback.
//
// How a maintainer links it (no reference source is edited): add this file and libebvo_b200.so, and compile
// src/Temporal_Matches.cpp with -Dget_Temporal_Edge_Pairs_from_Quads=get_Temporal_Edge_Pairs_from_Quads_cpu, which keeps
// the CPU body under another name and every other member (add_edges_to_spatial_grid, build_Veridical_Quads, the
// individual apply_* filters, the writers) as it is.  dropin/Makefile does this with the reference source in place.
//
// Scope: the SIFT gate and the SIFT best-nearly-best pass run when every mate carries descriptor pairs (the stereo drop-in's
// default flow fills them); with a mate lacking them the reference's min_sift returns 900 for it (:482-483), which fails the
// 200 gate for every quad, whereas this drop-in then runs SIFT-off (EBVO_DROPIN_SIFT=0 selects that explicitly).  The per-stage Evaluate_Temporal_Edge_Pairs_on_Quads metrics
// (ground-truth diagnostics, has_gt() only) are not produced: the returned Frame_Evaluation_Metrics is empty.
// Which keyframe mates take part is read from the argument exactly as the reference does: those whose
// veridical_quads list (built on the host by build_Veridical_Quads from ground-truth poses) is non-empty (:345).
// The spatial grids passed in are not read: the library bins the current frame's mates itself (GRID_SIZE cells).
//
// State left behind, as after apply_temporal_edge_clustering_quads: candidate_cluster_pairs_[g] holds the surviving
// (left, right) Temporal_CF_Edge_Cluster pairs of group g and quads[g].candidate_quads points at them.
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

#include "Temporal_Matches.h"    // the reference header
#include "ebvo_b200.h"
#include "ebvo_dropin_common.hpp"

Frame_Evaluation_Metrics Temporal_Matches::get_Temporal_Edge_Pairs_from_Quads(
    std::vector<KF_Temporal_Edge_Quads>& quads, const std::vector<final_stereo_edge_pair>& KF,
    const std::vector<final_stereo_edge_pair>& CF, const SpatialGrid& left_spatial_grids, const SpatialGrid& right_spatial_grids,
    Stereo_Edge_Pairs& last_keyframe_stereo, Stereo_Edge_Pairs& current_frame_stereo, const StereoFrame& keyframe,
    const StereoFrame& current_frame, size_t keyframe_idx, size_t current_frame_idx)
{
    (void)right_spatial_grids; (void)last_keyframe_stereo; (void)current_frame_stereo; (void)keyframe_idx; (void)current_frame_idx;
    Frame_Evaluation_Metrics frame_metrics;
    candidate_cluster_pairs_.assign(quads.size(), {});        // :340
    const int W = current_frame.left_image.cols, H = current_frame.left_image.rows;
    const int n_kf = (int)KF.size(), n_cf = (int)CF.size();
    if (quads.empty() || n_kf == 0) return frame_metrics;
    ebvo_dropin::Trace tr("get_Temporal_Edge_Pairs_from_Quads");
    ebvo_dropin::Lease lease(W, H, 1);
    ebvo_ctx* ctx = lease.ctx;
    if (!ctx) return frame_metrics;

    std::vector<unsigned char> mask((size_t)n_kf, 0);
    std::vector<int> group_of((size_t)n_kf, -1);
    for (size_t g = 0; g < quads.size(); ++g) {
        const ptrdiff_t i = quads[g].KF_stereo_mate - KF.data();
        if (i < 0 || i >= n_kf) continue;
        group_of[(size_t)i] = (int)g;
        if (!quads[g].veridical_quads.empty()) mask[(size_t)i] = 1;      // :345
    }
    auto mates = [](const std::vector<final_stereo_edge_pair>& v) {
        std::vector<ebvo_mate> m(v.size());
        for (size_t i = 0; i < v.size(); ++i)
            m[i] = ebvo_mate{(int)i, 0, v[i].left_edge.location.x, v[i].left_edge.location.y, v[i].left_edge.orientation,
                             v[i].right_edge.location.x, v[i].right_edge.location.y, v[i].right_edge.orientation, 0.0};
        return m;
    };
    const std::vector<ebvo_mate> kf = mates(KF), cf = mates(CF);
    auto pack = [&](const cv::Mat& m) { return ebvo_dropin::packed_u8(m.data, H, W, m.step); };
    const std::vector<unsigned char> kL = pack(keyframe.left_image), kLu = pack(keyframe.left_image_undistorted), kRu = pack(keyframe.right_image_undistorted);
    const std::vector<unsigned char> cL = pack(current_frame.left_image), cLu = pack(current_frame.left_image_undistorted), cRu = pack(current_frame.right_image_undistorted);
    ebvo_quad_params qp{left_spatial_grids.cell_size, 0, 30.0, 10.0, 0.8, 0.8, 200.0};     // thresholds as written at :185-196
    // descriptor pairs of the mates (filled by the stereo stage's drop-in in its default SIFT-on flow): all present => SIFT-on.
    // They are staged in page-locked buffers the drop-in keeps (4 x 26 MB per KITTI-shape pair; fresh zero-filled vectors and
    // pageable uploads cost more than the kernels).
    ebvo_dropin::Shared& S = ebvo_dropin::shared();
    auto descs = [](const std::vector<final_stereo_edge_pair>& v, bool right, ebvo_dropin::HostBuf& hb, const float*& ptr) {
        ptr = nullptr;
        if (v.empty()) return false;
        float* out = hb.floats(v.size() * 256);
        if (!out) return false;
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
        for (long long i = 0; i < (long long)v.size(); ++i) {
            const auto& p = right ? v[(size_t)i].right_edge_descriptors : v[(size_t)i].left_edge_descriptors;
            if (p.first.empty() || p.second.empty() || p.first.cols != 128 || p.second.cols != 128) { bad |= 1; continue; }
            std::memcpy(out + (size_t)i * 256, &p.first.at<float>(0, 0), 128 * sizeof(float));
            std::memcpy(out + (size_t)i * 256 + 128, &p.second.at<float>(0, 0), 128 * sizeof(float));
        }
        ptr = out;
        return bad == 0;
    };
    const float *dkl = nullptr, *dkr = nullptr, *dcl = nullptr, *dcr = nullptr;
    const bool sift_on = ebvo_dropin::sift_enabled() && descs(KF, false, S.tq_desc[0], dkl) && descs(KF, true, S.tq_desc[1], dkr) &&
                         descs(CF, false, S.tq_desc[2], dcl) && descs(CF, true, S.tq_desc[3], dcr);
    tr.mark("inputs");
    // result records: sized from the previous call (a keyframe mate keeps 4 quads on average, 128 at most); a call that needs
    // more says how many and is repeated once
    size_t cap = std::max<size_t>(std::max<size_t>((size_t)n_kf * 8, S.tq_last + S.tq_last / 2), 1024);
    cap = std::min<size_t>(cap, (size_t)n_kf * 128);
    int n = 0, rc = EBVO_OK;
    ebvo_quad* out = nullptr;
    for (int attempt = 0; attempt < 2; ++attempt) {
        out = static_cast<ebvo_quad*>(S.quads.ensure(cap * sizeof(ebvo_quad)));
        if (!out) { std::printf("\033[1;31m[ERROR] out of host memory for the quad records\033[0m\n"); return frame_metrics; }
        rc = ebvo_temporal_quads(ctx, kL.data(), kLu.data(), kRu.data(), cL.data(), cLu.data(), cRu.data(), W, H, W, kf.data(), n_kf,
                                 mask.data(), cf.data(), n_cf, sift_on ? dkl : nullptr, sift_on ? dkr : nullptr,
                                 sift_on ? dcl : nullptr, sift_on ? dcr : nullptr, &qp, out, (int)cap, &n);
        if (rc == EBVO_ERR_CAPACITY && (size_t)n > cap && attempt == 0) { cap = (size_t)n; continue; }
        break;
    }
    tr.mark("ebvo_temporal_quads");
    if (rc != EBVO_OK) {
        std::printf("\033[1;31m[ERROR] ebvo_temporal_quads failed (%d): %s\033[0m\n", rc, ebvo_last_error(ctx));
        return frame_metrics;
    }
    S.tq_last = (size_t)n;
    size_t num_quads = 0;
    for (const auto& kvq : quads) num_quads += kvq.veridical_quads.size();
    std::cout << "Veridical quads: " << quads.size() << " KF groups, " << num_quads << " total quads" << std::endl;     // :182
    for (int k = 0; k < n; ++k) {
        const ebvo_quad& q = out[k];
        const int g = group_of[(size_t)q.kf_index];
        if (g < 0) continue;
        Temporal_CF_Edge_Cluster l, r;
        l.cf_stereo_edge_mate_index = r.cf_stereo_edge_mate_index = q.cf_index;
        l.contributing_cf_stereo_indices = r.contributing_cf_stereo_indices = {q.cf_index};
        l.center_edge = CF[(size_t)q.cf_index].left_edge; r.center_edge = CF[(size_t)q.cf_index].right_edge;
        l.center_edge.location = cv::Point2d(q.lx, q.ly); l.center_edge.orientation = q.ltheta;
        r.center_edge.location = cv::Point2d(q.rx, q.ry); r.center_edge.orientation = q.rtheta;
        l.matching_scores = scores{q.ncc_left, q.sift_left}; r.matching_scores = scores{q.ncc_right, q.sift_right};
        l.refine_final_score = q.score_left; r.refine_final_score = q.score_right;
        l.refine_validity = r.refine_validity = q.valid != 0;
        candidate_cluster_pairs_[(size_t)g].emplace_back(std::move(l), std::move(r));
    }
    for (size_t g = 0; g < quads.size(); ++g) {
        quads[g].candidate_quads.clear();
        for (auto& p : candidate_cluster_pairs_[g]) quads[g].candidate_quads.push_back({&p.first, &p.second});
    }
    tr.mark("containers");
    return frame_metrics;
}
