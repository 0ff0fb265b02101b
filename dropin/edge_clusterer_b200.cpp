// Drop-in replacement for the reference translation unit src/EdgeClusterer.cpp: implements the class declared by
// the reference's own header (include/EdgeClusterer.h:30-63, unmodified) - constructor
// EdgeClusterer(std::vector<Edge>, std::vector<int>, bool), performClustering(), public returned_clusters /
// cluster_labels / clusters / Num_Of_Clusters / Epip_Correct_Edges - over the C ABI (ebvo_cluster,
// include/ebvo_b200.h).  Same merges, same Gaussian-weighted centres, same cluster order as
// src/EdgeClusterer.cpp:119-302.  No CPU fallback.
//
// One call clusters ONE candidate set (n <= 128 edges) with one warp, so this adapter exists for API compatibility
// (Temporal_Matches.cpp:658-665 and tests); inside the stereo path the GPU clusters every left edge's candidate set
// in one launch (cluster_kernel), which is where the speed comes from.
#include <algorithm>
#include <cstdio>
#include <map>
#include <numeric>
#include <vector>
#include <opencv2/opencv.hpp>
#include <Eigen/Dense>

#include "EdgeClusterer.h"       // the reference header
#include "ebvo_b200.h"
#include "ebvo_dropin_common.hpp"

EdgeClusterer::EdgeClusterer(std::vector<Edge> edge_set, std::vector<int> toed_indices, bool by_orientation)
    : Epip_Correct_Edges(edge_set), b_cluster_by_orientation(by_orientation), toed_indices_of_shifted_edges(toed_indices)
{
    Num_Of_Epipolar_Corrected_H2_Edges = (int)edge_set.size();
    Num_Of_Clusters = 0;
    H1_edge_idx = -1;
    cluster_labels.resize(edge_set.size());
    std::iota(cluster_labels.begin(), cluster_labels.end(), 0);     // EdgeClusterer.cpp:11-13
}

void EdgeClusterer::performClustering()
{
    const int n = Num_Of_Epipolar_Corrected_H2_Edges;
    returned_clusters.clear(); clusters.clear(); Num_Of_Clusters = 0;
    if (n == 0) return;
    ebvo_dropin::Lease lease(64, 64, 1024);
    ebvo_ctx* ctx = lease.ctx;
    if (!ctx) return;
    const std::vector<Edge> shifted_edges = Epip_Correct_Edges;
    std::vector<ebvo_edge> in((size_t)n), centers((size_t)n);
    for (int k = 0; k < n; ++k) in[k] = ebvo_edge{shifted_edges[k].location.x, shifted_edges[k].location.y, shifted_edges[k].orientation, k, 0};
    std::vector<int> labels((size_t)n);
    int ncl = 0;
    const int rc = ebvo_cluster(ctx, in.data(), n, b_cluster_by_orientation ? 1 : 0, centers.data(), labels.data(), &ncl);
    if (rc != EBVO_OK) {
        std::printf("\033[1;31m[ERROR] ebvo_cluster failed (%d): %s\033[0m\n", rc, ebvo_last_error(ctx));
        return;
    }
    // the reference keeps "label = original index of one member" (EdgeClusterer.cpp:131-204); the library returns the
    // renumbered labels 0..ncl-1 in ascending order of that label, which is what every later step uses (:275-286)
    std::vector<int> first_member((size_t)ncl, -1);
    clusters.assign((size_t)ncl, std::vector<int>());
    for (int k = 0; k < n; ++k) {
        if (first_member[labels[k]] < 0) first_member[labels[k]] = k;
        clusters[labels[k]].push_back(k);
    }
    for (int k = 0; k < n; ++k) cluster_labels[k] = labels[k];
    Num_Of_Clusters = (unsigned)ncl;
    returned_clusters.resize((size_t)ncl);
    for (int c = 0; c < ncl; ++c) {
        Edge centre{cv::Point2d(centers[c].x, centers[c].y), centers[c].theta, false, 0};      // :243
        returned_clusters[c].center_edge = centre;
        for (int k : clusters[c]) {
            Epip_Correct_Edges[k] = centre;                                                      // :262-267
            returned_clusters[c].contributing_edges.push_back(shifted_edges[k]);                 // :297 (toed indices not filled, :299)
        }
    }
}
