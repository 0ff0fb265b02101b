// TEST INFRASTRUCTURE ONLY (oracle/): C entry point around the UNMODIFIED reference
// class ThirdOrderEdgeDetectionCPU (/root/reference/include/toed/cpu_toed.hpp:70-115,
// /root/reference/src/toed/cpu_toed.cpp:24-582). Linked with the reference source
// compiled in place; output goes to oracle/_ref/libtoed_ref.so (git-ignored).
// Used by tests/ (checker) and by bench.py's cpu_baseline / --impl reference leg.
#include <cstring>
#include <memory>
#include <omp.h>
#include <opencv2/opencv.hpp>          // the shim in third_party_shim
#include "toed/cpu_toed.hpp"           // the reference header, from /root/reference/include

extern "C" {

// Runs the reference detector once.
//   edges_xyt : capacity*3 doubles, receives (x, y, orientation) of toed_edges in order
//   all4      : optional (may be NULL) capacity_all*4 doubles: subpix_edge_pts_final rows
// Returns number of toed_edges (border-filtered); *n_total = Total_Num_Of_TOED.
int toed_ref_run(const unsigned char* img, int H, int W, int stride,
                 double* edges_xyt, int capacity, int* n_total,
                 double* all4, int capacity_all,
                 double* time_conv, double* time_nms, int omp_threads)
{
    ThirdOrderEdgeDetectionCPU det(H, W);
    if (omp_threads > 0) det.omp_threads = omp_threads;
    cv::Mat m(H, W, CV_8UC1, (void*)img, (size_t)stride);
    det.get_Third_Order_Edges(m);
    int n = (int)det.toed_edges.size();
    for (int k = 0; k < n && k < capacity; ++k) {
        edges_xyt[3 * k + 0] = det.toed_edges[k].location.x;
        edges_xyt[3 * k + 1] = det.toed_edges[k].location.y;
        edges_xyt[3 * k + 2] = det.toed_edges[k].orientation;
    }
    if (n_total) *n_total = det.Total_Num_Of_TOED;
    if (all4) {
        int m4 = det.Total_Num_Of_TOED < capacity_all ? det.Total_Num_Of_TOED : capacity_all;
        std::memcpy(all4, det.subpix_edge_pts_final, sizeof(double) * 4 * (size_t)m4);
    }
    if (time_conv) *time_conv = det.time_conv;
    if (time_nms) *time_nms = det.time_nms;
    return n;
}

// Persistent detector for timing (constructor cost excluded, as in Pipeline.h:89-101
// where the detector is built once and reused for every image).
void* toed_ref_create(int H, int W, int omp_threads)
{
    auto* d = new ThirdOrderEdgeDetectionCPU(H, W);
    if (omp_threads > 0) d->omp_threads = omp_threads;
    return d;
}
int toed_ref_detect(void* h, const unsigned char* img, int H, int W, int stride,
                    double* time_conv, double* time_nms)
{
    auto* d = static_cast<ThirdOrderEdgeDetectionCPU*>(h);
    cv::Mat m(H, W, CV_8UC1, (void*)img, (size_t)stride);
    d->get_Third_Order_Edges(m);
    if (time_conv) *time_conv = d->time_conv;
    if (time_nms) *time_nms = d->time_nms;
    return (int)d->toed_edges.size();
}
int toed_ref_fetch(void* h, double* edges_xyt, int capacity)
{
    auto* d = static_cast<ThirdOrderEdgeDetectionCPU*>(h);
    int n = (int)d->toed_edges.size();
    for (int k = 0; k < n && k < capacity; ++k) {
        edges_xyt[3 * k + 0] = d->toed_edges[k].location.x;
        edges_xyt[3 * k + 1] = d->toed_edges[k].location.y;
        edges_xyt[3 * k + 2] = d->toed_edges[k].orientation;
    }
    return n;
}
void toed_ref_destroy(void* h) { delete static_cast<ThirdOrderEdgeDetectionCPU*>(h); }
int toed_ref_num_procs() { return omp_get_num_procs(); }

}  // extern "C"
