// Drop-in replacement for the reference translation unit src/toed/cpu_toed.cpp.
//
// It implements the class DECLARED BY THE REFERENCE'S OWN HEADER (include/toed/cpu_toed.hpp:70-115, unmodified):
// same constructor, same get_Third_Order_Edges(cv::Mat), same public members (toed_edges, Total_Num_Of_TOED,
// time_conv, time_nms, omp_threads, subpix_edge_pts_final, edge_pt_list_idx, num_of_edge_data), so
// Pipeline::ProcessEdges (src/Pipeline.cpp:24-29) and test_third_order_edges.hpp compile and behave unchanged.
// The work is done on the GPU through the C ABI (include/ebvo_b200.h: ebvo_create / ebvo_toed); there is no CPU
// fallback - without a CUDA device the constructor reports the error the reference way (LOG_ERROR-style print)
// and every call yields an empty edge list.
//
// Build: add this file to the library instead of src/toed/cpu_toed.cpp and link libebvo_b200.so:
//   g++ -std=c++17 -I<reference>/include -I<ebvo-b200>/include -c dropin/cpu_toed_b200.cpp
#include <algorithm>
#include <cstdio>
#include <map>
#include <mutex>
#include <vector>

#include <opencv2/opencv.hpp>
#include "toed/cpu_toed.hpp"   // the reference header (leaks img()/Ix()/... macros: keep locals clear of those names)
#include "ebvo_b200.h"
#include "ebvo_dropin_common.hpp"   // the process-wide context shared with the matcher drop-ins (no second set of device buffers)

// edges per image the shared context is sized for: one per 8 input pixels (a dense synthetic scene yields one per 14), at
// least 65 536 (KITTI / EuRoC / ETH3D shapes), at most 2 M (the 4K stress shape); beyond it ebvo_toed reports EBVO_ERR_CAPACITY
static int edge_capacity(int H, int W) { return std::min(1 << 21, std::max(1 << 16, H * W / 8)); }

ThirdOrderEdgeDetectionCPU::ThirdOrderEdgeDetectionCPU(int H, int W)
{
    img_height = H;
    img_width = W;
    kernel_sz = 17;          // TOED_KERNEL_SIZE (definitions.h:76)
    shifted_kernel_sz = kernel_sz + 2;
    g_sig = 2;               // TOED_SIGMA (definitions.h:77)
    interp_img_height = H * 2;
    interp_img_width = W * 2;
    omp_threads = 0;         // no host threads are used
    time_conv = time_nms = 0.0;
    Total_Num_Of_TOED = 0;
    edge_pt_list_idx = 0;
    num_of_edge_data = 4;
    // the dense maps of the reference (cpu_toed.cpp:48-63) do not exist on this path
    img = Ix = Iy = I_grad_mag = I_orient = nullptr;
    subpix_pos_x_map = subpix_pos_y_map = subpix_grad_mag_map = nullptr;
    subpix_edge_pts_final = new double[(size_t)4 * 4 * H * W]();   // (x, y, theta, 0) rows of the LAST call's edges
    ebvo_dropin::Lease lease(W, H, edge_capacity(H, W));             // creates the shared context now, so that a missing GPU is reported here
}

ThirdOrderEdgeDetectionCPU::~ThirdOrderEdgeDetectionCPU()
{
    delete[] subpix_edge_pts_final;
}

void ThirdOrderEdgeDetectionCPU::get_Third_Order_Edges(cv::Mat image)
{
    toed_edges.clear();
    Total_Num_Of_TOED = 0;
    edge_pt_list_idx = 0;
    ebvo_dropin::Lease lease(img_width, img_height, edge_capacity(img_height, img_width));
    ebvo_ctx* c = lease.ctx;
    if (!c) return;
    if (image.rows != img_height || image.cols != img_width) {
        std::printf("\033[1;31m[ERROR] image size differs from the detector's (H, W)\033[0m\n");
        return;
    }
    const int cap = std::max(1 << 16, ebvo_dropin::shared().edges);      // the context's edge capacity (EBVO_ERR_CAPACITY beyond it)
    std::vector<ebvo_edge> out((size_t)cap);
    int n = 0, n_total = 0;
    const unsigned char* first = &image.at<unsigned char>(0, 0);
    const int stride = image.rows > 1 ? (int)(&image.at<unsigned char>(1, 0) - first) : image.cols;
    int rc = ebvo_toed(c, first, img_width, img_height, stride, out.data(), cap, &n, &n_total);
    if (rc != EBVO_OK) {
        std::printf("\033[1;31m[ERROR] ebvo_toed failed (%d): %s\033[0m\n", rc, ebvo_last_error(c));
        return;
    }
    toed_edges.reserve((size_t)n);
    Edge e;   // default: b_isEmpty = true, frame_source = -1, exactly what cpu_toed.cpp:527,557-563 leaves
    for (int k = 0; k < n; ++k) {
        e.location = cv::Point2d(out[k].x, out[k].y);
        e.orientation = out[k].theta;
        e.index = out[k].index;
        toed_edges.push_back(e);
        subpix_edge_pts_final[4 * (size_t)k + 0] = out[k].x;
        subpix_edge_pts_final[4 * (size_t)k + 1] = out[k].y;
        subpix_edge_pts_final[4 * (size_t)k + 2] = out[k].theta;
        subpix_edge_pts_final[4 * (size_t)k + 3] = 0.0;
    }
    edge_pt_list_idx = n_total;
    Total_Num_Of_TOED = n_total;   // unfiltered count, as cpu_toed.cpp:76,581
    // the class's profiling members (seconds, cpu_toed.cpp:366-368, 516-518) from the kernels' CUDA-event times: the dense kernel
    // fuses the convolution with the NMS tests, the sparse ones (ordering, FP64 sub-pixel fit) are what is left of the NMS stage
    const char* conv[] = {"toed_grad_nms"};
    const char* nms[] = {"toed_scan", "toed_expand", "toed_refine", "toed_prune"};
    time_conv = ebvo_dropin::kernel_ms(c, conv, 1) * 1e-3;
    time_nms = ebvo_dropin::kernel_ms(c, nms, 4) * 1e-3;
}

// The three stages are fused on the GPU; the reference's separate entry points stay callable.
void ThirdOrderEdgeDetectionCPU::preprocessing(cv::Mat) {}
void ThirdOrderEdgeDetectionCPU::convolve_img() {}
int ThirdOrderEdgeDetectionCPU::non_maximum_suppresion() { return Total_Num_Of_TOED; }
void ThirdOrderEdgeDetectionCPU::read_array_from_file(std::string, double*, int, int) {}
void ThirdOrderEdgeDetectionCPU::write_array_to_file(std::string, double*, int, int) {}
