// Uses the detector exactly like Pipeline::ProcessEdges (src/Pipeline.cpp:24-29) and the reference's own
// test harness (test/test_include/test_third_order_edges.hpp:13-17): construct with (H, W), call
// get_Third_Order_Edges(cv::Mat), copy toed_edges.  Reads a raw u8 image, prints "n n_total" then x y theta rows.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>
#include <opencv2/opencv.hpp>
#include "toed/cpu_toed.hpp"

int main(int argc, char** argv)
{
    if (argc < 4) return 2;
    const int H = std::atoi(argv[2]), W = std::atoi(argv[3]);
    std::vector<unsigned char> buf((size_t)H * W);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(buf.data(), 1, buf.size(), f) != buf.size()) return 3;
    std::fclose(f);
    cv::Mat image(H, W, CV_8UC1, buf.data(), (size_t)W);
    ThirdOrderEdgeDetectionCPU::Ptr TOED = ThirdOrderEdgeDetectionCPU::Ptr(new ThirdOrderEdgeDetectionCPU(H, W));
    std::vector<Edge> edges;
    for (int rep = 0; rep < 2; ++rep) {          // the pipeline reuses one detector for the left and right image
        TOED->get_Third_Order_Edges(image);
        edges = TOED->toed_edges;
    }
    if (argc > 4) {   // timing mode: wall clock of Pipeline::ProcessEdges' two statements, argv[4] repetitions
        const int reps = std::atoi(argv[4]);
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; ++r) { TOED->get_Third_Order_Edges(image); edges = TOED->toed_edges; }
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
        std::printf("{\"what\": \"ThirdOrderEdgeDetectionCPU::get_Third_Order_Edges (GPU drop-in), one image\", \"H\": %d, \"W\": %d, \"edges\": %zu, "
                    "\"wall_ms\": %.3f, \"time_conv_ms\": %.3f, \"time_nms_ms\": %.3f}\n", H, W, edges.size(), ms, TOED->time_conv * 1e3, TOED->time_nms * 1e3);
        return 0;
    }
    std::printf("%zu %d\n", edges.size(), TOED->Total_Num_Of_TOED);
    for (const Edge& e : edges) std::printf("%.17g %.17g %.17g %d\n", e.location.x, e.location.y, e.orientation, e.index);
    return 0;
}
