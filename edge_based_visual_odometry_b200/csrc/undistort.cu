// cv::undistort for 8-bit single-channel images on sm_100a (Pipeline::prepare_Stereo_Images, reference
// src/Pipeline.cpp:78-79: cv::undistort(image, undistorted, K, distCoeffs) with the 4 coefficients k1 k2 p1 p2 of the YAMLs).
//
// The arithmetic lives in OpenCV (modules/calib3d/src/undistort.dispatch.cpp, modules/imgproc/src/imgwarp.cpp), not in
// the reference repository; it is restated here from its published algorithm and pinned bit for bit against cv2 4.13
// (tests/test_gpu_ops.py):
//   * initUndistortRectifyMap(A, dist, I, Ar, CV_16SC2) per stripe of min(max(1, 4096 / cols), rows) rows, Ar = A with
//     cy shifted by the stripe origin: normalised (x, y) of the destination pixel, radial + tangential model in FP64,
//     u = fx x_d + cx, v = fy y_d + cy, fixed point iu = cvRound(32 u), iv = cvRound(32 v)
//   * remap(INTER_LINEAR, BORDER_CONSTANT 0) in fixed point: integer cell (iu >> 5, iv >> 5), 5-bit fractions, bilinear
//     weights 32 (32 - fy)(32 - fx) ... (the entries of OpenCV's BilinearTab_i, which are exact), result
//     (sum + 2^14) >> 15; taps outside the image read 0.
// One thread per destination pixel; HBM-bound (1 B written, ~4 B gathered through L2 per pixel).
#include "ebvo_internal.cuh"

namespace ebvo {

__global__ void __launch_bounds__(256) undistort_kernel(const uint8_t* __restrict__ src, int srcPitch, uint8_t* __restrict__ dst, int dstPitch,
                                                        int W, int H, double fx, double fy, double cx, double cy, double k1, double k2, double p1, double p2, int stripe)
{
    const int j = blockIdx.x * 32 + (threadIdx.x & 31), r = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (j >= W || r >= H) return;
    const int y0 = (r / stripe) * stripe, i = r - y0;
    // iR = (Ar I)^-1 with Ar = [fx 0 cx; 0 fy cy - y0; 0 0 1]
    const double ir0 = 1.0 / fx, ir2 = -cx / fx, ir4 = 1.0 / fy, ir5 = -(cy - (double)y0) / fy;
    const double x = (double)j * ir0 + ir2, y = (double)i * ir4 + ir5;
    const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
    const double kr = (1 + ((0 * r2 + k2) * r2 + k1) * r2) / 1.0;
    const double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2), yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy;
    const double u = fx * xd + cx, v = fy * yd + cy;
    const int iu = __double2int_rn(u * 32.0), iv = __double2int_rn(v * 32.0);     // saturate_cast<int> = cvRound
    const int sx = (int)(short)(iu >> 5), sy = (int)(short)(iv >> 5), qx = iu & 31, qy = iv & 31;
    auto tap = [&](int yy, int xx) { return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (int)src[(size_t)yy * srcPitch + xx] : 0; };
    const int acc = tap(sy, sx) * ((32 - qy) * (32 - qx) * 32) + tap(sy, sx + 1) * ((32 - qy) * qx * 32) +
                    tap(sy + 1, sx) * (qy * (32 - qx) * 32) + tap(sy + 1, sx + 1) * (qy * qx * 32);
    dst[(size_t)r * dstPitch + j] = (uint8_t)min(max((acc + (1 << 14)) >> 15, 0), 255);
}

void launch_undistort(const uint8_t* src, int srcPitch, uint8_t* dst, int dstPitch, int W, int H, const double K[9], const double dist[4], cudaStream_t st)
{
    const int stripe = std::min(std::max(1, (1 << 12) / std::max(W, 1)), H);
    dim3 g((W + 31) / 32, (H + 7) / 8);
    undistort_kernel<<<g, 256, 0, st>>>(src, srcPitch, dst, dstPitch, W, H, K[0], K[4], K[2], K[5], dist[0], dist[1], dist[2], dist[3], stripe);
}

}  // namespace ebvo
