// SIFT descriptors at edge-derived keypoints on sm_100a (S4 / S7' / S13 of the stereo path, "SIFT-on").
//
// The reference calls cv::SIFT::compute(image, {kp+, kp-}, desc) once per edge with two hand-made keypoints
// (size 1, angle = deg(theta), octave 0) at p +- 8 (sin theta, -cos theta): augment_Edge_Data and apply_SIFT_filtering,
// src/Stereo_Matches.cpp:655-787, and finalize_stereo_edge_mates :1627-1635.  The arithmetic lives in OpenCV
// (modules/features2d/src/sift.dispatch.cpp / sift.simd.hpp, 4.x), not in the reference repository; it is restated
// here from its published algorithm and pinned against cv2 4.13 in tests/test_gpu_sift.py (99.97 % of the descriptor
// entries identical, the rest off by one: OpenCV's SIMD summation orders are not reproducible bit for bit):
//   * provided keypoints with octave 0 => firstOctave = 0: no image doubling; the descriptor image is
//     GaussianBlur(float(image), sigma = sqrt(1.6^2 - 0.5^2)), 13 taps, BORDER_REFLECT_101        (createInitialImage)
//   * calcSIFTDescriptor(img, pt, ori = 360 - angle, scl = size/2 = 0.5, d = 4, n = 8): radius 5 => 11 x 11 samples
//     around cvRound(pt), central-difference gradients, fastAtan2 polynomial, Gaussian weight exp(-(r'^2+c'^2)/8),
//     trilinear vote into a (d+2) x (d+2) x (n+2) histogram, circular fold of the orientation axis, clip at 0.2 |h|,
//     scale to 512 / |h|, saturate to 8 bits
//   * quirk kept: ori = 360 - angle exceeds 360 for negative angles, so floor(obin) can be below -n and the single
//     "o0 += n" wrap leaves a negative orientation index; the vote then lands 1..4 slots BEFORE the cell's first
//     bin in the flat histogram (the previous cell's upper slots), exactly as OpenCV's pointer arithmetic does;
//     votes that fall outside the array are dropped.
// Votes are accumulated in fixed point with shared-memory integer atomics, so the result does not depend on the order
// of the lanes (deterministic run to run).
#include "ebvo_internal.cuh"
#include <cfloat>
#include <cmath>

namespace ebvo {

constexpr int SK = 13, SR = 6;            // Gaussian taps / radius
__constant__ float c_sift_k[SK];

void upload_sift_tables()
{
    // cv::getGaussianKernel(13, sigma, CV_32F): exp(-x^2 / (2 sigma^2)) normalised in double, stored as float
    const double sigma = std::sqrt(std::max(1.6 * 1.6 - 0.5 * 0.5, 0.01));
    double k[SK], sum = 0;
    for (int i = 0; i < SK; ++i) { const double x = i - (SK - 1) * 0.5; k[i] = std::exp(-(x * x) / (2.0 * sigma * sigma)); sum += k[i]; }
    float kf[SK];
    for (int i = 0; i < SK; ++i) kf[i] = (float)(k[i] / sum);
    cudaMemcpyToSymbol(c_sift_k, kf, sizeof kf);
}

__device__ __forceinline__ int reflect101(int i, int n)   // BORDER_REFLECT_101; the clamp only matters for pixels that are never stored
{
    const int r = n == 1 ? 0 : (i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i));
    return min(max(r, 0), n - 1);
}

// 32 x 32 output tile per CTA (256 threads): u8 tile + 6 px halo -> row pass -> column pass, symmetric summation
// k[6] s[6] + sum_j k[6+j] (s[6+j] + s[6-j]) with fused multiply-adds (the order closest to OpenCV's: <= 3 float ulps)
__global__ void __launch_bounds__(256) sift_blur_kernel(DevBatch b)
{
    __shared__ float s_in[32 + 2 * SR][32 + 2 * SR + 1];
    __shared__ float s_row[32 + 2 * SR][33];
    const int img = blockIdx.z, x0 = blockIdx.x * 32, y0 = blockIdx.y * 32, tid = threadIdx.x;
    const uint8_t* I = b.und + (size_t)img * b.imgStride;
    for (int e = tid; e < (32 + 2 * SR) * (32 + 2 * SR); e += 256) {
        const int r = e / (32 + 2 * SR), c = e - r * (32 + 2 * SR);
        const int gy = reflect101(y0 - SR + r, b.H), gx = reflect101(x0 - SR + c, b.W);
        s_in[r][c] = (float)I[(size_t)gy * b.pitch + gx];
    }
    __syncthreads();
    for (int e = tid; e < (32 + 2 * SR) * 32; e += 256) {
        const int r = e >> 5, c = e & 31;
        const float* v = &s_in[r][c];
        float a = v[SR] * c_sift_k[SR];
#pragma unroll
        for (int j = 1; j <= SR; ++j) a = fmaf(v[SR + j] + v[SR - j], c_sift_k[SR + j], a);
        s_row[r][c] = a;
    }
    __syncthreads();
    float* out = b.blur + (size_t)img * b.blurStride;
    for (int e = tid; e < 32 * 32; e += 256) {
        const int r = e >> 5, c = e & 31;
        float a = s_row[r + SR][c] * c_sift_k[SR];
#pragma unroll
        for (int j = 1; j <= SR; ++j) a = fmaf(s_row[r + SR + j][c] + s_row[r + SR - j][c], c_sift_k[SR + j], a);
        if (y0 + r < b.H && x0 + c < b.W) out[(size_t)(y0 + r) * b.W + x0 + c] = a;
    }
}

// cv::hal::fastAtan2 (degrees), modules/core/src/mathfuncs_core.simd.hpp
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float p1 = 0.9997878412794807f * (float)(180 / 3.14159265358979323846), p3 = -0.3258083974640975f * (float)(180 / 3.14159265358979323846);
    const float p5 = 0.1555786518463281f * (float)(180 / 3.14159265358979323846), p7 = -0.04432655554792128f * (float)(180 / 3.14159265358979323846);
    const float ax = fabsf(x), ay = fabsf(y);
    float a;
    if (ax >= ay) {
        const float c = ay / (ax + (float)DBL_EPSILON), c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        const float c = ax / (ay + (float)DBL_EPSILON), c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

constexpr int SD = 4, SN = 8;                                  // SIFT_DESCR_WIDTH, SIFT_DESCR_HIST_BINS
constexpr int SHIST = (SD + 2) * (SD + 2) * (SN + 2);          // 360
constexpr float VFIX = 262144.f;                               // 2^18 fixed-point scale of the votes (a folded bin collects < 2^14: no overflow in 32 bits;
                                                               // 2^14 measurably lowers the agreement with cv2: 1.7e-3 of the entries differ instead of 3e-4)

// One warp per EDGE = two keypoints (side 0: p + 8 (sin, -cos), side 1: p - 8 (sin, -cos)) that share orientation, hence
// the rotated sampling pattern: which of the 11 x 11 samples fall inside the descriptor window, their bin coordinates and
// Gaussian weights are computed once and used for both sides.  The samples inside the window (46 % of the 121) are
// compacted first, so two passes of 32 lanes replace four; votes are 32-bit fixed-point shared-memory atomics (native
// ATOMS.ADD), so the histogram does not depend on the order of the lanes.
__global__ void __launch_bounds__(128) sift_desc_kernel(DevBatch b)
{
    __shared__ __align__(16) uint32_t s_h[4][2][SHIST];
    __shared__ uint8_t s_list[4][128];
    __shared__ float2 s_ij[128];               // sample k -> (i, j) = (k / 11 - 5, k % 11 - 5) as floats
    if (threadIdx.x < 121) s_ij[threadIdx.x] = make_float2((float)((int)threadIdx.x / 11 - 5), (float)((int)threadIdx.x % 11 - 5));
    __syncthreads();
    const int img = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = b.nE[img];
    const float* I = b.blur + (size_t)img * b.blurStride;
    const int rows = b.H, cols = b.W;
    const float bins_per_rad = SN / 360.f, exp_scale = -1.f / (SD * SD * 0.5f), hist_width = 3.f * 0.5f;
    // floor of |x| < 2^22 without the conversion unit: x + 1.5 * 2^23 holds round-to-nearest(x) in its low mantissa bits
    auto floor_if = [](float x, int& n) {
        const float t = x + 12582912.f;
        float f = t - 12582912.f;
        n = __float_as_int(t) - 0x4B400000;
        if (f > x) { f -= 1.f; --n; }
        return f;
    };
    // The scalar set-up of an edge (FP64 sincos for the keypoint offsets, the FP32 rotation) is done for SCH edges at once,
    // one edge per lane, and handed to the warp by shuffles: replicated over 32 lanes it was a fifth of the kernel.
    constexpr int SCH = 16;
    for (int e0 = (blockIdx.x * 4 + w) * SCH; e0 < n; e0 += gridDim.x * 4 * SCH) {
      const int cntE = min(SCH, n - e0);
      int l_ptx0 = 0, l_ptx1 = 0, l_pty0 = 0, l_pty1 = 0;
      float l_ori = 0.f, l_cos = 0.f, l_sin = 0.f;
      if (lane < cntE) {
          const size_t eo = (size_t)img * b.E + e0 + lane;
          const double x = b.ex[eo], y = b.ey[eo], th = b.eth[eo];
          double sn, cs;
          sincos(th, &sn, &cs);
          // utility.cpp:128-139 and the cv::KeyPoint(Point2d, 1, 180 / M_PI * theta) constructor (narrowing to float); cvRound
          l_ptx0 = __float2int_rn((float)(x + 8.0 * sn)); l_ptx1 = __float2int_rn((float)(x + 8.0 * (-sn)));
          l_pty0 = __float2int_rn((float)(y + 8.0 * (-cs))); l_pty1 = __float2int_rn((float)(y + 8.0 * cs));
          const float angle = (float)(180 / 3.14159265358979323846 * th);
          float ori = 360.f - angle;
          if (fabsf(ori - 360.f) < FLT_EPSILON) ori = 0.f;
          l_ori = ori;
          l_cos = cosf(ori * (float)(3.14159265358979323846 / 180)) / hist_width;
          l_sin = sinf(ori * (float)(3.14159265358979323846 / 180)) / hist_width;
      }
      for (int q = 0; q < cntE; ++q) {
        const size_t eo = (size_t)img * b.E + e0 + q;
        const int ptx[2] = {__shfl_sync(0xffffffffu, l_ptx0, q), __shfl_sync(0xffffffffu, l_ptx1, q)};
        const int pty[2] = {__shfl_sync(0xffffffffu, l_pty0, q), __shfl_sync(0xffffffffu, l_pty1, q)};
        const float ori = __shfl_sync(0xffffffffu, l_ori, q), cos_t = __shfl_sync(0xffffffffu, l_cos, q), sin_t = __shfl_sync(0xffffffffu, l_sin, q);
        {
            uint4* hz = reinterpret_cast<uint4*>(&s_h[w][0][0]);
            for (int k = lane; k < 2 * SHIST / 4; k += 32) hz[k] = make_uint4(0u, 0u, 0u, 0u);
        }
        // samples inside the descriptor window, in sample order
        int cnt = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = lane + 32 * t;
            const float2 ij = s_ij[k & 127];
            const float c_rot = ij.y * cos_t - ij.x * sin_t, r_rot = ij.y * sin_t + ij.x * cos_t;
            const float rbin = r_rot + SD / 2 - 0.5f, cbin = c_rot + SD / 2 - 0.5f;
            const bool in = k < 121 && rbin > -1 && rbin < SD && cbin > -1 && cbin < SD;
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (in) s_list[w][cnt + __popc(m & ((1u << lane) - 1u))] = (uint8_t)k;
            cnt += __popc(m);
        }
        __syncwarp();
        for (int p0 = 0; p0 < cnt; p0 += 32) {
            if (p0 + lane < cnt) {
                const int k = s_list[w][p0 + lane];
                const float2 ij = s_ij[k];
                const int i = k / 11 - 5, j = k - 11 * (i + 5) - 5;
                const float c_rot = ij.y * cos_t - ij.x * sin_t, r_rot = ij.y * sin_t + ij.x * cos_t;
                float rbin = r_rot + SD / 2 - 0.5f, cbin = c_rot + SD / 2 - 0.5f;
                const float wgt = expf((c_rot * c_rot + r_rot * r_rot) * exp_scale);
                int r0, c0;
                rbin -= floor_if(rbin, r0); cbin -= floor_if(cbin, c0);
                const int idx0 = ((r0 + 1) * (SD + 2) + c0 + 1) * (SN + 2);
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    const int r = pty[side] + i, c = ptx[side] + j;
                    if (r > 0 && r < rows - 1 && c > 0 && c < cols - 1) {
                        const float* Ip = I + (r * cols + c);
                        const float dx = Ip[1] - Ip[-1];
                        const float dy = Ip[-cols] - Ip[cols];
                        float obin = (fast_atan2_deg(dy, dx) - ori) * bins_per_rad;
                        const float mag = sqrtf(dx * dx + dy * dy) * wgt;
                        int o0;
                        obin -= floor_if(obin, o0);
                        if (o0 < 0) o0 += SN;
                        if (o0 >= SN) o0 -= SN;
                        const float v_r1 = mag * rbin, v_r0 = mag - v_r1;
                        const float v_rc11 = v_r1 * cbin, v_rc10 = v_r1 - v_rc11, v_rc01 = v_r0 * cbin, v_rc00 = v_r0 - v_rc01;
                        const float v111 = v_rc11 * obin, v110 = v_rc11 - v111, v101 = v_rc10 * obin, v100 = v_rc10 - v101;
                        const float v011 = v_rc01 * obin, v010 = v_rc01 - v011, v001 = v_rc00 * obin, v000 = v_rc00 - v001;
                        const int idx = idx0 + o0;
                        uint32_t* h = s_h[w][side] + idx;
                        auto fix = [](float v) { return __float2uint_rn(v * VFIX); };
                        if (idx >= 0) {        // the whole 2 x 2 x 2 cell block lies inside the histogram (its last slot is 287 + 71)
                            atomicAdd(h, fix(v000)); atomicAdd(h + 1, fix(v001));
                            atomicAdd(h + (SN + 2), fix(v010)); atomicAdd(h + (SN + 3), fix(v011));
                            atomicAdd(h + (SD + 2) * (SN + 2), fix(v100)); atomicAdd(h + (SD + 2) * (SN + 2) + 1, fix(v101));
                            atomicAdd(h + (SD + 3) * (SN + 2), fix(v110)); atomicAdd(h + (SD + 3) * (SN + 2) + 1, fix(v111));
                        } else {               // negative orientation index in the first cell: votes before the array are dropped
                            auto vote = [&](int o, float v) { if (idx + o >= 0 && idx + o < SHIST) atomicAdd(h + o, fix(v)); };
                            vote(0, v000); vote(1, v001);
                            vote(SN + 2, v010); vote(SN + 3, v011);
                            vote((SD + 2) * (SN + 2), v100); vote((SD + 2) * (SN + 2) + 1, v101);
                            vote((SD + 3) * (SN + 2), v110); vote((SD + 3) * (SN + 2) + 1, v111);
                        }
                    }
                }
            }
        }
        __syncwarp();
        // circular fold + copy: lane owns descriptor entries 4 lane .. 4 lane + 3 (cell lane / 2, orientations 4 (lane & 1) + t)
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const int cell = lane >> 1, ob = 4 * (lane & 1);
            const uint32_t* h = s_h[w][side] + ((cell / SD + 1) * (SD + 2) + (cell % SD + 1)) * (SN + 2) + ob;
            float val[4];
            float nrm2 = 0.f;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                uint32_t sacc = h[t];
                if (ob == 0 && t < 2) sacc += h[SN + t];
                val[t] = __uint2float_rn(sacc) * (1.f / VFIX);
                nrm2 += val[t] * val[t];
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) nrm2 += __shfl_xor_sync(0xffffffffu, nrm2, o);
            const float thr = sqrtf(nrm2) * 0.2f;                                   // SIFT_DESCR_MAG_THR
            float n2 = 0.f;
#pragma unroll
            for (int t = 0; t < 4; ++t) { val[t] = fminf(val[t], thr); n2 += val[t] * val[t]; }
#pragma unroll
            for (int o = 16; o; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
            const float scale = 512.f / fmaxf(sqrtf(n2), FLT_EPSILON);              // SIFT_INT_DESCR_FCTR
            uint32_t pk = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) pk |= (uint32_t)min(max(__float2int_rn(val[t] * scale), 0), 255) << (8 * t);   // saturate_cast<uchar>
            reinterpret_cast<uint32_t*>(b.desc8 + (eo * 2 + side) * 128)[lane] = pk;
        }
        __syncwarp();
      }
    }
}

void launch_sift_desc(const DevBatch& b, int nImages, cudaStream_t st, Prof* prof)   // descriptors only: the blurred images exist already
{
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    int gx = (sms * 16 + nImages - 1) / nImages;
    if (gx < 1) gx = 1;
    EBVO_KERNEL(prof, "sift_desc", st, (sift_desc_kernel<<<dim3(gx, nImages), 128, 0, st>>>(b)));
}

void launch_sift(const DevBatch& b, int nImages, cudaStream_t st, Prof* prof)
{
    dim3 gB((b.W + 31) / 32, (b.H + 31) / 32, nImages);
    EBVO_KERNEL(prof, "sift_blur", st, (sift_blur_kernel<<<gB, 256, 0, st>>>(b)));
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    int gx = (sms * 16 + nImages - 1) / nImages;
    if (gx < 1) gx = 1;
    EBVO_KERNEL(prof, "sift_desc", st, (sift_desc_kernel<<<dim3(gx, nImages), 128, 0, st>>>(b)));
}

}  // namespace ebvo
