// Third-order edge detection on sm_100a.
//
// Replaces ThirdOrderEdgeDetectionCPU::{preprocessing, convolve_img, non_maximum_suppresion}
// (reference src/toed/cpu_toed.cpp:82-120, 122-376, 386-582).  The dense pass (which samples are edges) is FP32; the
// values that reach the edge list (sub-pixel position, orientation) are re-evaluated in FP64 for the survivors.
//
// Work-efficient split (same results as the reference's dense evaluation):
//   K_A  toed_grad_nms   dense, tile-fused: uint8 tile + halo -> separable G/Gx row pass (3 tap variants) ->
//                         column pass for fx, fy on the four 2x sub-grids -> gradient magnitude -> octant NMS +
//                         parabola sub-pixel fit (cpu_toed.cpp:400-514) -> per-row ballot bitmask (+ sparse
//                         sub-pixel offsets).  No interp-grid map is ever written to HBM.
//   K_B  toed_scan/expand row counts -> exclusive scan -> ordered (i,j) list.  Order = row-major interp order,
//                         exactly the serial scan of cpu_toed.cpp:530-575, so edge indices match the reference.
//   K_C  toed_refine      sparse, FP64: gradient at the 3 x 3 samples around every surviving edge sample, sub-pixel fit,
//                         the seven remaining third-order responses and the orientation (cpu_toed.cpp:224-229, 418-510)
//                         are re-evaluated in double only where they reach the edge list (about 2% of the interp grid).
#include "ebvo_internal.cuh"
#include <algorithm>
#include <cmath>
#include <mutex>
#include <cuda.h>
#include <math_constants.h>

namespace ebvo {

// [variant][filter][tap]; variant 0 = unshifted, middle 17 taps (ends zero); 1 = unshifted 19; 2 = shifted 19
// filter 0..3 = G, Gx, Gxx, Gxxx (closed forms quoted at cpu_toed.cpp:137-140,151-154; sigma = 2)
__constant__ float c_T[3][4][19];
__constant__ double c_T64[3][4][19];     // the same tables in double (FP64 refinement of the surviving samples)

void upload_toed_tables()
{
    float T[3][4][19];
    double T64[3][4][19];
    const double sig = 2.0, s2 = sig * sig, c = std::sqrt(2.0 * 3.14159265358979323846);
    for (int v = 0; v < 3; ++v)
        for (int p = -9; p <= 9; ++p) {
            double s = p + (v == 2 ? 0.5 : 0.0), e = std::exp(-s * s / (2.0 * s2));
            double G = e / (c * sig), Gx = (-s * e) / (c * sig * s2), Gxx = ((s * s - s2) * e) / (c * sig * s2 * s2);
            double Gxxx = ((s * (3.0 * s2 - s * s)) * e) / (c * sig * s2 * s2 * s2);
            bool zero = (v == 0 && (p == -9 || p == 9));
            T[v][0][p + 9] = zero ? 0.f : (float)G;
            T[v][1][p + 9] = zero ? 0.f : (float)Gx;
            T[v][2][p + 9] = zero ? 0.f : (float)Gxx;
            T[v][3][p + 9] = zero ? 0.f : (float)Gxxx;
            T64[v][0][p + 9] = zero ? 0.0 : G; T64[v][1][p + 9] = zero ? 0.0 : Gx; T64[v][2][p + 9] = zero ? 0.0 : Gxx; T64[v][3][p + 9] = zero ? 0.0 : Gxxx;
        }
    cudaMemcpyToSymbol(c_T, T, sizeof(T));
    cudaMemcpyToSymbol(c_T64, T64, sizeof(T64));
}


constexpr int IN_WP = IN_W + 1;  // padded strides (floats)
constexpr int OWP = OW + 1;
constexpr int IWP = IW + 1;
constexpr int CR = 5;                           // output rows per thread in the column pass
constexpr int NSEG = (OH + CR - 1) / CR;        // 7
// Shared memory: the input tile + six row-pass planes are dead once the column pass has its sums in registers, and the
// three interp-grid planes (Ix, Iy, magnitude) are only written after that point, so the two groups ALIAS (one extra
// barrier): 56 KB instead of 111 KB per CTA => 4 CTAs (32 warps) per SM instead of 2, which is what hides the barriers
// between the four stages and the global-load latency of stage 0.
constexpr size_t TOED_SMEM_A = sizeof(float) * (IN_H * IN_WP + 6 * IN_H * OWP);
constexpr size_t TOED_SMEM_B = sizeof(float) * (3 * IH * IWP);
constexpr size_t TOED_SMEM = TOED_SMEM_A > TOED_SMEM_B ? TOED_SMEM_A : TOED_SMEM_B;

// ---- TMA plumbing for the input tile -----------------------------------------------------------------------
// The u8 tile + halo is fetched by ONE bulk tensor copy (cp.async.bulk.tensor.3d, SASS UTMALDG) from a tensor map over
// the whole image array [image][row][column]: the hardware clips the 64 x 52 box against the image and fills what lies
// outside with zeros, which is exactly the zero padding of the reference's convolution (cpu_toed.cpp:204-205), so the
// per-pixel bounds tests and the scalar gather loop disappear.  Completion is signalled on an mbarrier.
constexpr int TMA_BOX_W = 64;                         // box width in bytes (a multiple of 16 covering IN_W = 52 plus the alignment shift)
constexpr int TMA_XPAD = 16;                          // the box starts at x0 - 16, not x0 - HALO: its first byte must be 16-byte aligned
static_assert(TW % 16 == 0 && TMA_XPAD >= HALO && TMA_XPAD - HALO + IN_W <= TMA_BOX_W, "TMA box must cover the tile + halo");
constexpr int TMA_BYTES = TMA_BOX_W * IN_H;           // 3328
constexpr size_t TOED_U8_OFF = 11264;                 // 128-byte aligned offset inside the (not yet written) row-pass planes
static_assert(TOED_U8_OFF >= sizeof(float) * IN_H * IN_WP && TOED_U8_OFF % 128 == 0, "u8 tile must not overlap s_in");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(TOED_THREADS, 4) toed_grad_nms_kernel(DevBatch b, const __grid_constant__ CUtensorMap tmap, float magThresh, int border)
{
    extern __shared__ __align__(128) float smem[];
    float* s_in = smem;
    float* s_row = s_in + IN_H * IN_WP;  // planes: 0 G17, 1 Gx17, 2 G19, 3 Gx19, 4 Gs, 5 Gxs
    float* s_ix = smem;                  // aliases s_in / s_row (see TOED_SMEM)
    float* s_iy = s_ix + IH * IWP;
    float* s_mag = s_iy + IH * IWP;
    __shared__ int s_tot;

    __shared__ __align__(8) unsigned long long s_bar;
    const int tid = threadIdx.x, img = blockIdx.z;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    if (tid == 0) s_tot = 0;

    // ---- stage 0: uint8 tile + halo by TMA (zero fill outside the image), then u8 -> float ----
    const uint8_t* s_u8 = reinterpret_cast<const uint8_t*>(smem) + TOED_U8_OFF;
    const uint32_t bar = smem_u32(&s_bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TMA_BYTES) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(s_u8)), "l"(&tmap), "r"(x0 - TMA_XPAD), "r"(y0 - HALO), "r"(b.imgBase + img), "r"(bar) : "memory");
    }
    {   // wait for the bytes (phase 0); a bounded spin: a mis-programmed copy traps instead of hanging the GPU
        uint32_t done = 0;
        for (int spin = 0; !done && spin < (1 << 24); ++spin)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar) : "memory");
        if (!done) __trap();
    }
    for (int it = tid; it < IN_H * IN_W; it += TOED_THREADS) {
        const int r = it / IN_W, c = it - r * IN_W;
        s_in[r * IN_WP + c] = (float)s_u8[r * TMA_BOX_W + (TMA_XPAD - HALO) + c];
    }
    __syncthreads();

    // ---- stage 1: row pass (x direction).  Output column c <-> image column x0-1+c; taps v[k] = column c+k,
    //      tap q = 9-k, coefficient index q+9 = 18-k.
    for (int it = tid; it < IN_H * OW; it += TOED_THREADS) {
        int r = it / OW, c = it - r * OW;
        const float* v = s_in + r * IN_WP + c;
        float g17 = 0.f, gx17 = 0.f, gs = 0.f, gxs = 0.f;
        float v0 = v[0], v18 = v[18];
#pragma unroll
        for (int k = 1; k <= 17; ++k) {
            float x = v[k];
            g17 = fmaf(x, c_T[1][0][18 - k], g17);
            gx17 = fmaf(x, c_T[1][1][18 - k], gx17);
            gs = fmaf(x, c_T[2][0][18 - k], gs);
            gxs = fmaf(x, c_T[2][1][18 - k], gxs);
        }
        gs = fmaf(v0, c_T[2][0][18], gs);   gs = fmaf(v18, c_T[2][0][0], gs);
        gxs = fmaf(v0, c_T[2][1][18], gxs); gxs = fmaf(v18, c_T[2][1][0], gxs);
        float g19 = fmaf(v0, c_T[1][0][18], fmaf(v18, c_T[1][0][0], g17));
        float gx19 = fmaf(v0, c_T[1][1][18], fmaf(v18, c_T[1][1][0], gx17));
        int o = r * OWP + c;
        s_row[0 * IN_H * OWP + o] = g17;
        s_row[1 * IN_H * OWP + o] = gx17;
        s_row[2 * IN_H * OWP + o] = g19;
        s_row[3 * IN_H * OWP + o] = gx19;
        s_row[4 * IN_H * OWP + o] = gs;
        s_row[5 * IN_H * OWP + o] = gxs;
    }
    __syncthreads();

    // ---- stage 2: column pass for fx, fy on the four sub-grids; CR output rows per thread ----
    const int seg = tid / OW, c = tid - seg * OW;
    const int o0 = seg * CR;
    float fx00[CR], fy00[CR], fx01[CR], fy01[CR], fx10[CR], fy10[CR], fx11[CR], fy11[CR];
#pragma unroll
    for (int k = 0; k < CR; ++k) fx00[k] = fy00[k] = fx01[k] = fy01[k] = fx10[k] = fy10[k] = fx11[k] = fy11[k] = 0.f;
    if (tid < OW * NSEG) {
#pragma unroll
        for (int rr = 0; rr < CR + 18; ++rr) {
            int row = o0 + rr;
            if (row < IN_H) {
                int o = row * OWP + c;
                float rg17 = s_row[0 * IN_H * OWP + o], rgx17 = s_row[1 * IN_H * OWP + o];
                float rg19 = s_row[2 * IN_H * OWP + o], rgx19 = s_row[3 * IN_H * OWP + o];
                float rgs = s_row[4 * IN_H * OWP + o], rgxs = s_row[5 * IN_H * OWP + o];
#pragma unroll
                for (int oo = 0; oo < CR; ++oo) {
                    const int t = rr - oo;  // smem row o+t of output row o  <->  tap p = 9-t, coefficient 18-t
                    if (t >= 0 && t <= 18) {
                        if (t >= 1 && t <= 17) {
                            fx00[oo] = fmaf(rgx17, c_T[1][0][18 - t], fx00[oo]);   // Gx(x) * G(y), 17 taps
                            fy00[oo] = fmaf(rg17, c_T[1][1][18 - t], fy00[oo]);    // G(x) * Gx(y)
                        }
                        fx01[oo] = fmaf(rgxs, c_T[1][0][18 - t], fx01[oo]);        // x shifted, y unshifted 19
                        fy01[oo] = fmaf(rgs, c_T[1][1][18 - t], fy01[oo]);
                        fx10[oo] = fmaf(rgx19, c_T[2][0][18 - t], fx10[oo]);       // x unshifted 19, y shifted
                        fy10[oo] = fmaf(rg19, c_T[2][1][18 - t], fy10[oo]);
                        fx11[oo] = fmaf(rgxs, c_T[2][0][18 - t], fx11[oo]);        // both shifted
                        fy11[oo] = fmaf(rgs, c_T[2][1][18 - t], fy11[oo]);
                    }
                }
            }
        }
    }
    __syncthreads();   // every thread is done reading s_row: the interp planes may now overwrite it
    if (tid < OW * NSEG) {
#pragma unroll
        for (int oo = 0; oo < CR; ++oo) {
            int o = o0 + oo;
            if (o < OH) {
                int li = 2 * o, lj = 2 * c;
                s_ix[li * IWP + lj] = fx00[oo];           s_iy[li * IWP + lj] = fy00[oo];
                s_mag[li * IWP + lj] = sqrtf(fx00[oo] * fx00[oo] + fy00[oo] * fy00[oo]);
                s_ix[li * IWP + lj + 1] = fx01[oo];       s_iy[li * IWP + lj + 1] = fy01[oo];
                s_mag[li * IWP + lj + 1] = sqrtf(fx01[oo] * fx01[oo] + fy01[oo] * fy01[oo]);
                s_ix[(li + 1) * IWP + lj] = fx10[oo];     s_iy[(li + 1) * IWP + lj] = fy10[oo];
                s_mag[(li + 1) * IWP + lj] = sqrtf(fx10[oo] * fx10[oo] + fy10[oo] * fy10[oo]);
                s_ix[(li + 1) * IWP + lj + 1] = fx11[oo]; s_iy[(li + 1) * IWP + lj + 1] = fy11[oo];
                s_mag[(li + 1) * IWP + lj + 1] = sqrtf(fx11[oo] * fx11[oo] + fy11[oo] * fy11[oo]);
            }
        }
    }
    __syncthreads();

    // ---- stage 3: NMS + sub-pixel fit on the 64x64 interior; one warp per interp row, two 32-wide halves ----
    const int warp = tid >> 5, lane = tid & 31;
    uint32_t* mask = b.mask + (size_t)img * b.maskStride;
    int* rowcnt = b.rowcnt + (size_t)img * b.rowStride;
    int nall = 0;
    for (int rr = 0; rr < 8; ++rr) {
        const int li = 2 + warp * 8 + rr;
        const int gi = 2 * y0 + (li - 2);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int lj = 2 + 32 * h + lane;
            const int gj = 2 * x0 + (lj - 2);
            // `edge` counts the samples the reference's tests accept (FP32 evaluation: n_total); `keep` is a slightly WIDER set
            // (every threshold relaxed by far more than the FP32 error of the sums) that goes to toed_refine, where the same
            // tests are decided in FP64: the edge list is the reference's set, not the FP32 approximation of it
            bool edge = false, keep = false;
            if (gi >= border && gi < b.H2 - border && gj >= border && gj < b.W2 - border) {
                const float* M = s_mag + li * IWP + lj;
                float m = M[0], gx = s_ix[li * IWP + lj], gy = s_iy[li * IWP + lj];
                if (m > magThresh - 1e-3f && !(fabsf(gx) < 10e-6f && fabsf(gy) < 10e-6f)) {
                    float nx = gx / m, ny = gy / m, slope, fp, fm;
                    // octant table, cpu_toed.cpp:418-477 (M[+-IWP] = row i+-1, M[+-1] = column j+-1)
                    if (gx >= 0.f && gy >= 0.f) {
                        if (gx >= gy) { slope = ny / nx; fp = M[1] * (1 - slope) + M[IWP + 1] * slope; fm = M[-1] * (1 - slope) + M[-IWP - 1] * slope; }
                        else { slope = nx / ny; fp = M[IWP] * (1 - slope) + M[IWP + 1] * slope; fm = M[-IWP] * (1 - slope) + M[-IWP - 1] * slope; }
                    } else if (gx < 0.f && gy >= 0.f) {
                        if (fabsf(gx) < gy) { slope = -nx / ny; fp = M[IWP] * (1 - slope) + M[IWP - 1] * slope; fm = M[-IWP] * (1 - slope) + M[-IWP + 1] * slope; }
                        else { slope = -ny / nx; fp = M[-1] * (1 - slope) + M[IWP - 1] * slope; fm = M[1] * (1 - slope) + M[-IWP + 1] * slope; }
                    } else if (gx < 0.f && gy < 0.f) {
                        if (fabsf(gx) >= fabsf(gy)) { slope = ny / nx; fp = M[-1] * (1 - slope) + M[-IWP - 1] * slope; fm = M[1] * (1 - slope) + M[IWP + 1] * slope; }
                        else { slope = nx / ny; fp = M[-IWP] * (1 - slope) + M[-IWP - 1] * slope; fm = M[IWP] * (1 - slope) + M[IWP + 1] * slope; }
                    } else {
                        if (gx < fabsf(gy)) { slope = -nx / ny; fp = M[-IWP] * (1 - slope) + M[-IWP + 1] * slope; fm = M[IWP] * (1 - slope) + M[IWP - 1] * slope; }
                        else { slope = -ny / nx; fp = M[1] * (1 - slope) + M[-IWP + 1] * slope; fm = M[-1] * (1 - slope) + M[IWP - 1] * slope; }
                    }
                    const float tol = 1e-4f * m + 1e-4f;
                    if (m >= fm - tol && m >= fp - tol) {
                        const bool strict = m > magThresh && ((m > fm && m > fp) || (m > fm && m >= fp) || (m >= fm && m > fp));
                        float s = sqrtf(1.f + slope * slope);
                        float A = (fm + fp - 2.f * m) / (2.f * s * s), B = (fp - fm) / (2.f * s);
                        float ss = -B / (2.f * A);
                        if (fabsf(ss) <= 1.41421356237f + 1e-2f) {
                            edge = strict && fabsf(ss) <= 1.41421356237f;
                            const float dx = ss * nx, dy = ss * ny;
                            double X = ((double)gj + (double)dx - 1.0) * 0.5, Y = ((double)gi + (double)dy - 1.0) * 0.5;
                            // border filter of cpu_toed.cpp:553-554 (relaxed by 0.01 px; decided in toed_refine)
                            keep = (X > (double)border - 1e-2) && (X < (double)(b.W - border) + 1e-2) && (Y > (double)border - 1e-2) && (Y < (double)(b.H - border) + 1e-2);
                        }
                    }
                }
            }
            unsigned all = __ballot_sync(0xffffffffu, edge), kp = __ballot_sync(0xffffffffu, keep);
            nall += __popc(all);
            if (lane == 0) {
                mask[(size_t)gi * b.maskPitch + ((2 * x0 + 32 * h) >> 5)] = kp;
                if (kp && gi < b.H2) atomicAdd(&rowcnt[gi], __popc(kp));
            }
        }
    }
    if (lane == 0 && nall) atomicAdd(&s_tot, nall);
    __syncthreads();
    if (tid == 0 && s_tot) atomicAdd(&b.nTot[img], s_tot);
}

// ---- K_B: exclusive scan of the per-row counts (one CTA per image) -------------------------------------
__global__ void __launch_bounds__(1024) toed_scan_kernel(DevBatch b)
{
    const int img = blockIdx.x, tid = threadIdx.x;
    const int* rowcnt = b.rowcnt + (size_t)img * b.rowStride;
    int* rowoff = b.rowoff + (size_t)img * b.rowStride;
    const int per = (b.H2 + 1023) / 1024;
    const int r0 = tid * per;
    int local = 0;
    for (int k = 0; k < per; ++k) { int r = r0 + k; if (r < b.H2) local += rowcnt[r]; }
    __shared__ int s_w[32];
    int v = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, d); if ((tid & 31) >= d) v += t; }
    if ((tid & 31) == 31) s_w[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
        int w = s_w[tid];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, w, d); if (tid >= d) w += t; }
        s_w[tid] = w;
    }
    __syncthreads();
    int excl = v - local + ((tid >> 5) ? s_w[(tid >> 5) - 1] : 0);
    for (int k = 0; k < per; ++k) { int r = r0 + k; if (r < b.H2) { rowoff[r] = excl; excl += rowcnt[r]; } }
    if (tid == 1023) {
        int total = excl;
        rowoff[b.H2] = total;
        if (total > b.E) { atomicExch(b.errFlag + (img >> 1), 1); total = b.E; }
        b.nE[img] = total;
    }
}

// ---- K_B: expand the bitmask into the ordered (i,j) list; one warp per interp row ----------------------
__global__ void __launch_bounds__(256) toed_expand_kernel(DevBatch b)
{
    const int img = blockIdx.y, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= b.H2) return;
    const int* rowoff = b.rowoff + (size_t)img * b.rowStride;
    int base = rowoff[row];
    if (rowoff[row + 1] == base) return;
    const uint32_t* mask = b.mask + (size_t)img * b.maskStride + (size_t)row * b.maskPitch;
    uint32_t* coords = b.coords + (size_t)img * b.E;
    for (int w0 = 0; w0 < b.maskPitch; w0 += 32) {
        int w = w0 + lane;
        uint32_t bits = (w < b.maskPitch) ? mask[w] : 0u;
        int cnt = __popc(bits), incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        int pos = base + incl - cnt;
        while (bits) {
            int bit = __ffs(bits) - 1;
            bits &= bits - 1;
            if (pos < b.E) coords[pos] = ((uint32_t)row << 16) | (uint32_t)(w * 32 + bit);
            ++pos;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// ---- K_C: FP64 refinement of the surviving edge samples; one warp per edge -------------------------------
// The dense kernel decides WHICH interp samples are edges in FP32 (2 % of the grid survive); this kernel re-evaluates, for
// those samples only, everything that reaches the edge list in FP64 and in the reference's own formulas: the gradient
// (fx, fy) at the sample and its 8 neighbours (cpu_toed.cpp:200-222), the octant interpolation and parabola fit
// (:418-510), the seven third-order responses and the orientation (:224-229).  The edge list then carries the
// reference's FP64 values (measured |dx|, |dy| < 1e-11 px, |dtheta| < 1e-11 rad against the FP64 CPU restatement instead of 5e-5 px /
// 3e-5 rad for the FP32 values), so the matcher downstream sees the same input as the CPU path and the 0.1 % budget for
// near-threshold flips is not spent on detector noise.  Separable form per edge: lane = image row (21 rows x 21
// columns cover the 3 x 3 samples' 19 x 19 supports): 1-D row sums for the three sample columns (G, Gx; the 17-tap and
// 19-tap variants of sub-grid (0,0) share their middle taps) and Gxx, Gxxx for the centre column -> shared memory ->
// lane = (sample, response): 19-tap column sums -> every lane evaluates the closed forms, lane 0 writes.
template <int B, int A>   // B = gj & 1, A = gi & 1 (x / y phase of the sample): fix where the three sample columns sit in the 21-pixel row
__device__ __forceinline__ void refine_row_pass(const double (&px)[21], double* out /* 14 */)
{
    // column dj = -1, 0, +1: centre offset oc in px[], x phase b'
    constexpr int OC[3] = {9, B ? 9 : 10, 10};
    constexpr int BP[3] = {B ? 0 : 1, B ? 1 : 0, B ? 0 : 1};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (BP[c]) {           // shifted tables, 19 taps
            double g = 0, gx = 0;
#pragma unroll
            for (int q = -9; q <= 9; ++q) { g = fma(px[OC[c] - q], c_T64[2][0][q + 9], g); gx = fma(px[OC[c] - q], c_T64[2][1][q + 9], gx); }
            out[4 * c + 0] = g; out[4 * c + 1] = gx;
        } else {               // unshifted: 17 taps for sample rows of sub-grid (0,0), 19 taps for sub-grid (1,0)
            double g = 0, gx = 0;
#pragma unroll
            for (int q = -8; q <= 8; ++q) { g = fma(px[OC[c] - q], c_T64[1][0][q + 9], g); gx = fma(px[OC[c] - q], c_T64[1][1][q + 9], gx); }
            out[4 * c + 0] = g; out[4 * c + 1] = gx;
            out[4 * c + 2] = fma(px[OC[c] - 9], c_T64[1][0][18], fma(px[OC[c] + 9], c_T64[1][0][0], g));
            out[4 * c + 3] = fma(px[OC[c] - 9], c_T64[1][1][18], fma(px[OC[c] + 9], c_T64[1][1][0], gx));
        }
    }
    // centre column: Gxx, Gxxx with the centre sample's own x variant (0: 17 taps, 1: 19 taps, 2: shifted)
    constexpr int XV = B ? 2 : A;
    double gxx = 0, gxxx = 0;
#pragma unroll
    for (int q = -9; q <= 9; ++q) { gxx = fma(px[OC[1] - q], c_T64[XV][2][q + 9], gxx); gxxx = fma(px[OC[1] - q], c_T64[XV][3][q + 9], gxxx); }
    out[12] = gxx; out[13] = gxxx;
}

__global__ void __launch_bounds__(128) toed_refine_kernel(DevBatch b, double magThresh, int border)
{
    __shared__ double s_T[3][4][19];
    __shared__ double s_rc[4][21][14];
    __shared__ double s_out[4][32][25];
    __shared__ uint32_t s_win[4][21][7];      // u8 window, 28-byte rows (7 words: odd stride, conflict-free row-parallel reads)
    for (int k = threadIdx.x; k < 3 * 4 * 19; k += 128) (&s_T[0][0][0])[k] = (&c_T64[0][0][0])[k];
    __syncthreads();
    const int img = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = b.nE[img];
    const uint8_t* src = b.und + (size_t)img * b.imgStride;
    int nrej = 0;
    for (int e0 = (blockIdx.x * 4 + w) * 32; e0 < n; e0 += gridDim.x * 4 * 32) {
        const int cnt = min(32, n - e0);
        const uint32_t ijl = lane < cnt ? b.coords[(size_t)img * b.E + e0 + lane] : 0u;
        // ---- 1-D sums of the 32 edges of this pass, one after the other ----
        for (int k = 0; k < cnt; ++k) {
            const uint32_t ij = __shfl_sync(0xffffffffu, ijl, k);
            const int gi = ij >> 16, gj = ij & 0xffff;
            const int a = gi & 1, bb = gj & 1;
            const int rbase = ((gi - 1) >> 1) - 9, cbase = ((gj - 1) >> 1) - 9;
            // the 21 x 21 u8 window (rows rbase.., columns cbase..) goes through shared memory: coalesced 32-bit loads of the
            // words that cover it (6 per row), zero outside the image (the reference's zero padding, cpu_toed.cpp:204-205)
            const int c4 = cbase & ~3;
            for (int t = lane; t < 21 * 6; t += 32) {
                const int r = t / 6, wd = t - 6 * r;
                const int row = rbase + r, col = c4 + 4 * wd;
                uint32_t v = 0;
                if (row >= 0 && row < b.H) {
                    const uint8_t* rp = src + (size_t)row * b.pitch;
                    if (col >= 0 && col + 3 < b.W) v = __ldg(reinterpret_cast<const uint32_t*>(rp + col));
                    else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) if (col + q >= 0 && col + q < b.W) v |= (uint32_t)__ldg(rp + col + q) << (8 * q);
                    }
                }
                s_win[w][r][wd] = v;
            }
            __syncwarp();
            if (lane < 21) {       // row pass: lane = image row
                const uint8_t* wp = reinterpret_cast<const uint8_t*>(&s_win[w][lane][0]) + (cbase - c4);
                double px[21];
#pragma unroll
                for (int c = 0; c < 21; ++c) px[c] = __hiloint2double(0x43300000, (int)wp[c]) - 4503599627370496.0;   // exact u8 -> double
                double* out = s_rc[w][lane];
                if (bb) { if (a) refine_row_pass<1, 1>(px, out); else refine_row_pass<1, 0>(px, out); }
                else { if (a) refine_row_pass<0, 1>(px, out); else refine_row_pass<0, 0>(px, out); }
            }
            __syncwarp();
            // column pass: lane l < 18 -> sample l / 2 (di = s / 3 - 1, dj = s % 3 - 1), fx (l even) or fy (l odd);
            // lanes 18..24 -> fxx, fyy, fxy, fxxy, fxyy, fxxx, fyyy at the centre sample
            if (lane < 25) {
                int di = 0, dj = 0, xk, yk;
                if (lane < 18) { const int sidx = lane >> 1; di = sidx / 3 - 1; dj = sidx % 3 - 1; xk = (lane & 1) ? 0 : 1; yk = (lane & 1) ? 1 : 0; }
                else {
                    const int t = lane - 18;             // x kernels 2 0 1 2 1 3 0, y kernels 0 2 1 1 2 0 3 (cpu_toed.cpp:207-216)
                    xk = (0x0312102 >> (4 * t)) & 15; yk = (0x3021120 >> (4 * t)) & 15;
                }
                const int ap = (gi + di) & 1, bp = (gj + dj) & 1, cid = (gi + di) >> 1;
                const int yv = ap ? 2 : (bp ? 1 : 0);
                const int slot = xk >= 2 ? 10 + xk : 4 * (dj + 1) + xk + ((bp == 0 && ap) ? 2 : 0);   // 19-tap version for sub-grid (1,0)
                const int r0 = cid - rbase;                                                           // row of tap p = 0
                EBVO_ASSERT(b.errFlag + (img >> 1), r0 - 9 >= 0 && r0 + 9 <= 20 && slot >= 0 && slot < 14 && cbase - (cbase & ~3) + 20 < 28);
                double acc = 0;
#pragma unroll
                for (int p = -9; p <= 9; ++p) acc = fma(s_rc[w][r0 - p][slot], s_T[yv][yk][p + 9], acc);
                s_out[w][k][lane] = acc;
            }
            __syncwarp();
        }
        // ---- closed forms: lane = edge ----
        bool ok = false;
        double ex = 0, ey = 0, eth = 0;
        if (lane < cnt) {
            const double* o = s_out[w][lane];
            const int gi = ijl >> 16, gj = ijl & 0xffff;
            auto FX = [&](int di, int dj) { return o[2 * ((di + 1) * 3 + dj + 1)]; };
            auto FY = [&](int di, int dj) { return o[2 * ((di + 1) * 3 + dj + 1) + 1]; };
            auto M = [&](int di, int dj) { const double x = FX(di, dj), y = FY(di, dj); return sqrt(x * x + y * y); };
            const double gx = FX(0, 0), gy = FY(0, 0);
            const double m = sqrt(gx * gx + gy * gy);
            if (m > magThresh && !(fabs(gx) < 10e-6 && fabs(gy) < 10e-6)) {         // cpu_toed.cpp:406-411
                const double nx = gx / m, ny = gy / m;
                double slope, fp, fm;
                // octant table, cpu_toed.cpp:418-477 (M(di, dj): row i + di, column j + dj)
                if (gx >= 0 && gy >= 0) {
                    if (gx >= gy) { slope = ny / nx; fp = M(0, 1) * (1 - slope) + M(1, 1) * slope; fm = M(0, -1) * (1 - slope) + M(-1, -1) * slope; }
                    else { slope = nx / ny; fp = M(1, 0) * (1 - slope) + M(1, 1) * slope; fm = M(-1, 0) * (1 - slope) + M(-1, -1) * slope; }
                } else if (gx < 0 && gy >= 0) {
                    if (fabs(gx) < gy) { slope = -nx / ny; fp = M(1, 0) * (1 - slope) + M(1, -1) * slope; fm = M(-1, 0) * (1 - slope) + M(-1, 1) * slope; }
                    else { slope = -ny / nx; fp = M(0, -1) * (1 - slope) + M(1, -1) * slope; fm = M(0, 1) * (1 - slope) + M(-1, 1) * slope; }
                } else if (gx < 0 && gy < 0) {
                    if (fabs(gx) >= fabs(gy)) { slope = ny / nx; fp = M(0, -1) * (1 - slope) + M(-1, -1) * slope; fm = M(0, 1) * (1 - slope) + M(1, 1) * slope; }
                    else { slope = nx / ny; fp = M(-1, 0) * (1 - slope) + M(-1, -1) * slope; fm = M(1, 0) * (1 - slope) + M(1, 1) * slope; }
                } else {
                    if (gx < fabs(gy)) { slope = -nx / ny; fp = M(-1, 0) * (1 - slope) + M(-1, 1) * slope; fm = M(1, 0) * (1 - slope) + M(1, -1) * slope; }
                    else { slope = -ny / nx; fp = M(0, 1) * (1 - slope) + M(-1, 1) * slope; fm = M(0, -1) * (1 - slope) + M(1, -1) * slope; }
                }
                if ((m > fm && m > fp) || (m > fm && m >= fp) || (m >= fm && m > fp)) {   // :481
                    const double s = sqrt(1 + slope * slope);
                    const double A = (fm + fp - 2 * m) / (2 * s * s), Bc = (fp - fm) / (2 * s);
                    const double ss = -Bc / (2 * A);
                    if (fabs(ss) <= sqrt(2.0)) {                                          // :496
                        const double X = (double)gj + ss * nx, Y = (double)gi + ss * ny;  // :499-500
                        ex = (X - 1) / 2; ey = (Y - 1) / 2;                               // :538, :542
                        ok = X != 0.0 && ex > (double)border && ex < (double)(b.W - border) && ey > (double)border && ey < (double)(b.H - border);   // :534, :553-554
                        // third-order orientation, cpu_toed.cpp:224-229
                        const double fx = gx, fy = gy, fxx = o[18], fyy = o[19], fxy = o[20], fxxy = o[21], fxyy = o[22], fxxx = o[23], fyyy = o[24];
                        double tx = fx * (2 * fxx * fxx + 2 * fxy * fxy) + fy * (2 * fxx * fxy + 2 * fyy * fxy) + 2 * fx * fy * fxxy + fy * fy * fxyy + fx * fx * fxxx;
                        double ty = fx * (2 * fxx * fxy + 2 * fyy * fxy) + fy * (2 * fyy * fyy + 2 * fxy * fxy) + 2 * fx * fy * fxyy + fx * fx * fxxy + fy * fy * fyyy;
                        const double tm = sqrt(tx * tx + ty * ty);
                        tx /= tm; ty /= tm;
                        eth = atan2(tx, -ty);
                    }
                }
            }
            const size_t oo = (size_t)img * b.E + e0 + lane;
            b.ex[oo] = ok ? ex : CUDART_NAN; b.ey[oo] = ey; b.eth[oo] = eth;     // NaN x marks a prefilter survivor the FP64 tests reject
            if (!ok) ++nrej;
        }
        __syncwarp();
    }
    nrej = __reduce_add_sync(0xffffffffu, nrej);
    if (lane == 0 && nrej) atomicAdd(&b.nRej[img], nrej);
}

// Stable in-place compaction of an image's edge list after the FP64 tests (one CTA per image; nothing to do - and nothing
// done - for the usual image without rejections).  Chunks are read whole before they are written and destinations never
// lie beyond their sources, so the compaction is safe in place and keeps the row-major order of cpu_toed.cpp:530-575.
__global__ void __launch_bounds__(1024) toed_prune_kernel(DevBatch b)
{
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (b.nRej[img] == 0) return;
    const int n = b.nE[img];
    double *ex = b.ex + (size_t)img * b.E, *ey = b.ey + (size_t)img * b.E, *eth = b.eth + (size_t)img * b.E;
    __shared__ int s_w[32];
    __shared__ int s_base;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += 1024) {
        const int i = c0 + tid;
        double x = 0, y = 0, t = 0;
        if (i < n) { x = ex[i]; y = ey[i]; t = eth[i]; }
        const bool keep = i < n && x == x;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_w[w] = __popc(m);
        __syncthreads();
        if (w == 0) {
            const int v = s_w[lane];
            int s = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, s, d); if (lane >= d) s += u; }
            s_w[lane] = s - v;
        }
        __syncthreads();
        const int base = s_base;
        const int pos = base + s_w[w] + __popc(m & ((1u << lane) - 1));
        if (keep) { ex[pos] = x; ey[pos] = y; eth[pos] = t; }
        __syncthreads();
        if (tid == 1023) s_base = base + s_w[31] + __popc(m);
        __syncthreads();
    }
    if (tid == 0) b.nE[img] = s_base;
}

// Tensor map over the context's image array: dims (W, H, images) of u8, strides (pitch, imgStride) bytes, box 64 x 52 x 1.
int make_toed_tensor_map(void* out128, const uint8_t* base, int W, int H, int pitch, size_t imgStride, int nImages)
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    static std::once_flag once;      // contexts may be created from several host threads (one per GPU)
    std::call_once(once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn) encode = (EncodeFn)fn;
    });
    if (!encode) return -1;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nImages};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)imgStride};
    const cuuint32_t box[3] = {(cuuint32_t)TMA_BOX_W, (cuuint32_t)IN_H, 1}, estr[3] = {1, 1, 1};
    const CUresult r = encode(reinterpret_cast<CUtensorMap*>(out128), CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

// per-DEVICE initialisation (constant tables, function attributes): called by ebvo_create after cudaSetDevice, so that
// every GPU a process opens a context on is set up (ebvo_stereo_batch_multi drives several devices from one process)
void init_toed_device()
{
    cudaFuncSetAttribute(toed_grad_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TOED_SMEM);
    cudaFuncSetAttribute(toed_grad_nms_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    upload_toed_tables();
}

void launch_toed(const DevBatch& b, const DevParams& p, int nImages, cudaStream_t st, Prof* prof)
{
    cudaMemsetAsync(b.rowcnt, 0, sizeof(int) * b.rowStride * nImages, st);
    cudaMemsetAsync(b.nTot, 0, sizeof(int) * nImages, st);
    cudaMemsetAsync(b.nRej, 0, sizeof(int) * nImages, st);
    dim3 gA(b.tilesX, b.tilesY, nImages);
    EBVO_KERNEL(prof, "toed_grad_nms", st, (toed_grad_nms_kernel<<<gA, TOED_THREADS, TOED_SMEM, st>>>(b, *reinterpret_cast<const CUtensorMap*>(b.tmap), p.toed_mag_thresh, p.toed_border)));
    EBVO_KERNEL(prof, "toed_scan", st, (toed_scan_kernel<<<nImages, 1024, 0, st>>>(b)));
    dim3 gE((b.H2 + 7) / 8, nImages);
    EBVO_KERNEL(prof, "toed_expand", st, (toed_expand_kernel<<<gE, 256, 0, st>>>(b)));
    int sms = 0, dev = 0;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    dim3 gC(std::max(1, std::min((b.E + 127) / 128, (sms * 12 + nImages - 1) / nImages)), nImages);
    EBVO_KERNEL(prof, "toed_refine", st, (toed_refine_kernel<<<gC, 128, 0, st>>>(b, (double)p.toed_mag_thresh, p.toed_border)));
    EBVO_KERNEL(prof, "toed_prune", st, (toed_prune_kernel<<<nImages, 1024, 0, st>>>(b)));
}

}  // namespace ebvo
