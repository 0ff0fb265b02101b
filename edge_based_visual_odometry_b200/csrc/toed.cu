// Third-order edge detection on sm_100a.
//
// Replaces ThirdOrderEdgeDetectionCPU::{preprocessing, convolve_img, non_maximum_suppresion}
// (reference src/toed/cpu_toed.cpp:82-120, 122-376, 386-582).  FP32 arithmetic; final coordinates are
// assembled in FP64 so that the edge list handed to the matcher keeps sub-1e-4 px resolution.
//
// Work-efficient split (same results as the reference's dense evaluation):
//   K_A  toed_grad_nms   dense, tile-fused: uint8 tile + halo -> separable G/Gx row pass (3 tap variants) ->
//                         column pass for fx, fy on the four 2x sub-grids -> gradient magnitude -> octant NMS +
//                         parabola sub-pixel fit (cpu_toed.cpp:400-514) -> per-row ballot bitmask (+ sparse
//                         sub-pixel offsets).  No interp-grid map is ever written to HBM.
//   K_B  toed_scan/expand row counts -> exclusive scan -> ordered (i,j) list.  Order = row-major interp order,
//                         exactly the serial scan of cpu_toed.cpp:530-575, so edge indices match the reference.
//   K_C  toed_orient      sparse: the seven remaining third-order responses + orientation (cpu_toed.cpp:224-229)
//                         are evaluated only at the surviving edge samples (about 2% of the interp grid).
#include "ebvo_internal.cuh"
#include <cmath>
#include <cuda.h>

namespace ebvo {

// [variant][filter][tap]; variant 0 = unshifted, middle 17 taps (ends zero); 1 = unshifted 19; 2 = shifted 19
// filter 0..3 = G, Gx, Gxx, Gxxx (closed forms quoted at cpu_toed.cpp:137-140,151-154; sigma = 2)
__constant__ float c_T[3][4][19];

void upload_toed_tables()
{
    float T[3][4][19];
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
        }
    cudaMemcpyToSymbol(c_T, T, sizeof(T));
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
    float2* sp = b.sp + (size_t)img * b.spStride;
    int* rowcnt = b.rowcnt + (size_t)img * b.rowStride;
    int nall = 0;
    for (int rr = 0; rr < 8; ++rr) {
        const int li = 2 + warp * 8 + rr;
        const int gi = 2 * y0 + (li - 2);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int lj = 2 + 32 * h + lane;
            const int gj = 2 * x0 + (lj - 2);
            bool edge = false, keep = false;
            float dx = 0.f, dy = 0.f;
            if (gi >= border && gi < b.H2 - border && gj >= border && gj < b.W2 - border) {
                const float* M = s_mag + li * IWP + lj;
                float m = M[0], gx = s_ix[li * IWP + lj], gy = s_iy[li * IWP + lj];
                if (m > magThresh && !(fabsf(gx) < 10e-6f && fabsf(gy) < 10e-6f)) {
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
                    if ((m > fm && m > fp) || (m > fm && m >= fp) || (m >= fm && m > fp)) {
                        float s = sqrtf(1.f + slope * slope);
                        float A = (fm + fp - 2.f * m) / (2.f * s * s), B = (fp - fm) / (2.f * s);
                        float ss = -B / (2.f * A);
                        if (fabsf(ss) <= 1.41421356237f) {
                            edge = true;
                            dx = ss * nx; dy = ss * ny;
                            double X = ((double)gj + (double)dx - 1.0) * 0.5, Y = ((double)gi + (double)dy - 1.0) * 0.5;
                            // border filter of cpu_toed.cpp:553-554; X==0 map sentinel (:534) cannot occur for gj >= 10
                            keep = (X > (double)border) && (X < (double)(b.W - border)) && (Y > (double)border) && (Y < (double)(b.H - border));
                        }
                    }
                }
            }
            unsigned all = __ballot_sync(0xffffffffu, edge), kp = __ballot_sync(0xffffffffu, keep);
            nall += __popc(all);
            if (keep) sp[(size_t)gi * b.W2 + gj] = make_float2(dx, dy);
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
        if (total > b.E) { atomicExch(b.errFlag, 1); total = b.E; }
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

// ---- K_C: third-order orientation at the edge samples; one thread per edge ------------------------------
__global__ void __launch_bounds__(128) toed_orient_kernel(DevBatch b)
{
    __shared__ float s_T[3][4][19];
    for (int k = threadIdx.x; k < 3 * 4 * 19; k += 128) (&s_T[0][0][0])[k] = (&c_T[0][0][0])[k];
    __syncthreads();
    const int img = blockIdx.y;
    const int n = b.nE[img];
    const int e = blockIdx.x * 128 + threadIdx.x;
    if (e >= n) return;
    const uint8_t* src = b.und + (size_t)img * b.imgStride;
    const uint32_t ij = b.coords[(size_t)img * b.E + e];
    const int gi = ij >> 16, gj = ij & 0xffff;
    const int a = gi & 1, bb = gj & 1, ci = gi >> 1, cj = gj >> 1;
    const int xv = bb ? 2 : (a ? 1 : 0), yv = a ? 2 : (bb ? 1 : 0);
    const float(*X)[19] = s_T[xv];
    const float(*Y)[19] = s_T[yv];
    float fx = 0, fy = 0, fxx = 0, fyy = 0, fxy = 0, fxxy = 0, fxyy = 0, fxxx = 0, fyyy = 0;
    for (int p = -9; p <= 9; ++p) {
        int row = ci - p;
        if (row < 0 || row >= b.H) continue;
        const uint8_t* rp = src + (size_t)row * b.pitch;
        float rG = 0, rGx = 0, rGxx = 0, rGxxx = 0;
#pragma unroll
        for (int q = -9; q <= 9; ++q) {
            int col = cj - q;
            float v = (col >= 0 && col < b.W) ? (float)rp[col] : 0.f;
            rG = fmaf(v, X[0][q + 9], rG);
            rGx = fmaf(v, X[1][q + 9], rGx);
            rGxx = fmaf(v, X[2][q + 9], rGxx);
            rGxxx = fmaf(v, X[3][q + 9], rGxxx);
        }
        float yG = Y[0][p + 9], yGx = Y[1][p + 9], yGxx = Y[2][p + 9], yGxxx = Y[3][p + 9];
        fx = fmaf(rGx, yG, fx);       fy = fmaf(rG, yGx, fy);
        fxx = fmaf(rGxx, yG, fxx);    fxy = fmaf(rGx, yGx, fxy);   fyy = fmaf(rG, yGxx, fyy);
        fxxy = fmaf(rGxx, yGx, fxxy); fxyy = fmaf(rGx, yGxx, fxyy);
        fxxx = fmaf(rGxxx, yG, fxxx); fyyy = fmaf(rG, yGxxx, fyyy);
    }
    // cpu_toed.cpp:224-229
    float tx = fx * (2 * fxx * fxx + 2 * fxy * fxy) + fy * (2 * fxx * fxy + 2 * fyy * fxy) + 2 * fx * fy * fxxy + fy * fy * fxyy + fx * fx * fxxx;
    float ty = fx * (2 * fxx * fxy + 2 * fyy * fxy) + fy * (2 * fyy * fyy + 2 * fxy * fxy) + 2 * fx * fy * fxyy + fx * fx * fxxy + fy * fy * fyyy;
    float tm = sqrtf(tx * tx + ty * ty);
    tx /= tm; ty /= tm;
    float th = atan2f(tx, -ty);
    float2 d = b.sp[(size_t)img * b.spStride + (size_t)gi * b.W2 + gj];
    size_t o = (size_t)img * b.E + e;
    b.ex[o] = ((double)gj + (double)d.x - 1.0) * 0.5;   // cpu_toed.cpp:538
    b.ey[o] = ((double)gi + (double)d.y - 1.0) * 0.5;   // cpu_toed.cpp:542
    b.eth[o] = (double)th;
}

// Tensor map over the context's image array: dims (W, H, images) of u8, strides (pitch, imgStride) bytes, box 64 x 52 x 1.
int make_toed_tensor_map(void* out128, const uint8_t* base, int W, int H, int pitch, size_t imgStride, int nImages)
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -1;
        encode = (EncodeFn)fn;
    }
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
    dim3 gA(b.tilesX, b.tilesY, nImages);
    EBVO_KERNEL(prof, "toed_grad_nms", st, (toed_grad_nms_kernel<<<gA, TOED_THREADS, TOED_SMEM, st>>>(b, *reinterpret_cast<const CUtensorMap*>(b.tmap), p.toed_mag_thresh, p.toed_border)));
    EBVO_KERNEL(prof, "toed_scan", st, (toed_scan_kernel<<<nImages, 1024, 0, st>>>(b)));
    dim3 gE((b.H2 + 7) / 8, nImages);
    EBVO_KERNEL(prof, "toed_expand", st, (toed_expand_kernel<<<gE, 256, 0, st>>>(b)));
    dim3 gC((b.E + 127) / 128, nImages);
    EBVO_KERNEL(prof, "toed_orient", st, (toed_orient_kernel<<<gC, 128, 0, st>>>(b)));
}

}  // namespace ebvo
