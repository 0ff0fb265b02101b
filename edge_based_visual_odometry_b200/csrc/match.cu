// Stereo edge correspondence on sm_100a: one warp per left edge (candidate set) or per candidate.
//
// Replaces Stereo_Matches::get_Stereo_Edge_Pairs + finalize_stereo_edge_mates, no-GT branch
// (reference src/Stereo_Matches.cpp:1360-1653) with the helpers it calls in src/utility.cpp,
// include/utility.h and src/EdgeClusterer.cpp.  Stage -> kernel map:
//   sobel_kernel           util_compute_Img_Gradients            include/utility.h:131-141
//   bounds_kernel(+scan)   index over the right edges (replaces the brute-force scan of :91-109)
//   gate_kernel            S1 epipolar distance, S2 disparity, S3 orientation   :381-419, 534-553, 863-915
//   sift_gate_kernel       S4 with injected descriptors           :691-787
//   patch_kernel           S5 oriented 7x7 patches of every edge, normalised once per edge   utility.cpp:82-212, 165-178
//   ncc_bnb_kernel         S6 NCC gate, S7 (and S7') best-nearly-best   :555-616, 789-862
//   shift_kernel           S8 epipolar shift (and the second shift of S10)   :26-89, 967-1037
//   gn*_kernel             S9 Gauss-Newton along the epipolar line   :1159-1358
//   cluster_kernel         S10 EdgeClusterer                      :1483, EdgeClusterer.cpp:119-302
//   ncc2_best_kernel       S11 NCC on the cluster centres, S12 arg-max   :1500, :916-965
//   compact_kernel         S13 remove_empty_clusters + finalize   :1543-1653
// Candidate lists live in a per-frame pool (CSR with explicit start/count per left edge); every list keeps
// the reference's order (ascending right-edge index, then the re-orderings the reference applies).
#include "ebvo_internal.cuh"
#include <math_constants.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <algorithm>

namespace ebvo {

constexpr int MAXC = 128;     // max candidates per left edge held in shared memory by the warp kernels
constexpr int WPB = 4;        // warps per block in the warp-per-item kernels
constexpr unsigned FULL = 0xffffffffu;
constexpr int GEO = 8;        // doubles per left edge in DevBatch::lines: a, b, c, dirx, diry, sin(thL), cos(thL), pad
constexpr int EB = 8;         // right edges per index block (bounds_kernel / gate_kernel): DevBatch::NB = E / EB

// 64-bit shuffles without the `asm volatile` register moves of the CUDA header's double overloads (those cost two MOVs per
// shuffle that ptxas may not remove: 6 % of the Gauss-Newton kernel's instructions); pure data movement, same values.
__device__ __forceinline__ double shfl_xor_d(double v, int m)
{
    return __hiloint2double(__shfl_xor_sync(FULL, __double2hiint(v), m), __shfl_xor_sync(FULL, __double2loint(v), m));
}
__device__ __forceinline__ double shfl_xor_dm(double v, int m, unsigned mask)   // for code that only part of the warp executes
{
    return __hiloint2double(__shfl_xor_sync(mask, __double2hiint(v), m), __shfl_xor_sync(mask, __double2loint(v), m));
}
__device__ __forceinline__ double shfl_idx_d(double v, int src)
{
    return __hiloint2double(__shfl_sync(FULL, __double2hiint(v), src), __shfl_sync(FULL, __double2loint(v), src));
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}
__device__ __forceinline__ void warp_sum2(double& a, double& b)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) { a += shfl_xor_d(a, o); b += shfl_xor_d(b, o); }
}
// Sums of FOUR values over the warp, every lane receiving all four: the butterfly is transposed - after the xor-16 step a
// lane carries two of the four partial sums, after the xor-8 step one - so 10 64-bit shuffles and 6 additions replace
// the 20 + 20 of two warp_sum2 calls.  The additions pair the same lanes in the same tree as the plain butterfly:
// the results are bit-identical to it.
__device__ __forceinline__ void warp_sum4(double& a, double& b, double& c, double& d, int lane)
{
    const bool h16 = lane & 16, h8 = lane & 8;
    double k0 = h16 ? c : a, k1 = h16 ? d : b;
    k0 += shfl_xor_d(h16 ? a : c, 16); k1 += shfl_xor_d(h16 ? b : d, 16);
    double k = h8 ? k1 : k0;
    k += shfl_xor_d(h8 ? k0 : k1, 8);
    k += shfl_xor_d(k, 4); k += shfl_xor_d(k, 2); k += shfl_xor_d(k, 1);
    a = shfl_idx_d(k, 0); b = shfl_idx_d(k, 8); c = shfl_idx_d(k, 16); d = shfl_idx_d(k, 24);
}
// Two values, every lane receiving both: transposed likewise (7 shuffles + 5 additions instead of 10 + 10), bit-identical.
__device__ __forceinline__ void warp_sum2t(double& a, double& b, int lane)
{
    const bool h16 = lane & 16;
    double k = h16 ? b : a;
    k += shfl_xor_d(h16 ? a : b, 16);
    k += shfl_xor_d(k, 8); k += shfl_xor_d(k, 4); k += shfl_xor_d(k, 2); k += shfl_xor_d(k, 1);
    a = shfl_idx_d(k, 0); b = shfl_idx_d(k, 16);
}
__device__ __forceinline__ void warp_sum3(double& a, double& b, double& c)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) { a += shfl_xor_d(a, o); b += shfl_xor_d(b, o); c += shfl_xor_d(c, o); }
}

// ------------------------------------------------------------------------------------------------------
// Sobel 3x3 * 1/8 with BORDER_REFLECT_101 on the undistorted right image (exact in FP32 on 8-bit data)
// ------------------------------------------------------------------------------------------------------
__global__ void sobel_kernel(DevBatch b)
{
    const int f = blockIdx.z;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= b.W || y >= b.H) return;
    const uint8_t* I = b.und + (size_t)(2 * f + 1) * b.imgStride;
    auto R = [](int i, int n) { return n == 1 ? 0 : (i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i)); };
    int xm = R(x - 1, b.W), xp = R(x + 1, b.W), ym = R(y - 1, b.H), yp = R(y + 1, b.H);
    auto at = [&](int yy, int xx) { return (float)I[(size_t)yy * b.pitch + xx]; };
    float gx = (at(ym, xp) - at(ym, xm)) * 0.125f + (at(y, xp) - at(y, xm)) * 0.25f + (at(yp, xp) - at(yp, xm)) * 0.125f;
    float gy = (at(yp, xm) - at(ym, xm)) * 0.125f + (at(yp, x) - at(ym, x)) * 0.25f + (at(yp, xp) - at(ym, xp)) * 0.125f;
    const size_t o = (size_t)f * b.gStride + (size_t)y * b.W + x;
    if (b.pk) b.pk[o] = make_float4(at(y, x), gx, gy, 0.f);
    if (b.pkh) {    // {half I, -, half gx, half gy}: 8-bit intensities and Sobel/8 values (k/8, |k| <= 1020) are exact in fp16; 8 bytes per pixel
        const __half2 hg = __floats2half2_rn(gx, gy);
        b.pkh[o] = make_uint2((uint32_t)__half_as_ushort(__float2half_rn(at(y, x))), *reinterpret_cast<const uint32_t*>(&hg));
    }
    if (b.pk16) {   // {I, 8*gx, 8*gy} as exact 16-bit integers (|8*g| <= 1020): 8 bytes per pixel
        const int i8 = (int)at(y, x), gx8 = (int)(gx * 8.f), gy8 = (int)(gy * 8.f);
        b.pk16[o] = make_uint2((uint32_t)i8 | ((uint32_t)(gx8 & 0xffff) << 16), (uint32_t)(gy8 & 0xffff));
    }
}

// ------------------------------------------------------------------------------------------------------
// Bounding intervals of EB-edge blocks of the right edge list + monotone envelopes for the search.
// The TOED list is in row-major interp order, so blocks are short runs along image rows; any order is still correct.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) bounds_kernel(DevBatch b)
{
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int img = 2 * f + 1;
    const int nR = b.nE[img];
    const int nblk = (nR + EB - 1) / EB;
    const double* ex = b.ex + (size_t)img * b.E;
    const double* ey = b.ey + (size_t)img * b.E;
    float4* blk = b.blk + (size_t)f * b.NB;
    float* pmax = b.pmax + (size_t)f * b.NB;
    float* smin = b.smin + (size_t)f * b.NB;
    // four blocks of EB = 8 consecutive edges per warp pass: the TOED list is in row-major interp order, so a block is a short
    // run along one image row (x span ~ 100 px); 32-edge blocks span a third of the row and made the x window nearly useless
    for (int k4 = warp * 4; k4 < nblk; k4 += 128) {
        const int k = k4 + (lane >> 3);
        const int e = k * EB + (lane & 7);
        float ylo = CUDART_INF_F, yhi = -CUDART_INF_F, xlo = CUDART_INF_F, xhi = -CUDART_INF_F;
        if (k < nblk && e < nR) {
            double x = ex[e], y = ey[e];
            ylo = __double2float_rd(y); yhi = __double2float_ru(y);
            xlo = __double2float_rd(x); xhi = __double2float_ru(x);
        }
#pragma unroll
        for (int o = 4; o; o >>= 1) {
            ylo = fminf(ylo, __shfl_xor_sync(FULL, ylo, o)); yhi = fmaxf(yhi, __shfl_xor_sync(FULL, yhi, o));
            xlo = fminf(xlo, __shfl_xor_sync(FULL, xlo, o)); xhi = fmaxf(xhi, __shfl_xor_sync(FULL, xhi, o));
        }
        if ((lane & 7) == 0 && k < nblk) blk[k] = make_float4(ylo, yhi, xlo, xhi);
    }
    __syncthreads();
    // prefix max of yhi, suffix min of ylo (serial per chunk + block scan; nblk <= NB)
    __shared__ float s_a[1024], s_b[1024];
    const int per = (nblk + 1023) / 1024;
    float mx = -CUDART_INF_F, mn = CUDART_INF_F;
    for (int k = 0; k < per; ++k) {
        int i = tid * per + k;
        if (i < nblk) mx = fmaxf(mx, blk[i].y);
        int j = nblk - 1 - (tid * per + k);
        if (j >= 0) mn = fminf(mn, blk[j].x);
    }
    s_a[tid] = mx; s_b[tid] = mn;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        float a = (tid >= d) ? s_a[tid - d] : -CUDART_INF_F, c = (tid >= d) ? s_b[tid - d] : CUDART_INF_F;
        __syncthreads();
        s_a[tid] = fmaxf(s_a[tid], a); s_b[tid] = fminf(s_b[tid], c);
        __syncthreads();
    }
    float runmx = tid ? s_a[tid - 1] : -CUDART_INF_F, runmn = tid ? s_b[tid - 1] : CUDART_INF_F;
    for (int k = 0; k < per; ++k) {
        int i = tid * per + k;
        if (i < nblk) { runmx = fmaxf(runmx, blk[i].y); pmax[i] = runmx; }
        int j = nblk - 1 - (tid * per + k);
        if (j >= 0) { runmn = fminf(runmn, blk[j].x); smin[j] = runmn; }
    }
    __syncthreads();
    // per-row lookup tables on the two monotone envelopes (binary search per integer y; the gate reads one entry each)
    int* yt = b.ytab + (size_t)f * 2 * b.YT;
    for (int k = tid; k < b.YT; k += 1024) {
        const float key = (float)k;
        int lo = 0, hi = nblk;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (pmax[mid] >= key) hi = mid; else lo = mid + 1; }
        yt[k] = lo;
        lo = 0; hi = nblk;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (smin[mid] > key) hi = mid; else lo = mid + 1; }
        yt[b.YT + k] = lo;
    }
}

// ------------------------------------------------------------------------------------------------------
// Gates.  mode 0: S1 only; 1: S1+S2; 2: S1+S2+S3 (production).  One warp per left edge.
// ------------------------------------------------------------------------------------------------------
struct Fmat { double v[9]; };     // F21, row-major: a kernel parameter (constant bank), not a device buffer shared between streams
static Fmat fmat_of(const double* F21)
{
    Fmat F;
    for (int k = 0; k < 9; ++k) F.v[k] = F21[k];
    return F;
}
struct GateCtx {
    double a, b, c, nrm, xL, yL, thL;
    double ylo, yhi, xlo, xhi;
    double qlo, qhi, m2lo, m2hi;   // guard bands of the division- and sqrt-free forms of the S1 / S2 tests
    int blo, bhi;
};

// The reference's tests are fabs(a x + b y + c) / nrm < epi (Stereo_Matches.cpp:99) and sqrt(dx^2 + dy^2) <= maxdisp (:545-546).
// t / nrm < epi is decided by t against epi * nrm and sqrt(s) <= maxdisp by s against maxdisp^2 whenever the value lies
// outside a 2^-50 relative band around the threshold (a correctly rounded quotient / root cannot cross it there); inside
// the band (probability ~1e-15 per test) the reference's own expression is evaluated.  Decisions are identical by construction.
__device__ __forceinline__ bool gate_test(const GateCtx& g, const DevParams& p, int mode, double xr, double yr, double thr)
{
    const double t = fabs(g.a * xr + g.b * yr + g.c);
    bool ok = t < g.qlo;
    if (!ok && !(t > g.qhi)) ok = (t / g.nrm) < p.epi;
    if (mode >= 1) {
        const double dx = g.xL - xr, dy = g.yL - yr;
        const double s2 = dx * dx + dy * dy;
        bool ok2 = s2 < g.m2lo;
        if (!ok2 && !(s2 > g.m2hi)) ok2 = sqrt(s2) <= p.maxdisp;
        ok = ok && ok2;
    }
    if (mode >= 2) {
        double od = fabs((g.thL - thr) * (180.0 / 3.14159265358979323846));   // :887-901
        if (od > 180.0) od = 360.0 - od;
        ok = ok && (od < p.orient_deg || fabs(od - 180.0) < p.orient_deg);
    }
    return ok;
}

__device__ __forceinline__ void gate_setup(GateCtx& g, const DevBatch& b, const DevParams& p, const Fmat& Fm, int f, int i, int mode)
{
    const int imgL = 2 * f;
    g.xL = b.ex[(size_t)imgL * b.E + i]; g.yL = b.ey[(size_t)imgL * b.E + i]; g.thL = b.eth[(size_t)imgL * b.E + i];
    const double* F = Fm.v;
    g.a = F[0] * g.xL + F[1] * g.yL + F[2] * 1.0;   // Stereo_Matches.cpp:15-16
    g.b = F[3] * g.xL + F[4] * g.yL + F[5] * 1.0;
    g.c = F[6] * g.xL + F[7] * g.yL + F[8] * 1.0;
    g.nrm = sqrt((g.a * g.a) + (g.b * g.b));
    {
        const double q = p.epi * g.nrm, m2 = p.maxdisp * p.maxdisp, eps = 8.8817841970012523e-16;   // 2^-50
        g.qlo = q * (1.0 - eps); g.qhi = q * (1.0 + eps); g.m2lo = m2 * (1.0 - eps); g.m2hi = m2 * (1.0 + eps);
    }
    // conservative search window
    if (mode >= 1) { g.xlo = g.xL - p.maxdisp - 1e-6; g.xhi = g.xL + p.maxdisp + 1e-6; g.ylo = g.yL - p.maxdisp - 1e-6; g.yhi = g.yL + p.maxdisp + 1e-6; }
    else { g.xlo = -1.0; g.xhi = (double)b.W + 1.0; g.ylo = -1.0; g.yhi = (double)b.H + 1.0; }
    if (fabs(g.b) > 1e-9 * g.nrm) {
        double y1 = -(g.a * g.xlo + g.c) / g.b, y2 = -(g.a * g.xhi + g.c) / g.b;
        double m = p.epi * g.nrm / fabs(g.b) + 1e-6;
        g.ylo = fmax(g.ylo, fmin(y1, y2) - m);
        g.yhi = fmin(g.yhi, fmax(y1, y2) + m);
    }
    // block range from the per-row tables bounds_kernel built (one load each instead of two multi-round searches on the
    // envelopes): ytab[0][k] = first block with pmax >= k, ytab[1][k] = first block with smin > k; floor / ceil keep it conservative
    const int* yt = b.ytab + (size_t)f * 2 * b.YT;
    const int klo = min(max((int)floorf((float)g.ylo - 1e-3f), 0), b.YT - 1), khi = min(max((int)ceilf((float)g.yhi + 1e-3f), 0), b.YT - 1);
    g.blo = yt[klo];
    g.bhi = yt[b.YT + khi];
}

// scan: calls emit(rank, ridx) in ascending ridx order for passing right edges; returns the count
template <typename Emit>
__device__ __forceinline__ int gate_scan(const GateCtx& g, const DevBatch& b, const DevParams& p, int f, int mode, int lane, Emit emit)
{
    const int imgR = 2 * f + 1;
    const int nR = b.nE[imgR];
    const double* ex = b.ex + (size_t)imgR * b.E;
    const double* ey = b.ey + (size_t)imgR * b.E;
    const double* eth = b.eth + (size_t)imgR * b.E;
    const float4* blk = b.blk + (size_t)f * b.NB;
    const float ylo = (float)g.ylo - 1e-3f, yhi = (float)g.yhi + 1e-3f, xlo = (float)g.xlo - 1e-3f, xhi = (float)g.xhi + 1e-3f;
    int count = 0;
    // the bounds of 32 blocks (256 right edges) are tested at once, then the surviving blocks are scanned four at a time: lane
    // l tests edge l & 7 of the (l >> 3)-th surviving block, so ascending lanes are ascending right-edge indices
    for (int k0 = g.blo; k0 < g.bhi; k0 += 32) {
        const int kk = k0 + lane;
        bool hit = false;
        if (kk < g.bhi) { const float4 bb = blk[kk]; hit = !(bb.y < ylo || bb.x > yhi || bb.w < xlo || bb.z > xhi); }
        unsigned bm = __ballot_sync(FULL, hit);
        while (bm) {
            const unsigned pos = __fns(bm, 0, (lane >> 3) + 1);             // position of this lane group's block, 0xffffffff if none
#pragma unroll
            for (int r = 0; r < 4; ++r) bm &= bm - 1;                       // (bm & (bm - 1) of 0 stays 0)
            const int e = (k0 + (int)pos) * EB + (lane & 7);
            const bool in = pos != 0xffffffffu && e < nR;
            double xa = 0, ya = 0, ta = 0;
            if (in) { xa = ex[e]; ya = ey[e]; ta = eth[e]; }
            const bool ok = in && gate_test(g, p, mode, xa, ya, ta);
            const unsigned m = __ballot_sync(FULL, ok);
            if (ok) emit(count + __popc(m & ((1u << lane) - 1)), e);
            count += __popc(m);
        }
    }
    return count;
}

__global__ void __launch_bounds__(32 * WPB, 8) gate_kernel(DevBatch b, DevParams p, const Fmat F)
{
    __shared__ int s_buf[WPB][64];
    const int f = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nL = b.nE[2 * f];
    int* cstart = b.cstart + (size_t)f * b.E;
    int* ccount = b.ccount + (size_t)f * b.E;
    int* c_ridx = b.c_ridx + (size_t)f * b.P;
    unsigned long long pairs = 0;
    for (int i = blockIdx.x * WPB + w; i < nL; i += gridDim.x * WPB) {
        GateCtx g;
        gate_setup(g, b, p, F, f, i, 2);
        if (lane == 0) {   // per-left-edge geometry used by every later stage
            double* l = b.lines + ((size_t)f * b.E + i) * GEO;
            double dirx = -g.b, diry = g.a;                         // Stereo_Matches.cpp:1330-1335
            const double nn = sqrt(dirx * dirx + diry * diry);
            dirx /= nn; diry /= nn;
            double sn, cs;
            sincos(g.thL, &sn, &cs);
            l[0] = g.a; l[1] = g.b; l[2] = g.c; l[3] = dirx; l[4] = diry; l[5] = sn; l[6] = cs; l[7] = 0.0;
        }
        int* buf = s_buf[w];
        int n = gate_scan(g, b, p, f, 2, lane, [&](int rank, int e) { EBVO_ASSERT(b.errFlag + f, e >= 0 && e < b.nE[2 * f + 1]); if (rank < 64) buf[rank] = e; });
        int start = 0;
        if (lane == 0 && n > 0) start = atomicAdd(&b.poolUsed[f], n);
        start = __shfl_sync(FULL, start, 0);
        if (n > 0 && start + n > b.P) {  // pool exhausted
            if (lane == 0) atomicExch(b.errFlag + f, 2);
            n = 0;
        }
        __syncwarp();
        if (n > 0) {
            if (n <= 64) { for (int k = lane; k < n; k += 32) c_ridx[start + k] = buf[k]; }
            else gate_scan(g, b, p, f, 2, lane, [&](int rank, int e) { EBVO_ASSERT(b.errFlag + f, start + rank < b.P); c_ridx[start + rank] = e; });
        }
        if (lane == 0) { cstart[i] = start; ccount[i] = n; }
        pairs += n;
        __syncwarp();
    }
    if (lane == 0 && pairs) atomicAdd(&b.counters[(size_t)f * 8 + 0], pairs);
}

// debug variants for the stage dumps (frame 0): count, then fill at host-scanned offsets
__global__ void __launch_bounds__(32 * WPB) gate_count_kernel(DevBatch b, DevParams p, const Fmat F, int mode, int* counts)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nL = b.nE[0];
    for (int i = blockIdx.x * WPB + w; i < nL; i += gridDim.x * WPB) {
        GateCtx g;
        gate_setup(g, b, p, F, 0, i, mode);
        int n = gate_scan(g, b, p, 0, mode, lane, [&](int, int) {});
        if (lane == 0) counts[i] = n;
    }
}
__global__ void __launch_bounds__(32 * WPB) gate_fill_kernel(DevBatch b, DevParams p, const Fmat F, int mode, const int* offsets, int* ridx)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nL = b.nE[0];
    for (int i = blockIdx.x * WPB + w; i < nL; i += gridDim.x * WPB) {
        GateCtx g;
        gate_setup(g, b, p, F, 0, i, mode);
        int o = offsets[i];
        gate_scan(g, b, p, 0, mode, lane, [&](int rank, int e) { ridx[o + rank] = e; });
    }
}

// ------------------------------------------------------------------------------------------------------
// Oriented 7x7 patches + NCC.  Lane l owns patch cells t = l and t = l + 32 (< 49) of BOTH the "+" and "-"
// patch, so the four cross dot products are lane-local before the shuffle reduction.
// ------------------------------------------------------------------------------------------------------
// include/utility.h:81-104 on an 8-bit image (convertTo CV_64F is exact): NaN outside the image or when a
// coordinate is an exact integer (0/0 weights in the reference formula).  FP64 blend, stored as float
// (utility.cpp:206-209), so that patch values are bit-identical to the reference's.
__device__ __forceinline__ float bilinear_u8(const uint8_t* __restrict__ I, int pitch, int W, int H, double px, double py)
{
    if (!(px == px) || !(py == py)) return CUDART_NAN_F;
    // floor and the integer cell without conversion instructions (the conversion unit - 16 lanes per SM and clock - is what bound
    // the patch kernels: floor, double -> int and four u8 -> double per sample): a round-down add of 1.5 * 2^52 leaves floor(p) in
    // the low word and, minus the constant, as a double (exact for |p| < 2^31; an infinite p fails the fx == px test below)
    const double MAGIC = 6755399441055744.0;
    const double tx = __dadd_rd(px, MAGIC), ty = __dadd_rd(py, MAGIC);
    const double fx = tx - MAGIC, fy = ty - MAGIC;
    const int ix = __double2loint(tx), iy = __double2loint(ty);
    if (ix < 0 || iy < 0 || fx == px || fy == py) return CUDART_NAN_F;
    if (ix + 1 >= W || iy + 1 >= H) return CUDART_NAN_F;
    const double wx2 = px - fx, wx1 = (fx + 1.0) - px, wy2 = py - fy, wy1 = (fy + 1.0) - py;   // denominators are exactly 1
    const uint8_t* r0 = I + (size_t)iy * pitch + ix;
    auto u8d = [](uint8_t v) { return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0; };   // exact u8 -> double
    const double v12 = u8d(__ldg(r0)), v22 = u8d(__ldg(r0 + 1)), v11 = u8d(__ldg(r0 + pitch)), v21 = u8d(__ldg(r0 + pitch + 1));
    const double f1 = wx1 * v11 + wx2 * v21;   // row ceil(y)
    const double f2 = wx1 * v12 + wx2 * v22;   // row floor(y)
    return (float)(wy2 * f1 + wy1 * f2);
}

struct Patches {      // normalised (zero-mean, unit-norm) cells owned by this lane
    float p[2], m[2];
    bool flatP, flatM; // sum of squares < 1e-10  => similarity -1 (utility.cpp:170-172)
};

__device__ __forceinline__ void raw_patches(const uint8_t* I, int pitch, int W, int H, double x, double y, double th, double shift,
                                            int lane, float (&vp)[2], float (&vm)[2])
{
    double s, c;
    sincos(th, &s, &c);
    // utility.cpp:84-87
    double pxp = x + shift * s, pyp = y + shift * (-c), pxm = x + shift * (-s), pym = y + shift * c;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        int t = lane + 32 * u;
        vp[u] = 0.f; vm[u] = 0.f;
        if (t < 49) {
            int i = t / 7 - 3, j = t % 7 - 3;
            double ox = c * (double)i - s * (double)j, oy = s * (double)i + c * (double)j;   // utility.cpp:151
            vp[u] = bilinear_u8(I, pitch, W, H, ox + pxp, oy + pyp);
            vm[u] = bilinear_u8(I, pitch, W, H, ox + pxm, oy + pym);
        }
    }
}

// utility.cpp:165-178 with OpenCV's CV_32F type mix (double mean/sums, float centred & normalised values)
__device__ __forceinline__ void normalise_patches(const float (&vp)[2], const float (&vm)[2], int lane, Patches& P)
{
    const bool has1 = (lane + 32) < 49;
    double sp = (double)vp[0] + (has1 ? (double)vp[1] : 0.0), sm = (double)vm[0] + (has1 ? (double)vm[1] : 0.0);
    warp_sum2t(sp, sm, lane);
    float mp = (float)(sp / 49.0), mm = (float)(sm / 49.0);
    float dp0 = vp[0] - mp, dp1 = has1 ? vp[1] - mp : 0.f, dm0 = vm[0] - mm, dm1 = has1 ? vm[1] - mm : 0.f;
    double ssp = (double)(dp0 * dp0) + (double)(dp1 * dp1), ssm = (double)(dm0 * dm0) + (double)(dm1 * dm1);
    warp_sum2t(ssp, ssm, lane);
    P.flatP = ssp < 1e-10; P.flatM = ssm < 1e-10;
    float ip = (float)(1.0 / sqrt(ssp)), im = (float)(1.0 / sqrt(ssm));
    P.p[0] = dp0 * ip; P.p[1] = dp1 * ip; P.m[0] = dm0 * im; P.m[1] = dm1 * im;
}

// max of the four similarities with std::max({..}) NaN semantics (Stereo_Matches.cpp:592-596)
__device__ __forceinline__ double ncc_score(const Patches& A, const Patches& B)
{
    double pp = (double)A.p[0] * (double)B.p[0] + (double)A.p[1] * (double)B.p[1];
    double nn = (double)A.m[0] * (double)B.m[0] + (double)A.m[1] * (double)B.m[1];
    double pn = (double)A.p[0] * (double)B.m[0] + (double)A.p[1] * (double)B.m[1];
    double np = (double)A.m[0] * (double)B.p[0] + (double)A.m[1] * (double)B.p[1];
    warp_sum4(pp, nn, pn, np, threadIdx.x & 31);
    if (A.flatP || B.flatP) pp = -1.0;
    if (A.flatM || B.flatM) nn = -1.0;
    if (A.flatP || B.flatM) pn = -1.0;
    if (A.flatM || B.flatP) np = -1.0;
    double m = pp;
    if (m < nn) m = nn;
    if (m < pn) m = pn;
    if (m < np) m = np;
    return m;
}

// Best-nearly-best on a warp-private score list (Stereo_Matches.cpp:789-862).  On return order[k] (k < keep)
// lists the surviving positions in the order the reference leaves them; returns keep.
__device__ __forceinline__ int bnb_select(const double* sc, int n, double thr, bool is_ncc, int lane, int* order, bool always_sorted = false)
{
    if (n < 2) { if (lane == 0 && n == 1) order[0] = 0; __syncwarp(); return n; }
    // best = max (NCC) or min (SIFT distance)
    double best = is_ncc ? -CUDART_INF : CUDART_INF;
    for (int k = lane; k < n; k += 32) best = is_ncc ? fmax(best, sc[k]) : fmin(best, sc[k]);
#pragma unroll
    for (int o = 16; o; o >>= 1) { double t = shfl_xor_d(best, o); best = is_ncc ? fmax(best, t) : fmin(best, t); }
    int keep;
    if (best == 0.0) keep = 1;
    else {
        int c = 0;
        for (int k = lane; k < n; k += 32) { double r = is_ncc ? sc[k] / best : best / sc[k]; c += (r >= thr) ? 1 : 0; }
#pragma unroll
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
        keep = c < 1 ? 1 : c;
    }
    // the stereo stage leaves a list it keeps whole in its original order; the quad stage always returns sorted order
    // (Temporal_Matches.cpp:556-561 rebuilds the list from the sorted indices)
    if (keep >= n && !always_sorted) { for (int k = lane; k < n; k += 32) order[k] = k; __syncwarp(); return n; }
    if (keep > n) keep = n;
    // stable rank in sorted order
    for (int k = lane; k < n; k += 32) {
        double s = sc[k];
        int r = 0;
        for (int q = 0; q < n; ++q) {
            double t = sc[q];
            bool before = is_ncc ? (t > s) : (t < s);
            r += (before || (t == s && q < k)) ? 1 : 0;
        }
        if (r < keep) order[r] = k;
    }
    __syncwarp();
    return keep;
}

__device__ __forceinline__ void dump_put(const DumpBuf& d, int o, int ridx, double x, double y, double th, double score)
{
    d.ridx[o] = ridx; d.x[o] = x; d.y[o] = y; d.th[o] = th; d.score[o] = score;
}

// ------------------------------------------------------------------------------------------------------
// Quarter-warp (8-lane) helpers of the patch / NCC kernels: lane q of a group owns cells t = q + 8 r, r < 7 (t < 49) of
// both the "+" and the "-" patch; four items (edges, candidate pairs, cluster centres) advance per warp instruction and
// a reduction is 3 butterfly steps.  A full warp per item left a third of the lanes idle in the second round of cells
// and paid 5-step reductions per item.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void group8_sum2(double& a, double& b, int q, int base)   // sums over the 8 lanes; both results in every lane
{
    const bool h = q & 4;
    double k = h ? b : a;
    k += shfl_xor_d(h ? a : b, 4);
    k += shfl_xor_d(k, 2); k += shfl_xor_d(k, 1);
    a = shfl_idx_d(k, base); b = shfl_idx_d(k, base + 4);
}
__device__ __forceinline__ void group8_sum4(double& a, double& b, double& c, double& d, int q, int base)
{
    const bool h4 = q & 4, h2 = q & 2;
    double k0 = h4 ? c : a, k1 = h4 ? d : b;
    k0 += shfl_xor_d(h4 ? a : c, 4); k1 += shfl_xor_d(h4 ? b : d, 4);
    double k = h2 ? k1 : k0;
    k += shfl_xor_d(h2 ? k0 : k1, 2);
    k += shfl_xor_d(k, 1);
    a = shfl_idx_d(k, base); b = shfl_idx_d(k, base + 2); c = shfl_idx_d(k, base + 4); d = shfl_idx_d(k, base + 6);
}

// raw "+"/"-" samples of the cells this lane owns (utility.cpp:82-93, 141-159); act = false leaves zeros
__device__ __forceinline__ void raw_patches8(const uint8_t* I, int pitch, int W, int H, double x, double y, double s, double c, double shift,
                                             int q, bool act, float (&vp)[7], float (&vm)[7])
{
    const double pxp = x + shift * s, pyp = y + shift * (-c), pxm = x + shift * (-s), pym = y + shift * c;   // utility.cpp:84-87
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        const int t = q + 8 * r;
        vp[r] = 0.f; vm[r] = 0.f;
        if (act && t < 49) {
            const int i = t / 7 - 3, j = t % 7 - 3;
            const double ox = c * (double)i - s * (double)j, oy = s * (double)i + c * (double)j;   // utility.cpp:151
            vp[r] = bilinear_u8(I, pitch, W, H, ox + pxp, oy + pyp);
            vm[r] = bilinear_u8(I, pitch, W, H, ox + pxm, oy + pym);
        }
    }
}
// utility.cpp:165-178 with OpenCV's CV_32F type mix (double mean / sums, float centred and normalised values); in place
__device__ __forceinline__ void normalise_patches8(float (&vp)[7], float (&vm)[7], int q, int base, bool& flatP, bool& flatM)
{
    const int nr = q == 0 ? 7 : 6;
    double sp = 0, sm = 0;
#pragma unroll
    for (int r = 0; r < 7; ++r) if (r < nr) { sp += (double)vp[r]; sm += (double)vm[r]; }
    group8_sum2(sp, sm, q, base);
    const float mp = (float)(sp / 49.0), mm = (float)(sm / 49.0);
    double ssp = 0, ssm = 0;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        if (r < nr) { vp[r] -= mp; vm[r] -= mm; ssp += (double)(vp[r] * vp[r]); ssm += (double)(vm[r] * vm[r]); }
    }
    group8_sum2(ssp, ssm, q, base);
    flatP = ssp < 1e-10; flatM = ssm < 1e-10;          // => similarity -1 (utility.cpp:170-172)
    const float ip = (float)(1.0 / sqrt(ssp)), im = (float)(1.0 / sqrt(ssm));
#pragma unroll
    for (int r = 0; r < 7; ++r) if (r < nr) { vp[r] *= ip; vm[r] *= im; }
}
// max of the four similarities with std::max({..}) NaN semantics (Stereo_Matches.cpp:592-596); sums already reduced
__device__ __forceinline__ double ncc_max4(double pp, double nn, double pn, double np, bool aP, bool aM, bool bP, bool bM)
{
    if (aP || bP) pp = -1.0;
    if (aM || bM) nn = -1.0;
    if (aP || bM) pn = -1.0;
    if (aM || bP) np = -1.0;
    double m = pp;
    if (m < nn) m = nn;
    if (m < pn) m = pn;
    if (m < np) m = np;
    return m;
}

// ------------------------------------------------------------------------------------------------------
// S5: the "+"/"-" patches of EVERY edge of both views, sampled from the RAW image (Stereo_Matches.cpp:562-563)
// and normalised once.  A right edge is a candidate of ~5 left edges; the reference re-samples it for each pair.
// Layout: npatch[img][e][0..48] = "+" cells, [52..100] = "-" cells (zero-mean, unit-norm floats, zero padding to
// 16-byte rows); pflag bit0/bit1 = flat "+"/"-" patch.  A warp takes 32 edges: every lane evaluates sincos for one
// of them (one pass of the FP64 trigonometry per 32 edges), then four edges at a time are sampled by a quarter-warp each.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * WPB) patch_kernel(DevBatch b, DevParams p, int img0)
{
    const int img = blockIdx.y + img0, lane = threadIdx.x & 31;
    const int g = lane >> 3, q = lane & 7, base = lane & ~7;
    const int n = b.nE[img];
    const uint8_t* I = b.raw + (size_t)img * b.imgStride;
    const double *ex = b.ex + (size_t)img * b.E, *ey = b.ey + (size_t)img * b.E, *eth = b.eth + (size_t)img * b.E;
    float* np_ = b.npatch + (size_t)img * b.E * NPF;
    uint8_t* pf = b.pflag + (size_t)img * b.E;
    for (int e0 = (blockIdx.x * WPB + (threadIdx.x >> 5)) * 32; e0 < n; e0 += gridDim.x * WPB * 32) {
        double xl = 0, yl = 0, sl = 0, cl = 0;
        if (e0 + lane < n) { xl = ex[e0 + lane]; yl = ey[e0 + lane]; sincos(eth[e0 + lane], &sl, &cl); }
        for (int k = 0; k < 8 && e0 + 4 * k < n; ++k) {
            const int src = 4 * k + g, e = e0 + src;
            const double x = shfl_idx_d(xl, src), y = shfl_idx_d(yl, src), s = shfl_idx_d(sl, src), c = shfl_idx_d(cl, src);
            const bool act = e < n;
            float vp[7], vm[7];
            bool flatP, flatM;
            raw_patches8(I, b.pitch, b.W, b.H, x, y, s, c, p.shift_mag, q, act, vp, vm);
            normalise_patches8(vp, vm, q, base, flatP, flatM);
            if (act) {
                float* o = np_ + (size_t)e * NPF;
#pragma unroll
                for (int r = 0; r < 6; ++r) { o[q + 8 * r] = vp[r]; o[52 + q + 8 * r] = vm[r]; }
                if (q == 0) { o[48] = vp[6]; o[100] = vm[6]; pf[e] = (flatP ? 1 : 0) | (flatM ? 2 : 0); }
                else if (q < 4) { o[48 + q] = 0.f; o[100 + q] = 0.f; }
            }
        }
    }
}

__device__ __forceinline__ void load_patches(const float* __restrict__ np_, const uint8_t* __restrict__ pf, int e, int lane, Patches& P)
{
    const float* o = np_ + (size_t)e * 98;
    const bool has1 = lane + 32 < 49;
    P.p[0] = o[lane]; P.m[0] = o[49 + lane];
    P.p[1] = has1 ? o[lane + 32] : 0.f; P.m[1] = has1 ? o[49 + lane + 32] : 0.f;
    const int fl = pf[e];
    P.flatP = fl & 1; P.flatM = fl & 2;
}

// S6 + S7 (+ S7'): one warp per left edge, four candidates at a time (a quarter-warp per pair).  Lane q owns the float4
// chunks q and q + 8 (< 13) of each 52-float half of a patch row, so the four cross dot products are lane-local FP64
// sums of exact products before a 3-step reduction; the right patches of the next four candidates are in flight while
// the current four are scored.
struct PatchRegs { float4 p0, p1, m0, m1; int flags; };
__device__ __forceinline__ void load_patch_row(const float* __restrict__ np_, const uint8_t* __restrict__ pf, int e, int q, PatchRegs& P)
{
    const float4* o = reinterpret_cast<const float4*>(np_ + (size_t)e * NPF);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    P.p0 = __ldg(o + q); P.m0 = __ldg(o + 13 + q);
    P.p1 = q < 5 ? __ldg(o + 8 + q) : z; P.m1 = q < 5 ? __ldg(o + 21 + q) : z;
    P.flags = pf[e];
}
__global__ void __launch_bounds__(32 * WPB, 6) ncc_bnb_kernel(DevBatch b, DevParams p, int use_sift, int f0)
{
    __shared__ double s_sc[WPB][MAXC];
    __shared__ double s_cf[WPB][MAXC];
    __shared__ double s_tmp[WPB][MAXC];
    __shared__ int s_ri[WPB][MAXC];
    __shared__ int s_or[WPB][MAXC];
    __shared__ int s_or2[WPB][MAXC];
    __shared__ int s_or3[WPB][MAXC];
    const int f = blockIdx.y + f0, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane >> 3, q = lane & 7, base = lane & ~7;
    const int imgL = 2 * f, imgR = 2 * f + 1;
    const int nL = b.nE[imgL];
    const float *npL = b.npatch + (size_t)imgL * b.E * NPF, *npR = b.npatch + (size_t)imgR * b.E * NPF;
    const uint8_t *pfL = b.pflag + (size_t)imgL * b.E, *pfR = b.pflag + (size_t)imgR * b.E;
    const double *exR = b.ex + (size_t)imgR * b.E, *eyR = b.ey + (size_t)imgR * b.E, *ethR = b.eth + (size_t)imgR * b.E;
    const int* cstart = b.cstart + (size_t)f * b.E;
    int* ccount = b.ccount + (size_t)f * b.E;
    int* c_ridx = b.c_ridx + (size_t)f * b.P;
    double *c_x = b.c_x + (size_t)f * b.P, *c_y = b.c_y + (size_t)f * b.P, *c_th = b.c_th + (size_t)f * b.P;
    double *c_score = b.c_score + (size_t)f * b.P, *c_conf = b.c_conf + (size_t)f * b.P;
    int* c_owner = b.c_owner + (size_t)f * b.P;
    const bool dumps = b.dumps && f == 0;
    unsigned long long kept = 0;
    for (int i = blockIdx.x * WPB + w; i < nL; i += gridDim.x * WPB) {
        const int n = ccount[i];
        if (n == 0) { if (dumps && lane == 0) { b.dump[DUMP_S6].n[i] = 0; b.dump[DUMP_S7].n[i] = 0; } continue; }
        const int st = cstart[i];
        PatchRegs L, R, N;
        load_patch_row(npL, pfL, i, q, L);
        const double lp[8] = {L.p0.x, L.p0.y, L.p0.z, L.p0.w, L.p1.x, L.p1.y, L.p1.z, L.p1.w};
        const double lm[8] = {L.m0.x, L.m0.y, L.m0.z, L.m0.w, L.m1.x, L.m1.y, L.m1.z, L.m1.w};
        const bool lP = L.flags & 1, lM = L.flags & 2;
        int ns = 0;
        int rn = g < n ? c_ridx[st + g] : -1;
        EBVO_ASSERT(b.errFlag + f, st >= 0 && st + n <= b.P && rn < b.nE[imgR]);
        N = L;
        if (rn >= 0) load_patch_row(npR, pfR, rn, q, N);
        for (int j0 = 0; j0 < n; j0 += 4) {
            const int r = rn;
            R = N;
            if (j0 + 4 < n) {      // prefetch: the loads overlap the sums and reductions below
                rn = j0 + 4 + g < n ? c_ridx[st + j0 + 4 + g] : -1;
                if (rn >= 0) load_patch_row(npR, pfR, rn, q, N);
            }
            const double rp[8] = {R.p0.x, R.p0.y, R.p0.z, R.p0.w, R.p1.x, R.p1.y, R.p1.z, R.p1.w};
            const double rm[8] = {R.m0.x, R.m0.y, R.m0.z, R.m0.w, R.m1.x, R.m1.y, R.m1.z, R.m1.w};
            double pp = 0, nn = 0, pn = 0, np = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) { pp = fma(lp[u], rp[u], pp); nn = fma(lm[u], rm[u], nn); pn = fma(lp[u], rm[u], pn); np = fma(lm[u], rp[u], np); }
            group8_sum4(pp, nn, pn, np, q, base);
            const double sc = ncc_max4(pp, nn, pn, np, lP, lM, R.flags & 1, R.flags & 2);
            const bool pass = r >= 0 && sc > p.ncc_thresh;       // NCC_THRESH gate, :597
            const unsigned m = __ballot_sync(FULL, pass && q == 0);
            if (pass && q == 0) {
                const int o = ns + __popc(m & ((1u << lane) - 1));
                if (o < MAXC) { s_sc[w][o] = sc; s_ri[w][o] = r; s_cf[w][o] = use_sift ? c_conf[st + j0 + g] : 0.0; }
            }
            ns += __popc(m);
        }
        if (ns > MAXC) { if (lane == 0) atomicExch(b.errFlag + f, 3); ns = MAXC; }
        __syncwarp();
        if (dumps) {
            for (int k = lane; k < ns; k += 32) { int r = s_ri[w][k]; dump_put(b.dump[DUMP_S6], st + k, r, exR[r], eyR[r], ethR[r], s_sc[w][k]); }
            if (lane == 0) b.dump[DUMP_S6].n[i] = ns;
        }
        // S7: best-nearly-best on NCC
        int keep = bnb_select(s_sc[w], ns, p.bnb_ncc, true, lane, s_or[w]);
        if (dumps) {
            for (int k = lane; k < keep; k += 32) { int o = s_or[w][k], r = s_ri[w][o]; dump_put(b.dump[DUMP_S7], st + k, r, exR[r], eyR[r], ethR[r], s_sc[w][o]); }
            if (lane == 0) b.dump[DUMP_S7].n[i] = keep;
        }
        if (use_sift && keep >= 2) {
            // S7': order the survivors by SIFT distance; compose the two selections
            for (int k = lane; k < keep; k += 32) s_tmp[w][k] = s_cf[w][s_or[w][k]];
            __syncwarp();
            const int keep2 = bnb_select(s_tmp[w], keep, p.bnb_sift, false, lane, s_or2[w]);
            for (int k = lane; k < keep2; k += 32) s_or3[w][k] = s_or[w][s_or2[w][k]];
            __syncwarp();
            for (int k = lane; k < keep2; k += 32) s_or[w][k] = s_or3[w][k];
            keep = keep2;
            __syncwarp();
        }
        for (int k = lane; k < keep; k += 32) {
            const int o = s_or[w][k], r = s_ri[w][o];
            c_ridx[st + k] = r;
            c_x[st + k] = exR[r]; c_y[st + k] = eyR[r]; c_th[st + k] = ethR[r];
            c_score[st + k] = s_sc[w][o];
            c_conf[st + k] = s_cf[w][o];
            c_owner[st + k] = i;
        }
        for (int k = keep + lane; k < n; k += 32) c_owner[st + k] = -1;   // dead slots of this segment
        if (lane == 0) ccount[i] = keep;
        kept += keep;
        __syncwarp();
    }
    if (lane == 0 && kept) atomicAdd(&b.counters[(size_t)f * 8 + 1], kept);
}

// ------------------------------------------------------------------------------------------------------
// S4 (optional): SIFT gate with caller-supplied descriptors (frame 0 only).  One warp per left edge.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * WPB) sift_gate_kernel(DevBatch b, DevParams p)
{
    const int f = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nL = b.nE[2 * f];
    const int* cstart = b.cstart + (size_t)f * b.E;
    int* ccount = b.ccount + (size_t)f * b.E;
    int* c_ridx = b.c_ridx + (size_t)f * b.P;
    double* c_conf = b.c_conf + (size_t)f * b.P;
    for (int i = blockIdx.x * WPB + w; i < nL; i += gridDim.x * WPB) {
        const int n = ccount[i];
        if (n == 0) continue;
        const int st = cstart[i];
        const float* l1 = b.descL + (size_t)i * 256;
        const float4 a1 = reinterpret_cast<const float4*>(l1)[lane], a2 = reinterpret_cast<const float4*>(l1 + 128)[lane];
        int ns = 0;
        for (int j = 0; j < n; ++j) {
            const int r = c_ridx[st + j];
            const float* r1 = b.descR + (size_t)r * 256;
            const float4 b1 = reinterpret_cast<const float4*>(r1)[lane], b2 = reinterpret_cast<const float4*>(r1 + 128)[lane];
            auto d2 = [](float4 u, float4 v) {
                double x = (double)u.x - (double)v.x, y = (double)u.y - (double)v.y, z = (double)u.z - (double)v.z, q = (double)u.w - (double)v.w;
                return x * x + y * y + z * z + q * q;
            };
            double d11 = d2(a1, b1), d21 = d2(a2, b1), d12 = d2(a1, b2), d22 = d2(a2, b2);
            warp_sum4(d11, d21, d12, d22, lane);
            const double d = fmin(fmin(sqrt(d11), sqrt(d21)), fmin(sqrt(d12), sqrt(d22)));   // :736-740
            __syncwarp();
            if (d < p.sift_thresh) {
                if (lane == 0) { c_ridx[st + ns] = r; c_conf[st + ns] = d; }
                ++ns;
            }
            __syncwarp();
        }
        // the slots the gate dropped stay allocated in the pool: mark them dead, or the slot-driven kernels (shift, Gauss-Newton) would
        // work on whatever owner index an earlier call left there
        for (int k = ns + lane; k < n; k += 32) b.c_owner[(size_t)f * b.P + st + k] = -1;
        if (lane == 0) ccount[i] = ns;
    }
}

// The same gate on descriptors computed on the device (sift.cu; every frame): 8-bit entries, so the four squared
// distances are exact integer sums (cv::norm accumulates the float entries in double: the same value), and the smallest
// of the four distances is the root of the smallest squared one.  One warp per left edge, one quarter-warp per candidate
// (four candidates per step): a lane holds 16 bytes of each of the two left and the two right descriptors, byte
// differences and their squares are SIMD-in-a-word instructions (vabsdiffu4, dp4a).
__global__ void __launch_bounds__(32 * WPB) sift_gate8_kernel(DevBatch b, DevParams p)
{
    const int f = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane >> 3, gl = lane & 7;
    const unsigned gmask = 0xffu << (8 * g);
    const int nL = b.nE[2 * f];
    const int* cstart = b.cstart + (size_t)f * b.E;
    int* ccount = b.ccount + (size_t)f * b.E;
    int* c_ridx = b.c_ridx + (size_t)f * b.P;
    double* c_conf = b.c_conf + (size_t)f * b.P;
    const uint8_t* dL = b.desc8 + (size_t)(2 * f) * b.E * 256;
    const uint8_t* dR = b.desc8 + (size_t)(2 * f + 1) * b.E * 256;
    auto ssd = [](const uint4& u, const uint4& v) {
        unsigned d, acc;
        d = __vabsdiffu4(u.x, v.x); acc = __dp4a(d, d, 0u);
        d = __vabsdiffu4(u.y, v.y); acc = __dp4a(d, d, acc);
        d = __vabsdiffu4(u.z, v.z); acc = __dp4a(d, d, acc);
        d = __vabsdiffu4(u.w, v.w); return __dp4a(d, d, acc);
    };
    for (int i = blockIdx.x * WPB + w; i < nL; i += gridDim.x * WPB) {
        const int n = ccount[i];
        if (n == 0) continue;
        const int st = cstart[i];
        const uint4 a1 = reinterpret_cast<const uint4*>(dL + (size_t)i * 256)[gl], a2 = reinterpret_cast<const uint4*>(dL + (size_t)i * 256 + 128)[gl];
        int ns = 0;
        for (int j0 = 0; j0 < n; j0 += 4) {
            const int j = j0 + g;
            const bool live = j < n;
            const int r = live ? c_ridx[st + j] : 0;
            unsigned dmin = 0xffffffffu;
            if (live) {
                const uint4 b1 = reinterpret_cast<const uint4*>(dR + (size_t)r * 256)[gl], b2 = reinterpret_cast<const uint4*>(dR + (size_t)r * 256 + 128)[gl];
                const unsigned d11 = __reduce_add_sync(gmask, ssd(a1, b1)), d21 = __reduce_add_sync(gmask, ssd(a2, b1));
                const unsigned d12 = __reduce_add_sync(gmask, ssd(a1, b2)), d22 = __reduce_add_sync(gmask, ssd(a2, b2));
                dmin = min(min(d11, d21), min(d12, d22));                         // :736-740
            }
            const double d = sqrt((double)dmin);
            const bool pass = live && d < p.sift_thresh;
            const unsigned pm = __ballot_sync(FULL, pass && gl == 0);             // bit 8 g = candidate j0 + g passes
            __syncwarp();                                                         // every candidate of this step is read before a slot is rewritten
            if (pass && gl == 0) {
                const int pos = st + ns + __popc(pm & ((1u << lane) - 1u));
                c_ridx[pos] = r; c_conf[pos] = d;
            }
            ns += __popc(pm);
            __syncwarp();
        }
        // the slots the gate dropped stay allocated in the pool: mark them dead, or the slot-driven kernels (shift, Gauss-Newton) would
        // work on whatever owner index an earlier call left there
        for (int k = ns + lane; k < n; k += 32) b.c_owner[(size_t)f * b.P + st + k] = -1;
        if (lane == 0) ccount[i] = ns;
    }
}

// ------------------------------------------------------------------------------------------------------
// S8 epipolar shift (Stereo_Matches.cpp:26-89; utility.cpp:46-74).  One THREAD per live pool slot: scalar FP64
// trigonometry is 32x cheaper here than replicated across the lanes of a warp-per-candidate kernel.
// which = 0: first shift (S8, positions + orientation written back, right-edge index dropped as at :993-997);
// which = 1: second shift at the head of S10 (Stereo_Matches.cpp:1483 -> :981-998).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double tangential(double a, double b, double c, double x, double y, double th, double& xi, double& yi)
{
    const double ae = tan(th), be = -1.0, ce = -(ae * x - y);
    xi = (b * ce - be * c) / (a * be - ae * b);
    yi = (c * ae - ce * a) / (a * be - ae * b);
    return sqrt((xi - x) * (xi - x) + (yi - y) * (yi - y));
}
__device__ __forceinline__ void shift_to_line(double a, double b, double c, const DevParams& p, double& x, double& y, double& th)
{
    const double ex = x - a * (a * x + b * y + c) / (a * a + b * b);
    const double ey = y - b * (a * x + b * y + c) / (a * a + b * b);
    const double dn = sqrt((x - ex) * (x - ex) + (y - ey) * (y - ey));
    if (dn < p.loc_pert) { x = ex; y = ey; return; }
    double xi, yi;
    if (tangential(a, b, c, x, y, th, xi, yi) < p.tang_displ) { x = xi; y = yi; return; }
    double s, co;
    sincos(th, &s, &co);
    const double pt = a * co + b * s, dpt = -a * s + b * co;
    double t2 = th;
    if (pt > 0 && dpt < 0) t2 -= p.orient_pert;
    else if (pt < 0 && dpt < 0) t2 -= p.orient_pert;
    else if (pt > 0 && dpt > 0) t2 += p.orient_pert;
    else if (pt < 0 && dpt > 0) t2 += p.orient_pert;
    if (tangential(a, b, c, x, y, t2, xi, yi) < p.tang_displ) { x = xi; y = yi; th = t2; }
}

__global__ void __launch_bounds__(128) shift_kernel(DevBatch b, DevParams p, int which)
{
    const int f = blockIdx.y;
    const int used = min(b.poolUsed[f], b.P);
    const bool dumps = b.dumps && f == 0 && which == 0;
    const int* c_owner = b.c_owner + (size_t)f * b.P;
    double *c_x = b.c_x + (size_t)f * b.P, *c_y = b.c_y + (size_t)f * b.P, *c_th = b.c_th + (size_t)f * b.P;
    for (int q = blockIdx.x * 128 + threadIdx.x; q < used; q += gridDim.x * 128) {
        const int i = c_owner[q];
        if (i < 0) continue;
        const double* ln = b.lines + ((size_t)f * b.E + i) * GEO;
        double x = c_x[q], y = c_y[q], th = c_th[q];
        shift_to_line(ln[0], ln[1], ln[2], p, x, y, th);
        c_x[q] = x; c_y[q] = y; c_th[q] = th;
        if (which == 0) b.c_ridx[(size_t)f * b.P + q] = -1;
        if (dumps) dump_put(b.dump[DUMP_S8], q, -1, x, y, th, b.c_score[q]);
    }
}

// ------------------------------------------------------------------------------------------------------
// S9 Gauss-Newton along the epipolar line (Stereo_Matches.cpp:1159-1358).  Three kernels (ebvo_params.gn_mode):
//   gn_lerp64_kernel (0, default)  reference arithmetic (FP64 blends rounded to float, FP64 residuals / weights /
//                              normal equations); persistent warps pulling pool-slot chunks, right-view samples
//                              served from warp-private shared-memory tiles that are kept across the candidates
//                              of a left edge, 3 sample rounds + a cooperative 49th sample
//   gn64_kernel (1)            the same arithmetic, one warp per candidate, samples gathered from global memory
//                              (the simple form; kept as the cross-check of the tiled kernel)
//   gn32_kernel (2)            everything FP32: 0.02 % of the mates move by > 1e-3 px (non-converging GN); opt-in
// Gauss-Newton sequences that do not converge amplify last-bit differences by up to ~1e8, and the clusterer that
// follows decides nearest-neighbour merges among candidates that converged to the same point within ~1e-6 px, so
// anything less than double precision in ANY channel changes 0.05 % of the final mates' orientations.
// ------------------------------------------------------------------------------------------------------
// include/utility.h:159-172 on an 8-bit image viewed as CV_32F (convertTo is exact): clamped sampler, FP32 blend
__device__ __forceinline__ void cell_of(double x, int n, int& x0, int& x1, float& a)
{
    x = fmin(fmax(x, 0.0), (double)n - 1.0);
    x0 = __double2int_rd(x);
    x1 = min(x0 + 1, n - 1);
    a = (float)(x - (double)x0);
}
__device__ __forceinline__ float sample_u8(const uint8_t* __restrict__ I, int pitch, int w, int h, double x, double y)
{
    int x0, x1, y0, y1;
    float a, bb;
    cell_of(x, w, x0, x1, a);
    cell_of(y, h, y0, y1, bb);
    const float v00 = (float)__ldg(I + (size_t)y0 * pitch + x0), v10 = (float)__ldg(I + (size_t)y0 * pitch + x1);
    const float v01 = (float)__ldg(I + (size_t)y1 * pitch + x0), v11 = (float)__ldg(I + (size_t)y1 * pitch + x1);
    return (1.f - a) * (1.f - bb) * v00 + a * (1.f - bb) * v10 + (1.f - a) * bb * v01 + a * bb * v11;
}
// the same sampler with the FP64 blend of the reference, result rounded to float (utility.h:171)
__device__ __forceinline__ double sample_u8_exact(const uint8_t* __restrict__ I, int pitch, int W, int H, double x, double y)
{
    x = fmin(fmax(x, 0.0), (double)W - 1.0); y = fmin(fmax(y, 0.0), (double)H - 1.0);
    const int x0 = __double2int_rd(x), y0 = __double2int_rd(y);
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    const double a = x - (double)x0, bb = y - (double)y0;
    const double v00 = (double)__ldg(I + (size_t)y0 * pitch + x0), v10 = (double)__ldg(I + (size_t)y0 * pitch + x1);
    const double v01 = (double)__ldg(I + (size_t)y1 * pitch + x0), v11 = (double)__ldg(I + (size_t)y1 * pitch + x1);
    return (double)(float)((1 - a) * (1 - bb) * v00 + a * (1 - bb) * v10 + (1 - a) * bb * v01 + a * bb * v11);
}

// (double)(float)v without conversion instructions: Veltkamp split keeping 24 significant bits (round to nearest).
// Valid for |v| in the float normal range, which holds for 8-bit image samples and their Sobel responses.
__device__ __forceinline__ double round_to_float(double v)
{
    // the same rounding on the bit pattern (integer pipe instead of three FP64 instructions): drop the low 29 mantissa bits
    // to nearest, ties to even; a mantissa carry runs into the exponent, which is the correct result
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    u += 0x0FFFFFFFull + ((u >> 29) & 1ull);
    return __longlong_as_double((long long)(u & ~0x1FFFFFFFull));
}
// exact integer -> double without conversion instructions (2^52 magic); fields of the packed right-view pixel
__device__ __forceinline__ double pk_i(uint2 u) { return __hiloint2double(0x43300000, (int)(u.x & 0xffffu)) - 4503599627370496.0; }
__device__ __forceinline__ double pk_gx(uint2 u) { return __hiloint2double(0x43300000, (((int)u.x) >> 16) ^ 0x80000000) - 4503601774854144.0; }
__device__ __forceinline__ double pk_gy(uint2 u) { return __hiloint2double(0x43300000, ((int)(u.y << 16) >> 16) ^ 0x80000000) - 4503601774854144.0; }

// floor + fraction of a clamped coordinate without conversions (utility.h:161-166): a round-down add of 1.5*2^52 puts
// floor(x) in the low word.  The clamp to [0, n-1] is applied to the cell; d1 (offset of the second corner) is 0
// when the coordinate was clamped, which makes both corners the same pixel, so the fraction needs no fix-up
// (the reference sets it to 0 there: (1-a)v + a v = v).
__device__ __forceinline__ void cell_magic(double x, int n, int stride, int& x0, int& d1, double& a)
{
    const double t = __dadd_rd(x, 6755399441055744.0);
    const int xu = __double2loint(t);
    a = x - (t - 6755399441055744.0);
    d1 = ((unsigned)xu < (unsigned)(n - 1)) ? stride : 0;
    x0 = min(max(xu, 0), n - 1);
}

struct GnSetup {
    double dirx, diry, xr, yr;
    double Bx[4], By[4], Lc[4];
};

// common prologue: geometry, per-sample base coordinates of the right patches, centred left samples
__device__ __forceinline__ void gn_setup(const DevBatch& b, int f, int i, int q, int lane, GnSetup& g)
{
    const int imgL = 2 * f;
    const uint8_t* IL = b.und + (size_t)imgL * b.imgStride;   // GN uses the UNDISTORTED images (:1293-1294)
    const double* ln = b.lines + ((size_t)f * b.E + i) * GEO;
    g.dirx = ln[3]; g.diry = ln[4];
    const double st_ = ln[5], ct_ = ln[6];
    const double xL = b.ex[(size_t)imgL * b.E + i], yL = b.ey[(size_t)imgL * b.E + i];
    const double side = 7 / 2.0 + 1.0;                         // :1171
    const double nxs = -st_ * side, nys = ct_ * side;          // n * side, n = (-t.y, t.x) (:1169-1170)
    g.xr = b.c_x[(size_t)f * b.P + q]; g.yr = b.c_y[(size_t)f * b.P + q];
    double sumP = 0, sumM = 0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int s = lane + 32 * m;
        const bool neg = s >= 49;
        const int t = s - (neg ? 49 : 0);
        const int ii = t / 7 - 3, jj = t % 7 - 3;
        const double cx = neg ? -nxs : nxs, cy = neg ? -nys : nys;        // +-n*side
        const double rx = ct_ * ii - st_ * jj, ry = st_ * ii + ct_ * jj;  // rotated cell (utility.h:154)
        g.Bx[m] = (g.xr + cx) + rx; g.By[m] = (g.yr + cy) + ry;           // right: + alpha*dir per iteration (:1203-1204)
        g.Lc[m] = 0.0;
        if (s < 98) {
            g.Lc[m] = sample_u8_exact(IL, b.pitch, b.W, b.H, (xL + cx) + rx, (yL + cy) + ry);
            if (neg) sumM += g.Lc[m]; else sumP += g.Lc[m];
        }
    }
    warp_sum2(sumP, sumM);
    const double mLp = sumP / 49.0, mLm = sumM / 49.0;
#pragma unroll
    for (int m = 0; m < 4; ++m) { const int s = lane + 32 * m; if (s < 98) g.Lc[m] -= (s >= 49) ? mLm : mLp; }
}

__device__ __forceinline__ void gn_store(const DevBatch& b, int f, int q, const GnSetup& g, double alpha, double score, double conf)
{
    const size_t o = (size_t)f * b.P + q;
    b.c_x[o] = g.xr + alpha * g.dirx;          // :1350-1352 (moved regardless of validity)
    b.c_y[o] = g.yr + alpha * g.diry;
    b.c_score[o] = score; b.c_conf[o] = conf;
}

// ------------------------------------------------------------------------------------------------------
// gn_lerp64_kernel (gn_mode 0, default).  Persistent warps pull chunks of pool slots from per-frame cursors; a warp works
// on one candidate at a time and samples the left patches once per LEFT EDGE (consecutive slots belong to one edge).
// Lanes 0-15 own the "+" patch, lanes 16-31 the "-" patch (cell t = hl + 16 m, m < 3), so the two patch means reduce
// inside half-warps; the 49th cell (3,3) of each patch is evaluated cooperatively by its own half-warp: lane hl < 6
// handles (channel, cell row), the pairs meet by one shuffle.  Per candidate each half-warp stages the pixels its
// patch can reach while alpha stays within +-R px of the build position into a warp-private tile of packed
// {half I, -, half gx, half gy} pixels (exact values, 8 B each, widened to double per corner by F2F.F64.F16).
// Coordinates beyond the image read the clamped border pixel, which is what util_bilinear_Sample_F's coordinate clamp
// produces (utility.h:161-166).  The tile is rebuilt when alpha leaves the window (3 % of the candidates) and SURVIVES
// to the next candidate of the same left edge when that one's patch fits in it.  Arithmetic: the four-corner blend as
// two horizontal interpolations and one vertical one, top = v00 + a (v10 - v00), bot = v01 + a (v11 - v01),
// v = top + b (bot - top), per channel in FP64 (corner differences exact in fp16: |dI| <= 255, |8 dg| <= 2040 < 2^11,
// taken by HSUB2 on the packed pixels), each rounded to float (util_bilinear_Sample_F returns float) on the integer
// pipe; FP64 residuals, Huber weights and normal equations; divisions by 49, |r| and H by reciprocal + FMA
// correction (within 1 ulp; same class as the summation order).  Against the reference: 9e-7 px at the GN stage
// over 128 913 candidates (one non-converging candidate; p99.9 7e-13), 6e-8 px on the final mates.
// What bounds it and what was tried in round 2: profiles/r02_gn_whatif.md, profiles/r02_gn_pair_experiment.md.
// ------------------------------------------------------------------------------------------------------
#ifndef GNL_MINB
#define GNL_MINB 5
#endif
#ifndef GNL_PXV
#define GNL_PXV 256
#endif
constexpr int GNL_PX = GNL_PXV;
#ifndef GN_CHUNKV
#define GN_CHUNKV 32
#endif
constexpr int GN_CHUNK = GN_CHUNKV;        // pool slots fetched per atomic (at most 32: one lane per slot)
#ifndef GN_RV
#define GN_RV 5
#endif
constexpr int GN_R = GN_RV;            // tile reach along the epipolar direction (px) before a rebuild

__device__ __forceinline__ double half_sum(double v)   // sum over the 16 lanes of a half-warp, result in every lane
{
#pragma unroll
    for (int o = 8; o; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}
// a / 49 correctly rounded without the division sequence (Markstein: q = a*y, r = a - 49 q exactly, q + r*y)
__device__ __forceinline__ double div49(double a)
{
    const double y = 1.0 / 49.0;
    const double q = a * y;
    const double r = fma(-49.0, q, a);
    return fma(r, y, q);
}
// 1 / x to ~1 ulp: MUFU.RCP64H seed (about 20 bits) + two Newton steps; x is a finite positive normal number here
__device__ __forceinline__ double rcp_fast(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);        // one third-order step: y (1 + e + e^2), |e| < 2^-20 => error 2^-60
    return fma(y, fma(e, e, e), y);
}
// a / b (b finite, positive, normal) correctly rounded in all but pathological cases, without the division sequence
__device__ __forceinline__ double div_fast(double a, double b)
{
    const double y = rcp_fast(b);
    const double q = a * y;
    const double r = fma(-b, q, a);
    return fma(r, y, q);
}
// exact fp16 -> fp64 (one F2F.F64.F16, the half selected in place from the low 16 bits of the argument)
__device__ __forceinline__ double h2d(unsigned int lo16)
{
    double d;
    asm("cvt.f64.f16 %0, %1;" : "=d"(d) : "h"((unsigned short)lo16));
    return d;
}
__device__ __forceinline__ unsigned hsub2_u32(unsigned a, unsigned b)   // a - b on both halves
{
    unsigned d;
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
template <int GT64_MAXPX, int MINB>
__global__ void __launch_bounds__(32 * WPB, MINB) gn_lerp64_kernel(DevBatch b, DevParams p, int Rmax, int nFrames)
{
    __shared__ uint2 s_tile[WPB][2][GT64_MAXPX];
    // the per-left-edge sample geometry and centred left samples live in shared memory (lane-contiguous, conflict-free reads): the 24
    // registers they would hold are worth more to the scheduler's interleaving of the three sample chains (-2.6 %; a fifth CTA per
    // SM bought with them instead is not: 20 warps run no faster than 16)
    __shared__ double s_state[WPB][12][32];
    double (*SS)[32] = s_state[threadIdx.x >> 5];
#define RX(m) SS[m][lane]
#define RY(m) SS[3 + (m)][lane]
#define LC(m) SS[6 + (m)][lane]
#define RX48 SS[9][lane]
#define RY48 SS[10][lane]
#define LC48 SS[11][lane]
    uint2* tF = s_tile[threadIdx.x >> 5][(threadIdx.x & 31) >> 4];
    const int lane = threadIdx.x & 31;
    const int hw = lane >> 4, hl = lane & 15;
    // cooperative lanes of the left-over sample of THIS half-warp's patch: lane hl < 6 = (channel, cell row), u = 2 channel + row
    // (lanes 6-15 repeat them), so every lane works on its own half-warp's tile and patch centre
    const int u = hl % 6;
    const int cRow = u & 1, cCh = u >> 1;
    const int W = b.W, H = b.H;
    const double huber = p.gn_huber;
    const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: a round-down add leaves floor(x) in the low word
    const int f0 = (int)(((long long)blockIdx.x * nFrames) / gridDim.x);
    for (int ff = 0; ff < nFrames; ++ff) {
        const int f = (f0 + ff) % nFrames;
        const int imgL = 2 * f;
        const uint8_t* IL = b.und + (size_t)imgL * b.imgStride;         // GN uses the UNDISTORTED images (:1293-1294)
        const uint2* __restrict__ PK = b.pkh + (size_t)f * b.gStride;   // right view: {half I, -, half gx, half gy}
        const int* c_owner = b.c_owner + (size_t)f * b.P;
        double *c_x = b.c_x + (size_t)f * b.P, *c_y = b.c_y + (size_t)f * b.P;
        double *c_score = b.c_score + (size_t)f * b.P, *c_conf = b.c_conf + (size_t)f * b.P;
        unsigned long long* cursor = b.counters + (size_t)f * 8 + 7;
        const int used = min(b.poolUsed[f], b.P);
        unsigned long long npairs = 0, niters = 0, nbuilds = 0;
        for (;;) {
            int q0 = 0;
            if (lane == 0) q0 = (int)atomicAdd(cursor, (unsigned long long)GN_CHUNK);
            q0 = __shfl_sync(FULL, q0, 0);
            if (q0 >= used) break;
            const int q1 = min(q0 + GN_CHUNK, used);
            // the chunk's slots are fetched by the first GN_CHUNK lanes at once and handed out by shuffles: no global-memory
            // round trip between two candidates
            int own_l = -1;
            double x_l = 0, y_l = 0;
            if (lane < GN_CHUNK && q0 + lane < q1) { own_l = c_owner[q0 + lane]; x_l = c_x[q0 + lane]; y_l = c_y[q0 + lane]; }
            int owner = -1;
            double dirx = 0, diry = 0, cx = 0, cy = 0, hext = 0;
            int Rint = 0;
            bool tileValid = false;
            int TWp = 0, THp = 0, npx = 0, ox = 0, oy = 0;
            float invTW = 0.f;
            for (int q = q0; q < q1; ++q) {
                const int i = __shfl_sync(FULL, own_l, q - q0);
                const double xr = shfl_idx_d(x_l, q - q0), yr = shfl_idx_d(y_l, q - q0);
                if (i < 0) continue;          // dead slot (dropped by NCC / best-nearly-best)
                EBVO_ASSERT(b.errFlag + f, i < b.nE[imgL] && q < b.P);
                if (i != owner) {
                    owner = i;
                    // ---- per left edge: geometry, centred left samples, tile shape ----
                    const double* ln = b.lines + ((size_t)f * b.E + i) * GEO;
                    dirx = ln[3]; diry = ln[4];
                    const double st_ = ln[5], ct_ = ln[6];
                    const double xL = b.ex[(size_t)imgL * b.E + i], yL = b.ey[(size_t)imgL * b.E + i];
                    const double side = 7 / 2.0 + 1.0;                               // :1171
                    cx = hw ? st_ * side : -st_ * side;                              // +-n*side, n = (-t.y, t.x) (:1169-1170)
                    cy = hw ? -ct_ * side : ct_ * side;
                    double sumL = 0, lcv[3];
#pragma unroll
                    for (int m = 0; m < 3; ++m) {
                        const int t = hl + 16 * m;
                        const int ii = t / 7 - 3, jj = t % 7 - 3;
                        const double rxm = ct_ * ii - st_ * jj, rym = st_ * ii + ct_ * jj;    // rotated cell (utility.h:154)
                        RX(m) = rxm; RY(m) = rym;
                        lcv[m] = sample_u8_exact(IL, b.pitch, W, H, (xL + cx) + rxm, (yL + cy) + rym); sumL += lcv[m];
                    }
                    const double rx48v = ct_ * 3 - st_ * 3, ry48v = st_ * 3 + ct_ * 3;              // cell (3, 3), t = 48
                    RX48 = rx48v; RY48 = ry48v;
                    const double lc48v = sample_u8_exact(IL, b.pitch, W, H, (xL + cx) + rx48v, (yL + cy) + ry48v);
                    sumL = half_sum(sumL) + lc48v;
                    const double mL = sumL / 49.0;
#pragma unroll
                    for (int m = 0; m < 3; ++m) LC(m) = lcv[m] - mL;
                    LC48 = lc48v - mL;
                    // tile: every sample of this patch stays within (centre +- (R|dir| + hext)) while |alpha - alpha0| <= R
                    hext = 3.0 * (fabs(ct_) + fabs(st_)) + 1e-6;
                    int R = Rmax;
                    for (;;) {
                        TWp = (int)ceil(2.0 * (R * fabs(dirx) + hext)) + 2;
                        THp = (int)ceil(2.0 * (R * fabs(diry) + hext)) + 2;
                        // row pitch in 8-byte pixels: residues 0, +-1, +-2 and 8 (mod 16 bank pairs) fold neighbouring rows onto the same banks
                        while ((0xC107 >> (TWp & 15)) & 1) ++TWp;
                        if (TWp * THp <= GT64_MAXPX || R == 0) break;
                        --R;
                    }
                    Rint = R;
                    npx = TWp * THp;
                    invTW = 1.0f / (float)TWp;
                    tileValid = false;           // the tile shape belongs to the left edge
                }
                {
                const double xc = xr + cx, yc = yr + cy;      // patch centre at alpha = 0 (:1203-1204)
                double alpha = 0.0, score = 0.0, conf = 0.0, alpha0 = CUDART_NAN;
                double Rv = (double)Rint - 1e-6;
                if (tileValid) {
                    // 78 % of the consecutive candidates of a left edge lie within 1 px of each other: keep the tile of the
                    // previous candidate when this one's patch (at alpha = 0) sits inside it, with the reach the margins leave
                    const double mx = fmin((xc - hext) - (double)ox, ((double)(ox + TWp - 1) - hext) - xc);
                    const double my = fmin((yc - hext) - (double)oy, ((double)(oy + THp - 1) - hext) - yc);
                    double reach = fmin(fmin(mx / fabs(dirx), my / fabs(diry)), (double)Rint);   // x / 0 = inf, 0 / 0 = NaN is ignored by fmin
                    if (!(mx >= 0.0 && my >= 0.0)) reach = -1.0;
                    reach = fmin(reach, shfl_xor_d(reach, 16));
                    if (reach >= 1.0) { alpha0 = 0.0; Rv = reach - 1e-6; }
                }
                // rectified pairs (KITTI): the first row of F is exactly zero, so every epipolar direction is (+-1, 0) and the vertical Sobel
                // channel enters the Jacobian with the factor 0 (the Sobel values are finite): same bits without its interpolation
                const bool useGy = diry != 0.0;
                for (int it = 0; it < p.gn_max_iter; ++it) {
                    const double xs = xc + alpha * dirx, ys = yc + alpha * diry;     // (location +- n*side) + shift (:1203-1204)
                    if (!(fabs(alpha - alpha0) <= Rv)) {
                        // ---- (re)build this half-warp's sub-tile around the current position ----
                        alpha0 = alpha; Rv = (double)Rint - 1e-6; tileValid = true;
                        ox = __double2int_rd(xs - (Rint * fabs(dirx) + hext));
                        oy = __double2int_rd(ys - (Rint * fabs(diry) + hext));
                        ++nbuilds;
                        __syncwarp();
                        for (int e = hl; e < npx; e += 16) {
                            const int py = (int)(((float)e + 0.5f) * invTW), px = e - py * TWp;
                            const int X = min(max(ox + px, 0), W - 1), Y = min(max(oy + py, 0), H - 1);
                            tF[e] = __ldg(PK + (Y * W + X));
                        }
                        __syncwarp();
                    }
                    double vi[3], vg[3];       // (kept in registers: staging them in shared memory as well costs 4 %)
#define VI(m) vi[m]
#define VG(m) vg[m]
                    double sR = 0;
#pragma unroll
                    for (int m = 0; m < 3; ++m) {
                        const double x = xs + RX(m), y = ys + RY(m);
                        const double tx = __dadd_rd(x, MAGIC), ty = __dadd_rd(y, MAGIC);
                        const double a = x - (tx - MAGIC), bb = y - (ty - MAGIC);
                        const int xi = (int)min((unsigned)(__double2loint(tx) - ox), (unsigned)(TWp - 2));
                        const int yi = (int)min((unsigned)(__double2loint(ty) - oy), (unsigned)(THp - 2));
                        const int o = yi * TWp + xi;
                        EBVO_ASSERT(b.errFlag + f, o >= 0 && o + TWp + 1 < npx && npx <= GT64_MAXPX);
                        const uint2 p00 = tF[o], p10 = tF[o + 1], p01 = tF[o + TWp], p11 = tF[o + TWp + 1];
                        const unsigned d0x = hsub2_u32(p10.x, p00.x), d0y = hsub2_u32(p10.y, p00.y);
                        const unsigned d1x = hsub2_u32(p11.x, p01.x), d1y = hsub2_u32(p11.y, p01.y);
                        // FP64 interpolation rounded to float = util_bilinear_Sample_F (utility.h:159-172) per channel
                        double top = fma(a, h2d(d0x), h2d(p00.x)), bot = fma(a, h2d(d1x), h2d(p01.x));
                        const double vim = round_to_float(fma(bb, bot - top, top));
                        VI(m) = vim;
                        top = fma(a, h2d(d0y), h2d(p00.y)); bot = fma(a, h2d(d1y), h2d(p01.y));
                        const double gx = round_to_float(fma(bb, bot - top, top));
                        if (useGy) {
                            top = fma(a, h2d(d0y >> 16), h2d(p00.y >> 16)); bot = fma(a, h2d(d1y >> 16), h2d(p01.y >> 16));
                            const double gy = round_to_float(fma(bb, bot - top, top));
                            VG(m) = -gx * dirx + gy * diry;                               // :1240
                        } else VG(m) = -gx * dirx;     // horizontal epipolar line: gy * 0 adds an exact zero, its four-corner blend is skipped
                        sR += vim;
                    }
                    // ---- the 49th sample of both patches, one (patch, channel, cell row) per lane ----
                    double vi48, vg48;
                    {
                        const double x = xs + RX48, y = ys + RY48;
                        const double tx = __dadd_rd(x, MAGIC), ty = __dadd_rd(y, MAGIC);
                        const double a = x - (tx - MAGIC), bb = y - (ty - MAGIC);
                        const int xi = (int)min((unsigned)(__double2loint(tx) - ox), (unsigned)(TWp - 2));
                        const int yi = (int)min((unsigned)(__double2loint(ty) - oy), (unsigned)(THp - 2));
                        const int o = (yi + cRow) * TWp + xi;
                        EBVO_ASSERT(b.errFlag + f, o >= 0 && o + 1 < npx);
                        const uint2 p0 = tF[o], p1 = tF[o + 1];
                        const unsigned s0 = cCh == 0 ? p0.x : (cCh == 1 ? p0.y : p0.y >> 16);
                        const unsigned s1 = cCh == 0 ? p1.x : (cCh == 1 ? p1.y : p1.y >> 16);
                        const double lin = fma(a, h2d(hsub2_u32(s1, s0)), h2d(s0));       // top (row 0) or bottom (row 1)
                        const double oth = shfl_xor_d(lin, 1);
                        const double top = cRow ? oth : lin, bot = cRow ? lin : oth;
                        const double v = round_to_float(fma(bb, bot - top, top));
                        vi48 = shfl_idx_d(v, hw << 4);
                        const double gx = shfl_idx_d(v, (hw << 4) + 2), gy = shfl_idx_d(v, (hw << 4) + 4);
                        vg48 = -gx * dirx + gy * diry;
                    }
                    sR = half_sum(sR) + vi48;
                    const double mR = div49(sR);
                    double Hh = 0, bb_ = 0, cost = 0;
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const double r = (m < 3 ? LC(m < 3 ? m : 0) : LC48) - ((m < 3 ? VI(m < 3 ? m : 0) : vi48) - mR);
                        const double gg = m < 3 ? VG(m < 3 ? m : 0) : vg48;
                        const double ar = fabs(r);
                        double wgt = (ar <= huber) ? 1.0 : huber * rcp_fast(ar);
                        if (m == 3 && hl != 0) wgt = 0.0;                                 // sample 48 counts once per patch
                        const double wg = wgt * gg;
                        Hh = fma(wg, gg, Hh); bb_ = fma(wg, r, bb_); cost = fma(wgt * r, r, cost);
                    }
                    warp_sum2t(Hh, bb_, lane);
                    ++niters;
                    if (Hh < 1e-8) break;                          // :1253 (outputs stay at their initial values)
                    const double delta = -div_fast(bb_, Hh);
                    alpha += delta;
                    if (fabs(delta) < p.gn_tol || it == p.gn_max_iter - 1) {
                        cost = warp_sum(cost);
                        const double rms = sqrt(cost / 98.0);
                        score = rms; conf = exp(-rms / p.gn_huber);
                        break;
                    }
                }
                ++npairs;
                if (lane == 0) {
                    c_x[q] = xr + alpha * dirx;          // :1350-1352 (moved regardless of validity)
                    c_y[q] = yr + alpha * diry;
                    c_score[q] = score; c_conf[q] = conf;
                }
                }
            }
        }
        if (lane == 0 && npairs) {
            atomicAdd(&b.counters[(size_t)f * 8 + 2], npairs); atomicAdd(&b.counters[(size_t)f * 8 + 3], niters);
            atomicAdd(&b.counters[(size_t)f * 8 + 5], nbuilds);
        }
    }
}
#undef VI
#undef VG
#undef RX
#undef RY
#undef LC
#undef RX48
#undef RY48
#undef LC48

__global__ void __launch_bounds__(32 * WPB, 5) gn64_kernel(DevBatch b, DevParams p)
{
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const uint2* __restrict__ PK16 = b.pk16 + (size_t)f * b.gStride;   // right view: {I, 8*Sobel gx, 8*Sobel gy} as int16
    const int* c_owner = b.c_owner + (size_t)f * b.P;
    const int used = min(b.poolUsed[f], b.P);
    const int W = b.W, H = b.H;
    unsigned long long npairs = 0, niters = 0;
    for (int q = blockIdx.x * WPB + (threadIdx.x >> 5); q < used; q += gridDim.x * WPB) {
        const int i = c_owner[q];
        if (i < 0) continue;
        GnSetup g;
        gn_setup(b, f, i, q, lane, g);
        double alpha = 0.0, score = 0.0, conf = 0.0;
        for (int it = 0; it < p.gn_max_iter; ++it) {
            const double sx = alpha * g.dirx, sy = alpha * g.diry;
            double vi[4], vg[4];
            double sRp = 0, sRm = 0;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                vi[m] = 0.0; vg[m] = 0.0;
                const int s = lane + 32 * m;
                if (s < 98) {
                    int x0, y0, dx1, dy1;
                    double a, bb;
                    cell_magic(g.Bx[m] + sx, W, 1, x0, dx1, a);
                    cell_magic(g.By[m] + sy, H, W, y0, dy1, bb);
                    const uint2* c00 = PK16 + (y0 * W + x0);
                    const uint2 u00 = __ldg(c00), u10 = __ldg(c00 + dx1), u01 = __ldg(c00 + dy1), u11 = __ldg(c00 + dy1 + dx1);
                    const double w00 = (1 - a) * (1 - bb), w10 = a * (1 - bb), w01 = (1 - a) * bb, w11 = a * bb;
                    // FP64 blend, then rounded to float as util_bilinear_Sample_F returns float (Veltkamp split: 24-bit RN)
                    const double bi = w00 * pk_i(u00) + w10 * pk_i(u10) + w01 * pk_i(u01) + w11 * pk_i(u11);
                    const double bgx = w00 * pk_gx(u00) + w10 * pk_gx(u10) + w01 * pk_gx(u01) + w11 * pk_gx(u11);   // 8 * gx
                    const double bgy = w00 * pk_gy(u00) + w10 * pk_gy(u10) + w01 * pk_gy(u01) + w11 * pk_gy(u11);   // 8 * gy
                    vi[m] = round_to_float(bi);
                    vg[m] = (-round_to_float(bgx) * g.dirx + round_to_float(bgy) * g.diry) * 0.125;   // :1240 (x 1/8: exact)
                    if (s >= 49) sRm += vi[m]; else sRp += vi[m];
                }
            }
            warp_sum2(sRp, sRm);
            const double mRp = sRp / 49.0, mRm = sRm / 49.0;
            double Hh = 0, bb_ = 0, cost = 0;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int s = lane + 32 * m;
                if (s < 98) {
                    const double r = g.Lc[m] - (vi[m] - ((s >= 49) ? mRm : mRp));
                    const double gg = vg[m];
                    const double ar = fabs(r);
                    const double wgt = (ar <= p.gn_huber) ? 1.0 : p.gn_huber / ar;
                    Hh += wgt * gg * gg; bb_ += wgt * gg * r; cost += wgt * r * r;
                }
            }
            warp_sum3(Hh, bb_, cost);
            ++niters;
            if (Hh < 1e-8) break;                          // :1253 (outputs stay at their initial values)
            const double delta = -bb_ / Hh;
            alpha += delta;
            if (fabs(delta) < p.gn_tol || it == p.gn_max_iter - 1) {
                const double rms = sqrt(cost / 98.0);
                score = rms; conf = exp(-rms / p.gn_huber);
                break;
            }
        }
        ++npairs;
        if (lane == 0) gn_store(b, f, q, g, alpha, score, conf);
    }
    if (lane == 0 && npairs) { atomicAdd(&b.counters[(size_t)f * 8 + 2], npairs); atomicAdd(&b.counters[(size_t)f * 8 + 3], niters); }
}

__global__ void __launch_bounds__(32 * WPB, 6) gn32_kernel(DevBatch b, DevParams p)
{
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int imgL = 2 * f;
    const uint8_t* IL = b.und + (size_t)imgL * b.imgStride;
    const float4* __restrict__ PK = b.pk + (size_t)f * b.gStride;   // right view: {I, Sobel gx, Sobel gy}
    const int* c_owner = b.c_owner + (size_t)f * b.P;
    const int used = min(b.poolUsed[f], b.P);
    const int W = b.W, H = b.H;
    const float huber = (float)p.gn_huber, tol = (float)p.gn_tol;
    unsigned long long npairs = 0, niters = 0;
    for (int q = blockIdx.x * WPB + (threadIdx.x >> 5); q < used; q += gridDim.x * WPB) {
        const int i = c_owner[q];
        if (i < 0) continue;
        const double* ln = b.lines + ((size_t)f * b.E + i) * GEO;
        const double dirx = ln[3], diry = ln[4], st_ = ln[5], ct_ = ln[6];
        const double xL = b.ex[(size_t)imgL * b.E + i], yL = b.ey[(size_t)imgL * b.E + i];
        const double side = 7 / 2.0 + 1.0;
        const double nxs = -st_ * side, nys = ct_ * side;
        const double xr = b.c_x[(size_t)f * b.P + q], yr = b.c_y[(size_t)f * b.P + q];
        // Right samples: FP32 offsets from the integer anchor (ax, ay) of the candidate; alpha*dir added in FP32
        const int ax = __double2int_rd(xr), ay = __double2int_rd(yr);
        const float fxr = (float)(xr - (double)ax), fyr = (float)(yr - (double)ay);
        float ox[4], oy[4], Lc[4];
        float sumP = 0.f, sumM = 0.f;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int s = lane + 32 * m;
            const bool neg = s >= 49;
            const int t = s - (neg ? 49 : 0);
            const int ii = t / 7 - 3, jj = t % 7 - 3;
            const double dox = (neg ? -nxs : nxs) + (ct_ * ii - st_ * jj);
            const double doy = (neg ? -nys : nys) + (st_ * ii + ct_ * jj);
            ox[m] = fxr + (float)dox; oy[m] = fyr + (float)doy;
            Lc[m] = 0.f;
            if (s < 98) {
                Lc[m] = sample_u8(IL, b.pitch, W, H, xL + dox, yL + doy);
                if (neg) sumM += Lc[m]; else sumP += Lc[m];
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { sumP += __shfl_xor_sync(FULL, sumP, o); sumM += __shfl_xor_sync(FULL, sumM, o); }
        const float mLp = sumP * (1.f / 49.f), mLm = sumM * (1.f / 49.f);
#pragma unroll
        for (int m = 0; m < 4; ++m) { const int s = lane + 32 * m; if (s < 98) Lc[m] -= (s >= 49) ? mLm : mLp; }

        double alpha = 0.0;
        float score = 0.f, conf = 0.f;
        const float fdx = (float)dirx, fdy = (float)diry;
        for (int it = 0; it < p.gn_max_iter; ++it) {
            const float sx = (float)(alpha * dirx), sy = (float)(alpha * diry);
            float vi[4], vg[4];
            float sRp = 0.f, sRm = 0.f;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                vi[m] = 0.f; vg[m] = 0.f;
                const int s = lane + 32 * m;
                if (s < 98) {
                    const float xf = ox[m] + sx, yf = oy[m] + sy;
                    const float flx = floorf(xf), fly = floorf(yf);
                    int x0 = ax + (int)flx, y0 = ay + (int)fly;
                    float a = xf - flx, bb = yf - fly;
                    if (x0 < 0) { x0 = 0; a = 0.f; } else if (x0 >= W - 1) { x0 = W - 1; a = 0.f; }
                    if (y0 < 0) { y0 = 0; bb = 0.f; } else if (y0 >= H - 1) { y0 = H - 1; bb = 0.f; }
                    const int dx1 = (x0 < W - 1) ? 1 : 0, dy1 = (y0 < H - 1) ? W : 0;
                    const float4* c00 = PK + (y0 * W + x0);
                    const float4 p00 = __ldg(c00), p10 = __ldg(c00 + dx1), p01 = __ldg(c00 + dy1), p11 = __ldg(c00 + dy1 + dx1);
                    const float w00 = (1.f - a) * (1.f - bb), w10 = a * (1.f - bb), w01 = (1.f - a) * bb, w11 = a * bb;
                    vi[m] = w00 * p00.x + w10 * p10.x + w01 * p01.x + w11 * p11.x;
                    const float gx = w00 * p00.y + w10 * p10.y + w01 * p01.y + w11 * p11.y;
                    const float gy = w00 * p00.z + w10 * p10.z + w01 * p01.z + w11 * p11.z;
                    vg[m] = -gx * fdx + gy * fdy;
                    if (s >= 49) sRm += vi[m]; else sRp += vi[m];
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) { sRp += __shfl_xor_sync(FULL, sRp, o); sRm += __shfl_xor_sync(FULL, sRm, o); }
            const float mRp = sRp * (1.f / 49.f), mRm = sRm * (1.f / 49.f);
            float Hh = 0.f, bb_ = 0.f, cost = 0.f;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int s = lane + 32 * m;
                if (s < 98) {
                    const float r = Lc[m] - (vi[m] - ((s >= 49) ? mRm : mRp));
                    const float gg = vg[m];
                    const float ar = fabsf(r);
                    const float wgt = (ar <= huber) ? 1.f : __fdividef(huber, ar);
                    const float wg = wgt * gg;
                    Hh = fmaf(wg, gg, Hh); bb_ = fmaf(wg, r, bb_); cost = fmaf(wgt * r, r, cost);
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                Hh += __shfl_xor_sync(FULL, Hh, o); bb_ += __shfl_xor_sync(FULL, bb_, o); cost += __shfl_xor_sync(FULL, cost, o);
            }
            ++niters;
            if (Hh < 1e-8f) break;
            const float delta = -bb_ / Hh;
            alpha += (double)delta;
            if (fabsf(delta) < tol || it == p.gn_max_iter - 1) {
                const float rms = sqrtf(cost * (1.f / 98.f));
                score = rms; conf = __expf(-rms / huber);
                break;
            }
        }
        ++npairs;
        if (lane == 0) {
            const size_t o = (size_t)f * b.P + q;
            b.c_x[o] = xr + alpha * dirx; b.c_y[o] = yr + alpha * diry;
            b.c_score[o] = (double)score; b.c_conf[o] = (double)conf;
        }
    }
    if (lane == 0 && npairs) { atomicAdd(&b.counters[(size_t)f * 8 + 2], npairs); atomicAdd(&b.counters[(size_t)f * 8 + 3], niters); }
}

// ------------------------------------------------------------------------------------------------------
// S10 EdgeClusterer, then S11 NCC on the cluster centres + S12 arg-max.  One warp per left edge.
// ------------------------------------------------------------------------------------------------------
// EdgeClusterer::performClustering (EdgeClusterer.cpp:119-302) on n <= MAXC warp-private candidates.
// Returns the number of clusters; centres go to (ox, oy, oth)[0..ncl) in ascending-label order (the order of
// returned_clusters); lab[] receives the renumbered label of every input; csz[] (n ints) and dk[], gk[] (n doubles) are scratch.
//
// The reference repeats: scan the points in index order; the first point whose nearest admissible neighbour
// (other cluster, d < 1 px, |dtheta| < 20 deg when clustering by orientation; first minimum on ties) can be merged
// without exceeding MAX_CLUSTER_SIZE triggers a merge and restarts the scan.  Here every lane evaluates "its"
// point's nearest admissible neighbour at once and a ballot picks the first feasible point: same merges, same
// order, O(n) work per merge instead of O(n^2).  Distances are compared squared (sqrt is monotone; d < 1 <=> d^2 < 1).
// adm != nullptr (the launch for big sets, n > 48): the points do not move while clusters merge, so which pairs are
// admissible (d < 1 px, |dtheta| < 20 deg) is decided once, as a 2 x 64-bit mask per point, and every rescan visits only
// a point's admissible neighbours in ascending index order (O(degree) instead of O(n) per rescan; for the usual
// handful of candidates the mask bookkeeping costs more than it saves, so the small launch passes nullptr).
__device__ int warp_cluster(const double* sx, const double* sy, const double* sth, int n, bool by_orient, const DevParams& p,
                            int lane, int* lab, int* csz, double* dk, double* gk, double* ox, double* oy, double* oth,
                            unsigned long long* adm = nullptr)
{
    const double d2max = p.clus_dist * p.clus_dist;
    for (int i = lane; i < n; i += 32) {
        lab[i] = i; csz[i] = 1;
        if (adm) {
            const double xi = sx[i], yi = sy[i], ti = sth[i];
            unsigned long long m0 = 0, m1 = 0;
            for (int j = 0; j < n; ++j) {
                const double dx = xi - sx[j], dy = yi - sy[j];
                const double d2 = dx * dx + dy * dy;
                if (j != i && d2 < d2max && (!by_orient || fabs(ti - sth[j]) < p.clus_orient_rad)) { if (j < 64) m0 |= 1ull << j; else m1 |= 1ull << (j - 64); }
            }
            adm[2 * i] = m0; adm[2 * i + 1] = m1;
        }
    }
    __syncwarp();
    for (;;) {
        int pick_i = -1, pick_j = -1;
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int i = c0 + lane;
            int bj = -1;
            if (i < n) {
                const int li = lab[i];
                const double xi = sx[i], yi = sy[i], ti = sth[i];
                double best = CUDART_INF;
                if (adm) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        unsigned long long wbits = adm[2 * i + h], live = wbits;
                        while (wbits) {
                            const int bit = __ffsll((long long)wbits) - 1, j = bit + 64 * h;
                            wbits &= wbits - 1;
                            if (lab[j] == li) { live &= ~(1ull << bit); continue; }      // same cluster now, and for good: never looked at again
                            const double dx = xi - sx[j], dy = yi - sy[j];
                            const double d2 = dx * dx + dy * dy;
                            if (d2 < best) { best = d2; bj = j; }
                        }
                        adm[2 * i + h] = live;
                    }
                } else {
                    for (int j = 0; j < n; ++j) {
                        if (lab[j] == li) continue;
                        const double dx = xi - sx[j], dy = yi - sy[j];
                        const double d2 = dx * dx + dy * dy;
                        if (d2 < best && d2 < d2max && (!by_orient || fabs(ti - sth[j]) < p.clus_orient_rad)) { best = d2; bj = j; }
                    }
                }
                if (bj >= 0 && csz[li] + csz[lab[bj]] > p.clus_max) bj = -1;   // MAX_CLUSTER_SIZE, EdgeClusterer.cpp:179
            }
            const unsigned m = __ballot_sync(FULL, bj >= 0);
            if (m) {
                const int src = __ffs(m) - 1;
                pick_i = c0 + src;
                pick_j = __shfl_sync(FULL, bj, src);
                break;
            }
        }
        if (pick_i < 0) break;
        const int li = lab[pick_i], lold = lab[pick_j];
        const int merged_size = csz[li] + csz[lold];
        __syncwarp();
        for (int k = lane; k < n; k += 32) if (lab[k] == lold) lab[k] = li;
        if (lane == 0) csz[li] = merged_size;
        __syncwarp();
    }
    // Gaussian-weighted centre of every cluster (:43-117).  A live label L always labels element L itself, so "label L
    // exists" <=> lab[L] == L.  The transcendental work (sqrt, exp) is done by one lane per MEMBER; the sums are done
    // by one lane per CLUSTER walking its members in index order, i.e. in the reference's summation order.
    for (int L = lane; L < n; L += 32) if (lab[L] == L) {          // centroid (:55-74), label-indexed scratch in ox/oy
        double sxx = 0, syy = 0;
        for (int k = 0; k < n; ++k) if (lab[k] == L) { sxx += sx[k]; syy += sy[k]; }
        ox[L] = sxx / csz[L]; oy[L] = syy / csz[L];
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) { const int L = lab[k]; const double dx = sx[k] - ox[L], dy = sy[k] - oy[L]; dk[k] = sqrt(dx * dx + dy * dy); }
    __syncwarp();
    for (int L = lane; L < n; L += 32) if (lab[L] == L) {          // mean distance from the centroid (:77-88)
        double tot = 0;
        for (int k = 0; k < n; ++k) if (lab[k] == L) tot += dk[k];
        oth[L] = tot / csz[L];
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) { const double z = (dk[k] - oth[lab[k]]) / p.clus_sigma; gk[k] = exp(-0.5 * (z * z)); }   // :103
    __syncwarp();
    for (int L = lane; L < n; L += 32) if (lab[L] == L) {          // weighted means (:105-114)
        double wx = 0, wy = 0, wt = 0, ww = 0;
        for (int k = 0; k < n; ++k) if (lab[k] == L) { const double g = gk[k]; wx += g * sx[k]; wy += g * sy[k]; wt += g * sth[k]; ww += g; }
        ox[L] = wx / ww; oy[L] = wy / ww; oth[L] = wt / ww;
    }
    __syncwarp();
    // clusters in ascending label order (std::map, :209-222): ordered compaction label index -> cluster index
    int ncl = 0;
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int L = c0 + lane;
        const bool live = L < n && lab[L] == L;
        const unsigned m = __ballot_sync(FULL, live);
        const int c = ncl + __popc(m & ((1u << lane) - 1));
        double vx = 0, vy = 0, vt = 0;
        if (live) { vx = ox[L]; vy = oy[L]; vt = oth[L]; }
        __syncwarp();
        if (live) { ox[c] = vx; oy[c] = vy; oth[c] = vt; csz[L] = -1 - c; }
        __syncwarp();
        ncl += __popc(m);
    }
    for (int k = lane; k < n; k += 32) { const int L = lab[k]; lab[k] = -1 - csz[L]; }
    __syncwarp();
    return ncl;
}

// Sets of at most 8 candidates (the usual case: ~4 per left edge after best-nearly-best) are clustered by a QUARTER-warp each,
// four left edges per warp in lock step: a full warp per set left 75 % of the lanes idle and spent ~1000 warp instructions per
// set on serial phases.  Same merges, same order, same summation order as warp_cluster (lane gl owns point gl; "label L
// exists" <=> lab[L] == L; first feasible point in index order wins the ballot).  Larger sets are appended to a per-frame
// work list (indices in mateFlag, free until ncc2_best; count in counters[6]) for the warp-per-set launches below.
__global__ void __launch_bounds__(32 * WPB, 8) cluster8_kernel(DevBatch b, DevParams p)
{
    __shared__ double s_x[WPB][4][8], s_y[WPB][4][8], s_t[WPB][4][8];
    __shared__ double s_ox[WPB][4][8], s_oy[WPB][4][8], s_om[WPB][4][8], s_dk[WPB][4][8], s_gk[WPB][4][8];
    __shared__ int s_lab[WPB][4][8], s_csz[WPB][4][8];
    const int f = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane >> 3, gl = lane & 7;
    const int nL = b.nE[2 * f];
    const int* cstart = b.cstart + (size_t)f * b.E;
    int* ccount = b.ccount + (size_t)f * b.E;
    double *c_x = b.c_x + (size_t)f * b.P, *c_y = b.c_y + (size_t)f * b.P, *c_th = b.c_th + (size_t)f * b.P;
    const bool dumps = b.dumps && f == 0;
    int* big = b.mateFlag + (size_t)f * b.E;
    unsigned long long* nbig = b.counters + (size_t)f * 8 + 6;
    const int small = min(8, p.clus_small);
    const double d2max = p.clus_dist * p.clus_dist;
    double *X = s_x[w][g], *Y = s_y[w][g], *T = s_t[w][g], *OX = s_ox[w][g], *OY = s_oy[w][g], *OM = s_om[w][g], *DK = s_dk[w][g], *GK = s_gk[w][g];
    int *LAB = s_lab[w][g], *CSZ = s_csz[w][g];
    for (int i0 = (blockIdx.x * WPB + w) * 4; i0 < nL; i0 += gridDim.x * WPB * 4) {      // warp-uniform trip count
        const int i = i0 + g;
        int n = i < nL ? ccount[i] : 0;
        if (n > small) { if (gl == 0) big[(int)atomicAdd(nbig, 1ull)] = i; n = 0; }      // left for the warp-per-set launches
        else if (n == 0 && i < nL && dumps && gl == 0) b.dump[DUMP_S10].n[i] = 0;
        const int st = n ? cstart[i] : 0;
        EBVO_ASSERT(b.errFlag + f, n >= 0 && n <= 8 && st >= 0 && st + n <= b.P);
        const bool valid = gl < n;
        double xi = 0, yi = 0, ti = 0;
        if (valid) { xi = c_x[st + gl]; yi = c_y[st + gl]; ti = c_th[st + gl]; }        // after the second shift
        X[gl] = xi; Y[gl] = yi; T[gl] = ti; LAB[gl] = gl; CSZ[gl] = 1;
        __syncwarp();
        int li = gl;
        for (;;) {
            int bj = -1;
            if (valid) {
                double best = CUDART_INF;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j >= n || LAB[j] == li) continue;
                    const double dx = xi - X[j], dy = yi - Y[j];
                    const double d2 = dx * dx + dy * dy;
                    if (d2 < best && d2 < d2max && fabs(ti - T[j]) < p.clus_orient_rad) { best = d2; bj = j; }
                }
                if (bj >= 0 && CSZ[li] + CSZ[LAB[bj]] > p.clus_max) bj = -1;             // MAX_CLUSTER_SIZE, EdgeClusterer.cpp:179
            }
            const unsigned m = __ballot_sync(FULL, bj >= 0);
            if (!m) break;
            const unsigned mg = (m >> (8 * g)) & 0xffu;
            const int src = mg ? __ffs(mg) - 1 : 0;
            const int pj = __shfl_sync(FULL, bj, 8 * g + src);
            EBVO_ASSERT(b.errFlag + f, !mg || (pj >= 0 && pj < n && src < n));
            int lnew = 0, lold = -1, merged = 0;
            if (mg) { lnew = LAB[src]; lold = LAB[pj]; merged = CSZ[lnew] + CSZ[lold]; }
            __syncwarp();
            if (mg) {
                if (valid && li == lold) { li = lnew; LAB[gl] = lnew; }
                if (gl == 0) CSZ[lnew] = merged;
            }
            __syncwarp();
        }
        // Gaussian-weighted centre of every cluster (EdgeClusterer.cpp:43-117), sums in member index order
        const bool isL = valid && li == gl;
        if (isL) {
            double sxx = 0, syy = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) if (k < n && LAB[k] == gl) { sxx += X[k]; syy += Y[k]; }
            OX[gl] = sxx / CSZ[gl]; OY[gl] = syy / CSZ[gl];
        }
        __syncwarp();
        double dk = 0;
        if (valid) { const double dx = xi - OX[li], dy = yi - OY[li]; dk = sqrt(dx * dx + dy * dy); DK[gl] = dk; }
        __syncwarp();
        if (isL) {
            double tot = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) if (k < n && LAB[k] == gl) tot += DK[k];
            OM[gl] = tot / CSZ[gl];
        }
        __syncwarp();
        if (valid) { const double z = (dk - OM[li]) / p.clus_sigma; GK[gl] = exp(-0.5 * (z * z)); }   // :103
        __syncwarp();
        double vx = 0, vy = 0, vt = 0;
        if (isL) {
            double wx = 0, wy = 0, wt = 0, ww = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) if (k < n && LAB[k] == gl) { const double gg = GK[k]; wx += gg * X[k]; wy += gg * Y[k]; wt += gg * T[k]; ww += gg; }
            vx = wx / ww; vy = wy / ww; vt = wt / ww;
        }
        // clusters in ascending label order (std::map, :209-222)
        const unsigned ml = (__ballot_sync(FULL, isL) >> (8 * g)) & 0xffu;
        if (isL) {
            const int c = __popc(ml & ((1u << gl) - 1));
            c_x[st + c] = vx; c_y[st + c] = vy; c_th[st + c] = vt;
            if (dumps) dump_put(b.dump[DUMP_S10], st + c, -1, vx, vy, vt, CUDART_NAN);
        }
        if (gl == 0 && n > 0) { ccount[i] = __popc(ml); if (dumps) b.dump[DUMP_S10].n[i] = __popc(ml); }
        __syncwarp();
    }
}

// The warp-per-set launches work through the list cluster8_kernel left behind.  CAP = shared-memory capacity per warp.
// <48> (4.6 KB per warp) takes the sets with n <= min(48, clus_small) and passes the others on in a second list (indices in the
// mates staging area, count in nMates[f]: both free until ncc2_best / compact); <MAXC> takes that list.  Both decide pair
// admissibility once (bit masks) and drop a neighbour from a point's mask for good once the two share a cluster: a rescan
// costs the number of neighbours still in OTHER clusters instead of n.
template <int CAP, int MINB>
__global__ void __launch_bounds__(32 * WPB, MINB) cluster_kernel(DevBatch b, DevParams p)
{
    __shared__ double s_x[WPB][CAP], s_y[WPB][CAP], s_t[WPB][CAP];
    __shared__ double s_ox[WPB][CAP], s_oy[WPB][CAP], s_ot[WPB][CAP];
    __shared__ int s_lab[WPB][CAP], s_csz[WPB][CAP];
    __shared__ double s_dk[WPB][CAP], s_gk[WPB][CAP];
    __shared__ unsigned long long s_adm[WPB][2 * CAP];
    const int f = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int* cstart = b.cstart + (size_t)f * b.E;
    int* ccount = b.ccount + (size_t)f * b.E;
    double *c_x = b.c_x + (size_t)f * b.P, *c_y = b.c_y + (size_t)f * b.P, *c_th = b.c_th + (size_t)f * b.P;
    const bool dumps = b.dumps && f == 0;
    const int* listA = b.mateFlag + (size_t)f * b.E;
    int* listB = reinterpret_cast<int*>(b.mates + (size_t)f * b.E);
    const int* work = CAP < MAXC ? listA : listB;
    const int nwork = CAP < MAXC ? (int)b.counters[(size_t)f * 8 + 6] : b.nMates[f];
    // sets are handed out one at a time from a per-frame cursor: their cost grows faster than n^2, so a static split leaves the
    // kernel waiting for the warp that drew the largest ones
    int* cursor = b.wcur + (size_t)f * 4 + (CAP < MAXC ? 0 : 1);
    for (;;) {
        int wi = 0;
        if (lane == 0) wi = atomicAdd(cursor, 1);
        wi = __shfl_sync(FULL, wi, 0);
        if (wi >= nwork) break;
        const int i = work[wi];
        EBVO_ASSERT(b.errFlag + f, i >= 0 && i < b.nE[2 * f]);
        int n = ccount[i];
        if (CAP < MAXC && (n > CAP || n > p.clus_small)) {          // left for the MAXC launch
            if (lane == 0) listB[atomicAdd(&b.nMates[f], 1)] = i;
            continue;
        }
        if (n > CAP) { if (lane == 0) atomicExch(b.errFlag + f, 4); n = CAP; }
        const int st = cstart[i];
        for (int k = lane; k < n; k += 32) { s_x[w][k] = c_x[st + k]; s_y[w][k] = c_y[st + k]; s_t[w][k] = c_th[st + k]; }   // after the second shift
        __syncwarp();
        const int ncl = warp_cluster(s_x[w], s_y[w], s_t[w], n, true, p, lane, s_lab[w], s_csz[w], s_dk[w], s_gk[w], s_ox[w], s_oy[w], s_ot[w], s_adm[w]);
        __syncwarp();
        for (int k = lane; k < ncl; k += 32) {
            c_x[st + k] = s_ox[w][k]; c_y[st + k] = s_oy[w][k]; c_th[st + k] = s_ot[w][k];
            if (dumps) dump_put(b.dump[DUMP_S10], st + k, -1, s_ox[w][k], s_oy[w][k], s_ot[w][k], CUDART_NAN);
        }
        if (lane == 0) { ccount[i] = ncl; if (dumps) b.dump[DUMP_S10].n[i] = ncl; }
        __syncwarp();
    }
}

// S11 NCC of every cluster centre against the left patches (raw images, :1500) + S12 arg-max (first maximum wins, :941-951).
// A quarter-warp per LEFT EDGE, four left edges per warp in lock step (a left edge has one or two cluster centres, so a
// full warp per edge idles on memory latency and a quarter-warp per centre of one edge finds nothing to do).
__global__ void __launch_bounds__(32 * WPB, 5) ncc2_best_kernel(DevBatch b, DevParams p)
{
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int g = lane >> 3, q = lane & 7, base = lane & ~7;
    const int imgL = 2 * f, imgR = 2 * f + 1;
    const int nL = b.nE[imgL];
    const uint8_t* IR = b.raw + (size_t)imgR * b.imgStride;
    const float* npL = b.npatch + (size_t)imgL * b.E * NPF;
    const uint8_t* pfL = b.pflag + (size_t)imgL * b.E;
    const double *exL = b.ex + (size_t)imgL * b.E, *eyL = b.ey + (size_t)imgL * b.E, *ethL = b.eth + (size_t)imgL * b.E;
    const int* cstart = b.cstart + (size_t)f * b.E;
    int* ccount = b.ccount + (size_t)f * b.E;
    double *c_x = b.c_x + (size_t)f * b.P, *c_y = b.c_y + (size_t)f * b.P, *c_th = b.c_th + (size_t)f * b.P;
    double* c_score = b.c_score + (size_t)f * b.P;
    int* mateFlag = b.mateFlag + (size_t)f * b.E;
    ebvo_mate* mates = b.mates + (size_t)f * b.E;   // staging: slot i
    const bool dumps = b.dumps && f == 0;
    unsigned long long n2 = 0;
    for (int i0 = (blockIdx.x * WPB + (threadIdx.x >> 5)) * 4; i0 < nL; i0 += gridDim.x * WPB * 4) {     // warp-uniform trip count
        const int i = i0 + g;
        const int ncl = i < nL ? ccount[i] : 0;
        if (i < nL && ncl == 0 && q == 0) { mateFlag[i] = 0; if (dumps) b.dump[DUMP_S11].n[i] = 0; }
        const int maxn = __reduce_max_sync(FULL, ncl);
        if (maxn == 0) continue;
        const int st = ncl ? cstart[i] : 0;
        double lp[7], lm[7];      // left patches: the cells this lane owns, as doubles
        bool lP = false, lM = false;
#pragma unroll
        for (int r = 0; r < 7; ++r) { lp[r] = 0.0; lm[r] = 0.0; }
        if (ncl) {
            const float* o = npL + (size_t)i * NPF;
#pragma unroll
            for (int r = 0; r < 7; ++r) if (r < 6 || q == 0) { lp[r] = (double)o[q + 8 * r]; lm[r] = (double)o[52 + q + 8 * r]; }
            const int fl = pfL[i];
            lP = fl & 1; lM = fl & 2;
        }
        double bestS = -1.0, bx = 0, by = 0, bt = 0;
        int bestK = -1, nsurv = 0;
        for (int k = 0; k < maxn; ++k) {
            const bool act = k < ncl;
            double cx = 0, cy = 0, ct = 0, sn = 0, cs = 1;
            if (act) { cx = c_x[st + k]; cy = c_y[st + k]; ct = c_th[st + k]; sincos(ct, &sn, &cs); }
            float vp[7], vm[7];
            bool rP, rM;
            raw_patches8(IR, b.pitch, b.W, b.H, cx, cy, sn, cs, p.shift_mag, q, act, vp, vm);
            normalise_patches8(vp, vm, q, base, rP, rM);
            double pp = 0, nn = 0, pn = 0, np = 0;
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                if (r < 6 || q == 0) {
                    const double rp = (double)vp[r], rm = (double)vm[r];
                    pp = fma(lp[r], rp, pp); nn = fma(lm[r], rm, nn); pn = fma(lp[r], rm, pn); np = fma(lm[r], rp, np);
                }
            }
            group8_sum4(pp, nn, pn, np, q, base);
            const double sc = ncc_max4(pp, nn, pn, np, lP, lM, rP, rM);
            if (act) {
                ++n2;
                if (sc > p.ncc_thresh) {
                    if (dumps && q == 0) dump_put(b.dump[DUMP_S11], st + nsurv, -1, cx, cy, ct, sc);
                    ++nsurv;
                    if (sc > bestS) { bestS = sc; bestK = k; bx = cx; by = cy; bt = ct; }
                }
            }
        }
        __syncwarp();
        if (q == 0 && ncl) {
            if (dumps) b.dump[DUMP_S11].n[i] = nsurv;
            if (bestK >= 0) {
                ebvo_mate m;
                m.left_index = i; m.reserved = 0;
                m.lx = exL[i]; m.ly = eyL[i]; m.ltheta = ethL[i];
                m.rx = bx; m.ry = by; m.rtheta = bt;
                m.score = bestS;
                mates[i] = m;
                mateFlag[i] = 1;
                c_x[st] = bx; c_y[st] = by; c_th[st] = bt; c_score[st] = bestS;
                ccount[i] = 1;
            } else { mateFlag[i] = 0; ccount[i] = 0; }
        }
        __syncwarp();
    }
    if (q == 0 && n2) atomicAdd(&b.counters[(size_t)f * 8 + 4], n2);
}

// S13: ordered compaction of the per-left-edge staging slots into the mate list (one CTA per frame)
__global__ void __launch_bounds__(1024) compact_kernel(DevBatch b, ebvo_mate* out, int outStride)
{
    // one CTA per frame walks the left-edge list in chunks of 1024 (coalesced flag reads and record copies): ballot + popc
    // inside a warp, the 32 warp totals scanned through shared memory, a running base across chunks
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nL = b.nE[2 * f];
    const int* flag = b.mateFlag + (size_t)f * b.E;
    const uint4* stg = reinterpret_cast<const uint4*>(b.mates + (size_t)f * b.E);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)f * outStride);
    __shared__ int s_w[32];
    __shared__ int s_base;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < nL; c0 += 1024) {
        const int i = c0 + tid;
        const bool keep = i < nL && flag[i];
        const unsigned m = __ballot_sync(FULL, keep);
        if (lane == 0) s_w[w] = __popc(m);
        __syncthreads();
        if (w == 0) {
            const int v = s_w[lane];
            int x = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, x, d); if (lane >= d) x += t; }
            s_w[lane] = x - v;                                  // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int base = s_base;
        const int pos = base + s_w[w] + __popc(m & ((1u << lane) - 1));
        if (keep && pos < outStride) {
            const uint4* r = stg + 4 * (size_t)i;               // 64-byte records
            uint4* o = dst + 4 * (size_t)pos;
            o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
        }
        __syncthreads();
        if (tid == 1023) s_base = base + s_w[31] + __popc(m);   // total of this chunk (warp 31's prefix + its count)
        __syncthreads();
    }
    if (tid == 0) b.nMates[f] = s_base < outStride ? s_base : outStride;
}

// debug: copy frame 0's live pool (src < 0) or dump[src] into compact arrays at host-scanned offsets
__global__ void snapshot_kernel(DevBatch b, int src, const int* offsets, int* ridx, double* x, double* y, double* th, double* score)
{
    const int nL = b.nE[0];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nL) return;
    const int st = b.cstart[i];
    const int n = offsets[i + 1] - offsets[i];
    for (int k = 0; k < n; ++k) {
        const int o = offsets[i] + k;
        if (src < 0) { ridx[o] = b.c_ridx[st + k]; x[o] = b.c_x[st + k]; y[o] = b.c_y[st + k]; th[o] = b.c_th[st + k]; score[o] = b.c_score[st + k]; }
        else { const DumpBuf& d = b.dump[src]; ridx[o] = d.ridx[st + k]; x[o] = d.x[st + k]; y[o] = d.y[st + k]; th[o] = d.th[st + k]; score[o] = d.score[st + k]; }
    }
}

// ------------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------------
static int g_sms = 0;     // SM count (every GPU of a box is the same model); set by init_match_device, i.e. before any launch

void init_match_device()   // per-device function attributes, called by ebvo_create after cudaSetDevice
{
    cudaFuncSetAttribute(gn_lerp64_kernel<GNL_PX, GNL_MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms > 0) g_sms = sms;
}

static dim3 warp_grid(int nFrames)
{
    if (!g_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev); }
    int gx = (g_sms * 16 + nFrames - 1) / nFrames;   // about 16 CTAs of 4 warps per SM over the whole batch
    if (gx < 1) gx = 1;
    return dim3(gx, nFrames);
}

void launch_sobel(const DevBatch& b, int nFrames, cudaStream_t st, Prof* prof)
{
    dim3 g((b.W + 31) / 32, (b.H + 7) / 8, nFrames), t(32, 8);
    EBVO_KERNEL(prof, "sobel", st, (sobel_kernel<<<g, t, 0, st>>>(b)));
}


void match_prologue(const DevBatch& b, const DevParams& p, const double* F21, int nFrames, cudaStream_t st, Prof* prof)
{
    (void)p;
    cudaMemsetAsync(b.poolUsed, 0, sizeof(int) * nFrames, st);
    cudaMemsetAsync(b.counters, 0, sizeof(unsigned long long) * 8 * nFrames, st);
    cudaMemsetAsync(b.nMates, 0, sizeof(int) * nFrames, st);      // doubles as the count of the clusterer's second work list until compact writes it
    cudaMemsetAsync(b.wcur, 0, sizeof(int) * 4 * nFrames, st);
    launch_sobel(b, nFrames, st, prof);
    EBVO_KERNEL(prof, "bounds", st, (bounds_kernel<<<nFrames, 1024, 0, st>>>(b)));
}
void match_gate(const DevBatch& b, const DevParams& p, const double* F21, int nFrames, cudaStream_t st, Prof* prof)
{
    EBVO_KERNEL(prof, "gate", st, (gate_kernel<<<warp_grid(nFrames), 32 * WPB, 0, st>>>(b, p, fmat_of(F21))));
}
void match_sift(const DevBatch& b, const DevParams& p, int nFrames, cudaStream_t st, Prof* prof)
{
    if (b.siftDev) EBVO_KERNEL(prof, "sift_gate", st, (sift_gate8_kernel<<<warp_grid(nFrames), 32 * WPB, 0, st>>>(b, p)));
    else EBVO_KERNEL(prof, "sift_gate", st, (sift_gate_kernel<<<warp_grid(nFrames), 32 * WPB, 0, st>>>(b, p)));
}
void match_ncc(const DevBatch& b, const DevParams& p, int nFrames, bool sift, cudaStream_t st, Prof* prof)
{
    // (launching the two kernels per group of 1 / 2 / 4 / 8 frames, so that a group's patch rows are still in the L2 when the NCC kernel
    // reads them back, was measured: 18.9 / 14.2 / 13.0 / 12.2 ms per 160-frame step against 11.4 for one launch each - the DRAM
    // round trip of the rows is not what the NCC kernel waits for)
    dim3 gp(warp_grid(nFrames).x, 2 * nFrames);
    EBVO_KERNEL(prof, "patch", st, (patch_kernel<<<gp, 32 * WPB, 0, st>>>(b, p, 0)));
    EBVO_KERNEL(prof, "ncc_bnb", st, (ncc_bnb_kernel<<<warp_grid(nFrames), 32 * WPB, 0, st>>>(b, p, sift ? 1 : 0, 0)));
}
static dim3 slot_grid(int nFrames)
{
    dim3 g = warp_grid(nFrames);
    g.x = (g.x + 3) / 4;
    return g;
}
void match_gn(const DevBatch& b, const DevParams& p, int nFrames, cudaStream_t st, Prof* prof)
{
    EBVO_KERNEL(prof, "shift", st, (shift_kernel<<<slot_grid(nFrames), 128, 0, st>>>(b, p, 0)));
    if (p.gn_mode == 2) EBVO_KERNEL(prof, "gn32", st, (gn32_kernel<<<warp_grid(nFrames), 32 * WPB, 0, st>>>(b, p)));
    else if (p.gn_mode == 1) EBVO_KERNEL(prof, "gn64", st, (gn64_kernel<<<warp_grid(nFrames), 32 * WPB, 0, st>>>(b, p)));
    else {
        if (!g_sms) warp_grid(1);
        EBVO_KERNEL(prof, "gn", st, (gn_lerp64_kernel<GNL_PX, GNL_MINB><<<g_sms * GNL_MINB, 32 * WPB, 0, st>>>(b, p, GN_R, nFrames)));
    }
}
void match_cluster(const DevBatch& b, const DevParams& p, int nFrames, cudaStream_t st, Prof* prof)
{
    EBVO_KERNEL(prof, "shift", st, (shift_kernel<<<slot_grid(nFrames), 128, 0, st>>>(b, p, 1)));
    constexpr int CL_SMALL = 48;
    EBVO_KERNEL(prof, "cluster8", st, (cluster8_kernel<<<warp_grid(nFrames), 32 * WPB, 0, st>>>(b, p)));
    EBVO_KERNEL(prof, "cluster", st, (cluster_kernel<CL_SMALL, 8><<<warp_grid(nFrames), 32 * WPB, 0, st>>>(b, p)));
    EBVO_KERNEL(prof, "cluster_big", st, (cluster_kernel<MAXC, 4><<<dim3(16, nFrames), 32 * WPB, 0, st>>>(b, p)));
    EBVO_KERNEL(prof, "ncc2_best", st, (ncc2_best_kernel<<<warp_grid(nFrames), 32 * WPB, 0, st>>>(b, p)));
}

void launch_match(const DevBatch& b, const DevParams& p, const double* F21, int nFrames, bool sift, cudaStream_t st, Prof* prof)
{
    match_prologue(b, p, F21, nFrames, st, prof);
    if (sift && b.siftDev) launch_sift(b, 2 * nFrames, st, prof);
    match_gate(b, p, F21, nFrames, st, prof);
    if (sift) match_sift(b, p, nFrames, st, prof);
    match_ncc(b, p, nFrames, sift, st, prof);
    match_gn(b, p, nFrames, st, prof);
    match_cluster(b, p, nFrames, st, prof);
}

void launch_compact(const DevBatch& b, int nFrames, ebvo_mate* d_out, int stride, cudaStream_t st, Prof* prof)
{
    EBVO_KERNEL(prof, "compact", st, (compact_kernel<<<nFrames, 1024, 0, st>>>(b, d_out, stride)));
}

// Unpadded result of a batch: the mates of frames 0 .. nFrames-1 back to back in `dst` (64-byte records), offsets[f] = first
// record of frame f, offsets[nFrames] = total.  One CTA per frame; every CTA sums the counts before its own frame (<= a few
// hundred ints).  This is what travels in the final gather of a sharded batch (sharding.gather_packed).
__global__ void __launch_bounds__(256) pack_kernel(const ebvo_mate* src, int srcStride, const int* nMates, int nFrames, ebvo_mate* dst,
                                                    long long cap, int* offsets)
{
    __shared__ int s_off;
    const int f = blockIdx.x, tid = threadIdx.x;
    if (tid < 32) {
        int acc = 0;
        for (int g = tid; g < f; g += 32) acc += nMates[g];
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
        if (tid == 0) s_off = acc;
    }
    __syncthreads();
    const int off = s_off, n = nMates[f];
    if (tid == 0) { offsets[f] = off; if (f == nFrames - 1) offsets[nFrames] = off + n; }
    const uint4* r = reinterpret_cast<const uint4*>(src + (size_t)f * srcStride);
    uint4* o = reinterpret_cast<uint4*>(dst + off);
    const long long room = cap - off;                          // records that still fit
    const int m = (int)max(0ll, min((long long)n, room));
    for (int k = tid; k < 4 * m; k += 256) o[k] = r[k];
}
void launch_pack(const ebvo_mate* src, int srcStride, const int* nMates, int nFrames, ebvo_mate* dst, long long cap, int* offsets, cudaStream_t st, Prof* prof)
{
    if (nFrames > 0) EBVO_KERNEL(prof, "pack", st, (pack_kernel<<<nFrames, 256, 0, st>>>(src, srcStride, nMates, nFrames, dst, cap, offsets)));
}

void launch_gate_count(const DevBatch& b, const DevParams& p, const double* F21, int mode, int* d_counts, cudaStream_t st)
{
    gate_count_kernel<<<592, 32 * WPB, 0, st>>>(b, p, fmat_of(F21), mode, d_counts);
}
void launch_gate_fill(const DevBatch& b, const DevParams& p, const double* F21, int mode, const int* d_offsets, int* d_ridx, cudaStream_t st)
{
    gate_fill_kernel<<<592, 32 * WPB, 0, st>>>(b, p, fmat_of(F21), mode, d_offsets, d_ridx);
}
void launch_snapshot(const DevBatch& b, int src, const int* d_offsets, int* ridx, double* x, double* y, double* th, double* score, cudaStream_t st)
{
    snapshot_kernel<<<(b.E + 127) / 128, 128, 0, st>>>(b, src, d_offsets, ridx, x, y, th, score);
}

// ---- small single-purpose entry points ----------------------------------------------------------------
__global__ void edge_patches_kernel(const uint8_t* I, int w, int h, int pitch, const double* ex, const double* ey, const double* eth, int n,
                                    double shift, float* plus, float* minus)
{
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= n) return;
    float vp[2], vm[2];
    raw_patches(I, pitch, w, h, ex[e], ey[e], eth[e], shift, lane, vp, vm);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int t = lane + 32 * u;
        if (t < 49) { plus[(size_t)e * 49 + t] = vp[u]; minus[(size_t)e * 49 + t] = vm[u]; }
    }
}
void launch_edge_patches(const uint8_t* d_img, int w, int h, int pitch, const double* ex, const double* ey, const double* eth, int n,
                         double shift, float* plus, float* minus, cudaStream_t st)
{
    if (n > 0) edge_patches_kernel<<<(n + 3) / 4, 128, 0, st>>>(d_img, w, h, pitch, ex, ey, eth, n, shift, plus, minus);
}

// finalisation helpers of ebvo_stereo_match_full: coordinate arrays of the mates, descriptor gathers
__global__ void mates_to_edges_kernel(const ebvo_mate* m, int n, double* lx, double* ly, double* lt, double* rx, double* ry, double* rt)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const ebvo_mate a = m[k];
    lx[k] = a.lx; ly[k] = a.ly; lt[k] = a.ltheta; rx[k] = a.rx; ry[k] = a.ry; rt[k] = a.rtheta;
}
void launch_mates_to_edges(const ebvo_mate* m, int n, double* lx, double* ly, double* lt, double* rx, double* ry, double* rt, cudaStream_t st)
{
    if (n > 0) mates_to_edges_kernel<<<(n + 127) / 128, 128, 0, st>>>(m, n, lx, ly, lt, rx, ry, rt);
}
__global__ void gather_desc_kernel(const uint8_t* desc8, const ebvo_mate* m, int n, float* out)   // one CTA of 256 threads per mate
{
    const int k = blockIdx.x;
    if (k >= n) return;
    out[(size_t)k * 256 + threadIdx.x] = (float)desc8[(size_t)(m ? m[k].left_index : k) * 256 + threadIdx.x];
}
void launch_gather_desc(const uint8_t* desc8, const ebvo_mate* m, int n, float* out, cudaStream_t st)
{
    if (n > 0) gather_desc_kernel<<<n, 256, 0, st>>>(desc8, m, n, out);
}
void launch_desc_to_float(const uint8_t* desc8, int n, float* out, cudaStream_t st)
{
    if (n > 0) gather_desc_kernel<<<n, 256, 0, st>>>(desc8, nullptr, n, out);
}

__global__ void ncc_pairs_kernel(const float* p1, const float* p2, int n, double* out)
{
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= n) return;
    float a[2], z[2] = {0.f, 0.f}, c[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int t = lane + 32 * u;
        a[u] = t < 49 ? p1[(size_t)e * 49 + t] : 0.f;
        c[u] = t < 49 ? p2[(size_t)e * 49 + t] : 0.f;
    }
    Patches A, B;
    normalise_patches(a, z, lane, A);
    normalise_patches(c, z, lane, B);
    double pp = (double)A.p[0] * (double)B.p[0] + (double)A.p[1] * (double)B.p[1];
    pp = warp_sum(pp);
    if (A.flatP || B.flatP) pp = -1.0;
    if (lane == 0) out[e] = pp;
}
void launch_ncc_pairs(const float* p1, const float* p2, int n, double* out, cudaStream_t st)
{
    if (n > 0) ncc_pairs_kernel<<<(n + 3) / 4, 128, 0, st>>>(p1, p2, n, out);
}

__global__ void cluster_one_kernel(const double* x, const double* y, const double* th, int n, int by_orient, DevParams p,
                                   double* cx, double* cy, double* cth, int* labels, int* nclusters)
{
    __shared__ double s_x[MAXC], s_y[MAXC], s_t[MAXC], s_ox[MAXC], s_oy[MAXC], s_ot[MAXC];
    __shared__ int s_lab[MAXC], s_csz[MAXC];
    __shared__ double s_dk[MAXC], s_gk[MAXC];
    const int lane = threadIdx.x;
    for (int k = lane; k < n; k += 32) { s_x[k] = x[k]; s_y[k] = y[k]; s_t[k] = th[k]; }
    __syncwarp();
    const int ncl = warp_cluster(s_x, s_y, s_t, n, by_orient != 0, p, lane, s_lab, s_csz, s_dk, s_gk, s_ox, s_oy, s_ot);
    __syncwarp();
    for (int k = lane; k < ncl; k += 32) { cx[k] = s_ox[k]; cy[k] = s_oy[k]; cth[k] = s_ot[k]; }
    for (int k = lane; k < n; k += 32) labels[k] = s_lab[k];
    if (lane == 0) *nclusters = ncl;
}
void launch_cluster_one(const double* x, const double* y, const double* th, int n, int by_orient, const DevParams& p,
                        double* cx, double* cy, double* cth, int* labels, int* nclusters, cudaStream_t st)
{
    cluster_one_kernel<<<1, 32, 0, st>>>(x, y, th, n, by_orient, p, cx, cy, cth, labels, nclusters);
}

#include "temporal.inl"

}  // namespace ebvo
