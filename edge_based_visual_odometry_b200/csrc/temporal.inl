// Keyframe -> current-frame quad tracking on the device: the stages of
// Temporal_Matches::get_Temporal_Edge_Pairs_from_Quads (/root/reference/src/Temporal_Matches.cpp:168-218) over the
// finalised stereo mates of two frames.  Included at the end of match.cu (same translation unit: it reuses the patch,
// NCC, best-nearly-best, FP64 sampler and clusterer device code of the stereo stage).
//
//   tq_cells / tq_scan / tq_fill / tq_sort   SpatialGrid of the CF mates' left edges as a CSR (cell 15 px), ascending
//                                            CF index inside a cell = the reference's push_back order (:18-55)
//   tq_patch          normalised "+"/"-" patches of the KF and CF mates (left: raw left image, right: undistorted
//                     right image), once per mate
//   tq_gate           one warp per KF mate: candidates in the reference's order (5x5 cell block, cells row-major),
//                     right-cell membership (= the unordered_set test of :353-358), orientation gate on both views
//                     (:385-414), NCC > 0.8 on both views (:416-469), best-nearly-best on the left score (:517-570)
//   tq_gn             2-D Gauss-Newton on both views (:572-634, :735-851), one warp per KF mate looping over its quads:
//                     the keyframe patches are sampled once per mate and side; FP64 blends rounded to float, FP64 normal
//                     equations, Eigen's pivoted 2x2 LDLT restated
//   tq_cluster        EdgeClusterer by orientation on the refined left edges + the right-centre means (:636-733)
//   tq_gather         ordered compaction into ebvo_quad records
// The SIFT gate (:471-515) and the SIFT best-nearly-best pass run inside tq_gate on caller-supplied descriptor pairs
// (the mates' left / right_edge_descriptors); without them both are skipped ("SIFT-off", as the stereo stage's default).

static_assert(TQ_CAP <= MAXC, "the warp-private lists of the quad kernels reuse the stereo stage's capacity");

__device__ __forceinline__ bool tq_orient_ok(double a, double b, double thr)   // Temporal_Matches.cpp:394-404
{
    double d = fabs((a - b) * (180.0 / M_PI));
    if (d > 180.0) d = 360.0 - d;
    return d < thr || fabs(d - 180.0) < thr;
}

__global__ void tq_cells_kernel(TqDev d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n_cf) return;
    const double* m = d.cf + 6 * (size_t)i;
    const int lx = (int)m[0] / d.cell, ly = (int)m[1] / d.cell;      // static_cast<int>(x) / cell_size (:30-31)
    int lc = -1;
    if (lx >= 0 && lx < d.gw && ly >= 0 && ly < d.gh) { lc = ly * d.gw + lx; atomicAdd(&d.cellCount[lc], 1); }
    d.lcell[i] = lc;
    const int rx = (int)m[3] / d.cell, ry = (int)m[4] / d.cell;
    const bool ok = rx >= 0 && rx < d.gw && ry >= 0 && ry < d.gh;
    d.rcx[i] = ok ? rx : -1000000; d.rcy[i] = ok ? ry : -1000000;
}
// exclusive scan of n ints by one block (n up to a few 10^4 cells / keyframe mates)
__global__ void __launch_bounds__(1024) tq_scan_kernel(const int* in, int* out, int n, int* total)
{
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? in[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, x, o); if (lane >= o) x += t; }
        if (lane == 31) s_w[w] = x;
        __syncthreads();
        if (w == 0) {
            int y = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, y, o); if (lane >= o) y += t; }
            s_w[lane] = y;
        }
        __syncthreads();
        const int excl = s_carry + (w ? s_w[w - 1] : 0) + x - v;
        if (i < n) out[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[n] = s_carry; if (total) *total = s_carry; }
}
__global__ void tq_fill_kernel(TqDev d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n_cf) return;
    const int c = d.lcell[i];
    if (c >= 0) d.cellList[d.cellStart[c] + atomicAdd(&d.cellCursor[c], 1)] = i;
}
__global__ void tq_sort_kernel(TqDev d)    // ascending CF index inside every cell (the reference pushes mates in index order)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d.gw * d.gh) return;
    int* a = d.cellList + d.cellStart[c];
    const int n = d.cellCount[c];
    for (int i = 1; i < n; ++i) {
        const int v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
        a[j + 1] = v;
    }
}

// {I, 8 gx, 8 gy} as int16 (exact), Sobel 3x3 / 8 with BORDER_REFLECT_101 (utility.h:131-141), for one image
__global__ void tq_pack_kernel(const uint8_t* I, int W, int H, int pitch, uint2* out, uint2* outh)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    auto R = [](int i, int n) { return n == 1 ? 0 : (i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i)); };
    const int xm = R(x - 1, W), xp = R(x + 1, W), ym = R(y - 1, H), yp = R(y + 1, H);
    auto at = [&](int yy, int xx) { return (int)I[(size_t)yy * pitch + xx]; };
    const int gx8 = (at(ym, xp) - at(ym, xm)) + 2 * (at(y, xp) - at(y, xm)) + (at(yp, xp) - at(yp, xm));
    const int gy8 = (at(yp, xm) - at(ym, xm)) + 2 * (at(yp, x) - at(ym, x)) + (at(yp, xp) - at(ym, xp));
    out[(size_t)y * W + x] = make_uint2((uint32_t)at(y, x) | ((uint32_t)(gx8 & 0xffff) << 16), (uint32_t)(gy8 & 0xffff));
    // 8-bit intensities and Sobel/8 values (k/8, |k| <= 1020) are exact in fp16
    const __half2 hg = __floats2half2_rn((float)gx8 * 0.125f, (float)gy8 * 0.125f);
    outh[(size_t)y * W + x] = make_uint2((uint32_t)__half_as_ushort(__float2half_rn((float)at(y, x))), *reinterpret_cast<const uint32_t*>(&hg));
}

__global__ void __launch_bounds__(32 * WPB) tq_patch_kernel(TqDev d, DevParams p)
{
    const int set = blockIdx.y, lane = threadIdx.x & 31;
    const bool isKf = set < 2, right = set & 1;
    const int n = isKf ? d.n_kf : d.n_cf;
    const double* m = isKf ? d.kf : d.cf;
    const uint8_t* I = isKf ? (right ? d.kfRund : d.kfLraw) : (right ? d.cfRund : d.cfLraw);
    for (int e = blockIdx.x * WPB + (threadIdx.x >> 5); e < n; e += gridDim.x * WPB) {
        float vp[2], vm[2];
        Patches P;
        const double* q = m + 6 * (size_t)e + (right ? 3 : 0);
        raw_patches(I, d.pitch, d.W, d.H, q[0], q[1], q[2], p.shift_mag, lane, vp, vm);
        normalise_patches(vp, vm, lane, P);
        float* o = d.np[set] + (size_t)e * 98;
        o[lane] = P.p[0]; o[49 + lane] = P.m[0];
        if (lane + 32 < 49) { o[lane + 32] = P.p[1]; o[49 + lane + 32] = P.m[1]; }
        if (lane == 0) d.pf[set][e] = (P.flatP ? 1 : 0) | (P.flatM ? 2 : 0);
    }
}

// mode 0: grid stage only, 1: + orientation, 2: + NCC gate, 3: + SIFT gate, 4: + best-nearly-best on NCC, 5: + on SIFT.
// Modes 0 / 1 serve the stage dumps: they count (offs == nullptr) or write the CF indices at offs[i] (two passes);
// modes >= 2 fill pool 1.
__global__ void __launch_bounds__(32 * WPB) tq_gate_kernel(TqDev d, int mode, int* counts, const int* offs, int* outCf)
{
    __shared__ int s_cf[WPB][TQ_CAP];
    __shared__ double s_nl[WPB][TQ_CAP], s_nr[WPB][TQ_CAP];
    __shared__ double s_sl[WPB][TQ_CAP], s_sr[WPB][TQ_CAP], s_tmp[WPB][TQ_CAP];
    __shared__ int s_or[WPB][TQ_CAP], s_or2[WPB][TQ_CAP], s_or3[WPB][TQ_CAP];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool sift = d.desc[0] != nullptr && mode >= 3;
    unsigned long long nGrid = 0, nOri = 0, nKeep = 0;
    for (int i = blockIdx.x * WPB + w; i < d.n_kf; i += gridDim.x * WPB) {
        if (d.kf_mask && !d.kf_mask[i]) { if (lane == 0) { if (mode < 2) { if (!offs) counts[i] = 0; } else d.cnt[i] = 0; } continue; }
        const double* k = d.kf + 6 * (size_t)i;
        const double thL = k[2], thR = k[5];
        const int gx0 = (int)k[0] / d.cell, gy0 = (int)k[1] / d.cell;      // Dataset.h:94-95
        const int rx0 = (int)k[3] / d.cell, ry0 = (int)k[4] / d.cell;
        Patches KL, KR;
        if (mode >= 2) { load_patches(d.np[0], d.pf[0], i, lane, KL); load_patches(d.np[1], d.pf[1], i, lane, KR); }
        float4 kl1 = make_float4(0, 0, 0, 0), kl2 = kl1, kr1 = kl1, kr2 = kl1;      // keyframe descriptor pairs, 4 floats per lane
        if (sift) {
            const float4* a = reinterpret_cast<const float4*>(d.desc[0] + (size_t)i * 256); kl1 = a[lane]; kl2 = a[32 + lane];
            const float4* b = reinterpret_cast<const float4*>(d.desc[1] + (size_t)i * 256); kr1 = b[lane]; kr2 = b[32 + lane];
        }
        int n = 0;                       // entries produced so far (warp-uniform)
        int ns = 0;
        const int obase = (mode < 2 && offs) ? offs[i] : 0;
        for (int dy = -d.sr; dy <= d.sr; ++dy) {
            const int cy = gy0 + dy;
            if (cy < 0 || cy >= d.gh) continue;
            for (int dx = -d.sr; dx <= d.sr; ++dx) {
                const int cx = gx0 + dx;
                if (cx < 0 || cx >= d.gw) continue;
                const int c = cy * d.gw + cx, st = d.cellStart[c], len = d.cellCount[c];
                for (int b0 = 0; b0 < len; b0 += 32) {
                    const int e = b0 + lane;
                    int cf = -1;
                    bool ok = false;
                    if (e < len) {
                        cf = d.cellList[st + e];
                        // member of the right grid's block around the KF right edge (right_set.count, :353-358)
                        ok = abs(d.rcx[cf] - rx0) <= d.sr && abs(d.rcy[cf] - ry0) <= d.sr;
                    }
                    nGrid += __popc(__ballot_sync(FULL, ok));
                    if (mode >= 1 && ok) {
                        const double* m = d.cf + 6 * (size_t)cf;
                        ok = tq_orient_ok(thL, m[2], d.orient_deg) && tq_orient_ok(thR, m[5], d.orient_deg);
                    }
                    const unsigned bal = __ballot_sync(FULL, ok);
                    if (mode >= 1) nOri += __popc(bal);
                    if (mode < 2) {
                        if (offs && ok) outCf[obase + n + __popc(bal & ((1u << lane) - 1))] = cf;
                        n += __popc(bal);
                        continue;
                    }
                    unsigned rem = bal;
                    while (rem) {                      // survivors of this chunk, in order
                        const int src = __ffs(rem) - 1;
                        rem &= rem - 1;
                        const int c2 = __shfl_sync(FULL, cf, src);
                        Patches CL;
                        load_patches(d.np[2], d.pf[2], c2, lane, CL);
                        const double sl = ncc_score(KL, CL);
                        if (!(sl > d.ncc_thresh)) continue;
                        Patches CR;
                        load_patches(d.np[3], d.pf[3], c2, lane, CR);
                        const double sr_ = ncc_score(KR, CR);
                        if (!(sr_ > d.ncc_thresh)) continue;
                        double fl = 900.0, fr = 900.0;          // scores{-1.0, 900.0} (:364)
                        if (sift) {
                            // min of the four L2 distances on both views (:481-489), gate < sift_thresh (:499)
                            auto d2 = [](float4 u, float4 v) {
                                const double x = (double)u.x - (double)v.x, y = (double)u.y - (double)v.y, z = (double)u.z - (double)v.z, q = (double)u.w - (double)v.w;
                                return x * x + y * y + z * z + q * q;
                            };
                            const float4* cl = reinterpret_cast<const float4*>(d.desc[2] + (size_t)c2 * 256);
                            const float4 b1 = cl[lane], b2 = cl[32 + lane];
                            double d11 = d2(kl1, b1), d12 = d2(kl1, b2), d21 = d2(kl2, b1), d22 = d2(kl2, b2);
                            warp_sum4(d11, d12, d21, d22, lane);
                            fl = fmin(fmin(sqrt(d11), sqrt(d12)), fmin(sqrt(d21), sqrt(d22)));
                            if (!(fl < d.sift_thresh)) continue;
                            const float4* cr = reinterpret_cast<const float4*>(d.desc[3] + (size_t)c2 * 256);
                            const float4 c1 = cr[lane], c2_ = cr[32 + lane];
                            d11 = d2(kr1, c1); d12 = d2(kr1, c2_); d21 = d2(kr2, c1); d22 = d2(kr2, c2_);
                            warp_sum4(d11, d12, d21, d22, lane);
                            fr = fmin(fmin(sqrt(d11), sqrt(d12)), fmin(sqrt(d21), sqrt(d22)));
                            if (!(fr < d.sift_thresh)) continue;
                        }
                        if (ns < TQ_CAP) { if (lane == 0) { s_cf[w][ns] = c2; s_nl[w][ns] = sl; s_nr[w][ns] = sr_; s_sl[w][ns] = fl; s_sr[w][ns] = fr; } ++ns; }
                        else if (lane == 0) atomicExch(d.errFlag, 5);
                    }
                }
            }
        }
        if (mode < 2) { if (!offs && lane == 0) counts[i] = n; continue; }
        __syncwarp();
        int keep = ns;
        if (mode >= 4) keep = bnb_select(s_nl[w], ns, d.bnb_thresh, true, lane, s_or[w], true);
        else for (int q = lane; q < ns; q += 32) s_or[w][q] = q;
        __syncwarp();
        if (mode >= 5 && sift && keep >= 2) {      // second pass on the left SIFT distance of the survivors, in their new order (:196)
            for (int q = lane; q < keep; q += 32) s_tmp[w][q] = s_sl[w][s_or[w][q]];
            __syncwarp();
            const int keep2 = bnb_select(s_tmp[w], keep, d.bnb_thresh, false, lane, s_or2[w], true);
            for (int q = lane; q < keep2; q += 32) s_or3[w][q] = s_or[w][s_or2[w][q]];
            __syncwarp();
            for (int q = lane; q < keep2; q += 32) s_or[w][q] = s_or3[w][q];
            keep = keep2;
            __syncwarp();
        }
        for (int q = lane; q < keep; q += 32) {
            const int o = s_or[w][q], c2 = s_cf[w][o];
            const size_t e = (size_t)i * TQ_CAP + q;
            const double* m = d.cf + 6 * (size_t)c2;
            d.q_cf[e] = c2; d.q_ncc[2 * e] = s_nl[w][o]; d.q_ncc[2 * e + 1] = s_nr[w][o];
            d.q_sift[2 * e] = s_sl[w][o]; d.q_sift[2 * e + 1] = s_sr[w][o];
            d.q_l[3 * e] = m[0]; d.q_l[3 * e + 1] = m[1]; d.q_l[3 * e + 2] = m[2];
            d.q_r[3 * e] = m[3]; d.q_r[3 * e + 1] = m[4]; d.q_r[3 * e + 2] = m[5];
            d.q_sc[2 * e] = 1e6; d.q_sc[2 * e + 1] = 1e6; d.q_valid[e] = 0;      // Dataset.h:325-326
        }
        if (lane == 0) d.cnt[i] = keep;
        nKeep += keep;
        __syncwarp();
    }
    if (lane == 0) {
        if (nGrid) atomicAdd(&d.counters[3], nGrid);
        if (nOri) atomicAdd(&d.counters[4], nOri);
        if (nKeep) atomicAdd(&d.counters[0], nKeep);
    }
}

// Eigen::LDLT<Matrix2d> compute + solve (lower triangle, pivot on the largest |diagonal|, first maximum): H x = b
__device__ __forceinline__ void ldlt2_solve(double h00, double h10, double h11, double b0, double b1, double& x0, double& x1)
{
    const bool swp = fabs(h11) > fabs(h00);
    const double d0 = swp ? h11 : h00, a11 = swp ? h00 : h11;
    double l = h10, d1 = a11;
    double y0 = swp ? b1 : b0, y1 = swp ? b0 : b1;
    if (fabs(d0) > 0.0) { l = l / d0; d1 = a11 - l * (d0 * l); }
    y1 -= l * y0;
    const double tol = 2.2250738585072014e-308;
    y0 = fabs(d0) > tol ? y0 / d0 : 0.0;
    y1 = fabs(d1) > tol ? y1 / d1 : 0.0;
    y0 -= l * y1;
    x0 = swp ? y1 : y0; x1 = swp ? y0 : y1;
}

// Temporal_Matches.cpp:572-634 + :735-851.  Lane owns samples s = lane + 32 m (s < 49: "+" patch, else "-").
__global__ void __launch_bounds__(32 * WPB, 4) tq_gn_kernel(TqDev d, DevParams p)
{
    const int lane = threadIdx.x & 31;
    const int W = d.W, H = d.H;
    const double side = 7 / 2.0 + 1.0;
    unsigned long long nprob = 0, niter = 0;
    // persistent warps: (keyframe mate, view) items are pulled from a cursor, so a mate with 100 quads does not hold up the
    // warps that drew mates with two (counters[5 + view] = cursor)
    unsigned long long* cursor = d.counters + 5 + blockIdx.y;
    for (;;) {
        int i = 0;
        if (lane == 0) i = (int)atomicAdd(cursor, 1ull);
        i = __shfl_sync(FULL, i, 0);
        if (i >= d.n_kf) break;
        const int n = d.cnt[i];
        if (n == 0) continue;
        {
            const int sd = blockIdx.y;      // left and right views are refined by different warps
            const double* k = d.kf + 6 * (size_t)i + 3 * sd;
            const double kx = k[0], ky = k[1], kth = k[2];
            const uint8_t* Ikf = sd ? d.kfRund : d.kfLund;
            const uint2* __restrict__ PK = d.pk16[sd];
            // keyframe patches, once per mate and side (:745-768)
            double Lc[4];
            {
                double st_, ct_;
                sincos(kth, &st_, &ct_);
                const double nxs = -st_ * side, nys = ct_ * side;
                double sumP = 0, sumM = 0;
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int s = lane + 32 * m;
                    const bool neg = s >= 49;
                    const int t = s - (neg ? 49 : 0);
                    const int ii = t / 7 - 3, jj = t % 7 - 3;
                    Lc[m] = 0.0;
                    if (s < 98) {
                        const double cx = neg ? kx - nxs : kx + nxs, cy = neg ? ky - nys : ky + nys;
                        Lc[m] = sample_u8_exact(Ikf, d.pitch, W, H, cx + ct_ * ii - st_ * jj, cy + st_ * ii + ct_ * jj);
                        if (neg) sumM += Lc[m]; else sumP += Lc[m];
                    }
                }
                warp_sum2(sumP, sumM);
                const double mLp = sumP / 49.0, mLm = sumM / 49.0;
#pragma unroll
                for (int m = 0; m < 4; ++m) { const int s = lane + 32 * m; if (s < 98) Lc[m] -= (s >= 49) ? mLm : mLp; }
            }
            for (int q = 0; q < n; ++q) {
                const size_t e = (size_t)i * TQ_CAP + q;
                const double* c = d.cf + 6 * (size_t)d.q_cf[e] + 3 * sd;
                const double cfx = c[0], cfy = c[1], cth = c[2];
                double sc_, cc_;
                sincos(cth, &sc_, &cc_);
                double ox[4], oy[4];       // sample offsets from the moving edge location: +-n_cf * side + rotated cell
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int s = lane + 32 * m;
                    const bool neg = s >= 49;
                    const int t = s - (neg ? 49 : 0);
                    const int ii = t / 7 - 3, jj = t % 7 - 3;
                    ox[m] = (neg ? sc_ * side : -sc_ * side) + (cc_ * ii - sc_ * jj);
                    oy[m] = (neg ? -cc_ * side : cc_ * side) + (sc_ * ii + cc_ * jj);
                }
                double d0 = kx - cfx, d1 = ky - cfy;       // init_disp (:602-603)
                double score = 0.0;
                bool valid = false;
                for (int it = 0; it < p.gn_max_iter; ++it) {
                    const double lx = kx - d0, ly = ky - d1;
                    double vi[4], gx[4], gy[4];
                    double sRp = 0, sRm = 0;
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        vi[m] = 0.0; gx[m] = 0.0; gy[m] = 0.0;
                        const int s = lane + 32 * m;
                        if (s < 98) {
                            int x0, y0, dx1, dy1;
                            double a, bb;
                            cell_magic(lx + ox[m], W, 1, x0, dx1, a);
                            cell_magic(ly + oy[m], H, W, y0, dy1, bb);
                            const uint2* c00 = PK + (y0 * W + x0);
                            const uint2 u00 = __ldg(c00), u10 = __ldg(c00 + dx1), u01 = __ldg(c00 + dy1), u11 = __ldg(c00 + dy1 + dx1);
                            const double w00 = (1 - a) * (1 - bb), w10 = a * (1 - bb), w01 = (1 - a) * bb, w11 = a * bb;
                            vi[m] = round_to_float(w00 * pk_i(u00) + w10 * pk_i(u10) + w01 * pk_i(u01) + w11 * pk_i(u11));
                            gx[m] = round_to_float(w00 * pk_gx(u00) + w10 * pk_gx(u10) + w01 * pk_gx(u01) + w11 * pk_gx(u11)) * 0.125;
                            gy[m] = round_to_float(w00 * pk_gy(u00) + w10 * pk_gy(u10) + w01 * pk_gy(u01) + w11 * pk_gy(u11)) * 0.125;
                            if (s >= 49) sRm += vi[m]; else sRp += vi[m];
                        }
                    }
                    warp_sum2(sRp, sRm);
                    const double mRp = sRp / 49.0, mRm = sRm / 49.0;
                    double h00 = 0, h10 = 0, h11 = 0, b0 = 0, b1 = 0, cost = 0;
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int s = lane + 32 * m;
                        if (s < 98) {
                            const double r = Lc[m] - (vi[m] - ((s >= 49) ? mRm : mRp));
                            const double ar = fabs(r);
                            const double wgt = (ar < p.gn_huber) ? 1.0 : p.gn_huber / ar;     // strict <, :808
                            const double wjx = wgt * gx[m], wjy = wgt * gy[m];
                            h00 += wjx * gx[m]; h10 += wjy * gx[m]; h11 += wjy * gy[m];
                            b0 += wjx * r; b1 += wjy * r; cost += wgt * r * r;
                        }
                    }
                    warp_sum3(h00, h10, h11);
                    warp_sum2(b0, b1);
                    h00 += 98 * 1e-6; h11 += 98 * 1e-6;       // H += 1e-6 * Identity for each of the 98 samples (:811)
                    ++niter;
                    double s0, s1;
                    ldlt2_solve(h00, h10, h11, b0, b1, s0, s1);
                    const double e0 = -s0, e1 = -s1;
                    d0 += e0; d1 += e1;
                    if (sqrt(e0 * e0 + e1 * e1) < p.gn_tol || it == p.gn_max_iter - 1) {
                        const double rms = sqrt(warp_sum(cost) / 98.0);       // the residual is only read at the last iteration (:843-847)
                        valid = !((rms > p.gn_huber * 2.0) || (it < 1)); score = rms;
                        break;
                    }
                }
                ++nprob;
                if (lane == 0) {
                    d.q_sc[2 * e + sd] = score;
                    if (valid) { double* o = (sd ? d.q_r : d.q_l) + 3 * e; o[0] = kx - d0; o[1] = ky - d1; }     // :623-632
                    if (valid) atomicOr(&d.q_valid[e], 1 << sd);                                        // refine_validity = both bits (valid_left && valid_right)
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0 && nprob) { atomicAdd(&d.counters[1], nprob); atomicAdd(&d.counters[2], niter); }
}

// ------------------------------------------------------------------------------------------------------
// tq_gn_tile_kernel (default): the 2-D refinement with the data path of gn_lerp64_kernel.  Lanes 0-15 own the "+" patch,
// 16-31 the "-" patch (3 sample rounds + the cooperative 49th sample); per quad each half-warp stages the pixels its patch
// can reach while the edge stays within +-R px (both axes) of the build position in a warp-private tile of packed
// {half I, -, half gx, half gy} pixels; interpolation form on exact fp16 corner differences; FP64 throughout.
// ------------------------------------------------------------------------------------------------------
constexpr int TQ_TILE_PX = 320;      // 19 x 16 pixels: reach 2.5 px for every orientation
__global__ void __launch_bounds__(32 * WPB, 4) tq_gn_tile_kernel(TqDev d, DevParams p)
{
    __shared__ uint2 s_tile[WPB][2][TQ_TILE_PX];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int hw = lane >> 4, hl = lane & 15;
    uint2* tF = s_tile[w][hw];
    const int u = hl % 6;                                      // cooperative lanes of this half-warp's left-over sample
    const int cRow = u & 1, cCh = u >> 1;
    const int W = d.W, H = d.H;
    const double side = 7 / 2.0 + 1.0, huber = p.gn_huber;
    const double MAGIC = 6755399441055744.0;
    const int sd = blockIdx.y;
    const uint8_t* Ikf = sd ? d.kfRund : d.kfLund;
    const uint2* __restrict__ PK = d.pkh[sd];
    unsigned long long nprob = 0, niter = 0;
    unsigned long long* cursor = d.counters + 5 + sd;
    for (;;) {
        int i = 0;
        if (lane == 0) i = (int)atomicAdd(cursor, 1ull);
        i = __shfl_sync(FULL, i, 0);
        if (i >= d.n_kf) break;
        const int n = d.cnt[i];
        if (n == 0) continue;
        const double* k = d.kf + 6 * (size_t)i + 3 * sd;
        const double kx = k[0], ky = k[1];
        // keyframe patches, once per mate and view (:745-768)
        double Lc[3], Lc48;
        {
            double st_, ct_;
            sincos(k[2], &st_, &ct_);
            const double cx = hw ? st_ * side : -st_ * side, cy = hw ? -ct_ * side : ct_ * side;      // +-n*side, n = (-t.y, t.x)
            double sumL = 0;
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const int t = hl + 16 * m, ii = t / 7 - 3, jj = t % 7 - 3;
                Lc[m] = sample_u8_exact(Ikf, d.pitch, W, H, (kx + cx) + (ct_ * ii - st_ * jj), (ky + cy) + (st_ * ii + ct_ * jj));
                sumL += Lc[m];
            }
            Lc48 = sample_u8_exact(Ikf, d.pitch, W, H, (kx + cx) + (ct_ * 3 - st_ * 3), (ky + cy) + (st_ * 3 + ct_ * 3));
            sumL = half_sum(sumL) + Lc48;
            const double mL = sumL / 49.0;
#pragma unroll
            for (int m = 0; m < 3; ++m) Lc[m] -= mL;
            Lc48 -= mL;
        }
        for (int q = 0; q < n; ++q) {
            const size_t e = (size_t)i * TQ_CAP + q;
            const double* c = d.cf + 6 * (size_t)d.q_cf[e] + 3 * sd;
            double sc_, cc_;
            sincos(c[2], &sc_, &cc_);
            const double cxp = hw ? sc_ * side : -sc_ * side, cyp = hw ? -cc_ * side : cc_ * side;    // this half-warp's patch centre offset
            double rx[3], ry[3];
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const int t = hl + 16 * m, ii = t / 7 - 3, jj = t % 7 - 3;
                rx[m] = cc_ * ii - sc_ * jj; ry[m] = sc_ * ii + cc_ * jj;
            }
            const double rx48 = cc_ * 3 - sc_ * 3, ry48 = sc_ * 3 + cc_ * 3;
            // tile shape: the patch stays inside while the edge is within +-R px of the build position on both axes
            const double hext = 3.0 * (fabs(cc_) + fabs(sc_)) + 1e-6;
            double R = 2.5;
            int TWp, THp;
            for (;;) {
                TWp = (int)ceil(2.0 * (R + hext)) + 2;
                THp = TWp;
                while ((0xC107 >> (TWp & 15)) & 1) ++TWp;      // row pitch off the bank-folding residues (see gn_tile64_kernel)
                if (TWp * THp <= TQ_TILE_PX || R <= 0.0) break;
                R -= 0.5;
            }
            const int npx = TWp * THp;
            const float invTW = 1.0f / (float)TWp;
            const double Rv = R - 1e-6, ext = R + hext;
            double d0 = kx - c[0], d1 = ky - c[1];       // init_disp (:602-603)
            double lx0 = CUDART_NAN, ly0 = CUDART_NAN, score = 0.0;
            int ox = 0, oy = 0;
            bool valid = false;
            for (int it = 0; it < p.gn_max_iter; ++it) {
                const double lx = kx - d0, ly = ky - d1;
                const double xs = lx + cxp, ys = ly + cyp;
                if (!(fabs(lx - lx0) <= Rv && fabs(ly - ly0) <= Rv)) {
                    lx0 = lx; ly0 = ly;
                    ox = __double2int_rd(xs - ext); oy = __double2int_rd(ys - ext);
                    __syncwarp();
                    for (int t = hl; t < npx; t += 16) {
                        const int py = (int)(((float)t + 0.5f) * invTW), px = t - py * TWp;
                        const int X = min(max(ox + px, 0), W - 1), Y = min(max(oy + py, 0), H - 1);
                        tF[t] = __ldg(PK + (Y * W + X));
                    }
                    __syncwarp();
                }
                double vi[3], gxv[3], gyv[3];
                double sR = 0;
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    const double x = xs + rx[m], y = ys + ry[m];
                    const double tx = __dadd_rd(x, MAGIC), ty = __dadd_rd(y, MAGIC);
                    const double a = x - (tx - MAGIC), bb = y - (ty - MAGIC);
                    const int xi = (int)min((unsigned)(__double2loint(tx) - ox), (unsigned)(TWp - 2));
                    const int yi = (int)min((unsigned)(__double2loint(ty) - oy), (unsigned)(THp - 2));
                    const int o = yi * TWp + xi;
                    const uint2 p00 = tF[o], p10 = tF[o + 1], p01 = tF[o + TWp], p11 = tF[o + TWp + 1];
                    const unsigned d0x = hsub2_u32(p10.x, p00.x), d0y = hsub2_u32(p10.y, p00.y);
                    const unsigned d1x = hsub2_u32(p11.x, p01.x), d1y = hsub2_u32(p11.y, p01.y);
                    double top = fma(a, h2d(d0x), h2d(p00.x)), bot = fma(a, h2d(d1x), h2d(p01.x));
                    vi[m] = round_to_float(fma(bb, bot - top, top));
                    top = fma(a, h2d(d0y), h2d(p00.y)); bot = fma(a, h2d(d1y), h2d(p01.y));
                    gxv[m] = round_to_float(fma(bb, bot - top, top));
                    top = fma(a, h2d(d0y >> 16), h2d(p00.y >> 16)); bot = fma(a, h2d(d1y >> 16), h2d(p01.y >> 16));
                    gyv[m] = round_to_float(fma(bb, bot - top, top));
                    sR += vi[m];
                }
                double vi48, gx48, gy48;
                {
                    const double x = xs + rx48, y = ys + ry48;
                    const double tx = __dadd_rd(x, MAGIC), ty = __dadd_rd(y, MAGIC);
                    const double a = x - (tx - MAGIC), bb = y - (ty - MAGIC);
                    const int xi = (int)min((unsigned)(__double2loint(tx) - ox), (unsigned)(TWp - 2));
                    const int yi = (int)min((unsigned)(__double2loint(ty) - oy), (unsigned)(THp - 2));
                    const int o = (yi + cRow) * TWp + xi;
                    const uint2 p0 = tF[o], p1 = tF[o + 1];
                    const unsigned s0 = cCh == 0 ? p0.x : (cCh == 1 ? p0.y : p0.y >> 16);
                    const unsigned s1 = cCh == 0 ? p1.x : (cCh == 1 ? p1.y : p1.y >> 16);
                    const double lin = fma(a, h2d(hsub2_u32(s1, s0)), h2d(s0));
                    const double oth = shfl_xor_d(lin, 1);
                    const double top = cRow ? oth : lin, bot = cRow ? lin : oth;
                    const double v = round_to_float(fma(bb, bot - top, top));
                    vi48 = shfl_idx_d(v, hw << 4); gx48 = shfl_idx_d(v, (hw << 4) + 2); gy48 = shfl_idx_d(v, (hw << 4) + 4);
                }
                sR = half_sum(sR) + vi48;
                const double mR = div49(sR);
                double h00 = 0, h10 = 0, h11 = 0, b0 = 0, b1 = 0, cost = 0;
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const double r = (m < 3 ? Lc[m] : Lc48) - ((m < 3 ? vi[m] : vi48) - mR);
                    const double gx = m < 3 ? gxv[m] : gx48, gy = m < 3 ? gyv[m] : gy48;
                    const double ar = fabs(r);
                    double wgt = (ar < huber) ? 1.0 : huber * rcp_fast(ar);              // strict <, :808
                    if (m == 3 && hl != 0) wgt = 0.0;                                     // sample 48 counts once per patch
                    const double wjx = wgt * gx, wjy = wgt * gy;
                    h00 = fma(wjx, gx, h00); h10 = fma(wjy, gx, h10); h11 = fma(wjy, gy, h11);
                    b0 = fma(wjx, r, b0); b1 = fma(wjy, r, b1); cost = fma(wgt * r, r, cost);
                }
                warp_sum4(h00, h10, h11, b0, lane);
                b1 = warp_sum(b1);
                h00 += 98 * 1e-6; h11 += 98 * 1e-6;       // H += 1e-6 * Identity for each of the 98 samples (:811)
                ++niter;
                double s0, s1;
                ldlt2_solve(h00, h10, h11, b0, b1, s0, s1);
                const double e0 = -s0, e1 = -s1;
                d0 += e0; d1 += e1;
                if (sqrt(e0 * e0 + e1 * e1) < p.gn_tol || it == p.gn_max_iter - 1) {
                    const double rms = sqrt(warp_sum(cost) / 98.0);
                    valid = !((rms > huber * 2.0) || (it < 1)); score = rms;
                    break;
                }
            }
            ++nprob;
            if (lane == 0) {
                d.q_sc[2 * e + sd] = score;
                if (valid) { double* o = (sd ? d.q_r : d.q_l) + 3 * e; o[0] = kx - d0; o[1] = ky - d1; }     // :623-632
                if (valid) atomicOr(&d.q_valid[e], 1 << sd);
            }
            __syncwarp();
        }
    }
    if (lane == 0 && nprob) { atomicAdd(&d.counters[1], nprob); atomicAdd(&d.counters[2], niter); }
}

// Temporal_Matches.cpp:636-733
__global__ void __launch_bounds__(32 * WPB) tq_cluster_kernel(TqDev d, DevParams p)
{
    __shared__ double s_x[WPB][TQ_CAP], s_y[WPB][TQ_CAP], s_t[WPB][TQ_CAP];
    __shared__ double s_ox[WPB][TQ_CAP], s_oy[WPB][TQ_CAP], s_ot[WPB][TQ_CAP];
    __shared__ double s_dk[WPB][TQ_CAP], s_gk[WPB][TQ_CAP];
    __shared__ int s_lab[WPB][TQ_CAP], s_csz[WPB][TQ_CAP], s_near[WPB][TQ_CAP];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long* cursor = d.counters + 7;      // keyframe mates are pulled from a cursor: their lists hold 0 .. 128 quads
    for (;;) {
        int i = 0;
        if (lane == 0) i = (int)atomicAdd(cursor, 1ull);
        i = __shfl_sync(FULL, i, 0);
        if (i >= d.n_kf) break;
        const int n = d.cnt[i];
        const size_t base = (size_t)i * TQ_CAP;
        if (n < 2) {          // :642-643: lists of fewer than two quads are left alone
            if (lane == 0) {
                d.cnt2[i] = n;
                if (n == 1) {
                    d.r_cf[base] = d.q_cf[base]; d.r_valid[base] = d.q_valid[base];
                    for (int k = 0; k < 2; ++k) { d.r_ncc[2 * base + k] = d.q_ncc[2 * base + k]; d.r_sc[2 * base + k] = d.q_sc[2 * base + k]; d.r_sift[2 * base + k] = d.q_sift[2 * base + k]; }
                    for (int k = 0; k < 3; ++k) { d.r_l[3 * base + k] = d.q_l[3 * base + k]; d.r_r[3 * base + k] = d.q_r[3 * base + k]; }
                }
            }
            continue;
        }
        for (int k = lane; k < n; k += 32) { s_x[w][k] = d.q_l[3 * (base + k)]; s_y[w][k] = d.q_l[3 * (base + k) + 1]; s_t[w][k] = d.q_l[3 * (base + k) + 2]; }
        __syncwarp();
        const int ncl = warp_cluster(s_x[w], s_y[w], s_t[w], n, true, p, lane, s_lab[w], s_csz[w], s_dk[w], s_gk[w], s_ox[w], s_oy[w], s_ot[w]);
        __syncwarp();
        // the shifted_left edge closest to every contributor, first minimum (:670-680); a contributor is at distance 0 from itself
        for (int m = lane; m < n; m += 32) {
            int ci = -1; double cd = CUDART_INF;     // std::numeric_limits<double>::max() in the reference; distances are finite
            for (int k = 0; k < n; ++k) {
                const double dx = s_x[w][m] - s_x[w][k], dy = s_y[w][m] - s_y[w][k];
                const double dd = sqrt(dx * dx + dy * dy);
                if (dd < cd) { cd = dd; ci = k; }
            }
            s_near[w][m] = ci;
        }
        __syncwarp();
        for (int c = lane; c < ncl; c += 32) {       // one lane per cluster, members in index order
            double sx = 0, sy = 0, st = 0;
            int cnt = 0, best = -1;
            for (int m = 0; m < n; ++m) if (s_lab[w][m] == c) {
                const int ci = s_near[w][m];
                if (ci >= 0) { sx += d.q_r[3 * (base + ci)]; sy += d.q_r[3 * (base + ci) + 1]; st += d.q_r[3 * (base + ci) + 2]; ++cnt; best = ci; }
            }
            // ascending cluster index = order of returned_clusters; every cluster has members, so none is skipped (:684)
            const size_t o = base + c, b = base + best;
            d.r_cf[o] = d.q_cf[b]; d.r_valid[o] = d.q_valid[b];
            d.r_ncc[2 * o] = d.q_ncc[2 * b]; d.r_ncc[2 * o + 1] = d.q_ncc[2 * b + 1];
            d.r_sc[2 * o] = d.q_sc[2 * b]; d.r_sc[2 * o + 1] = d.q_sc[2 * b + 1];
            d.r_sift[2 * o] = d.q_sift[2 * b]; d.r_sift[2 * o + 1] = d.q_sift[2 * b + 1];
            d.r_l[3 * o] = s_ox[w][c]; d.r_l[3 * o + 1] = s_oy[w][c]; d.r_l[3 * o + 2] = s_ot[w][c];
            if (cnt == 1) { d.r_r[3 * o] = sx; d.r_r[3 * o + 1] = sy; d.r_r[3 * o + 2] = st; }
            else { d.r_r[3 * o] = sx / cnt; d.r_r[3 * o + 1] = sy / cnt; d.r_r[3 * o + 2] = st / cnt; }
        }
        if (lane == 0) d.cnt2[i] = ncl;
        __syncwarp();
    }
}

// ordered compaction of one pool into ebvo_quad records; which = 1 (pool after gate / GN) or 2 (after clustering)
__global__ void tq_gather_kernel(TqDev d, int which, const int* offs, ebvo_quad* out, int cap)
{
    const int i = blockIdx.x;
    const int* cnt = which == 2 ? d.cnt2 : d.cnt;
    const int n = cnt[i], o0 = offs[i];
    const int* cf = which == 2 ? d.r_cf : d.q_cf; const int* va = which == 2 ? d.r_valid : d.q_valid;
    const double* sf = which == 2 ? d.r_sift : d.q_sift;
    const double *ncc = which == 2 ? d.r_ncc : d.q_ncc, *sc = which == 2 ? d.r_sc : d.q_sc, *l = which == 2 ? d.r_l : d.q_l, *r = which == 2 ? d.r_r : d.q_r;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        if (o0 + k >= cap) break;
        const size_t e = (size_t)i * TQ_CAP + k;
        ebvo_quad q;
        q.kf_index = i; q.cf_index = cf[e];
        q.lx = l[3 * e]; q.ly = l[3 * e + 1]; q.ltheta = l[3 * e + 2];
        q.rx = r[3 * e]; q.ry = r[3 * e + 1]; q.rtheta = r[3 * e + 2];
        q.ncc_left = ncc[2 * e]; q.ncc_right = ncc[2 * e + 1];
        q.sift_left = sf[2 * e]; q.sift_right = sf[2 * e + 1];
        q.score_left = sc[2 * e]; q.score_right = sc[2 * e + 1];
        q.valid = va[e] == 3; q.reserved = 0;
        out[o0 + k] = q;
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
void tq_prepare(const TqDev& d, const DevParams& p, cudaStream_t st, Prof* prof)
{
    const int ncell = d.gw * d.gh;
    cudaMemsetAsync(d.cellCount, 0, sizeof(int) * ncell, st);
    cudaMemsetAsync(d.cellCursor, 0, sizeof(int) * ncell, st);
    cudaMemsetAsync(d.counters, 0, sizeof(unsigned long long) * 8, st);
    if (d.n_cf > 0) EBVO_KERNEL(prof, "tq_cells", st, (tq_cells_kernel<<<(d.n_cf + 255) / 256, 256, 0, st>>>(d)));
    EBVO_KERNEL(prof, "tq_scan", st, (tq_scan_kernel<<<1, 1024, 0, st>>>(d.cellCount, d.cellStart, ncell, nullptr)));
    if (d.n_cf > 0) EBVO_KERNEL(prof, "tq_fill", st, (tq_fill_kernel<<<(d.n_cf + 255) / 256, 256, 0, st>>>(d)));
    EBVO_KERNEL(prof, "tq_sort", st, (tq_sort_kernel<<<(ncell + 127) / 128, 128, 0, st>>>(d)));
}
static int tq_warp_blocks(int n)
{
    if (!g_sms) warp_grid(1);
    const int want = (n + WPB - 1) / WPB;
    return want < 1 ? 1 : (want > g_sms * 16 ? g_sms * 16 : want);
}
void tq_patches(const TqDev& d, const DevParams& p, cudaStream_t st, Prof* prof)
{
    const int n = d.n_kf > d.n_cf ? d.n_kf : d.n_cf;
    if (n > 0) EBVO_KERNEL(prof, "tq_patch", st, (tq_patch_kernel<<<dim3(tq_warp_blocks(n), 4), 32 * WPB, 0, st>>>(d, p)));
    const dim3 g((d.W + 31) / 32, (d.H + 7) / 8), t(32, 8);
    EBVO_KERNEL(prof, "tq_pack", st, (tq_pack_kernel<<<g, t, 0, st>>>(d.cfLund, d.W, d.H, d.pitch, d.pk16[0], d.pkh[0])));
    EBVO_KERNEL(prof, "tq_pack", st, (tq_pack_kernel<<<g, t, 0, st>>>(d.cfRund, d.W, d.H, d.pitch, d.pk16[1], d.pkh[1])));
}
void tq_gate(const TqDev& d, int mode, int* counts, const int* offs, int* outCf, cudaStream_t st, Prof* prof)
{
    if (d.n_kf > 0) EBVO_KERNEL(prof, "tq_gate", st, (tq_gate_kernel<<<tq_warp_blocks(d.n_kf), 32 * WPB, 0, st>>>(d, mode, counts, offs, outCf)));
}
void tq_gn(const TqDev& d, const DevParams& p, cudaStream_t st, Prof* prof)
{
    if (!g_sms) warp_grid(1);
    if (d.n_kf > 0 && d.gn_gather) EBVO_KERNEL(prof, "tq_gn", st, (tq_gn_kernel<<<dim3(g_sms * 2, 2), 32 * WPB, 0, st>>>(d, p)));
    else if (d.n_kf > 0) EBVO_KERNEL(prof, "tq_gn", st, (tq_gn_tile_kernel<<<dim3(g_sms * 2, 2), 32 * WPB, 0, st>>>(d, p)));
}
void tq_cluster(const TqDev& d, const DevParams& p, cudaStream_t st, Prof* prof)
{
    if (d.n_kf > 0) EBVO_KERNEL(prof, "tq_cluster", st, (tq_cluster_kernel<<<tq_warp_blocks(d.n_kf), 32 * WPB, 0, st>>>(d, p)));
}
void tq_scan(const int* in, int* out, int n, cudaStream_t st, Prof* prof)
{
    EBVO_KERNEL(prof, "tq_scan", st, (tq_scan_kernel<<<1, 1024, 0, st>>>(in, out, n, nullptr)));
}
void tq_gather(const TqDev& d, int which, const int* offs, ebvo_quad* out, int cap, cudaStream_t st, Prof* prof)
{
    if (d.n_kf > 0) EBVO_KERNEL(prof, "tq_gather", st, (tq_gather_kernel<<<d.n_kf, 32, 0, st>>>(d, which, offs, out, cap)));
}
