/*
 * TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement (C++17, FP64 where the reference is FP64) of the
 * reference's stereo edge-correspondence path.  It is the checker for the CUDA path; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (edge_based_visual_odometry_b200/csrc) never links or calls anything in this directory.
 *
 * PARITY PINNING: pinned against the reference's OWN stereo sources.  /root/reference/src/Stereo_Matches.cpp,
 * src/utility.cpp and src/EdgeClusterer.cpp are compiled in place, unmodified, into oracle/_ref/libstereo_ref.so
 * (oracle/Makefile) against stand-in headers for the absent third-party libraries (third_party_shim: a minimal
 * cv::Mat/MatExpr/Sobel/mean/sum/dot, fixed-size Eigen matrices, a yaml-cpp stub) and driven stage by stage by
 * oracle/ref_stereo_harness.cpp in the order of get_Stereo_Edge_Pairs.  On the KITTI-shape pair (32.6 k edges per
 * view) every stage's candidate lists, the Gauss-Newton iterates, the cluster centres and the 27 095 final mates of
 * this restatement are identical to that build (positions bit-identical; NCC values within 4e-6, which is the
 * spread between two equally valid readings of OpenCV's CV_32F type mix).  tests/test_oracle_stereo.py checks this
 * live when oracle/_ref exists and against tests/golden/stereo_ref_small.npz (reference output) otherwise.
 * What stays unpinned: the arithmetic INSIDE OpenCV/Eigen (the stand-ins restate it; Sobel is checked bit-exact and
 * the NCC type mix to 2e-6 against cv2 4.13) and cv::SIFT (not run: SIFT-off).
 *
 * What it follows (all paths under /root/reference):
 *   S0  F21 = Kr^-T [T]x R Kl^-1                       src/Dataset.cpp:102-112, src/utility.cpp:33-43
 *   S1  epipolar-distance gate                          src/Stereo_Matches.cpp:10-20,91-109,381-419
 *   S2  max-disparity gate                              src/Stereo_Matches.cpp:534-553
 *   S3  orientation gate                                src/Stereo_Matches.cpp:863-915
 *   S4  SIFT gate (optional: descriptors injected)      src/Stereo_Matches.cpp:655-787
 *   S5  oriented 7x7 patches                            src/utility.cpp:82-93,141-161,182-212; include/utility.h:81-104
 *   S6  NCC + gate                                      src/utility.cpp:163-180; src/Stereo_Matches.cpp:555-616
 *   S7  best-nearly-best (NCC 0.9; SIFT 0.4 optional)   src/Stereo_Matches.cpp:789-862
 *   S8  shift to the epipolar line                      src/Stereo_Matches.cpp:26-89,967-1037; src/utility.cpp:46-74
 *   S9  Gauss-Newton refinement along the line          src/Stereo_Matches.cpp:1159-1358; include/utility.h:131-192
 *   S10 second shift + EdgeClusterer                    src/Stereo_Matches.cpp:1483,967-1037; src/EdgeClusterer.cpp:1-302
 *   S11 NCC on cluster centres                          src/Stereo_Matches.cpp:1500,555-616
 *   S12 arg-max                                         src/Stereo_Matches.cpp:916-965
 *   S13 clean-up + finalisation                         src/Stereo_Matches.cpp:1543-1653
 * Constants: include/definitions.h:17-36.  Call order: src/Stereo_Matches.cpp:1360-1540.
 *
 * Documented deviations from the reference:
 *   (1) the dead read contributing_edges_toed_indices[0] at Stereo_Matches.cpp:991 on the second
 *       consolidate call (vectors emptied by the first call => undefined behaviour) is skipped;
 *   (2) SIFT descriptors are not computed here: mode 0 skips S4 and S7' entirely ("SIFT-off");
 *       mode 1 takes per-edge descriptors computed by the caller (cv2.SIFT in the test harness);
 *   (3) outputs the reference leaves uninitialised when H < 1e-8 in the first GN iteration
 *       (Stereo_Matches.cpp:1253) are written as score = 0, confidence = 0, validity = false.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <numeric>
#include <unordered_set>
#include <vector>
#include <omp.h>

namespace {

struct P {  // the reference's #define knobs (definitions.h:17-36)
    double epi_dist = 0.5, max_disp = 25.0, orient_deg = 10.0, ncc_thresh = 0.6;
    double shift_mag = 5.0;  // ORTHOGONAL_SHIFT_MAG
    double loc_pert = 0.4, tang_displ = 3.0, orient_pert = 0.174533;
    double clus_dist = 1.0, clus_orient_deg = 20.0, clus_sigma = 2.0;
    int clus_max = 10;
    double bnb_ncc = 0.9, bnb_sift = 0.4, sift_thresh = 500.0;
    int gn_max_iter = 20;
    double gn_tol = 1e-3, gn_huber = 3.0;  // defaults of Stereo_Matches.h:84
};

struct E { double x, y, th; };

struct Cand {
    E e;          // center_edge
    int ridx;     // contributing_edges_toed_indices[0]; -1 once dropped (S8)
};

struct Cl {  // Stereo_Matching_Edge_Clusters (Dataset.h:169-179)
    std::vector<Cand> c;
    std::vector<double> score, conf;
    std::vector<char> valid;
    std::vector<int> iters;   // diagnostics: GN iterations per candidate
};

struct Img {
    int H = 0, W = 0;
    std::vector<double> d;  // CV_64F copy (convertTo, Stereo_Matches.cpp:562-563)
    std::vector<float> f;   // CV_32F copy (:1293-1294)
};

const double NaN = std::numeric_limits<double>::quiet_NaN();

// include/utility.h:81-104 (Bilinear_Interpolation<double> on a CV_64F image)
inline double bilinear64(const Img &I, double px, double py)
{
    if (!(px == px) || !(py == py)) return NaN;
    double fx = std::floor(px), cx = std::ceil(px), fy = std::floor(py), cy = std::ceil(py);
    // Q12=(fx,fy) Q22=(cx,fy) Q11=(fx,cy) Q21=(cx,cy)
    if (fx < 0 || cy < 0 || cx >= I.W || cy >= I.H || fx < 0 || fy < 0 || cx >= I.W || fy >= I.H) return NaN;
    const double *d = I.d.data();
    double v11 = d[(size_t)(int)cy * I.W + (int)fx], v21 = d[(size_t)(int)cy * I.W + (int)cx];
    double v12 = d[(size_t)(int)fy * I.W + (int)fx], v22 = d[(size_t)(int)fy * I.W + (int)cx];
    double f1 = ((cx - px) / (cx - fx)) * v11 + ((px - fx) / (cx - fx)) * v21;
    double f2 = ((cx - px) / (cx - fx)) * v12 + ((px - fx) / (cx - fx)) * v22;
    return ((fy - py) / (fy - cy)) * f1 + ((py - cy) / (fy - cy)) * f2;
}

// src/utility.cpp:82-93,141-161,182-212 -> two 7x7 float patches ("+" then "-")
inline void edge_patches(const E &e, const Img &I, double shift, float *plus, float *minus)
{
    double s = std::sin(e.th), c = std::cos(e.th);
    double px[2] = {e.x + shift * s, e.x + shift * (-s)};
    double py[2] = {e.y + shift * (-c), e.y + shift * c};
    float *out[2] = {plus, minus};
    for (int k = 0; k < 2; ++k)
        for (int i = -3; i <= 3; ++i)
            for (int j = -3; j <= 3; ++j) {
                double rx = std::cos(e.th) * (i)-std::sin(e.th) * (j) + px[k];
                double ry = std::sin(e.th) * (i) + std::cos(e.th) * (j) + py[k];
                out[k][(i + 3) * 7 + (j + 3)] = (float)bilinear64(I, rx, ry);
            }
}

// src/utility.cpp:163-180 with OpenCV's CV_32F type mix
inline double patch_similarity(const float *a, const float *b)
{
    double sa = 0, sb = 0;
    for (int k = 0; k < 49; ++k) { sa += a[k]; sb += b[k]; }
    double ma = sa / 49.0, mb = sb / 49.0;
    float fa = (float)ma, fb = (float)mb;
    double ssa = 0, ssb = 0;
    float da[49], db[49];
    for (int k = 0; k < 49; ++k) {
        da[k] = a[k] - fa; db[k] = b[k] - fb;
        ssa += (double)(float)(da[k] * da[k]);
        ssb += (double)(float)(db[k] * db[k]);
    }
    if (ssa < 1e-10 || ssb < 1e-10) return -1.0;
    float ia = (float)(1.0 / std::sqrt(ssa)), ib = (float)(1.0 / std::sqrt(ssb));
    double dot = 0;
    for (int k = 0; k < 49; ++k) dot += (double)(float)(da[k] * ia) * (double)(float)(db[k] * ib);
    return dot;
}

// std::max({pp,nn,pn,np}) (Stereo_Matches.cpp:596): left fold with operator<
inline double max4(double a, double b, double c, double d)
{
    double m = a;
    if (m < b) m = b;
    if (m < c) m = c;
    if (m < d) m = d;
    return m;
}

// src/utility.cpp:46-54
inline double normal_dist(const double *l, double x, double y, double &ex, double &ey)
{
    double a = l[0], b = l[1], c = l[2];
    ex = x - a * (a * x + b * y + c) / (std::pow(a, 2) + std::pow(b, 2));
    ey = y - b * (a * x + b * y + c) / (std::pow(a, 2) + std::pow(b, 2));
    return std::sqrt(std::pow(x - ex, 2) + std::pow(y - ey, 2));
}
// src/utility.cpp:63-74
inline double tangential_dist(const double *l, double x, double y, double th, double &xi, double &yi)
{
    double ae = std::tan(th), be = -1, ce = -(ae * x - y);
    double a = l[0], b = l[1], c = l[2];
    xi = (b * ce - be * c) / (a * be - ae * b);
    yi = (c * ae - ce * a) / (a * be - ae * b);
    return std::sqrt((xi - x) * (xi - x) + (yi - y) * (yi - y));
}
// src/Stereo_Matches.cpp:26-89
inline E shift_to_line(const E &o, const double *l, const P &p)
{
    double ex, ey;
    if (normal_dist(l, o.x, o.y, ex, ey) < p.loc_pert) return E{ex, ey, o.th};
    double xi, yi;
    if (tangential_dist(l, o.x, o.y, o.th, xi, yi) < p.tang_displ) return E{xi, yi, o.th};
    double th = o.th;
    double pt = l[0] * std::cos(th) + l[1] * std::sin(th);
    double dpt = -l[0] * std::sin(th) + l[1] * std::cos(th);
    if (pt > 0 && dpt < 0) th -= p.orient_pert;
    else if (pt < 0 && dpt < 0) th -= p.orient_pert;
    else if (pt > 0 && dpt > 0) th += p.orient_pert;
    else if (pt < 0 && dpt > 0) th += p.orient_pert;
    if (tangential_dist(l, o.x, o.y, th, xi, yi) < p.tang_displ) return E{xi, yi, th};
    return o;
}

// include/utility.h:159-172 (clamped, float image, returns float)
inline float sampleF(const float *I, int w, int h, double x, double y)
{
    x = std::clamp(x, 0.0, (double)w - 1.0);
    y = std::clamp(y, 0.0, (double)h - 1.0);
    int x0 = (int)std::floor(x), y0 = (int)std::floor(y);
    int x1 = std::min(x0 + 1, w - 1), y1 = std::min(y0 + 1, h - 1);
    double a = x - x0, b = y - y0;
    float v00 = I[(size_t)y0 * w + x0], v10 = I[(size_t)y0 * w + x1], v01 = I[(size_t)y1 * w + x0], v11 = I[(size_t)y1 * w + x1];
    return (float)((1 - a) * (1 - b) * v00 + a * (1 - b) * v10 + (1 - a) * b * v01 + a * b * v11);
}
// include/utility.h:143-157,174-181
inline void sample_patch(const float *I, int w, int h, double cx, double cy, double th, double *v)
{
    double ct = std::cos(th), st = std::sin(th);
    int k = 0;
    for (int i = -3; i <= 3; ++i)
        for (int j = -3; j <= 3; ++j) v[k++] = (double)sampleF(I, w, h, cx + ct * i - st * j, cy + st * i + ct * j);
}
inline double mean49(const double *v)
{
    double s = 0;
    for (int k = 0; k < 49; ++k) s += v[k];
    return s / 49;
}

// include/utility.h:131-141: cv::Sobel(CV_32F, ksize 3, scale 1/8, BORDER_REFLECT_101); exact on uint8 data
void sobel(const std::vector<float> &I, int H, int W, std::vector<float> &gx, std::vector<float> &gy)
{
    gx.assign((size_t)H * W, 0.f);
    gy.assign((size_t)H * W, 0.f);
    auto R = [](int i, int n) { if (n == 1) return 0; if (i < 0) return -i; if (i >= n) return 2 * n - 2 - i; return i; };
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        int ym = R(y - 1, H), yp = R(y + 1, H);
        for (int x = 0; x < W; ++x) {
            int xm = R(x - 1, W), xp = R(x + 1, W);
            auto at = [&](int yy, int xx) { return I[(size_t)yy * W + xx]; };
            gx[(size_t)y * W + x] = ((at(ym, xp) - at(ym, xm)) * 0.125f + (at(y, xp) - at(y, xm)) * 0.25f + (at(yp, xp) - at(yp, xm)) * 0.125f);
            gy[(size_t)y * W + x] = ((at(yp, xm) - at(ym, xm)) * 0.125f + (at(yp, x) - at(ym, x)) * 0.25f + (at(yp, xp) - at(ym, xp)) * 0.125f);
        }
    }
}

// src/Stereo_Matches.cpp:1159-1288
void gn_refine(const E &L, const E &Rc, double dx, double dy, const Img &IL, const Img &IR, const std::vector<float> &gxR,
               const std::vector<float> &gyR, const P &p, double &alpha_out, double &score, double &conf, bool &valid, int &iters)
{
    double tx = std::cos(L.th), ty = std::sin(L.th);
    double nx = -ty, ny = tx;
    const double side = 7 / 2.0 + 1.0;
    double Lp[49], Lm[49];
    sample_patch(IL.f.data(), IL.W, IL.H, L.x + nx * side, L.y + ny * side, L.th, Lp);
    sample_patch(IL.f.data(), IL.W, IL.H, L.x - nx * side, L.y - ny * side, L.th, Lm);
    double mLp = mean49(Lp), mLm = mean49(Lm);
    for (int k = 0; k < 49; ++k) { Lp[k] -= mLp; Lm[k] -= mLm; }
    double alpha = 0.0;
    score = 0; conf = 0; valid = false; iters = 0;  // deviation (3)
    int logn = 0;
    for (int it = 0; it < p.gn_max_iter; ++it) {
        double sx = alpha * dx, sy = alpha * dy;
        double Rp[49], Rm[49], gxp[49], gxm[49], gyp[49], gym[49];
        double cpx = (Rc.x + nx * side) + sx, cpy = (Rc.y + ny * side) + sy;
        double cmx = (Rc.x - nx * side) + sx, cmy = (Rc.y - ny * side) + sy;
        sample_patch(IR.f.data(), IR.W, IR.H, cpx, cpy, L.th, Rp);
        sample_patch(IR.f.data(), IR.W, IR.H, cmx, cmy, L.th, Rm);
        sample_patch(gxR.data(), IR.W, IR.H, cpx, cpy, L.th, gxp);
        sample_patch(gxR.data(), IR.W, IR.H, cmx, cmy, L.th, gxm);
        sample_patch(gyR.data(), IR.W, IR.H, cpx, cpy, L.th, gyp);
        sample_patch(gyR.data(), IR.W, IR.H, cmx, cmy, L.th, gym);
        double mRp = mean49(Rp), mRm = mean49(Rm);
        double Hh = 0, b = 0, cost = 0;
        auto acc = [&](const double *Lc, const double *Rf, const double *gxf, const double *gyf, double mR) {
            for (int k = 0; k < 49; ++k) {
                double r = Lc[k] - (Rf[k] - mR);
                double g = -gxf[k] * dx + gyf[k] * dy;
                double ar = std::abs(r);
                double w = (ar <= p.gn_huber) ? 1.0 : p.gn_huber / ar;
                Hh += w * g * g; b += w * g * r; cost += w * r * r;
            }
        };
        acc(Lp, Rp, gxp, gyp, mRp);
        acc(Lm, Rm, gxm, gym, mRm);
        if (Hh < 1e-8) break;
        double delta = -b / Hh;
        alpha += delta;
        double rms = std::sqrt(cost / 98.0);
        ++logn;
        iters = it + 1;
        bool outlier = (rms > p.gn_huber * 2.0) || (logn < 2);
        if (std::abs(delta) < p.gn_tol || it == p.gn_max_iter - 1) {
            valid = !outlier; score = rms; conf = std::exp(-rms / p.gn_huber);
            break;
        }
    }
    alpha_out = alpha;
}

// src/EdgeClusterer.cpp:43-117 (single-label form used at :234)
void gaussian_average(const std::vector<E> &e, const std::vector<int> &lab, int label, double &gx, double &gy, double &gth, const P &p)
{
    int n = (int)e.size(), cnt = 0;
    double sx = 0, sy = 0;
    for (int i = 0; i < n; ++i) if (lab[i] == label) { sx += e[i].x; sy += e[i].y; cnt++; }
    if (cnt == 0) { gx = gy = gth = 0; return; }
    double cx = sx / cnt, cy = sy / cnt, tot = 0;
    for (int i = 0; i < n; ++i) if (lab[i] == label) { double dx = e[i].x - cx, dy = e[i].y - cy; tot += std::sqrt(dx * dx + dy * dy); }
    double md = tot / cnt, wx = 0, wy = 0, wt = 0, w = 0;
    for (int i = 0; i < n; ++i) if (lab[i] == label) {
        double dx = e[i].x - cx, dy = e[i].y - cy, d = std::sqrt(dx * dx + dy * dy);
        double g = std::exp(-0.5 * std::pow((d - md) / p.clus_sigma, 2));
        wx += g * e[i].x; wy += g * e[i].y; wt += g * e[i].th; w += g;
    }
    gx = wx / w; gy = wy / w; gth = wt / w;
}

// src/EdgeClusterer.cpp:119-302 ; returns cluster centres in returned_clusters order; labels_out = renumbered labels
void cluster_edges(const std::vector<E> &in, bool by_orient, const P &p, std::vector<E> &centers, std::vector<int> &labels_out)
{
    int n = (int)in.size();
    std::vector<int> lab(n);
    std::iota(lab.begin(), lab.end(), 0);
    auto csize = [&](int l) { int s = 0; for (int i = 0; i < n; ++i) s += (lab[i] == l); return s; };
    const double oth = p.clus_orient_deg * (M_PI / 180.0);
    bool merged = true;
    while (merged) {
        merged = false;
        for (int i = 0; i < n; ++i) {
            double md = std::numeric_limits<double>::max();
            int nearest = -1;
            for (int j = 0; j < n; ++j) {
                if (lab[i] == lab[j]) continue;
                double dx = in[i].x - in[j].x, dy = in[i].y - in[j].y;
                double dist = std::sqrt(dx * dx + dy * dy);
                if (by_orient) {
                    if (dist < md && dist < p.clus_dist && std::abs(in[i].th - in[j].th) < oth) { md = dist; nearest = j; }
                } else if (dist < md && dist < p.clus_dist) { md = dist; nearest = j; }
            }
            if (nearest != -1) {
                int oldl = lab[nearest], newl = lab[i];
                if (csize(oldl) + csize(newl) <= p.clus_max) {
                    for (int k = 0; k < n; ++k) if (lab[k] == oldl) lab[k] = newl;
                    merged = true;
                    break;
                }
            }
        }
    }
    std::map<int, std::vector<int>> l2c;
    for (int i = 0; i < n; ++i) l2c[lab[i]].push_back(i);
    centers.clear();
    labels_out.assign(n, 0);
    int c = 0;
    for (auto &kv : l2c) {
        double gx, gy, gth;
        gaussian_average(in, lab, kv.first, gx, gy, gth, p);
        centers.push_back(E{gx, gy, gth});
        for (int i : kv.second) labels_out[i] = c;
        ++c;
    }
}

struct StageDump {  // one ragged list per left edge
    std::vector<int> off;      // nL+1
    std::vector<int> ridx;
    std::vector<double> x, y, th, score;
};

struct Result {
    int nL = 0;
    std::vector<StageDump> stages;  // see ST_* below
    std::vector<int> mate_left;     // left edge index per final mate
    std::vector<double> mate_rx, mate_ry, mate_rth, mate_score;
    std::vector<double> lines;      // nL*3
    std::vector<double> t_stage;    // seconds per stage
    long gn_pairs = 0, gn_iters = 0, ncc_pairs1 = 0, ncc_pairs2 = 0, s1_total = 0;
    std::vector<int> gn_iter_list;   // per GN candidate, in stage order
};

enum { ST_EPI = 0, ST_DISP, ST_ORIENT, ST_SIFT, ST_NCC, ST_BNB_NCC, ST_BNB_SIFT, ST_SHIFT, ST_GN, ST_CLUSTER, ST_NCC2, ST_BEST, ST_COUNT };

void dump(Result &r, int st, const std::vector<Cl> &cl, bool want)
{
    if (!want) return;
    StageDump &d = r.stages[st];
    int n = (int)cl.size();
    d.off.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) d.off[i + 1] = d.off[i] + (int)cl[i].c.size();
    size_t tot = d.off[n];
    d.ridx.resize(tot); d.x.resize(tot); d.y.resize(tot); d.th.resize(tot); d.score.resize(tot);
    for (int i = 0; i < n; ++i)
        for (size_t j = 0; j < cl[i].c.size(); ++j) {
            size_t k = d.off[i] + j;
            d.ridx[k] = cl[i].c[j].ridx; d.x[k] = cl[i].c[j].e.x; d.y[k] = cl[i].c[j].e.y; d.th[k] = cl[i].c[j].e.th;
            d.score[k] = j < cl[i].score.size() ? cl[i].score[j] : NaN;
        }
}

// src/Stereo_Matches.cpp:555-616.  Serial in the reference (orphaned "omp for"); `par` lets the
// cpu_baseline run it either way.
void ncc_filter(std::vector<Cl> &cl, const std::vector<E> &L, const Img &IL, const Img &IR, const P &p, long &pairs,
                std::vector<float> &left_patches)
{
    int nL = (int)L.size();
    left_patches.resize((size_t)nL * 98);
    long cnt = 0;
    for (int i = 0; i < nL; ++i) {
        float *lp = &left_patches[(size_t)i * 98], *lm = lp + 49;
        edge_patches(L[i], IL, p.shift_mag, lp, lm);
        Cl &c = cl[i];
        std::vector<Cand> sc;
        std::vector<double> ss, sf;
        std::vector<char> sv;
        for (size_t j = 0; j < c.c.size(); ++j) {
            float rp[49], rm[49];
            edge_patches(c.c[j].e, IR, p.shift_mag, rp, rm);
            double pp = patch_similarity(lp, rp), nn = patch_similarity(lm, rm);
            double pn = patch_similarity(lp, rm), np = patch_similarity(lm, rp);
            double s = max4(pp, nn, pn, np);
            ++cnt;
            if (s > p.ncc_thresh) {
                sc.push_back(c.c[j]); ss.push_back(s);
                sf.push_back(j < c.conf.size() ? c.conf[j] : 0.0);
                sv.push_back(1);
            }
        }
        c.c = std::move(sc); c.score = std::move(ss); c.conf = std::move(sf); c.valid = std::move(sv);
    }
    pairs = cnt;
}

// src/Stereo_Matches.cpp:789-862
void bnb(std::vector<Cl> &cl, double thr, bool is_ncc)
{
    int nL = (int)cl.size();
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < nL; ++i) {
        Cl &c = cl[i];
        size_t n = c.c.size();
        if (n < 2) continue;
        std::vector<size_t> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        if (is_ncc) std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return c.score[a] > c.score[b]; });
        else std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return c.conf[a] < c.conf[b]; });
        size_t keep = 1;
        double best = is_ncc ? c.score[idx[0]] : c.conf[idx[0]];
        for (size_t j = 0; j + 1 < n; ++j) {
            double next = is_ncc ? c.score[idx[j + 1]] : c.conf[idx[j + 1]];
            if (best == 0) break;
            double ratio = is_ncc ? next / best : best / next;
            if (ratio >= thr) keep++; else break;
        }
        if (keep < n) {
            Cl o;
            for (size_t k = 0; k < keep; ++k) {
                o.c.push_back(c.c[idx[k]]); o.score.push_back(c.score[idx[k]]);
                o.conf.push_back(c.conf[idx[k]]); o.valid.push_back(c.valid[idx[k]]);
            }
            c = std::move(o);
        }
    }
}

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

extern "C" {

// S0: Dataset.cpp:102-112 ; K, R row-major 3x3; F out row-major
void so_fundamental(const double *Kl, const double *Kr, const double *R21, const double *T21, double *F21, double *F12)
{
    auto inv3 = [](const double *m, double *o) {
        double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
        o[0] = (m[4] * m[8] - m[5] * m[7]) / det; o[1] = (m[2] * m[7] - m[1] * m[8]) / det; o[2] = (m[1] * m[5] - m[2] * m[4]) / det;
        o[3] = (m[5] * m[6] - m[3] * m[8]) / det; o[4] = (m[0] * m[8] - m[2] * m[6]) / det; o[5] = (m[2] * m[3] - m[0] * m[5]) / det;
        o[6] = (m[3] * m[7] - m[4] * m[6]) / det; o[7] = (m[1] * m[6] - m[0] * m[7]) / det; o[8] = (m[0] * m[4] - m[1] * m[3]) / det;
    };
    auto mul = [](const double *a, const double *b, double *o) {
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += a[i * 3 + k] * b[k * 3 + j]; o[i * 3 + j] = s; }
    };
    auto tr = [](const double *a, double *o) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o[i * 3 + j] = a[j * 3 + i]; };
    auto skew = [](const double *t, double *o) { o[0] = 0; o[1] = -t[2]; o[2] = t[1]; o[3] = t[2]; o[4] = 0; o[5] = -t[0]; o[6] = -t[1]; o[7] = t[0]; o[8] = 0; };
    double Kli[9], Kri[9], KriT[9], KliT[9], S[9], SR[9], tmp[9];
    inv3(Kl, Kli); inv3(Kr, Kri); tr(Kri, KriT); tr(Kli, KliT);
    skew(T21, S); mul(S, R21, SR); mul(KriT, SR, tmp); mul(tmp, Kli, F21);
    double R12[9], T12[3];
    tr(R21, R12);
    for (int i = 0; i < 3; ++i) T12[i] = -(R12[i * 3 + 0] * T21[0] + R12[i * 3 + 1] * T21[1] + R12[i * 3 + 2] * T21[2]);
    skew(T12, S); mul(S, R12, SR); mul(KliT, SR, tmp); mul(tmp, Kri, F12);
}

void so_sobel(const uint8_t *img, int H, int W, float *gx, float *gy)
{
    std::vector<float> I((size_t)H * W), a, b;
    for (size_t k = 0; k < I.size(); ++k) I[k] = (float)img[k];
    sobel(I, H, W, a, b);
    std::memcpy(gx, a.data(), a.size() * 4);
    std::memcpy(gy, b.data(), b.size() * 4);
}

void so_edge_patches(const uint8_t *img, int H, int W, double x, double y, double th, float *plus49, float *minus49)
{
    Img I; I.H = H; I.W = W; I.d.resize((size_t)H * W);
    for (size_t k = 0; k < I.d.size(); ++k) I.d[k] = (double)img[k];
    P p;
    edge_patches(E{x, y, th}, I, p.shift_mag, plus49, minus49);
}
double so_patch_similarity(const float *a, const float *b) { return patch_similarity(a, b); }

// cluster n edges (xyt: n*3); centers_xyt: n*3 capacity; labels: n. returns number of clusters
int so_cluster(const double *xyt, int n, int by_orientation, double *centers_xyt, int *labels)
{
    std::vector<E> in(n), c;
    for (int i = 0; i < n; ++i) in[i] = E{xyt[3 * i], xyt[3 * i + 1], xyt[3 * i + 2]};
    std::vector<int> lab;
    P p;
    cluster_edges(in, by_orientation != 0, p, c, lab);
    for (size_t k = 0; k < c.size(); ++k) { centers_xyt[3 * k] = c[k].x; centers_xyt[3 * k + 1] = c[k].y; centers_xyt[3 * k + 2] = c[k].th; }
    for (int i = 0; i < n; ++i) labels[i] = lab[i];
    return (int)c.size();
}
void so_shift_to_line(const double *line3, double x, double y, double th, double *out3)
{
    P p;
    E r = shift_to_line(E{x, y, th}, line3, p);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.th;
}

/*
 * The whole stereo stage (get_Stereo_Edge_Pairs + finalize, no-GT branch).  Images are row-major uint8
 * H x W (raw and undistorted, both views).  sift_mode 0: S4/S7' skipped.  sift_mode 1: descL (nL*2*128
 * floats) and descR (nR*2*128) are the per-edge descriptor pairs of augment_Edge_Data/apply_SIFT_filtering.
 * want_dumps != 0 keeps every stage's ragged candidate lists.  Returns an opaque handle.
 */
void *so_run(const uint8_t *Lraw, const uint8_t *Rraw, const uint8_t *Lund, const uint8_t *Rund, int H, int W,
             const double *Lxyt, int nL, const double *Rxyt, int nR, const double *F21, int sift_mode, const float *descL,
             const float *descR, int want_dumps, int omp_threads)
{
    if (omp_threads > 0) omp_set_num_threads(omp_threads);
    P p;
    Result *res = new Result;
    res->nL = nL;
    res->stages.resize(ST_COUNT);
    res->t_stage.assign(ST_COUNT + 2, 0.0);
    std::vector<E> L(nL), R(nR);
    for (int i = 0; i < nL; ++i) L[i] = E{Lxyt[3 * i], Lxyt[3 * i + 1], Lxyt[3 * i + 2]};
    for (int i = 0; i < nR; ++i) R[i] = E{Rxyt[3 * i], Rxyt[3 * i + 1], Rxyt[3 * i + 2]};
    auto mk = [&](const uint8_t *s, bool dbl, bool flt) {
        Img I; I.H = H; I.W = W;
        if (dbl) { I.d.resize((size_t)H * W); for (size_t k = 0; k < I.d.size(); ++k) I.d[k] = (double)s[k]; }
        if (flt) { I.f.resize((size_t)H * W); for (size_t k = 0; k < I.f.size(); ++k) I.f[k] = (float)s[k]; }
        return I;
    };
    double t0 = now();
    // S1 (omp parallel in the reference, :395-418)
    res->lines.resize((size_t)nL * 3);
    for (int i = 0; i < nL; ++i)
        for (int r = 0; r < 3; ++r) res->lines[3 * (size_t)i + r] = F21[3 * r] * L[i].x + F21[3 * r + 1] * L[i].y + F21[3 * r + 2] * 1.0;
    std::vector<Cl> cl(nL);
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < nL; ++i) {
        const double *l = &res->lines[3 * (size_t)i];
        for (int k = 0; k < nR; ++k) {
            double d = std::abs(l[0] * R[k].x + l[1] * R[k].y + l[2]) / std::sqrt((l[0] * l[0]) + (l[1] * l[1]));
            if (d < p.epi_dist) cl[i].c.push_back(Cand{R[k], k});
        }
    }
    for (int i = 0; i < nL; ++i) res->s1_total += (long)cl[i].c.size();
    double t1 = now(); res->t_stage[ST_EPI] = t1 - t0; t0 = t1;
    dump(*res, ST_EPI, cl, want_dumps);
    t0 = now();
    // S2 (serial, :534-553)
    for (int i = 0; i < nL; ++i) {
        std::vector<Cand> s;
        for (auto &c : cl[i].c) {
            double dx = L[i].x - c.e.x, dy = L[i].y - c.e.y;
            if (std::sqrt(dx * dx + dy * dy) <= p.max_disp) s.push_back(c);
        }
        cl[i].c = std::move(s);
    }
    t1 = now(); res->t_stage[ST_DISP] = t1 - t0;
    dump(*res, ST_DISP, cl, want_dumps);
    t0 = now();
    // S3 (omp, :863-915)
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < nL; ++i) {
        std::vector<Cand> s;
        for (auto &c : cl[i].c) {
            double d = std::abs((L[i].th - c.e.th) * (180.0 / M_PI));
            if (d > 180.0) d = 360.0 - d;
            if (d < p.orient_deg || std::abs(d - 180.0) < p.orient_deg) s.push_back(c);
        }
        cl[i].c = std::move(s);
    }
    t1 = now(); res->t_stage[ST_ORIENT] = t1 - t0;
    dump(*res, ST_ORIENT, cl, want_dumps);
    t0 = now();
    // S4 (optional)
    if (sift_mode == 1) {
        auto l2 = [](const float *a, const float *b) { double s = 0; for (int k = 0; k < 128; ++k) { double d = (double)a[k] - (double)b[k]; s += d * d; } return std::sqrt(s); };
#pragma omp parallel for schedule(dynamic, 64)
        for (int i = 0; i < nL; ++i) {
            const float *l1 = descL + (size_t)i * 256, *l2p = l1 + 128;
            std::vector<Cand> s; std::vector<double> cf;
            for (auto &c : cl[i].c) {
                const float *r1 = descR + (size_t)c.ridx * 256, *r2 = r1 + 128;
                double d = std::min({l2(l1, r1), l2(l2p, r1), l2(l1, r2), l2(l2p, r2)});
                if (d < p.sift_thresh) { s.push_back(c); cf.push_back(d); }
            }
            cl[i].c = std::move(s); cl[i].conf = std::move(cf);
        }
    }
    t1 = now(); res->t_stage[ST_SIFT] = t1 - t0;
    dump(*res, ST_SIFT, cl, want_dumps);
    // S5/S6 (serial)
    Img ILraw = mk(Lraw, true, false), IRraw = mk(Rraw, true, false);
    t0 = now();
    std::vector<float> lpatch;
    ncc_filter(cl, L, ILraw, IRraw, p, res->ncc_pairs1, lpatch);
    t1 = now(); res->t_stage[ST_NCC] = t1 - t0;
    dump(*res, ST_NCC, cl, want_dumps);
    t0 = now();
    bnb(cl, p.bnb_ncc, true);
    t1 = now(); res->t_stage[ST_BNB_NCC] = t1 - t0;
    dump(*res, ST_BNB_NCC, cl, want_dumps);
    t0 = now();
    if (sift_mode == 1) bnb(cl, p.bnb_sift, false);
    t1 = now(); res->t_stage[ST_BNB_SIFT] = t1 - t0;
    dump(*res, ST_BNB_SIFT, cl, want_dumps);
    t0 = now();
    // S8 (serial, :967-1037 with shift=true, cluster=false)
    for (int i = 0; i < nL; ++i)
        for (auto &c : cl[i].c) { c.e = shift_to_line(c.e, &res->lines[3 * (size_t)i], p); c.ridx = -1; }
    t1 = now(); res->t_stage[ST_SHIFT] = t1 - t0;
    dump(*res, ST_SHIFT, cl, want_dumps);
    // S9 (omp, :1290-1358)
    Img ILu = mk(Lund, false, true), IRu = mk(Rund, false, true);
    std::vector<float> gxR, gyR;
    sobel(IRu.f, H, W, gxR, gyR);
    t0 = now();
    long gnp = 0, gni = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : gnp, gni)
    for (int i = 0; i < nL; ++i) {
        Cl &c = cl[i];
        if (c.c.empty()) continue;
        c.score.clear(); c.conf.clear(); c.valid.clear(); c.iters.clear();
        const double *l = &res->lines[3 * (size_t)i];
        double dx = -l[1], dy = l[0], nn = std::sqrt(dx * dx + dy * dy);
        dx /= nn; dy /= nn;
        for (auto &k : c.c) {
            double a, s, cf; bool v; int iters;
            gn_refine(L[i], k.e, dx, dy, ILu, IRu, gxR, gyR, p, a, s, cf, v, iters);
            c.score.push_back(s); c.conf.push_back(cf); c.valid.push_back(v); c.iters.push_back(iters);
            k.e.x += a * dx; k.e.y += a * dy;
            gnp++; gni += iters;
        }
    }
    res->gn_pairs = gnp; res->gn_iters = gni;
    for (int i = 0; i < nL; ++i) for (int it : cl[i].iters) res->gn_iter_list.push_back(it);
    t1 = now(); res->t_stage[ST_GN] = t1 - t0;
    dump(*res, ST_GN, cl, want_dumps);
    t0 = now();
    // S10 (serial, second shift + clustering by orientation)
    for (int i = 0; i < nL; ++i) {
        Cl &c = cl[i];
        if (c.c.empty()) continue;
        std::vector<E> sh;
        for (auto &k : c.c) sh.push_back(shift_to_line(k.e, &res->lines[3 * (size_t)i], p));
        std::vector<E> cen; std::vector<int> lab;
        cluster_edges(sh, true, p, cen, lab);
        c.c.clear();
        for (auto &e : cen) c.c.push_back(Cand{e, -1});  // score/conf/valid keep their old lengths (A.8)
    }
    t1 = now(); res->t_stage[ST_CLUSTER] = t1 - t0;
    dump(*res, ST_CLUSTER, cl, want_dumps);
    t0 = now();
    ncc_filter(cl, L, ILraw, IRraw, p, res->ncc_pairs2, lpatch);
    t1 = now(); res->t_stage[ST_NCC2] = t1 - t0;
    dump(*res, ST_NCC2, cl, want_dumps);
    t0 = now();
    // S12 (serial, :916-965)
    for (int i = 0; i < nL; ++i) {
        Cl &c = cl[i];
        if (c.c.empty()) continue;
        int best = 0; double mx = -1.0;
        for (int j = 0; j < (int)c.c.size(); ++j) if (c.score[j] > mx) { mx = c.score[j]; best = j; }
        Cand kc = c.c[best]; double ks = c.score[best], kf = c.conf[best]; char kv = c.valid[best];
        c.c = {kc}; c.score = {ks}; c.conf = {kf}; c.valid = {kv};
    }
    dump(*res, ST_BEST, cl, want_dumps);
    // S13
    for (int i = 0; i < nL; ++i)
        if (!cl[i].c.empty()) {
            res->mate_left.push_back(i);
            res->mate_rx.push_back(cl[i].c[0].e.x); res->mate_ry.push_back(cl[i].c[0].e.y); res->mate_rth.push_back(cl[i].c[0].e.th);
            res->mate_score.push_back(cl[i].score[0]);
        }
    t1 = now(); res->t_stage[ST_BEST] = t1 - t0;
    return res;
}

int so_num_mates(void *h) { return (int)((Result *)h)->mate_left.size(); }
void so_get_mates(void *h, int *left, double *rx, double *ry, double *rth, double *score)
{
    Result *r = (Result *)h;
    size_t n = r->mate_left.size();
    std::memcpy(left, r->mate_left.data(), n * 4);
    std::memcpy(rx, r->mate_rx.data(), n * 8); std::memcpy(ry, r->mate_ry.data(), n * 8);
    std::memcpy(rth, r->mate_rth.data(), n * 8); std::memcpy(score, r->mate_score.data(), n * 8);
}
int so_stage_total(void *h, int st) { Result *r = (Result *)h; return r->stages[st].off.empty() ? -1 : r->stages[st].off.back(); }
void so_get_stage(void *h, int st, int *off, int *ridx, double *x, double *y, double *th, double *score)
{
    StageDump &d = ((Result *)h)->stages[st];
    std::memcpy(off, d.off.data(), d.off.size() * 4);
    size_t n = d.ridx.size();
    std::memcpy(ridx, d.ridx.data(), n * 4);
    std::memcpy(x, d.x.data(), n * 8); std::memcpy(y, d.y.data(), n * 8);
    std::memcpy(th, d.th.data(), n * 8); std::memcpy(score, d.score.data(), n * 8);
}
void so_get_lines(void *h, double *lines) { Result *r = (Result *)h; std::memcpy(lines, r->lines.data(), r->lines.size() * 8); }
void so_get_stats(void *h, double *t_stage /*ST_COUNT*/, long *counts /*5: s1_total,ncc1,ncc2,gn_pairs,gn_iters*/)
{
    Result *r = (Result *)h;
    for (int k = 0; k < ST_COUNT; ++k) t_stage[k] = r->t_stage[k];
    counts[0] = r->s1_total; counts[1] = r->ncc_pairs1; counts[2] = r->ncc_pairs2; counts[3] = r->gn_pairs; counts[4] = r->gn_iters;
}
void so_get_gn_iters(void *h, int *out) { Result *r = (Result *)h; std::memcpy(out, r->gn_iter_list.data(), r->gn_iter_list.size() * 4); }
void so_free(void *h) { delete (Result *)h; }
int so_stage_count() { return ST_COUNT; }

}  // extern "C"

// keyframe -> current-frame quad tracking (Temporal_Matches.cpp), same translation unit
#include "temporal_oracle.inl"
