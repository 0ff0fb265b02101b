// TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the reference's keyframe -> current-frame quad tracking,
// Temporal_Matches::get_Temporal_Edge_Pairs_from_Quads (/root/reference/src/Temporal_Matches.cpp:168-218), included at the
// end of stereo_oracle.cpp (same translation unit: it reuses the samplers, patches, NCC and clusterer restated there).
//
// Stage map, every stage keeps the candidate order of the reference:
//   TQ_GRID    add_edges_to_spatial_grid (:18-55) + apply_spatial_grid_filtering_quads (:335-383), SpatialGrid
//              (include/Dataset.h:22-114): CF mates whose LEFT edge lies in the 5x5 cell block around the KF left
//              edge's cell (cell 15 px, "radius" 30 px -> ceil(30/15) = 2 cells; no distance test) and whose RIGHT edge
//              lies in the block around the KF right edge's cell; order = cells row-major (dy, dx), ascending CF
//              index inside a cell (push_back order)
//   TQ_ORIENT  apply_orientation_filtering_quads (:385-414), 10 deg on both views (wrapped at 180)
//   TQ_NCC     apply_NCC_filtering_quads (:416-469): max of the 4 patch similarities > 0.8 on both views; KF/CF left
//              patches come from the RAW left images (Stereo_Matches.cpp:562,578), right patches from the UNDISTORTED
//              right images (:1580,1622)
//   TQ_SIFT    apply_SIFT_filtering_quads (:471-515): min of the 4 descriptor distances < 200 on both views (only when
//              the caller supplies descriptors; a pass-through otherwise)
//   TQ_BNB     apply_best_nearly_best_filtering_quads (:517-570) on the left NCC score, ratio 0.8
//   TQ_BNB_SIFT the same on the left SIFT distance, ratio 0.8 (pass-through without descriptors)
//   TQ_GN      apply_photometric_refinement_quads (:572-634) + min_Edge_Photometric_Residual_by_Gauss_Newton (:735-851):
//              2-D Gauss-Newton on both views (Huber 3, <= 20 iterations, 2x2 LDLT as Eigen does it)
//   TQ_CLUSTER apply_temporal_edge_clustering_quads (:636-733): EdgeClusterer by orientation on the refined left
//              edges, right centre = plain mean of the members' right edges
// Documented deviations: descriptors are not computed here (cv::SIFT is OpenCV code): without caller-supplied descriptor
// pairs the two SIFT stages are skipped ("SIFT-off", as in the stereo oracle); which KF mates take part (the reference takes those with a non-empty veridical_quads list, a
// ground-truth construct, :57-166) is an input mask.
//
// Pinned against the reference's own Temporal_Matches.cpp compiled in place (oracle/ref_temporal_harness.cpp ->
// oracle/_ref/libtemporal_ref.so) by tests/test_oracle_temporal.py.

namespace {

enum { TQ_GRID = 0, TQ_ORIENT, TQ_NCC, TQ_SIFT, TQ_BNB, TQ_BNB_SIFT, TQ_GN, TQ_CLUSTER, TQ_COUNT };

struct Mate { E l, r; };
struct Quad {
    int cf = -1;          // cf_stereo_edge_mate_index
    E l, r;               // Temporal_CF_Edge_Cluster.center_edge (left / right)
    double ncc_l = -1, ncc_r = -1;
    double sift_l = 900, sift_r = 900;   // scores{-1.0, 900.0} (:364)
    double sc_l = 1e6, sc_r = 1e6;   // refine_final_score (Dataset.h:325: 1e6 until the refinement runs)
    int valid = 0;               // refine_validity (valid_left && valid_right)
};
struct TResult {
    int n_kf = 0;
    std::vector<std::vector<Quad>> stage[TQ_COUNT];
    long gn_pairs = 0, gn_iters = 0;
    double t_stage[TQ_COUNT] = {0};
};

inline double wrapped_deg(double a, double b)   // Temporal_Matches.cpp:394-396
{
    double d = std::abs((a - b) * (180.0 / M_PI));
    if (d > 180.0) d = 360.0 - d;
    return d;
}

// Eigen::LDLT<Matrix2d>::compute + solve (Eigen 3.4 LDLT.h, lower triangle, pivot on the largest |diagonal|,
// first maximum wins): returns x with H x = b
inline void ldlt2_solve(double h00, double h10, double h11, double b0, double b1, double &x0, double &x1)
{
    const bool swap = std::abs(h11) > std::abs(h00);
    double d0 = swap ? h11 : h00, a11 = swap ? h00 : h11, l = h10;
    double y0 = swap ? b1 : b0, y1 = swap ? b0 : b1;
    double d1;
    if (std::abs(d0) > 0.0) { l = l / d0; d1 = a11 - l * (d0 * l); }
    else { d1 = a11; }   // whole diagonal zero: Eigen returns early with the matrix untouched
    y1 -= l * y0;                                            // L
    const double tol = std::numeric_limits<double>::min();   // D (pseudo-inverse)
    y0 = std::abs(d0) > tol ? y0 / d0 : 0.0;
    y1 = std::abs(d1) > tol ? y1 / d1 : 0.0;
    y0 -= l * y1;                                            // L^T
    x0 = swap ? y1 : y0; x1 = swap ? y0 : y1;
}

// Temporal_Matches.cpp:735-851
void gn_refine_2d(const E &kf, const E &cf, const Img &Ikf, const Img &Icf, const std::vector<float> &gx, const std::vector<float> &gy,
                  const P &p, double &dx_out, double &dy_out, double &score, bool &valid, int &iters)
{
    const double side = 7 / 2.0 + 1.0;
    double tx = std::cos(kf.th), ty = std::sin(kf.th);
    double nx = -ty, ny = tx;
    double Lp[49], Lm[49];
    sample_patch(Ikf.f.data(), Ikf.W, Ikf.H, kf.x + nx * side, kf.y + ny * side, kf.th, Lp);
    sample_patch(Ikf.f.data(), Ikf.W, Ikf.H, kf.x - nx * side, kf.y - ny * side, kf.th, Lm);
    double mLp = mean49(Lp), mLm = mean49(Lm);
    for (int k = 0; k < 49; ++k) { Lp[k] -= mLp; Lm[k] -= mLm; }
    double tcx = std::cos(cf.th), tcy = std::sin(cf.th);
    double ncx = -tcy, ncy = tcx;
    double d0 = kf.x - cf.x, d1 = kf.y - cf.y;      // init_disp (:602-603)
    score = 0; valid = false; iters = 0;
    int logn = 0;
    for (int it = 0; it < p.gn_max_iter; ++it) {
        const double lx = kf.x - d0, ly = kf.y - d1;
        double Rp[49], Rm[49], gxp[49], gxm[49], gyp[49], gym[49];
        const double cpx = lx + ncx * side, cpy = ly + ncy * side, cmx = lx - ncx * side, cmy = ly - ncy * side;
        sample_patch(Icf.f.data(), Icf.W, Icf.H, cpx, cpy, cf.th, Rp);
        sample_patch(Icf.f.data(), Icf.W, Icf.H, cmx, cmy, cf.th, Rm);
        sample_patch(gx.data(), Icf.W, Icf.H, cpx, cpy, cf.th, gxp);
        sample_patch(gx.data(), Icf.W, Icf.H, cmx, cmy, cf.th, gxm);
        sample_patch(gy.data(), Icf.W, Icf.H, cpx, cpy, cf.th, gyp);
        sample_patch(gy.data(), Icf.W, Icf.H, cmx, cmy, cf.th, gym);
        double mRp = mean49(Rp), mRm = mean49(Rm);
        double h00 = 0, h01 = 0, h10 = 0, h11 = 0, b0 = 0, b1 = 0, cost = 0;
        auto acc = [&](const double *Lc, const double *Rf, const double *gxf, const double *gyf, double mR) {
            for (int k = 0; k < 49; ++k) {
                double r = Lc[k] - (Rf[k] - mR);
                double ar = std::abs(r);
                double w = (ar < p.gn_huber) ? 1.0 : p.gn_huber / ar;     // strict <, :808
                double wjx = w * gxf[k], wjy = w * gyf[k];                 // (w * J) first, then * J^T (:810)
                h00 += wjx * gxf[k]; h01 += wjx * gyf[k]; h10 += wjy * gxf[k]; h11 += wjy * gyf[k];
                h00 += 1e-6; h11 += 1e-6;                                  // H += 1e-6 * Identity, per sample (:811)
                b0 += wjx * r; b1 += wjy * r;
                cost += w * r * r;
            }
        };
        acc(Lp, Rp, gxp, gyp, mRp);
        acc(Lm, Rm, gxm, gym, mRm);
        (void)h01;
        double s0, s1;
        ldlt2_solve(h00, h10, h11, b0, b1, s0, s1);
        const double e0 = -s0, e1 = -s1;
        d0 += e0; d1 += e1;
        double rms = std::sqrt(cost / 98.0);
        ++logn;
        iters = it + 1;
        bool outlier = (rms > p.gn_huber * 2.0) || (logn < 2);
        if (std::sqrt(e0 * e0 + e1 * e1) < p.gn_tol || it == p.gn_max_iter - 1) {
            valid = !outlier; score = rms;
            break;
        }
    }
    dx_out = d0; dy_out = d1;
}

}  // namespace

extern "C" {

// kf, cf: n x 6 doubles (left x, y, theta, right x, y, theta); kf_mask: n_kf bytes or NULL (= every KF mate)
void *to_run(const uint8_t *kfLraw, const uint8_t *kfLund, const uint8_t *kfRund, const uint8_t *cfLraw, const uint8_t *cfLund,
             const uint8_t *cfRund, int H, int W, const double *kf, int n_kf, const uint8_t *kf_mask, const double *cf, int n_cf,
             int cell_size, double grid_radius, double orient_deg, double ncc_thresh, double bnb_thresh,
             const float *descKfL, const float *descKfR, const float *descCfL, const float *descCfR, double sift_thresh)
{
    const bool sift_on = descKfL && descKfR && descCfL && descCfR;   // n x 2 x 128 floats per array (first, second descriptor)
    P p;
    TResult *res = new TResult;
    res->n_kf = n_kf;
    auto mk = [&](const uint8_t *u8) { Img I; I.H = H; I.W = W; I.d.resize((size_t)H * W); I.f.resize((size_t)H * W); for (size_t k = 0; k < (size_t)H * W; ++k) { I.d[k] = u8[k]; I.f[k] = u8[k]; } return I; };
    const Img IkfLraw = mk(kfLraw), IkfLund = mk(kfLund), IkfRund = mk(kfRund), IcfLraw = mk(cfLraw), IcfLund = mk(cfLund), IcfRund = mk(cfRund);
    std::vector<float> gxL, gyL, gxR, gyR;     // Pipeline.cpp:82-84: gradients of the undistorted current-frame views
    sobel(IcfLund.f, H, W, gxL, gyL);
    sobel(IcfRund.f, H, W, gxR, gyR);
    auto mates = [](const double *m, int n) { std::vector<Mate> v(n); for (int i = 0; i < n; ++i) { v[i].l = E{m[6 * i], m[6 * i + 1], m[6 * i + 2]}; v[i].r = E{m[6 * i + 3], m[6 * i + 4], m[6 * i + 5]}; } return v; };
    const std::vector<Mate> KF = mates(kf, n_kf), CF = mates(cf, n_cf);

    // ---- SpatialGrid(W, H, cell) + add_edges_to_spatial_grid ----
    const int gw = (W + cell_size - 1) / cell_size, gh = (H + cell_size - 1) / cell_size;
    std::vector<std::vector<int>> gridL((size_t)gw * gh), gridR((size_t)gw * gh);
    auto cell_of = [&](double x, double y, int &cx, int &cy) { cx = static_cast<int>(x) / cell_size; cy = static_cast<int>(y) / cell_size; return cx >= 0 && cx < gw && cy >= 0 && cy < gh; };
    for (int i = 0; i < n_cf; ++i) {
        int cx, cy;
        if (cell_of(CF[i].l.x, CF[i].l.y, cx, cy)) gridL[(size_t)cy * gw + cx].push_back(i);
        if (cell_of(CF[i].r.x, CF[i].r.y, cx, cy)) gridR[(size_t)cy * gw + cx].push_back(i);
    }
    auto within = [&](const std::vector<std::vector<int>> &g, double x, double y) {
        std::vector<int> out;
        const int gx0 = static_cast<int>(x) / cell_size, gy0 = static_cast<int>(y) / cell_size;
        const int sr = static_cast<int>(std::ceil(grid_radius / cell_size));
        for (int dy = -sr; dy <= sr; ++dy)
            for (int dx = -sr; dx <= sr; ++dx) {
                const int nx = gx0 + dx, ny = gy0 + dy;
                if (nx >= 0 && nx < gw && ny >= 0 && ny < gh) { const auto &c = g[(size_t)ny * gw + nx]; out.insert(out.end(), c.begin(), c.end()); }
            }
        return out;
    };

    // patches of every mate: left from the raw left image, right from the undistorted right image
    std::vector<float> pkfL((size_t)n_kf * 98), pkfR((size_t)n_kf * 98), pcfL((size_t)n_cf * 98), pcfR((size_t)n_cf * 98);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_kf; ++i) { edge_patches(KF[i].l, IkfLraw, p.shift_mag, &pkfL[(size_t)i * 98], &pkfL[(size_t)i * 98 + 49]); edge_patches(KF[i].r, IkfRund, p.shift_mag, &pkfR[(size_t)i * 98], &pkfR[(size_t)i * 98 + 49]); }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_cf; ++i) { edge_patches(CF[i].l, IcfLraw, p.shift_mag, &pcfL[(size_t)i * 98], &pcfL[(size_t)i * 98 + 49]); edge_patches(CF[i].r, IcfRund, p.shift_mag, &pcfR[(size_t)i * 98], &pcfR[(size_t)i * 98 + 49]); }
    auto sim4 = [&](const float *A, const float *B) {
        return max4(patch_similarity(A, B), patch_similarity(A, B + 49), patch_similarity(A + 49, B), patch_similarity(A + 49, B + 49));
    };

    for (int s = 0; s < TQ_COUNT; ++s) res->stage[s].assign(n_kf, {});
    long gnp = 0, gni = 0;
    double t0 = now();
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : gnp, gni)
    for (int i = 0; i < n_kf; ++i) {
        if (kf_mask && !kf_mask[i]) continue;
        // ---- TQ_GRID ----
        std::vector<Quad> q;
        {
            std::vector<int> lc = within(gridL, KF[i].l.x, KF[i].l.y), rc = within(gridR, KF[i].r.x, KF[i].r.y);
            std::unordered_set<int> rs(rc.begin(), rc.end());
            for (int c : lc) {
                if (!rs.count(c)) continue;
                Quad e; e.cf = c; e.l = CF[c].l; e.r = CF[c].r; e.ncc_l = -1; e.ncc_r = -1;
                q.push_back(e);
            }
        }
        res->stage[TQ_GRID][i] = q;
        // ---- TQ_ORIENT ----
        {
            std::vector<Quad> o;
            for (const Quad &e : q) {
                const double dl = wrapped_deg(KF[i].l.th, e.l.th), dr = wrapped_deg(KF[i].r.th, e.r.th);
                if ((dl < orient_deg || std::abs(dl - 180.0) < orient_deg) && (dr < orient_deg || std::abs(dr - 180.0) < orient_deg)) o.push_back(e);
            }
            q.swap(o);
        }
        res->stage[TQ_ORIENT][i] = q;
        // ---- TQ_NCC ----
        {
            std::vector<Quad> o;
            for (Quad e : q) {
                const double sl = sim4(&pkfL[(size_t)i * 98], &pcfL[(size_t)e.cf * 98]);
                const double sr = sim4(&pkfR[(size_t)i * 98], &pcfR[(size_t)e.cf * 98]);
                if (sl > ncc_thresh && sr > ncc_thresh) { e.ncc_l = sl; e.ncc_r = sr; o.push_back(e); }
            }
            q.swap(o);
        }
        res->stage[TQ_NCC][i] = q;
        // ---- TQ_SIFT ----
        if (sift_on) {
            auto l2 = [](const float *a, const float *b) { double t = 0; for (int k = 0; k < 128; ++k) { double d = (double)a[k] - (double)b[k]; t += d * d; } return std::sqrt(t); };
            auto min_sift = [&](const float *a, const float *b) { return std::min({l2(a, b), l2(a, b + 128), l2(a + 128, b), l2(a + 128, b + 128)}); };
            std::vector<Quad> o;
            for (Quad e : q) {
                const double sl = min_sift(descKfL + (size_t)i * 256, descCfL + (size_t)e.cf * 256);
                const double sr = min_sift(descKfR + (size_t)i * 256, descCfR + (size_t)e.cf * 256);
                if (sl < sift_thresh && sr < sift_thresh) { e.sift_l = sl; e.sift_r = sr; o.push_back(e); }
            }
            q.swap(o);
        }
        res->stage[TQ_SIFT][i] = q;
        // ---- TQ_BNB / TQ_BNB_SIFT (std::sort on the left score; ties keep index order here) ----
        for (int pass = 0; pass < 2; ++pass) {
            const bool is_ncc = pass == 0;
            if (q.size() >= 2 && (is_ncc || sift_on)) {
                std::vector<size_t> idx(q.size());
                std::iota(idx.begin(), idx.end(), 0);
                if (is_ncc) std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return q[a].ncc_l > q[b].ncc_l; });
                else std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return q[a].sift_l < q[b].sift_l; });
                const double best = is_ncc ? q[idx[0]].ncc_l : q[idx[0]].sift_l;
                size_t keep = 1;
                for (size_t j = 0; j + 1 < q.size(); ++j) {
                    if (best == 0) break;
                    const double next = is_ncc ? q[idx[j + 1]].ncc_l : q[idx[j + 1]].sift_l;
                    const double ratio = is_ncc ? next / best : best / next;
                    if (ratio >= bnb_thresh) ++keep; else break;
                }
                std::vector<Quad> o;
                for (size_t k = 0; k < keep; ++k) o.push_back(q[idx[k]]);
                q.swap(o);
            }
            res->stage[is_ncc ? TQ_BNB : TQ_BNB_SIFT][i] = q;
        }
        // ---- TQ_GN ----
        for (Quad &e : q) {
            double dlx, dly, drx, dry, sl, sr; bool vl, vr; int il, ir;
            gn_refine_2d(KF[i].l, CF[e.cf].l, IkfLund, IcfLund, gxL, gyL, p, dlx, dly, sl, vl, il);
            gn_refine_2d(KF[i].r, CF[e.cf].r, IkfRund, IcfRund, gxR, gyR, p, drx, dry, sr, vr, ir);
            gnp += 2; gni += il + ir;
            e.sc_l = sl; e.sc_r = sr; e.valid = (vl && vr) ? 1 : 0;
            if (vl) { e.l.x = KF[i].l.x - dlx; e.l.y = KF[i].l.y - dly; }
            if (vr) { e.r.x = KF[i].r.x - drx; e.r.y = KF[i].r.y - dry; }
        }
        res->stage[TQ_GN][i] = q;
        // ---- TQ_CLUSTER ----
        if (q.size() >= 2) {
            std::vector<E> sl(q.size());
            for (size_t k = 0; k < q.size(); ++k) sl[k] = q[k].l;
            std::vector<E> centers; std::vector<int> lab;
            cluster_edges(sl, true, p, centers, lab);
            std::vector<Quad> o;
            for (size_t c = 0; c < centers.size(); ++c) {
                int best_idx = -1; std::vector<int> sub;
                for (size_t m = 0; m < q.size(); ++m) {
                    if (lab[m] != (int)c) continue;
                    int ci = -1; double cd = std::numeric_limits<double>::max();   // closest shifted_left edge to this contributor (:670-680)
                    for (size_t k = 0; k < sl.size(); ++k) {
                        const double dx = sl[m].x - sl[k].x, dy = sl[m].y - sl[k].y;
                        const double d = std::sqrt(dx * dx + dy * dy);
                        if (d < cd) { cd = d; ci = (int)k; }
                    }
                    if (ci >= 0) { sub.push_back(ci); best_idx = ci; }
                }
                if (best_idx < 0 || sub.empty()) continue;
                E rc;
                if (sub.size() == 1) rc = q[sub[0]].r;
                else {
                    double sx = 0, sy = 0, st = 0;
                    for (int k : sub) { sx += q[k].r.x; sy += q[k].r.y; st += q[k].r.th; }
                    const int n = (int)sub.size();
                    rc = E{sx / n, sy / n, st / n};
                }
                Quad e = q[best_idx];
                e.l = centers[c]; e.r = rc; e.cf = q[best_idx].cf;
                o.push_back(e);
            }
            q.swap(o);
        }
        res->stage[TQ_CLUSTER][i] = q;
    }
    res->gn_pairs = gnp; res->gn_iters = gni;
    res->t_stage[0] = now() - t0;
    return res;
}

int to_stage_total(void *h, int st)
{
    TResult *r = (TResult *)h; long t = 0;
    for (auto &v : r->stage[st]) t += (long)v.size();
    return (int)t;
}
// off: n_kf + 1; per entry: cf, l[3], r[3], ncc[2], sc[2], valid, sift[2]
void to_get_stage(void *h, int st, int *off, int *cf, double *l, double *r, double *ncc, double *sc, int *valid, double *sift)
{
    TResult *R = (TResult *)h; int o = 0;
    for (int i = 0; i < R->n_kf; ++i) {
        off[i] = o;
        for (const Quad &e : R->stage[st][i]) {
            cf[o] = e.cf; l[3 * o] = e.l.x; l[3 * o + 1] = e.l.y; l[3 * o + 2] = e.l.th; r[3 * o] = e.r.x; r[3 * o + 1] = e.r.y; r[3 * o + 2] = e.r.th;
            ncc[2 * o] = e.ncc_l; ncc[2 * o + 1] = e.ncc_r; sift[2 * o] = e.sift_l; sift[2 * o + 1] = e.sift_r; sc[2 * o] = e.sc_l; sc[2 * o + 1] = e.sc_r; valid[o] = e.valid;
            ++o;
        }
    }
    off[R->n_kf] = o;
}
void to_get_stats(void *h, double *seconds, long *counts /*2: gn_pairs, gn_iters*/)
{
    TResult *r = (TResult *)h; *seconds = r->t_stage[0]; counts[0] = r->gn_pairs; counts[1] = r->gn_iters;
}
void to_free(void *h) { delete (TResult *)h; }
int to_stage_count() { return TQ_COUNT; }

}  // extern "C"
