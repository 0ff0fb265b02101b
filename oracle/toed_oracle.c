/*
 * TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement, in plain C / FP64, of the reference's
 * third-order edge detector.  It is the checker for the CUDA path; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.  The product
 * (edge_based_visual_odometry_b200/csrc) never links or calls anything in this directory.
 *
 * Pinned against the UNMODIFIED reference compiled in place (oracle/_ref/libtoed_ref.so,
 * see oracle/Makefile and tests/test_oracle_toed.py): identical edge count and order,
 * |dx|,|dy|,|dtheta| < 1e-9 on every synthetic shape tested.
 *
 * Follows /root/reference/src/toed/cpu_toed.cpp:
 *   - preprocessing   (:82-120)   uint8 -> double
 *   - convolve_img    (:122-376)  nine Gaussian-derivative responses on the 2x grid.  The reference
 *       evaluates the 2-D sums directly; out-of-image taps are skipped (:204-205) = zero padding, and
 *       every 2-D kernel is an outer product of 1-D tables, so the sums are restated here in separable
 *       form (row pass, then column pass).  Tables = the closed forms quoted at :137-140,151-154 with
 *       sigma = TOED_SIGMA = 2 (definitions.h:77), s = p (unshifted) or p + 0.5 (shifted), p = -9..9;
 *       sub-grid (0,0) uses only the middle 17 entries (:200-207, index q+cent+1).
 *   - third-order orientation (:224-229)
 *   - non_maximum_suppresion (:386-515) octant NMS + parabola sub-pixel fit
 *   - compaction (:525-581) row-major, x=(X-1)/2, y=(Y-1)/2, 10 px border filter
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NT 19 /* shifted_kernel_sz = TOED_KERNEL_SIZE + 2 (cpu_toed.cpp:30-31) */
#define SIGMA 2.0
#define PI_ 3.14159265358979323846

static void make_tables(double delta, double *G, double *Gx, double *Gxx, double *Gxxx)
{
    const double s2 = SIGMA * SIGMA, c = sqrt(2.0 * PI_);
    for (int p = -9; p <= 9; ++p) {
        double s = p + delta, e = exp(-s * s / (2.0 * s2));
        G[p + 9] = e / (c * SIGMA);
        Gx[p + 9] = (-s * e) / (c * SIGMA * SIGMA * SIGMA);
        Gxx[p + 9] = ((s * s - s2) * e) / (c * SIGMA * SIGMA * SIGMA * SIGMA * SIGMA);
        Gxxx[p + 9] = ((s * (3.0 * s2 - s * s)) * e) / (c * SIGMA * SIGMA * SIGMA * SIGMA * SIGMA * SIGMA * SIGMA);
    }
}

/* out(i,j) = sum_{q=-R..R} in(i, j-q) * f[q+9]   (zero padding) */
static void row_filter(const double *in, double *out, int H, int W, const double *f, int R)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            double acc = 0;
            for (int q = -R; q <= R; ++q) {
                int jj = j - q;
                if (jj < 0 || jj >= W) continue;
                acc += in[(size_t)i * W + jj] * f[q + 9];
            }
            out[(size_t)i * W + j] = acc;
        }
}
/* out(i,j) = sum_{p=-R..R} in(i-p, j) * f[p+9] */
static void col_filter(const double *in, double *out, int H, int W, const double *f, int R)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            double acc = 0;
            for (int p = -R; p <= R; ++p) {
                int ii = i - p;
                if (ii < 0 || ii >= H) continue;
                acc += in[(size_t)ii * W + j] * f[p + 9];
            }
            out[(size_t)i * W + j] = acc;
        }
}

/*
 * maps (optional): 7 planes of 2H*2W doubles: Ix, Iy, mag, orient, subpix_x, subpix_y, subpix_mag.
 * edges_xyt: cap*3 (x, y, theta) in reference order; all4 (optional): cap_all*4 rows of
 * subpix_edge_pts_final.  Returns the number of border-filtered edges (toed_edges.size()).
 */
int toed_oracle(const uint8_t *img8, int H, int W, int stride, double *edges_xyt, int cap, int *n_total,
                double *maps, double *all4, int cap_all)
{
    const int H2 = 2 * H, W2 = 2 * W;
    const size_t N = (size_t)H * W, N2 = (size_t)H2 * W2;
    double T[2][4][NT]; /* [shifted?][G,Gx,Gxx,Gxxx][tap] */
    make_tables(0.0, T[0][0], T[0][1], T[0][2], T[0][3]);
    make_tables(0.5, T[1][0], T[1][1], T[1][2], T[1][3]);

    double *img = (double *)malloc(N * sizeof(double));
    double *rowf[4], *tmp = (double *)malloc(N * sizeof(double));
    for (int k = 0; k < 4; ++k) rowf[k] = (double *)malloc(N * sizeof(double));
    double *Ix = (double *)calloc(N2, sizeof(double)), *Iy = (double *)calloc(N2, sizeof(double));
    double *mag = (double *)calloc(N2, sizeof(double)), *ori = (double *)calloc(N2, sizeof(double));
    double *spx = (double *)calloc(N2, sizeof(double)), *spy = (double *)calloc(N2, sizeof(double));
    double *spm = (double *)calloc(N2, sizeof(double));
    double *resp[9];
    for (int k = 0; k < 9; ++k) resp[k] = (double *)malloc(N * sizeof(double));

    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) img[(size_t)i * W + j] = (double)img8[(size_t)i * stride + j];

    /* sub-grid (a,b): interp sample (2i+a, 2j+b); a shifts y, b shifts x (cpu_toed.cpp:243,287,331) */
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            const int R = (a == 0 && b == 0) ? 8 : 9;
            const double(*TX)[NT] = T[b], (*TY)[NT] = T[a];
            for (int k = 0; k < 4; ++k) row_filter(img, rowf[k], H, W, TX[k], R);
            /* response order: fx fy fxx fyy fxy fxxy fxyy fxxx fyyy ; (x-table, y-table) pairs :207-216 */
            const int xt[9] = {1, 0, 2, 0, 1, 2, 1, 3, 0};
            const int yt[9] = {0, 1, 0, 2, 1, 1, 2, 0, 3};
            for (int k = 0; k < 9; ++k) col_filter(rowf[xt[k]], resp[k], H, W, TY[yt[k]], R);
#pragma omp parallel for schedule(static)
            for (int i = 0; i < H; ++i)
                for (int j = 0; j < W; ++j) {
                    size_t s = (size_t)i * W + j, d = (size_t)(2 * i + a) * W2 + (2 * j + b);
                    double fx = resp[0][s], fy = resp[1][s], fxx = resp[2][s], fyy = resp[3][s], fxy = resp[4][s];
                    double fxxy = resp[5][s], fxyy = resp[6][s], fxxx = resp[7][s], fyyy = resp[8][s];
                    Ix[d] = fx;
                    Iy[d] = fy;
                    mag[d] = sqrt(fx * fx + fy * fy);
                    double tx = fx * (2 * fxx * fxx + 2 * fxy * fxy) + fy * (2 * fxx * fxy + 2 * fyy * fxy) + 2 * fx * fy * fxxy + fy * fy * fxyy + fx * fx * fxxx;
                    double ty = fx * (2 * fxx * fxy + 2 * fyy * fxy) + fy * (2 * fyy * fyy + 2 * fxy * fxy) + 2 * fx * fy * fxyy + fx * fx * fxxy + fy * fy * fyyy;
                    double tm = sqrt(tx * tx + ty * ty);
                    tx /= tm;
                    ty /= tm;
                    ori[d] = atan2(tx, -ty);
                }
        }

        /* NMS (cpu_toed.cpp:400-514) */
#define M(i, j) mag[(size_t)(i) * W2 + (j)]
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 10; i < H2 - 10; ++i)
        for (int j = 10; j < W2 - 10; ++j) {
            size_t d = (size_t)i * W2 + j;
            double m = mag[d], gx = Ix[d], gy = Iy[d];
            if (m <= 2) continue;
            if (fabs(gx) < 10e-6 && fabs(gy) < 10e-6) continue;
            double nx = gx / m, ny = gy / m, slope = 0, fp = 0, fm = 0;
            if (gx >= 0 && gy >= 0) {
                if (gx >= gy) { slope = ny / nx; fp = M(i, j + 1) * (1 - slope) + M(i + 1, j + 1) * slope; fm = M(i, j - 1) * (1 - slope) + M(i - 1, j - 1) * slope; }
                else { slope = nx / ny; fp = M(i + 1, j) * (1 - slope) + M(i + 1, j + 1) * slope; fm = M(i - 1, j) * (1 - slope) + M(i - 1, j - 1) * slope; }
            } else if (gx < 0 && gy >= 0) {
                if (fabs(gx) < gy) { slope = -nx / ny; fp = M(i + 1, j) * (1 - slope) + M(i + 1, j - 1) * slope; fm = M(i - 1, j) * (1 - slope) + M(i - 1, j + 1) * slope; }
                else { slope = -ny / nx; fp = M(i, j - 1) * (1 - slope) + M(i + 1, j - 1) * slope; fm = M(i, j + 1) * (1 - slope) + M(i - 1, j + 1) * slope; }
            } else if (gx < 0 && gy < 0) {
                if (fabs(gx) >= fabs(gy)) { slope = ny / nx; fp = M(i, j - 1) * (1 - slope) + M(i - 1, j - 1) * slope; fm = M(i, j + 1) * (1 - slope) + M(i + 1, j + 1) * slope; }
                else { slope = nx / ny; fp = M(i - 1, j) * (1 - slope) + M(i - 1, j - 1) * slope; fm = M(i + 1, j) * (1 - slope) + M(i + 1, j + 1) * slope; }
            } else if (gx >= 0 && gy < 0) {
                if (gx < fabs(gy)) { slope = -nx / ny; fp = M(i - 1, j) * (1 - slope) + M(i - 1, j + 1) * slope; fm = M(i + 1, j) * (1 - slope) + M(i + 1, j - 1) * slope; }
                else { slope = -ny / nx; fp = M(i, j + 1) * (1 - slope) + M(i - 1, j + 1) * slope; fm = M(i, j - 1) * (1 - slope) + M(i + 1, j - 1) * slope; }
            }
            double s = sqrt(1 + slope * slope);
            if ((m > fm && m > fp) || (m > fm && m >= fp) || (m >= fm && m > fp)) {
                double A = (fm + fp - 2 * m) / (2 * s * s), B = (fp - fm) / (2 * s), C = m;
                double ss = -B / (2 * A);
                double maxf = A * ss * ss + B * ss + C;
                if (fabs(ss) <= sqrt(2.0)) {
                    double sgx = maxf * nx, sgy = maxf * ny;
                    spx[d] = j + ss * nx;
                    spy[d] = i + ss * ny;
                    spm[d] = sqrt(sgx * sgx + sgy * sgy);
                }
            }
        }
#undef M

    /* compaction (cpu_toed.cpp:526-575) */
    int idx_all = 0, idx = 0;
    for (int i = 10; i < H2 - 10; ++i)
        for (int j = 10; j < W2 - 10; ++j) {
            size_t d = (size_t)i * W2 + j;
            if (spx[d] != 0) {
                double x = (spx[d] - 1) / 2, y = (spy[d] - 1) / 2, th = ori[d];
                if (all4 && idx_all < cap_all) {
                    all4[4 * (size_t)idx_all + 0] = x;
                    all4[4 * (size_t)idx_all + 1] = y;
                    all4[4 * (size_t)idx_all + 2] = th;
                    all4[4 * (size_t)idx_all + 3] = spm[d];
                }
                if (x > 10 && x < W - 10 && y > 10 && y < H - 10) {
                    if (idx < cap) {
                        edges_xyt[3 * (size_t)idx + 0] = x;
                        edges_xyt[3 * (size_t)idx + 1] = y;
                        edges_xyt[3 * (size_t)idx + 2] = th;
                    }
                    idx++;
                }
                idx_all++;
            }
        }
    if (n_total) *n_total = idx_all;

    if (maps) {
        double *src[7] = {Ix, Iy, mag, ori, spx, spy, spm};
        for (int k = 0; k < 7; ++k) memcpy(maps + (size_t)k * N2, src[k], N2 * sizeof(double));
    }
    free(img); free(tmp);
    for (int k = 0; k < 4; ++k) free(rowf[k]);
    for (int k = 0; k < 9; ++k) free(resp[k]);
    free(Ix); free(Iy); free(mag); free(ori); free(spx); free(spy); free(spm);
    return idx;
}
