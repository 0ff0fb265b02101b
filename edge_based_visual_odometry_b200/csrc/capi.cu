// C ABI of libebvo_b200.so (see include/ebvo_b200.h).  Host-side context, buffers, copies and launch order.
#include "ebvo_internal.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>

namespace ebvo {

int Prof::index_of(const char* name)
{
    for (size_t k = 0; k < names.size(); ++k) if (names[k] == name) return (int)k;
    names.push_back(name); ms.push_back(0.f); launches.push_back(0);
    return (int)names.size() - 1;
}
void Prof::begin(const char* name, cudaStream_t st)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, st);
    pending.push_back({a, b});
    pendingIdx.push_back(index_of(name));
}
void Prof::end(cudaStream_t st) { cudaEventRecord(pending.back().second, st); }
void Prof::collect()
{
    for (size_t k = 0; k < pending.size(); ++k) {
        float t = 0.f;
        cudaEventSynchronize(pending[k].second);
        cudaEventElapsedTime(&t, pending[k].first, pending[k].second);
        ms[pendingIdx[k]] += t; launches[pendingIdx[k]] += 1;
        cudaEventDestroy(pending[k].first); cudaEventDestroy(pending[k].second);
    }
    pending.clear(); pendingIdx.clear();
    cnames.clear();
    for (auto& s : names) cnames.push_back(s.c_str());
}
void Prof::reset()
{
    for (auto& e : pending) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    pending.clear(); pendingIdx.clear();
    std::fill(ms.begin(), ms.end(), 0.f); std::fill(launches.begin(), launches.end(), 0);
}

struct StageData {
    bool valid = false;
    std::vector<int> off, ridx;
    std::vector<double> x, y, th, score;
};

}  // namespace ebvo

using namespace ebvo;

struct ebvo_ctx {
    int device = 0, maxW = 0, maxH = 0, maxB = 0, E = 0, P = 0;
    ebvo_params params;
    DevParams dp;
    cudaStream_t st = nullptr;
    std::string err;
    DevBatch b;
    uint8_t* d_raw = nullptr; uint8_t* d_und = nullptr;   // d_und: 2 images, only for ebvo_stereo_match with distinct undistorted inputs
    size_t imgStrideMax = 0;
    ebvo_mate* d_out = nullptr;   // [maxB][E] compacted mates
    float *d_descL = nullptr, *d_descR = nullptr; size_t descCap = 0;
    std::vector<void*> allocs;
    Prof prof;
    bool dumpsEnabled = false;
    bool dumpAllocated = false;
    StageData stages[EBVO_STAGE_COUNT];
    int stageNL = 0;
    int curFrames = 0;
    std::vector<int> h_counts;
    // pipelined batch call: copy-in / copy-out streams and per-sub-batch events
    alignas(64) unsigned char tmap[128];          // CUtensorMap of the image array for the current geometry
    int tmapW = 0, tmapH = 0;
    cudaStream_t stIn = nullptr, stOut = nullptr, st2 = nullptr;   // st2: second compute stream (odd sub-batches)
    cudaStream_t st3 = nullptr, st4 = nullptr;                     // further compute streams of the pipelined batch call (created on first use)
    std::vector<cudaEvent_t> evIn, evDone;
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    long long tqCounters[8] = {0};   // work counters of the last quad-tracking call
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char buf[512];                                                                         \
            snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            ctx->err = buf;                                                                        \
            return e__ == cudaErrorMemoryAllocation ? EBVO_ERR_NOMEM : EBVO_ERR_CUDA;              \
        }                                                                                          \
    } while (0)

template <typename T>
cudaError_t dalloc(ebvo_ctx* ctx, T** p, size_t n)
{
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 256);
    if (e == cudaSuccess) { ctx->allocs.push_back(q); *p = (T*)q; }
    return e;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void set_devparams(ebvo_ctx* ctx)
{
    const ebvo_params& q = ctx->params;
    DevParams& d = ctx->dp;
    d.epi = q.epipolar_line_dist_thresh; d.maxdisp = q.max_disparity; d.orient_deg = q.orientation_thresh_deg;
    d.shift_mag = q.orthogonal_shift_mag; d.ncc_thresh = q.ncc_thresh; d.bnb_ncc = q.bnb_ncc; d.bnb_sift = q.bnb_sift;
    d.sift_thresh = q.sift_threshold; d.loc_pert = q.location_perturbation; d.tang_displ = q.epip_tangency_displ_thresh;
    d.orient_pert = q.orient_perturbation; d.clus_dist = q.cluster_dist_thresh;
    d.clus_orient_rad = q.cluster_orient_thresh_deg * (M_PI / 180.0);   // deg_to_rad, include/utility.h:293-297
    d.clus_sigma = q.cluster_orient_gauss_sigma; d.clus_max = q.max_cluster_size; d.gn_max_iter = q.gn_max_iter;
    d.gn_tol = q.gn_tol; d.gn_huber = q.gn_huber_delta; d.toed_mag_thresh = (float)q.toed_mag_thresh; d.toed_border = q.toed_border;
    d.gn_mode = q.gn_mode; d.sift_mode = q.sift_mode;
    d.clus_small = 48;
    if (const char* e = getenv("EBVO_CLUSTER_SMALL")) d.clus_small = std::min(48, std::max(1, atoi(e)));
}

// geometry-dependent fields of the device view
int configure(ebvo_ctx* ctx, int w, int h, int nFrames)
{
    if (w <= 0 || h <= 0 || w > ctx->maxW || h > ctx->maxH || w >= 32768 || h >= 32768) { ctx->err = "image size exceeds the context's max_w/max_h"; return EBVO_ERR_INVALID; }
    if (nFrames < 1 || nFrames > ctx->maxB) { ctx->err = "n_frames exceeds the context's max_batch"; return EBVO_ERR_INVALID; }
    DevBatch& b = ctx->b;
    b.W = w; b.H = h; b.pitch = (int)align_up(w, 16);
    b.W2 = 2 * w; b.H2 = 2 * h;
    b.tilesX = (w + TW - 1) / TW; b.tilesY = (h + TH - 1) / TH;
    b.maskPitch = 2 * b.tilesX; b.maskRows = 2 * TH * b.tilesY;
    b.nFrames = nFrames; b.nImages = 2 * nFrames;
    b.raw = ctx->d_raw; b.und = ctx->d_raw;
    b.descL = nullptr; b.descR = nullptr;
    b.siftDev = ctx->params.sift_mode == 1 ? 1 : 0;
    b.imgBase = 0; b.tmap = ctx->tmap;
    if (w != ctx->tmapW || h != ctx->tmapH) {
        const int r = make_toed_tensor_map(ctx->tmap, ctx->d_raw, w, h, b.pitch, b.imgStride, 2 * ctx->maxB);
        if (r) { ctx->err = "cuTensorMapEncodeTiled failed (" + std::to_string(r) + ")"; return EBVO_ERR_CUDA; }
        ctx->tmapW = w; ctx->tmapH = h;
    }
    b.dumps = 0;
    ctx->curFrames = nFrames;
    return EBVO_OK;
}

int upload_image(ebvo_ctx* ctx, uint8_t* base, int slot, const uint8_t* src, int stride)
{
    const DevBatch& b = ctx->b;
    CK(cudaMemcpy2DAsync(base + (size_t)slot * b.imgStride, b.pitch, src, stride, b.W, b.H, cudaMemcpyHostToDevice, ctx->st));
    return EBVO_OK;
}

// Capacity overflows are recorded per FRAME (DevBatch::errFlag[f]): a dense or repetitive frame that exhausts a fixed
// capacity fails alone, the other frames of the batch keep their results.  Returns EBVO_ERR_CAPACITY when any frame of
// [0, nFrames) overflowed; `failed` (optional, nFrames entries) receives the per-frame codes.
int check_err_flag(ebvo_ctx* ctx, int nFrames = 1, std::vector<int>* failed = nullptr)
{
    nFrames = std::max(1, std::min(nFrames, ctx->maxB));
    std::vector<int> flags((size_t)nFrames, 0);
    CK(cudaMemcpyAsync(flags.data(), ctx->b.errFlag, sizeof(int) * nFrames, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    int first = -1, count = 0;
    for (int f = 0; f < nFrames; ++f) if (flags[f]) { if (first < 0) first = f; ++count; }
    if (failed) *failed = flags;
    if (first >= 0) {
        cudaMemsetAsync(ctx->b.errFlag, 0, sizeof(int) * nFrames, ctx->st);
        static const char* what[] = {"", "edge capacity (max_edges) exceeded", "candidate pool exhausted", "more than 128 NCC survivors for one left edge",
                                     "more than 128 candidates entering the clusterer for one left edge"};
        const int flag = flags[first];
        ctx->err = std::string(flag == 99 ? "internal bounds assertion (EBVO_CHECKED build)" : "capacity: ") + what[flag > 0 && flag < 5 ? flag : 0] + " (frame " + std::to_string(first) + (count > 1 ? ", " + std::to_string(count) + " frames in all" : "") + ")";
        return EBVO_ERR_CAPACITY;
    }
    return EBVO_OK;
}

int ensure_dump_buffers(ebvo_ctx* ctx)
{
    if (ctx->dumpAllocated) return EBVO_OK;
    for (int k = 0; k < DUMP_COUNT; ++k) {
        DumpBuf& d = ctx->b.dump[k];
        CK(dalloc(ctx, &d.n, (size_t)ctx->E)); CK(dalloc(ctx, &d.ridx, (size_t)ctx->P));
        CK(dalloc(ctx, &d.x, (size_t)ctx->P)); CK(dalloc(ctx, &d.y, (size_t)ctx->P));
        CK(dalloc(ctx, &d.th, (size_t)ctx->P)); CK(dalloc(ctx, &d.score, (size_t)ctx->P));
    }
    ctx->dumpAllocated = true;
    return EBVO_OK;
}

// host exclusive scan of per-left-edge counts -> offsets on host and device
int scan_counts(ebvo_ctx* ctx, const int* d_counts, int nL, std::vector<int>& off, int** d_off)
{
    std::vector<int> cnt(nL);
    if (nL) CK(cudaMemcpyAsync(cnt.data(), d_counts, sizeof(int) * nL, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    off.assign(nL + 1, 0);
    for (int i = 0; i < nL; ++i) off[i + 1] = off[i] + cnt[i];
    CK(cudaMalloc(d_off, sizeof(int) * (nL + 1)));
    CK(cudaMemcpyAsync(*d_off, off.data(), sizeof(int) * (nL + 1), cudaMemcpyHostToDevice, ctx->st));
    return EBVO_OK;
}

int snapshot_stage(ebvo_ctx* ctx, int stage, int src, int nL)
{
    StageData& S = ctx->stages[stage];
    const int* d_counts = (src < 0 || src == DUMP_S8) ? ctx->b.ccount : ctx->b.dump[src].n;   // S8 keeps the S7 counts
    int* d_off = nullptr;
    int rc = scan_counts(ctx, d_counts, nL, S.off, &d_off);
    if (rc) return rc;
    const size_t tot = S.off[nL];
    S.ridx.assign(tot, -1); S.x.assign(tot, 0); S.y.assign(tot, 0); S.th.assign(tot, 0); S.score.assign(tot, 0);
    if (tot) {
        int* d_r; double *d_x, *d_y, *d_t, *d_s;
        CK(cudaMalloc(&d_r, tot * 4)); CK(cudaMalloc(&d_x, tot * 8)); CK(cudaMalloc(&d_y, tot * 8)); CK(cudaMalloc(&d_t, tot * 8)); CK(cudaMalloc(&d_s, tot * 8));
        launch_snapshot(ctx->b, src, d_off, d_r, d_x, d_y, d_t, d_s, ctx->st);
        CK(cudaMemcpyAsync(S.ridx.data(), d_r, tot * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(S.x.data(), d_x, tot * 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(S.y.data(), d_y, tot * 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(S.th.data(), d_t, tot * 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(S.score.data(), d_s, tot * 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        cudaFree(d_r); cudaFree(d_x); cudaFree(d_y); cudaFree(d_t); cudaFree(d_s);
    }
    cudaFree(d_off);
    S.valid = true;
    return EBVO_OK;
}

int gate_stage(ebvo_ctx* ctx, int stage, int mode, const double* F21, int nL, const std::vector<double>& rx, const std::vector<double>& ry,
               const std::vector<double>& rth)
{
    StageData& S = ctx->stages[stage];
    int* d_counts = nullptr;
    CK(cudaMalloc(&d_counts, sizeof(int) * (nL + 1)));
    launch_gate_count(ctx->b, ctx->dp, F21, mode, d_counts, ctx->st);
    int* d_off = nullptr;
    int rc = scan_counts(ctx, d_counts, nL, S.off, &d_off);
    if (rc) return rc;
    const size_t tot = S.off[nL];
    S.ridx.assign(tot, -1);
    if (tot) {
        int* d_r;
        CK(cudaMalloc(&d_r, tot * 4));
        launch_gate_fill(ctx->b, ctx->dp, F21, mode, d_off, d_r, ctx->st);
        CK(cudaMemcpyAsync(S.ridx.data(), d_r, tot * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        cudaFree(d_r);
    }
    cudaFree(d_off); cudaFree(d_counts);
    S.x.resize(tot); S.y.resize(tot); S.th.resize(tot);
    S.score.assign(tot, std::numeric_limits<double>::quiet_NaN());
    for (size_t k = 0; k < tot; ++k) { int r = S.ridx[k]; S.x[k] = rx[r]; S.y[k] = ry[r]; S.th[k] = rth[r]; }
    S.valid = true;
    return EBVO_OK;
}

int upload_edges(ebvo_ctx* ctx, int img, const ebvo_edge* e, int n, std::vector<double>* keepx = nullptr, std::vector<double>* keepy = nullptr,
                 std::vector<double>* keept = nullptr)
{
    if (n > ctx->E) { ctx->err = "edge list longer than max_edges"; return EBVO_ERR_CAPACITY; }
    std::vector<double> x(n), y(n), t(n);
    for (int k = 0; k < n; ++k) { x[k] = e[k].x; y[k] = e[k].y; t[k] = e[k].theta; }
    const DevBatch& b = ctx->b;
    if (n) {
        CK(cudaMemcpyAsync(b.ex + (size_t)img * b.E, x.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(b.ey + (size_t)img * b.E, y.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(b.eth + (size_t)img * b.E, t.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->st));
    }
    CK(cudaMemcpyAsync(b.nE + img, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    if (keepx) { *keepx = x; *keepy = y; *keept = t; }
    return EBVO_OK;
}

int download_edges(ebvo_ctx* ctx, int img, ebvo_edge* out, int cap, int* n_edges, int* n_total)
{
    const DevBatch& b = ctx->b;
    int n = 0, nt = 0;
    CK(cudaMemcpyAsync(&n, b.nE + img, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(&nt, b.nTot + img, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    if (n_edges) *n_edges = n;
    if (n_total) *n_total = nt;
    if (!out) return EBVO_OK;
    const int m = std::min(n, cap);
    std::vector<double> x(m), y(m), t(m);
    if (m) {
        CK(cudaMemcpyAsync(x.data(), b.ex + (size_t)img * b.E, 8 * (size_t)m, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(y.data(), b.ey + (size_t)img * b.E, 8 * (size_t)m, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(t.data(), b.eth + (size_t)img * b.E, 8 * (size_t)m, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
    }
    for (int k = 0; k < m; ++k) { out[k].x = x[k]; out[k].y = y[k]; out[k].theta = t[k]; out[k].index = k; out[k].frame_source = -1; }
    if (n > cap) { ctx->err = "output edge buffer too small"; return EBVO_ERR_CAPACITY; }
    return EBVO_OK;
}

int download_mates(ebvo_ctx* ctx, int nFrames, ebvo_mate* out, int cap, int* n_mates)
{
    const DevBatch& b = ctx->b;
    ctx->h_counts.resize(nFrames);
    CK(cudaMemcpyAsync(ctx->h_counts.data(), b.nMates, sizeof(int) * nFrames, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    int rc = EBVO_OK;
    for (int f = 0; f < nFrames; ++f) {
        int n = ctx->h_counts[f];
        if (n_mates) n_mates[f] = n;
        int m = std::min(n, cap);
        if (n > cap) { ctx->err = "output mate buffer too small"; rc = EBVO_ERR_CAPACITY; }
        if (out && m) CK(cudaMemcpyAsync(out + (size_t)f * cap, ctx->d_out + (size_t)f * b.E, sizeof(ebvo_mate) * (size_t)m, cudaMemcpyDeviceToHost, ctx->st));
    }
    CK(cudaStreamSynchronize(ctx->st));
    return rc;
}

}  // namespace

namespace {
struct TqScratch {     // device allocations of one call (stream-ordered; the driver's pool makes repeated calls cheap)
    cudaStream_t st; std::vector<void*> ptrs;
    template <typename T> cudaError_t get(T** p, size_t n) { cudaError_t e = cudaMallocAsync((void**)p, (n ? n : 1) * sizeof(T), st); if (e == cudaSuccess) ptrs.push_back(*p); return e; }
    ~TqScratch() { for (void* q : ptrs) cudaFreeAsync(q, st); cudaStreamSynchronize(st); }
};
}  // namespace

extern "C" {

int ebvo_params_default(ebvo_params* p)
{
    if (!p) return EBVO_ERR_INVALID;
    p->epipolar_line_dist_thresh = 0.5; p->max_disparity = 25.0; p->orientation_thresh_deg = 10.0; p->orthogonal_shift_mag = 5.0;
    p->ncc_thresh = 0.6; p->bnb_ncc = 0.9; p->bnb_sift = 0.4; p->sift_threshold = 500.0; p->location_perturbation = 0.4;
    p->epip_tangency_displ_thresh = 3.0; p->orient_perturbation = 0.174533; p->cluster_dist_thresh = 1.0;
    p->cluster_orient_thresh_deg = 20.0; p->cluster_orient_gauss_sigma = 2.0; p->max_cluster_size = 10; p->gn_max_iter = 20;
    p->gn_tol = 1e-3; p->gn_huber_delta = 3.0; p->toed_mag_thresh = 2.0; p->toed_border = 10; p->gn_mode = 0; p->sift_mode = 0;
    return EBVO_OK;
}

int ebvo_create(ebvo_ctx** out, int device, int max_w, int max_h, int max_batch, int max_edges, const ebvo_params* params)
{
    if (!out || max_w <= 0 || max_h <= 0 || max_batch <= 0 || max_edges <= 0) return EBVO_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return EBVO_ERR_NO_DEVICE;
    ebvo_ctx* ctx = new ebvo_ctx;
    ctx->device = device; ctx->maxW = max_w; ctx->maxH = max_h; ctx->maxB = max_batch;
    ctx->E = (int)align_up(max_edges, 1024);
    ctx->P = ctx->E * 8;
    if (params) ctx->params = *params; else ebvo_params_default(&ctx->params);
    set_devparams(ctx);
    *out = ctx;   // returned even on failure so that ebvo_last_error() can be read; caller destroys it
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking));
    {   // the stream-ordered scratch of the quad-tracking and finalisation calls (cudaMallocAsync) stays in the device's pool between
        // calls instead of going back to the driver at every synchronisation: after the first call the "allocations" are pointer bumps
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    {   // the quad-tracking call takes its scratch from the stream-ordered allocator: keep freed blocks in the pool between calls
        cudaMemPool_t mp; unsigned long long keep = ~0ull;
        if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    CK(cudaStreamCreateWithFlags(&ctx->st2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->evJoin, cudaEventDisableTiming));
    // per-device set-up (constant tables, opt-in shared memory, the TMA driver entry point): a device or driver that cannot grant
    // them is reported HERE, not as a launch failure in some later call
    cudaGetLastError();
    init_toed_device();
    init_match_device();
    if (ctx->params.sift_mode == 1) upload_sift_tables();
    CK(cudaGetLastError());
    DevBatch& b = ctx->b;
    memset(&b, 0, sizeof b);
    const int B = max_batch, nImg = 2 * B;
    b.E = ctx->E; b.P = ctx->P; b.NB = (ctx->E + 7) / 8;   // index blocks of 8 right edges (match.cu: EB)
    const int tilesX = (max_w + TW - 1) / TW, tilesY = (max_h + TH - 1) / TH;
    b.imgStride = align_up((size_t)align_up(max_w, 16) * max_h, 256);
    b.maskStride = (size_t)(2 * tilesX) * (2 * TH * tilesY);
    b.rowStride = align_up((size_t)2 * max_h + 2, 32);
    b.gStride = align_up((size_t)max_w * max_h, 64);
    CK(dalloc(ctx, &ctx->d_raw, b.imgStride * nImg));
    CK(dalloc(ctx, &ctx->d_und, b.imgStride * 2));
    CK(dalloc(ctx, &b.mask, b.maskStride * nImg));
    CK(dalloc(ctx, &b.rowcnt, b.rowStride * nImg));
    CK(dalloc(ctx, &b.rowoff, b.rowStride * nImg));
    CK(dalloc(ctx, &b.coords, (size_t)b.E * nImg));
    CK(dalloc(ctx, &b.ex, (size_t)b.E * nImg)); CK(dalloc(ctx, &b.ey, (size_t)b.E * nImg)); CK(dalloc(ctx, &b.eth, (size_t)b.E * nImg));
    CK(dalloc(ctx, &b.nE, (size_t)nImg)); CK(dalloc(ctx, &b.nTot, (size_t)nImg)); CK(dalloc(ctx, &b.nRej, (size_t)nImg));
    if (ctx->params.gn_mode == 2) CK(dalloc(ctx, &b.pk, b.gStride * B));
    else if (ctx->params.gn_mode == 1) CK(dalloc(ctx, &b.pk16, b.gStride * B));
    else CK(dalloc(ctx, &b.pkh, b.gStride * B));
    CK(dalloc(ctx, &b.npatch, (size_t)b.E * NPF * nImg)); CK(dalloc(ctx, &b.pflag, (size_t)b.E * nImg));
    CK(dalloc(ctx, &b.blk, (size_t)b.NB * B)); CK(dalloc(ctx, &b.pmax, (size_t)b.NB * B)); CK(dalloc(ctx, &b.smin, (size_t)b.NB * B));
    CK(dalloc(ctx, &b.lines, (size_t)b.E * 8 * B));
    CK(dalloc(ctx, &b.cstart, (size_t)b.E * B)); CK(dalloc(ctx, &b.ccount, (size_t)b.E * B));
    CK(dalloc(ctx, &b.poolUsed, (size_t)B));
    CK(dalloc(ctx, &b.c_ridx, (size_t)b.P * B));
    CK(dalloc(ctx, &b.c_x, (size_t)b.P * B)); CK(dalloc(ctx, &b.c_y, (size_t)b.P * B)); CK(dalloc(ctx, &b.c_th, (size_t)b.P * B));
    CK(dalloc(ctx, &b.c_score, (size_t)b.P * B)); CK(dalloc(ctx, &b.c_conf, (size_t)b.P * B));
    CK(dalloc(ctx, &b.c_owner, (size_t)b.P * B));
    CK(dalloc(ctx, &b.mates, (size_t)b.E * B)); CK(dalloc(ctx, &b.nMates, (size_t)B)); CK(dalloc(ctx, &b.mateFlag, (size_t)b.E * B));
    CK(dalloc(ctx, &b.errFlag, (size_t)B + 4)); CK(dalloc(ctx, &b.counters, (size_t)8 * B));
    CK(dalloc(ctx, &ctx->d_out, (size_t)b.E * B));
    if (ctx->params.sift_mode == 1) {
        b.blurStride = align_up((size_t)max_w * max_h, 64);
        CK(dalloc(ctx, &b.blur, b.blurStride * nImg));
        CK(dalloc(ctx, &b.desc8, (size_t)b.E * 256 * nImg));
    }
    b.YT = max_h + 3;
    CK(dalloc(ctx, &b.wcur, (size_t)4 * B));
    CK(dalloc(ctx, &b.ytab, (size_t)2 * b.YT * B));
    CK(cudaMemsetAsync(b.errFlag, 0, sizeof(int) * ((size_t)B + 4), ctx->st));
    CK(cudaMemsetAsync(b.nE, 0, sizeof(int) * nImg, ctx->st));
    CK(cudaMemsetAsync(b.nMates, 0, sizeof(int) * B, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return EBVO_OK;
}

void ebvo_destroy(ebvo_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->st) cudaStreamSynchronize(ctx->st);
    ctx->prof.reset();
    for (void* p : ctx->allocs) cudaFree(p);
    if (ctx->d_descL) cudaFree(ctx->d_descL);
    if (ctx->d_descR) cudaFree(ctx->d_descR);
    for (cudaEvent_t e : ctx->evIn) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->evDone) cudaEventDestroy(e);
    if (ctx->stIn) cudaStreamDestroy(ctx->stIn);
    if (ctx->stOut) cudaStreamDestroy(ctx->stOut);
    if (ctx->st2) cudaStreamDestroy(ctx->st2);
    if (ctx->st3) cudaStreamDestroy(ctx->st3);
    if (ctx->st4) cudaStreamDestroy(ctx->st4);
    if (ctx->evFork) cudaEventDestroy(ctx->evFork);
    if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
    if (ctx->st) cudaStreamDestroy(ctx->st);
    delete ctx;
}

const char* ebvo_last_error(const ebvo_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int ebvo_fundamental(const ebvo_calib* c, double F21[9], double F12[9])
{
    if (!c || !F21) return EBVO_ERR_INVALID;
    auto inv3 = [](const double* m, double* o) {
        double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
        o[0] = (m[4] * m[8] - m[5] * m[7]) / det; o[1] = (m[2] * m[7] - m[1] * m[8]) / det; o[2] = (m[1] * m[5] - m[2] * m[4]) / det;
        o[3] = (m[5] * m[6] - m[3] * m[8]) / det; o[4] = (m[0] * m[8] - m[2] * m[6]) / det; o[5] = (m[2] * m[3] - m[0] * m[5]) / det;
        o[6] = (m[3] * m[7] - m[4] * m[6]) / det; o[7] = (m[1] * m[6] - m[0] * m[7]) / det; o[8] = (m[0] * m[4] - m[1] * m[3]) / det;
    };
    auto mul = [](const double* a, const double* b, double* o) {
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += a[i * 3 + k] * b[k * 3 + j]; o[i * 3 + j] = s; }
    };
    auto tr = [](const double* a, double* o) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o[i * 3 + j] = a[j * 3 + i]; };
    auto skew = [](const double* t, double* o) { o[0] = 0; o[1] = -t[2]; o[2] = t[1]; o[3] = t[2]; o[4] = 0; o[5] = -t[0]; o[6] = -t[1]; o[7] = t[0]; o[8] = 0; };
    double Kli[9], Kri[9], KriT[9], KliT[9], S[9], SR[9], tmp[9];
    inv3(c->Kl, Kli); inv3(c->Kr, Kri); tr(Kri, KriT); tr(Kli, KliT);
    skew(c->T21, S); mul(S, c->R21, SR); mul(KriT, SR, tmp); mul(tmp, Kli, F21);   // Dataset.cpp:105
    if (F12) {
        double R12[9], T12[3];
        tr(c->R21, R12);                                                           // Dataset.cpp:107-108
        for (int i = 0; i < 3; ++i) T12[i] = -(R12[i * 3] * c->T21[0] + R12[i * 3 + 1] * c->T21[1] + R12[i * 3 + 2] * c->T21[2]);
        skew(T12, S); mul(S, R12, SR); mul(KliT, SR, tmp); mul(tmp, Kri, F12);     // Dataset.cpp:111
    }
    return EBVO_OK;
}

int ebvo_toed(ebvo_ctx* ctx, const uint8_t* img, int w, int h, int stride, ebvo_edge* out, int cap, int* n_edges, int* n_total)
{
    if (!ctx || !img) return EBVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int rc = configure(ctx, w, h, 1);
    if (rc) return rc;
    ctx->prof.reset();
    if ((rc = upload_image(ctx, ctx->d_raw, 0, img, stride))) return rc;
    launch_toed(ctx->b, ctx->dp, 1, ctx->st, &ctx->prof);
    CK(cudaGetLastError());
    if ((rc = check_err_flag(ctx))) return rc;
    ctx->prof.collect();
    return download_edges(ctx, 0, out, cap, n_edges, n_total);
}

static int run_match_with_dumps(ebvo_ctx* ctx, const double* F21, bool sift, int nL, const std::vector<double>& rx, const std::vector<double>& ry,
                                const std::vector<double>& rth)
{
    int rc = ensure_dump_buffers(ctx);
    if (rc) return rc;
    ctx->b.dumps = 1;
    for (auto& s : ctx->stages) s.valid = false;
    ctx->stageNL = nL;
    match_prologue(ctx->b, ctx->dp, F21, 1, ctx->st, &ctx->prof);
    if ((rc = gate_stage(ctx, EBVO_STAGE_EPI, 0, F21, nL, rx, ry, rth))) return rc;
    if ((rc = gate_stage(ctx, EBVO_STAGE_DISP, 1, F21, nL, rx, ry, rth))) return rc;
    if ((rc = gate_stage(ctx, EBVO_STAGE_ORIENT, 2, F21, nL, rx, ry, rth))) return rc;
    match_gate(ctx->b, ctx->dp, F21, 1, ctx->st, &ctx->prof);
    if (sift) {
        if (ctx->b.siftDev) launch_sift(ctx->b, 2, ctx->st, &ctx->prof);
        match_sift(ctx->b, ctx->dp, 1, ctx->st, &ctx->prof);
        // positions are not materialised before the NCC kernel: take them from the right-edge list
        if ((rc = snapshot_stage(ctx, EBVO_STAGE_SIFT, -1, nL))) return rc;
        StageData& S = ctx->stages[EBVO_STAGE_SIFT];
        for (size_t k = 0; k < S.ridx.size(); ++k) { int r = S.ridx[k]; S.x[k] = rx[r]; S.y[k] = ry[r]; S.th[k] = rth[r]; S.score[k] = std::numeric_limits<double>::quiet_NaN(); }
    } else ctx->stages[EBVO_STAGE_SIFT] = ctx->stages[EBVO_STAGE_ORIENT];
    match_ncc(ctx->b, ctx->dp, 1, sift, ctx->st, &ctx->prof);
    if ((rc = snapshot_stage(ctx, EBVO_STAGE_NCC, DUMP_S6, nL))) return rc;
    if ((rc = snapshot_stage(ctx, EBVO_STAGE_BNB_NCC, DUMP_S7, nL))) return rc;
    if ((rc = snapshot_stage(ctx, EBVO_STAGE_BNB_SIFT, -1, nL))) return rc;
    match_gn(ctx->b, ctx->dp, 1, ctx->st, &ctx->prof);
    if ((rc = snapshot_stage(ctx, EBVO_STAGE_SHIFT, DUMP_S8, nL))) return rc;
    if ((rc = snapshot_stage(ctx, EBVO_STAGE_GN, -1, nL))) return rc;
    match_cluster(ctx->b, ctx->dp, 1, ctx->st, &ctx->prof);
    if ((rc = snapshot_stage(ctx, EBVO_STAGE_CLUSTER, DUMP_S10, nL))) return rc;
    if ((rc = snapshot_stage(ctx, EBVO_STAGE_NCC2, DUMP_S11, nL))) return rc;
    if ((rc = snapshot_stage(ctx, EBVO_STAGE_BEST, -1, nL))) return rc;
    ctx->b.dumps = 0;
    return EBVO_OK;
}

static int stereo_match_core(ebvo_ctx* ctx, const ebvo_calib* calib, const uint8_t* L_raw, const uint8_t* R_raw, const uint8_t* L_und,
                      const uint8_t* R_und, int w, int h, int stride, const ebvo_edge* L, int nL, const ebvo_edge* R, int nR,
                      const float* descL, const float* descR)
{
    if (!ctx || !calib || !L_raw || !R_raw || (nL && !L) || (nR && !R) || nL < 0 || nR < 0) return EBVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int rc = configure(ctx, w, h, 1);
    if (rc) return rc;
    ctx->prof.reset();
    if (!L_und) L_und = L_raw;
    if (!R_und) R_und = R_raw;
    if ((rc = upload_image(ctx, ctx->d_raw, 0, L_raw, stride))) return rc;
    if ((rc = upload_image(ctx, ctx->d_raw, 1, R_raw, stride))) return rc;
    if (L_und != L_raw || R_und != R_raw) {
        if ((rc = upload_image(ctx, ctx->d_und, 0, L_und, stride))) return rc;
        if ((rc = upload_image(ctx, ctx->d_und, 1, R_und, stride))) return rc;
        ctx->b.und = ctx->d_und;
    }
    std::vector<double> rx, ry, rth;
    if ((rc = upload_edges(ctx, 0, L, nL))) return rc;
    if ((rc = upload_edges(ctx, 1, R, nR, &rx, &ry, &rth))) return rc;
    const bool injected = descL && descR;
    const bool sift = injected || ctx->params.sift_mode == 1;
    ctx->b.siftDev = (!injected && ctx->params.sift_mode == 1) ? 1 : 0;
    if (injected) {
        size_t need = (size_t)std::max(nL, nR) * 256;
        if (need > ctx->descCap) {
            if (ctx->d_descL) cudaFree(ctx->d_descL);
            if (ctx->d_descR) cudaFree(ctx->d_descR);
            CK(cudaMalloc(&ctx->d_descL, need * 4 + 256)); CK(cudaMalloc(&ctx->d_descR, need * 4 + 256));
            ctx->descCap = need;
        }
        if (nL) CK(cudaMemcpyAsync(ctx->d_descL, descL, (size_t)nL * 256 * 4, cudaMemcpyHostToDevice, ctx->st));
        if (nR) CK(cudaMemcpyAsync(ctx->d_descR, descR, (size_t)nR * 256 * 4, cudaMemcpyHostToDevice, ctx->st));
        ctx->b.descL = ctx->d_descL; ctx->b.descR = ctx->d_descR;
    }
    double F21[9];
    ebvo_fundamental(calib, F21, nullptr);
    if (ctx->dumpsEnabled) {
        if ((rc = run_match_with_dumps(ctx, F21, sift, nL, rx, ry, rth))) return rc;
    } else launch_match(ctx->b, ctx->dp, F21, 1, sift, ctx->st, &ctx->prof);
    launch_compact(ctx->b, 1, ctx->d_out, ctx->b.E, ctx->st, &ctx->prof);
    CK(cudaGetLastError());
    if ((rc = check_err_flag(ctx))) return rc;
    ctx->prof.collect();
    return EBVO_OK;
}


int ebvo_stereo_match(ebvo_ctx* ctx, const ebvo_calib* calib, const uint8_t* L_raw, const uint8_t* R_raw, const uint8_t* L_und,
                      const uint8_t* R_und, int w, int h, int stride, const ebvo_edge* L, int nL, const ebvo_edge* R, int nR,
                      const float* descL, const float* descR, ebvo_mate* out, int cap, int* n_mates)
{
    int rc = stereo_match_core(ctx, calib, L_raw, R_raw, L_und, R_und, w, h, stride, L, nL, R, nR, descL, descR);
    if (rc) return rc;
    return download_mates(ctx, 1, out, cap, n_mates);
}

int ebvo_stereo_match_full(ebvo_ctx* ctx, const ebvo_calib* calib, const uint8_t* L_raw, const uint8_t* R_raw, const uint8_t* L_und,
                           const uint8_t* R_und, int w, int h, int stride, const ebvo_edge* L, int nL, const ebvo_edge* R, int nR,
                           ebvo_mate* out, int cap, int* n_mates, float* l_plus49, float* l_minus49, float* r_plus49, float* r_minus49,
                           float* l_desc256, float* r_desc256)
{
    int rc = stereo_match_core(ctx, calib, L_raw, R_raw, L_und, R_und, w, h, stride, L, nL, R, nR, nullptr, nullptr);
    if (rc) return rc;
    int n = 0;
    if ((rc = download_mates(ctx, 1, out, cap, &n))) return rc;
    if (n_mates) *n_mates = n;
    n = std::min(n, cap);
    if (n == 0) return EBVO_OK;
    const bool wantDesc = l_desc256 || r_desc256;
    if (wantDesc && ctx->params.sift_mode != 1) { ctx->err = "descriptors need a context created with sift_mode = 1"; return EBVO_ERR_INVALID; }
    const DevBatch& b = ctx->b;
    // scratch: six coordinate arrays, four patch arrays, two descriptor arrays of n entries (stream-ordered, pooled by the driver)
    double* xy = nullptr; float* pt = nullptr; float* ds = nullptr;
    CK(cudaMallocAsync(&xy, sizeof(double) * 6 * (size_t)n, ctx->st));
    CK(cudaMallocAsync(&pt, sizeof(float) * 4 * 49 * (size_t)n, ctx->st));
    if (wantDesc) CK(cudaMallocAsync(&ds, sizeof(float) * 2 * 256 * (size_t)n, ctx->st));
    double *lx = xy, *ly = xy + n, *lt = xy + 2 * (size_t)n, *rx = xy + 3 * (size_t)n, *ry = xy + 4 * (size_t)n, *rt = xy + 5 * (size_t)n;
    launch_mates_to_edges(ctx->d_out, n, lx, ly, lt, rx, ry, rt, ctx->st);
    // left patches from the RAW left image (Stereo_Matches.cpp:570-576), right patches from the UNDISTORTED right image (:1580-1582, :1622)
    launch_edge_patches(ctx->d_raw, w, h, b.pitch, lx, ly, lt, n, ctx->dp.shift_mag, pt, pt + 49 * (size_t)n, ctx->st);
    launch_edge_patches(b.und + b.imgStride, w, h, b.pitch, rx, ry, rt, n, ctx->dp.shift_mag, pt + 98 * (size_t)n, pt + 147 * (size_t)n, ctx->st);
    if (wantDesc) {
        // left descriptor pairs: those of the matched left edges, already computed for the SIFT gate (:655-689)
        launch_gather_desc(b.desc8, ctx->d_out, n, ds, ctx->st);
        // right descriptor pairs at the mates (:1627-1635): the right view's edge slots are free now, so the mates take their place
        CK(cudaMemcpyAsync(b.ex + b.E, rx, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaMemcpyAsync(b.ey + b.E, ry, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaMemcpyAsync(b.eth + b.E, rt, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaMemcpyAsync(b.nE + 1, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->st));
        DevBatch v = b;
        v.ex += b.E; v.ey += b.E; v.eth += b.E; v.nE += 1; v.blur += b.blurStride; v.desc8 += (size_t)b.E * 256;
        launch_sift_desc(v, 1, ctx->st, &ctx->prof);
        launch_desc_to_float(v.desc8, n, ds + 256 * (size_t)n, ctx->st);
    }
    CK(cudaGetLastError());
    const size_t pb = sizeof(float) * 49 * (size_t)n;
    if (l_plus49) CK(cudaMemcpyAsync(l_plus49, pt, pb, cudaMemcpyDeviceToHost, ctx->st));
    if (l_minus49) CK(cudaMemcpyAsync(l_minus49, pt + 49 * (size_t)n, pb, cudaMemcpyDeviceToHost, ctx->st));
    if (r_plus49) CK(cudaMemcpyAsync(r_plus49, pt + 98 * (size_t)n, pb, cudaMemcpyDeviceToHost, ctx->st));
    if (r_minus49) CK(cudaMemcpyAsync(r_minus49, pt + 147 * (size_t)n, pb, cudaMemcpyDeviceToHost, ctx->st));
    if (l_desc256) CK(cudaMemcpyAsync(l_desc256, ds, sizeof(float) * 256 * (size_t)n, cudaMemcpyDeviceToHost, ctx->st));
    if (r_desc256) CK(cudaMemcpyAsync(r_desc256, ds + 256 * (size_t)n, sizeof(float) * 256 * (size_t)n, cudaMemcpyDeviceToHost, ctx->st));
    cudaFreeAsync(xy, ctx->st); cudaFreeAsync(pt, ctx->st);
    if (ds) cudaFreeAsync(ds, ctx->st);
    CK(cudaStreamSynchronize(ctx->st));
    ctx->prof.collect();
    return EBVO_OK;
}

// The same device view restricted to frames [f0, f0 + n): every per-image / per-frame base pointer advanced, so that
// the kernels (which index frames from 0) work on a sub-batch while copies of other sub-batches are in flight.
static DevBatch frame_view(const DevBatch& b, int f0, int n)
{
    DevBatch v = b;
    const size_t i0 = 2 * (size_t)f0, F0 = (size_t)f0, E = (size_t)b.E, P = (size_t)b.P;
    v.nFrames = n; v.nImages = 2 * n;
    v.imgBase = b.imgBase + 2 * f0;
    v.raw += i0 * b.imgStride; v.und += i0 * b.imgStride;
    v.mask += i0 * b.maskStride; v.rowcnt += i0 * b.rowStride; v.rowoff += i0 * b.rowStride;
    v.coords += i0 * E; v.ex += i0 * E; v.ey += i0 * E; v.eth += i0 * E; v.nE += i0; v.nTot += i0; v.nRej += i0;
    if (v.pkh) v.pkh += F0 * b.gStride;
    if (v.pk16) v.pk16 += F0 * b.gStride;
    if (v.pk) v.pk += F0 * b.gStride;
    v.npatch += i0 * E * NPF; v.pflag += i0 * E;
    v.blk += F0 * b.NB; v.pmax += F0 * b.NB; v.smin += F0 * b.NB; v.ytab += F0 * 2 * b.YT;
    v.lines += F0 * E * 8;
    v.cstart += F0 * E; v.ccount += F0 * E; v.poolUsed += F0; v.wcur += F0 * 4;
    v.c_ridx += F0 * P; v.c_x += F0 * P; v.c_y += F0 * P; v.c_th += F0 * P; v.c_score += F0 * P; v.c_conf += F0 * P; v.c_owner += F0 * P;
    v.mates += F0 * E; v.nMates += F0; v.mateFlag += F0 * E;
    v.counters += F0 * 8; v.errFlag += F0;
    if (v.blur) v.blur += i0 * b.blurStride;
    if (v.desc8) v.desc8 += i0 * E * 256;
    return v;
}

static int run_frames(ebvo_ctx* ctx, const ebvo_calib* calib, int nFrames, int do_match)
{
    double F21[9];
    ebvo_fundamental(calib, F21, nullptr);
    auto run = [&](const DevBatch& v, int f0, int n, cudaStream_t cs) {
        launch_toed(v, ctx->dp, 2 * n, cs, &ctx->prof);
        if (do_match) {
            launch_match(v, ctx->dp, F21, n, ctx->params.sift_mode == 1, cs, &ctx->prof);
            launch_compact(v, n, ctx->d_out + (size_t)f0 * ctx->b.E, ctx->b.E, cs, &ctx->prof);
        }
    };
    if (nFrames >= 32 && !ctx->prof.enabled) {
        // big batches: slices alternating between two compute streams forked from / joined into the context stream, so that the
        // tail of every kernel (last wave, slowest warp of a persistent kernel) is filled by the other stream's work.  With
        // per-kernel profiling enabled the batch stays on one stream so that kernel durations are exclusive.
        static const int NS_ENV = getenv("EBVO_SLICES") ? std::max(2, atoi(getenv("EBVO_SLICES"))) : 0;
        const int ns = std::min(NS_ENV ? NS_ENV : 2, nFrames / 8);
        CK(cudaEventRecord(ctx->evFork, ctx->st));
        CK(cudaStreamWaitEvent(ctx->st2, ctx->evFork, 0));
        for (int k = 0; k < ns; ++k) {
            const int f0 = (int)((long long)nFrames * k / ns), f1 = (int)((long long)nFrames * (k + 1) / ns);
            run(frame_view(ctx->b, f0, f1 - f0), f0, f1 - f0, (k & 1) ? ctx->st2 : ctx->st);
        }
        CK(cudaEventRecord(ctx->evJoin, ctx->st2));
        CK(cudaStreamWaitEvent(ctx->st, ctx->evJoin, 0));
    } else run(ctx->b, 0, nFrames, ctx->st);
    CK(cudaGetLastError());
    return EBVO_OK;
}

int ebvo_stereo_frame(ebvo_ctx* ctx, const ebvo_calib* calib, const uint8_t* L_img, const uint8_t* R_img, int w, int h, int stride,
                      ebvo_mate* out, int cap, int* n_mates, ebvo_edge* L_edges, int* nL, ebvo_edge* R_edges, int* nR, int edge_cap)
{
    if (!ctx || !calib || !L_img || !R_img) return EBVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int rc = configure(ctx, w, h, 1);
    if (rc) return rc;
    ctx->prof.reset();
    if ((rc = upload_image(ctx, ctx->d_raw, 0, L_img, stride))) return rc;
    if ((rc = upload_image(ctx, ctx->d_raw, 1, R_img, stride))) return rc;
    if ((rc = run_frames(ctx, calib, 1, 1))) return rc;
    if ((rc = check_err_flag(ctx))) return rc;
    ctx->prof.collect();
    if ((rc = download_mates(ctx, 1, out, cap, n_mates))) return rc;
    if (L_edges || nL) if ((rc = download_edges(ctx, 0, L_edges, edge_cap, nL, nullptr))) return rc;
    if (R_edges || nR) if ((rc = download_edges(ctx, 1, R_edges, edge_cap, nR, nullptr))) return rc;
    return EBVO_OK;
}

int ebvo_batch_upload(ebvo_ctx* ctx, int n_frames, const uint8_t* const* L_imgs, const uint8_t* const* R_imgs, int w, int h, int stride)
{
    if (!ctx || !L_imgs || !R_imgs) return EBVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int rc = configure(ctx, w, h, n_frames);
    if (rc) return rc;
    for (int f = 0; f < n_frames; ++f) {
        if ((rc = upload_image(ctx, ctx->d_raw, 2 * f, L_imgs[f], stride))) return rc;
        if ((rc = upload_image(ctx, ctx->d_raw, 2 * f + 1, R_imgs[f], stride))) return rc;
    }
    return EBVO_OK;
}

int ebvo_batch_run(ebvo_ctx* ctx, const ebvo_calib* calib, int do_match)
{
    if (!ctx || !calib || ctx->curFrames < 1) return EBVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    return run_frames(ctx, calib, ctx->curFrames, do_match);
}

int ebvo_batch_sync(ebvo_ctx* ctx)
{
    if (!ctx) return EBVO_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->st));
    int rc = check_err_flag(ctx, ctx->curFrames);
    ctx->prof.collect();
    return rc;
}

int ebvo_batch_download(ebvo_ctx* ctx, ebvo_mate* out, int cap, int* n_mates)
{
    if (!ctx) return EBVO_ERR_INVALID;
    return download_mates(ctx, ctx->curFrames, out, cap, n_mates);
}

int ebvo_batch_pack(ebvo_ctx* ctx, void* d_dst, long long cap_records, int* d_offsets, long long* total)
{
    if (!ctx || !d_dst || !d_offsets || ctx->curFrames < 1) return EBVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const int F = ctx->curFrames;
    launch_pack(ctx->d_out, ctx->b.E, ctx->b.nMates, F, (ebvo_mate*)d_dst, cap_records, d_offsets, ctx->st, &ctx->prof);
    CK(cudaGetLastError());
    int tot = 0;
    CK(cudaMemcpyAsync(&tot, d_offsets + F, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    if (total) *total = tot;
    if (tot > cap_records) { ctx->err = "packed mate buffer too small"; return EBVO_ERR_CAPACITY; }
    return EBVO_OK;
}

int ebvo_batch_counts(ebvo_ctx* ctx, int* nL, int* nR, int* n_mates, long long* stage_counts)
{
    if (!ctx) return EBVO_ERR_INVALID;
    const int F = ctx->curFrames;
    std::vector<int> nE(2 * F), nM(F);
    CK(cudaMemcpyAsync(nE.data(), ctx->b.nE, sizeof(int) * 2 * F, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(nM.data(), ctx->b.nMates, sizeof(int) * F, cudaMemcpyDeviceToHost, ctx->st));
    if (stage_counts) CK(cudaMemcpyAsync(stage_counts, ctx->b.counters, sizeof(long long) * 8 * F, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    for (int f = 0; f < F; ++f) { if (nL) nL[f] = nE[2 * f]; if (nR) nR[f] = nE[2 * f + 1]; if (n_mates) n_mates[f] = nM[f]; }
    return EBVO_OK;
}

// dense = false: frame f's mates at out + f * cap (ebvo_stereo_batch); dense = true: frame f's mates follow frame f - 1's, at most
// cap_records in all (ebvo_stereo_batch_packed)
static int stereo_batch_impl(ebvo_ctx* ctx, const ebvo_calib* calib, int n_frames, const uint8_t* const* L_imgs, const uint8_t* const* R_imgs, int w,
                             int h, int stride, ebvo_mate* out, int cap, int* n_mates, bool dense, long long cap_records, long long* total)
{
    if (!ctx || !calib || !L_imgs || !R_imgs) return EBVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    ctx->prof.reset();
    int rc = configure(ctx, w, h, n_frames);
    if (rc) return rc;
    // (SB = 16: swept 8 .. 80 at 160 frames, profiles/r02_sub_batch_sweep.txt; EBVO_SUB_BATCH overrides it)
    // Software pipeline over sub-batches of SB frames: the images of sub-batch k+1 are copied in and the mates of the oldest
    // sub-batch in flight copied out while the kernels of the others run (a copy-in, a copy-out and four compute streams, events in between).
    static const int SB_ENV = getenv("EBVO_SUB_BATCH") ? std::max(1, atoi(getenv("EBVO_SUB_BATCH"))) : 0;
    const int SB = SB_ENV ? SB_ENV : 16, nsb = (n_frames + SB - 1) / SB;
    if (!ctx->stIn) {
        CK(cudaStreamCreateWithFlags(&ctx->stIn, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&ctx->stOut, cudaStreamNonBlocking));
    }
    while ((int)ctx->evIn.size() < nsb) {
        cudaEvent_t a, b;
        CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        ctx->evIn.push_back(a); ctx->evDone.push_back(b);
    }
    const DevBatch& b = ctx->b;
    double F21[9];
    ebvo_fundamental(calib, F21, nullptr);
    ctx->h_counts.assign(n_frames, 0);
    auto upload = [&](int k) -> int {
        const int f0 = k * SB, f1 = std::min(n_frames, f0 + SB);
        for (int f = f0; f < f1; ++f) {
            CK(cudaMemcpy2DAsync(ctx->d_raw + (size_t)(2 * f) * b.imgStride, b.pitch, L_imgs[f], stride, b.W, b.H, cudaMemcpyHostToDevice, ctx->stIn));
            CK(cudaMemcpy2DAsync(ctx->d_raw + (size_t)(2 * f + 1) * b.imgStride, b.pitch, R_imgs[f], stride, b.W, b.H, cudaMemcpyHostToDevice, ctx->stIn));
        }
        CK(cudaEventRecord(ctx->evIn[k], ctx->stIn));
        return EBVO_OK;
    };
    int over = EBVO_OK;
    long long written = 0;
    auto download = [&](int k) -> int {
        const int f0 = k * SB, f1 = std::min(n_frames, f0 + SB);
        CK(cudaStreamWaitEvent(ctx->stOut, ctx->evDone[k], 0));
        CK(cudaMemcpyAsync(ctx->h_counts.data() + f0, b.nMates + f0, sizeof(int) * (f1 - f0), cudaMemcpyDeviceToHost, ctx->stOut));
        CK(cudaStreamSynchronize(ctx->stOut));
        for (int f = f0; f < f1; ++f) {
            const int n = ctx->h_counts[f];
            const int m = dense ? (int)std::max<long long>(0, std::min<long long>(n, cap_records - written)) : std::min(n, cap);
            if (n_mates) n_mates[f] = n;
            if (n > m) { ctx->err = "output mate buffer too small"; over = EBVO_ERR_CAPACITY; }
            ebvo_mate* dst = dense ? out + written : out + (size_t)f * cap;
            if (out && m) CK(cudaMemcpyAsync(dst, ctx->d_out + (size_t)f * b.E, sizeof(ebvo_mate) * (size_t)m, cudaMemcpyDeviceToHost, ctx->stOut));
            written += m;
        }
        return EBVO_OK;
    };
    // an error return must not leave copies or kernels of this call in flight on any of the four streams
    auto bail = [&](int r) {
        cudaStreamSynchronize(ctx->stIn); cudaStreamSynchronize(ctx->st); cudaStreamSynchronize(ctx->st2);
        if (ctx->st3) cudaStreamSynchronize(ctx->st3);
        if (ctx->st4) cudaStreamSynchronize(ctx->st4);
        cudaStreamSynchronize(ctx->stOut);
        return r;
    };
    // the copy-in stream must not overwrite images a previous call's kernels may still read: calls are synchronous at return
    // sub-batches in flight, one compute stream each (EBVO_INFLIGHT overrides: 2 .. 4)
    static const int NFL = getenv("EBVO_INFLIGHT") ? std::max(2, std::min(4, atoi(getenv("EBVO_INFLIGHT")))) : 4;
    if (NFL >= 3 && !ctx->st3) CK(cudaStreamCreateWithFlags(&ctx->st3, cudaStreamNonBlocking));
    if (NFL >= 4 && !ctx->st4) CK(cudaStreamCreateWithFlags(&ctx->st4, cudaStreamNonBlocking));
    cudaStream_t css[4] = {ctx->st, ctx->st2, ctx->st3, ctx->st4};
    if ((rc = upload(0))) return bail(rc);
    for (int k = 0; k < nsb; ++k) {
        if (k + 1 < nsb && (rc = upload(k + 1))) return bail(rc);
        const int f0 = k * SB, n = std::min(n_frames, f0 + SB) - f0;
        const DevBatch v = frame_view(b, f0, n);
        // sub-batches rotate over the compute streams: the tail of one sub-batch's kernels (a persistent kernel waits for
        // its slowest warp) is filled by the next ones' kernels; the sub-batches touch disjoint buffers.  Measured end to end, 160 frames:
        // two in flight 1054 frames/s, three 1069, four 1074 (the host blocks on the OLDEST one's counts before it enqueues the next)
        cudaStream_t cs = css[k % NFL];
        CK(cudaStreamWaitEvent(cs, ctx->evIn[k], 0));
        launch_toed(v, ctx->dp, 2 * n, cs, &ctx->prof);
        launch_match(v, ctx->dp, F21, n, ctx->params.sift_mode == 1, cs, &ctx->prof);
        launch_compact(v, n, ctx->d_out + (size_t)f0 * b.E, b.E, cs, &ctx->prof);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->evDone[k], cs));
        if (k >= NFL - 1 && (rc = download(k - (NFL - 1)))) return bail(rc);
    }
    for (int k = std::max(0, nsb - (NFL - 1)); k < nsb; ++k)
        if ((rc = download(k))) return bail(rc);
    CK(cudaStreamSynchronize(ctx->stOut));
    CK(cudaStreamSynchronize(ctx->st)); CK(cudaStreamSynchronize(ctx->st2));
    if (ctx->st3) CK(cudaStreamSynchronize(ctx->st3));
    if (ctx->st4) CK(cudaStreamSynchronize(ctx->st4));
    {   // a frame that exhausted a capacity fails ALONE: its count becomes -1, every other frame keeps its mates
        std::vector<int> failed;
        rc = check_err_flag(ctx, n_frames, &failed);
        if (rc && n_mates) for (int f = 0; f < n_frames; ++f) if (failed[f]) n_mates[f] = -1;
        if (rc) { ctx->prof.collect(); return rc; }
    }
    ctx->prof.collect();
    if (total) *total = written;
    return over;
}

int ebvo_stereo_batch(ebvo_ctx* ctx, const ebvo_calib* calib, int n_frames, const uint8_t* const* L_imgs, const uint8_t* const* R_imgs, int w,
                      int h, int stride, ebvo_mate* out, int cap, int* n_mates)
{
    return stereo_batch_impl(ctx, calib, n_frames, L_imgs, R_imgs, w, h, stride, out, cap, n_mates, false, 0, nullptr);
}

int ebvo_stereo_batch_packed(ebvo_ctx* ctx, const ebvo_calib* calib, int n_frames, const uint8_t* const* L_imgs, const uint8_t* const* R_imgs, int w,
                             int h, int stride, ebvo_mate* out, long long cap_records, int* n_mates, long long* total)
{
    if (!out || cap_records < 0) return EBVO_ERR_INVALID;
    return stereo_batch_impl(ctx, calib, n_frames, L_imgs, R_imgs, w, h, stride, out, 0, n_mates, true, cap_records, total);
}

int ebvo_stereo_batch_multi(ebvo_ctx* const* ctxs, int n_ctx, const ebvo_calib* calib, int n_frames, const uint8_t* const* L_imgs,
                            const uint8_t* const* R_imgs, int w, int h, int stride, ebvo_mate* out, int cap, int* n_mates)
{
    if (!ctxs || n_ctx < 1 || !calib || n_frames < 0 || (n_frames && (!L_imgs || !R_imgs))) return EBVO_ERR_INVALID;
    for (int g = 0; g < n_ctx; ++g) if (!ctxs[g]) return EBVO_ERR_INVALID;
    const int per = (n_frames + n_ctx - 1) / n_ctx;
    std::vector<int> rcs(n_ctx, EBVO_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < n_ctx; ++g) {
        const int f0 = std::min(g * per, n_frames), f1 = std::min(f0 + per, n_frames);
        if (f1 <= f0) continue;
        th.emplace_back([=, &rcs]() {
            rcs[g] = ebvo_stereo_batch(ctxs[g], calib, f1 - f0, L_imgs + f0, R_imgs + f0, w, h, stride, out ? out + (size_t)f0 * cap : nullptr, cap,
                                       n_mates ? n_mates + f0 : nullptr);
        });
    }
    for (auto& t : th) t.join();
    for (int rc : rcs) if (rc != EBVO_OK) return rc;
    return EBVO_OK;
}

int ebvo_undistort(ebvo_ctx* ctx, const uint8_t* img, int w, int h, int stride, const double K[9], const double dist[4], uint8_t* out, int out_stride)
{
    if (!ctx || !img || !K || !dist || !out) return EBVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int rc = configure(ctx, w, h, 1);
    if (rc) return rc;
    if ((rc = upload_image(ctx, ctx->d_raw, 0, img, stride))) return rc;
    const DevBatch& b = ctx->b;
    uint8_t* d_dst = ctx->d_raw + b.imgStride;      // image slot 1 of the context as the destination
    launch_undistort(ctx->d_raw, b.pitch, d_dst, b.pitch, w, h, K, dist, ctx->st);
    CK(cudaGetLastError());
    CK(cudaMemcpy2DAsync(out, out_stride, d_dst, b.pitch, w, h, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return EBVO_OK;
}

int ebvo_sift_descriptors(ebvo_ctx* ctx, const uint8_t* img, int w, int h, int stride, const ebvo_edge* edges, int n, float* out)
{
    if (!ctx || !img || (n && !edges) || n < 0 || (n && !out)) return EBVO_ERR_INVALID;
    if (ctx->params.sift_mode != 1) { ctx->err = "ebvo_sift_descriptors needs a context created with sift_mode = 1"; return EBVO_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    int rc = configure(ctx, w, h, 1);
    if (rc) return rc;
    if (n > ctx->E) { ctx->err = "more edges than the context's max_edges"; return EBVO_ERR_CAPACITY; }
    if ((rc = upload_image(ctx, ctx->d_raw, 0, img, stride))) return rc;
    std::vector<double> x(n), y(n), t(n);
    for (int k = 0; k < n; ++k) { x[k] = edges[k].x; y[k] = edges[k].y; t[k] = edges[k].theta; }
    const DevBatch& b = ctx->b;
    if (n) {
        CK(cudaMemcpyAsync(b.ex, x.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(b.ey, y.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(b.eth, t.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->st));
    }
    CK(cudaMemcpyAsync(b.nE, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->st));
    launch_sift(b, 1, ctx->st, nullptr);
    CK(cudaGetLastError());
    std::vector<uint8_t> d8((size_t)n * 256);
    if (n) CK(cudaMemcpyAsync(d8.data(), b.desc8, d8.size(), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    for (size_t k = 0; k < d8.size(); ++k) out[k] = (float)d8[k];
    return EBVO_OK;
}

int ebvo_edge_patches(ebvo_ctx* ctx, const uint8_t* img, int w, int h, int stride, const ebvo_edge* edges, int n, float* plus49, float* minus49)
{
    if (!ctx || !img || !edges || n < 0) return EBVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int rc = configure(ctx, w, h, 1);
    if (rc) return rc;
    if ((rc = upload_image(ctx, ctx->d_raw, 0, img, stride))) return rc;
    if ((rc = upload_edges(ctx, 0, edges, n))) return rc;
    if (n == 0) return EBVO_OK;
    float *dp, *dm;
    CK(cudaMalloc(&dp, (size_t)n * 49 * 4)); CK(cudaMalloc(&dm, (size_t)n * 49 * 4));
    launch_edge_patches(ctx->d_raw, w, h, ctx->b.pitch, ctx->b.ex, ctx->b.ey, ctx->b.eth, n, ctx->dp.shift_mag, dp, dm, ctx->st);
    CK(cudaMemcpyAsync(plus49, dp, (size_t)n * 49 * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(minus49, dm, (size_t)n * 49 * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    cudaFree(dp); cudaFree(dm);
    return EBVO_OK;
}

int ebvo_ncc_patch_pair(ebvo_ctx* ctx, const float* p1, const float* p2, int n_pairs, double* out)
{
    if (!ctx || !p1 || !p2 || !out || n_pairs < 0) return EBVO_ERR_INVALID;
    if (n_pairs == 0) return EBVO_OK;
    CK(cudaSetDevice(ctx->device));
    float *d1, *d2; double* dout;
    const size_t nb = (size_t)n_pairs * 49 * 4;
    CK(cudaMalloc(&d1, nb)); CK(cudaMalloc(&d2, nb)); CK(cudaMalloc(&dout, (size_t)n_pairs * 8));
    CK(cudaMemcpyAsync(d1, p1, nb, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(d2, p2, nb, cudaMemcpyHostToDevice, ctx->st));
    launch_ncc_pairs(d1, d2, n_pairs, dout, ctx->st);
    CK(cudaMemcpyAsync(out, dout, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    cudaFree(d1); cudaFree(d2); cudaFree(dout);
    return EBVO_OK;
}

int ebvo_cluster(ebvo_ctx* ctx, const ebvo_edge* edges, int n, int by_orientation, ebvo_edge* centers, int* labels, int* n_clusters)
{
    if (!ctx || !edges || n < 0 || !centers || !labels || !n_clusters) return EBVO_ERR_INVALID;
    if (n > 128) { ctx->err = "ebvo_cluster handles at most 128 edges per set"; return EBVO_ERR_CAPACITY; }
    *n_clusters = 0;
    if (n == 0) return EBVO_OK;
    CK(cudaSetDevice(ctx->device));
    std::vector<double> h(6 * (size_t)n);
    for (int k = 0; k < n; ++k) { h[k] = edges[k].x; h[n + k] = edges[k].y; h[2 * n + k] = edges[k].theta; }
    double* d; int* dl;
    CK(cudaMalloc(&d, 6 * (size_t)n * 8)); CK(cudaMalloc(&dl, ((size_t)n + 1) * 4));
    CK(cudaMemcpyAsync(d, h.data(), 3 * (size_t)n * 8, cudaMemcpyHostToDevice, ctx->st));
    launch_cluster_one(d, d + n, d + 2 * n, n, by_orientation, ctx->dp, d + 3 * n, d + 4 * n, d + 5 * n, dl, dl + n, ctx->st);
    std::vector<int> hl(n + 1);
    CK(cudaMemcpyAsync(h.data(), d, 6 * (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(hl.data(), dl, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    cudaFree(d); cudaFree(dl);
    const int nc = hl[n];
    for (int k = 0; k < nc; ++k) { centers[k].x = h[3 * n + k]; centers[k].y = h[4 * n + k]; centers[k].theta = h[5 * n + k]; centers[k].index = k; centers[k].frame_source = 0; }
    for (int k = 0; k < n; ++k) labels[k] = hl[k];
    *n_clusters = nc;
    return EBVO_OK;
}

// ------------------------------------------------------------------------------------------------------
// Quad tracking (Temporal_Matches.cpp:168-218): host side of ebvo_temporal_quads / ebvo_temporal_quads_stage
// ------------------------------------------------------------------------------------------------------

int ebvo_temporal_quads_stage(ebvo_ctx* ctx, const uint8_t* kf_Lraw, const uint8_t* kf_Lund, const uint8_t* kf_Rund,
                              const uint8_t* cf_Lraw, const uint8_t* cf_Lund, const uint8_t* cf_Rund, int w, int h, int stride,
                              const ebvo_mate* kf, int n_kf, const uint8_t* kf_mask, const ebvo_mate* cf, int n_cf,
                              const float* desc_kf_l, const float* desc_kf_r, const float* desc_cf_l, const float* desc_cf_r,
                              const ebvo_quad_params* qp, int stage, int* off, ebvo_quad* out, int cap, int* n_quads)
{
    if (!ctx) return EBVO_ERR_INVALID;
    if (!kf_Lraw || !kf_Lund || !kf_Rund || !cf_Lraw || !cf_Lund || !cf_Rund || w <= 0 || h <= 0 || stride < w || n_kf < 0 || n_cf < 0 ||
        (n_kf && !kf) || (n_cf && !cf) || stage < 0 || stage >= EBVO_TQ_COUNT || cap < 0 || (cap && !out) || !n_quads) {
        ctx->err = "ebvo_temporal_quads: invalid argument"; return EBVO_ERR_INVALID;
    }
    ebvo_quad_params q{15, 0, 30.0, 10.0, 0.8, 0.8, 200.0};
    const float* hdesc[4] = {desc_kf_l, desc_kf_r, desc_cf_l, desc_cf_r};
    const int ndesc = (desc_kf_l != nullptr) + (desc_kf_r != nullptr) + (desc_cf_l != nullptr) + (desc_cf_r != nullptr);
    if (ndesc != 0 && ndesc != 4) { ctx->err = "ebvo_temporal_quads: pass all four descriptor arrays or none"; return EBVO_ERR_INVALID; }
    const bool sift_on = ndesc == 4 && n_kf > 0 && n_cf > 0;
    if (qp) q = *qp;
    if (q.cell_size <= 0 || !(q.grid_radius >= 0)) { ctx->err = "ebvo_temporal_quads: invalid grid parameters"; return EBVO_ERR_INVALID; }
    *n_quads = 0;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->st;
    TqScratch S{st, {}};
    TqDev d{};
    d.W = w; d.H = h; d.pitch = (w + 15) & ~15;
    d.n_kf = n_kf; d.n_cf = n_cf;
    d.cell = q.cell_size; d.gw = (w + q.cell_size - 1) / q.cell_size; d.gh = (h + q.cell_size - 1) / q.cell_size;   // Dataset.h:33-34
    d.sr = (int)std::ceil(q.grid_radius / q.cell_size);                                                             // Dataset.h:96
    d.orient_deg = q.orient_deg; d.ncc_thresh = q.ncc_thresh; d.bnb_thresh = q.bnb_thresh; d.sift_thresh = q.sift_thresh;
    const int ncell = d.gw * d.gh;
    // images
    uint8_t* img[6];
    const uint8_t* himg[6] = {kf_Lraw, kf_Lund, kf_Rund, cf_Lraw, cf_Lund, cf_Rund};
    for (int k = 0; k < 6; ++k) {
        CK(S.get(&img[k], (size_t)d.pitch * h));
        CK(cudaMemcpy2DAsync(img[k], d.pitch, himg[k], stride, w, h, cudaMemcpyHostToDevice, st));
    }
    d.kfLraw = img[0]; d.kfLund = img[1]; d.kfRund = img[2]; d.cfLraw = img[3]; d.cfLund = img[4]; d.cfRund = img[5];
    // mates
    std::vector<double> hk(6 * (size_t)n_kf), hc(6 * (size_t)n_cf);
    for (int i = 0; i < n_kf; ++i) { double* m = &hk[6 * (size_t)i]; m[0] = kf[i].lx; m[1] = kf[i].ly; m[2] = kf[i].ltheta; m[3] = kf[i].rx; m[4] = kf[i].ry; m[5] = kf[i].rtheta; }
    for (int i = 0; i < n_cf; ++i) { double* m = &hc[6 * (size_t)i]; m[0] = cf[i].lx; m[1] = cf[i].ly; m[2] = cf[i].ltheta; m[3] = cf[i].rx; m[4] = cf[i].ry; m[5] = cf[i].rtheta; }
    double *dkf, *dcf; uint8_t* dmask = nullptr;
    CK(S.get(&dkf, hk.size())); CK(S.get(&dcf, hc.size()));
    if (n_kf) CK(cudaMemcpyAsync(dkf, hk.data(), hk.size() * 8, cudaMemcpyHostToDevice, st));
    if (n_cf) CK(cudaMemcpyAsync(dcf, hc.data(), hc.size() * 8, cudaMemcpyHostToDevice, st));
    if (kf_mask && n_kf) { CK(S.get(&dmask, (size_t)n_kf)); CK(cudaMemcpyAsync(dmask, kf_mask, (size_t)n_kf, cudaMemcpyHostToDevice, st)); }
    d.kf = dkf; d.cf = dcf; d.kf_mask = dmask;
    for (int k = 0; k < 4; ++k) d.desc[k] = nullptr;
    if (sift_on && stage >= EBVO_TQ_SIFT) for (int k = 0; k < 4; ++k) {
        float* dd; const size_t n = (size_t)(k < 2 ? n_kf : n_cf) * 256;
        CK(S.get(&dd, n)); CK(cudaMemcpyAsync(dd, hdesc[k], n * sizeof(float), cudaMemcpyHostToDevice, st));
        d.desc[k] = dd;
    }
    // grid, patches, pools
    CK(S.get(&d.cellCount, (size_t)ncell)); CK(S.get(&d.cellStart, (size_t)ncell + 1)); CK(S.get(&d.cellCursor, (size_t)ncell));
    CK(S.get(&d.cellList, (size_t)n_cf)); CK(S.get(&d.lcell, (size_t)n_cf)); CK(S.get(&d.rcx, (size_t)n_cf)); CK(S.get(&d.rcy, (size_t)n_cf));
    CK(S.get(&d.errFlag, 1)); CK(S.get(&d.counters, 8));
    CK(cudaMemsetAsync(d.errFlag, 0, sizeof(int), st));
    int *counts, *offs;
    CK(S.get(&counts, (size_t)n_kf)); CK(S.get(&offs, (size_t)n_kf + 1));
    tq_prepare(d, ctx->dp, st, &ctx->prof);
    std::vector<int> hoff((size_t)n_kf + 1, 0);
    int total = 0, rc = EBVO_OK;
    if (stage <= EBVO_TQ_ORIENT) {
        // large lists (every mate in the 5x5 cell block): count, scan, fill; the quads are the CF mates themselves
        tq_gate(d, stage, counts, nullptr, nullptr, st, &ctx->prof);
        tq_scan(counts, offs, n_kf, st, &ctx->prof);
        CK(cudaMemcpyAsync(hoff.data(), offs, hoff.size() * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        total = hoff[n_kf];
        int* dcfidx;
        CK(S.get(&dcfidx, (size_t)total));
        tq_gate(d, stage, nullptr, offs, dcfidx, st, &ctx->prof);
        std::vector<int> hcf((size_t)total);
        if (total) CK(cudaMemcpyAsync(hcf.data(), dcfidx, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ctx->tqCounters, d.counters, 8 * sizeof(long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (total > cap) { ctx->err = "output quad buffer too small"; rc = EBVO_ERR_CAPACITY; }
        else for (int i = 0; i < n_kf; ++i) for (int e = hoff[i]; e < hoff[i + 1]; ++e) {
            const ebvo_mate& m = cf[hcf[e]];
            out[e] = ebvo_quad{i, hcf[e], m.lx, m.ly, m.ltheta, m.rx, m.ry, m.rtheta, -1.0, -1.0, 900.0, 900.0, 1e6, 1e6, 0, 0};   // scores{-1, 900}, Dataset.h:325-326
        }
    } else {
        for (int k = 0; k < 4; ++k) { const size_t n = (k < 2) ? n_kf : n_cf; CK(S.get(&d.np[k], n * 98)); CK(S.get(&d.pf[k], n)); }
        CK(S.get(&d.pk16[0], (size_t)w * h)); CK(S.get(&d.pk16[1], (size_t)w * h));
        CK(S.get(&d.pkh[0], (size_t)w * h)); CK(S.get(&d.pkh[1], (size_t)w * h));
        d.gn_gather = ctx->params.gn_mode == 1 ? 1 : 0;      // gn_mode 1 selects the gather kernels of both stages
        const size_t pool = (size_t)n_kf * TQ_CAP;
        CK(S.get(&d.cnt, (size_t)n_kf)); CK(S.get(&d.cnt2, (size_t)n_kf));
        CK(S.get(&d.q_cf, pool)); CK(S.get(&d.q_valid, pool)); CK(S.get(&d.q_ncc, 2 * pool)); CK(S.get(&d.q_sift, 2 * pool)); CK(S.get(&d.q_sc, 2 * pool)); CK(S.get(&d.q_l, 3 * pool)); CK(S.get(&d.q_r, 3 * pool));
        if (stage == EBVO_TQ_CLUSTER) { CK(S.get(&d.r_cf, pool)); CK(S.get(&d.r_valid, pool)); CK(S.get(&d.r_ncc, 2 * pool)); CK(S.get(&d.r_sift, 2 * pool)); CK(S.get(&d.r_sc, 2 * pool)); CK(S.get(&d.r_l, 3 * pool)); CK(S.get(&d.r_r, 3 * pool)); }
        tq_patches(d, ctx->dp, st, &ctx->prof);
        tq_gate(d, stage >= EBVO_TQ_BNB_SIFT ? 5 : stage, nullptr, nullptr, nullptr, st, &ctx->prof);      // gate modes 2..5 = stages NCC..BNB_SIFT
        if (stage >= EBVO_TQ_GN) tq_gn(d, ctx->dp, st, &ctx->prof);
        if (stage == EBVO_TQ_CLUSTER) tq_cluster(d, ctx->dp, st, &ctx->prof);
        const int which = stage == EBVO_TQ_CLUSTER ? 2 : 1;
        tq_scan(which == 2 ? d.cnt2 : d.cnt, offs, n_kf, st, &ctx->prof);
        ebvo_quad* dout;
        CK(S.get(&dout, (size_t)cap));
        tq_gather(d, which, offs, dout, cap, st, &ctx->prof);
        int herr = 0;
        CK(cudaMemcpyAsync(hoff.data(), offs, hoff.size() * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&herr, d.errFlag, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ctx->tqCounters, d.counters, 8 * sizeof(long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        total = hoff[n_kf];
        if (herr) { ctx->err = "quad tracking: more than 128 quads of one keyframe mate passed the NCC gate"; rc = EBVO_ERR_CAPACITY; }
        else if (total > cap) { ctx->err = "output quad buffer too small"; rc = EBVO_ERR_CAPACITY; }
        else if (total) { CK(cudaMemcpyAsync(out, dout, (size_t)total * sizeof(ebvo_quad), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); }
    }
    if (ctx->prof.enabled) ctx->prof.collect();
    CK(cudaGetLastError());
    *n_quads = total;
    if (off) std::memcpy(off, hoff.data(), hoff.size() * 4);
    return rc;
}

int ebvo_temporal_quads(ebvo_ctx* ctx, const uint8_t* kf_Lraw, const uint8_t* kf_Lund, const uint8_t* kf_Rund,
                        const uint8_t* cf_Lraw, const uint8_t* cf_Lund, const uint8_t* cf_Rund, int w, int h, int stride,
                        const ebvo_mate* kf, int n_kf, const uint8_t* kf_mask, const ebvo_mate* cf, int n_cf,
                        const float* desc_kf_l, const float* desc_kf_r, const float* desc_cf_l, const float* desc_cf_r,
                        const ebvo_quad_params* qp, ebvo_quad* out, int cap, int* n_quads)
{
    return ebvo_temporal_quads_stage(ctx, kf_Lraw, kf_Lund, kf_Rund, cf_Lraw, cf_Lund, cf_Rund, w, h, stride, kf, n_kf, kf_mask, cf, n_cf,
                                     desc_kf_l, desc_kf_r, desc_cf_l, desc_cf_r, qp, EBVO_TQ_CLUSTER, nullptr, out, cap, n_quads);
}

int ebvo_temporal_counters(ebvo_ctx* ctx, long long* out8)
{
    if (!ctx || !out8) return EBVO_ERR_INVALID;
    std::memcpy(out8, ctx->tqCounters, sizeof ctx->tqCounters);
    return EBVO_OK;
}

int ebvo_sobel(ebvo_ctx* ctx, const uint8_t* img, int w, int h, int stride, float* gx, float* gy)
{
    if (!ctx || !img || !gx || !gy) return EBVO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int rc = configure(ctx, w, h, 1);
    if (rc) return rc;
    if ((rc = upload_image(ctx, ctx->d_raw, 1, img, stride))) return rc;   // slot 1 = right view of frame 0
    launch_sobel(ctx->b, 1, ctx->st, nullptr);
    const size_t npx = (size_t)w * h;
    if (ctx->b.pk) {
        std::vector<float> pk(npx * 4);
        CK(cudaMemcpyAsync(pk.data(), ctx->b.pk, pk.size() * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        for (size_t k = 0; k < npx; ++k) { gx[k] = pk[4 * k + 1]; gy[k] = pk[4 * k + 2]; }
    } else {
        std::vector<uint32_t> pk(npx * 2);
        CK(cudaMemcpyAsync(pk.data(), ctx->b.pk16 ? ctx->b.pk16 : ctx->b.pkh, pk.size() * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        auto h2f = [](uint16_t hbits) {   // IEEE half -> float (values here are k/8, |k| <= 1020: normal or zero)
            const uint32_t s = (hbits >> 15) & 1, e = (hbits >> 10) & 31, mnt = hbits & 1023;
            if (e == 0) return (s ? -1.f : 1.f) * (float)mnt * (1.f / 16777216.f);
            return (s ? -1.f : 1.f) * std::ldexp(1.f + (float)mnt / 1024.f, (int)e - 15);
        };
        for (size_t k = 0; k < npx; ++k) {
            if (ctx->b.pk16) { gx[k] = (float)(int16_t)(pk[2 * k] >> 16) * 0.125f; gy[k] = (float)(int16_t)(pk[2 * k + 1] & 0xffff) * 0.125f; }
            else { gx[k] = h2f((uint16_t)(pk[2 * k + 1] & 0xffff)); gy[k] = h2f((uint16_t)(pk[2 * k + 1] >> 16)); }
        }
    }
    return EBVO_OK;
}

int ebvo_set_stage_dumps(ebvo_ctx* ctx, int enable)
{
    if (!ctx) return EBVO_ERR_INVALID;
    ctx->dumpsEnabled = enable != 0;
    return EBVO_OK;
}
int ebvo_stage_size(ebvo_ctx* ctx, int stage, int* n_left, int* total)
{
    if (!ctx || stage < 0 || stage >= EBVO_STAGE_COUNT || !ctx->stages[stage].valid) return EBVO_ERR_INVALID;
    if (n_left) *n_left = ctx->stageNL;
    if (total) *total = ctx->stages[stage].off.back();
    return EBVO_OK;
}
int ebvo_stage_fetch(ebvo_ctx* ctx, int stage, int* offsets, int* ridx, double* x, double* y, double* theta, double* score)
{
    if (!ctx || stage < 0 || stage >= EBVO_STAGE_COUNT || !ctx->stages[stage].valid) return EBVO_ERR_INVALID;
    const StageData& S = ctx->stages[stage];
    const size_t n = S.ridx.size();
    if (offsets) memcpy(offsets, S.off.data(), S.off.size() * 4);
    if (ridx && n) memcpy(ridx, S.ridx.data(), n * 4);
    if (x && n) memcpy(x, S.x.data(), n * 8);
    if (y && n) memcpy(y, S.y.data(), n * 8);
    if (theta && n) memcpy(theta, S.th.data(), n * 8);
    if (score && n) memcpy(score, S.score.data(), n * 8);
    return EBVO_OK;
}

long long ebvo_launch_count(ebvo_ctx* ctx) { return ctx ? ctx->prof.launchCount : -1; }

int ebvo_set_profiling(ebvo_ctx* ctx, int enable)
{
    if (!ctx) return EBVO_ERR_INVALID;
    ctx->prof.enabled = enable != 0;
    ctx->prof.reset();
    return EBVO_OK;
}
int ebvo_get_kernel_times(ebvo_ctx* ctx, const char*** names, const float** ms, const int** launches, int* n)
{
    if (!ctx || !n) return EBVO_ERR_INVALID;
    ctx->prof.cnames.clear();
    for (auto& s : ctx->prof.names) ctx->prof.cnames.push_back(s.c_str());
    if (names) *names = ctx->prof.cnames.data();
    if (ms) *ms = ctx->prof.ms.data();
    if (launches) *launches = ctx->prof.launches.data();
    *n = (int)ctx->prof.names.size();
    return EBVO_OK;
}
void* ebvo_stream(ebvo_ctx* ctx) { return ctx ? (void*)ctx->st : nullptr; }

void* ebvo_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void ebvo_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
