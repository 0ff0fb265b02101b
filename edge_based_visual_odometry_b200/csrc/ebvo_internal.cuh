// Internal declarations shared by the translation units of libebvo_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/ebvo_b200.h"

namespace ebvo {

// ---- TOED tiling (K_A) ---------------------------------------------------------------------------------
constexpr int TW = 32;             // input pixels per tile, x
constexpr int TH = 32;             // input pixels per tile, y
constexpr int HALO = 10;           // 9 taps + 1 pixel for the NMS ring
constexpr int IN_W = TW + 2 * HALO;  // 52
constexpr int IN_H = TH + 2 * HALO;  // 52
constexpr int OW = TW + 2;         // conv output columns kept per tile (x0-1 .. x0+TW)
constexpr int OH = TH + 2;
constexpr int IW = 2 * OW;         // interp samples per tile incl. ring: 68
constexpr int IH = 2 * OH;
constexpr int TOED_THREADS = 256;
constexpr int NPF = 104;           // floats per edge in DevBatch::npatch: "+" cells at 0..48, "-" cells at 52..100, zero padding (16-byte rows)

// Debug stage dumps (frame 0 only): in-kernel snapshots of lists that do not survive the fused kernels.
// Every buffer is addressed with the left edge's pool segment offset (cstart[i]); n[i] = entries of edge i.
struct DumpBuf { int* n; int* ridx; double *x, *y, *th, *score; };
enum { DUMP_S6 = 0, DUMP_S7, DUMP_S8, DUMP_S10, DUMP_S11, DUMP_COUNT };

// Device view of one batch: base pointers + per-image / per-frame strides (elements of the pointed type).
// Image index = 2*frame + view (view 0 = left, 1 = right).
struct DevBatch {
    int W, H, pitch;       // image size; pitch in bytes of the device copies
    int W2, H2;            // interp grid
    int tilesX, tilesY;
    int maskPitch;         // uint32 words per interp row (= 2*tilesX)
    int maskRows;          // rows in the mask (= 64*tilesY)
    int E;                 // edge capacity per image
    int P;                 // candidate pool capacity per frame
    int NB;                // edge blocks per image (= E/32)
    int nImages, nFrames;
    int imgBase;           // index of this view's first image in the context's image array (sub-batch views; TMA coordinate)
    const void* tmap;      // host pointer to the 128-byte CUtensorMap of the image array (toed.cu)
    const uint8_t* raw; const uint8_t* und; size_t imgStride;
    uint32_t* mask; size_t maskStride;
    int* rowcnt; int* rowoff; size_t rowStride;
    uint32_t* coords;                     // [img][E] packed (i<<16 | j)
    double *ex, *ey, *eth;                // [img][E]
    int *nE, *nTot;                       // [img]
    int* nRej;                            // [img] prefilter survivors the FP64 tests of toed_refine rejected
    // right-view image packed with its Sobel/8 gradients, one of (per ebvo_params.gn_mode), [frame][H*W]:
    uint2* pkh;    // {u16 I, half gx, half gy, 0}   default mixed-precision GN kernel
    uint2* pk16;   // {I, 8gx, 8gy} as int16           FP64 GN kernel
    float4* pk;    // {I, gx, gy, 0} floats            FP32 GN kernel
    size_t gStride;
    float* npatch; uint8_t* pflag;        // normalised NCC patches [img][E][NPF] + flat flags [img][E]
    float4* blk; float* pmax; float* smin;  // right-edge block bounds [frame][NB]
    double* lines;                        // [frame][E][8]: a, b, c, dirx, diry, sin(thL), cos(thL), pad
    int *cstart, *ccount;                 // [frame][E]
    int* poolUsed;                        // [frame]
    int* wcur;                            // [frame][4] work cursors of the list-driven clusterer launches (zeroed per run)
    int* c_ridx; double *c_x, *c_y, *c_th, *c_score, *c_conf;  // [frame][P]
    int* c_owner;                         // [frame][P] left-edge index of a live pool slot, -1 for dead slots
    ebvo_mate* mates; int* nMates;        // [frame][E], [frame]
    int* mateFlag;                        // [frame][E]
    int* errFlag;                         // [frame] capacity overflows (0 = none): only the frame that overflowed is reported as failed
    unsigned long long* counters;         // [frame][8] work counters (s3 pairs, ncc pairs, gn pairs, gn iters, ncc2 pairs ...)
    const float* descL; const float* descR; // optional caller-supplied SIFT descriptors (frame 0 only), may be null
    float* blur; size_t blurStride;         // sift_mode 1: Gaussian-blurred float images [img][H*W] (descriptor image of cv::SIFT)
    uint8_t* desc8;                         // sift_mode 1: descriptors computed on the device [img][E][2][128]
    int siftDev;                            // != 0: use desc8 (all frames) instead of descL/descR
    int* ytab; int YT;                    // [frame][2][YT] first index block per integer image row (bounds_kernel -> gate_kernel), YT = H + 3
    int dumps;                            // != 0: fill dump[] for frame 0
    DumpBuf dump[DUMP_COUNT];
};

struct DevParams {
    double epi, maxdisp, orient_deg, shift_mag, ncc_thresh, bnb_ncc, bnb_sift, sift_thresh;
    double loc_pert, tang_displ, orient_pert, clus_dist, clus_orient_rad, clus_sigma;
    int clus_max, gn_max_iter;
    int clus_small;   // sets up to this size go to the small-capacity clusterer launch (48; EBVO_CLUSTER_SMALL lowers it for tests)
    double gn_tol, gn_huber;
    float toed_mag_thresh; int toed_border;
    int gn_mode;   // 0 FP64 tiled (default), 1 FP64 gather, 2 FP32
    int sift_mode; // 0: SIFT stages only with caller-supplied descriptors; 1: descriptors computed on the device
};

// kernel launchers (defined in toed.cu / match.cu); all asynchronous on `st`
void launch_toed(const DevBatch& b, const DevParams& p, int nImages, cudaStream_t st, struct Prof* prof);
void launch_match(const DevBatch& b, const DevParams& p, const double* F21 /*host 9*/, int nFrames, bool sift, cudaStream_t st, struct Prof* prof);
void launch_sobel(const DevBatch& b, int nFrames, cudaStream_t st, struct Prof* prof);
void launch_sift(const DevBatch& b, int nImages, cudaStream_t st, struct Prof* prof);
void launch_sift_desc(const DevBatch& b, int nImages, cudaStream_t st, struct Prof* prof);
void launch_mates_to_edges(const ebvo_mate* m, int n, double* lx, double* ly, double* lt, double* rx, double* ry, double* rt, cudaStream_t st);
void launch_gather_desc(const uint8_t* desc8, const ebvo_mate* m, int n, float* out, cudaStream_t st);
void launch_desc_to_float(const uint8_t* desc8, int n, float* out, cudaStream_t st);
void launch_undistort(const uint8_t* src, int srcPitch, uint8_t* dst, int dstPitch, int W, int H, const double K[9], const double dist[4], cudaStream_t st);   // undistort.cu   // blur + descriptors of every edge (sift.cu)
void upload_toed_tables();
void upload_sift_tables();
void init_toed_device();    // per-device: constant tables + function attributes (call after cudaSetDevice)
void init_match_device();
int make_toed_tensor_map(void* out128, const uint8_t* base, int W, int H, int pitch, size_t imgStride, int nImages);

// stage-dump support (debug): gate lists for stages 0..2 on frame 0
void launch_gate_count(const DevBatch& b, const DevParams& p, const double* F21, int mode, int* d_counts, cudaStream_t st);
void launch_gate_fill(const DevBatch& b, const DevParams& p, const double* F21, int mode, const int* d_offsets, int* d_ridx, cudaStream_t st);
// snapshot of the current candidate CSR of frame 0 into compact arrays (offsets computed on host)
// src < 0: live pool (counts = ccount); else dump[src]
void launch_snapshot(const DevBatch& b, int src, const int* d_offsets, int* ridx, double* x, double* y, double* th, double* score, cudaStream_t st);
void launch_compact(const DevBatch& b, int nFrames, ebvo_mate* d_out, int stride, cudaStream_t st, struct Prof* prof);
void launch_pack(const ebvo_mate* src, int srcStride, const int* nMates, int nFrames, ebvo_mate* dst, long long cap, int* offsets, cudaStream_t st, struct Prof* prof);
// individual matching stages (used by the stage-dump path, which snapshots between them)
void match_prologue(const DevBatch& b, const DevParams& p, const double* F21, int nFrames, cudaStream_t st, struct Prof* prof);
void match_gate(const DevBatch& b, const DevParams& p, const double* F21, int nFrames, cudaStream_t st, struct Prof* prof);
void match_sift(const DevBatch& b, const DevParams& p, int nFrames, cudaStream_t st, struct Prof* prof);
void match_ncc(const DevBatch& b, const DevParams& p, int nFrames, bool sift, cudaStream_t st, struct Prof* prof);
void match_gn(const DevBatch& b, const DevParams& p, int nFrames, cudaStream_t st, struct Prof* prof);
void match_cluster(const DevBatch& b, const DevParams& p, int nFrames, cudaStream_t st, struct Prof* prof);

// small helpers
void launch_edge_patches(const uint8_t* d_img, int w, int h, int pitch, const double* ex, const double* ey, const double* eth,
                         int n, double shift, float* plus, float* minus, cudaStream_t st);
void launch_ncc_pairs(const float* p1, const float* p2, int n, double* out, cudaStream_t st);
void launch_cluster_one(const double* x, const double* y, const double* th, int n, int by_orient, const DevParams& p,
                        double* cx, double* cy, double* cth, int* labels, int* nclusters, cudaStream_t st);

constexpr int TQ_CAP = 128;      // quads per keyframe mate that survive the NCC gate (errFlag 5 beyond)

struct TqDev {
    int W, H, pitch;
    const uint8_t *kfLraw, *kfLund, *kfRund, *cfLraw, *cfLund, *cfRund;
    int n_kf, n_cf;
    const double *kf, *cf;           // n x 6: left x, y, theta, right x, y, theta
    const uint8_t* kf_mask;          // n_kf or nullptr
    int cell, gw, gh, sr;
    double orient_deg, ncc_thresh, bnb_thresh, sift_thresh;
    const float* desc[4];            // 0 KF left, 1 KF right, 2 CF left, 3 CF right: n x 2 x 128 floats, or all nullptr (SIFT-off)
    int *cellCount, *cellStart, *cellCursor, *cellList, *lcell, *rcx, *rcy;
    float* np[4]; uint8_t* pf[4];    // 0 KF left, 1 KF right, 2 CF left, 3 CF right
    uint2* pk16[2];                  // CF left / right undistorted view packed with its Sobel gradients (int16; gather kernel)
    uint2* pkh[2];                   // the same as {half I, -, half gx, half gy} (exact; tiled kernel)
    int gn_gather;                   // 1: tq_gn_kernel (global-memory gathers, cross-check), 0: tq_gn_tile_kernel
    // pool 1 (after gate .. GN) and pool 2 (after clustering): [n_kf][TQ_CAP]
    int *cnt, *cnt2, *q_cf, *q_valid, *r_cf, *r_valid;
    double *q_ncc, *q_sc, *q_l, *q_r, *r_ncc, *r_sc, *r_l, *r_r;   // ncc/sc: 2 per entry, l/r: 3 per entry
    double *q_sift, *r_sift;         // 2 per entry
    int* errFlag;
    unsigned long long* counters;    // 0: gate survivors, 1: GN problems, 2: GN iterations, 3: grid candidates, 4: orientation survivors
};

// quad tracking launchers (temporal.inl, compiled inside match.cu); all asynchronous on `st`
void tq_prepare(const TqDev& d, const DevParams& p, cudaStream_t st, struct Prof* prof);
void tq_patches(const TqDev& d, const DevParams& p, cudaStream_t st, struct Prof* prof);
void tq_gate(const TqDev& d, int mode, int* counts, const int* offs, int* outCf, cudaStream_t st, struct Prof* prof);
void tq_gn(const TqDev& d, const DevParams& p, cudaStream_t st, struct Prof* prof);
void tq_cluster(const TqDev& d, const DevParams& p, cudaStream_t st, struct Prof* prof);
void tq_scan(const int* in, int* out, int n, cudaStream_t st, struct Prof* prof);
void tq_gather(const TqDev& d, int which, const int* offs, ebvo_quad* out, int cap, cudaStream_t st, struct Prof* prof);

// per-kernel event profiling
struct Prof {
    bool enabled = false;
    long long launchCount = 0;   // kernels launched through EBVO_KERNEL since the context was created (counted even when disabled)
    std::vector<std::string> names;
    std::vector<const char*> cnames;
    std::vector<float> ms;
    std::vector<int> launches;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    std::vector<int> pendingIdx;
    int index_of(const char* name);
    void begin(const char* name, cudaStream_t st);
    void end(cudaStream_t st);
    void collect();
    void reset();
};

// Checked build (-DEBVO_CHECKED, scripts/checked_run.sh): compute-sanitizer is closed on the GPU pool, so the kernels carry their own
// bounds assertions on every shared-memory tile / list / pool index; a violation sets the frame's error flag to 99 and the call fails
// with EBVO_ERR_CAPACITY "internal bounds assertion".  Compiled out of the product build.
#ifdef EBVO_CHECKED
#define EBVO_ASSERT(errp, cond) do { if (!(cond)) atomicExch((errp), 99); } while (0)
#else
#define EBVO_ASSERT(errp, cond) do { } while (0)
#endif

#define EBVO_KERNEL(prof, name, st, ...)            \
    do {                                            \
        if (prof) ++(prof)->launchCount;            \
        if ((prof) && (prof)->enabled) (prof)->begin(name, st); \
        __VA_ARGS__;                                \
        if ((prof) && (prof)->enabled) (prof)->end(st);         \
    } while (0)

}  // namespace ebvo
