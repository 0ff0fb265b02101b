/*
 * ebvo_b200.h - C ABI of libebvo_b200.so: the B200 (sm_100a) implementation of the per-frame hot path of
 * Brown-LEMS/Edge_Based_Visual_Odometry: third-order edge detection + epipolar-gated stereo edge
 * correspondence with oriented-patch NCC, Gauss-Newton refinement and edge clustering.
 *
 * This is the drop-in boundary.  Plain pointers and sizes only; no C++/torch
 * types.  Every function returns 0 on success or a negative EBVO_ERR_* code; ebvo_last_error() gives the
 * text.  There is NO CPU fallback: without a CUDA device ebvo_create() fails with EBVO_ERR_NO_DEVICE.
 * A context is bound to one GPU and one host thread; calls on a context are synchronous at return.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference repo).
 */
#ifndef EBVO_B200_H
#define EBVO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EBVO_OK 0
#define EBVO_ERR_NO_DEVICE (-1)
#define EBVO_ERR_INVALID (-2)
#define EBVO_ERR_CUDA (-3)
#define EBVO_ERR_CAPACITY (-4)
#define EBVO_ERR_NOMEM (-5)

typedef struct ebvo_ctx ebvo_ctx;

/* struct Edge, include/toed/cpu_toed.hpp:26-48 (location.x, location.y, orientation, index, frame_source). */
typedef struct ebvo_edge {
    double x, y, theta;
    int32_t index;
    int32_t frame_source;
} ebvo_edge;

/* Numeric content of a reference YAML (config/kitti.yaml:13-28 etc.); 3x3 matrices row-major. */
typedef struct ebvo_calib {
    double Kl[9], Kr[9], R21[9], T21[3];
} ebvo_calib;

/* The reference's compile-time knobs, include/definitions.h:17-36,76-77 and Stereo_Matches.h:84. */
typedef struct ebvo_params {
    double epipolar_line_dist_thresh; /* EPIPOLAR_LINE_DIST_THRESH 0.5 */
    double max_disparity;             /* MAX_DISPARITY 25 */
    double orientation_thresh_deg;    /* apply_orientation_filter(10.0), Stereo_Matches.cpp:1399 */
    double orthogonal_shift_mag;      /* ORTHOGONAL_SHIFT_MAG 5 */
    double ncc_thresh;                /* NCC_THRESH 0.6 */
    double bnb_ncc;                   /* BNB_NCC 0.9 */
    double bnb_sift;                  /* BNB_SIFT 0.4 */
    double sift_threshold;            /* SIFT_THRESHOLD 500 */
    double location_perturbation;     /* LOCATION_PERTURBATION 0.4 */
    double epip_tangency_displ_thresh;/* EPIP_TANGENCY_DISPL_THRESH 3 */
    double orient_perturbation;       /* ORIENT_PERTURBATION 0.174533 */
    double cluster_dist_thresh;       /* CLUSTER_DIST_THRESH 1 */
    double cluster_orient_thresh_deg; /* CLUSTER_ORIENT_THRESH 20 */
    double cluster_orient_gauss_sigma;/* CLUSTER_ORIENT_GAUSS_SIGMA 2 */
    int32_t max_cluster_size;         /* MAX_CLUSTER_SIZE 10 */
    int32_t gn_max_iter;              /* 20 */
    double gn_tol;                    /* 1e-3 */
    double gn_huber_delta;            /* 3.0 */
    double toed_mag_thresh;           /* "I_grad_mag <= 2" cpu_toed.cpp:406 */
    int32_t toed_border;              /* 10, cpu_toed.cpp:401-403,553 */
    int32_t gn_mode;                  /* Gauss-Newton kernel: 0 (default) reference arithmetic (FP64), shared-memory tiles;
                                         1: the same arithmetic, global-memory gathers (cross-check); 2: all FP32 (looser parity) */
    int32_t sift_mode;                /* 0 (default): the SIFT gate / BNB-SIFT run only with caller-supplied descriptors ("SIFT-off" otherwise);
                                         1: descriptors computed on the device (cv::SIFT::compute at the reference's keypoints, restated),
                                            S4 and S7' always run (Stereo_Matches.cpp:655-787,1452) */
} ebvo_params;

/* One finalised stereo mate: final_stereo_edge_pair.left_edge / right_edge, include/Dataset.h:291-309. */
typedef struct ebvo_mate {
    int32_t left_index; /* index of the left edge in the left TOED list */
    int32_t reserved;
    double lx, ly, ltheta;
    double rx, ry, rtheta;
    double score; /* NCC of the surviving cluster (refine_final_scores[0]) */
} ebvo_mate;

/* Stage identifiers for debug dumps (get_Stereo_Edge_Pairs stage order, Stereo_Matches.cpp:1360-1540). */
enum {
    EBVO_STAGE_EPI = 0,     /* apply_Epipolar_Line_Distance_Filtering */
    EBVO_STAGE_DISP = 1,    /* apply_Disparity_Filtering */
    EBVO_STAGE_ORIENT = 2,  /* apply_orientation_filter */
    EBVO_STAGE_SIFT = 3,    /* apply_SIFT_filtering (descriptors injected) */
    EBVO_STAGE_NCC = 4,     /* apply_NCC_Filtering (1st) */
    EBVO_STAGE_BNB_NCC = 5, /* apply_Best_Nearly_Best_Test(NCC) */
    EBVO_STAGE_BNB_SIFT = 6,/* apply_Best_Nearly_Best_Test(SIFT) */
    EBVO_STAGE_SHIFT = 7,   /* consolidate_redundant_edge_hypothesis(shift) */
    EBVO_STAGE_GN = 8,      /* refine_edge_disparity */
    EBVO_STAGE_CLUSTER = 9, /* consolidate_redundant_edge_hypothesis(shift+cluster) */
    EBVO_STAGE_NCC2 = 10,   /* apply_NCC_Filtering (2nd) */
    EBVO_STAGE_BEST = 11,   /* apply_Lowe_Ratio_Test (arg-max) */
    EBVO_STAGE_COUNT = 12
};

/* One surviving quad of the keyframe -> current-frame tracking: a keyframe stereo mate paired with a cluster of
 * current-frame stereo mates (KF_Temporal_Edge_Quads::candidate_quads, include/Temporal_Matches.h:16-28; the two
 * Temporal_CF_Edge_Cluster of a Candidate_Quad_Entry, include/Dataset.h:318-327). */
typedef struct ebvo_quad {
    int32_t kf_index;            /* index of the keyframe mate */
    int32_t cf_index;            /* cf_stereo_edge_mate_index */
    double lx, ly, ltheta;       /* CF_left->center_edge */
    double rx, ry, rtheta;       /* CF_right->center_edge */
    double ncc_left, ncc_right;  /* matching_scores.ncc_score of the two clusters */
    double sift_left, sift_right; /* matching_scores.sift_score (900 when the SIFT stages did not run) */
    double score_left, score_right; /* refine_final_score (1e6 before the photometric refinement) */
    int32_t valid;               /* refine_validity (valid_left && valid_right) */
    int32_t reserved;
} ebvo_quad;

/* Knobs of Temporal_Matches::get_Temporal_Edge_Pairs_from_Quads as written at src/Temporal_Matches.cpp:185-213 and
 * GRID_SIZE (include/definitions.h:45). */
typedef struct ebvo_quad_params {
    int32_t cell_size;      /* 15 */
    int32_t reserved;
    double grid_radius;     /* 30 */
    double orient_deg;      /* 10 */
    double ncc_thresh;      /* 0.8 */
    double bnb_thresh;      /* 0.8 (both best-nearly-best passes) */
    double sift_thresh;     /* 200 */
} ebvo_quad_params;

/* Stage identifiers of the quad tracking (stage order of get_Temporal_Edge_Pairs_from_Quads). */
enum {
    EBVO_TQ_GRID = 0,    /* add_edges_to_spatial_grid + apply_spatial_grid_filtering_quads */
    EBVO_TQ_ORIENT = 1,  /* apply_orientation_filtering_quads */
    EBVO_TQ_NCC = 2,     /* apply_NCC_filtering_quads */
    EBVO_TQ_SIFT = 3,    /* apply_SIFT_filtering_quads (pass-through without descriptors) */
    EBVO_TQ_BNB = 4,     /* apply_best_nearly_best_filtering_quads("NCC") */
    EBVO_TQ_BNB_SIFT = 5,/* apply_best_nearly_best_filtering_quads("SIFT") (pass-through without descriptors) */
    EBVO_TQ_GN = 6,      /* apply_photometric_refinement_quads */
    EBVO_TQ_CLUSTER = 7, /* apply_temporal_edge_clustering_quads */
    EBVO_TQ_COUNT = 8
};

/* Fills *p with the reference defaults. */
int ebvo_params_default(ebvo_params* p);

/* Context: owns device buffers for up to max_batch stereo frames of up to max_w x max_h pixels and up to
 * max_edges edges per image.  params may be NULL (defaults).  Replaces the ThirdOrderEdgeDetectionCPU
 * constructor/destructor (src/toed/cpu_toed.cpp:24-64,649-663) and Stereo_Matches() (Stereo_Matches.h:52). */
int ebvo_create(ebvo_ctx** out, int device, int max_w, int max_h, int max_batch, int max_edges, const ebvo_params* params);
void ebvo_destroy(ebvo_ctx* ctx);
const char* ebvo_last_error(const ebvo_ctx* ctx);

/* S0: F21 = Kr^-T [T21]x R21 Kl^-1 and F12 (src/Dataset.cpp:102-112).  Host FP64, row-major out. */
int ebvo_fundamental(const ebvo_calib* calib, double F21[9], double F12[9]);

/* ThirdOrderEdgeDetectionCPU::get_Third_Order_Edges (src/toed/cpu_toed.cpp:66-77) on one 8-bit image.
 * out receives toed_edges in the reference's row-major order; *n_total = Total_Num_Of_TOED. */
int ebvo_toed(ebvo_ctx* ctx, const uint8_t* img, int w, int h, int stride, ebvo_edge* out, int cap, int* n_edges,
              int* n_total);

/* Stereo_Matches::get_Stereo_Edge_Pairs + finalize_stereo_edge_mates (src/Stereo_Matches.cpp:1360-1653),
 * no-GT branch, on caller-supplied edge lists.  L_und/R_und may equal L_raw/R_raw (zero distortion).
 * descL/descR: optional per-edge SIFT descriptor pairs (n*2*128 floats, augment_Edge_Data layout); NULL =>
 * the SIFT gate and BNB-SIFT are skipped ("SIFT-off"). */
int ebvo_stereo_match(ebvo_ctx* ctx, const ebvo_calib* calib, const uint8_t* L_raw, const uint8_t* R_raw,
                      const uint8_t* L_und, const uint8_t* R_und, int w, int h, int stride, const ebvo_edge* L, int nL,
                      const ebvo_edge* R, int nR, const float* descL, const float* descR, ebvo_mate* out, int cap,
                      int* n_mates);

/* ebvo_stereo_match followed, in the SAME call and on the images already on the device, by everything the reference's
 * get_Stereo_Edge_Pairs leaves behind per surviving left edge and finalize_stereo_edge_mates adds per mate
 * (Stereo_Matches.cpp:570-576, 655-689, 1578-1653): the "+"/"-" 7x7 patches of the matched left edge from the RAW left image
 * (l_plus49 / l_minus49, 49 floats per mate), the mate's patches from the UNDISTORTED right image (r_plus49 / r_minus49) and -
 * for a context created with sift_mode = 1 - the two 128-entry descriptors of the left edge and of the mate (l_desc256 /
 * r_desc256, 256 floats per mate).  Any of the six outputs may be NULL; they hold `cap` entries.  One upload of the four
 * images and one synchronisation per frame instead of the five calls the members would otherwise make. */
int ebvo_stereo_match_full(ebvo_ctx* ctx, const ebvo_calib* calib, const uint8_t* L_raw, const uint8_t* R_raw, const uint8_t* L_und,
                           const uint8_t* R_und, int w, int h, int stride, const ebvo_edge* L, int nL, const ebvo_edge* R, int nR,
                           ebvo_mate* out, int cap, int* n_mates, float* l_plus49, float* l_minus49, float* r_plus49, float* r_minus49,
                           float* l_desc256, float* r_desc256);

/* Pipeline::prepare_Stereo_Images edge part + get_Stereo_Edge_Correspondences (src/Pipeline.cpp:93-131):
 * TOED on both views + matching, edges stay on the device.  Optional edge outputs may be NULL. */
int ebvo_stereo_frame(ebvo_ctx* ctx, const ebvo_calib* calib, const uint8_t* L_img, const uint8_t* R_img, int w, int h,
                      int stride, ebvo_mate* out, int cap, int* n_mates, ebvo_edge* L_edges, int* nL, ebvo_edge* R_edges,
                      int* nR, int edge_cap);

/* Batch of independent stereo frames (the StereoIterator::getNext loop of cmd/main_VO.cpp:99-113 with
 * frames already in memory).  L_imgs/R_imgs: n_frames host pointers.  out: n_frames*cap mates;
 * n_mates: n_frames counts.  Host buffers in, host buffers out (copies are part of the call).
 * Capacities (max_edges per image, the candidate pool of 8 x max_edges per frame, 128 NCC survivors / clusterer inputs per left
 * edge) are checked PER FRAME: a frame that exhausts one is reported with n_mates[f] = -1 and the call returns
 * EBVO_ERR_CAPACITY (ebvo_last_error names the first such frame), while every other frame of the batch keeps its mates. */
int ebvo_stereo_batch(ebvo_ctx* ctx, const ebvo_calib* calib, int n_frames, const uint8_t* const* L_imgs,
                      const uint8_t* const* R_imgs, int w, int h, int stride, ebvo_mate* out, int cap, int* n_mates);

/* Pipeline::get_Temporal_Edge_Correspondences (src/Pipeline.cpp:147-165): SpatialGrid of the current frame's mates
 * (Temporal_Matches::add_edges_to_spatial_grid, src/Temporal_Matches.cpp:18-55) + the filter chain of
 * Temporal_Matches::get_Temporal_Edge_Pairs_from_Quads (:168-218) on the stereo mates of a keyframe (kf) and of
 * the current frame (cf), e.g. the outputs of two ebvo_stereo_frame calls.  *_Lraw: raw left image (left patches,
 * Stereo_Matches.cpp:562,578), *_Lund / *_Rund: undistorted views (right patches :1582, Gauss-Newton :578-582).
 * kf_mask (n_kf bytes or NULL): which keyframe mates take part; the reference takes those with a non-empty
 * veridical_quads list (build_Veridical_Quads, :57-166, a ground-truth construct that stays on the host).
 * desc_kf_l / desc_kf_r / desc_cf_l / desc_cf_r: the mates' left_edge_descriptors / right_edge_descriptors as n x 2 x 128
 * floats (final_stereo_edge_pair, include/Dataset.h:299-300); all four NULL => the SIFT gate (:471-515) and the SIFT
 * best-nearly-best pass are skipped ("SIFT-off").  qp may be NULL (reference values).  out receives the surviving quads grouped by keyframe mate in index order. */
int ebvo_temporal_quads(ebvo_ctx* ctx, const uint8_t* kf_Lraw, const uint8_t* kf_Lund, const uint8_t* kf_Rund,
                        const uint8_t* cf_Lraw, const uint8_t* cf_Lund, const uint8_t* cf_Rund, int w, int h, int stride,
                        const ebvo_mate* kf, int n_kf, const uint8_t* kf_mask, const ebvo_mate* cf, int n_cf,
                        const float* desc_kf_l, const float* desc_kf_r, const float* desc_cf_l, const float* desc_cf_r,
                        const ebvo_quad_params* qp, ebvo_quad* out, int cap, int* n_quads);

/* The same call stopped after `stage` (EBVO_TQ_*): the candidate quads as the reference holds them at that point, for
 * stage-by-stage parity tests.  off (n_kf + 1 ints, may be NULL) receives the start of every keyframe mate's list. */
int ebvo_temporal_quads_stage(ebvo_ctx* ctx, const uint8_t* kf_Lraw, const uint8_t* kf_Lund, const uint8_t* kf_Rund,
                              const uint8_t* cf_Lraw, const uint8_t* cf_Lund, const uint8_t* cf_Rund, int w, int h, int stride,
                              const ebvo_mate* kf, int n_kf, const uint8_t* kf_mask, const ebvo_mate* cf, int n_cf,
                              const float* desc_kf_l, const float* desc_kf_r, const float* desc_cf_l, const float* desc_cf_r,
                              const ebvo_quad_params* qp, int stage, int* off, ebvo_quad* out, int cap, int* n_quads);

/* Work counters of the last quad-tracking call: 0 gate survivors, 1 Gauss-Newton problems, 2 Gauss-Newton iterations,
 * 3 grid candidates, 4 orientation survivors (8 values). */
int ebvo_temporal_counters(ebvo_ctx* ctx, long long* out8);

/* ebvo_stereo_batch with the mates of the whole batch back to back: frame f's records follow frame f - 1's in `out`
 * (n_mates[f] of them; *total = records written, at most cap_records, EBVO_ERR_CAPACITY beyond).  Every sub-batch's records
 * are copied out while the next sub-batches compute, so when `out` is page-locked memory shared by the processes of a box
 * (one slice per GPU: edge_based_visual_odometry_b200/sharding.py HostGather) the call IS the final result gather of a
 * batch sharded over the GPUs, overlapped with the computation (Pipeline.cpp:109-131 per frame; BASELINE configs[2]). */
int ebvo_stereo_batch_packed(ebvo_ctx* ctx, const ebvo_calib* calib, int n_frames, const uint8_t* const* L_imgs,
                             const uint8_t* const* R_imgs, int w, int h, int stride, ebvo_mate* out, long long cap_records,
                             int* n_mates, long long* total);
/* The same batch over several contexts (normally one per GPU of the box): frames are split into contiguous blocks of
 * ceil(n_frames / n_ctx), one host thread per context, no exchange between devices (frames are independent:
 * src/Pipeline.cpp:64-145 reads nothing from other frames); results land in out / n_mates at their global frame index.
 * Returns the first non-zero status of any block (the message is on that block's context). */
int ebvo_stereo_batch_multi(ebvo_ctx* const* ctxs, int n_ctx, const ebvo_calib* calib, int n_frames, const uint8_t* const* L_imgs,
                            const uint8_t* const* R_imgs, int w, int h, int stride, ebvo_mate* out, int cap, int* n_mates);

/* Device-resident variant used to time the kernels alone: upload once, run many times, download once. */
int ebvo_batch_upload(ebvo_ctx* ctx, int n_frames, const uint8_t* const* L_imgs, const uint8_t* const* R_imgs, int w,
                      int h, int stride);
int ebvo_batch_run(ebvo_ctx* ctx, const ebvo_calib* calib, int do_match); /* async on the context stream */
int ebvo_batch_sync(ebvo_ctx* ctx);
int ebvo_batch_download(ebvo_ctx* ctx, ebvo_mate* out, int cap, int* n_mates);
int ebvo_batch_counts(ebvo_ctx* ctx, int* nL, int* nR, int* n_mates, long long* stage_counts /* n_frames*8 or NULL */);
/* The results of the last batch (ebvo_batch_run, or ebvo_stereo_batch called with out = NULL), still on the device, packed
 * back to back without padding into the caller's DEVICE buffer d_dst (cap_records ebvo_mate records); d_offsets (device,
 * n_frames + 1 ints) receives the first record of every frame and the total.  This is the payload of the one exchange a
 * sharded batch has - the final gather of the per-frame results that the reference's frame loop accumulates on one host
 * (cmd/main_VO.cpp:99-113 -> Pipeline::ProcessStereoFrame) - so that it can go GPU to GPU (NCCL) at its real size. */
int ebvo_batch_pack(ebvo_ctx* ctx, void* d_dst, long long cap_records, int* d_offsets, long long* total);

/* Utility::get_edge_patches (src/utility.cpp:182-212): 7x7 "+" and "-" patches of n edges of one image. */
int ebvo_edge_patches(ebvo_ctx* ctx, const uint8_t* img, int w, int h, int stride, const ebvo_edge* edges, int n,
                      float* plus49, float* minus49);
/* Utility::get_patch_similarity / MatlabNCCComputer::computeNCC (src/utility.cpp:163-180,
 * src/MatlabNCCComputer.cpp:58-90): NCC of n pairs of 49-vectors. */
int ebvo_ncc_patch_pair(ebvo_ctx* ctx, const float* p1, const float* p2, int n_pairs, double* out);
/* EdgeClusterer::performClustering (src/EdgeClusterer.cpp:119-302) on one candidate set. */
int ebvo_cluster(ebvo_ctx* ctx, const ebvo_edge* edges, int n, int by_orientation, ebvo_edge* centers, int* labels,
                 int* n_clusters);
/* augment_Edge_Data (src/Stereo_Matches.cpp:655-689): the two SIFT descriptors of n edges of one image, keypoints at
 * p +- 8 (sin theta, -cos theta), size 1, angle deg(theta); out = n*2*128 floats (values 0..255).  Needs sift_mode 1. */
int ebvo_sift_descriptors(ebvo_ctx* ctx, const uint8_t* img, int w, int h, int stride, const ebvo_edge* edges, int n, float* out);
/* cv::undistort(image, out, K, dist) of Pipeline::prepare_Stereo_Images (src/Pipeline.cpp:78-79) for one 8-bit image:
 * K row-major 3x3 (the intrinsics of the config YAML files), dist = {k1, k2, p1, p2}; bit-identical to OpenCV 4.x (fixed-point remap). */
int ebvo_undistort(ebvo_ctx* ctx, const uint8_t* img, int w, int h, int stride, const double K[9], const double dist[4], uint8_t* out,
                   int out_stride);
/* util_compute_Img_Gradients (include/utility.h:131-141): Sobel 3x3 * 1/8, reflect-101 border. */
int ebvo_sobel(ebvo_ctx* ctx, const uint8_t* img, int w, int h, int stride, float* gx, float* gy);

/* Debug stage dumps of the LAST ebvo_stereo_match call (enable before the call).  Ragged per-left-edge lists:
 * offsets[nL+1]; per entry right-edge index (or -1), x, y, theta, score. */
int ebvo_set_stage_dumps(ebvo_ctx* ctx, int enable);
int ebvo_stage_size(ebvo_ctx* ctx, int stage, int* n_left, int* total);
int ebvo_stage_fetch(ebvo_ctx* ctx, int stage, int* offsets, int* ridx, double* x, double* y, double* theta,
                     double* score);

/* Per-kernel device times (ms) of the last batch/frame call, measured with CUDA events on the context
 * stream when profiling is enabled.  names: array of const char* owned by the library. */
int ebvo_set_profiling(ebvo_ctx* ctx, int enable);
int ebvo_get_kernel_times(ebvo_ctx* ctx, const char*** names, const float** ms, const int** launches, int* n);
/* Number of kernels the context has launched since it was created (monotonic; counted with or without profiling). */
long long ebvo_launch_count(ebvo_ctx* ctx);
/* The CUDA stream (cudaStream_t) the context launches on, for external event timing. */
void* ebvo_stream(ebvo_ctx* ctx);
/* Page-locked host memory for the buffers a caller hands to the entry points above (images in, mates / patches /
 * descriptors out): copies to and from such buffers are direct DMA transfers instead of staged ones.  Plain memory works
 * everywhere; this is an optimisation the C++ drop-ins use for their per-frame result buffers (the reference has no
 * counterpart: its results never leave host memory).  ebvo_host_alloc returns NULL when the allocation fails. */
void* ebvo_host_alloc(size_t bytes);
void ebvo_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* EBVO_B200_H */
