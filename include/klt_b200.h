/*
 * klt_b200.h -- C ABI of libklt_b200.so: B200 (sm_100a) pyramid build + pyramidal Lucas-Kanade.
 *
 * Drop-in boundary for the one hot path of JonasFrey96/Visual-Odom-Pipeline that this project
 * replaces: the four `cv2.calcOpticalFlowPyrLK(im0, im1, p, None, winSize, maxLevel, criteria)`
 * calls per frame at reference src/extractor/extractor.py:44,45 (extend_tracks) and :65,66
 * (extend_landmarks), with parameters from src/extractor/extractor.py:16-19.  The reference has no
 * plugin API for this path -- the interface *is* the OpenCV function -- so the entry points below
 * are what a ctypes binding of that function (and of cv2.buildOpticalFlowPyramid, which OpenCV
 * runs inside it) binds.  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions: plain C, no exceptions / STL / torch types across the ABI.  Every function returns
 * a klt_status (0 = ok, <0 = KLT_ERR_*, >0 = cudaError_t passed through).  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  Device pointers are owned by the
 * caller.  Functions named *_host take HOST pointers, do their own H2D/D2H and are synchronous on
 * return (like the cv2 call); all others are asynchronous on `stream`.
 *
 * There is no CPU fallback anywhere in this library.
 */
#ifndef KLT_B200_H
#define KLT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KLT_B200_VERSION 120 /* 0.1.2: + bilateral pre-filter; thread-safe contexts; pageable inputs staged in the library */

#define KLT_MAX_LEVELS 16   /* pyramid levels incl. level 0 (cv2 stops when a level <= winSize) */
#define KLT_MAX_WIN_AREA 4096 /* win_w * win_h supported by the LK kernel (cv2 default 21x21) */

/* cv2.TERM_CRITERIA_* and cv2.OPTFLOW_* values (same numbers as OpenCV) */
#define KLT_TERM_COUNT 1
#define KLT_TERM_EPS 2
#define KLT_OPTFLOW_USE_INITIAL_FLOW 4
#define KLT_OPTFLOW_LK_GET_MIN_EIGENVALS 8

typedef int klt_status;
#define KLT_OK 0
#define KLT_ERR_INVALID_ARG (-1)   /* cv2 would raise error -215 (assertion) */
#define KLT_ERR_UNSUPPORTED (-2)   /* valid for cv2, outside this library's limits */
#define KLT_ERR_NO_DEVICE (-3)     /* no CUDA device / wrong architecture: never falls back */
#define KLT_ERR_OUT_OF_MEMORY (-4)
#define KLT_ERR_INTERNAL (-5)

/* Opaque context: device id, two streams, device workspace, pinned staging, helper threads for pageable inputs.
 * Threads: a context may be shared.  The *_host entry points use the context's workspaces and streams and are
 * serialised by a mutex inside the context (concurrent callers queue up; one context per thread avoids that).  The
 * device-pointer entry points keep no per-call state in the context and may run concurrently on different streams.
 * Every entry point makes the context's device current for the duration of the call and restores the caller's. */
typedef struct klt_ctx klt_ctx;

/* One pyramid level as the kernels see it (row `y` of batch item `b` starts at
 * data + b*batch_stride + y*pitch). */
typedef struct klt_level {
    int32_t w, h;
    int64_t pitch;        /* bytes between rows */
    int64_t batch_stride; /* bytes between batch items */
    int64_t offset;       /* byte offset of this level inside the pyramid buffer (levels >= 1) */
} klt_level;

/* Geometry of a pyramid: level 0 is the caller's image batch (never copied), levels 1..top live
 * in ONE caller-provided buffer of `bytes` bytes.  Mirrors what cv2.buildOpticalFlowPyramid
 * returns (retval = top, list of levels) -- SURVEY.md s8b / A.2. */
typedef struct klt_pyr_layout {
    int32_t top;    /* index of the last level built (cv2's retval; may be < max_level) */
    int32_t batch;
    klt_level level[KLT_MAX_LEVELS]; /* level[0].pitch/batch_stride are filled by the caller */
    int64_t bytes;  /* size of the levels>=1 buffer for the whole batch */
} klt_pyr_layout;

/* ---- context -------------------------------------------------------------------------------- */
klt_status klt_create(int device, klt_ctx** out);
klt_status klt_destroy(klt_ctx* ctx);
int klt_version(void);
const char* klt_status_string(klt_status s);
/* number of SMs / name of the device the context is bound to (for bench & grid sizing) */
klt_status klt_device_info(klt_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len);

/* ---- pinned host memory (so numpy inputs can be DMA'd without a staging copy) ---------------- */
klt_status klt_host_alloc(void** ptr, int64_t bytes);
klt_status klt_host_free(void* ptr);

/* ---- pyramid (replaces cv2.buildOpticalFlowPyramid / pyrDown) ----------------- */
/* Fills `out` for a w x h image batch: number of levels per cv2's rule (stop before a level whose
 * width <= win_w or height <= win_h), per-level size / pitch / offset.  Host-only, no CUDA call. */
klt_status klt_pyr_plan(int w, int h, int win_w, int win_h, int max_level, int batch, klt_pyr_layout* out);

/* Builds levels 1..top of batch items [first_item, first_item + n_items) (n_items <= 0: all of
 * them).  d_img: level 0 of item 0 (u8, pitch/batch_stride from layout->level[0]); d_pyr: buffer of
 * layout->bytes.  One kernel launch per level for the whole item range. */
klt_status klt_pyr_build(klt_ctx* ctx, const uint8_t* d_img, const klt_pyr_layout* layout,
                         uint8_t* d_pyr, int first_item, int n_items, void* stream);

/* Single pyrDown step (cv2.pyrDown on 1-channel u8, BORDER_REFLECT_101), device pointers. */
klt_status klt_pyr_down(klt_ctx* ctx, const uint8_t* d_src, int w, int h, int64_t src_pitch,
                        int64_t src_batch_stride, uint8_t* d_dst, int64_t dst_pitch,
                        int64_t dst_batch_stride, int batch, void* stream);

/* ---- Lucas-Kanade (replaces the per-level Scharr + LKTrackerInvoker loop) -- */
typedef struct klt_lk_params {
    int32_t win_w, win_h;
    int32_t crit_type;      /* KLT_TERM_COUNT | KLT_TERM_EPS, as in cv2's criteria[0] */
    int32_t crit_max_count; /* criteria[1] */
    double crit_eps;        /* criteria[2] */
    int32_t flags;          /* KLT_OPTFLOW_*; every other bit must be 0 (0x100 / 0x200 are used by diagnostics and the launcher) */
    double min_eig_threshold; /* cv2 default 1e-4 */
} klt_lk_params;

/* Tracks n_per_pair points in each of n_pairs frame pairs, all pyramid levels in ONE launch.
 * `layout` describes both pyramid buffers (they may be the same buffer): pair i uses batch item
 * prev_first + i*pair_stride of (d_prev_img, d_prev_pyr) and item next_first + i*pair_stride of
 * (d_next_img, d_next_pyr), so one batched pyramid build can hold previous and next frames side by
 * side (interleaved: prev_first 0, next_first 1, pair_stride 2; a frame ring: 0, 1, 1).
 * d_prev_pts / d_next_pts: float2 [n_pairs][n_per_pair]; d_status: u8; d_err: float;
 * d_iters (optional, may be NULL): int32 LK iterations executed per point over all levels.
 * d_next_pts is read only with KLT_OPTFLOW_USE_INITIAL_FLOW.  Points whose err cv2 leaves
 * uninitialised get err = 0. */
klt_status klt_lk_track(klt_ctx* ctx,
                        const uint8_t* d_prev_img, const uint8_t* d_prev_pyr,
                        const uint8_t* d_next_img, const uint8_t* d_next_pyr,
                        const klt_pyr_layout* layout, int prev_first, int next_first, int pair_stride, int n_pairs,
                        const float* d_prev_pts, float* d_next_pts, uint8_t* d_status, float* d_err,
                        int32_t* d_iters, int n_per_pair, const klt_lk_params* params, void* stream);

/* ---- host-pointer entry points: exactly what a binding of the cv2 functions needs ------------ */
/* cv2.calcOpticalFlowPyrLK(prevImg, nextImg, prevPts, nextPts, winSize, maxLevel, criteria[, flags,
 * minEigThreshold]) -> nextPts, status, err.  HOST buffers (pinned or pageable); images u8 with
 * arbitrary row pitch; points float32 [n][2].  Synchronous.  top_level_out (optional) receives the
 * last pyramid level used.  Pinned images are DMA'd in place; pageable ones (numpy arrays of the un-edited
 * reference) are staged through a pinned landing zone by the calling thread and three helper threads of the
 * context while the DMA of the first image already runs. */
klt_status klt_calc_optical_flow_pyr_lk_host(klt_ctx* ctx,
                        const uint8_t* prev_img, int64_t prev_pitch,
                        const uint8_t* next_img, int64_t next_pitch, int w, int h,
                        const float* prev_pts, float* next_pts, uint8_t* status, float* err, int n,
                        int max_level, const klt_lk_params* params, int* top_level_out);

/* ---- the reference's tracking step, fused ------------------------------ */
/* Bidirectional-error / bounds filter, device pointers: reference src/extractor/extractor.py:46-47,53 (extend_tracks)
 * and :67-68,75 (extend_landmarks):  bidir = max(|p0 - p0r|) over x, y;  keep = bidir < max_bidir_error and
 * 0 <= p1.x <= w and 0 <= p1.y <= h (inclusive, as in the reference; NaN fails).  d_bidir_err may be NULL. */
klt_status klt_track_filter(klt_ctx* ctx, const float* d_p0, const float* d_p1, const float* d_p0r, int64_t n,
                            float max_bidir_error, int w, int h, uint8_t* d_keep, float* d_bidir_err, void* stream);

/* One call for what extend_tracks / extend_landmarks do with cv2 (extractor.py:43-53, 64-75): p1 = LK(im0, im1, p0),
 * p0r = LK(im0, im1, p1) (the reference's second call runs in the same direction, started from the forward result),
 * then the filter above.  HOST buffers, synchronous.  One upload and ONE pyramid build per image instead of the four
 * OpenCV builds two cv2 calls make.  next_pts / status / err are those of the first call (bit-identical to cv2's). */
klt_status klt_track_bidirectional_host(klt_ctx* ctx,
                        const uint8_t* prev_img, int64_t prev_pitch,
                        const uint8_t* next_img, int64_t next_pitch, int w, int h,
                        const float* prev_pts, int n, int max_level, const klt_lk_params* params,
                        float max_bidir_error, float* next_pts, uint8_t* status, float* err,
                        uint8_t* keep, float* bidir_err);

/* cv2.buildOpticalFlowPyramid(img, winSize, maxLevel) without derivatives / borders: writes level
 * l (l = 0..top) tightly packed (pitch = level width) at out + level_offsets[l].  Pass out = NULL
 * to query `top` and the offsets / total size only. */
klt_status klt_build_optical_flow_pyramid_host(klt_ctx* ctx, const uint8_t* img, int64_t pitch, int w, int h,
                        int win_w, int win_h, int max_level, uint8_t* out, int64_t* level_offsets /*[KLT_MAX_LEVELS+1]*/,
                        int* top_out);

/* ---- Shi-Tomasi corner detection -------------------------------------------------------
 * Replaces cv2.goodFeaturesToTrack(img, mask=mask, maxCorners=1000, qualityLevel=0.03, minDistance, blockSize=31)
 * as called at reference src/extractor/extractor.py:110-111 (parameters :21-24; once per frame from
 * src/pipeline/pipeline.py:159-163) and the cv2.cornerMinEigenVal it runs inside (gradient size 3, no Harris).
 * Arithmetic: oracle/gftt_oracle.c G.1-G.8, bit-exact with the cv2 wheel of the image. */

/* Bytes of device scratch klt_corner_min_eigen_val needs for a batch of w x h images. */
int64_t klt_corner_ws_bytes(int w, int h, int batch);

/* cv2.cornerMinEigenVal(src, blockSize, ksize=3), device pointers, batched, asynchronous.  d_img: u8, any pitch;
 * d_eig: float32, eig_pitch / eig_batch_stride in ELEMENTS.  d_max (optional): one uint32 per batch item, zeroed by
 * the caller, receives the order-preserving encoding of the maximum eigenvalue over mask != 0 (d_mask optional =
 * everywhere); it is the input of klt_corner_candidates.  d_ws: klt_corner_ws_bytes() bytes, 256-byte aligned. */
klt_status klt_corner_min_eigen_val(klt_ctx* ctx, const uint8_t* d_img, int w, int h, int64_t pitch, int64_t batch_stride,
                        int batch, int block_size, float* d_eig, int64_t eig_pitch, int64_t eig_batch_stride,
                        const uint8_t* d_mask, int64_t mask_pitch, int64_t mask_batch_stride, uint32_t* d_max,
                        void* d_ws, int64_t ws_bytes, void* stream);

/* Threshold (max * quality_level), 3x3 dilation and local-maximum test of goodFeaturesToTrack fused: appends one
 * 64-bit key per candidate, (order-preserving float encoding << 32) | (y << 16) | x, in arbitrary order, to d_keys
 * (capacity keys per batch item, keys_batch_stride in ELEMENTS).  d_count: one uint32 per batch item, zeroed by the
 * caller, receives the number of candidates (may exceed capacity: the excess is dropped). */
klt_status klt_corner_candidates(klt_ctx* ctx, const float* d_eig, int64_t eig_pitch, int64_t eig_batch_stride, int w, int h,
                        int batch, const uint8_t* d_mask, int64_t mask_pitch, int64_t mask_batch_stride,
                        const uint32_t* d_max, double quality_level, uint64_t* d_keys, int64_t keys_batch_stride,
                        int capacity, uint32_t* d_count, void* stream);

/* The detection mask of reference src/extractor/extractor.py:102-107, built on the device: 255 everywhere, then a
 * filled cv2.circle(mask, np.int32((x, y)), radius, 0, -1) around each of the n points (float32 x, y pairs, device).
 * Bit-identical to OpenCV's filled circle (midpoint circle), centres outside the image included.  radius <= 127. */
klt_status klt_corner_mask_from_points(klt_ctx* ctx, const float* d_points, int n, int radius, int w, int h,
                        uint8_t* d_mask, int64_t mask_pitch, void* stream);

/* The sequential tail of goodFeaturesToTrack on the HOST: sorts the keys of klt_corner_candidates in place (strongest
 * first, equal values: the later pixel in raster order first, like OpenCV) and runs the greedy minimum-distance
 * selection.  corners: capacity x (x, y) floats; *n_out = corners found (<= max_corners if max_corners > 0). */
klt_status klt_select_corners_host(uint64_t* keys, int64_t n_keys, int w, int h, int max_corners, double min_distance,
                        float* corners, int capacity, int* n_out);

/* cv2.cornerMinEigenVal(img, blockSize, ksize=3) with HOST buffers, synchronous.  eig: h x w float32, packed. */
klt_status klt_corner_min_eigen_val_host(klt_ctx* ctx, const uint8_t* img, int64_t pitch, int w, int h, int block_size,
                        float* eig);

/* cv2.goodFeaturesToTrack(img, maxCorners, qualityLevel, minDistance, mask, blockSize) with HOST buffers, synchronous.
 * mask may be NULL.  corners: capacity x (x, y) float32, strongest first; *n_out = number found (cv2 returns None for
 * 0).  KLT_ERR_INVALID_ARG where cv2 asserts (quality_level <= 0, min_distance < 0, max_corners < 0). */
klt_status klt_good_features_to_track_host(klt_ctx* ctx, const uint8_t* img, int64_t pitch, int w, int h,
                        const uint8_t* mask, int64_t mask_pitch, int max_corners, double quality_level,
                        double min_distance, int block_size, float* corners, int capacity, int* n_out);

/* Caller-side fusion of the reference's extract() step (src/extractor/extractor.py:102-111): instead of a host mask the
 * caller passes the keypoints it tracks (HOST float32 x, y pairs) and mask_radius; the mask is rasterised on the device
 * (klt_corner_mask_from_points), so the call uploads n_points * 8 bytes instead of a w x h mask.  Same result as
 * building the mask with cv2.circle and calling klt_good_features_to_track_host. */
klt_status klt_good_features_to_track_points_host(klt_ctx* ctx, const uint8_t* img, int64_t pitch, int w, int h,
                        const float* points, int n_points, int mask_radius, int max_corners, double quality_level,
                        double min_distance, int block_size, float* corners, int capacity, int* n_out);

/* ---- bilateral pre-filter ------------------------------------------------------------------
 * Replaces cv2.bilateralFilter(img, d=5, sigmaColor=1.5, sigmaSpace=1.5) as the reference's loader applies it to every
 * frame (src/loader/loader.py:16-20,86): 8-bit single channel, BORDER_REFLECT_101 (cv2's default), d <= 15.
 * Arithmetic: oracle/bilateral_oracle.c B.1-B.6 = OpenCV's own code path (an OpenCV build without IPP); differs from
 * it only on exact rounding ties (< 1e-5 of the pixels, by 1), and by at most 1 from the IPP-enabled wheel. */

/* Device pointers, batched, asynchronous on `stream`.  d_dst must not alias d_src. */
klt_status klt_bilateral_filter(klt_ctx* ctx, const uint8_t* d_src, int w, int h, int64_t src_pitch, int64_t src_batch_stride,
                        uint8_t* d_dst, int64_t dst_pitch, int64_t dst_batch_stride, int batch, int d, double sigma_color,
                        double sigma_space, void* stream);

/* HOST buffers, synchronous. */
klt_status klt_bilateral_filter_host(klt_ctx* ctx, const uint8_t* img, int64_t pitch, int w, int h, int d, double sigma_color,
                        double sigma_space, uint8_t* out, int64_t out_pitch);

#ifdef __cplusplus
}
#endif
#endif /* KLT_B200_H */
