/* ivln_map.h -- C ABI of the B200-native semantic-map update (libivlnmap.so).
 *
 * The reference (jacobkrantz/IVLN-CE) is pure Python and has no FFI on this
 * path; these entry points are what a binding for
 *   ivlnce_baselines/common/mapping_module/mapper.py:904-947  (MappingModule.forward)
 * would call.  Plain C: PODs, raw device pointers, a CUDA stream handle, int
 * status codes (0 = ok).  No torch types, no C++ exceptions, no allocation and
 * no synchronisation behind the caller's back (except ivm_read_status, which is
 * documented to synchronise the given stream).  A context is not re-entrant.
 *
 * Which reference code each call replaces:
 *   ivm_step_iterative  UpdateWorldSemanticPointcloud.forward    mapper.py:825-848
 *                       (+ GenerateSemanticPointCloud 398-425, KeepHighest 428-474,
 *                        projector/core.py:117-230, PredictSemantics argmax 795-798)
 *                       FilterPointCloudByRobotHeight.forward     mapper.py:884-901
 *                       OccupancySemanticMapMemory.update         mapper.py:555-636
 *   ivm_known_load      SemanticPointcloud.from_npz_file + GetGTWorldSemanticPointcloud
 *                                                                 mapper.py:283-294, 862-881
 *   ivm_step_known      same raster as above on the loaded scene clouds
 *   ivm_export_world    MappingModule.get_world_semantic_pointcloud  mapper.py:946-947
 */
#ifndef IVLN_MAP_H
#define IVLN_MAP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ivm_ctx ivm_ctx;
typedef void *ivm_stream_t; /* a cudaStream_t */

enum {
    IVM_OK = 0,
    IVM_E_INVALID = 1,      /* bad argument */
    IVM_E_WORKSPACE = 2,    /* workspace too small / misaligned */
    IVM_E_CUDA = 3,         /* a CUDA runtime call failed; see ivm_last_cuda_error */
    IVM_E_STEP_OVERFLOW = 4 /* 2^24-1 steps reached; call ivm_rebase_stamps */
};

typedef struct ivm_config {
    int32_t max_envs;     /* envs this context can hold (batch size upper bound) */
    int32_t height, width; /* depth image size (CameraParameters.features_spatial_dimensions) */
    int32_t map_rows, map_cols; /* MapDimensions.num_rows / num_cols */
    float res;            /* (float) resolution_meters */
    float half_res;       /* (float)(resolution_meters / 2): de-dup cell, mapper.py:464 */
    float half_h;         /* (float)(height_meters / 2), mapper.py:112 */
    float half_w;         /* (float)(width_meters / 2) */
    int32_t store_rows, store_cols; /* world store extent per env, in half-cells */
    int32_t mode;         /* 0 = iterative (depth ingested every step), 1 = known map */
    int64_t known_capacity; /* known mode: max points per env */
    int32_t tile_rows, tile_cols;   /* ego tile per raster CTA; 0 = choose */
    int32_t reserved[4];  /* [0] step variant: 0 = auto (the persistent step kernel when it applies), 1 = four
                           * kernels with register-staged score loads, 2 = four kernels with the bulk-async ring;
                           * [1] profiling / test switches of the persistent kernel (0 in production; bit 128 =
                           * two tiles per chunk, exercises the multi-chunk path; bits 8..11 = ring depth); [2] 2 = stage raster tiles in
                           * shared memory with cp.async (measured slower; off by default);
                           * [3] test switch: period of the candidate-plane stamp (0 = the longest the pixel index leaves
                           * room for, 65 535 steps at 256x256; the plane is cleared once per period) */
} ivm_config;

typedef struct ivm_status {
    uint32_t error_flags; /* 1 = point outside world store, 2 = edge list overflow, 4 = known cloud overflow,
                           * 8 = grid barrier time-out in the fused kernel, 16 = frame candidate table full */
    uint32_t pad;
    uint64_t stats[8];    /* valid pixels, frame survivors, world records, rasterised records, e1, e2, merged, - */
} ivm_status;

/* Bytes of device workspace a context with this config needs.  The caller
 * allocates it (e.g. a torch uint8 tensor), ZERO-FILLED, 256-byte aligned. */
size_t ivm_workspace_bytes(const ivm_config *cfg);

int ivm_create(const ivm_config *cfg, void *workspace_dev, size_t workspace_bytes, ivm_ctx **out);
int ivm_destroy(ivm_ctx *ctx);

/* Camera scale tables x_scale[width], y_scale[height] (projector/core.py:86-107),
 * device pointers, copied into the workspace on `stream`. */
int ivm_set_camera(ivm_ctx *ctx, const float *xs_dev, const float *ys_dev, ivm_stream_t stream);

/* One map update for `num_envs` envs (iterative / episodic mode).  All pointers
 * are device pointers:
 *   depth   f32 [B,H,W]   normalised depth (Observations.depth_normalized)
 *   labels  u8  [B,H,W]   GT labels, or NULL when `logits` is given
 *   logits  f32 [B,num_classes,H,W]  class scores (NCHW) or NULL; argmax'ed in-kernel
 *   labels_out u8 [B,H,W] receives the argmax labels (required with logits)
 *   T12     f32 [B,12]    rows 0..2 of the camera->world matrix (core.py:6-37)
 *   pose    f32 [B,3]     world_robot_pose
 *   cs      f32 [B,2]     cos(-heading), sin(-heading) as f32 (mapper.py:38-48,264-266)
 *   orientation [B,2]     (elevation, heading) as f64 or f32 (world_robot_orientation).  When
 *                         given, T12 and cs may be NULL: they are then derived on the device in
 *                         the angles' dtype (same libdevice sin/cos torch's CUDA ops use).
 *   masks   u8  [B]       not_done_masks; 0 wipes that env first (mapper.py:320-326)
 *   occ,sem u8  [B,R,C]   outputs (OccupancySemanticMapMemory.occupancy / .semantic)
 * Envs with index >= num_envs are dropped (mapper.py:315-318). */
int ivm_step_iterative(ivm_ctx *ctx, int32_t num_envs, const float *depth, const uint8_t *labels,
                       const float *logits, int32_t num_classes, uint8_t *labels_out, const float *T12,
                       const float *pose, const float *cs, const void *orientation, int32_t orientation_is_f64,
                       const uint8_t *masks, uint8_t *occ, uint8_t *sem, ivm_stream_t stream);

/* Known-map mode.  ivm_known_load replaces env `env`'s scene cloud
 * (xyz f32 [n,3], sem u8 [n], device pointers; list order = npz order);
 * origin_row/origin_col place the store window (absolute half-cell of store cell 0,0). */
int ivm_known_load(ivm_ctx *ctx, int32_t env, int64_t n, const float *xyz, const uint8_t *sem, int32_t origin_row,
                   int32_t origin_col, ivm_stream_t stream);
int ivm_known_clear(ivm_ctx *ctx, int32_t env, ivm_stream_t stream);
int ivm_step_known(ivm_ctx *ctx, int32_t num_envs, const float *pose, const float *cs, const void *orientation,
                   int32_t orientation_is_f64, uint8_t *occ, uint8_t *sem, ivm_stream_t stream);

/* Compacts the live world records into (env i64, xyz f32 x3, label u8, list key u64)
 * arrays of capacity `cap`; *count_dev (u64, device) receives the number of records.
 * Records come out grouped by env in (half-row, half-col) order; sorting by `key`
 * (stable) yields the reference's list order. */
int ivm_export_world(ivm_ctx *ctx, int32_t num_envs, int64_t cap, int64_t *env_out, float *xyz_out, uint8_t *label_out,
                     uint64_t *key_out, uint64_t *count_dev, ivm_stream_t stream);

/* Copies error flags and statistics of the last step to host memory; synchronises `stream`. */
int ivm_read_status(ivm_ctx *ctx, ivm_status *host_out, ivm_stream_t stream);

/* Error flags without a synchronisation: enqueues a 4-byte copy of the flags word into PINNED host memory on
 * `stream` (the caller looks at it once an event recorded behind the copy has completed); ivm_clear_error_flags
 * zeroes the word on `stream`.  The flags are sticky until cleared.  What the flags mean for the results:
 *   1  a point (or a known-map cloud point, 4) fell outside the env's store window: that point is DROPPED -- the
 *      reference's unbounded cloud would have kept it; enlarge store_rows / store_cols;
 *   2  the edge lists overflowed: collisions on the bounding-box edge may be unresolved (results invalid);
 *   8  a grid barrier timed out: the step's results are invalid and the context must be re-created. */
int ivm_copy_error_flags_async(ivm_ctx *ctx, uint32_t *host_pinned_out, ivm_stream_t stream);
int ivm_clear_error_flags(ivm_ctx *ctx, ivm_stream_t stream);

/* Per-kernel device timing: when enabled, CUDA events bracket each kernel of the next
 * steps; ivm_stage_times returns the accumulated milliseconds per stage since the last
 * reset (synchronises the events).  Stages: 0 prep, 1 ingest-scatter, 2 ingest-resolve,
 * 3 edge fix-up, 4 raster.  When a step runs as the single persistent kernel
 * (config.reserved[0] == 0 and the image tiles evenly), the whole step is booked under
 * stage 1 and stages 2-4 stay empty; ivm_read_phase_ns gives the split inside it. */
int ivm_set_timing(ivm_ctx *ctx, int32_t enabled);
int ivm_stage_times(ivm_ctx *ctx, float *ms_out5, int32_t *launches_out5, int32_t reset);

/* Persistent step kernel only: %globaltimer (ns) at the phase boundaries of the LAST step --
 * [0] start, [1] grid barrier 1 passed (depth scatter done), [2] grid barrier 2 passed (score stream and
 * resolve done), [3] edge fix-up done on CTA 0 (it runs beside the raster), [4] raster release issued,
 * [5] end (max over CTAs), [6..7] unused; [8..23] milestones inside the edge fix-up (start, stage-1 classes +
 * merges, bbox + segments, scan start, stage-2 start, stage-2 classes, end; rest unused).  [0..7] are all
 * zero if the last step took the multi-kernel path.  `ns_out24` has room for 24 values.  Synchronises `stream`. */
int ivm_read_phase_ns(ivm_ctx *ctx, uint64_t *ns_out24, ivm_stream_t stream);

/* Persistent step kernel only: per-CTA timeline of the LAST step, 16 %globaltimer values (ns) per CTA for the
 * first `num_ctas` CTAs (<= 1024): [7] kernel start, [1] slots prepared, [2] scatter passes done, [12] argmax
 * group's scatter share done, [10] arrived at grid barrier 1, [0] barrier 1 passed, [3] resolve done,
 * [6] argmax warps done, [5] barrier 2 passed, [11] first raster pass done, [8] released, [9] end; rest
 * unused.  Synchronises `stream`. */
int ivm_read_cta_trace(ivm_ctx *ctx, uint64_t *ns_out, int32_t num_ctas, ivm_stream_t stream);

/* Consumer-side epilogue: SemanticMapEncoder.generate_map_features
 * (ivlnce_baselines/models/encoders/map_encoder.py:85-90) = cat(occupancy, one_hot(semantic, num_classes)) as
 * float32 [B, 1 + num_classes, R, C].  occ, sem: u8 [B,R,C] (the outputs of a step); out: f32 device buffer.
 * Stateless (no context).  err_flag_dev (u32, device, may be NULL): bit 0 is set if a semantic value is
 * >= num_classes (F.one_hot raises there); the planes of such a cell are all zero. */
int ivm_map_features(const uint8_t *occ, const uint8_t *sem, int32_t num_envs, int32_t rows, int32_t cols,
                     int32_t num_classes, float *out, uint32_t *err_flag_dev, ivm_stream_t stream);

/* Segmentation front end: PredictSemantics.forward up to the network
 * (ivlnce_baselines/common/mapping_module/mapper.py:715-736, 788-793) as one kernel.  rgb: u8 device pointer to a
 * [B,3,rgb_h,rgb_w] tensor with ELEMENT strides rgb_strides4 (host array: batch, channel, row, column -- the sensor's
 * NHWC layout is strides {h*w*3, 1, w*3, 3}), or NULL; depth: f32 [B,1,height,width] or NULL.
 * rgb_out f32 [B,3,height,width] = ((bilinear resize of rgb / 255) - mean) / std, depth_out f32 [B,1,height,width] =
 * (depth - 0.213) / 0.285 (contiguous NCHW, what RedNet consumes).  Stateless.  Floating point: within 1e-5 of the
 * reference's torch ops (F.interpolate mode="bilinear", align_corners=False). */
int ivm_rednet_preprocess(const uint8_t *rgb, const int64_t *rgb_strides4, int32_t num_envs, int32_t rgb_h, int32_t rgb_w,
                          const float *depth, int32_t height, int32_t width, float *rgb_out, float *depth_out,
                          ivm_stream_t stream);

/* Number of kernels launched by this context so far. */
int64_t ivm_kernel_launches(const ivm_ctx *ctx);

/* Rewrites all live stamps to 1 (call when IVM_E_STEP_OVERFLOW is returned). */
int ivm_rebase_stamps(ivm_ctx *ctx, ivm_stream_t stream);

/* Pipelined stepping.  Off (default): a step touches nothing -- not even its inputs -- before the previous kernel
 * of the stream has completed.  On: the caller promises that the INPUT buffers of every ivm_step_iterative call
 * (depth, labels / logits, pose, angles, masks) are complete before the PREVIOUS step of this context was enqueued
 * (true for resident or replayed inputs, and for inputs staged by copy-engine transfers rather than kernels).
 * Consecutive steps enqueued back to back on one stream then overlap: the next step's score stream, pose matrices
 * and depth filter run while the last ego tiles of the current step are rastered, and the next step waits for the
 * current one through a counter of finished CTAs instead of the kernel boundary.  Results are identical. */
int ivm_set_pipelined(ivm_ctx *ctx, int32_t enabled);

/* Test hook: sets the 24-bit step counter of a context that has not stepped yet (so that a test can reach the
 * IVM_E_STEP_OVERFLOW / ivm_rebase_stamps path without 2^24 calls). */
int ivm_debug_set_step(ivm_ctx *ctx, uint32_t step);

/* Copies the whole map state of `src` into `dst` (same config except dst.max_envs >=
 * src.max_envs; dst freshly created over a zero-filled workspace).  Lets the host grow the
 * batch without losing the world stores (the reference's world cloud survives a batch
 * grow, mapper.py:146-162, 533-553). */
int ivm_copy_state(ivm_ctx *dst, const ivm_ctx *src, ivm_stream_t stream);

const char *ivm_last_cuda_error(const ivm_ctx *ctx);
const char *ivm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* IVLN_MAP_H */
