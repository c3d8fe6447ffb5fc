/* se3ds_geom.h -- C ABI of the B200-native geometric guidance path of SE3DS.
 *
 * The reference (google-research/se3ds) has no FFI: the path sits behind plain
 * Python functions.  Each entry point below names the reference function(s) it
 * replaces (file:line relative to the reference checkout).  The Python shims in
 * se3ds_b200/utils (pano_utils.py, point_cloud_utils.py) keep the reference signatures and call these through
 * ctypes with raw device pointers; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns an se3ds_status (0 = ok); no C++ exceptions cross
 *     the boundary; se3ds_last_error() gives a thread-local message.
 *   - all tensor pointers are DEVICE pointers unless the name ends in _host;
 *     tensors are dense, row-major, in the layouts of the reference.
 *   - the caller owns every input and output buffer; the library owns only the
 *     opaque workspace (z-buffer, feature buffer, per-point scratch, bins,
 *     angle tables).  A workspace is bound to one device and may be used by one
 *     stream at a time; distinct workspaces are independent (re-entrant).
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued
 *     asynchronously on it unless stated otherwise.
 *   - inputs must be finite; NaN/Inf depth or coordinates are rejected points.
 */
#ifndef SE3DS_GEOM_H_
#define SE3DS_GEOM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SE3DS_GEOM_VERSION 201 /* major*100 + minor; 200: round-2 ABI (apply_bin flags, ring / compact / expand / quantize entry points); 201: + SE3DS_FLAG_HOST_ASYNC, se3ds_ws_host_wait */

typedef struct se3ds_ws se3ds_ws;

typedef enum {
  SE3DS_OK = 0,
  SE3DS_ERR_BAD_SHAPE = 1, /* reference raises ValueError / AssertionError on these */
  SE3DS_ERR_BAD_DTYPE = 2, /* e.g. unsigned features with a negative void class   */
  SE3DS_ERR_BAD_ARG = 3,
  SE3DS_ERR_CUDA = 4,
  SE3DS_ERR_NOMEM = 5
} se3ds_status;

typedef enum { SE3DS_U8 = 0, SE3DS_I32 = 1, SE3DS_F32 = 2 } se3ds_dtype;

/* se3ds_reproject flags */
#define SE3DS_FLAG_FILTER_VOID 1u /* drop points whose channels all equal unproject_void:
                                     the compaction of models/models.py:229-237 (batch 1) */
#define SE3DS_FLAG_BIN_PER_JOB 2u /* every (item, pose) job behaves like its own reference call
                                     (batch 1): rejected points land on ITS pixel (0,0).  Default:
                                     the whole call is one reference call; the global reject bin
                                     (utils/point_cloud_utils.py:150-153) is job 0's pixel (0,0). */

#define SE3DS_FLAG_KEY64 4u       /* always splat with the 64-bit packed depth|point-index key.  By
                                     default that key is used when winner_out is requested; without
                                     it a 32-bit depth-only key gives the same guidance tensors
                                     with half the z-buffer traffic. */

#define SE3DS_FLAG_INPUTS_READY 8u /* promise: rgb / depth / positions were complete before the kernel
                                      that precedes this call on the stream was launched (static or
                                      long-uploaded buffers).  Lets the projection math of this call
                                      overlap the tail of that kernel (programmatic dependent launch);
                                      without the flag the call waits for it first. */

#define SE3DS_FLAG_RAW_FEATURES 16u /* proj_image receives the raw per-channel maxima (the
                                       `projected_feat` of point_cloud_utils.py:173-178) instead of
                                       clip(x / 255, 0, 1): e.g. semantic class ids replicated into
                                       the three channels (models/models.py:276-278). */

#define SE3DS_FLAG_COMPACT_OUT 32u /* compact guidance: proj_image points to a UINT8 (J,H,W,3) plane that
                                     receives the per-channel maxima clamped to [0, 255], proj_depth stays
                                     float32, proj_mask may be NULL and is not written -- 7 instead of 20
                                     bytes per pixel (device->host copies, the NCCL all-gather).  Nothing
                                     is lost: clip(x / 255, 0, 1) is a function of that byte and
                                     mask = 0 < depth < 1; se3ds_expand_guidance restores the float32
                                     tensors bit for bit.  Not combinable with SE3DS_FLAG_RAW_FEATURES. */
#define SE3DS_FLAG_HOST_ASYNC 64u /* se3ds_reproject_host only: return as soon as the copies and kernels are
                                     enqueued on the workspace's own streams; se3ds_ws_host_wait completes the
                                     call.  Two workspaces used in turn keep the host link busy in both
                                     directions: batch i+1 travels to the device while batch i travels back.
                                     The batch then moves in one DMA transfer per tensor (the blocking call
                                     pipelines item by item inside the call instead). */

int se3ds_version(void);
const char* se3ds_status_string(int status);
const char* se3ds_last_error(void);

/* Workspace.  max_bytes bounds the device memory the workspace may hold (0 = default 2 GiB);
 * larger calls are processed in job chunks.  l2_chunk_bytes is the preferred size of one chunk's
 * working set (0 = default, tuned for the 126 MB L2). */
int se3ds_ws_create(int device, size_t max_bytes, size_t l2_chunk_bytes, se3ds_ws** out);
int se3ds_ws_destroy(se3ds_ws* ws);
int se3ds_ws_bytes(const se3ds_ws* ws, size_t* bytes);

/* Projection mode of the fused path.  0: every point takes the canonical (IEEE-division)
 * projection; 1 (default): certified fast path -- MUFU approximations, accepted only when the column
 * coordinate is farther than the margin from an integer and z / rad lies inside the row's cosine
 * interval shrunk by the margin, canonical fallback otherwise, so the results are the canonical ones
 * bit for bit; 2: verify -- both are evaluated and compared (slow; tests).  margin_scale > 0 overrides
 * the margin (dx = W*scale pixels, dy = 2*H*scale pixels = 2*pi*scale radians; default 1e-6 -- the
 * measured deviation of the fast path is a tenth of that; smaller margins are for experiments in
 * verify mode and void the bit-exactness of mode 1).  se3ds_ws_verify_read returns {points, certified,
 * certified-but-different} and the largest distance by which a certified fast coordinate fell outside
 * its canonical pixel (x, y; in pixels) since the last read. */
int se3ds_ws_projection_mode(se3ds_ws* ws, int mode, float margin_scale);

/* Programmatic dependent launch between the kernels of the fused path (default on): the next
 * kernel's blocks are scheduled into the tail of the running one and wait on-device for its
 * completion (griddepcontrol), instead of the stream serialising whole grids. */
int se3ds_ws_pdl(se3ds_ws* ws, int enable);

/* Concurrent chunk lanes (1..4, default 2): the job chunks of one se3ds_reproject call are dealt
 * round-robin to `lanes` streams (the caller's stream plus internal ones, forked and joined with
 * events, so the call still behaves as one unit of work on the caller's stream).  The projection
 * kernel is bound by instruction issue, the other two by memory: chunks in different phases overlap.
 * Every lane has its own slice of the workspace; the L2 budget is shared between the lanes.  A call
 * uses fewer lanes when it cannot give each one `min_chunks_per_lane` chunks (0 = default, 2) of at
 * least `min_points_per_lane` source points (0 = default, 2^20); se3ds_reproject_host and profiled
 * calls run on one lane. */
int se3ds_ws_lanes(se3ds_ws* ws, int lanes, long long min_points_per_lane, int min_chunks_per_lane);

/* The chunk / lane plan se3ds_reproject would use for a call of this shape with these knobs (0 = the
 * defaults of se3ds_ws_create / se3ds_ws_lanes); pure host arithmetic, no device needed.
 * plan = {lanes used, batch items per chunk, poses per chunk, jobs per chunk, number of chunks}. */
int se3ds_plan_chunks(size_t l2_chunk_bytes, int lanes, long long min_points_per_lane, int min_chunks_per_lane,
                      int n, int s, int p, int h, int w, long long plan[5]);
int se3ds_ws_verify_read(se3ds_ws* ws, unsigned long long counts[3], float max_dev[2]);

/* Measurement hooks (bench.py).  mode 1: se3ds_reproject brackets its three kernel classes with cudaEvents
 * on the caller's stream (programmatic dependent launch is switched off meanwhile, so the kernels run
 * back to back without overlap); se3ds_ws_profile_read synchronises, returns the accumulated device
 * milliseconds of {splat_depth, splat_feat, resolve} since the last read and the number of kernels this
 * workspace has launched so far (counted always, profiling or not).
 * mode 2: the pipeline runs as it does in production (programmatic dependent launch on, lanes on); every
 * block stamps %globaltimer when it ends, and a kernel's share of the step is its last stamp minus the
 * last stamp of the kernel launched before it.  se3ds_ws_profile_read_stamps synchronises and returns the
 * summed shares in milliseconds over `chunks` job chunks (the first chunk after enabling has no
 * predecessor and is skipped; at most 2048 chunks are recorded per read).  mode 0: off. */
int se3ds_ws_profile(se3ds_ws* ws, int mode);
int se3ds_ws_profile_read(se3ds_ws* ws, float ms[3], unsigned long long* launches);
int se3ds_ws_profile_read_stamps(se3ds_ws* ws, double ms[3], long long* chunks);

/* utils/pano_utils.py:245-265  mask_pano(pano, proportion, masked_region_value).
 * pano/out (N,H,W,C) of `dtype`; rows r < int(H*p) or r > H - int(H*p) become the value. */
int se3ds_mask_pano(const void* pano, int dtype, int n, int h, int w, int c, double proportion,
                    double masked_region_value, void* out, void* stream);

/* utils/pano_utils.py:164-242  equirectangular_to_pointcloud (size_mult == 1).
 * feats (N,H,W,C) in_dtype, depth (N,H,W) f32 -> xyz1 (N,4,H*W) f32 planar, feats_out (N,H*W,C)
 * out_dtype ('nearest' keeps the dtype, 'bilinear' -> f32).  size_mult != 1: resize first with
 * se3ds_resize, then pass the scaled (h, w). */
int se3ds_unproject_equirect(se3ds_ws* ws, const void* feats, int in_dtype, const float* depth, int n,
                             int h, int w, int c, double void_class, float depth_scale,
                             float* xyz1_out, void* feats_out, int out_dtype, void* stream);

/* utils/pano_utils.py:117-161 project_feats_to_equirectangular (mode 0: coords are cartesian) and
 * utils/point_cloud_utils.py:90-183 project_to_feat (mode 1: coords are "transformed_coords").
 * coords (N,4,M) f32, feats (N,M,C) feat_dtype (cast to f32 like the reference) ->
 * depth_out (N,H,W) f32 in [0,1], feats_out (N,H,W,C) f32, winner_out (N,H,W) int32 or NULL
 * (index m of the nearest point, lowest index on ties, -1 = none).  M may be 0. */
int se3ds_project_cloud(se3ds_ws* ws, const float* coords, const void* feats, int feat_dtype, int n,
                        long long m, int c, int h, int w, int mode, float input_void_class,
                        float output_void_class, float depth_scale, float* depth_out,
                        float* feats_out, int32_t* winner_out, void* stream);

/* The fused path: replaces, for S source frames and P target poses per batch item,
 *   mask_pano (pano_utils.py:245-265, first `mask_frames` frames) ->
 *   equirectangular_to_pointcloud (pano_utils.py:164-242) -> xyz1 += src_pos ->
 *   [compaction, models.py:229-237] -> concat over frames -> coords - tgt_pos
 *   (models.py:225-226,273-275; gan_manager.py:474-475,549) ->
 *   project_feats_to_equirectangular (pano_utils.py:117-161 + point_cloud_utils.py:90-183) ->
 *   guidance assembly (models.py:282-293; gan_manager.py:484-494)
 * without materialising the cloud.
 *   rgb (N,S,H,W,3) rgb_dtype (U8, or I32 with values in [-2048, 2048] -- the per-channel maxima are reduced in
 *   float16, exact for those integers; the reference only produces [-1, 255]; the Python shim refuses anything
 *   else); depth (N,S,H,W) f32;
 *   src_pos (N,S,3) f32; tgt_pos (N,P,3) f32 (device).  Jobs are (n,p) row-major, J = N*P.
 *   proj_image (J,H,W,3) f32 = clip(rgb/255,0,1); proj_depth (J,H,W,1) f32; proj_mask (J,H,W,1)
 *   f32 in {0,1}; winner_out (J,H,W) int32 or NULL: index s*H*W + r*W + c of the nearest valid
 *   point of that pixel (lowest index on ties), -1 = none.
 *   unproject_void: feature given to depth-invalid pixels; project_void: feature value that
 *   makes a point invalid (any channel).  SE3DSModel: -1/-1 + FILTER_VOID; gan_manager: 0/-1;
 *   eval_metric: -1/-1. */
int se3ds_reproject(se3ds_ws* ws, const void* rgb, int rgb_dtype, const float* depth,
                    const float* src_pos, const float* tgt_pos, int n, int s, int p, int h, int w,
                    float depth_scale, double mask_proportion, int mask_frames, int unproject_void,
                    int project_void, unsigned flags, float* proj_image, float* proj_depth,
                    float* proj_mask, int32_t* winner_out, float* bin_out, void* stream);

/* Full SE(3) target poses (the reference only translates points and rotates
 * afterwards in image space, utils/pano_utils.py:306-341): same as se3ds_reproject, but every point
 * is additionally rotated into the target camera frame, q = R[n,p] * ((local + src) - tgt), with
 * tgt_rot (N,P,3,3) f32 row-major on the device.  Canonical arithmetic per row:
 * fma(r2, z, fma(r1, y, r0 * x)).  tgt_rot == NULL is se3ds_reproject bit for bit. */
int se3ds_reproject_se3(se3ds_ws* ws, const void* rgb, int rgb_dtype, const float* depth,
                        const float* src_pos, const float* tgt_pos, const float* tgt_rot, int n, int s,
                        int p, int h, int w, float depth_scale, double mask_proportion, int mask_frames,
                        int unproject_void, int project_void, unsigned flags, float* proj_image,
                        float* proj_depth, float* proj_mask, int32_t* winner_out, float* bin_out,
                        void* stream);

/* The memory of a trajectory as a frame RING (models/models.py:127-152,239-245 and the rollout loops
 * trainers/gan_manager.py:462-463,550-551, utils/eval_metric.py:146-147,238-239 grow a concatenated
 * cloud by tf.concat every frame: O(M) per step).  Same as se3ds_reproject_se3, but rgb / depth /
 * src_pos are allocated for `s_capacity` frames per item -- (N,s_capacity,H,W,3), (N,s_capacity,H,W),
 * (N,s_capacity,3) -- of which the first `s` are in use: a caller appends a frame by writing slot s in
 * place and calling again with s + 1, nothing is copied or re-stacked. */
int se3ds_reproject_ring(se3ds_ws* ws, const void* rgb, int rgb_dtype, const float* depth,
                         const float* src_pos, const float* tgt_pos, const float* tgt_rot, int n, int s,
                         int s_capacity, int p, int h, int w, float depth_scale, double mask_proportion,
                         int mask_frames, int unproject_void, int project_void, unsigned flags,
                         float* proj_image, float* proj_depth, float* proj_mask, int32_t* winner_out,
                         float* bin_out, void* stream);

/* Feedback of a generated frame into the memory (trainers/gan_manager.py:539-542,
 * utils/eval_metric.py:227-230): out = clip_by_value(cast(image * 255, int32), -1, 255), cast truncating
 * toward zero.  image: n dense items of elems_per_item float32 (= H*W*3); item i is written to
 * out + i * out_item_stride (int32 elements), e.g. frame slot s of an (N,S_cap,H,W,3) ring. */
int se3ds_quantize_rgb(const float* image, int n, long long elems_per_item, int32_t* out, long long out_item_stride,
                       void* stream);

/* Compact guidance (SE3DS_FLAG_COMPACT_OUT) -> the float32 tensors of models/models.py:282-293:
 * proj_image (njobs,px_per_job,3) = clip(rgb_u8 / 255, 0, 1), proj_mask (njobs,px_per_job) = 0 < depth < 1.
 * Bit-identical to what se3ds_reproject writes without the flag.  job_map (device, njobs int32, or NULL =
 * identity): source job s is written as destination job job_map[s] -- a gather buffer that arrived piece by
 * piece (se3ds_b200/parallel.py) is put into job order on the way; depth_out (or NULL) receives the depth
 * plane in destination order. */
int se3ds_expand_guidance(const uint8_t* rgb_u8, const float* proj_depth, long long njobs, long long px_per_job,
                          const int32_t* job_map, float* proj_image, float* depth_out, float* proj_mask, void* stream);

/* Multi-GPU support for the global reject bin.  When se3ds_reproject is given bin_out (device,
 * 5 floats) the call's reject bin is NOT applied to job 0's pixel (0,0) but exported as
 * (min depth or +inf, max R, max G, max B, depth of that pixel's own winner or +inf); ranks reduce
 * the first four (min / max), the owner of global job 0 keeps its own fifth value and applies the
 * result with se3ds_apply_bin.  clip() and the divisions are monotone, so patching the finished
 * outputs is bit-identical to the single-call result; `winner` (job 0's winner plane, may be NULL)
 * gets -1 at the pixel when a rejected point is nearer than its own winner.  `flags`: pass
 * SE3DS_FLAG_RAW_FEATURES / SE3DS_FLAG_COMPACT_OUT when the outputs were produced with them (proj_mask
 * may then be NULL). */
int se3ds_apply_bin(const float* bin, float depth_scale, unsigned flags, float* proj_image, float* proj_depth,
                    float* proj_mask, int32_t* winner, void* stream);

/* Same as se3ds_reproject with HOST buffers (pinned memory recommended): copies the inputs to the
 * device, runs the fused path and copies the guidance tensors back, pipelined over groups of batch items
 * (four stages) on the workspace's own streams.  Blocks until the outputs are complete, unless SE3DS_FLAG_HOST_ASYNC is
 * given: the call then returns once everything is enqueued, the host buffers (inputs and outputs) belong
 * to the library until se3ds_ws_host_wait(ws) returns, and a further host call on the same workspace
 * waits for the pending one first. */
int se3ds_reproject_host(se3ds_ws* ws, const void* rgb_host, int rgb_dtype, const float* depth_host,
                         const float* src_pos_host, const float* tgt_pos_host, int n, int s, int p,
                         int h, int w, float depth_scale, double mask_proportion, int mask_frames,
                         int unproject_void, int project_void, unsigned flags,
                         float* proj_image_host, float* proj_depth_host, float* proj_mask_host,
                         int32_t* winner_out_host);

/* Completes a pending SE3DS_FLAG_HOST_ASYNC call of this workspace (no-op when nothing is pending). */
int se3ds_ws_host_wait(se3ds_ws* ws);

/* tf.image.resize with half-pixel centres, as utils/pano_utils.py:203-208 uses it when
 * size_mult != 1: in (N,H,W,C) of `dtype` -> out (N,out_h,out_w,C); bilinear == 0: 'nearest', the
 * dtype is kept; bilinear != 0: 'bilinear', out is f32. */
int se3ds_resize(const void* in, int dtype, int n, int h, int w, int c, int out_h, int out_w, int bilinear,
                 void* out, void* stream);

/* tensorflow_addons.image.interpolate_bilinear as the reference calls it (utils/pano_utils.py:339
 * rotate_pano, :412 project_perspective_image, :472 get_perspective_from_equirectangular_image;
 * SURVEY 8f rank 2).  grid (B,H,W,C) f32, query_points (B,Nq,2) f32 in (y,x) order, or (x,y) when
 * indexing_xy != 0 -> out (B,Nq,C) f32. */
int se3ds_interpolate_bilinear(const float* grid, const float* query_points, int b, int h, int w, int c,
                               long long num_queries, int indexing_xy, float* out, void* stream);

/* utils/point_cloud_utils.py:32-87 get_filtered_coords_and_feats (legacy perspective unprojection;
 * only the reference's tests call it).  feats (N,H,W,C) dtype, depth (N,H,W) f32 -> xyz (N,4,H*W) f32,
 * feats_out (N,H*W,C) f32; kinv_x / kinv_y = the diagonal of inv(get_intrinsic_matrix(HFOV)). */
int se3ds_filtered_coords_and_feats(const void* feats, int dtype, const float* depth, int n, int h, int w, int c,
                                    float depth_scale, float kinv_x, float kinv_y, float* xyz_out, float* feats_out,
                                    void* stream);

/* Fused forms of the resampling functions (coordinates computed in-kernel, then the same bilinear
 * sample).  All tensors f32 on the device; the 3x3 matrices of the last two are HOST arrays
 * (row-major), computed by the caller exactly as the reference computes them.
 *   se3ds_pixel_rays                 utils/pano_utils.py:92-114  -> out (3, H*2H)
 *   se3ds_rotate_pano                utils/pano_utils.py:306-341 pano (N,H,W,C), matrix (N,3,3) device
 *                                    -> out (N,OH,2*OH,C)
 *   se3ds_project_perspective_image  utils/pano_utils.py:344-417 image (h,w,C), world_to_image =
 *                                    get_world_to_image_transform(...); pad = 1 for pad_mode
 *                                    'constant' / 'mean' (pad_value = constant or image mean), 0 for
 *                                    'reflect' -> out (OH,2*OH,C)
 *   se3ds_perspective_from_equirect  utils/pano_utils.py:443-476 image (EH,EW,C), kinv_t = inv(K)^T,
 *                                    rotation -> out (height,width,C) */
int se3ds_pixel_rays(int output_height, float* out, void* stream);
int se3ds_rotate_pano(const float* pano, const float* matrix, int n, int h, int w, int c, int output_height,
                      float* out, void* stream);
int se3ds_project_perspective_image(const float* image, int h, int w, int c, const float world_to_image[9],
                                    int output_height, int pad, float pad_value, int round_to_nearest, float* out,
                                    void* stream);
int se3ds_perspective_from_equirect(const float* image, int eq_h, int eq_w, int c, const float kinv_t[9],
                                    const float rotation[9], int height, int width, float* out, void* stream);

/* inference/perturbation_utils.py:23-71 get_proportion_invalid_for_depth, batched over P offsets.
 * offsets (P,3) f32 device, depth (H,W) f32 device -> out (P,) f32 device. */
int se3ds_proportion_invalid(const float* offsets, int p, const float* depth, int h, int w,
                             float distance_padding, float depth_scale, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SE3DS_GEOM_H_ */
