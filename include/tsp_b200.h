/*
 * tsp_b200.h - C ABI of the B200-native surface-projection library (libtsp_b200.so).
 *
 * One entry point per call the reference makes on its projection hot path.  The reference is
 * pure Python (no FFI of its own); the interface each function replaces is cited as
 * file:line of kasirershahartau/tissue_image_processing:
 *
 *   SP  = tissue_analyzing_tool/surface_projection.py
 *   SPM = tissue_analyzing_tool/surface_proj_m.py
 *   BIM = tissue_analyzing_tool/basic_image_manipulations.py
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * TSP_ERR_* code and never throws; tsp_last_error() gives the text for the calling thread.
 * "d_" pointers are device memory on the handle's GPU, "h_" pointers are host memory (pinned
 * host memory makes the copies asynchronous DMA; an input stack in pageable memory is staged by
 * the library through pinned chunks with a pool of host threads - TSP_COPY_THREADS, default the
 * CPUs of the process / LOCAL_WORLD_SIZE, at most 16 - and the submitting call returns when the
 * last chunk has been handed to the DMA engine).  Stacks are uint16, C-contiguous (C, Z, Y, X) with X fastest - the layout the
 * reference operator works on after BIM:199-231 / SP:21-26.
 */
#ifndef TSP_B200_H
#define TSP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSP_ABI_VERSION 3

/* error codes */
#define TSP_OK 0
#define TSP_ERR_INVALID (-1)      /* bad argument (shape, channel, mode ...)                      */
#define TSP_ERR_CUDA (-2)         /* CUDA runtime failure, text in tsp_last_error()               */
#define TSP_ERR_WORKSPACE (-3)    /* workspace smaller than tsp_project_workspace_bytes()         */
#define TSP_ERR_BAND_INDEX (-4)   /* height map indexes past the cropped stack: the reference     */
                                  /* raises IndexError at SP:68-69 (min_z>0 or atoh_shift)        */
#define TSP_ERR_CHOOSE_LIMIT (-5) /* surface_projection_m with 64 or more planes: np.choose      */
                                  /* raises ValueError at SPM:40                                  */

/* score-stage variants (all feed the same argmax + band projection) */
#define TSP_MODE_FAST 0      /* multirate sigma=30 stage, one uint16 read; |score err| ~1e-5 rel  */
#define TSP_MODE_EXACT 1     /* direct separable FIR, fp32 accumulate, reference pass order       */
#define TSP_MODE_BITEXACT 2  /* direct FIR, fp64 accumulate in scipy's summation order, f32 store */

typedef struct tsp_handle tsp_handle; /* per-GPU context: tables, status words, staging buffers */

/* score methods for bin_size > 1 (SP:39-53) */
#define TSP_METHOD_MAX_AVERAGES 0  /* block mean of the sigma=30 score                            */
#define TSP_METHOD_MAX_STD 1       /* block variance of the pre-blurred reference channel         */
#define TSP_METHOD_MULTI_CHANNEL 2 /* block variance (reference) x block mean (next channel)      */

/* desc.flags.  By default the ten kernels of a frame are chained with programmatic dependent launches: each is
 * queued while its predecessor drains and only waits (griddepcontrol.wait) before touching memory - the lowest
 * latency for ONE stack (0.372 -> 0.351 ms at 2048x2048x64).  When frames of other streams are in flight on the same
 * GPU the early-queued CTAs only take SM slots from them: set TSP_FRAME_CONCURRENT and the kernels launch the
 * ordinary way (tsp_frame_submit sets it by itself while another slot is busy). */
#define TSP_FRAME_CONCURRENT 1
/* Host-buffer calls only (tsp_frame_submit / tsp_project_frame_host): the outputs are converted on the device to what
 * the movie driver stores - projection and height map as uint16, the C cast numpy's astype("uint16") performs
 * (BIM:481, SP:229-231) - so a frame returns 2 + 2 bytes per pixel instead of 8 + 8.  h_proj / h_zmap then point at
 * uint16 arrays. */
#define TSP_FRAME_OUT_U16 2

/* The constants the reference hard-codes (SP:28, SP:35, SP:37, SP:55, SP:70-71).  tsp_default_params() fills the
 * reference's values; a frame descriptor with has_params == 0 uses them too.  Non-default sigmas leave the
 * fast-mode tables (built for sigma 1 / 30 / (1,2,2)): the score then runs through the direct FIR (TSP_MODE_FAST
 * behaves like TSP_MODE_EXACT) and a non-default sigma_mask through the materialised band mask. */
typedef struct tsp_params {
    float percentile;       /* SP:35   95: clip at this percentile of the non-zero reference voxels, [0, 100]   */
    int32_t pedestal;       /* SP:28   10000: subtracted (and clamped at 0) when desc.airyscan != 0             */
    float sigma_pre[3];     /* SP:37   (0.5, 1, 1)   z, y, x                                                    */
    float sigma_score[3];   /* SP:55   (0.5, 30, 30)                                                            */
    float sigma_mask[3];    /* SP:70   (1, 2, 2): the band whose weighted maximum is projected                  */
    int32_t reserved[5];    /* must be 0                                                                        */
} tsp_params;

/* Frame descriptor = the arguments of time_point_surface_projection (SP:17-19) that reach the
 * arithmetic.  bin_size <= 1 and build_manifold == 0 (the zeroed defaults) select the plain
 * operator; bin_size > 1 bins the score with `method` (SP:39-53) and resamples it with order-1
 * interpolation (SP:59-65); build_manifold != 0 replaces the argmax by the region growing of
 * SP:87-165 (at most 254 planes).  Both run the direct-FIR score (TSP_MODE_FAST behaves like
 * TSP_MODE_EXACT for them). */
typedef struct tsp_frame_desc {
    int32_t channels, planes, rows, cols; /* C, Z, Y, X of the uint16 stack                      */
    int32_t reference_channel;            /* SP:32                                               */
    int32_t min_z, max_z;                 /* SP:30-31: crop [min_z, max_z) when max_z > 0        */
    int32_t airyscan;                     /* SP:27-29: subtract 10000, clamp at 0                */
    int32_t atoh_shift;                   /* SP:62: plane offset for the non-reference channels  */
    int32_t mode;                         /* TSP_MODE_*                                          */
    int32_t bin_size;                     /* SP:39: <= 1 none                                    */
    int32_t method;                       /* TSP_METHOD_* (only read when bin_size > 1)          */
    int32_t build_manifold;               /* SP:56-57                                            */
    int32_t flags;                        /* TSP_FRAME_* (0 = defaults)                           */
    int32_t has_params;                   /* 0: reference constants; 1: `params` below            */
    int32_t reserved;                     /* must be 0                                           */
    tsp_params params;                    /* read only when has_params != 0                      */
} tsp_frame_desc;

/* What the operator learned about the frame (filled by the *_host calls and tsp_get_frame_status) */
typedef struct tsp_frame_status {
    int32_t band_index_error; /* 1 when the reference would raise IndexError (SP:68-69)          */
    int32_t has_nonzero;      /* 0 when the reference channel is all zero after SP:27-29 (no clip)*/
    float percentile95;       /* clip value of SP:35 (undefined when has_nonzero == 0)           */
    int32_t zmap_min, zmap_max;
    int64_t nonzero_count;    /* voxels > 0 of the reference channel after SP:27-31              */
    int32_t near_tie_pixels;  /* fast mode: pixels whose top-2 score gap is below 2.5e-4 rel.    */
    int32_t reserved[5];
} tsp_frame_status;

int tsp_abi_version(void);
const char* tsp_last_error(void);

void tsp_default_params(tsp_params* out);

int tsp_create(int device, tsp_handle** out);
int tsp_destroy(tsp_handle* h);

/* ---- the operator: SP:17-85 (time_point_surface_projection) ------------------------------- */

/* Bytes of device scratch tsp_project_frame needs for this frame shape and mode. */
size_t tsp_project_workspace_bytes(const tsp_frame_desc* desc);

/* Device-resident call.  Repeated calls with the same descriptor and buffers (a frame slot, a movie loop) replay the
 * frame's launch sequence as a CUDA graph: one launch per frame instead of a dozen.
 * d_stack (C,Z,Y,X) uint16; d_proj (C,Y,X) float32 = SP:72-79 (the
 * reference's float64 values are float32-exact); d_zmap (Y,X) int32 = chosen_z of SP:61
 * (min_z already added).  Stream ordered, returns without synchronising; call
 * tsp_get_frame_status() afterwards to learn about TSP_ERR_BAND_INDEX. */
int tsp_project_frame(tsp_handle* h, const tsp_frame_desc* desc, const uint16_t* d_stack,
                      float* d_proj, int32_t* d_zmap, void* d_workspace, size_t workspace_bytes,
                      void* cuda_stream);

/* Synchronises the stream and reports on the last tsp_project_frame issued with d_workspace.
 * Returns TSP_ERR_BAND_INDEX when the reference would have raised IndexError. */
int tsp_get_frame_status(tsp_handle* h, const void* d_workspace, void* cuda_stream,
                     tsp_frame_status* out);

/* Host-buffer call = the plugin boundary `result = apply_function(chunk, **params)` of BIM:129.
 * h_stack (C,Z,Y,X) uint16 in host memory; h_proj (C,Y,X) float64 and h_zmap (Y,X) int64 are
 * the dtypes the reference returns (SP:73, SP:61) - uint16 both with TSP_FRAME_OUT_U16.  Copies in, runs,
 * copies out, synchronises.  Device scratch is owned (and reused) by the handle; the call has a frame slot
 * of its own, so it may be issued from any thread, also while a slot pipeline is running (calls from several
 * threads take turns). */
int tsp_project_frame_host(tsp_handle* h, const tsp_frame_desc* desc, const uint16_t* h_stack,
                           void* h_proj, void* h_zmap, tsp_frame_status* status);

/* Pipelined form of the same call for movies (SP:205-212 projects the time points of a movie one by
 * one; they are independent).  A handle owns TSP_MAX_SLOTS frame slots, each with its own stream
 * and device memory: tsp_frame_submit enqueues copy-in, operator and copy-out of one frame and
 * returns at once; tsp_frame_wait blocks until that slot's outputs are in h_proj / h_zmap.  With
 * pinned host buffers the copy-in of frame t+1 overlaps the kernels of frame t.
 * tsp_project_frame_host == submit + wait on slot 0. */
#define TSP_MAX_SLOTS 4
int tsp_frame_submit(tsp_handle* h, int slot, const tsp_frame_desc* desc, const uint16_t* h_stack,
                     void* h_proj, void* h_zmap);
int tsp_frame_wait(tsp_handle* h, int slot, tsp_frame_status* status);

/* ---- building blocks (also used one by one by the parity tests) --------------------------- */

/* BIM:373-390 blur_image = scipy.ndimage.gaussian_filter(mode='nearest') on a 3-D float32
 * volume (z,y,x), passes z -> y -> x with a float32 store between passes.  fp64_accumulate != 0
 * reproduces scipy's float64 line sums bit for bit.  d_tmp: scratch of the same size. */
int tsp_gaussian_blur_f32(tsp_handle* h, const float* d_in, float* d_out, float* d_tmp,
                          int planes, int rows, int cols, const double sigma[3],
                          int fp64_accumulate, void* cuda_stream);

/* Same on uint16 with scipy's integer-output behaviour: every pass truncates toward zero
 * (SPM:18 blurs the uint16 stack). */
int tsp_gaussian_blur_u16(tsp_handle* h, const uint16_t* d_in, uint16_t* d_out, uint16_t* d_tmp,
                          int planes, int rows, int cols, const double sigma[3],
                          void* cuda_stream);

/* SP:32-36: 95th percentile (numpy 'linear', float32 index arithmetic) of the voxels that are
 * > 0 after the optional airyscan pedestal.  Synchronises; results through `out`. */
int tsp_percentile95_nonzero_u16(tsp_handle* h, const uint16_t* d_volume, size_t count,
                                 int airyscan, void* cuda_stream, tsp_frame_status* out);
/* The same for any percentile in [0, 100] and pedestal (0 = none). */
int tsp_percentile_nonzero_u16(tsp_handle* h, const uint16_t* d_volume, size_t count, float percentile,
                               int pedestal, void* cuda_stream, tsp_frame_status* out);

/* SP:26-37 + SP:55: the focus score volume (planes,rows,cols) float32 of one channel, in the
 * exact / bit-exact variants (the fast variant never materialises it).  d_tmp: same size. */
int tsp_focus_score_f32(tsp_handle* h, const uint16_t* d_channel, float* d_score, float* d_tmp,
                        int planes, int rows, int cols, int airyscan, int fp64_accumulate,
                        void* cuda_stream);

/* SP:61: first-maximum argmax over z (+ z_offset). */
int tsp_argmax_z_f32(tsp_handle* h, const float* d_score, int32_t* d_zmap, int planes, int rows,
                     int cols, int z_offset, void* cuda_stream);

/* SP:87-165 build_continues_manifold: d_chosen (rows,cols) int32 = the height map grown in square
 * rings around the global maximum of d_score (planes,rows,cols), every pixel within +-1 plane of its
 * already-placed neighbours (find_pixel_plane, quirks included).  planes <= 254. */
size_t tsp_manifold_workspace_bytes(void);
int tsp_build_manifold(tsp_handle* h, const float* d_score, int32_t* d_chosen, int planes, int rows,
                       int cols, void* d_workspace, size_t workspace_bytes, void* cuda_stream);

/* SP:41-50 skimage.measure.block_reduce(volume, (1,b,b), np.mean | np.var): zero padded to a multiple
 * of bin_size; d_out (planes, ceil(rows/b), ceil(cols/b)) float32. */
int tsp_block_reduce_f32(tsp_handle* h, const float* d_volume, float* d_out, int planes, int rows,
                         int cols, int bin_size, int variance, void* cuda_stream);

/* SP:59-61 resize(score, (Z,Y,X)) [order 1, reflect; = scipy.ndimage.zoom(order=1, mode='mirror',
 * grid_mode=True)] fused with the first-maximum argmax over z (+ z_offset).  d_score is the binned
 * (planes, coarse_rows, coarse_cols) volume; d_workspace >= 256 bytes. */
int tsp_resize_argmax_f32(tsp_handle* h, const float* d_score, int32_t* d_zmap, int planes, int rows,
                          int cols, int coarse_rows, int coarse_cols, int z_offset, void* d_workspace,
                          size_t workspace_bytes, void* cuda_stream);

/* SP:63-65 np.round(resize(chosen_z.astype('float32'), (Y,X))).astype('int') of a coarse height map. */
int tsp_resize_round_i32(tsp_handle* h, const int32_t* d_coarse, int32_t* d_zmap, int rows, int cols,
                         int coarse_rows, int coarse_cols, void* cuda_stream);

/* SP:62-81: band mask from a height map and weighted max projection of every channel, without
 * materialising the one-hot volume.  d_zmap indexes the (already cropped) stack.  Channels other
 * than reference_channel use clip(zmap + atoh_shift, 0, planes).  Out-of-range indices set the
 * handle's band_index_error status (see tsp_get_frame_status). */
int tsp_band_project(tsp_handle* h, const uint16_t* d_stack, const int32_t* d_zmap, float* d_proj,
                     int channels, int planes, int rows, int cols, int reference_channel,
                     int atoh_shift, int airyscan, void* d_workspace, size_t workspace_bytes,
                     void* cuda_stream);
size_t tsp_band_workspace_bytes(int channels, int planes, int rows, int cols);

/* ---- SPM:14-35 surface_projection_m ------------------------------------------------------- */
/* d_channel: (planes,rows,cols) uint16 = image[reference_channel][min_z:max_z]; method 0 =
 * "max_averages" (block mean), 1 = "max_std" (block variance); d_out (rows,cols) uint16. */
size_t tsp_project_m_workspace_bytes(int planes, int rows, int cols, int bin_size);
int tsp_project_m(tsp_handle* h, const uint16_t* d_channel, uint16_t* d_out, int planes, int rows,
                  int cols, int method, int bin_size, void* d_workspace, size_t workspace_bytes,
                  void* cuda_stream);

/* ---- introspection used by tests ---------------------------------------------------------- */
/* Copies the fast-mode coarse FIR taps c[0..n) (symmetric half, c[0] = centre) designed by the
 * library; returns the number of taps, or a negative error. */
int tsp_debug_coarse_taps(double* out, int capacity);
/* Number of CUDA kernels launched through this handle so far (bench.py reports it). */
int64_t tsp_launch_count(const tsp_handle* h);
/* Kernel-variant switches for tests and A/B measurements (the product path never needs them; they replace the
 * environment variables of ABI 2, which were read on every launch).  Keys: "no_ring" (strip decimation instead of
 * the TMA ring), "band_variant" (0 auto, 2 register-prefetch kernel, 3 TMA ring kernel for every tile), "no_chain"
 * (plain launches everywhere), "interp_rows" (2, 4 or 8 image rows per thread of the interpolation stage),
 * "interp_global" (interpolation stage with its control points read from L2 instead of staged in shared memory),
 * "graphs" (0: never replay a frame as a CUDA graph).  The same keys are read ONCE at tsp_create from the
 * environment as TSP_NO_RING, TSP_BAND_VARIANT, TSP_NO_CHAIN, TSP_INTERP_ROWS, TSP_INTERP_GLOBAL, TSP_NO_GRAPHS. */
int tsp_debug_set(tsp_handle* h, const char* key, int value);

/* Optional per-stage device timing: when enabled, tsp_project_frame records CUDA events on the
 * launching stream between its stages.  tsp_get_stage_times synchronises the device, folds the
 * recorded intervals into per-stage totals (milliseconds, number of intervals) for
 * tsp_stage_count() stages named by tsp_stage_name(), and resets the totals when reset != 0. */
int tsp_set_profiling(tsp_handle* h, int enable);
int tsp_stage_count(void);
const char* tsp_stage_name(int stage);
int tsp_get_stage_times(tsp_handle* h, double* ms_out, int64_t* count_out, int reset);

#ifdef __cplusplus
}
#endif
#endif /* TSP_B200_H */
