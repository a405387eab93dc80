/*
 * mcgaze_b200 — C ABI of the B200-native MCGaze per-clip forward.
 *
 * The reference (zgchen33/MCGaze) has no FFI today: the hot path is pure Python on top of
 * torch/cuDNN/mmcv.  This header is the boundary a maintainer would bind from the reference's
 * `MultiClueGaze.simple_test` (mmdet/models/detectors/multiclue_gaze.py:105-131); each entry
 * point cites the reference interface it replaces.  See INTEGRATION.md for the ctypes stub.
 *
 * Conventions: plain C, no torch types.  Every function returns 0 on success or a negative
 * error code; the message is available from mcg_last_error() (thread-local).  Nothing throws
 * across the ABI.  One handle per (device, stream); a handle is not re-entrant.
 */
#ifndef MCGAZE_B200_H_
#define MCGAZE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcg_engine* mcg_handle;

/* One named fp32 tensor of a reference checkpoint `state_dict` (HOST memory, contiguous).
 * Key layout: SURVEY.md section 8b, i.e. exactly what mmcv.runner.load_checkpoint feeds
 * `MultiClueGaze` (mmdet/apis/inference.py:45). Unknown keys are ignored (non-strict). */
typedef struct {
  const char* name;
  const float* data;
  int ndim;
  int64_t shape[4];
} mcg_tensor;

enum {
  MCG_PRECISION_FP16X3 = 0, /* tcgen05, split-fp16 operands (hi+lo), 3 MMAs/k-step: fp32-equivalent; parity mode */
  MCG_PRECISION_FP16 = 1,   /* tcgen05, single fp16 operands, fp32 accumulate: fast mode */
  MCG_PRECISION_SIMT = 2,   /* fp32 CUDA-core kernels only (cross-check / bring-up) */
  MCG_PRECISION_FP16C8 = 3  /* tcgen05, fp16 hi*hi + both rounding corrections as e4m3 MMAs (2 instead of 3 MMA
                               units per k-step), activations stored fp16 hi + e4m3 lo8 + e4m3 hi8: parity mode,
                               the default of the Python surface */
};

enum {
  MCG_OK = 0,
  MCG_ERR_INVALID = -1,
  MCG_ERR_CUDA = -2,
  MCG_ERR_MISSING_WEIGHT = -3,
  MCG_ERR_UNSUPPORTED = -4
};

/* Replaces build_detector(cfg.model) + load_checkpoint (mmdet/apis/inference.py:17-56):
 * folds BN into the convolutions, repacks weights to K-major split-fp16 and uploads them. */
int mcg_create(mcg_handle* out, int device, const mcg_tensor* weights, int n_weights, int precision);
int mcg_destroy(mcg_handle h);

/* Replaces MultiClueGaze.forward(return_loss=False) -> simple_test
 * (mmdet/models/detectors/multiclue_gaze.py:105-131 -> roi_heads/multiclue_gaze_roi_head.py:287-384)
 * for B clips of T frames each (the reference's test path is B == 1; B > 1 uses clip_length = T
 * exactly like forward_train, multiclue_gaze_roi_head.py:229).
 *   img          DEVICE fp32 [B*T, 3, H, W]   (NCHW, the reference layout; H, W multiples of 32)
 *   img_hw       HOST   fp32 [B*T, 2]         unpadded (h, w) = meta['img_shape'] ; NULL -> (H, W)
 *   scale_factor HOST   fp32 [B*T, 4]         meta['scale_factor'] (rescale=True) ; NULL -> no rescale
 *   out_gaze     DEVICE fp32 [B*T, 4, 3]      unit vectors: fused, face, eyes, head
 *   out_boxes    DEVICE fp32 [B*T, 3, 4]      xyxy (face, eyes, head)
 *   out_scores   DEVICE fp32 [B*T, 3]         sigmoid scores
 * Asynchronous on `stream` (a cudaStream_t). */
int mcg_forward(mcg_handle h, const float* img, int B, int T, int H, int W, const float* img_hw,
                const float* scale_factor, float* out_gaze, float* out_boxes, float* out_scores, void* stream);

/* Same, with HOST input/output buffers: copies img host->device, runs, copies results back and
 * synchronises.  This is the call tools/test_gaze360_gaze.py:107 maps to (scatter + forward +
 * .cpu()).  Pinned host memory gives asynchronous copies. */
int mcg_forward_host(mcg_handle h, const float* img_host, int B, int T, int H, int W, const float* img_hw,
                     const float* scale_factor, float* out_gaze_host, float* out_boxes_host,
                     float* out_scores_host);

/* Pipelined form of mcg_forward_host (two slots): mcg_submit_host enqueues the host->device copy on a
 * copy stream and the forward + device->host read-back on the compute stream and returns a ticket;
 * mcg_wait_host blocks until that submission's results are on the host.  Submitting batch i+1 before
 * waiting for batch i overlaps its H2D copy with batch i's forward (what a DataLoader-fed test loop,
 * mmdet/apis/test.py:107-109, gets from pinned memory + non_blocking copies).  At most two
 * submissions may be in flight; the host image buffer must stay valid until the matching wait. */
int mcg_submit_host(mcg_handle h, const float* img_host, int B, int T, int H, int W, const float* img_hw,
                    const float* scale_factor, int* ticket);
int mcg_wait_host(mcg_handle h, int ticket, float* out_gaze_host, float* out_boxes_host, float* out_scores_host);

/* Per-op parity support: copy a named intermediate of the LAST forward as dense fp32 into `dst`
 * (DEVICE).  Activations are returned NCHW like the reference's tensors.  Names: "stem", "pool",
 * "layer{1-4}.{i}", "fpn{0-3}", "stage{0-3}.roi_feat" ([R,49,256]), "stage{s}.attn",
 * "stage{s}.obj", "stage{s}.boxes", "stage{s}.delta", "stage{s}.cls".  shape_out receives up to
 * 4 dims (unused = 0).  Returns MCG_ERR_INVALID for unknown names or too small capacity. */
int mcg_get_intermediate(mcg_handle h, const char* name, float* dst, int64_t capacity, int64_t shape_out[4]);

/* Number of kernels this library launched in the last forward. */
int mcg_last_launch_count(mcg_handle h);

/* tcgen05 GEMM statistics of the last EAGER forward: out[0] = launches, out[1] = algorithmic FLOPs
 * (2*M*N*K, independent of the precision mode), out[2] = summed device time in ms measured with
 * CUDA events around each launch (needs option "time_kernels" = 1; synchronises). */
int mcg_last_umma_stats(mcg_handle h, double out[3]);

/* Per-launch device times (ms, launch order) of the tcgen05 GEMMs of the last EAGER forward (needs "time_kernels");
 * returns the number of entries written (<= capacity) or a negative error code. */
int mcg_last_umma_times(mcg_handle h, double* out_ms, int capacity);

/* Device time of EVERY kernel launch of the last EAGER forward (needs "time_kernels"), as text lines
 * "kernel-name<TAB>milliseconds\n" in launch order (tcgen05 GEMMs are named "umma:<layer>").  Returns the number of
 * bytes needed including the terminating NUL (call again with a larger buffer if > capacity) or a negative error. */
int mcg_last_kernel_profile(mcg_handle h, char* buf, int capacity);

/* Operand-range check of the fp16c8 precision mode: its e4m3 correction planes are exact only for
 * activations in [2^-6, 448] and BN-folded weights in [2^-10, 28].  Scans the trunk / FPN activations the LAST forward
 * left in the workspace and reports them with the weight statistics taken at mcg_create, as text lines
 * "name<TAB>total<TAB>nonzero<TAB>over<TAB>under<TAB>nonfinite<TAB>maxabs<TAB>energy<TAB>energy_under\n"
 * (activations: over = |v| > 448, under = 0 < |v| < 2^-8; weight rows "w:<checkpoint key>": over = |w| > 28,
 * under = 0 < |w| < 2^-10; energy = sum v^2, energy_under = its part from the `under` elements).  The reference has
 * no counterpart (it computes in fp32: mmdet/models/backbones/resnet.py:263-302); callers use it to refuse a
 * checkpoint whose statistics leave the window instead of silently degrading to fp16 accuracy.  Synchronises.
 * Returns the bytes needed including the NUL (call again with a larger buffer if > capacity) or a negative error. */
int mcg_range_report(mcg_handle h, char* buf, int capacity);

/* Capture the forward for the current shape in a CUDA graph and replay it on later calls
 * (on = 1) or launch kernels eagerly (on = 0, default). */
int mcg_set_graph_mode(mcg_handle h, int on);

/* Options: "keep_intermediates" (1: snapshot per-stage head buffers for mcg_get_intermediate),
 * "time_kernels" (1: CUDA events around every tcgen05 GEMM launch),
 * "head_tensor_cores" (0: run the head's large Linear layers on the fp32 CUDA-core kernel),
 * "fused_stem" (0: im2col -> GEMM -> max-pool chain instead of the fused stem kernel; exposes "stem"),
 * "fuse_downsample" (0: a layer's first bottleneck runs its downsample branch as its own convolution and conv3 adds
 * it as a residual, like mmdet/models/backbones/resnet.py:286-295 literally; default 1: conv3 and the downsample
 * branch are one GEMM over the concatenated K),
 * "fuse_bottleneck" (0: conv2 and conv3 + identity of layer1 / layer2 bottlenecks as separate launches),
 * "split_layers" (bit l set: the bottlenecks of layer l+1 run as two chains over the two halves of the frames, on two
 * streams - the other half's launches fill the idle last wave of a persistent launch; results are bit-identical;
 * -1: default), "split_min_frames" (smallest batch that is split; -1: default 64). */
int mcg_set_option(mcg_handle h, const char* key, int value);

/* Stand-alone convolution / GEMM with the fused epilogue, for kernel-level parity tests.
 *   engine: MCG_PRECISION_*    x: DEVICE fp32 NHWC [NB,H,W,C]    w: DEVICE fp32 [Cout, R*S*C] (r,s,c order)
 *   res: DEVICE fp32 NHWC residual or NULL; res_mode 0 none / 1 same size / 2 nearest 2x upsample
 *   out: DEVICE fp32 NHWC [NB,P,Q,Cout].  force_im2col != 0 routes 1x1/s1 through the im2col TMA path.
 *   out_mode 0: the tensor-core engines write split-fp16 planes through the smem-staged TMA-store epilogue (the
 *   trunk's format; with single fp16 the result carries one fp16 rounding) and are converted to fp32 afterwards;
 *   out_mode 1: fp32 direct-store epilogue (the head's format). */
int mcg_debug_conv(int engine, const float* x, int NB, int H, int W, int C, const float* w, int Cout, int R, int S,
                   int stride, int pad, const float* bias, const float* res, int res_mode, int relu,
                   int force_im2col, int force_block_n, int out_mode, float* out, void* stream);

/* ---- test-time image pipeline -------------------------------------------------
 * One decoded frame and the geometry the reference's test_pipeline applies to it
 * (configs/_base_/datasets/gaze360.py:27-36): the CenterCrop window (mmdet/datasets/pipelines/transforms.py:1036-1047;
 * the whole image when the config has no CenterCrop, as in multiclue_gaze_r50_l2cs.py:31-39) and the size
 * Resize(keep_ratio=True) gives it (mmcv.imrescale; transforms.py:217-228). */
typedef struct {
  const uint8_t* src;  /* DEVICE uint8 [src_h, src_w, 3], channel order as decoded by LoadImageFromFile (BGR) */
  int64_t src_stride;  /* bytes per source row (>= 3 * src_w) */
  int32_t src_h, src_w;
  int32_t crop_y, crop_x, crop_h, crop_w;
  int32_t dst_h, dst_w;
} mcg_frame;

/* Replaces Resize -> RandomFlip(0.0) -> Normalize -> Pad(size_divisor) -> DefaultFormatBundle -> collate of the
 * reference's test_pipeline (transforms.py:213-241, 739-754, 665-681; formatting.py:96; the crop of CenterCrop is the
 * window in mcg_frame) for n frames: bit-exact cv2.resize(INTER_LINEAR) of the crop window to (dst_h, dst_w),
 * BGR->RGB when to_rgb, (x - mean) / std in mmcv.imnormalize's arithmetic, zeros outside (dst_h, dst_w).
 *   frames  HOST   n descriptors (device source pointers inside)
 *   mean, std  HOST fp32 [3], output-channel order (img_norm_cfg)
 *   out     DEVICE fp32 [n, 3, Hp, Wp], 16-byte aligned, Wp % 4 == 0, every dst_h <= Hp, dst_w <= Wp;
 *           this is the `img` argument of mcg_forward
 * Asynchronous on `stream`; needs no engine handle. */
int mcg_preprocess(const mcg_frame* frames, int n, const float* mean, const float* std, int to_rgb, float* out,
                   int Hp, int Wp, void* stream);

/* ---- PNG decode on the device ("decode -> ...") -------------------------------------
 * Replaces the decode inside LoadImageFromFile (mmdet/datasets/pipelines/loading.py:58-69: mmcv.imfrombytes ->
 * cv2.imdecode(IMREAD_COLOR)) for the frames data/gaze360/<split>_rawframes/<vid>/<00000>.png holds
 * (tools/gaze360_img_reorganize.py:108,137 writes them with cv2.imwrite): 8-bit, non-interlaced PNG of colour type
 * 0 (gray), 2 (RGB), 3 (palette), 4 (gray + alpha) or 6 (RGBA).  Output = what cv2 returns for the file, bit for bit:
 * BGR uint8, gray replicated, palette expanded, alpha (and tRNS) dropped, gamma / colour-space chunks ignored.
 * The host only walks the chunk list; inflate and scanline reconstruction run on the device. */
typedef struct {
  int32_t width, height;
  int32_t bit_depth, color_type, interlace;
  int32_t channels;      /* samples per scanline pixel */
  int32_t has_palette;
  int32_t supported;     /* 1 when mcg_png_decode takes the image (bit depth 8, not interlaced) */
  int64_t idat_bytes;    /* total IDAT payload = length of the zlib stream */
  uint8_t palette[768];  /* PLTE entries (RGB), zero padded */
} mcg_png_info;

/* HOST function (no GPU needed).  Walks the chunks of one PNG file image: signature, IHDR first, PLTE, IDAT, IEND;
 * checks every chunk CRC when check_crc != 0 (libpng does).  Fills `info`; when zdata != NULL the IDAT payloads are
 * concatenated there (zcap bytes available; MCG_ERR_INVALID with info->idat_bytes set when it is too small, so a
 * first call with zdata == NULL sizes the buffer).  The H2D copy of zdata is the caller's. */
int mcg_png_parse(const uint8_t* file, int64_t nbytes, int check_crc, mcg_png_info* info, uint8_t* zdata, int64_t zcap);

/* HOST functions for a batch of FILES (what a loader thread does for LoadImageFromFile's `filename`): sizes first, so the
 * caller can lay out one pinned staging block (slot i needs sizes[i] bytes: a file's IDAT payload is shorter than the
 * file), then every file is mapped, parsed like mcg_png_parse and its payload copied from the page cache straight into
 * block + slot_off[i], on `threads` worker threads inside the call (no per-file work in the caller's language).
 *   results[i]  0 = staged (infos[i] valid), 1 = unreadable, 2 = not a PNG / malformed, 3 = valid PNG that
 *               mcg_png_decode does not take (infos[i].supported == 0: decode it on the host)
 * Returns MCG_OK when every file was staged, else MCG_ERR_INVALID (mcg_last_error names the first one). */
int mcg_png_file_sizes(const char* const* paths, int n, int64_t* sizes);
int mcg_png_stage_files(const char* const* paths, int n, int check_crc, int threads, const int64_t* slot_off,
                        const int64_t* slot_cap, uint8_t* block, mcg_png_info* infos, int32_t* results);

typedef struct {
  const uint8_t* zdata;  /* DEVICE zlib stream (the concatenated IDAT payloads) */
  int64_t zbytes;
  int32_t width, height;
  int32_t color_type;    /* 0, 2, 3, 4, 6; bit depth 8, not interlaced */
  int32_t reserved;
  const uint8_t* palette;/* DEVICE [768] RGB for colour type 3, else NULL */
  uint8_t* scan;         /* DEVICE scratch, height * (1 + width * channels) bytes: the filtered scanlines */
  uint8_t* dst;          /* DEVICE uint8 [height, width, 3] BGR (the `src` of mcg_frame) */
  int64_t dst_stride;    /* bytes per output row (>= 3 * width) */
} mcg_png_job;

/* per-image results in `status` (DEVICE int32 [n], written asynchronously): 0 = decoded */
enum {
  MCG_PNG_OK = 0, MCG_PNG_BAD_ZLIB_HEADER = 1, MCG_PNG_BAD_BLOCK_TYPE = 2, MCG_PNG_BAD_STORED_LEN = 3,
  MCG_PNG_BAD_CODE_LENGTHS = 4, MCG_PNG_BAD_SYMBOL = 5, MCG_PNG_BAD_DISTANCE = 6, MCG_PNG_OUTPUT_OVERFLOW = 7,
  MCG_PNG_INPUT_EXHAUSTED = 8, MCG_PNG_OUTPUT_SHORT = 9, MCG_PNG_BAD_FILTER = 10, MCG_PNG_BAD_JOB = 11,
  MCG_PNG_BAD_CHECKSUM = 12
};

/* n images, one warp each, ONE launch per stage whatever n is: inflate (RFC 1950 / 1951: stored, fixed and dynamic
 * blocks) into `scan`, then scanline reconstruction (None / Sub / Up / Average / Paeth) + colour conversion into `dst`.
 *   jobs    DEVICE array [n] (upload it with the compressed bytes; the kernels read the descriptors from HBM, so a batch
 *           of thousands of images is one launch and fills the GPU: a single stream is a serial decode, the throughput
 *           is in the number of images in flight)
 *   status  DEVICE int32 [n]: MCG_PNG_OK, a stream error, or MCG_PNG_BAD_JOB for a descriptor with a null buffer, a
 *           colour type outside {0, 2, 3, 4, 6}, a non-positive size or dst_stride < 3 * width
 * Asynchronous on `stream`; needs no engine handle.  A corrupt stream sets its status and leaves `dst` undefined, it
 * never writes outside `scan` / `dst`.  The zlib stream's Adler-32 trailer is verified on the device against the inflated
 * bytes (MCG_PNG_BAD_CHECKSUM), so a caller may skip the host-side chunk CRCs (check_crc = 0) and still catch corrupt
 * pixel data. */
int mcg_png_decode(const mcg_png_job* jobs, int n, int32_t* status, void* stream);

/* ---- overlap merge------------------------------------------------------------------
 * Replaces the clip-to-video merge of tools/test_gaze360_gaze.py:129-201 for per-clip results that are on the device:
 * windows of clip_len frames every `stride` frames, the last one right-aligned (:73-86); frames a clip adds are copied
 * with their boxes zeroed where that clip's score < 0.5 (:135-141), frames it shares with earlier clips are averaged
 * with the running values - coordinates zeroed when either score < 0.5 (:170-174), scores and gaze vectors averaged
 * without re-normalisation (:182-183).
 *   rows         DEVICE fp32 [n_clips, clip_len, 27]  per clip and frame: boxes [3,4] xyxy, scores [3], gaze [4,3]
 *                (fused, face, eyes, head); clips of a video are consecutive, short clips are padded
 *   clip_start   DEVICE int32 [n_videos + 1]   first clip of each video
 *   frame_start  DEVICE int32 [n_videos + 1]   first frame of each video in the outputs
 *   det          DEVICE fp32 [F, 3, 5]  (x1, y1, x2, y2, score)      gaze  DEVICE fp32 [F, 4, 3]
 * Asynchronous on `stream`; needs no engine handle. */
int mcg_merge_clips(const float* rows, const int32_t* clip_start, const int32_t* frame_start, int n_videos, int clip_len,
                    int stride, float* det, float* gaze, void* stream);

/* ---- scorer ---------------------------------------------------------------------
 * Replaces gaze_error of tools/calculate_mae_gaze360.py (:110-188: smooth_filter :16-29 with alpha 0.6,
 * compute_angular_error :77-94, compute_yaw_angular :69-74) and of tools/calculate_mae_l2cs.py (:96-182) for
 * per-frame gaze vectors that are on the device.
 *   pred, gt     DEVICE fp32 [F, 3]   predicted / ground-truth vectors, the frames of all videos concatenated
 *   video_start  DEVICE int32 [n_videos + 1]   first frame of each video, video_start[n_videos] = F
 *   variant      MCG_SCORER_GAZE360: front-20 = |yaw(gt)| <= 20 deg; MCG_SCORER_L2CS: and |pitch(gt)| <= 20 deg
 *                (calculate_mae_l2cs.py:139; its ground truth sits at annotations[3 * video], :110 - the caller's job)
 *   out          DEVICE double [6]    {sum, frames} for 360, front-180 (|yaw(gt)| <= 90 deg), front-20;
 *                                     sum = per-video mean angle in degrees x frames of that video; MAE = sum / frames
 *                                     (sums of several calls / ranks add up: one all-reduce gives the sharded MAE)
 * Asynchronous on `stream`; needs no engine handle. */
enum { MCG_SCORER_GAZE360 = 0, MCG_SCORER_L2CS = 1 };
int mcg_gaze_error(const float* pred, const float* gt, const int32_t* video_start, int n_videos, int variant, double* out,
                   void* stream);

const char* mcg_last_error(void);
const char* mcg_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MCGAZE_B200_H_ */
