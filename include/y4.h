/*
 * y4.h — C-ABI of the B200-native YOLOv4 inference hot path (liby4.so).
 *
 * The reference (taipingeric/yolo-v4-tf.keras) has no FFI: its boundary is the Keras-model call surface
 * used by `class Yolov4`.  Every entry point below names the reference call site it replaces
 * (paths relative to the reference repo).  Plain pointers and sizes only; no torch / numpy types.
 *
 * Conventions: every function returns 0 on success, <0 on error (never throws across the ABI);
 * y4_last_error() gives the message.  The caller owns all host buffers; the engine owns device memory
 * and frees it in y4_destroy().  One engine is bound to one CUDA device and one stream and is NOT
 * thread-safe; independent engines (one per GPU) may be used concurrently.  There is no CPU fallback:
 * every compute entry point fails with Y4_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef Y4_H_
#define Y4_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Y4_OK                 0
#define Y4_ERR_ARG          (-1)   /* bad argument / config                                      */
#define Y4_ERR_CUDA         (-2)   /* CUDA runtime / driver failure, or no usable sm_100 device   */
#define Y4_ERR_WEIGHTS      (-3)   /* darknet file has the wrong byte count (utils.py:50-53)      */
#define Y4_ERR_STATE        (-4)   /* e.g. predict before weights were loaded                     */
#define Y4_ERR_CAPACITY     (-5)   /* reserved (round 1: candidate-list overflow; no longer produced: images with more than
                                      Y4_MAX_CANDIDATES candidates take an exact slower path, as TF has no limit)  */
#define Y4_ERR_COMM         (-6)   /* NCCL failure                                                */

#define Y4_MAX_CANDIDATES   8192   /* per image, after the score filter: capacity of the FAST NMS path         */

/* conv-stack arithmetic.  Decode + NMS are always fp32. */
#define Y4_PREC_FP32        0      /* fp32 activations, CUDA-core FFMA implicit GEMM (parity mode)          */
#define Y4_PREC_FP16        1      /* fp16 activations/weights, fp32 accumulate in TMEM, tcgen05 + TMA       */
#define Y4_PREC_FP16_SIMT   2      /* fp16 storage, CUDA-core kernels (debug reference for the tcgen05 path) */
#define Y4_PREC_FP16X3      3      /* tensor-core parity mode: activations and weights as fp16 hi+lo pairs, three
                                      tcgen05 MMAs per k-step (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi), fp32 accumulate  */

typedef struct y4_engine y4_engine;

/* Mirrors config.py:1-17 (`yolo_config`) + the ctor args of models.py:18-38. */
typedef struct y4_config {
    int32_t img_size;         /* square, multiple of 32 (asserts models.py:23-24); grid = img_size/stride     */
    int32_t num_classes;      /* len(class_names), models.py:27                                                */
    int32_t max_batch;        /* device buffers are sized for this many images                                 */
    int32_t precision;        /* Y4_PREC_*                                                                     */
    int32_t device;           /* CUDA ordinal                                                                  */
    int32_t max_boxes;        /* config.py 'max_boxes' = max_output_size_per_class = max_total_size            */
    int32_t strides[3];       /* config.py 'strides'                                                           */
    int32_t reserved_;
    float   anchors[18];      /* config.py 'anchors', reshape(3,3,2) (models.py:29)                            */
    double  xyscale[3];       /* config.py 'xyscale' — double: the reference evaluates 0.5*(xyscale-1) in
                                 Python double before the float32 cast (custom_layers.py:251)                  */
    float   iou_threshold;    /* config.py 'iou_threshold'                                                     */
    float   score_threshold;  /* config.py 'score_threshold'                                                   */
} y4_config;

/* Per-conv description, for tests and tooling. */
typedef struct y4_layer_info {
    int32_t idx, cin, cout, ksize, stride, batch_norm, activation /*0 linear,1 leaky,2 mish*/;
    int32_t out_hw;           /* output spatial size at the engine's img_size */
    int32_t kernel_kind;      /* 0 = CUDA-core implicit GEMM, 1 = tcgen05 flat GEMM, 2 = tcgen05 strided-box, 3 / 4 = conv 0 direct /
                                 tcgen05; steps only: 5 = SPP max-pools, 6 = conv 0 + conv 1 fused (stem_tc.cuh, out_name "c0+c1") */
    int32_t tile_n;           /* N tile of the tcgen05 kernel (0 if kernel_kind==0) */
    int64_t flops;            /* 2*MAC per image */
    char    out_name[16];     /* name of the tensor this conv materialises (r<k> when the residual add is fused) */
    int32_t tc_mode;          /* tcgen05 plan: 1 flat (one TMA per tap), 2 strided box, 3 flat with A-patch reuse, 4 CTA pair (cta_group::2), 5 CTA pair with A-patch reuse */
    int32_t tc_epilogue;      /* 0 per-thread global stores; 32 / 64: swizzled smem slab + TMA store in groups of that many channels */
    int32_t tc_stages, tc_group, tc_ctas_per_sm, tc_bk;   /* ring depth, k-blocks per barrier, persistent CTAs per SM, K block */
    int32_t tc_epi_warps;     /* 4 or 8 epilogue warps; 44 = four warps, lean variant compiled for up to four CTAs per SM */
    int32_t tc_resident_w;    /* 1: the weight matrix stays in shared memory for the life of the CTA */
} y4_layer_info;

/* Fills *cfg with the reference defaults (config.py) at 416x416, 80 classes, max_batch 1, fp16. */
int  y4_default_config(y4_config* cfg);

/* ctor / dtor — replaces Yolov4.__init__/build_model graph construction (models.py:18-73). */
int  y4_create(y4_engine** out, const y4_config* cfg);
void y4_destroy(y4_engine* e);

/* Last error message of this engine (or of the last failed y4_create when e == NULL). */
const char* y4_last_error(const y4_engine* e);

/* Darknet .weights loader — replaces load_weights(self.yolo_model, path) (models.py:77, utils.py:12-53).
 * Sequential, Keras creation order; verifies the exact byte count (utils.py:50-53) -> Y4_ERR_WEIGHTS. */
int  y4_load_darknet(y4_engine* e, const char* path);
int  y4_load_darknet_from_memory(y4_engine* e, const void* data, size_t nbytes);

/* inference_model.predict(imgs) — models.py:113,159.
 * imgs: host (batch, S, S, 3) float32, RGB in [0,1].  Outputs (caller-allocated host buffers):
 * boxes (batch,max_boxes,4) x1,y1,x2,y2 normalised & clipped to [0,1]; scores (batch,max_boxes) descending;
 * classes (batch,max_boxes) as float32; valid (batch) int32; zero padded.
 * cand_idx (nullable): (batch,max_boxes) int32 flat candidate index n = off_scale + (row*g+col)*3 + a, -1 padded. */
int  y4_predict(y4_engine* e, const float* imgs, int32_t batch,
                float* boxes, float* scores, float* classes, int32_t* valid, int32_t* cand_idx);

/* yolo_model.predict(imgs) — models.py:514,521: the three raw head tensors (batch,g,g,3*(5+nc)) float32. */
int  y4_forward_heads(y4_engine* e, const float* imgs, int32_t batch,
                      float* head_s, float* head_m, float* head_l);

/* yolov4_head(...) + nms(..., iou_threshold, score_threshold) — models.py:522-523,
 * custom_layers.py:201-298: decode + score filter + per-class NMS from caller-supplied head tensors
 * with runtime thresholds.  Same outputs as y4_predict. */
int  y4_decode_nms(y4_engine* e, const float* head_s, const float* head_m, const float* head_l,
                   int32_t batch, float iou_threshold, float score_threshold,
                   float* boxes, float* scores, float* classes, int32_t* valid, int32_t* cand_idx);

/* ---- pipelined host path (batched callers such as export_prediction, models.py:141-179) ------------------------- */
/* y4_submit enqueues H2D of `imgs` (host, ideally pinned: y4_host_alloc) on a copy stream, then forward + decode + NMS and
 * the D2H of the results into pinned staging, and returns immediately; y4_collect blocks until the OLDEST outstanding
 * submit has finished and copies its results out.  Up to two submits may be outstanding, so the H2D of batch i+1
 * overlaps the compute of batch i.  Same outputs as y4_predict. */
int  y4_submit(y4_engine* e, const float* imgs, int32_t batch);
int  y4_collect(y4_engine* e, int32_t batch, float* boxes, float* scores, float* classes, int32_t* valid, int32_t* cand_idx);

/* ---- device-resident path (benchmarks, data-parallel serving) ------------------------------------- */
/* Fill the engine's device input with synthetic images: pixel value = hash(seed, global element index),
 * image i depends only on (seed, first_index + i) so results do not depend on how images are sharded. */
int  y4_synth_fill(y4_engine* e, uint64_t seed, int64_t first_index, int32_t batch);
/* Enqueue forward + decode + NMS on the resident input (async on the engine stream). */
int  y4_run_resident(y4_engine* e, int32_t batch);
/* Enqueue only the 110-conv forward (heads stay on device) / only decode+NMS on the resident heads. */
int  y4_run_forward_resident(y4_engine* e, int32_t batch);
int  y4_run_decode_nms_resident(y4_engine* e, int32_t batch);
/* Upload head tensors (host, packed (batch,g,g,3*(5+nc))) into the resident head buffers. */
int  y4_upload_heads(y4_engine* e, const float* head_s, const float* head_m, const float* head_l, int32_t batch);
/* Copy the last results to host (synchronises the stream). */
int  y4_fetch_results(y4_engine* e, int32_t batch,
                      float* boxes, float* scores, float* classes, int32_t* valid, int32_t* cand_idx);
int  y4_sync(y4_engine* e);
/* CUDA-event timer on the engine stream: begin/end bracket enqueued work; end synchronises and returns ms. */
int  y4_timer_begin(y4_engine* e);
int  y4_timer_end(y4_engine* e, float* ms);
/* Write > L2 bytes to evict the 126 MB L2 between timed iterations. */
int  y4_flush_l2(y4_engine* e);
/* Number of kernels this engine has launched since creation. */
int64_t y4_launch_count(const y4_engine* e);
/* Per-step CUDA-event timing of one forward: ms[y4_num_steps()], in schedule order (y4_describe_step). */
int  y4_profile_layers(y4_engine* e, int32_t batch, float* ms, int32_t n);

/* Pinned host memory for callers that want async H2D/D2H. */
/* ---- raw 8-bit images: preprocess_img + predict (models.py:95-98, 109-127, 141-179) --------------------
 * imgs[i]: host uint8 HWC (heights[i], widths[i], 3), densely packed (what cv2.imread returns).  Each image is resized on
 * the GPU to (S, S) with OpenCV's 8-bit INTER_LINEAR fixed-point arithmetic (bit-exact with cv2.resize), divided by 255 in
 * float64 and rounded to float32 (Keras' input cast); aspect ratio is not preserved (models.py:96).
 * reverse_channels = 1 reproduces predict()'s BGR->RGB flip (models.py:126); export_prediction / predict_raw pass 0.
 * H2D traffic is 1 byte per source pixel-channel instead of 4 bytes per resized one. */
int  y4_predict_u8(y4_engine* e, const uint8_t* const* imgs, const int32_t* heights, const int32_t* widths, int32_t batch,
                   int32_t reverse_channels, float* boxes, float* scores, float* classes, int32_t* valid, int32_t* cand_idx);
/* Only the preprocessing: out (nullable) receives the (batch, S, S, 3) float32 network input (parity tests). */
int  y4_preprocess_u8(y4_engine* e, const uint8_t* const* imgs, const int32_t* heights, const int32_t* widths, int32_t batch,
                      int32_t reverse_channels, float* out);
/* Pipelined form of y4_predict_u8: pair with y4_collect, same rules as y4_submit. */
int  y4_submit_u8(y4_engine* e, const uint8_t* const* imgs, const int32_t* heights, const int32_t* widths, int32_t batch,
                  int32_t reverse_channels);

void* y4_host_alloc(size_t nbytes);
void  y4_host_free(void* p);

/* ---- introspection --------------------------------------------------------------------------------- */
int  y4_num_layers(const y4_engine* e);
int  y4_describe_layer(const y4_engine* e, int32_t idx, y4_layer_info* info);
/* The launch schedule of one forward (what y4_profile_layers times, in order): one entry per kernel launch.  A step is one of
 * the 110 convs, a fused pair of sibling convs (csp_block's route / main 1x1, custom_layers.py:59-60: idx >= 110, out_name
 * "c2+c3", flops of both), conv 0 + conv 1 in one kernel (kernel_kind 6, idx -2) or the SPP max-pools (kernel_kind 5, idx -1). */
int  y4_num_steps(const y4_engine* e);
int  y4_describe_step(const y4_engine* e, int32_t step, y4_layer_info* info);
int64_t y4_num_boxes(const y4_engine* e);   /* N = 3 * sum(g^2) */
/* Copy a named intermediate (e.g. "c0", "r1", "cat6", "c93") to host as unpadded NHWC float32
 * (batch,H,W,C); returns element count written, <0 on error.  out may be NULL to query the count. */
int64_t y4_debug_get_tensor(y4_engine* e, const char* name, int32_t batch, float* out, int64_t capacity);

/* Re-run one conv on the buffers as they are, with the CUDA-core kernel (use_tc = 0) or its tcgen05 plan (1):
 * lets tests compare both kernels on identical inputs.  Synchronises. */
int  y4_debug_run_conv(y4_engine* e, int32_t idx, int32_t batch, int32_t use_tc);

/* clock64 phase timestamps of one tcgen05 conv launch, 16 int64 per CTA (see y4_engine.cu); tooling only. */
int  y4_debug_trace_conv(y4_engine* e, int32_t idx, int32_t batch, int64_t* out, int32_t max_ctas);

/* ---- multi-GPU (one process per GPU; images shard, weights replicate) ------------------------------ */
/* uid: 128-byte ncclUniqueId produced on rank 0 and broadcast by the host launcher. */
int  y4_comm_unique_id(void* uid128);
int  y4_comm_init(y4_engine* e, int32_t rank, int32_t nranks, const void* uid128);
/* ncclAllGather of the per-image result records of the last run (2,404 B/img at max_boxes=100, +cand_idx):
 * outputs are (nranks*batch, ...) host buffers, rank-major. */
int  y4_allgather_results(y4_engine* e, int32_t batch,
                          float* boxes, float* scores, float* classes, int32_t* valid, int32_t* cand_idx);

#ifdef __cplusplus
}
#endif
#endif /* Y4_H_ */
