/*
 * svb200.h -- C ABI of libsvb200.so, the B200-native VideoMixer compute path.
 *
 * The reference (unpause-live/SwiftVideo @113d3d9) has no bespoke C ABI on this path: its FFI is the CUDA
 * driver API re-exported by Sources/CCUDA/shim.h:1-4, driven from Swift in compute.cuda.swift.  A
 * replacement therefore plugs in at two levels, and this header declares both:
 *
 *  (1) DEVICE MODULE.  svb_kernel_module_image() returns an sm_100a cubin for cuModuleLoadData whose entry
 *      points carry the ComputeKernel case names and the parameter convention of
 *      compute.cuda.swift:294-303 -- the reference's own runComputeKernel can launch them unchanged
 *      ("level 1").
 *  (2) HOST OPERATORS.  One C function per Swift free function / method of the hot path, same names (snake
 *      case), argument meaning and error behaviour, so that a thin Swift module (`CSVB200`, a modulemap over
 *      this header like CCUDA's) lets compute.swift / mix.video.swift forward to it ("level 2").
 *      Citations below are file:line under /root/reference/Sources/SwiftVideo/.
 *
 * Handles are opaque; every function returns svb_status (0 = ok) unless noted and records a message for
 * svb_last_error().  No CPU fallback exists: without a B200 and its driver every compute entry point fails
 * with SVB_ERROR_DEVICE_NOT_AVAILABLE.
 */
#ifndef SVB200_H
#define SVB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ComputeError (compute.swift:22-39), in declaration order */
typedef enum svb_status {
    SVB_OK = 0,
    SVB_ERROR_INVALID_PLATFORM = 1,
    SVB_ERROR_INVALID_DEVICE = 2,
    SVB_ERROR_INVALID_OPERATION = 3,
    SVB_ERROR_INVALID_VALUE = 4,
    SVB_ERROR_INVALID_PROGRAM = 5,
    SVB_ERROR_INVALID_CONTEXT = 6,
    SVB_ERROR_DEVICE_NOT_AVAILABLE = 7,
    SVB_ERROR_OUT_OF_MEMORY = 8,
    SVB_ERROR_COMPILER_NOT_AVAILABLE = 9,
    SVB_ERROR_COMPUTE_KERNEL_NOT_FOUND = 10,
    SVB_ERROR_BAD_TARGET = 11,
    SVB_ERROR_BAD_INPUT_DATA = 12,
    SVB_ERROR_BAD_CONTEXT_STATE = 13,
    SVB_ERROR_COMPILER_ERROR = 14,
    SVB_ERROR_UNKNOWN = 15,
    SVB_ERROR_NOT_IMPLEMENTED = 16
} svb_status;

/* PixelFormat (sample.pict.swift:20-33), in declaration order */
typedef enum svb_pixel_format {
    SVB_PIXEL_NV12 = 0, SVB_PIXEL_NV21, SVB_PIXEL_YUVS, SVB_PIXEL_ZVUY, SVB_PIXEL_Y420P, SVB_PIXEL_Y422P, SVB_PIXEL_Y444P,
    SVB_PIXEL_RGBA, SVB_PIXEL_BGRA, SVB_PIXEL_SHAPE, SVB_PIXEL_TEXT, SVB_PIXEL_INVALID,
    /* ours (upstream: "TODO: Higher bit-depth formats", sample.pict.swift:19): NV12's plane shapes, 16-bit little-endian
     * words with ten bits in the MSBs.  An input of svb_scale_convert_picture only. */
    SVB_PIXEL_P010
} svb_pixel_format;

/* BufferType (sample.pict.swift:58-63) */
typedef enum svb_buffer_type { SVB_BUFFER_SHARED = 0, SVB_BUFFER_CPU, SVB_BUFFER_GPU, SVB_BUFFER_INVALID } svb_buffer_type;

/* ComputeDeviceType (compute.swift:41-46) */
typedef enum svb_device_type { SVB_DEVICE_GPU = 0, SVB_DEVICE_CPU, SVB_DEVICE_ACCELERATOR, SVB_DEVICE_DEFAULT } svb_device_type;

/* ComputeKernel (compute.swift:49-74), in declaration order; SVB_KERNEL_CUSTOM takes a name */
typedef enum svb_compute_kernel {
    SVB_KERNEL_IMG_NV12_NV12 = 0, SVB_KERNEL_IMG_BGRA_NV12, SVB_KERNEL_IMG_RGBA_NV12, SVB_KERNEL_IMG_BGRA_BGRA,
    SVB_KERNEL_IMG_Y420P_Y420P, SVB_KERNEL_IMG_Y420P_NV12, SVB_KERNEL_IMG_CLEAR_NV12, SVB_KERNEL_IMG_CLEAR_YUVS,
    SVB_KERNEL_IMG_CLEAR_BGRA, SVB_KERNEL_IMG_CLEAR_Y420P, SVB_KERNEL_IMG_CLEAR_RGBA, SVB_KERNEL_IMG_RGBA_Y420P,
    SVB_KERNEL_IMG_BGRA_Y420P, SVB_KERNEL_SND_S16I_S16I, SVB_KERNEL_ME_FULLSEARCH, SVB_KERNEL_CUSTOM,
    /* ours: sources the reference's PixelFormat names but has no kernel for, named by findKernel's own rule */
    SVB_KERNEL_IMG_NV21_NV12, SVB_KERNEL_IMG_Y422P_NV12, SVB_KERNEL_IMG_Y444P_NV12, SVB_KERNEL_IMG_Y422P_Y420P, SVB_KERNEL_IMG_Y444P_Y420P
} svb_compute_kernel;

/* VideoMixer compose strategy (ours; the reference only has the per-layer sequence) */
/* SVB_MIX_FUSED_GATHER: the fused compositor fetching its taps through the texture unit (svb_mix_gather) wherever every
 * staged layer's planes can be bound as textures (base and pitch alignment); same bytes, see DESIGN.md section 5 */
typedef enum svb_mix_mode { SVB_MIX_FUSED = 0, SVB_MIX_PER_LAYER = 1, SVB_MIX_GENERIC = 2, SVB_MIX_FUSED_GATHER = 3, SVB_MIX_FUSED_TILED = 4, SVB_MIX_FUSED_RING = 5 } svb_mix_mode;

typedef struct svb_context svb_context; /* ComputeContext  (compute.cuda.swift:60-73) */
typedef struct svb_picture svb_picture; /* PictureSample   (sample.pict.linux.swift:105-249), immutable */
typedef struct svb_mixer svb_mixer;     /* VideoMixer      (mix.video.swift:21) */
typedef struct svb_timer svb_timer;     /* a pair of CUDA events on the context's streams */

/* ImageUniforms (compute.swift:76-86): 236 bytes, what applyComputeImage uploads */
typedef struct svb_image_uniforms {
    float transform[16], texture_transform[16], border_matrix[16];
    float fill_color[4];
    float input_size[2], output_size[2];
    float opacity, image_time, target_time;
} svb_image_uniforms;

/* Plane (sample.pict.swift:46-56) + where its bytes are */
typedef struct svb_plane_info {
    float width, height;
    int32_t stride, bit_depth, components;
    void* host;                 /* NULL when the sample has no CPU buffer */
    unsigned long long device;  /* CUdeviceptr, 0 when the sample has no GPU buffer */
    size_t size;                /* stride * height */
} svb_plane_info;

typedef struct svb_picture_info {
    int32_t pixel_format, buffer_type;
    float width, height;
    int32_t plane_count;
    svb_plane_info planes[3];
    float matrix[16], texture_matrix[16], border_matrix[16], fill_color[4], opacity;
    int32_t z_index;            /* zIndex(), sample.pict.linux.swift:116 */
    int64_t pts, time, timescale;
} svb_picture_info;

/* ---- errors ------------------------------------------------------------------------------------------ */
const char* svb_last_error(void);   /* message of the calling thread's last failure ("" if none) */
const char* svb_version(void);

/* ---- devices and contexts ---------------------------------------------------------------------------- */
int svb_available_compute_devices(void);                      /* availableComputeDevices().count, compute.cuda.swift:132-153 */
int svb_has_available_compute_devices(int device_type);       /* hasAvailableComputeDevices(forType:), compute.swift:112-119 */
/* makeComputeContext(forType:) compute.swift:121-129 (upstream always takes devices.first = index 0) */
svb_status svb_make_compute_context(int device_type, int device_index, svb_context** out);
svb_status svb_create_compute_context_sharing(const svb_context* sharing, svb_context** out); /* compute.cuda.swift:155-157 */
svb_status svb_destroy_compute_context(svb_context* ctx);      /* compute.cuda.swift:167-169 (+ frees the handle) */
svb_status svb_begin_compute_pass(svb_context* ctx);           /* compute.cuda.swift:308-311 */
svb_status svb_end_compute_pass(svb_context* ctx, int wait_for_completion); /* compute.cuda.swift:313-319 */
int svb_context_device_index(const svb_context* ctx);
int svb_context_sm_count(const svb_context* ctx);

/* ---- kernels ----------------------------------------------------------------------------------------- */
/* The sm_100a module with every kernel: feed it to cuModuleLoadData (compute.cuda.swift:193). */
svb_status svb_kernel_module_image(const void** image, size_t* size);
svb_status svb_default_compute_kernel_from_string(const char* name, int* kernel); /* compute.swift:90-110 */
const char* svb_compute_kernel_name(int kernel);               /* String(describing: kernel) */
/* buildComputeKernel compute.cuda.swift:171-201; image NULL = the built-in module, else a cubin/PTX image */
svb_status svb_build_compute_kernel(svb_context* ctx, const char* name, const void* image);
/* buildComputeKernel(_:name:source:) as upstream has it (compute.cuda.swift:171-201): CUDA C source -> NVRTC (--gpu-architecture=sm_100a where
 * upstream says compute_30, --fmad=false as upstream, :177) -> module -> function `name`, registered in the context's library under `name`:
 * svb_run_compute_kernel(..., SVB_KERNEL_CUSTOM, name, ...) launches it with the reference's argument convention ([out planes..., in planes...,
 * uniforms, inStride], block gcd(W,16) x gcd(H,16)); registered under a built-in's name it replaces that built-in for this context (:210-212).
 * SVB_ERROR_COMPILER_NOT_AVAILABLE without libnvrtc, SVB_ERROR_COMPILER_ERROR (log in svb_last_error) when the source does not compile. */
svb_status svb_build_compute_kernel_from_source(svb_context* ctx, const char* name, const char* source);
/* runComputeKernel<T> compute.cuda.swift:260-306 */
svb_status svb_run_compute_kernel(svb_context* ctx, const svb_picture* const* images, int image_count, const svb_picture* target,
                                  int kernel, const char* custom_name, int max_planes, const void* uniforms, size_t uniforms_size,
                                  int blends);
/* applyComputeImage compute.swift:145-170 */
svb_status svb_apply_compute_image(svb_context* ctx, const svb_picture* image, const svb_picture* target, int kernel);
/* the uniforms applyComputeImage would upload (compute.swift:149-161) */
svb_status svb_make_image_uniforms(const svb_picture* image, const svb_picture* target, svb_image_uniforms* out);

/* ---- pictures ---------------------------------------------------------------------------------------- */
/* createPictureSample sample.pict.linux.swift:254-273; pinned_from != NULL makes the CPU buffer page-locked */
svb_status svb_create_picture_sample(float width, float height, int pixel_format, const char* asset_id, const char* workspace_id,
                                     svb_context* pinned_from, svb_picture** out);
/* A CPU sample over caller-described planes with their own strides -- ImageBuffer(pixelFormat:bufferType:size:buffers:planes:)
 * + PictureSample(img, ...) sample.pict.linux.swift:23-39,160-189, as the FFmpeg decoder builds them with linesize strides
 * (SwiftVideo_FFmpeg/dec.video.ffmpeg.swift:144-220).  The bytes are copied. */
svb_status svb_picture_sample_from_planes(float width, float height, int pixel_format, const void* const* planes, const int32_t* strides,
                                          int plane_count, const char* asset_id, const char* workspace_id, svb_context* pinned_from,
                                          svb_picture** out);
/* PictureSample(other, matrix:textureMatrix:borderMatrix:fillColor:opacity:revision:assetId:) :194-226; NULL keeps other's value.
 * Matrices are 16 floats in VectorMath memory order (m11,m12,m13,m14,m21,...). */
svb_status svb_picture_with(const svb_picture* other, const float* matrix, const float* texture_matrix, const float* border_matrix,
                            const float* fill_color, const float* opacity, const char* revision, const char* asset_id,
                            svb_picture** out);
svb_status svb_picture_info_get(const svb_picture* pict, svb_picture_info* out);
/* revision() / assetId() sample.pict.linux.swift:130,142: valid while the handle lives */
const char* svb_picture_revision(const svb_picture* pict);
const char* svb_picture_asset_id(const svb_picture* pict);
unsigned long long svb_picture_identity(const svb_picture* pict); /* equal for two handles of the same (immutable) sample; 0 for NULL */
svb_status svb_picture_wait(const svb_picture* pict);          /* block until an asynchronously produced sample is complete */
void svb_picture_release(svb_picture* pict);
/* uploadComputePicture / downloadComputePicture compute.cuda.swift:359-402.  wait=1 is upstream's behaviour (synchronous copies,
 * endComputePass(ctx, true)).  wait=0 returns at once with the copies queued:
 *   upload:   the SOURCE sample's host bytes -- above all a page-locked staging buffer that a decoder refills -- must stay untouched
 *             until svb_picture_wait(result) returns (consumers on the GPU side are ordered behind the copy by the library itself);
 *   download: call svb_picture_wait(result) before touching the result's host bytes.  The result always owns fresh host buffers: a
 *             download never changes the bytes of the sample that was uploaded, as upstream's value-type Data guarantees.
 * A mixer may recycle the GPU sample that is being downloaded (its backing ring comes round every ten ticks): the next compose into
 * those planes waits for the copy on the device, so asynchronous consumers need no frame-period bookkeeping. */
svb_status svb_upload_compute_picture(svb_context* ctx, const svb_picture* pict, int max_planes, int retain_cpu_buffer, int wait, svb_picture** out);
/* the same for several pictures at once (a tick's layers).  CPU pictures that lie next to each other in page-locked host memory --
 * svb_create_picture_sample(pinned_from) hands out neighbours when called in sequence -- share one device block and travel as ONE copy
 * (a link busy in both directions gives 64 separate pictures 86 % of what it gives one copy); others go up one by one; GPU pictures pass
 * through.  svb_video_mixer_tick_many uploads its CPU layers this way. */
svb_status svb_upload_compute_pictures(svb_context* ctx, const svb_picture* const* picts, int count, int max_planes, int retain_cpu_buffer, int wait,
                                       svb_picture** outs);
svb_status svb_download_compute_picture(svb_context* ctx, const svb_picture* pict, int retain_gpu_buffer, int wait, svb_picture** out);
/* GPUBarrierUpload / GPUBarrierDownload compute.swift:175-198, :232-255: the pipeline stages around the two calls above.  A sample that
 * already lives on the right side passes through (*out is another handle of the SAME sample); a failure returns the status and fills
 * *err (may be NULL) with the event error upstream emits: EventError("barrier.upload" | "barrier.download", -1, "<error>", assetId:). */
typedef struct svb_event_error {
    char domain[32];
    int code;
    char description[256];
    char asset_id[128];
} svb_event_error;
svb_status svb_gpu_barrier_upload(svb_context* ctx, const svb_picture* pict, int retain_cpu_buffer, int wait, svb_picture** out, svb_event_error* err);
svb_status svb_gpu_barrier_download(svb_context* ctx, const svb_picture* pict, int retain_gpu_buffer, int wait, svb_picture** out, svb_event_error* err);

/* ---- convert + scale (an extension: no upstream counterpart) ------------------------------------------------ */
/* The reference has no resize filter beyond the OpenCL linear sampler (kernels.cl.swift:61), no high-bit-depth format
 * (sample.pict.swift:19) and, on Linux, no kernel that writes BGRA (compute.swift:54 reserves img_bgra_bgra; only a
 * half-written Metal body exists, kernels.metal:51-62).  BASELINE.json's configs 2 and 5 name this operator with libswscale
 * as comparator: src = a GPU NV12 or P010 sample, result = a new GPU BGRA sample of dst_width x dst_height, separable
 * bilinear or Lanczos-3 resize, BT.601 limited-range YUV -> full-range RGB.  Defined bit for bit by oracle/scale_oracle.c;
 * within one code value of libswscale 9.1 in its accurate mode.  wait=0 returns at once (svb_picture_wait before use). */
typedef enum svb_scale_filter { SVB_FILTER_BILINEAR = 0, SVB_FILTER_LANCZOS3 = 1 } svb_scale_filter;
svb_status svb_scale_convert_picture(svb_context* ctx, const svb_picture* src, float dst_width, float dst_height, int dst_pixel_format, int filter,
                                     int wait, svb_picture** out);
/* One axis of that resize: *taps weights per output sample (weights[dst_n * taps]) from source index first[x] on; call
 * with first = weights = NULL to learn *taps.  No device needed. */
svb_status svb_scale_filter_table(int filter, int src_n, int dst_n, int32_t* first, float* weights, int weights_capacity, int* taps);

/* ---- device hand-off: a composited frame consumed where it lies -------------------------------------------
 * Upstream downloads every mixed frame to feed libavcodec (its h264_nvenc line is commented out, enc.video.ffmpeg.swift:169-170) and relies
 * on "done within ten ticks" for the backing ring (mix.video.swift:152-164).  These three calls let an on-device consumer (an NVENC
 * session on the same context, or a peer GPU that gathers several mixers' frames) take the planes directly and hand them back explicitly. */
typedef struct svb_device_frame {
    int32_t device_index;            /* CUDA ordinal the planes live on */
    int32_t pixel_format;            /* svb_pixel_format */
    int32_t plane_count;
    float width, height;
    struct { unsigned long long ptr; int32_t pitch; int32_t width_bytes; int32_t rows; int32_t pad_; } planes[3];
    void* context;                   /* CUcontext (the device's primary context) the pointers belong to */
    void* ready_event;               /* CUevent: cuStreamWaitEvent(consumer_stream, ready_event, 0) before reading; NULL = nothing pending.
                                        Owned by the picture handle: valid until svb_picture_release */
} svb_device_frame;
svb_status svb_picture_device_frame(const svb_picture* pict, svb_device_frame* out);
/* the consumer has queued its reads on `consumer_stream` (a CUstream of frame.context): the producer's next write of these planes
 * (backing-ring reuse, pool reuse after release) is ordered behind that point */
svb_status svb_picture_consumed_on(const svb_picture* pict, void* consumer_stream);
/* copy a GPU sample that lives on another device into dst_ctx (cuMemcpyPeerAsync, direct over NVLink when the devices are peers),
 * ordered behind the sample's completion on its own device; a sample already on dst_ctx's device comes back as another handle of itself.
 * wait = 0 returns at once (svb_picture_wait / the result's ready_event to join). */
svb_status svb_gather_picture(svb_context* dst_ctx, const svb_picture* pict, int wait, svb_picture** out);

/* ---- PictureAnimator's state -> matrices (animator.pic.swift:107-128,207-272,326-333) ------------------ */
typedef struct svb_element_state {   /* the ElementState fields computePictureState reads (Proto/Composition.proto:56-71) */
    float pic_pos[3];
    float size[2];
    float texture_offset[2];
    float border_size[4];            /* left, top, right, bottom */
    float fill_color[4];
    float rotation, transparency;
    int32_t pic_aspect;              /* 0 none, 1 aspectFit, 2 aspectFill */
    int32_t pic_origin;              /* 0 centre, 1 top-left */
    int32_t has_fill_color;
    int32_t hidden;                  /* impl() emits nothing for a hidden element (animator.pic.swift:108-110) */
    uint32_t parent_anchors;         /* set of SVB_ANCHOR_*; 0 = { top-left } (:62) */
} svb_element_state;
enum { SVB_ANCHOR_TOP_LEFT = 1, SVB_ANCHOR_TOP_RIGHT = 2, SVB_ANCHOR_BOTTOM_LEFT = 4, SVB_ANCHOR_BOTTOM_RIGHT = 8 }; /* 1 << PictureAnchor */
typedef struct svb_computed_picture_state { /* ComputedPictureState animator.pic.swift:141-147; matrices unprojected, Matrix4 memory order */
    float matrix[16], texture_matrix[16], border_matrix[16];
    float fill_color[4];
    float opacity;
} svb_computed_picture_state;
/* PictureAnimator.impl: `pict` re-issued with matrix = ortho(canvas) * T(pos) * Rz(rotation) * S(size), textureMatrix by
 * aspect mode, borderMatrix, fillColor, opacity = (1 - transparency) * parent_opacity, revision (NULL keeps it). */
svb_status svb_animate_picture(const svb_picture* pict, float canvas_width, float canvas_height, const svb_element_state* state,
                               float parent_opacity, const char* revision, svb_picture** out);
/* computePictureState animator.pic.swift:229-272 as a pure function: `next` + `pct` (both non-NULL) interpolate the transition
 * (computeElementState :193-205); parent_matrix / initial_parent_matrix are the parent's unprojected ComputedPictureState.matrix now and
 * when the element first saw it (NULL = none: position 0, size 0); anchors move the parent's size change onto the element's edges
 * (computePositionSize :149-191). */
svb_status svb_compute_picture_state(float sample_width, float sample_height, const svb_element_state* current, const svb_element_state* next,
                                     const float* pct, uint32_t anchors, const float* parent_matrix, const float* initial_parent_matrix,
                                     svb_computed_picture_state* out);
/* PictureAnimator animator.pic.swift:29-139 without its Clock: the caller passes `now` (seconds) wherever the reference reads clock.current();
 * a transition ends -- next state becomes current, anchors re-read (:67-75) -- the first time `now` reaches start + duration. */
typedef struct svb_animator svb_animator;
svb_status svb_animator_create(float canvas_width, float canvas_height, svb_animator* parent, uint32_t parent_anchors, svb_animator** out); /* init :30-51 */
void svb_animator_destroy(svb_animator* animator);
const char* svb_animator_revision(const svb_animator* animator);
svb_status svb_animator_set_state(svb_animator* animator, const svb_element_state* state, double duration_seconds, double now);          /* setState :54-80 */
svb_status svb_animator_set_parent(svb_animator* animator, svb_animator* parent);                                                      /* setParent :104-106 */
/* computedState :82-102; SVB_ERROR_INVALID_VALUE ("noCurrentState") before the first setState */
svb_status svb_animator_computed_state(svb_animator* animator, float sample_width, float sample_height, double now,
                                       const svb_computed_picture_state* parent_state, svb_computed_picture_state* out);
/* impl :107-128: *out = the sample re-issued with projected matrices, fill colour, opacity x parent opacity and the animator's revision;
 * *out = NULL (status OK) when the reference returns .nothing (hidden element, no state on the element or its parent) */
svb_status svb_animator_apply(svb_animator* animator, const svb_picture* pict, double now, svb_picture** out);

/* ---- VideoMixer -------------------------------------------------------------------------------------- */
/* VideoMixer.init mix.video.swift:22-30; ctx NULL = makeComputeContext(forType: .GPU); asset_id NULL = generated */
svb_status svb_video_mixer_create(const svb_context* ctx, float width, float height, int pixel_format, const char* asset_id,
                                  const char* workspace_id, int64_t frame_duration, int64_t timescale, int64_t epoch, svb_mixer** out);
void svb_video_mixer_destroy(svb_mixer* mixer);
const char* svb_video_mixer_asset_id(const svb_mixer* mixer);
svb_status svb_video_mixer_set_mode(svb_mixer* mixer, int mode);
/* the ingest closure mix.video.swift:57-75: *stored = 1 when kept as a layer, 0 when passed through */
svb_status svb_video_mixer_push(svb_mixer* mixer, const svb_picture* pict, int* stored);
svb_status svb_video_mixer_push_many(svb_mixer* mixer, const svb_picture* const* picts, int count);
/* mix(at:) mix.video.swift:95-140: returns the emitted sample (a GPU PictureSample on the backing ring) */
svb_status svb_video_mixer_mix(svb_mixer* mixer, int64_t time, int wait, svb_picture** out);
/* several mixers of one context folded into one launch */
svb_status svb_video_mixer_mix_many(svb_mixer* const* mixers, int count, int64_t time, int wait, svb_picture** outs);
/* One tick of `count` mixers of one context, host buffers in and host buffers out, in ONE call (ours): every CPU sample of `layers`
 * (layer_counts[i] of them for mixer i, in order; they carry their placement like any pushed sample) goes through GPUBarrierUpload,
 * is pushed, the mixers are mixed in one launch, and each emitted frame goes through GPUBarrierDownload.  Nothing is waited for
 * unless wait != 0: outs[i] are CPU samples whose bytes are valid after svb_picture_wait(outs[i]), and the layers' host bytes must
 * stay untouched until then.  GPU samples among `layers` are pushed as they are.  (Per step of 8 mixers x 8 layers this replaces
 * 81 calls across the boundary by one.) */
svb_status svb_video_mixer_tick_many(svb_mixer* const* mixers, int count, const svb_picture* const* layers, const int* layer_counts, int64_t time,
                                     int wait, svb_picture** outs);
/* clear + fold with explicit uniforms (layers already in z-order) */
svb_status svb_compose(svb_context* ctx, const svb_picture* target, const svb_picture* const* layers, const svb_image_uniforms* uniforms,
                       int count, int mode);

/* ---- device-side timing (bench) ---------------------------------------------------------------------- */
svb_status svb_timer_create(svb_context* ctx, svb_timer** out);
svb_status svb_timer_start(svb_timer* t);  /* after everything queued so far on the context's streams */
svb_status svb_timer_stop(svb_timer* t);   /* after everything queued so far on the context's streams */
svb_status svb_timer_elapsed_ms(svb_timer* t, float* ms); /* waits for stop */
void svb_timer_destroy(svb_timer* t);
/* CUDA events around every fused-compose kernel on the compute stream: total device time and launch count
 * since timing was (re-)enabled.  Reading waits for the launches queued so far. */
svb_status svb_launch_timing(svb_context* ctx, int enable);
svb_status svb_launch_timing_read(svb_context* ctx, double* total_ms, unsigned long long* launches);
/* host time spent inside the fused compose calls while launch timing was on (planner + driver calls: what the calling thread pays per tick),
 * not counting the time the thread was blocked because it ran eight launches ahead of the GPU; that back-pressure is *wait_ms (may be NULL) */
svb_status svb_host_timing_read(svb_context* ctx, double* total_ms, unsigned long long* calls);
svb_status svb_host_timing_read2(svb_context* ctx, double* total_ms, unsigned long long* calls, double* wait_ms);
/* An extension (the reference recomputes every pixel's coordinates in every launch, kernels.cl.swift:70-78; it has nothing to cache): the
 * fused compositor derives per-column / per-row coordinate tables from the layers' uniforms in a pre-pass, and keeps them while a batch's
 * geometry -- frame sizes, every layer's ImageUniforms, source size and format -- stays what it was when the batch's buffer was last filled
 * (a mixer whose layout does not move: the steady state of a live mix).  Pixels are never cached.  enable = 0: run the pre-pass for every
 * launch.  Default: on. */
svb_status svb_table_cache(svb_context* ctx, int enable);
/* launches of our kernels issued by this process so far */
unsigned long long svb_kernel_launch_count(void);
/* 256 floats each: UNORM8 read by the division-free identity and by true division (device self-test) */
svb_status svb_selftest_unorm(svb_context* ctx, float* fast256, float* divided256);

#ifdef __cplusplus
}
#endif
#endif
