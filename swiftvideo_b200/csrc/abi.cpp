// abi.cpp -- the C ABI declared in include/svb200.h over the C++ host side.
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/svb200.h"
#include "compute.h"
#include "cu_driver.h"
#include "mix_video.h"
#include "animator.h"
#include "host_prof.h"

using namespace svb;

struct svb_context {
    ComputeContext c;
};
struct svb_picture {
    std::shared_ptr<const PictureSample> p;
};
struct svb_mixer {
    std::unique_ptr<VideoMixer> m;
};
struct svb_animator {
    std::shared_ptr<PictureAnimator> a;
};
struct svb_timer {
    std::shared_ptr<InternalContext> ic;
    CUevent e0 = nullptr, e1 = nullptr, tmp = nullptr;
};

static thread_local std::string g_err;

template <class F>
static svb_status guard(F&& f) {
    try {
        f();
        g_err.clear();
        return SVB_OK;
    } catch (const ComputeError& e) {
        g_err = e.what();
        return (svb_status)(int)e.code;
    } catch (const std::bad_alloc&) {
        g_err = "out of host memory";
        return SVB_ERROR_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        g_err = e.what();
        return SVB_ERROR_UNKNOWN;
    }
}
static void need(const void* p, const char* what) {
    if (!p) throw ComputeError(ErrorCode::invalidValue, std::string(what) + " is NULL");
}
static svb_picture* wrap(PictureSample&& s) { return new svb_picture{std::make_shared<const PictureSample>(std::move(s))}; }

#pragma GCC visibility push(default)
extern "C" {

const char* svb_last_error(void) { return g_err.c_str(); }
const char* svb_version(void) { return "svb200 0.1 (sm_100a)"; }

int svb_available_compute_devices(void) { return (int)availableComputeDevices().size(); }
int svb_has_available_compute_devices(int device_type) {
    int n = 0;
    for (const ComputeDevice& d : availableComputeDevices())
        if ((int)d.deviceType == device_type && d.available) ++n;
    return n > 0;
}
svb_status svb_make_compute_context(int device_type, int device_index, svb_context** out) {
    return guard([&] {
        need(out, "out");
        *out = new svb_context{makeComputeContext((ComputeDeviceType)device_type, device_index)};
    });
}
svb_status svb_create_compute_context_sharing(const svb_context* sharing, svb_context** out) {
    return guard([&] {
        need(sharing, "sharing");
        need(out, "out");
        *out = new svb_context{createComputeContext(sharing->c)};
    });
}
svb_status svb_destroy_compute_context(svb_context* ctx) {
    return guard([&] {
        if (!ctx) return;
        destroyComputeContext(ctx->c);
        delete ctx;
    });
}
svb_status svb_begin_compute_pass(svb_context* ctx) {
    return guard([&] {
        need(ctx, "ctx");
        ctx->c = beginComputePass(ctx->c);
    });
}
svb_status svb_end_compute_pass(svb_context* ctx, int wait) {
    return guard([&] {
        need(ctx, "ctx");
        if (!ctx->c.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
        ctx->c = endComputePass(ctx->c, wait != 0);
    });
}
int svb_context_device_index(const svb_context* ctx) { return ctx && ctx->c.ctx ? ctx->c.ctx->deviceIndex : -1; }
int svb_context_sm_count(const svb_context* ctx) { return ctx && ctx->c.ctx ? ctx->c.ctx->smCount : 0; }

svb_status svb_kernel_module_image(const void** image, size_t* size) {
    return guard([&] {
        need(image, "image");
        need(size, "size");
        kernelModuleImage(image, size);
    });
}
svb_status svb_default_compute_kernel_from_string(const char* name, int* kernel) {
    return guard([&] {
        need(name, "name");
        need(kernel, "kernel");
        *kernel = (int)defaultComputeKernelFromString(name);
    });
}
const char* svb_compute_kernel_name(int kernel) {
    return kernel >= 0 && kernel < (int)ComputeKernel::count_ ? computeKernelName((ComputeKernel)kernel) : "invalid";
}
svb_status svb_build_compute_kernel(svb_context* ctx, const char* name, const void* image) {
    return guard([&] {
        need(ctx, "ctx");
        need(name, "name");
        ctx->c = buildComputeKernel(ctx->c, name, image);
    });
}
svb_status svb_build_compute_kernel_from_source(svb_context* ctx, const char* name, const char* source) {
    return guard([&] {
        need(ctx, "ctx");
        need(name, "name");
        need(source, "source");
        ctx->c = buildComputeKernelFromSource(ctx->c, name, source);
    });
}
svb_status svb_run_compute_kernel(svb_context* ctx, const svb_picture* const* images, int image_count, const svb_picture* target, int kernel,
                                  const char* custom_name, int max_planes, const void* uniforms, size_t uniforms_size, int blends) {
    return guard([&] {
        need(ctx, "ctx");
        need(target, "target");
        std::vector<const PictureSample*> im;
        if (image_count < 0) throw ComputeError(ErrorCode::invalidValue, "negative image count");
        if (image_count > 0) need(images, "images");
        for (int i = 0; i < image_count; ++i) {
            need(images[i], "image");
            im.push_back(images[i]->p.get());
        }
        if (kernel < 0 || kernel >= (int)ComputeKernel::count_) throw ComputeError(ErrorCode::invalidValue, "bad kernel id");
        ctx->c = runComputeKernel(ctx->c, im, *target->p, (ComputeKernel)kernel, custom_name ? custom_name : "", max_planes, uniforms,
                                  uniforms_size, blends != 0);
    });
}
svb_status svb_apply_compute_image(svb_context* ctx, const svb_picture* image, const svb_picture* target, int kernel) {
    return guard([&] {
        need(ctx, "ctx");
        need(image, "image");
        need(target, "target");
        if (kernel < 0 || kernel >= (int)ComputeKernel::count_) throw ComputeError(ErrorCode::invalidValue, "bad kernel id");
        ctx->c = applyComputeImage(ctx->c, *image->p, *target->p, (ComputeKernel)kernel);
    });
}
svb_status svb_make_image_uniforms(const svb_picture* image, const svb_picture* target, svb_image_uniforms* out) {
    return guard([&] {
        need(image, "image");
        need(target, "target");
        need(out, "out");
        const ImageUniforms u = makeImageUniforms(*image->p, *target->p);
        static_assert(sizeof(svb_image_uniforms) == sizeof(ImageUniforms), "uniform layout");
        std::memcpy(out, &u, sizeof(u));
    });
}

svb_status svb_create_picture_sample(float width, float height, int pixel_format, const char* asset_id, const char* workspace_id,
                                     svb_context* pinned_from, svb_picture** out) {
    return guard([&] {
        need(out, "out");
        if (pixel_format < 0 || pixel_format > (int)PixelFormat::p010) throw ComputeError(ErrorCode::badInputData, "Invalid pixel format");
        *out = wrap(createPictureSample(Vector2{width, height}, (PixelFormat)pixel_format, asset_id ? asset_id : "", workspace_id ? workspace_id : "",
                                        pinned_from ? &pinned_from->c : nullptr));
    });
}
svb_status svb_picture_sample_from_planes(float width, float height, int pixel_format, const void* const* planes, const int32_t* strides,
                                          int plane_count, const char* asset_id, const char* workspace_id, svb_context* pinned_from, svb_picture** out) {
    return guard([&] {
        need(planes, "planes");
        need(strides, "strides");
        need(out, "out");
        if (pixel_format < 0 || pixel_format > (int)PixelFormat::p010) throw ComputeError(ErrorCode::badInputData, "Invalid pixel format");
        if (plane_count < 1 || plane_count > 3) throw ComputeError(ErrorCode::badInputData, "Input image must have 1, 2, or 3 planes");
        const uint8_t* p[3] = {nullptr, nullptr, nullptr};
        int st[3] = {0, 0, 0};
        for (int i = 0; i < plane_count; ++i) p[i] = (const uint8_t*)planes[i], st[i] = strides[i];
        *out = wrap(pictureSampleFromPlanes((PixelFormat)pixel_format, Vector2{width, height}, p, st, plane_count, asset_id ? asset_id : "",
                                            workspace_id ? workspace_id : "", pinned_from ? &pinned_from->c : nullptr));
    });
}
svb_status svb_picture_with(const svb_picture* other, const float* matrix, const float* texture_matrix, const float* border_matrix,
                            const float* fill_color, const float* opacity, const char* revision, const char* asset_id, svb_picture** out) {
    return guard([&] {
        need(other, "other");
        need(out, "out");
        PictureSample s = *other->p;  // sample.pict.linux.swift:194-226: every argument defaults to other's value
        if (matrix) s.transform = Matrix4::from_array(matrix);
        if (texture_matrix) s.texTransform = Matrix4::from_array(texture_matrix);
        if (border_matrix) s.borderTransform = Matrix4::from_array(border_matrix);
        if (fill_color) s.bgColor = Vector4{fill_color[0], fill_color[1], fill_color[2], fill_color[3]};
        if (opacity) s.alpha = *opacity;
        if (revision) s.idRevision = revision;
        if (asset_id) s.idAsset = asset_id;
        *out = wrap(std::move(s));
    });
}
svb_status svb_picture_info_get(const svb_picture* pict, svb_picture_info* out) {
    return guard([&] {
        need(pict, "pict");
        need(out, "out");
        const PictureSample& s = *pict->p;
        std::memset(out, 0, sizeof(*out));
        out->pixel_format = (int)s.pixelFormat();
        out->buffer_type = (int)s.bufferType();
        out->width = s.size().x;
        out->height = s.size().y;
        out->plane_count = (int)s.imgBuffer.planes.size();
        for (int i = 0; i < out->plane_count && i < 3; ++i) {
            const Plane& p = s.imgBuffer.planes[i];
            svb_plane_info& o = out->planes[i];
            o.width = p.size.x, o.height = p.size.y, o.stride = p.stride, o.bit_depth = p.bitDepth, o.components = (int)p.components.size();
            o.size = (size_t)p.stride * (size_t)(int)p.size.y;
            o.host = i < (int)s.imgBuffer.buffers.size() ? s.imgBuffer.buffers[i].ptr : nullptr;
            o.device = i < (int)s.imgBuffer.computeTextures.size() ? s.imgBuffer.computeTextures[i]->mem : 0;
        }
        std::memcpy(out->matrix, s.transform.data(), 64);
        std::memcpy(out->texture_matrix, s.texTransform.data(), 64);
        std::memcpy(out->border_matrix, s.borderTransform.data(), 64);
        out->fill_color[0] = s.bgColor.x, out->fill_color[1] = s.bgColor.y, out->fill_color[2] = s.bgColor.z, out->fill_color[3] = s.bgColor.w;
        out->opacity = s.alpha;
        out->z_index = s.zIndex();
        out->pts = s.ptsValue, out->time = s.timeValue, out->timescale = s.timescale;
    });
}
svb_status svb_picture_wait(const svb_picture* pict) {
    return guard([&] {
        need(pict, "pict");
        waitPicture(*pict->p);
    });
}
void svb_picture_release(svb_picture* pict) {
    SVB_PROF(23, "abi: picture_release");
    delete pict;
}
const char* svb_picture_revision(const svb_picture* pict) { return pict ? pict->p->revision().c_str() : nullptr; }
const char* svb_picture_asset_id(const svb_picture* pict) { return pict ? pict->p->assetId().c_str() : nullptr; }
unsigned long long svb_picture_identity(const svb_picture* pict) { return pict ? (unsigned long long)(uintptr_t)pict->p.get() : 0ull; }

svb_status svb_upload_compute_picture(svb_context* ctx, const svb_picture* pict, int max_planes, int retain_cpu_buffer, int wait, svb_picture** out) {
    return guard([&] {
        need(ctx, "ctx");
        need(pict, "pict");
        need(out, "out");
        *out = wrap(uploadComputePicture(ctx->c, *pict->p, max_planes, retain_cpu_buffer != 0, wait != 0));
    });
}
svb_status svb_upload_compute_pictures(svb_context* ctx, const svb_picture* const* picts, int count, int max_planes, int retain_cpu_buffer, int wait, svb_picture** outs) {
    return guard([&] {
        need(ctx, "ctx");
        if (count < 0) throw ComputeError(ErrorCode::invalidValue, "negative picture count");
        if (count > 0) {
            need(picts, "picts");
            need(outs, "outs");
        }
        std::vector<const PictureSample*> ps;
        for (int i = 0; i < count; ++i) {
            need(picts[i], "pict");
            ps.push_back(picts[i]->p.get());
        }
        std::vector<PictureSample> up = uploadComputePictures(ctx->c, ps, max_planes, retain_cpu_buffer != 0, wait != 0);
        for (int i = 0; i < count; ++i) outs[i] = wrap(std::move(up[(size_t)i]));
    });
}
static svb_status barrier(const BarrierResult& r, const svb_picture* pict, svb_picture** out, svb_event_error* err) {
    if (r.ok) {
        // an untouched sample passes through as another handle of the same sample
        *out = r.sample.bufferType() == pict->p->bufferType() ? new svb_picture{pict->p} : wrap(PictureSample(r.sample));
        g_err.clear();
        return SVB_OK;
    }
    g_err = r.error.description;
    if (err) {
        std::memset(err, 0, sizeof(*err));
        std::strncpy(err->domain, r.error.domain.c_str(), sizeof(err->domain) - 1);
        err->code = r.error.code;
        std::strncpy(err->description, r.error.description.c_str(), sizeof(err->description) - 1);
        std::strncpy(err->asset_id, r.error.assetId.c_str(), sizeof(err->asset_id) - 1);
    }
    return SVB_ERROR_UNKNOWN;
}
svb_status svb_gpu_barrier_upload(svb_context* ctx, const svb_picture* pict, int retain_cpu_buffer, int wait, svb_picture** out, svb_event_error* err) {
    svb_status st = SVB_OK;
    const svb_status g = guard([&] {
        need(ctx, "ctx");
        need(pict, "pict");
        need(out, "out");
        st = barrier(GPUBarrierUpload(ctx->c, retain_cpu_buffer != 0)(*pict->p, wait != 0), pict, out, err);
    });
    return g != SVB_OK ? g : st;
}
svb_status svb_gpu_barrier_download(svb_context* ctx, const svb_picture* pict, int retain_gpu_buffer, int wait, svb_picture** out, svb_event_error* err) {
    svb_status st = SVB_OK;
    const svb_status g = guard([&] {
        need(ctx, "ctx");
        need(pict, "pict");
        need(out, "out");
        st = barrier(GPUBarrierDownload(ctx->c, retain_gpu_buffer != 0)(*pict->p, wait != 0), pict, out, err);
    });
    return g != SVB_OK ? g : st;
}
svb_status svb_download_compute_picture(svb_context* ctx, const svb_picture* pict, int retain_gpu_buffer, int wait, svb_picture** out) {
    return guard([&] {
        need(ctx, "ctx");
        need(pict, "pict");
        need(out, "out");
        *out = wrap(downloadComputePicture(ctx->c, *pict->p, retain_gpu_buffer != 0, wait != 0));
    });
}

svb_status svb_scale_convert_picture(svb_context* ctx, const svb_picture* src, float dst_width, float dst_height, int dst_pixel_format, int filter, int wait,
                                     svb_picture** out) {
    return guard([&] {
        need(ctx, "ctx");
        need(src, "src");
        need(out, "out");
        if (filter != 0 && filter != 1) throw ComputeError(ErrorCode::invalidValue, "unknown scale filter");
        if (dst_pixel_format < 0 || dst_pixel_format > (int)PixelFormat::p010) throw ComputeError(ErrorCode::invalidValue, "unknown pixel format");
        *out = wrap(scaleConvertPicture(ctx->c, *src->p, Vector2{dst_width, dst_height}, (PixelFormat)dst_pixel_format, (ScaleFilter)filter, wait != 0));
    });
}
svb_status svb_scale_filter_table(int filter, int src_n, int dst_n, int32_t* first, float* weights, int weights_capacity, int* taps) {
    return guard([&] {
        need(taps, "taps");
        if (filter != 0 && filter != 1) throw ComputeError(ErrorCode::invalidValue, "unknown scale filter");
        const ScaleTable t = makeScaleTable((ScaleFilter)filter, src_n, dst_n);
        *taps = t.taps;
        if (first && weights) {
            if (weights_capacity < (int)t.weights.size()) throw ComputeError(ErrorCode::invalidValue, "weights_capacity too small");
            std::memcpy(first, t.first.data(), sizeof(int32_t) * t.first.size());
            std::memcpy(weights, t.weights.data(), sizeof(float) * t.weights.size());
        }
    });
}

svb_status svb_picture_device_frame(const svb_picture* pict, svb_device_frame* out) {
    return guard([&] {
        need(pict, "pict");
        need(out, "out");
        const PictureSample& p = *pict->p;
        if (p.bufferType() != BufferType::gpu || p.imgBuffer.computeTextures.empty()) throw ComputeError(ErrorCode::badInputData, "not a GPU sample");
        std::memset(out, 0, sizeof *out);
        const auto& ic = p.imgBuffer.computeTextures[0]->ctx;
        out->device_index = ic->deviceIndex;
        out->pixel_format = (int32_t)p.pixelFormat();
        out->width = p.size().x, out->height = p.size().y;
        out->plane_count = (int32_t)std::min<size_t>(3, p.imgBuffer.computeTextures.size());
        for (int i = 0; i < out->plane_count; ++i) {
            const Plane& pl = p.imgBuffer.planes[i];
            out->planes[i].ptr = (unsigned long long)p.imgBuffer.computeTextures[i]->mem;
            out->planes[i].pitch = pl.stride;
            out->planes[i].width_bytes = (int32_t)pl.size.x * (int32_t)pl.components.size() * ((pl.bitDepth + 7) / 8);
            out->planes[i].rows = (int32_t)pl.size.y;
        }
        out->context = ic->ctx;
        out->ready_event = pictureReadyEvent(p);
    });
}

svb_status svb_picture_consumed_on(const svb_picture* pict, void* consumer_stream) {
    return guard([&] {
        need(pict, "pict");
        pictureConsumedOn(*pict->p, (CUstream)consumer_stream);
    });
}

svb_status svb_gather_picture(svb_context* dst_ctx, const svb_picture* pict, int wait, svb_picture** out) {
    return guard([&] {
        need(dst_ctx, "dst_ctx");
        need(pict, "pict");
        need(out, "out");
        *out = wrap(gatherComputePicture(dst_ctx->c, *pict->p, wait != 0));
    });
}

static ElementState elementState(const svb_element_state* st) {
    ElementState e;
    e.picPos = Vector3{st->pic_pos[0], st->pic_pos[1], st->pic_pos[2]};
    e.size = Vector2{st->size[0], st->size[1]};
    e.textureOffset = Vector2{st->texture_offset[0], st->texture_offset[1]};
    e.borderSize = Vector4{st->border_size[0], st->border_size[1], st->border_size[2], st->border_size[3]};
    e.fillColor = Vector4{st->fill_color[0], st->fill_color[1], st->fill_color[2], st->fill_color[3]};
    e.rotation = st->rotation, e.transparency = st->transparency;
    if (st->pic_aspect < 0 || st->pic_aspect > 2 || st->pic_origin < 0 || st->pic_origin > 1 || st->parent_anchors > 15u)
        throw ComputeError(ErrorCode::invalidValue, "bad element state");
    e.picAspect = (AspectMode)st->pic_aspect, e.picOrigin = (PicOrigin)st->pic_origin, e.hasFillColor = st->has_fill_color != 0;
    e.hidden = st->hidden != 0, e.parentAnchor = st->parent_anchors;
    return e;
}

static void computedOut(const ComputedPictureState& cs, svb_computed_picture_state* out) {
    std::memcpy(out->matrix, cs.matrix.data(), 64);
    std::memcpy(out->texture_matrix, cs.textureMatrix.data(), 64);
    std::memcpy(out->border_matrix, cs.borderMatrix.data(), 64);
    out->fill_color[0] = cs.fillColor.x, out->fill_color[1] = cs.fillColor.y, out->fill_color[2] = cs.fillColor.z, out->fill_color[3] = cs.fillColor.w;
    out->opacity = cs.opacity;
}

svb_status svb_animate_picture(const svb_picture* pict, float canvas_width, float canvas_height, const svb_element_state* st, float parent_opacity,
                               const char* revision, svb_picture** out) {
    return guard([&] {
        need(pict, "pict");
        need(st, "state");
        need(out, "out");
        *out = wrap(animatePicture(*pict->p, Vector2{canvas_width, canvas_height}, elementState(st), parent_opacity, revision ? revision : ""));
    });
}

svb_status svb_compute_picture_state(float sample_width, float sample_height, const svb_element_state* current, const svb_element_state* next,
                                     const float* pct, uint32_t anchors, const float* parent_matrix, const float* initial_parent_matrix,
                                     svb_computed_picture_state* out) {
    return guard([&] {
        need(current, "current");
        need(out, "out");
        if (anchors > 15u) throw ComputeError(ErrorCode::invalidValue, "bad anchor set");
        const ElementState cur = elementState(current);
        ElementState nxt;
        Matrix4 parent, initial;
        PictureStateInputs in;
        if (next) nxt = elementState(next), in.next = &nxt;
        if (pct) in.pct = *pct;
        if (parent_matrix) parent = Matrix4::from_array(parent_matrix), in.parent = &parent;
        if (initial_parent_matrix) initial = Matrix4::from_array(initial_parent_matrix), in.initialParent = &initial;
        in.anchors = anchors ? anchors : (unsigned)anchorTopLeft;
        computedOut(computePictureState(Vector2{sample_width, sample_height}, cur, in), out);
    });
}

svb_status svb_animator_create(float canvas_width, float canvas_height, svb_animator* parent, uint32_t parent_anchors, svb_animator** out) {
    return guard([&] {
        need(out, "out");
        if (parent_anchors > 15u) throw ComputeError(ErrorCode::invalidValue, "bad anchor set");
        auto* h = new svb_animator;
        h->a = std::make_shared<PictureAnimator>(Vector2{canvas_width, canvas_height}, parent ? parent->a : nullptr, parent_anchors);
        *out = h;
    });
}

void svb_animator_destroy(svb_animator* animator) { delete animator; }

const char* svb_animator_revision(const svb_animator* animator) { return animator ? animator->a->revision().c_str() : nullptr; }

svb_status svb_animator_set_state(svb_animator* animator, const svb_element_state* state, double duration_seconds, double now) {
    return guard([&] {
        need(animator, "animator");
        need(state, "state");
        animator->a->setState(elementState(state), duration_seconds, now);
    });
}

svb_status svb_animator_set_parent(svb_animator* animator, svb_animator* parent) {
    return guard([&] {
        need(animator, "animator");
        animator->a->setParent(parent ? parent->a : nullptr);
    });
}

svb_status svb_animator_computed_state(svb_animator* animator, float sample_width, float sample_height, double now,
                                       const svb_computed_picture_state* parent_state, svb_computed_picture_state* out) {
    return guard([&] {
        need(animator, "animator");
        need(out, "out");
        ComputedPictureState ps;
        if (parent_state) {
            ps.matrix = Matrix4::from_array(parent_state->matrix);
            ps.opacity = parent_state->opacity;
        }
        computedOut(animator->a->computedState(Vector2{sample_width, sample_height}, now, parent_state ? &ps : nullptr), out);
    });
}

svb_status svb_animator_apply(svb_animator* animator, const svb_picture* pict, double now, svb_picture** out) {
    return guard([&] {
        need(animator, "animator");
        need(pict, "pict");
        need(out, "out");
        PictureSample res;
        *out = animator->a->apply(*pict->p, now, &res) ? wrap(std::move(res)) : nullptr;
    });
}

svb_status svb_video_mixer_create(const svb_context* ctx, float width, float height, int pixel_format, const char* asset_id,
                                  const char* workspace_id, int64_t frame_duration, int64_t timescale, int64_t epoch, svb_mixer** out) {
    return guard([&] {
        need(out, "out");
        if (pixel_format < 0 || pixel_format > (int)PixelFormat::invalid) throw ComputeError(ErrorCode::badInputData, "Invalid pixel format");
        *out = new svb_mixer{std::make_unique<VideoMixer>(ctx ? &ctx->c : nullptr, Vector2{width, height}, (PixelFormat)pixel_format,
                                                          asset_id ? asset_id : "", workspace_id ? workspace_id : "", frame_duration,
                                                          timescale, epoch)};
    });
}
void svb_video_mixer_destroy(svb_mixer* mixer) { delete mixer; }
const char* svb_video_mixer_asset_id(const svb_mixer* mixer) { return mixer ? mixer->m->assetId().c_str() : ""; }
svb_status svb_video_mixer_set_mode(svb_mixer* mixer, int mode) {
    return guard([&] {
        need(mixer, "mixer");
        if (mode < 0 || mode > 5) throw ComputeError(ErrorCode::invalidValue, "bad mix mode");
        mixer->m->setMode((VideoMixer::Mode)mode);
    });
}
svb_status svb_video_mixer_push(svb_mixer* mixer, const svb_picture* pict, int* stored) {
    return guard([&] {
        need(mixer, "mixer");
        need(pict, "pict");
        const bool s = mixer->m->push(pict->p);
        if (stored) *stored = s ? 1 : 0;
    });
}
svb_status svb_video_mixer_push_many(svb_mixer* mixer, const svb_picture* const* picts, int count) {
    SVB_PROF(20, "abi: push_many");
    return guard([&] {
        need(mixer, "mixer");
        for (int i = 0; i < count; ++i) {
            need(picts[i], "pict");
            mixer->m->push(picts[i]->p);
        }
    });
}
svb_status svb_video_mixer_mix(svb_mixer* mixer, int64_t time, int wait, svb_picture** out) {
    SVB_PROF(21, "abi: mix");
    return guard([&] {
        need(mixer, "mixer");
        need(out, "out");
        *out = wrap(mixer->m->mix(time, wait != 0));
    });
}
svb_status svb_video_mixer_mix_many(svb_mixer* const* mixers, int count, int64_t time, int wait, svb_picture** outs) {
    SVB_PROF(22, "abi: mix_many");
    return guard([&] {
        need(mixers, "mixers");
        need(outs, "outs");
        std::vector<VideoMixer*> ms;
        for (int i = 0; i < count; ++i) {
            need(mixers[i], "mixer");
            ms.push_back(mixers[i]->m.get());
        }
        std::vector<PictureSample> res(count);
        VideoMixer::mixMany(ms.data(), count, time, res.data(), wait != 0);
        for (int i = 0; i < count; ++i) outs[i] = wrap(std::move(res[i]));
    });
}
svb_status svb_video_mixer_tick_many(svb_mixer* const* mixers, int count, const svb_picture* const* layers, const int* layer_counts, int64_t time,
                                     int wait, svb_picture** outs) {
    return guard([&] {
        need(mixers, "mixers");
        need(outs, "outs");
        if (count <= 0) throw ComputeError(ErrorCode::invalidValue, "tick_many: no mixers");
        need(layer_counts, "layer_counts");
        std::vector<VideoMixer*> ms;
        size_t at = 0;
        std::vector<const PictureSample*> all;
        for (int i = 0; i < count; ++i) {
            need(mixers[i], "mixer");
            ms.push_back(mixers[i]->m.get());
            if (!ms.back()->computeContext()) throw ComputeError(ErrorCode::badContextState, "No context");
            if (layer_counts[i] < 0) throw ComputeError(ErrorCode::invalidValue, "negative layer count");
            if (layer_counts[i] > 0) need(layers, "layers");
            for (int k = 0; k < layer_counts[i]; ++k, ++at) {
                need(layers[at], "layer");
                all.push_back(layers[at]->p.get());
            }
        }
        // the tick's CPU layers go up together: neighbours in page-locked memory travel as one copy (mixers of one tick share a context:
        // mixMany checks it)
        for (VideoMixer* m : ms)
            if (m->computeContext()->ctx != ms[0]->computeContext()->ctx) throw ComputeError(ErrorCode::invalidContext, "tick_many: mixers must share one compute context");
        std::vector<PictureSample> up = uploadComputePictures(*ms[0]->computeContext(), all, 3, false, false);
        at = 0;
        for (int i = 0; i < count; ++i)
            for (int k = 0; k < layer_counts[i]; ++k, ++at) {
                if (layers[at]->p->bufferType() == BufferType::cpu) ms[i]->push(std::move(up[at]));
                else ms[i]->push(layers[at]->p);
            }
        std::vector<PictureSample> res(count);
        VideoMixer::mixMany(ms.data(), count, time, res.data(), false);
        for (int i = 0; i < count; ++i) outs[i] = wrap(downloadComputePicture(*ms[i]->computeContext(), res[i], true, wait != 0));
    });
}
svb_status svb_compose(svb_context* ctx, const svb_picture* target, const svb_picture* const* layers, const svb_image_uniforms* uniforms,
                       int count, int mode) {
    return guard([&] {
        need(ctx, "ctx");
        need(target, "target");
        if (mode < 0 || mode > 5) throw ComputeError(ErrorCode::invalidValue, "bad mix mode");
        if (count < 0) throw ComputeError(ErrorCode::invalidValue, "negative layer count");
        if (count > 0) {
            need(layers, "layers");
            need(uniforms, "uniforms");
        }
        std::vector<const PictureSample*> ls;
        for (int i = 0; i < count; ++i) {
            need(layers[i], "layer");
            ls.push_back(layers[i]->p.get());
        }
        ctx->c = VideoMixer::composeRaw(ctx->c, *target->p, ls, (const ImageUniforms*)uniforms, (VideoMixer::Mode)mode);
    });
}

svb_status svb_timer_create(svb_context* ctx, svb_timer** out) {
    return guard([&] {
        need(ctx, "ctx");
        need(out, "out");
        if (!ctx->c.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
        auto t = std::make_unique<svb_timer>();
        t->ic = ctx->c.ctx;
        CtxGuard g(t->ic);
        check(cu().cuEventCreate(&t->e0, CU_EVENT_DEFAULT), "cuEventCreate");
        check(cu().cuEventCreate(&t->e1, CU_EVENT_DEFAULT), "cuEventCreate");
        check(cu().cuEventCreate(&t->tmp, CU_EVENT_DISABLE_TIMING), "cuEventCreate");
        *out = t.release();
    });
}
// Join the upload and download streams into the compute stream, then stamp the compute stream.
static void stamp(svb_timer* t, CUevent e) {
    CtxGuard g(t->ic);
    for (CUstream s : {t->ic->upload, t->ic->download}) {
        check(cu().cuEventRecord(t->tmp, s), "cuEventRecord");
        check(cu().cuStreamWaitEvent(t->ic->compute, t->tmp, 0), "cuStreamWaitEvent");
    }
    check(cu().cuEventRecord(e, t->ic->compute), "cuEventRecord");
    // and nothing queued later on the side streams may start before the stamp
    check(cu().cuStreamWaitEvent(t->ic->upload, e, 0), "cuStreamWaitEvent");
    check(cu().cuStreamWaitEvent(t->ic->download, e, 0), "cuStreamWaitEvent");
}
svb_status svb_timer_start(svb_timer* t) {
    return guard([&] {
        need(t, "timer");
        stamp(t, t->e0);
    });
}
svb_status svb_timer_stop(svb_timer* t) {
    return guard([&] {
        need(t, "timer");
        stamp(t, t->e1);
    });
}
svb_status svb_timer_elapsed_ms(svb_timer* t, float* ms) {
    return guard([&] {
        need(t, "timer");
        need(ms, "ms");
        CtxGuard g(t->ic);
        check(cu().cuEventSynchronize(t->e1), "cuEventSynchronize");
        check(cu().cuEventElapsedTime(ms, t->e0, t->e1), "cuEventElapsedTime");
    });
}
void svb_timer_destroy(svb_timer* t) {
    if (!t) return;
    if (cu().ok) {
        cu().cuCtxPushCurrent(t->ic->ctx);
        for (CUevent e : {t->e0, t->e1, t->tmp})
            if (e) cu().cuEventDestroy(e);
        CUcontext old;
        cu().cuCtxPopCurrent(&old);
    }
    delete t;
}

svb_status svb_launch_timing(svb_context* ctx, int enable) {
    return guard([&] {
        need(ctx, "ctx");
        if (!ctx->c.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
        setLaunchTiming(ctx->c, enable != 0);
    });
}
svb_status svb_table_cache(svb_context* ctx, int enable) {
    return guard([&] {
        need(ctx, "ctx");
        if (!ctx->c.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
        setTableCache(ctx->c, enable != 0);
    });
}
svb_status svb_launch_timing_read(svb_context* ctx, double* total_ms, unsigned long long* launches) {
    return guard([&] {
        need(ctx, "ctx");
        need(total_ms, "total_ms");
        need(launches, "launches");
        if (!ctx->c.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
        readLaunchTiming(ctx->c, total_ms, launches);
    });
}

svb_status svb_host_timing_read(svb_context* ctx, double* total_ms, unsigned long long* calls) {
    return guard([&] {
        need(ctx, "ctx");
        need(total_ms, "total_ms");
        need(calls, "calls");
        if (!ctx->c.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
        readHostTiming(ctx->c, total_ms, calls);
    });
}

svb_status svb_host_timing_read2(svb_context* ctx, double* total_ms, unsigned long long* calls, double* wait_ms) {
    return guard([&] {
        need(ctx, "ctx");
        need(total_ms, "total_ms");
        need(calls, "calls");
        if (!ctx->c.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
        readHostTiming(ctx->c, total_ms, calls, wait_ms);
    });
}

unsigned long long svb_kernel_launch_count(void) { return kernelLaunchCount(); }

svb_status svb_selftest_unorm(svb_context* ctx, float* fast256, float* divided256) {
    return guard([&] {
        need(ctx, "ctx");
        need(fast256, "fast256");
        need(divided256, "divided256");
        if (!ctx->c.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
        auto ic = ctx->c.ctx;
        CtxGuard g(ic);
        CUdeviceptr buf = ic->alloc(2048);
        CUdeviceptr a = buf, b = buf + 1024;
        void* args[] = {&a, &b};
        check(cu().cuLaunchKernel(ic->builtin("svb_selftest_unorm"), 1, 1, 1, 256, 1, 1, 0, ic->compute, args, nullptr), "cuLaunchKernel");
        noteKernelLaunch();
        check(cu().cuStreamSynchronize(ic->compute), "cuStreamSynchronize");
        check(cu().cuMemcpyDtoH(fast256, a, 1024), "cuMemcpyDtoH");
        check(cu().cuMemcpyDtoH(divided256, b, 1024), "cuMemcpyDtoH");
        ic->release(buf, 2048);
    });
}

}  // extern "C"
#pragma GCC visibility pop
