// ctx.cu -- context, picture memory and profiling plumbing of the B200 rcn backend.
//
// Replaces, on the device side, the plane allocation of the reference's frame pool
// (libovvc/ovframepool.c:194-231 -> mempool.c:73-86): pictures live in HBM with 256-byte aligned
// pitches so that every row start is aligned for 128-bit vector access; host frames keep the OVFrame
// convention (data[3], linesize[3] in bytes, libovvc/ovframe.h:84-93).
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

int ovgpu_set_error(OVGPUCtx *ctx, int code, const char *fmt, ...)
{
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

StageScope::StageScope(OVGPUCtx *c, int stage, int nb_kernels) : ctx(c), slot(-1)
{
    ctx->launches += nb_kernels;
    if (!ctx->profile) return;
    if (!ctx->ev_pool_init) {
        for (int i = 0; i < 2 * OVGPUCtx::MAX_PENDING; i++) cudaEventCreate(&ctx->ev_pool[i]);
        ctx->ev_pool_init = 1;
    }
    if (ctx->nb_pending == OVGPUCtx::MAX_PENDING) {
        double ms; uint64_t n;
        ovgpu_profile_get(ctx, 0, &ms, &n);   // drains pending events
    }
    slot = ctx->nb_pending++;
    ctx->pending[slot].a = ctx->ev_pool[2 * slot];
    ctx->pending[slot].b = ctx->ev_pool[2 * slot + 1];
    ctx->pending[slot].stage = stage;
    cudaEventRecord(ctx->pending[slot].a, ctx->stream);
}

StageScope::~StageScope()
{
    if (slot >= 0) cudaEventRecord(ctx->pending[slot].b, ctx->stream);
}

#include <time.h>
unsigned long long ovgpu_ht_ns[HT_NB], ovgpu_ht_n[HT_NB];
int ovgpu_ht_on = getenv("OVGPU_HOST_TIMING") != NULL;
unsigned long long ovgpu_ht_now()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}

extern "C" {

int ovgpu_ctx_create(OVGPUCtx **out, int device, void *cuda_stream)
{
    if (!out) return OVGPU_EINVAL;
    *out = NULL;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return OVGPU_ENODEV;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return OVGPU_ENODEV;
    if (prop.major != 10) return OVGPU_ENODEV;   // kernels are built for sm_100a only
    OVGPUCtx *ctx = (OVGPUCtx *)calloc(1, sizeof(*ctx));
    if (!ctx) return OVGPU_ENOMEM;
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess) { free(ctx); return OVGPU_ECUDA; }
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { free(ctx); return OVGPU_ECUDA; }
        ctx->own_stream = 1;
    }
    *out = ctx;
    return 0;
}

void ovgpu_ctx_destroy(OVGPUCtx *ctx)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ev_pool_init)
        for (int i = 0; i < 2 * OVGPUCtx::MAX_PENDING; i++) cudaEventDestroy(ctx->ev_pool[i]);
    if (ctx->flush_buf) cudaFree(ctx->flush_buf);
    for (int i = 0; i < OVGPUCtx::NB_AUX; i++) {
        if (ctx->aux[i]) { cudaStreamSynchronize(ctx->aux[i]); cudaStreamDestroy(ctx->aux[i]); }
        if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    free(ctx);
}

int ovgpu_ctx_sync(OVGPUCtx *ctx)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx) return OVGPU_EINVAL;
    OVGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    OVGPU_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

const char *ovgpu_last_error(const OVGPUCtx *ctx) { return ctx ? ctx->err : "no context"; }
void ovgpu_host_timing_dump(void)
{
    static const char *const names[HT_NB] = { "mc commit: validate + route + items", "mc commit: copy calls", "tu commit: validation", "tu commit: item lists",
        "tu commit: copy calls", "tu anchors", "intra upload: validation", "intra upload: levels + maps", "intra upload: copy calls",
        "recon bind: lists", "recon bind: copy calls", "recon: motion compensation launches", "recon: other launches", "staging fence waits (inside the above)" };
    for (int i = 0; i < HT_NB; i++)
        if (ovgpu_ht_n[i]) fprintf(stderr, "host timing %-42s %8llu calls  %8.1f us/call\n", names[i], ovgpu_ht_n[i], 1e-3 * ovgpu_ht_ns[i] / ovgpu_ht_n[i]);
}

uint64_t ovgpu_launch_count(const OVGPUCtx *ctx) { return ctx ? ctx->launches : 0; }

int ovgpu_profile_enable(OVGPUCtx *ctx, int on)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx) return OVGPU_EINVAL;
    ctx->profile = on;
    return 0;
}

int ovgpu_profile_get(OVGPUCtx *ctx, int stage, double *ms, uint64_t *launches)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || stage < 0 || stage >= OVGPU_NB_STAGES) return OVGPU_EINVAL;
    if (ctx->nb_pending) {
        OVGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < ctx->nb_pending; i++) {
            float t = 0.f;
            cudaEventElapsedTime(&t, ctx->pending[i].a, ctx->pending[i].b);
            ctx->prof[ctx->pending[i].stage].ms += t;
            ctx->prof[ctx->pending[i].stage].n += 1;
        }
        ctx->nb_pending = 0;
    }
    if (ms) *ms = ctx->prof[stage].ms;
    if (launches) *launches = ctx->prof[stage].n;
    return 0;
}

int ovgpu_profile_reset(OVGPUCtx *ctx)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx) return OVGPU_EINVAL;
    double ms; uint64_t n;
    ovgpu_profile_get(ctx, 0, &ms, &n);
    memset(ctx->prof, 0, sizeof(ctx->prof));
    return 0;
}

int ovgpu_flush_l2(OVGPUCtx *ctx)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx) return OVGPU_EINVAL;
    if (!ctx->flush_buf) {
        ctx->flush_bytes = (size_t)256 << 20;   // 2x the 126 MB L2
        OVGPU_CUDA_CHECK(ctx, cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
    }
    OVGPU_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->flush_buf, 0x5a, ctx->flush_bytes, ctx->stream));
    return 0;
}

// ---- pictures ---------------------------------------------------------------------------------------
// Device-to-device picture copy on the SMs: a DMA copy would queue on the copy engines behind the pictures being downloaded,
// and the reference hand-off (the next pictures' motion compensation waits for it) must not.
struct CopyPlanes {
    const uint8_t *src[3];
    uint8_t *dst[3];
    uint32_t src_pitch[3], dst_pitch[3], row_bytes[3], rows[3];
};
__global__ void __launch_bounds__(256) ovgpu_copy_planes_kernel(CopyPlanes a)
{
    const int c = blockIdx.y;
    const uint32_t per_row = a.row_bytes[c] >> 3;
    const uint64_t n = (uint64_t)per_row * a.rows[c];
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(i / per_row), x = (uint32_t)(i - (uint64_t)y * per_row);
        *(uint2 *)(a.dst[c] + (size_t)y * a.dst_pitch[c] + 8 * x) = *(const uint2 *)(a.src[c] + (size_t)y * a.src_pitch[c] + 8 * x);
    }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int ovgpu_host_register(OVGPUCtx *ctx, void *ptr, size_t bytes)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !ptr || !bytes) return OVGPU_EINVAL;
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) { cudaGetLastError(); return ovgpu_set_error(ctx, OVGPU_ECUDA, "cudaHostRegister: %s", cudaGetErrorString(e)); }
    return 0;
}

int ovgpu_host_unregister(OVGPUCtx *ctx, void *ptr)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !ptr) return OVGPU_EINVAL;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return ovgpu_set_error(ctx, OVGPU_ECUDA, "cudaHostUnregister: %s", cudaGetErrorString(e)); }
    return 0;
}

int ovgpu_picture_create(OVGPUCtx *ctx, int width, int height, int bitdepth, OVGPUPicture **out)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !out || width <= 0 || height <= 0 || (width & 7) || (height & 7)) return OVGPU_EINVAL;
    if (bitdepth != 10 && bitdepth != 8) return OVGPU_EINVAL;
    OVGPUPicture *pic = (OVGPUPicture *)calloc(1, sizeof(*pic));
    if (!pic) return OVGPU_ENOMEM;
    pic->width = width; pic->height = height; pic->bitdepth = bitdepth; pic->device = ctx->device;
    pic->pitch[0] = align_up((size_t)width * 2, 256);
    pic->pitch[1] = pic->pitch[2] = align_up((size_t)(width >> 1) * 2, 256);
    size_t s0 = pic->pitch[0] * height, s1 = pic->pitch[1] * (height >> 1);
    cudaError_t e = cudaMalloc(&pic->base, s0 + 2 * s1);
    if (e != cudaSuccess) { free(pic); return ovgpu_set_error(ctx, OVGPU_ENOMEM, "cudaMalloc picture: %s", cudaGetErrorString(e)); }
    pic->plane[0] = (uint16_t *)pic->base;
    pic->plane[1] = (uint16_t *)((char *)pic->base + s0);
    pic->plane[2] = (uint16_t *)((char *)pic->base + s0 + s1);
    pic->owned = 1;
    *out = pic;
    return 0;
}

int ovgpu_picture_wrap(OVGPUCtx *ctx, int width, int height, int bitdepth,
                       void *const dev_planes[3], const size_t pitch_bytes[3], OVGPUPicture **out)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !out || !dev_planes || !pitch_bytes || width <= 0 || height <= 0 || (width & 7) || (height & 7) ||
        (bitdepth != 8 && bitdepth != 10)) return OVGPU_EINVAL;
    for (int c = 0; c < 3; c++)      // every kernel indexes the planes from width / height / pitch: a short pitch would be out-of-bounds writes
        if (!dev_planes[c] || ((uintptr_t)dev_planes[c] & 15) || (pitch_bytes[c] & 15) ||
            pitch_bytes[c] < (size_t)(c ? width >> 1 : width) * sizeof(uint16_t)) return OVGPU_EINVAL;
    OVGPUPicture *pic = (OVGPUPicture *)calloc(1, sizeof(*pic));
    if (!pic) return OVGPU_ENOMEM;
    pic->width = width; pic->height = height; pic->bitdepth = bitdepth; pic->device = ctx->device;
    for (int c = 0; c < 3; c++) { pic->plane[c] = (uint16_t *)dev_planes[c]; pic->pitch[c] = pitch_bytes[c]; }
    *out = pic;
    return 0;
}

void ovgpu_picture_destroy(OVGPUPicture *pic)
{
    if (!pic) return;
    if (pic->owned) { cudaSetDevice(pic->device); cudaFree(pic->base); }
    free(pic);
}

int ovgpu_picture_upload(OVGPUCtx *ctx, OVGPUPicture *pic, const void *const data[3], const size_t linesize[3])
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !pic || !data || !linesize) return OVGPU_EINVAL;
    for (int c = 0; c < 3; c++) {
        int w = c ? pic->width >> 1 : pic->width, h = c ? pic->height >> 1 : pic->height;
        OVGPU_CUDA_CHECK(ctx, cudaMemcpy2DAsync(pic->plane[c], pic->pitch[c], data[c], linesize[c],
                                                (size_t)w * 2, h, cudaMemcpyHostToDevice, ctx->stream));
    }
    return 0;
}

int ovgpu_picture_download(OVGPUCtx *ctx, const OVGPUPicture *pic, void *const data[3], const size_t linesize[3])
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !pic || !data || !linesize) return OVGPU_EINVAL;
    // OVGPU_DOWNLOAD_SM=1 (measurement knob): pinned (page-locked, hence device-mapped) destinations are written by an SM kernel
    // straight over PCIe instead of by the copy engine.  Measured on the 4K benchmark: 960 against 1100 pictures/s end to end -- the
    // kernel holds SM slots while it waits for the link -- so the DMA copy stays the default.
    {
        CopyPlanes a;
        static const int use_sm = getenv("OVGPU_DOWNLOAD_SM") != NULL;
        int mapped = use_sm;
        for (int c = 0; c < 3 && mapped; c++) {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, data[c]) != cudaSuccess) { cudaGetLastError(); mapped = 0; break; }
            if (at.type != cudaMemoryTypeHost || !at.devicePointer) { mapped = 0; break; }
            a.src[c] = (const uint8_t *)pic->plane[c]; a.dst[c] = (uint8_t *)at.devicePointer;
            a.src_pitch[c] = (uint32_t)pic->pitch[c]; a.dst_pitch[c] = (uint32_t)linesize[c];
            a.row_bytes[c] = (uint32_t)((c ? pic->width >> 1 : pic->width) * 2); a.rows[c] = (uint32_t)(c ? pic->height >> 1 : pic->height);
            mapped &= !(((uintptr_t)a.src[c] | (uintptr_t)a.dst[c] | a.src_pitch[c] | a.dst_pitch[c] | a.row_bytes[c]) & 7) && linesize[c] <= 0xffffffffu;
        }
        if (mapped) {
            ovgpu_copy_planes_kernel<<<dim3(64, 3, 1), 256, 0, ctx->stream>>>(a);
            OVGPU_CUDA_CHECK(ctx, cudaGetLastError());
            ctx->launches++;
            return 0;
        }
    }
    for (int c = 0; c < 3; c++) {
        int w = c ? pic->width >> 1 : pic->width, h = c ? pic->height >> 1 : pic->height;
        if (linesize[c] == (size_t)w * 2 && pic->pitch[c] == (size_t)w * 2)      // contiguous on both sides: one flat copy
            OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(data[c], pic->plane[c], (size_t)w * 2 * h, cudaMemcpyDeviceToHost, ctx->stream));
        else
            OVGPU_CUDA_CHECK(ctx, cudaMemcpy2DAsync(data[c], linesize[c], pic->plane[c], pic->pitch[c],
                                                    (size_t)w * 2, h, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return 0;
}

int ovgpu_picture_planes(const OVGPUPicture *pic, void *dev_planes[3], size_t pitch_bytes[3])
{
    if (!pic) return OVGPU_EINVAL;
    for (int c = 0; c < 3; c++) {
        if (dev_planes) dev_planes[c] = pic->plane[c];
        if (pitch_bytes) pitch_bytes[c] = pic->pitch[c];
    }
    return 0;
}

int ovgpu_picture_copy(OVGPUCtx *ctx, OVGPUPicture *dst, const OVGPUPicture *src)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !dst || !src || dst->width != src->width || dst->height != src->height) return OVGPU_EINVAL;
    CopyPlanes a;
    int aligned = 1;
    for (int c = 0; c < 3; c++) {
        a.src[c] = (const uint8_t *)src->plane[c]; a.dst[c] = (uint8_t *)dst->plane[c];
        a.src_pitch[c] = (uint32_t)src->pitch[c]; a.dst_pitch[c] = (uint32_t)dst->pitch[c];
        a.row_bytes[c] = (uint32_t)((c ? src->width >> 1 : src->width) * 2); a.rows[c] = (uint32_t)(c ? src->height >> 1 : src->height);
        aligned &= !(((uintptr_t)a.src[c] | (uintptr_t)a.dst[c] | a.src_pitch[c] | a.dst_pitch[c] | a.row_bytes[c]) & 7);
    }
    if (aligned) {
        ovgpu_copy_planes_kernel<<<dim3(148 * 2, 3, 1), 256, 0, ctx->stream>>>(a);
        OVGPU_CUDA_CHECK(ctx, cudaGetLastError());
        ctx->launches++;
        return 0;
    }
    for (int c = 0; c < 3; c++) {
        int w = c ? src->width >> 1 : src->width, h = c ? src->height >> 1 : src->height;
        OVGPU_CUDA_CHECK(ctx, cudaMemcpy2DAsync(dst->plane[c], dst->pitch[c], src->plane[c], src->pitch[c],
                                                (size_t)w * 2, h, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return 0;
}

} // extern "C"
