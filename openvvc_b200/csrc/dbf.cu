// dbf.cu -- C-ABI entry points of the deblocking filter (kernels in dbf_kernels.h).
// Replaces df.rcn_dbf_ctu / df.rcn_dbf_truncated_ctu (libovvc/rcn_df.c:2170-2228) for all CTUs of a picture.
#include "common.cuh"
#include "dbf_kernels.h"
#include <stdlib.h>

struct OVGPUDbfParams {
    int width, height, log2_ctu, device;
    int n_u, w_u, h_u, ncw, nch;
    size_t ver_words, hor_words, qp_bytes, total;    // per map / per component
    unsigned char *h_buf, *d_buf;
    int beta_offset, tc_offset, loaded;
    StagingFence fence;
};

static size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

extern "C" {

int ovgpu_dbf_params_create(OVGPUCtx *ctx, int width, int height, int log2_ctu_s, OVGPUDbfParams **out)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !out || width <= 0 || height <= 0 || (width & 7) || (height & 7) || log2_ctu_s < 5 || log2_ctu_s > 7)
        return OVGPU_EINVAL;
    OVGPUDbfParams *p = (OVGPUDbfParams *)calloc(1, sizeof(*p));
    if (!p) return OVGPU_ENOMEM;
    p->width = width; p->height = height; p->log2_ctu = log2_ctu_s; p->device = ctx->device;
    p->n_u = (1 << log2_ctu_s) >> 2; p->w_u = width >> 2; p->h_u = height >> 2;
    p->ncw = OVGPU_CEIL_DIV(p->w_u, p->n_u); p->nch = OVGPU_CEIL_DIV(p->h_u, p->n_u);
    p->ver_words = (size_t)p->nch * (p->w_u + 1); p->hor_words = (size_t)p->ncw * (p->h_u + 1);
    p->qp_bytes = (size_t)p->w_u * p->h_u;
    p->total = DBF_NB_MAPS * (align16(p->ver_words * 4) + align16(p->hor_words * 4)) + 3 * align16(p->qp_bytes);
    if (cudaMallocHost(&p->h_buf, p->total) != cudaSuccess || cudaMalloc(&p->d_buf, p->total) != cudaSuccess) {
        ovgpu_dbf_params_destroy(p);
        return ovgpu_set_error(ctx, OVGPU_ENOMEM, "DBF map buffers: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = p;
    return 0;
}

void ovgpu_dbf_params_destroy(OVGPUDbfParams *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->h_buf) cudaFreeHost(p->h_buf);
    if (p->d_buf) cudaFree(p->d_buf);
    fence_destroy(&p->fence);
    free(p);
}

int ovgpu_dbf_params_staging(OVGPUCtx *ctx, OVGPUDbfParams *p, OVGPUDbfMaps *maps)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !p || !maps) return OVGPU_EINVAL;
    OVGPU_CUDA_CHECK(ctx, fence_wait(&p->fence));
    maps->width = p->width; maps->height = p->height; maps->log2_ctu_s = p->log2_ctu;
    size_t off = 0;
    for (int k = 0; k < DBF_NB_MAPS; k++) { maps->ver[k] = (const uint32_t *)(p->h_buf + off); off += align16(p->ver_words * 4); }
    for (int k = 0; k < DBF_NB_MAPS; k++) { maps->hor[k] = (const uint32_t *)(p->h_buf + off); off += align16(p->hor_words * 4); }
    for (int c = 0; c < 3; c++) { maps->qp[c] = p->h_buf + off; off += align16(p->qp_bytes); }
    return 0;
}

int ovgpu_dbf_params_commit(OVGPUCtx *ctx, OVGPUDbfParams *p, int beta_offset, int tc_offset)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !p) return OVGPU_EINVAL;
    OVGPU_CUDA_CHECK(ctx, fence_wait(&p->fence));
    p->beta_offset = beta_offset; p->tc_offset = tc_offset;
    OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(p->d_buf, p->h_buf, p->total, cudaMemcpyHostToDevice, ctx->stream));
    OVGPU_CUDA_CHECK(ctx, fence_mark(&p->fence, ctx->stream));
    p->loaded = 1;
    return 0;
}

int ovgpu_dbf_params_upload(OVGPUCtx *ctx, OVGPUDbfParams *p, const OVGPUDbfMaps *maps)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !p || !maps) return OVGPU_EINVAL;
    if (maps->width != p->width || maps->height != p->height || maps->log2_ctu_s != p->log2_ctu)
        return ovgpu_set_error(ctx, OVGPU_EINVAL, "DBF maps geometry differs from the parameter object");
    for (int k = 0; k < DBF_NB_MAPS; k++) if (!maps->ver[k] || !maps->hor[k]) return OVGPU_EINVAL;
    for (int c = 0; c < 3; c++) if (!maps->qp[c]) return OVGPU_EINVAL;
    OVGPUDbfMaps st;
    int ret = ovgpu_dbf_params_staging(ctx, p, &st);
    if (ret) return ret;
    for (int k = 0; k < DBF_NB_MAPS; k++) {
        memcpy((void *)st.ver[k], maps->ver[k], p->ver_words * 4);
        memcpy((void *)st.hor[k], maps->hor[k], p->hor_words * 4);
    }
    for (int c = 0; c < 3; c++) memcpy((void *)st.qp[c], maps->qp[c], p->qp_bytes);
    return ovgpu_dbf_params_commit(ctx, p, maps->beta_offset, maps->tc_offset);
}

int ovgpu_dbf_picture(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPUDbfParams *p)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !pic || !p) return OVGPU_EINVAL;
    if (!p->loaded || pic->width != p->width || pic->height != p->height)
        return ovgpu_set_error(ctx, OVGPU_EINVAL, "DBF parameters not uploaded or of another geometry");
    DbfK<0>::Args a;
    memset(&a, 0, sizeof(a));
    for (int c = 0; c < 3; c++) { PlaneView v = plane_view(pic, c); a.plane[c] = v.p; a.pitch[c] = v.pitch; }
    a.bitdepth = pic->bitdepth;
    size_t off = 0;
    for (int k = 0; k < DBF_NB_MAPS; k++) { a.m.ver[k] = (const uint32_t *)(p->d_buf + off); off += align16(p->ver_words * 4); }
    for (int k = 0; k < DBF_NB_MAPS; k++) { a.m.hor[k] = (const uint32_t *)(p->d_buf + off); off += align16(p->hor_words * 4); }
    for (int c = 0; c < 3; c++) { a.m.qp[c] = p->d_buf + off; off += align16(p->qp_bytes); }
    a.m.n_u = p->n_u; a.m.w_u = p->w_u; a.m.h_u = p->h_u; a.m.beta_offset = p->beta_offset; a.m.tc_offset = p->tc_offset;
    dim3 grid(OVGPU_CEIL_DIV(p->w_u, DbfK<0>::BX), OVGPU_CEIL_DIV(p->h_u, DbfK<0>::BY), 3);
    {
        StageScope scope(ctx, OVGPU_STAGE_DBF_V, 1);
        OVGPU_CUDA_CHECK(ctx, ovd_launch<DbfK<0> >(a, grid, 0, ctx->stream));
    }
    {
        StageScope scope(ctx, OVGPU_STAGE_DBF_H, 1);
        DbfK<1>::Args b;
        memcpy(&b, &a, sizeof(b));
        OVGPU_CUDA_CHECK(ctx, ovd_launch<DbfK<1> >(b, grid, 0, ctx->stream));
    }
    return 0;
}

} // extern "C"
