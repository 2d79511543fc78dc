// sao.cu -- C-ABI entry points of the whole-picture SAO stage (kernel in sao_kernels.h).
// Replaces sao.rcn_sao_first_pix_rows + sao.rcn_sao_filter_line (libovvc/rcn_sao.c:190-293).
#include "common.cuh"
#include "sao_kernels.h"
#include <stdlib.h>

struct OVGPUSaoParams {
    int width, height, log2_ctu, nb_ctu_w, nb_ctu_h, nb_ctu;
    SaoCtuDev *h_ctu, *d_ctu;
    TileGridDev *h_tiles, *d_tiles;      // ovgpu_sao_params_set_tiles
    int has_tiles;
    int device;
    StagingFence fence;
};

static_assert(sizeof(SaoCtuDev) == sizeof(OVGPUSaoCtu), "SAO CTU parameter layout");
static_assert(sizeof(TileGridDev) == sizeof(OVGPUTileGrid), "tile grid layout");

extern "C" {

int ovgpu_sao_params_create(OVGPUCtx *ctx, int width, int height, int log2_ctu_s, OVGPUSaoParams **out)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !out || width <= 0 || height <= 0 || (width & 7) || (height & 7) || log2_ctu_s < 5 || log2_ctu_s > 7)
        return OVGPU_EINVAL;
    OVGPUSaoParams *p = (OVGPUSaoParams *)calloc(1, sizeof(*p));
    if (!p) return OVGPU_ENOMEM;
    int ctu_s = 1 << log2_ctu_s;
    p->width = width; p->height = height; p->log2_ctu = log2_ctu_s; p->device = ctx->device;
    p->nb_ctu_w = (width + ctu_s - 1) >> log2_ctu_s;
    p->nb_ctu_h = (height + ctu_s - 1) >> log2_ctu_s;
    p->nb_ctu = p->nb_ctu_w * p->nb_ctu_h;
    if (cudaMallocHost(&p->h_ctu, sizeof(SaoCtuDev) * p->nb_ctu) != cudaSuccess ||
        cudaMallocHost(&p->h_tiles, sizeof(TileGridDev)) != cudaSuccess || cudaMalloc(&p->d_tiles, sizeof(TileGridDev)) != cudaSuccess ||
        cudaMalloc(&p->d_ctu, sizeof(SaoCtuDev) * p->nb_ctu) != cudaSuccess) {
        ovgpu_sao_params_destroy(p);
        return ovgpu_set_error(ctx, OVGPU_ENOMEM, "SAO parameter buffers: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = p;
    return 0;
}

void ovgpu_sao_params_destroy(OVGPUSaoParams *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->h_ctu) cudaFreeHost(p->h_ctu);
    if (p->d_ctu) cudaFree(p->d_ctu);
    if (p->h_tiles) cudaFreeHost(p->h_tiles);
    if (p->d_tiles) cudaFree(p->d_tiles);
    fence_destroy(&p->fence);
    free(p);
}

int ovgpu_sao_params_upload(OVGPUCtx *ctx, OVGPUSaoParams *p, const OVGPUSaoCtu *ctu)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !p || !ctu) return OVGPU_EINVAL;
    // the kernel keeps the five offsets of a CTU as packed bytes (|offset| <= 31 at 10 bit, vcl_sao.c:69)
    for (int i = 0; i < p->nb_ctu; i++)
        for (int c = 0; c < 3; c++)
            for (int k = 0; k < 5; k++)
                if (ctu[i].offset_val[c][k] < -128 || ctu[i].offset_val[c][k] > 127)
                    return ovgpu_set_error(ctx, OVGPU_EINVAL, "SAO offset of CTU %d out of range", i);
    OVGPU_CUDA_CHECK(ctx, fence_wait(&p->fence));
    memcpy(p->h_ctu, ctu, sizeof(SaoCtuDev) * p->nb_ctu);
    OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(p->d_ctu, p->h_ctu, sizeof(SaoCtuDev) * p->nb_ctu,
                                          cudaMemcpyHostToDevice, ctx->stream));
    OVGPU_CUDA_CHECK(ctx, fence_mark(&p->fence, ctx->stream));
    return 0;
}

// boundaries must start at 0, increase and end at the picture's size in CTUs
int ovgpu_tile_grid_check(const OVGPUTileGrid *t, int nb_ctu_w, int nb_ctu_h)
{
    if (!t || t->nb_cols < 1 || t->nb_cols > 20 || t->nb_rows < 1 || t->nb_rows > 20) return -1;
    if (t->col_bd[0] != 0 || t->row_bd[0] != 0 || t->col_bd[t->nb_cols] != nb_ctu_w || t->row_bd[t->nb_rows] != nb_ctu_h) return -1;
    for (int i = 0; i < t->nb_cols; i++) if (t->col_bd[i + 1] <= t->col_bd[i]) return -1;
    for (int i = 0; i < t->nb_rows; i++) if (t->row_bd[i + 1] <= t->row_bd[i]) return -1;
    return 0;
}

int ovgpu_sao_params_set_tiles(OVGPUCtx *ctx, OVGPUSaoParams *p, const OVGPUTileGrid *tiles)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !p) return OVGPU_EINVAL;
    if (!tiles || tiles->nb_cols * tiles->nb_rows <= 1) { p->has_tiles = 0; return 0; }
    if (ovgpu_tile_grid_check(tiles, p->nb_ctu_w, p->nb_ctu_h)) return ovgpu_set_error(ctx, OVGPU_EINVAL, "tile grid does not cover the picture's %d x %d CTUs", p->nb_ctu_w, p->nb_ctu_h);
    OVGPU_CUDA_CHECK(ctx, fence_wait(&p->fence));
    memcpy(p->h_tiles, tiles, sizeof(TileGridDev));
    OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(p->d_tiles, p->h_tiles, sizeof(TileGridDev), cudaMemcpyHostToDevice, ctx->stream));
    OVGPU_CUDA_CHECK(ctx, fence_mark(&p->fence, ctx->stream));
    p->has_tiles = 1;
    return 0;
}

int ovgpu_sao_picture(OVGPUCtx *ctx, OVGPUPicture *dst, const OVGPUPicture *src, const OVGPUSaoParams *p)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !dst || !src || !p) return OVGPU_EINVAL;
    if (dst == src || dst->plane[0] == src->plane[0]) return ovgpu_set_error(ctx, OVGPU_EINVAL, "SAO needs dst != src");
    if (src->width != p->width || src->height != p->height || dst->width != p->width || dst->height != p->height)
        return ovgpu_set_error(ctx, OVGPU_EINVAL, "SAO picture/params size mismatch");
    StageScope scope(ctx, OVGPU_STAGE_SAO, 1);
    SaoK::Args a;
    for (int c = 0; c < 3; c++) {
        PlaneView s = plane_view(src, c), d = plane_view(dst, c);
        a.src[c] = s.p; a.dst[c] = d.p; a.src_pitch[c] = s.pitch; a.dst_pitch[c] = d.pitch;
    }
    a.pic_w = p->width; a.pic_h = p->height; a.log2_ctu = p->log2_ctu; a.bitdepth = src->bitdepth;
    a.nb_ctu_w = p->nb_ctu_w; a.nb_ctu_h = p->nb_ctu_h;
    a.ctu = p->d_ctu;
    a.tiles = p->has_tiles ? p->d_tiles : NULL;
    dim3 grid(OVGPU_CEIL_DIV(p->width, SaoK::BX * 8), OVGPU_CEIL_DIV(p->height, SaoK::BY), 3);
    OVGPU_CUDA_CHECK(ctx, ovd_launch<SaoK>(a, grid, 0, ctx->stream));
    return 0;
}

int ovgpu_sao_frame_host(OVGPUCtx *ctx, void *const dst_data[3], const size_t dst_linesize[3],
                         const void *const src_data[3], const size_t src_linesize[3],
                         int width, int height, int log2_ctu_s, int bitdepth, const OVGPUSaoCtu *ctu)
{
    OVGPU_DEVICE_GUARD(ctx);
    OVGPUPicture *src = NULL, *dst = NULL;
    OVGPUSaoParams *p = NULL;
    int ret = ovgpu_picture_create(ctx, width, height, bitdepth, &src);
    if (!ret) ret = ovgpu_picture_create(ctx, width, height, bitdepth, &dst);
    if (!ret) ret = ovgpu_sao_params_create(ctx, width, height, log2_ctu_s, &p);
    if (!ret) ret = ovgpu_picture_upload(ctx, src, src_data, src_linesize);
    if (!ret) ret = ovgpu_sao_params_upload(ctx, p, ctu);
    if (!ret) ret = ovgpu_sao_picture(ctx, dst, src, p);
    if (!ret) ret = ovgpu_picture_download(ctx, dst, dst_data, dst_linesize);
    if (!ret) ret = ovgpu_ctx_sync(ctx);
    else cudaStreamSynchronize(ctx->stream);
    ovgpu_sao_params_destroy(p);
    ovgpu_picture_destroy(dst);
    ovgpu_picture_destroy(src);
    return ret;
}

} // extern "C"
