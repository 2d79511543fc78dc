// alf.cu -- C-ABI entry points of the whole-picture ALF stage (kernels in alf_kernels.h).
// Replaces alf.rcn_alf_filter_line (libovvc/rcn_alf.c:1285) for every CTU line of a picture and
// alf.rcn_alf_reconstruct_coeff_APS (rcn_alf.c:264).
#include "common.cuh"
#include "alf_host.h"
#include <stdlib.h>

struct OVGPUAlfParams {
    int width, height, log2_ctu;
    AlfGeom g;
    int nb_ctu;
    // pinned staging
    uint32_t *h_tab;
    AlfCtuDev *h_ctu;
    // device
    uint32_t *d_tab;
    AlfCtuDev *d_ctu;
    AlfChromaTab chroma;
    int wide, cc_enabled[2];
    TileGridDev *h_tiles, *d_tiles;      // ovgpu_alf_params_set_tiles
    int has_tiles;
    int device;
    StagingFence fence;
};

// ---- fixed filters and APS expansion (host; same tables as the VVC spec / rcn_alf.c:61-155) -------------
static const int8_t k_fixed_coeff[64][12] = {
    { 0, 0, 2, -3, 1, -4, 1, 7, -1, 1, -1, 5 },        { 0, 0, 0, 0, 0, -1, 0, 1, 0, 0, -1, 2 },
    { 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0 },            { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -1, 1 },
    { 2, 2, -7, -3, 0, -5, 13, 22, 12, -3, -3, 17 },   { -1, 0, 6, -8, 1, -5, 1, 23, 0, 2, -5, 10 },
    { 0, 0, -1, -1, 0, -1, 2, 1, 0, 0, -1, 4 },        { 0, 0, 3, -11, 1, 0, -1, 35, 5, 2, -9, 9 },
    { 0, 0, 8, -8, -2, -7, 4, 4, 2, 1, -1, 25 },       { 0, 0, 1, -1, 0, -3, 1, 3, -1, 1, -1, 3 },
    { 0, 0, 3, -3, 0, -6, 5, -1, 2, 1, -4, 21 },       { -7, 1, 5, 4, -3, 5, 11, 13, 12, -8, 11, 12 },
    { -5, -3, 6, -2, -3, 8, 14, 15, 2, -7, 11, 16 },   { 2, -1, -6, -5, -2, -2, 20, 14, -4, 0, -3, 25 },
    { 3, 1, -8, -4, 0, -8, 22, 5, -3, 2, -10, 29 },    { 2, 1, -7, -1, 2, -11, 23, -5, 0, 2, -10, 29 },
    { -6, -3, 8, 9, -4, 8, 9, 7, 14, -2, 8, 9 },       { 2, 1, -4, -7, 0, -8, 17, 22, 1, -1, -4, 23 },
    { 3, 0, -5, -7, 0, -7, 15, 18, -5, 0, -5, 27 },    { 2, 0, 0, -7, 1, -10, 13, 13, -4, 2, -7, 24 },
    { 3, 3, -13, 4, -2, -5, 9, 21, 25, -2, -3, 12 },   { -5, -2, 7, -3, -7, 9, 8, 9, 16, -2, 15, 12 },
    { 0, -1, 0, -7, -5, 4, 11, 11, 8, -6, 12, 21 },    { 3, -2, -3, -8, -4, -1, 16, 15, -2, -3, 3, 26 },
    { 2, 1, -5, -4, -1, -8, 16, 4, -2, 1, -7, 33 },    { 2, 1, -4, -2, 1, -10, 17, -2, 0, 2, -11, 33 },
    { 1, -2, 7, -15, -16, 10, 8, 8, 20, 11, 14, 11 },  { 2, 2, 3, -13, -13, 4, 8, 12, 2, -3, 16, 24 },
    { 1, 4, 0, -7, -8, -4, 9, 9, -2, -2, 8, 29 },      { 1, 1, 2, -4, -1, -6, 6, 3, -1, -1, -3, 30 },
    { -7, 3, 2, 10, -2, 3, 7, 11, 19, -7, 8, 10 },     { 0, -2, -5, -3, -2, 4, 20, 15, -1, -3, -1, 22 },
    { 3, -1, -8, -4, -1, -4, 22, 8, -4, 2, -8, 28 },   { 0, 3, -14, 3, 0, 1, 19, 17, 8, -3, -7, 20 },
    { 0, 2, -1, -8, 3, -6, 5, 21, 1, 1, -9, 13 },      { -4, -2, 8, 20, -2, 2, 3, 5, 21, 4, 6, 1 },
    { 2, -2, -3, -9, -4, 2, 14, 16, 3, -6, 8, 24 },    { 2, 1, 5, -16, -7, 2, 3, 11, 15, -3, 11, 22 },
    { 1, 2, 3, -11, -2, -5, 4, 8, 9, -3, -2, 26 },     { 0, -1, 10, -9, -1, -8, 2, 3, 4, 0, 0, 29 },
    { 1, 2, 0, -5, 1, -9, 9, 3, 0, 1, -7, 20 },        { -2, 8, -6, -4, 3, -9, -8, 45, 14, 2, -13, 7 },
    { 1, -1, 16, -19, -8, -4, -3, 2, 19, 0, 4, 30 },   { 1, 1, -3, 0, 2, -11, 15, -5, 1, 2, -9, 24 },
    { 0, 1, -2, 0, 1, -4, 4, 0, 0, 1, -4, 7 },         { 0, 1, 2, -5, 1, -6, 4, 10, -2, 1, -4, 10 },
    { 3, 0, -3, -6, -2, -6, 14, 8, -1, -1, -3, 31 },   { 0, 1, 0, -2, 1, -6, 5, 1, 0, 1, -5, 13 },
    { 3, 1, 9, -19, -21, 9, 7, 6, 13, 5, 15, 21 },     { 2, 4, 3, -12, -13, 1, 7, 8, 3, 0, 12, 26 },
    { 3, 1, -8, -2, 0, -6, 18, 2, -2, 3, -10, 23 },    { 1, 1, -4, -1, 1, -5, 8, 1, -1, 2, -5, 10 },
    { 0, 1, -1, 0, 0, -2, 2, 0, 0, 1, -2, 3 },         { 1, 1, -2, -7, 1, -7, 14, 18, 0, 0, -7, 21 },
    { 0, 1, 0, -2, 0, -7, 8, 1, -2, 0, -3, 24 },       { 0, 1, 1, -2, 2, -10, 10, 0, -2, 1, -7, 23 },
    { 0, 2, 2, -11, 2, -4, -3, 39, 7, 1, -10, 9 },     { 1, 0, 13, -16, -5, -6, -1, 8, 6, 0, 6, 29 },
    { 1, 3, 1, -6, -4, -7, 9, 6, -3, -2, 3, 33 },      { 4, 0, -17, -1, -1, 5, 26, 8, -2, 3, -15, 30 },
    { 0, 1, -2, 0, 2, -8, 12, -6, 1, 1, -6, 16 },      { 0, 0, 0, -1, 1, -4, 4, 0, 0, 0, -3, 11 },
    { 0, 1, 2, -8, 2, -6, 5, 15, 0, 2, -7, 9 },        { 1, -1, 12, -15, -7, -2, 3, 6, 6, -1, 7, 30 },
};
static const uint8_t k_class_to_filter[16][25] = {
    { 8, 2, 2, 2, 3, 4, 53, 9, 9, 52, 4, 4, 5, 9, 2, 8, 10, 9, 1, 3, 39, 39, 10, 9, 52 },
    { 11, 12, 13, 14, 15, 30, 11, 17, 18, 19, 16, 20, 20, 4, 53, 21, 22, 23, 14, 25, 26, 26, 27, 28, 10 },
    { 16, 12, 31, 32, 14, 16, 30, 33, 53, 34, 35, 16, 20, 4, 7, 16, 21, 36, 18, 19, 21, 26, 37, 38, 39 },
    { 35, 11, 13, 14, 43, 35, 16, 4, 34, 62, 35, 35, 30, 56, 7, 35, 21, 38, 24, 40, 16, 21, 48, 57, 39 },
    { 11, 31, 32, 43, 44, 16, 4, 17, 34, 45, 30, 20, 20, 7, 5, 21, 22, 46, 40, 47, 26, 48, 63, 58, 10 },
    { 12, 13, 50, 51, 52, 11, 17, 53, 45, 9, 30, 4, 53, 19, 0, 22, 23, 25, 43, 44, 37, 27, 28, 10, 55 },
    { 30, 33, 62, 51, 44, 20, 41, 56, 34, 45, 20, 41, 41, 56, 5, 30, 56, 38, 40, 47, 11, 37, 42, 57, 8 },
    { 35, 11, 23, 32, 14, 35, 20, 4, 17, 18, 21, 20, 20, 20, 4, 16, 21, 36, 46, 25, 41, 26, 48, 49, 58 },
    { 12, 31, 59, 59, 3, 33, 33, 59, 59, 52, 4, 33, 17, 59, 55, 22, 36, 59, 59, 60, 22, 36, 59, 25, 55 },
    { 31, 25, 15, 60, 60, 22, 17, 19, 55, 55, 20, 20, 53, 19, 55, 22, 46, 25, 43, 60, 37, 28, 10, 55, 52 },
    { 12, 31, 32, 50, 51, 11, 33, 53, 19, 45, 16, 4, 4, 53, 5, 22, 36, 18, 25, 43, 26, 27, 27, 28, 10 },
    { 5, 2, 44, 52, 3, 4, 53, 45, 9, 3, 4, 56, 5, 0, 2, 5, 10, 47, 52, 3, 63, 39, 10, 9, 52 },
    { 12, 34, 44, 44, 3, 56, 56, 62, 45, 9, 56, 56, 7, 5, 0, 22, 38, 40, 47, 52, 48, 57, 39, 10, 9 },
    { 35, 11, 23, 14, 51, 35, 20, 41, 56, 62, 16, 20, 41, 56, 7, 16, 21, 38, 24, 40, 26, 26, 42, 57, 39 },
    { 33, 34, 51, 51, 52, 41, 41, 34, 62, 0, 41, 41, 56, 7, 5, 56, 38, 38, 40, 44, 37, 42, 57, 39, 10 },
    { 16, 31, 32, 15, 60, 30, 4, 17, 19, 25, 22, 20, 4, 53, 19, 21, 22, 46, 25, 55, 26, 48, 63, 58, 55 },
};
// coefficient slot permutations realising the 4 geometric transposes (rcn_alf.c:149-155)
static const uint8_t k_transpose[4][13] = {
    { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12 }, { 9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6, 12 },
    { 0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11, 12 }, { 9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6, 12 },
};

extern "C" {

int ovgpu_alf_reconstruct_coeff_aps(OVGPUAlfTables *t, const OVGPUAlfAps *luma_aps, int nb_luma_aps,
                                    const OVGPUAlfAps *chroma_aps, int bitdepth)
{
    if (!t || nb_luma_aps < 0 || nb_luma_aps > 8 || (nb_luma_aps && !luma_aps)) return OVGPU_EINVAL;
    const int16_t clip_lut[4] = { (int16_t)(1 << bitdepth), (int16_t)(1 << (bitdepth - 3)),
                                  (int16_t)(1 << (bitdepth - 5)), (int16_t)(1 << (bitdepth - 7)) };
    for (int s = 0; s < 16; s++)
        for (int tr = 0; tr < 4; tr++)
            for (int c = 0; c < 25; c++) {
                int16_t *co = &t->luma_coeff[s][(tr * 25 + c) * 13], *cl = &t->luma_clip[s][(tr * 25 + c) * 13];
                for (int k = 0; k < 12; k++) {
                    co[k] = k_fixed_coeff[k_class_to_filter[s][c]][k_transpose[tr][k]];
                    cl[k] = clip_lut[0];
                }
                co[12] = 128; cl[12] = clip_lut[0];
            }
    for (int a = 0; a < nb_luma_aps; a++) {
        const OVGPUAlfAps *aps = &luma_aps[a];
        for (int c = 0; c < 25; c++) {
            int f = aps->luma_coeff_delta_idx[c];
            int16_t co[13], cl[13];
            for (int k = 0; k < 12; k++) {
                co[k] = aps->luma_coeff[f][k];
                cl[k] = clip_lut[aps->luma_clip_flag ? (aps->luma_clip_idx[f][k] & 3) : 0];
            }
            co[12] = 128; cl[12] = clip_lut[0];
            for (int tr = 0; tr < 4; tr++)
                for (int k = 0; k < 13; k++) {
                    t->luma_coeff[16 + a][(tr * 25 + c) * 13 + k] = co[k_transpose[tr][k]];
                    t->luma_clip [16 + a][(tr * 25 + c) * 13 + k] = cl[k_transpose[tr][k]];
                }
        }
    }
    if (chroma_aps)
        for (int alt = 0; alt < chroma_aps->num_chroma_alts && alt < 8; alt++) {
            for (int k = 0; k < 6; k++) {
                t->chroma_coeff[alt][k] = chroma_aps->chroma_coeff[alt][k];
                t->chroma_clip [alt][k] = clip_lut[chroma_aps->chroma_clip_flag ? (chroma_aps->chroma_clip_idx[alt][k] & 3) : 0];
            }
            t->chroma_coeff[alt][6] = 128; t->chroma_clip[alt][6] = clip_lut[0];
        }
    return 0;
}

int ovgpu_alf_params_create(OVGPUCtx *ctx, int width, int height, int log2_ctu_s, OVGPUAlfParams **out)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !out || width <= 0 || height <= 0 || (width & 7) || (height & 7) || log2_ctu_s < 5 || log2_ctu_s > 7)
        return OVGPU_EINVAL;
    OVGPUAlfParams *p = (OVGPUAlfParams *)calloc(1, sizeof(*p));
    if (!p) return OVGPU_ENOMEM;
    p->width = width; p->height = height; p->log2_ctu = log2_ctu_s; p->device = ctx->device;
    p->g = alf_geom(width, height, log2_ctu_s);
    p->nb_ctu = p->g.nb_ctu_w * p->g.nb_ctu_h;
    size_t tab_bytes = (size_t)OVGPU_ALF_NUM_SETS * ALF_TAB_SET_WORDS * 4;
    if (cudaMallocHost(&p->h_tab, tab_bytes) != cudaSuccess ||
        cudaMallocHost(&p->h_ctu, sizeof(AlfCtuDev) * p->nb_ctu) != cudaSuccess ||
        cudaMalloc(&p->d_tab, tab_bytes) != cudaSuccess ||
        cudaMallocHost(&p->h_tiles, sizeof(TileGridDev)) != cudaSuccess || cudaMalloc(&p->d_tiles, sizeof(TileGridDev)) != cudaSuccess ||
        cudaMalloc(&p->d_ctu, sizeof(AlfCtuDev) * p->nb_ctu) != cudaSuccess) {
        ovgpu_alf_params_destroy(p);
        return ovgpu_set_error(ctx, OVGPU_ENOMEM, "ALF parameter buffers: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = p;
    return 0;
}

void ovgpu_alf_params_destroy(OVGPUAlfParams *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->h_tab) cudaFreeHost(p->h_tab);
    if (p->h_ctu) cudaFreeHost(p->h_ctu);
    if (p->d_tab) cudaFree(p->d_tab);
    if (p->d_ctu) cudaFree(p->d_ctu);
    if (p->h_tiles) cudaFreeHost(p->h_tiles);
    if (p->d_tiles) cudaFree(p->d_tiles);
    fence_destroy(&p->fence);
    free(p);
}

int ovgpu_alf_params_set_tiles(OVGPUCtx *ctx, OVGPUAlfParams *p, const OVGPUTileGrid *tiles)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !p) return OVGPU_EINVAL;
    if (!tiles || tiles->nb_cols * tiles->nb_rows <= 1) { p->has_tiles = 0; return 0; }
    if (ovgpu_tile_grid_check(tiles, p->g.nb_ctu_w, p->g.nb_ctu_h)) return ovgpu_set_error(ctx, OVGPU_EINVAL, "tile grid does not cover the picture's %d x %d CTUs", p->g.nb_ctu_w, p->g.nb_ctu_h);
    OVGPU_CUDA_CHECK(ctx, fence_wait(&p->fence));
    memcpy(p->h_tiles, tiles, sizeof(TileGridDev));
    OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(p->d_tiles, p->h_tiles, sizeof(TileGridDev), cudaMemcpyHostToDevice, ctx->stream));
    OVGPU_CUDA_CHECK(ctx, fence_mark(&p->fence, ctx->stream));
    p->has_tiles = 1;
    return 0;
}

int ovgpu_alf_params_upload(OVGPUCtx *ctx, OVGPUAlfParams *p, const OVGPUAlfTables *tab,
                            const OVGPUAlfCtu *ctu, const uint8_t *const cc_idx[2],
                            int cc_cb_enabled, int cc_cr_enabled)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !p || !tab || !ctu) return OVGPU_EINVAL;
    // the pinned staging buffers are reused: wait for the previous H2D copy out of them
    OVGPU_CUDA_CHECK(ctx, fence_wait(&p->fence));
    p->wide = alf_pack_luma_tab(p->h_tab, tab);
    alf_pack_chroma_tab(&p->chroma, tab);
    alf_pack_ctu(p->h_ctu, p->nb_ctu, ctu, cc_idx);
    p->cc_enabled[0] = cc_cb_enabled && cc_idx && cc_idx[0];
    p->cc_enabled[1] = cc_cr_enabled && cc_idx && cc_idx[1];
    OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(p->d_tab, p->h_tab, (size_t)OVGPU_ALF_NUM_SETS * ALF_TAB_SET_WORDS * 4,
                                          cudaMemcpyHostToDevice, ctx->stream));
    OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(p->d_ctu, p->h_ctu, sizeof(AlfCtuDev) * p->nb_ctu,
                                          cudaMemcpyHostToDevice, ctx->stream));
    OVGPU_CUDA_CHECK(ctx, fence_mark(&p->fence, ctx->stream));
    return 0;
}

int ovgpu_alf_picture(OVGPUCtx *ctx, OVGPUPicture *dst, const OVGPUPicture *src, const OVGPUAlfParams *p)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !dst || !src || !p) return OVGPU_EINVAL;
    if (dst == src || dst->plane[0] == src->plane[0]) return ovgpu_set_error(ctx, OVGPU_EINVAL, "ALF needs dst != src");
    if (src->width != p->width || src->height != p->height || dst->width != p->width || dst->height != p->height)
        return ovgpu_set_error(ctx, OVGPU_EINVAL, "ALF picture/params size mismatch");
    const AlfGeom &g = p->g;
    PlaneView sl = plane_view(src, 0), dl = plane_view(dst, 0);
    {
        StageScope scope(ctx, OVGPU_STAGE_ALF_LUMA, 1);
        dim3 grid(g.nb_ctu_w * g.l_tx, g.nb_ctu_h * g.l_ty, 1);
        if (p->wide) {
            typedef AlfLumaK<true> K;
            K::Args a = { sl.p, dl.p, sl.pitch, dl.pitch, p->width, p->height, p->log2_ctu, src->bitdepth,
                          g.l_tw, g.l_th, g.l_tx, g.l_ty, g.nb_ctu_w, p->d_tab, p->d_ctu, p->has_tiles ? p->d_tiles : NULL };
            OVGPU_CUDA_CHECK(ctx, ovd_launch<K>(a, grid, K::SMEM_BYTES, ctx->stream));
        } else {
            typedef AlfLumaK<false> K;
            K::Args a = { sl.p, dl.p, sl.pitch, dl.pitch, p->width, p->height, p->log2_ctu, src->bitdepth,
                          g.l_tw, g.l_th, g.l_tx, g.l_ty, g.nb_ctu_w, p->d_tab, p->d_ctu, p->has_tiles ? p->d_tiles : NULL };
            OVGPU_CUDA_CHECK(ctx, ovd_launch<K>(a, grid, K::SMEM_BYTES, ctx->stream));
        }
    }
    {
        StageScope scope(ctx, OVGPU_STAGE_ALF_CHROMA, 1);
        typedef AlfChromaK K;
        PlaneView s1 = plane_view(src, 1), s2 = plane_view(src, 2), d1 = plane_view(dst, 1), d2 = plane_view(dst, 2);
        if (s1.pitch != s2.pitch || d1.pitch != d2.pitch)
            return ovgpu_set_error(ctx, OVGPU_EINVAL, "Cb/Cr pitches must match");
        K::Args a;
        a.src[0] = s1.p; a.src[1] = s2.p; a.dst[0] = d1.p; a.dst[1] = d2.p;
        a.luma = sl.p;
        a.src_pitch = s1.pitch; a.dst_pitch = d1.pitch; a.luma_pitch = sl.pitch;
        a.pic_w = p->width; a.pic_h = p->height; a.log2_ctu = p->log2_ctu; a.bitdepth = src->bitdepth;
        a.tw = g.c_tw; a.th = g.c_th; a.tiles_x_per_ctu = g.c_tx; a.tiles_y_per_ctu = g.c_ty; a.nb_ctu_w = g.nb_ctu_w;
        a.cc_enabled[0] = p->cc_enabled[0]; a.cc_enabled[1] = p->cc_enabled[1];
        a.ctu = p->d_ctu;
        a.tab = p->chroma;
        a.tiles = p->has_tiles ? p->d_tiles : NULL;
        dim3 grid(g.nb_ctu_w * g.c_tx, g.nb_ctu_h * g.c_ty, 2);
        OVGPU_CUDA_CHECK(ctx, ovd_launch<K>(a, grid, K::SMEM_BYTES, ctx->stream));
    }
    return 0;
}

int ovgpu_alf_frame_host(OVGPUCtx *ctx, void *const dst_data[3], const size_t dst_linesize[3],
                         const void *const src_data[3], const size_t src_linesize[3],
                         int width, int height, int log2_ctu_s, int bitdepth,
                         const OVGPUAlfTables *tab, const OVGPUAlfCtu *ctu, const uint8_t *const cc_idx[2],
                         int cc_cb_enabled, int cc_cr_enabled)
{
    OVGPU_DEVICE_GUARD(ctx);
    OVGPUPicture *src = NULL, *dst = NULL;
    OVGPUAlfParams *p = NULL;
    int ret = ovgpu_picture_create(ctx, width, height, bitdepth, &src);
    if (!ret) ret = ovgpu_picture_create(ctx, width, height, bitdepth, &dst);
    if (!ret) ret = ovgpu_alf_params_create(ctx, width, height, log2_ctu_s, &p);
    if (!ret) ret = ovgpu_picture_upload(ctx, src, src_data, src_linesize);
    if (!ret) ret = ovgpu_alf_params_upload(ctx, p, tab, ctu, cc_idx, cc_cb_enabled, cc_cr_enabled);
    if (!ret) ret = ovgpu_alf_picture(ctx, dst, src, p);
    if (!ret) ret = ovgpu_picture_download(ctx, dst, dst_data, dst_linesize);
    if (!ret) ret = ovgpu_ctx_sync(ctx);
    else cudaStreamSynchronize(ctx->stream);
    ovgpu_alf_params_destroy(p);
    ovgpu_picture_destroy(dst);
    ovgpu_picture_destroy(src);
    return ret;
}

} // extern "C"
