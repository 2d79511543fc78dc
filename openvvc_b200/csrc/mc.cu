// mc.cu -- C-ABI entry points of motion compensation and LMCS (kernels in mc_kernels.h / mc_host.h).
// Replaces rcn_mcp, rcn_mcp_b, rcn_mcp_b_l, rcn_mcp_b_c (libovvc/rcn_inter.c:2750-2966) for all PUs of a picture,
// rcn_init_lmcs and lmcs_reshape_forward/backward (libovvc/rcn_lmcs.c).
#include "common.cuh"
#include "mc_host.h"
#include <stdlib.h>

static_assert(sizeof(McCmdDev) == sizeof(OVGPUMcCmd) && sizeof(McCmdDev) == 32, "MC command layout");

struct OVGPUMcBatch {
    int max_cmds, max_items, nb_cmds, nb_items, device;
    int nb_bdof, nb_prof_items, nb_prof;      // BDOF / PROF work items (appended to the item list), PROF table entries
    int nb_dmvr, nb_small_c, nb_big;
    DmvrMvDev *d_dmvr, *h_dmvr;               // refined MVs, one per DMVR item (device result + pinned read-back)
    McCmdDev *h_cmds, *d_cmds;
    McItemDev *h_items, *d_items;             // six regions of max_items entries: plain | BDOF | PROF | DMVR | small chroma | big tiles
    int max_x, max_y;                         // furthest sample any command writes (checked against the picture at run time)
    uint32_t used_refs[2];                    // reference indices the commands name, per list
    ProfInfoDev *h_prof, *d_prof;             // max_cmds / 16 + 1 entries (one per affine CU of >= 16 sub-blocks)
    int max_prof;
    StagingFence fence;
};
static_assert(sizeof(ProfInfoDev) == sizeof(OVGPUProfInfo), "PROF table layout");
static_assert(sizeof(DmvrMvDev) == sizeof(OVGPUDmvrMv), "DMVR result layout");

struct OVGPULmcs {
    OVGPULmcsLuts *h_luts, *d_luts;
    int device;
    int has_data, min_idx, max_idx_p1, crs_offset;      // ovgpu_lmcs_set_data (rcn_init_lmcs, rcn_lmcs.c:345-361)
    StagingFence fence;
};

extern "C" {

int ovgpu_lmcs_build_luts(OVGPULmcsLuts *luts, const OVGPULmcsData *data, int bitdepth)
{
    if (!luts || !data || (bitdepth != 10 && bitdepth != 8)) return OVGPU_EINVAL;
    if (data->lmcs_min_bin_idx > 15 || data->lmcs_delta_max_bin_idx > 15) return OVGPU_EINVAL;
    lmcs_build_luts_host(luts, data, bitdepth);
    return 0;
}

int ovgpu_lmcs_create(OVGPUCtx *ctx, OVGPULmcs **out)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !out) return OVGPU_EINVAL;
    OVGPULmcs *l = (OVGPULmcs *)calloc(1, sizeof(*l));
    if (!l) return OVGPU_ENOMEM;
    l->device = ctx->device;
    if (cudaMallocHost(&l->h_luts, sizeof(OVGPULmcsLuts)) != cudaSuccess ||
        cudaMalloc(&l->d_luts, sizeof(OVGPULmcsLuts)) != cudaSuccess) {
        ovgpu_lmcs_destroy(l);
        return ovgpu_set_error(ctx, OVGPU_ENOMEM, "LMCS buffers");
    }
    *out = l;
    return 0;
}

void ovgpu_lmcs_destroy(OVGPULmcs *l)
{
    if (!l) return;
    cudaSetDevice(l->device);
    if (l->h_luts) cudaFreeHost(l->h_luts);
    if (l->d_luts) cudaFree(l->d_luts);
    fence_destroy(&l->fence);
    free(l);
}

int ovgpu_lmcs_upload(OVGPUCtx *ctx, OVGPULmcs *l, const OVGPULmcsLuts *luts)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !l || !luts) return OVGPU_EINVAL;
    OVGPU_CUDA_CHECK(ctx, fence_wait(&l->fence));
    memcpy(l->h_luts, luts, sizeof(*luts));
    OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(l->d_luts, l->h_luts, sizeof(*luts), cudaMemcpyHostToDevice, ctx->stream));
    OVGPU_CUDA_CHECK(ctx, fence_mark(&l->fence, ctx->stream));
    return 0;
}

int ovgpu_lmcs_set_data(OVGPULmcs *l, const OVGPULmcsData *d)
{
    if (!l || !d || d->lmcs_min_bin_idx > 15 || d->lmcs_delta_max_bin_idx > 15) return OVGPU_EINVAL;
    l->min_idx = d->lmcs_min_bin_idx;
    l->max_idx_p1 = 16 - d->lmcs_delta_max_bin_idx;
    l->crs_offset = d->lmcs_delta_sign_crs_flag ? -(int)d->lmcs_delta_abs_crs : (int)d->lmcs_delta_abs_crs;
    l->has_data = 1;
    return 0;
}

int ovgpu_lmcs_reshape_picture(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPULmcs *l, int forward)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !pic || !l) return OVGPU_EINVAL;
    StageScope scope(ctx, OVGPU_STAGE_LMCS, 1);
    PlaneView v = plane_view(pic, 0);
    LmcsReshapeK::Args a = { v.p, v.pitch, v.w, v.h, pic->bitdepth, forward ? l->d_luts->fwd_lut : l->d_luts->bwd_lut };
    dim3 grid(OVGPU_CEIL_DIV(v.w, 256), OVGPU_CEIL_DIV(v.h, 8), 1);
    OVGPU_CUDA_CHECK(ctx, ovd_launch<LmcsReshapeK>(a, grid, 0, ctx->stream));
    return 0;
}

int ovgpu_mc_batch_create(OVGPUCtx *ctx, int max_cmds, int max_items, OVGPUMcBatch **out)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !out || max_cmds <= 0 || max_items < max_cmds) return OVGPU_EINVAL;
    OVGPUMcBatch *b = (OVGPUMcBatch *)calloc(1, sizeof(*b));
    if (!b) return OVGPU_ENOMEM;
    b->max_cmds = max_cmds; b->max_items = max_items; b->device = ctx->device;
    b->max_prof = max_cmds / 16 + 1;
    if (cudaMallocHost(&b->h_prof, sizeof(ProfInfoDev) * b->max_prof) != cudaSuccess ||
        cudaMalloc(&b->d_prof, sizeof(ProfInfoDev) * b->max_prof) != cudaSuccess ||
        cudaMallocHost(&b->h_cmds, sizeof(McCmdDev) * max_cmds) != cudaSuccess ||
        cudaMallocHost(&b->h_items, sizeof(McItemDev) * 6 * max_items) != cudaSuccess ||
        cudaMallocHost(&b->h_dmvr, sizeof(DmvrMvDev) * max_items) != cudaSuccess ||
        cudaMalloc(&b->d_dmvr, sizeof(DmvrMvDev) * max_items) != cudaSuccess ||
        cudaMalloc(&b->d_cmds, sizeof(McCmdDev) * max_cmds) != cudaSuccess ||
        cudaMalloc(&b->d_items, sizeof(McItemDev) * 6 * max_items) != cudaSuccess) {
        ovgpu_mc_batch_destroy(b);
        return ovgpu_set_error(ctx, OVGPU_ENOMEM, "MC batch buffers: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = b;
    return 0;
}

void ovgpu_mc_batch_destroy(OVGPUMcBatch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->h_cmds) cudaFreeHost(b->h_cmds);
    if (b->h_items) cudaFreeHost(b->h_items);
    if (b->d_cmds) cudaFree(b->d_cmds);
    if (b->d_items) cudaFree(b->d_items);
    if (b->h_dmvr) cudaFreeHost(b->h_dmvr);
    if (b->d_dmvr) cudaFree(b->d_dmvr);
    if (b->h_prof) cudaFreeHost(b->h_prof);
    if (b->d_prof) cudaFree(b->d_prof);
    fence_destroy(&b->fence);
    free(b);
}

int ovgpu_mc_batch_upload(OVGPUCtx *ctx, OVGPUMcBatch *b, const OVGPUMcCmd *cmds, int nb_cmds)
{
    OVGPU_DEVICE_GUARD(ctx);
    return ovgpu_mc_batch_upload2(ctx, b, cmds, nb_cmds, NULL, 0);
}

int ovgpu_mc_batch_staging(OVGPUCtx *ctx, OVGPUMcBatch *b, OVGPUMcCmd **cmds, OVGPUProfInfo **prof)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !b || !cmds) return OVGPU_EINVAL;
    OVGPU_CUDA_CHECK(ctx, fence_wait(&b->fence));           // previous H2D copy out of the staging buffers
    *cmds = (OVGPUMcCmd *)b->h_cmds;
    if (prof) *prof = (OVGPUProfInfo *)b->h_prof;
    return 0;
}

int ovgpu_mc_batch_commit(OVGPUCtx *ctx, OVGPUMcBatch *b, int nb_cmds, int nb_prof)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !b || nb_cmds < 0 || nb_cmds > b->max_cmds || nb_prof < 0) return OVGPU_EINVAL;
    if (nb_prof > b->max_prof) return ovgpu_set_error(ctx, OVGPU_EINVAL, "PROF table too large (%d > %d)", nb_prof, b->max_prof);
    OVGPU_CUDA_CHECK(ctx, fence_wait(&b->fence));
    // one pass over the commands: validate, split into <=16x16 work items per kernel, note extents and references
    const OVGPUMcCmd *cmds = (const OVGPUMcCmd *)b->h_cmds;
    McItemDev *ip = b->h_items, *ib = b->h_items + b->max_items, *iq = b->h_items + 2 * b->max_items, *id = b->h_items + 3 * b->max_items, *is = b->h_items + 4 * b->max_items, *ig = b->h_items + 5 * b->max_items;
    int np = 0, nb = 0, nq = 0, nd = 0, ns = 0, ng = 0, max_x = 0, max_y = 0;
    uint32_t used[2] = { 0, 0 };
    {
    HostSection hs_loop(HT_MC_LOOP);
    for (int i = 0; i < nb_cmds; i++) {
        const OVGPUMcCmd *c = &cmds[i];
        if (mc_check_cmd(c, nb_prof)) return ovgpu_set_error(ctx, OVGPU_EINVAL, "MC command %d malformed", i);
        const int w = 1 << c->log2_w, h = 1 << c->log2_h;
        if (c->x0 + w > max_x) max_x = c->x0 + w;
        if (c->y0 + h > max_y) max_y = c->y0 + h;
        if (c->tool == OVGPU_MC_TOOL_GPM) { used[(c->aux >> 8) & 1] |= 1u << c->ref_idx0; used[(c->aux >> 9) & 1] |= 1u << c->ref_idx1; }
        else {
            if (c->inter_dir & 1) used[0] |= 1u << c->ref_idx0;
            if (c->inter_dir & 2) used[1] |= 1u << c->ref_idx1;
        }
        const int route = mc_route(c);
        if (route == OVGPU_MC_TOOL_PLAIN) {
            if (np + ((w + 15) >> 4) * ((h + 15) >> 4) > b->max_items) return ovgpu_set_error(ctx, OVGPU_EINVAL, "MC batch: too many work items");
            for (int oy = 0; oy < h; oy += 16)
                for (int ox = 0; ox < w; ox += 16) {
                    McItemDev it = { (uint32_t)i, (uint8_t)ox, (uint8_t)oy, (uint8_t)(w - ox < 16 ? w - ox : 16), (uint8_t)(h - oy < 16 ? h - oy : 16) };
                    ip[np++] = it;
                }
        } else if (route == MC_ROUTE_BIG) {
            const int th = h >= 32 ? 32 : 16;
            if (ng + (w >> 5) * (h / th) > b->max_items) return ovgpu_set_error(ctx, OVGPU_EINVAL, "MC batch: too many work items");
            for (int oy = 0; oy < h; oy += th)
                for (int ox = 0; ox < w; ox += 32) {
                    McItemDev it = { (uint32_t)i, (uint8_t)ox, (uint8_t)oy, 32, (uint8_t)th };
                    ig[ng++] = it;
                }
        } else {
            McItemDev it = { (uint32_t)i, 0, 0, (uint8_t)w, (uint8_t)h };
            if (route == OVGPU_MC_TOOL_BDOF)      { if (nb >= b->max_items) return OVGPU_EINVAL; ib[nb++] = it; }
            else if (route == OVGPU_MC_TOOL_DMVR) { if (nd >= b->max_items) return OVGPU_EINVAL; id[nd++] = it; }
            else if (route == MC_ROUTE_SMALL_C)   { if (ns >= b->max_items) return OVGPU_EINVAL; is[ns++] = it; }
            else                                    { if (nq >= b->max_items) return OVGPU_EINVAL; iq[nq++] = it; }
        }
    }
    b->nb_cmds = nb_cmds; b->nb_prof = nb_prof; b->nb_items = np; b->nb_bdof = nb; b->nb_prof_items = nq; b->nb_dmvr = nd; b->nb_small_c = ns; b->nb_big = ng;
    b->max_x = max_x; b->max_y = max_y; b->used_refs[0] = used[0]; b->used_refs[1] = used[1];
    }
    HostSection hs_copies(HT_MC_COPIES);
    if (nb_cmds) {
        OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_cmds, b->h_cmds, sizeof(McCmdDev) * nb_cmds, cudaMemcpyHostToDevice, ctx->stream));
        if (np) OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_items, ip, sizeof(McItemDev) * np, cudaMemcpyHostToDevice, ctx->stream));
        if (nb) OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_items + b->max_items, ib, sizeof(McItemDev) * nb, cudaMemcpyHostToDevice, ctx->stream));
        if (nq) OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_items + 2 * b->max_items, iq, sizeof(McItemDev) * nq, cudaMemcpyHostToDevice, ctx->stream));
        if (nd) OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_items + 3 * b->max_items, id, sizeof(McItemDev) * nd, cudaMemcpyHostToDevice, ctx->stream));
        if (ns) OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_items + 4 * b->max_items, is, sizeof(McItemDev) * ns, cudaMemcpyHostToDevice, ctx->stream));
        if (ng) OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_items + 5 * b->max_items, ig, sizeof(McItemDev) * ng, cudaMemcpyHostToDevice, ctx->stream));
        if (nb_prof)
            OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_prof, b->h_prof, sizeof(ProfInfoDev) * nb_prof, cudaMemcpyHostToDevice, ctx->stream));
        OVGPU_CUDA_CHECK(ctx, fence_mark(&b->fence, ctx->stream));
    }
    return 0;
}

int ovgpu_mc_batch_upload2(OVGPUCtx *ctx, OVGPUMcBatch *b, const OVGPUMcCmd *cmds, int nb_cmds,
                           const OVGPUProfInfo *prof, int nb_prof)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !b || !cmds || nb_cmds < 0 || nb_cmds > b->max_cmds || nb_prof < 0 || (nb_prof && !prof)) return OVGPU_EINVAL;
    if (nb_prof > b->max_prof) return ovgpu_set_error(ctx, OVGPU_EINVAL, "PROF table too large (%d > %d)", nb_prof, b->max_prof);
    OVGPUMcCmd *hc; OVGPUProfInfo *hp;
    int ret = ovgpu_mc_batch_staging(ctx, b, &hc, &hp);
    if (ret) return ret;
    memcpy(hc, cmds, sizeof(McCmdDev) * nb_cmds);
    if (nb_prof) memcpy(hp, prof, sizeof(ProfInfoDev) * nb_prof);
    return ovgpu_mc_batch_commit(ctx, b, nb_cmds, nb_prof);
}

int ovgpu_mc_picture(OVGPUCtx *ctx, OVGPUPicture *dst, const OVGPUMcRefs *refs, const OVGPUMcBatch *b,
                     const OVGPULmcs *lmcs)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !dst || !refs || !b) return OVGPU_EINVAL;
    if (!(b->nb_items + b->nb_bdof + b->nb_prof_items + b->nb_dmvr + b->nb_small_c + b->nb_big)) return 0;
    McK::Args a;
    memset(&a, 0, sizeof(a));
    for (int c = 0; c < 3; c++) { PlaneView v = plane_view(dst, c); a.dst[c] = v.p; a.dst_pitch[c] = v.pitch; }
    a.pic_w = dst->width; a.pic_h = dst->height; a.bitdepth = dst->bitdepth;
    a.cmds = b->d_cmds; a.items = b->d_items; a.nb_items = b->nb_items;
    a.lmcs_fwd = lmcs ? lmcs->d_luts->fwd_lut : NULL;
    // every reference a command names must be present, and every command must lie inside the picture
    for (int l = 0; l < 2; l++)
        for (int i = 0; i < 16; i++)
            if (((b->used_refs[l] >> i) & 1) && !refs->rpl[l][i])
                return ovgpu_set_error(ctx, OVGPU_EINVAL, "MC commands reference a missing picture (list %d, index %d)", l, i);
    if (b->max_x > dst->width || b->max_y > dst->height)
        return ovgpu_set_error(ctx, OVGPU_EINVAL, "MC commands lie outside the picture");
    for (int l = 0; l < 2; l++)
        for (int i = 0; i < 16; i++) {
            const OVGPUPicture *p = refs->rpl[l][i];
            a.refs.poc[l][i] = refs->poc[l][i];
            if (!p) continue;
            if (p->width != dst->width || p->height != dst->height)
                return ovgpu_set_error(ctx, OVGPU_EINVAL, "reference picture size differs (RPR is not supported)");
            for (int c = 0; c < 3; c++) { a.refs.plane[l][i][c] = p->plane[c]; a.refs.pitch[l][i][c] = (int)(p->pitch[c] / 2); }
        }
    StageScope scope(ctx, OVGPU_STAGE_MC, (b->nb_items > 0) + (b->nb_bdof > 0) + (b->nb_prof_items > 0) + (b->nb_dmvr > 0) + (b->nb_small_c > 0) + (b->nb_big > 0));
    if (b->nb_big) {
        McK::Args t = a;
        t.items = b->d_items + 5 * b->max_items; t.nb_items = b->nb_big;
        OVGPU_CUDA_CHECK(ctx, ovd_launch<McBigK>(t, dim3(OVGPU_CEIL_DIV(b->nb_big, McBigK::WARPS), 1, 1), McBigK::SMEM_BYTES, ctx->stream));
    }
    if (b->nb_items)
        OVGPU_CUDA_CHECK(ctx, ovd_launch<McK>(a, dim3(OVGPU_CEIL_DIV(b->nb_items, McK::WARPS), 1, 1), McK::SMEM_BYTES, ctx->stream));
    if (b->nb_bdof) {
        McK::Args t = a;
        t.items = b->d_items + b->max_items; t.nb_items = b->nb_bdof;
        OVGPU_CUDA_CHECK(ctx, ovd_launch<BdofK>(t, dim3(OVGPU_CEIL_DIV(b->nb_bdof, BdofK::WARPS), 1, 1), BdofK::SMEM_BYTES, ctx->stream));
    }
    if (b->nb_small_c) {
        McK::Args t = a;
        t.items = b->d_items + 4 * b->max_items; t.nb_items = b->nb_small_c;
        OVGPU_CUDA_CHECK(ctx, ovd_launch<McSmallCK>(t, dim3(OVGPU_CEIL_DIV(b->nb_small_c, McSmallCK::WARPS * McSmallCK::SBS), 1, 1),
                                                    McSmallCK::SMEM_BYTES, ctx->stream));
    }
    if (b->nb_dmvr) {
        DmvrK::Args t;
        t.mc = a; t.mc.items = b->d_items + 3 * b->max_items; t.mc.nb_items = b->nb_dmvr; t.out = b->d_dmvr;
        OVGPU_CUDA_CHECK(ctx, ovd_launch<DmvrK>(t, dim3(OVGPU_CEIL_DIV(b->nb_dmvr, DmvrK::WARPS), 1, 1), DmvrK::SMEM_BYTES, ctx->stream));
    }
    if (b->nb_prof_items) {
        ProfK::Args t;
        t.mc = a; t.mc.items = b->d_items + 2 * b->max_items; t.mc.nb_items = b->nb_prof_items; t.prof = b->d_prof;
        OVGPU_CUDA_CHECK(ctx, ovd_launch<ProfK>(t, dim3(OVGPU_CEIL_DIV(b->nb_prof_items, ProfK::WARPS * ProfK::SBS), 1, 1),
                                                ProfK::SMEM_BYTES, ctx->stream));
    }
    return 0;
}

int ovgpu_mc_batch_read_dmvr(OVGPUCtx *ctx, OVGPUMcBatch *b, OVGPUDmvrMv *out, int max)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !b || (max > 0 && !out)) return OVGPU_EINVAL;
    const int n = b->nb_dmvr < max ? b->nb_dmvr : max;
    if (n > 0) {
        OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->h_dmvr, b->d_dmvr, sizeof(DmvrMvDev) * n, cudaMemcpyDeviceToHost, ctx->stream));
        OVGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(out, b->h_dmvr, sizeof(DmvrMvDev) * n);
    }
    return b->nb_dmvr;
}

} // extern "C"

int ovgpu_lmcs_scale_params(const OVGPULmcs *l, const uint16_t **d_wnd_bnd, int *min_idx, int *max_idx_p1, int *crs_offset)
{
    if (!l || !l->has_data) return -1;
    *d_wnd_bnd = l->d_luts->wnd_bnd; *min_idx = l->min_idx; *max_idx_p1 = l->max_idx_p1; *crs_offset = l->crs_offset;
    return 0;
}
