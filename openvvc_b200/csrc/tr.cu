// tr.cu -- C-ABI entry points of the residual path (kernels in tr_kernels.h, work lists in tr_host.h).
// Replaces tmp.rcn_tu_st / tmp.rcn_tu_c (libovvc/rcn_transform_tree.c:1228-1390) for all TUs of a picture.
#include "common.cuh"
#include "tr_host.h"
#include <stdlib.h>

static_assert(sizeof(TuCmdDev) == sizeof(OVGPUTuCmd) && sizeof(TuCmdDev) == 72, "TU command layout");

struct OVGPUTuBatch {
    int max_cmds, nb_cmds, nb_small, nb_large, device;
    int items_built;                    // the plain (non-reconstruction-order) item lists exist for this commit: built on first use
    uint64_t checked_serial; int checked_w, checked_h;      // last (commit, picture size) ovgpu_tu_check_picture passed
    size_t max_coeffs, nb_coeffs;
    OVGPUTuSlice slice;
    TuCmdDev *h_cmds, *d_cmds;
    TbItemDev *d_items;                 // small items first, then large ones (staged through h_bitems)
    int16_t *h_coeffs, *d_coeffs;
    StagingFence fence;                 // h_cmds, h_coeffs (staging / commit)
    // reconstruction order (ovgpu_recon_picture): items re-ordered by dependency group, see tr_host.h.  Built into staging
    // buffers of their own (h_bitems, h_ritems) behind their own fence, so that binding never waits for the coefficient
    // copy the commit has just queued.
    TuReconLists lists;
    TbItemDev *h_bitems;
    StagingFence fence_bind, fence_anchors;
    ReconItemDev *h_ritems, *d_ritems;  // [plane][max_ritems]
    int max_ritems, nb_ritems[2];
    uint64_t serial, bound_serial, bound_intra_serial;
    int bound_w_u, bound_h_u;
    // LMCS chroma scale anchors (ovgpu_tu_batch_set_anchors) and the scales LmcsScaleK derives for them
    LmcsAnchorDev *h_anchors, *d_anchors;
    int16_t *d_scales;
    int nb_anchors, max_anchors;
};
static_assert(sizeof(LmcsAnchorDev) == sizeof(OVGPULmcsAnchor) && sizeof(LmcsAnchorDev) == 8, "LMCS anchor layout");

template <int GROUP>
static cudaError_t launch_tr(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPUTuBatch *b, const TbItemDev *items, int n, cudaStream_t stream)
{
    typedef TrK<GROUP> K;
    typename K::Args a;
    memset(&a, 0, sizeof(a));
    for (int c = 0; c < 3; c++) { PlaneView v = plane_view(pic, c); a.dst[c] = v.p; a.dst_pitch[c] = v.pitch; }
    a.bitdepth = pic->bitdepth;
    a.cmds = b->d_cmds; a.items = items; a.nb_items = n; a.arena = b->d_coeffs;
    a.dep_quant = b->slice.dep_quant; a.sh_ts_disabled = b->slice.sh_ts_disabled; a.ict_type = b->slice.ict_type;
    a.scales = b->nb_anchors ? b->d_scales : NULL;
    return ovd_launch<K>(a, dim3(OVGPU_CEIL_DIV(n, K::GROUPS), 1, 1), K::SMEM_BYTES, stream);
}

extern "C" {

int ovgpu_tu_batch_create(OVGPUCtx *ctx, int max_cmds, size_t max_coeffs, OVGPUTuBatch **out)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !out || max_cmds <= 0 || !max_coeffs) return OVGPU_EINVAL;
    OVGPUTuBatch *b = (OVGPUTuBatch *)calloc(1, sizeof(*b));
    if (!b) return OVGPU_ENOMEM;
    b->max_cmds = max_cmds; b->max_coeffs = max_coeffs; b->device = ctx->device;
    if (cudaMallocHost(&b->h_cmds, sizeof(TuCmdDev) * max_cmds) != cudaSuccess ||
        cudaMallocHost(&b->h_bitems, sizeof(TbItemDev) * 3 * max_cmds) != cudaSuccess ||
        cudaMallocHost(&b->h_coeffs, sizeof(int16_t) * max_coeffs) != cudaSuccess ||
        cudaMalloc(&b->d_cmds, sizeof(TuCmdDev) * max_cmds) != cudaSuccess ||
        cudaMalloc(&b->d_items, sizeof(TbItemDev) * 3 * max_cmds) != cudaSuccess ||
        cudaMalloc(&b->d_coeffs, sizeof(int16_t) * max_coeffs) != cudaSuccess) {
        ovgpu_tu_batch_destroy(b);
        return ovgpu_set_error(ctx, OVGPU_ENOMEM, "TU batch buffers: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = b;
    return 0;
}

void ovgpu_tu_batch_destroy(OVGPUTuBatch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->h_cmds) cudaFreeHost(b->h_cmds);
    if (b->h_bitems) cudaFreeHost(b->h_bitems);
    if (b->h_coeffs) cudaFreeHost(b->h_coeffs);
    if (b->d_cmds) cudaFree(b->d_cmds);
    if (b->d_items) cudaFree(b->d_items);
    if (b->d_coeffs) cudaFree(b->d_coeffs);
    fence_destroy(&b->fence); fence_destroy(&b->fence_bind); fence_destroy(&b->fence_anchors);
    if (b->h_ritems) cudaFreeHost(b->h_ritems);
    if (b->d_ritems) cudaFree(b->d_ritems);
    if (b->h_anchors) cudaFreeHost(b->h_anchors);
    if (b->d_anchors) cudaFree(b->d_anchors);
    if (b->d_scales) cudaFree(b->d_scales);
    free(b);
}

int ovgpu_tu_batch_staging(OVGPUCtx *ctx, OVGPUTuBatch *b, OVGPUTuCmd **cmds, int16_t **coeffs)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !b || !cmds || !coeffs) return OVGPU_EINVAL;
    OVGPU_CUDA_CHECK(ctx, fence_wait(&b->fence));           // previous H2D copy out of the staging buffers
    *cmds = (OVGPUTuCmd *)b->h_cmds;
    *coeffs = b->h_coeffs;
    return 0;
}

int ovgpu_tu_batch_commit(OVGPUCtx *ctx, OVGPUTuBatch *b, int nb_cmds, size_t nb_coeffs, const OVGPUTuSlice *slice)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !b || !slice || nb_cmds < 0 || nb_cmds > b->max_cmds || nb_coeffs > b->max_coeffs) return OVGPU_EINVAL;
    const OVGPUTuCmd *cmds = (const OVGPUTuCmd *)b->h_cmds;
    {
        HostSection hs(HT_TU_CHECK);
        for (int i = 0; i < nb_cmds; i++)
            if (tu_check_cmd(&cmds[i], nb_coeffs))
                return ovgpu_set_error(ctx, OVGPU_EINVAL, "TU command %d malformed", i);
    }
    OVGPU_CUDA_CHECK(ctx, fence_wait(&b->fence));
    b->slice = *slice;
    b->nb_cmds = nb_cmds; b->nb_coeffs = nb_coeffs;
    b->serial++;
    b->lists.built = 0;
    for (int i = 0; i < nb_cmds; i++)
        if ((cmds[i].flags & OVGPU_TU_SCALE_ANCHOR) && cmds[i].scale_anchor >= b->nb_anchors)
            return ovgpu_set_error(ctx, OVGPU_EINVAL, "TU command %d names LMCS anchor %d of %d (set the anchors before the commands)",
                                   i, cmds[i].scale_anchor, b->nb_anchors);
    // the item lists of ovgpu_residual_picture are built when it first runs on this commit: ovgpu_recon_picture orders the
    // blocks its own way (ovgpu_tu_bind_recon) and never needs them
    b->nb_small = b->nb_large = 0; b->items_built = 0;
    if (!nb_cmds) { b->items_built = 1; return 0; }
    HostSection hs_copies(HT_TU_COPIES);
    OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_cmds, b->h_cmds, sizeof(TuCmdDev) * nb_cmds, cudaMemcpyHostToDevice, ctx->stream));
    OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_coeffs, b->h_coeffs, sizeof(int16_t) * nb_coeffs, cudaMemcpyHostToDevice, ctx->stream));
    OVGPU_CUDA_CHECK(ctx, fence_mark(&b->fence, ctx->stream));
    return 0;
}

int ovgpu_tu_batch_upload(OVGPUCtx *ctx, OVGPUTuBatch *b, const OVGPUTuCmd *cmds, int nb_cmds,
                          const int16_t *coeffs, size_t nb_coeffs, const OVGPUTuSlice *slice)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !b || !slice || nb_cmds < 0 || nb_cmds > b->max_cmds || nb_coeffs > b->max_coeffs ||
        (nb_cmds && (!cmds || !coeffs)))
        return OVGPU_EINVAL;
    OVGPUTuCmd *hc; int16_t *hk;
    int ret = ovgpu_tu_batch_staging(ctx, b, &hc, &hk);
    if (ret) return ret;
    if (nb_cmds) { memcpy(hc, cmds, sizeof(TuCmdDev) * nb_cmds); memcpy(hk, coeffs, sizeof(int16_t) * nb_coeffs); }
    return ovgpu_tu_batch_commit(ctx, b, nb_cmds, nb_coeffs, slice);
}

int ovgpu_tu_batch_set_anchors(OVGPUCtx *ctx, OVGPUTuBatch *b, const OVGPULmcsAnchor *anchors, int nb)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !b || nb < 0 || nb > 65536 || (nb && !anchors)) return OVGPU_EINVAL;
    HostSection hs(HT_TU_ANCHORS);
    OVGPU_CUDA_CHECK(ctx, fence_wait(&b->fence_anchors));
    if (nb > b->max_anchors) {
        if (b->h_anchors) cudaFreeHost(b->h_anchors);
        if (b->d_anchors) cudaFree(b->d_anchors);
        if (b->d_scales) cudaFree(b->d_scales);
        b->h_anchors = NULL; b->d_anchors = NULL; b->d_scales = NULL; b->max_anchors = 0;
        if (cudaMallocHost(&b->h_anchors, sizeof(LmcsAnchorDev) * nb) != cudaSuccess ||
            cudaMalloc(&b->d_anchors, sizeof(LmcsAnchorDev) * nb) != cudaSuccess ||
            cudaMalloc(&b->d_scales, sizeof(int16_t) * nb) != cudaSuccess)
            return ovgpu_set_error(ctx, OVGPU_ENOMEM, "LMCS anchor buffers");
        b->max_anchors = nb;
    }
    for (int i = 0; i < nb; i++)
        if (anchors[i].x0 < 0 || anchors[i].y0 < 0 || (anchors[i].x0 & 63) || (anchors[i].y0 & 63)) return ovgpu_set_error(ctx, OVGPU_EINVAL, "LMCS anchor %d malformed", i);
    b->nb_anchors = nb;
    if (nb) {
        memcpy(b->h_anchors, anchors, sizeof(LmcsAnchorDev) * nb);
        OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_anchors, b->h_anchors, sizeof(LmcsAnchorDev) * nb, cudaMemcpyHostToDevice, ctx->stream));
        OVGPU_CUDA_CHECK(ctx, fence_mark(&b->fence_anchors, ctx->stream));
    }
    return 0;
}

int ovgpu_residual_picture(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPUTuBatch *b)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !pic || !b) return OVGPU_EINVAL;
    if (pic->bitdepth != 10) return ovgpu_set_error(ctx, OVGPU_EINVAL, "residual path is built for 10-bit pictures");
    for (int i = 0; i < b->nb_cmds; i++) {
        const TuCmdDev *c = &b->h_cmds[i];
        if (c->x0 + (1 << c->log2_w) > pic->width || c->y0 + (1 << c->log2_h) > pic->height)
            return ovgpu_set_error(ctx, OVGPU_EINVAL, "TU command %d lies outside the picture", i);
    }
    if (b->lists.built)
        return ovgpu_set_error(ctx, OVGPU_EINVAL, "the batch is bound to ovgpu_recon_picture; commit it again to use ovgpu_residual_picture");
    if (!b->items_built) {          // first plain run on this commit (the lists are a cache inside the batch)
        OVGPUTuBatch *m = const_cast<OVGPUTuBatch *>(b);
        HostSection hs(HT_TU_ITEMS);
        OVGPU_CUDA_CHECK(ctx, fence_wait(&m->fence_bind));
        tu_build_items(NULL, NULL, &m->nb_small, &m->nb_large, (const OVGPUTuCmd *)m->h_cmds, m->nb_cmds);
        tu_build_items(m->h_bitems, m->h_bitems + m->nb_small, &m->nb_small, &m->nb_large, (const OVGPUTuCmd *)m->h_cmds, m->nb_cmds);
        OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(m->d_items, m->h_bitems, sizeof(TbItemDev) * (m->nb_small + m->nb_large), cudaMemcpyHostToDevice, ctx->stream));
        OVGPU_CUDA_CHECK(ctx, fence_mark(&m->fence_bind, ctx->stream));
        m->items_built = 1;
    }
    const int nk = (b->nb_small > 0) + (b->nb_large > 0);
    if (!nk) return 0;
    StageScope scope(ctx, OVGPU_STAGE_RESIDUAL, nk);
    if (b->nb_small) OVGPU_CUDA_CHECK(ctx, launch_tr<32>(ctx, pic, b, b->d_items, b->nb_small, ctx->stream));
    if (b->nb_large) OVGPU_CUDA_CHECK(ctx, launch_tr<128>(ctx, pic, b, b->d_items + b->nb_small, b->nb_large, ctx->stream));
    return 0;
}

} // extern "C"

// ---- internal: residual in reconstruction order (called by ovgpu_recon_picture, intra.cu) ---------------------------
// order[p] / nb[p]: the intra commands of plane p in queue order (NULL / 0 without intra commands)
int ovgpu_tu_bind_recon(OVGPUCtx *ctx, OVGPUTuBatch *b, uint64_t intra_serial, const uint32_t *pos_l, const uint32_t *pos_c,
                        int w_u, int h_u, const uint32_t *const order[2], const int nb[2])
{
    OVGPU_DEVICE_GUARD(ctx);
    if (b->lists.built && b->bound_serial == b->serial && b->bound_intra_serial == intra_serial && b->bound_w_u == w_u && b->bound_h_u == h_u) return 0;
    for (int i = 0; i < b->nb_cmds; i++) {
        const TuCmdDev *c = &b->h_cmds[i];
        if ((c->x0 >> 2) >= w_u || (c->y0 >> 2) >= h_u) return ovgpu_set_error(ctx, OVGPU_EINVAL, "TU command %d lies outside the picture", i);
    }
    OVGPU_CUDA_CHECK(ctx, fence_wait(&b->fence_bind));
    const int need = nb[0] > nb[1] ? nb[0] : nb[1];
    if (need > b->max_ritems) {
        if (b->h_ritems) cudaFreeHost(b->h_ritems);
        if (b->d_ritems) cudaFree(b->d_ritems);
        b->h_ritems = NULL; b->d_ritems = NULL; b->max_ritems = 0;
        if (cudaMallocHost(&b->h_ritems, sizeof(ReconItemDev) * 2 * need) != cudaSuccess ||
            cudaMalloc(&b->d_ritems, sizeof(ReconItemDev) * 2 * need) != cudaSuccess)
            return ovgpu_set_error(ctx, OVGPU_ENOMEM, "recon item buffers");
        b->max_ritems = need;
    }
    ReconItemDev *ritems[2] = { b->h_ritems, b->h_ritems + b->max_ritems };
    for (int p = 0; p < 2; p++) {
        b->nb_ritems[p] = nb[p];
        for (int q = 0; q < nb[p]; q++) ritems[p][q].cmd = order[p][q];
    }
    int n;
    { HostSection hs(HT_BIND_LISTS); n = tu_build_recon_lists(b->h_bitems, &b->lists, ritems, nb, (const OVGPUTuCmd *)b->h_cmds, b->nb_cmds, pos_l, pos_c, w_u); }
    HostSection hs_copies(HT_BIND_COPIES);
    if (n) OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_items, b->h_bitems, sizeof(TbItemDev) * n, cudaMemcpyHostToDevice, ctx->stream));
    if (need) OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_ritems, b->h_ritems, sizeof(ReconItemDev) * 2 * b->max_ritems, cudaMemcpyHostToDevice, ctx->stream));
    OVGPU_CUDA_CHECK(ctx, fence_mark(&b->fence_bind, ctx->stream));
    b->bound_serial = b->serial; b->bound_intra_serial = intra_serial; b->bound_w_u = w_u; b->bound_h_u = h_u;
    return 0;
}

// residual of the transform blocks of plane p that lie outside intra commands
int ovgpu_tu_run_inter(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPUTuBatch *b, int p)
{
    OVGPU_DEVICE_GUARD(ctx);
    const int ns = b->lists.n_small[p], nl = b->lists.n_large[p];
    if (!(ns + nl)) return 0;
    StageScope scope(ctx, OVGPU_STAGE_RESIDUAL, (ns > 0) + (nl > 0));
    // the two kernels work on different transform blocks: side by side (fork / join on a side stream of the context)
    static const int tr_par = getenv("OVGPU_TR_PARALLEL") ? atoi(getenv("OVGPU_TR_PARALLEL")) : 1;
    const int par = tr_par && ns && nl && ovgpu_ctx_aux(ctx) == 0;
    if (par) {
        OVGPU_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
        OVGPU_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->aux[0], ctx->ev_fork, 0));
    }
    if (nl) OVGPU_CUDA_CHECK(ctx, launch_tr<128>(ctx, pic, b, b->d_items + b->lists.start_large[p], nl, ctx->stream));
    if (ns) OVGPU_CUDA_CHECK(ctx, launch_tr<32>(ctx, pic, b, b->d_items + b->lists.start_small[p], ns, par ? ctx->aux[0] : ctx->stream));
    if (par) {
        OVGPU_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_join[0], ctx->aux[0]));
        OVGPU_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[0], 0));
    }
    return 0;
}

// the residual side of the dataflow kernel's arguments for plane p
void ovgpu_tu_recon_args(const OVGPUTuBatch *b, int p, ReconArgs *a)
{
    a->ritems = b->d_ritems + (size_t)p * b->max_ritems;
    a->tu_cmds = b->d_cmds; a->tb = b->d_items; a->arena = b->d_coeffs;
    a->dep_quant = b->slice.dep_quant; a->sh_ts_disabled = b->slice.sh_ts_disabled; a->ict_type = b->slice.ict_type;
    a->scales = b->nb_anchors ? b->d_scales : NULL;
}

// the chroma scales of the batch's anchors from the reconstructed luma plane of pic (all luma must be done)
int ovgpu_tu_derive_scales(OVGPUCtx *ctx, const OVGPUPicture *pic, const OVGPUTuBatch *b, const uint16_t *d_wnd_bnd,
                           int min_idx, int max_idx_p1, int crs_offset)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!b->nb_anchors) return 0;
    for (int i = 0; i < b->nb_anchors; i++) {
        const LmcsAnchorDev *a = &b->h_anchors[i];
        if (a->x0 >= pic->width || a->y0 >= pic->height || (a->abv_mask && !a->y0) || (a->lft_mask && !a->x0) ||
            a->x0 + 4 * (32 - __builtin_clz(a->abv_mask | 1u)) > pic->width + 3 * !a->abv_mask ||
            a->y0 + 4 * (32 - __builtin_clz(a->lft_mask | 1u)) > pic->height + 3 * !a->lft_mask)
            return ovgpu_set_error(ctx, OVGPU_EINVAL, "LMCS anchor %d reads outside the picture", i);
    }
    StageScope scope(ctx, OVGPU_STAGE_LMCS, 1);
    PlaneView v = plane_view(pic, 0);
    LmcsScaleK::Args a = { v.p, v.pitch, pic->bitdepth, b->d_anchors, b->nb_anchors, b->d_scales, d_wnd_bnd, min_idx, max_idx_p1, crs_offset };
    OVGPU_CUDA_CHECK(ctx, ovd_launch<LmcsScaleK>(a, dim3(OVGPU_CEIL_DIV(b->nb_anchors, LmcsScaleK::WARPS), 1, 1), LmcsScaleK::SMEM_BYTES, ctx->stream));
    return 0;
}

int ovgpu_tu_nb_anchors(const OVGPUTuBatch *b) { return b->nb_anchors; }

int ovgpu_tu_check_picture(OVGPUCtx *ctx, const OVGPUPicture *pic, const OVGPUTuBatch *b)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (pic->bitdepth != 10) return ovgpu_set_error(ctx, OVGPU_EINVAL, "residual path is built for 10-bit pictures");
    if (b->checked_serial == b->serial && b->checked_w == pic->width && b->checked_h == pic->height) return 0;      // same commit, same size
    for (int i = 0; i < b->nb_cmds; i++) {
        const TuCmdDev *c = &b->h_cmds[i];
        if (c->x0 + (1 << c->log2_w) > pic->width || c->y0 + (1 << c->log2_h) > pic->height)
            return ovgpu_set_error(ctx, OVGPU_EINVAL, "TU command %d lies outside the picture", i);
    }
    OVGPUTuBatch *m = const_cast<OVGPUTuBatch *>(b);
    m->checked_serial = b->serial; m->checked_w = pic->width; m->checked_h = pic->height;
    return 0;
}
