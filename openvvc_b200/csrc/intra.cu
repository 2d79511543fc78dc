// intra.cu -- C-ABI entry points of luma + chroma intra prediction (kernel in intra_kernels.h, scheduling in intra_host.h).
// Replaces the intra_pred / intra_pred_c slots (vvc_intra_pred, libovvc/rcn_intra.c:484; vvc_intra_pred_chroma, :1015)
// for the intra blocks of a picture.
#include "common.cuh"
#include "intra_host.h"
#include <stdlib.h>

static_assert(sizeof(IntraCmdDev) == sizeof(OVGPUIntraCmd) && sizeof(IntraCmdDev) == 32, "intra command layout");

struct OVGPUIntraBatch {
    int max_cmds, nb_cmds, device, pic_w, pic_h, log2_ctu;
    IntraCmdDev *h_cmds, *d_cmds;
    IntraLevels lv;                                 // per plane: order (pinned, 2 * max_cmds) and level starts (host only)
    uint32_t *h_order, *d_order;                    // [plane][max_cmds]
    uint16_t *map_l, *map_c;                        // host only: level of the command covering each 4x4 unit (luma / chroma tree)
    size_t map_units;
    uint64_t serial;
    StagingFence fence;
};

extern "C" {

int ovgpu_intra_batch_create(OVGPUCtx *ctx, int max_cmds, OVGPUIntraBatch **out)
{
    if (!ctx || !out || max_cmds <= 0) return OVGPU_EINVAL;
    OVGPUIntraBatch *b = (OVGPUIntraBatch *)calloc(1, sizeof(*b));
    if (!b) return OVGPU_ENOMEM;
    b->max_cmds = max_cmds; b->device = ctx->device;
    b->lv.start[0] = (uint32_t *)malloc(sizeof(uint32_t) * 2 * ((size_t)max_cmds + 3));
    if (!b->lv.start[0] ||
        cudaMallocHost(&b->h_cmds, sizeof(IntraCmdDev) * max_cmds) != cudaSuccess ||
        cudaMallocHost(&b->h_order, sizeof(uint32_t) * 2 * max_cmds) != cudaSuccess ||
        cudaMalloc(&b->d_cmds, sizeof(IntraCmdDev) * max_cmds) != cudaSuccess ||
        cudaMalloc(&b->d_order, sizeof(uint32_t) * 2 * max_cmds) != cudaSuccess) {
        ovgpu_intra_batch_destroy(b);
        return ovgpu_set_error(ctx, OVGPU_ENOMEM, "intra batch buffers: %s", cudaGetErrorString(cudaGetLastError()));
    }
    b->lv.start[1] = b->lv.start[0] + max_cmds + 3;
    b->lv.order[0] = b->h_order; b->lv.order[1] = b->h_order + max_cmds;
    *out = b;
    return 0;
}

void ovgpu_intra_batch_destroy(OVGPUIntraBatch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->h_cmds) cudaFreeHost(b->h_cmds);
    if (b->h_order) cudaFreeHost(b->h_order);
    if (b->d_cmds) cudaFree(b->d_cmds);
    if (b->d_order) cudaFree(b->d_order);
    free(b->lv.start[0]);
    free(b->map_l); free(b->map_c);
    fence_destroy(&b->fence);
    free(b);
}

int ovgpu_intra_batch_upload(OVGPUCtx *ctx, OVGPUIntraBatch *b, const OVGPUIntraCmd *cmds, int nb_cmds,
                             int pic_w, int pic_h, int log2_ctu_s)
{
    if (!ctx || !b || nb_cmds < 0 || nb_cmds > b->max_cmds || (nb_cmds && !cmds) || log2_ctu_s < 5 || log2_ctu_s > 7 ||
        pic_w <= 0 || pic_h <= 0 || (pic_w & 7) || (pic_h & 7))
        return OVGPU_EINVAL;
    for (int i = 0; i < nb_cmds; i++)
        if (intra_check_cmd(&cmds[i], pic_w, pic_h)) return ovgpu_set_error(ctx, OVGPU_EINVAL, "intra command %d malformed", i);
    OVGPU_CUDA_CHECK(ctx, fence_wait(&b->fence));
    b->nb_cmds = nb_cmds; b->pic_w = pic_w; b->pic_h = pic_h; b->log2_ctu = log2_ctu_s; b->lv.nb[0] = b->lv.nb[1] = 0;
    b->serial++;
    const size_t units = (size_t)(pic_w >> 2) * (pic_h >> 2);
    if (units > b->map_units) {
        free(b->map_l); free(b->map_c);
        b->map_l = (uint16_t *)malloc(units * sizeof(uint16_t)); b->map_c = (uint16_t *)malloc(units * sizeof(uint16_t));
        b->map_units = (b->map_l && b->map_c) ? units : 0;
        if (!b->map_units) return ovgpu_set_error(ctx, OVGPU_ENOMEM, "intra level maps");
    }
    if (!nb_cmds) { memset(b->map_l, 0, units * sizeof(uint16_t)); memset(b->map_c, 0, units * sizeof(uint16_t)); return 0; }
    memcpy(b->h_cmds, cmds, sizeof(IntraCmdDev) * nb_cmds);
    if (intra_build_levels(&b->lv, b->max_cmds, cmds, nb_cmds, pic_w, pic_h, b->map_l, b->map_c))
        return ovgpu_set_error(ctx, OVGPU_ENOMEM, "intra dependency levels");
    OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_cmds, b->h_cmds, sizeof(IntraCmdDev) * nb_cmds, cudaMemcpyHostToDevice, ctx->stream));
    OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_order, b->h_order, sizeof(uint32_t) * 2 * b->max_cmds, cudaMemcpyHostToDevice, ctx->stream));
    OVGPU_CUDA_CHECK(ctx, fence_mark(&b->fence, ctx->stream));
    return 0;
}

int ovgpu_intra_batch_levels(const OVGPUIntraBatch *b) { return b ? b->lv.nb[0] + b->lv.nb[1] : OVGPU_EINVAL; }

static int launch_intra(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPUIntraBatch *b, int plane, int start, int n)
{
    PlaneView v0 = plane_view(pic, 0), v1 = plane_view(pic, 1), v2 = plane_view(pic, 2);
    IntraK::Args a = { { v0.p, v1.p, v2.p }, { v0.pitch, v1.pitch, v2.pitch }, b->log2_ctu, pic->bitdepth, b->d_cmds,
                       b->d_order + (size_t)plane * b->max_cmds + start, n, plane };
    cudaError_t err = ovd_launch<IntraK>(a, dim3(OVGPU_CEIL_DIV(n, IntraK::WARPS), plane ? 2 : 1, 1), IntraK::SMEM_BYTES, ctx->stream);
    return err == cudaSuccess ? 0 : ovgpu_set_error(ctx, OVGPU_ECUDA, "intra launch: %s", cudaGetErrorString(err));
}

int ovgpu_intra_picture(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPUIntraBatch *b)
{
    if (!ctx || !pic || !b) return OVGPU_EINVAL;
    if (!b->nb_cmds) return 0;
    if (pic->width != b->pic_w || pic->height != b->pic_h)
        return ovgpu_set_error(ctx, OVGPU_EINVAL, "intra batch was built for another picture size");
    StageScope scope(ctx, OVGPU_STAGE_INTRA, b->lv.nb[0] + b->lv.nb[1]);
    return intra_for_each_launch(&b->lv, [&](int plane, int start, int n) { return launch_intra(ctx, pic, b, plane, start, n); });
}

static int recon_scales(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPULmcs *lmcs, const OVGPUTuBatch *tu)
{
    if (!tu || !ovgpu_tu_nb_anchors(tu)) return 0;
    const uint16_t *bnd; int mn, mx, off;
    if (ovgpu_lmcs_scale_params(lmcs, &bnd, &mn, &mx, &off))
        return ovgpu_set_error(ctx, OVGPU_EINVAL, "LMCS anchors need an LMCS object with ovgpu_lmcs_set_data");
    return ovgpu_tu_derive_scales(ctx, pic, tu, bnd, mn, mx, off);
}

/* Reconstruction of a picture in the reference's order: prediction of a CU, then its residual (coding_unit() ->
 * prediction -> rcn_transform_tree, vcl_coding_unit.c / rcn_transform_tree.c:1433), expressed as picture-level passes:
 * motion compensation, residual of the blocks outside intra CUs, then per dependency level intra prediction + residual. */
int ovgpu_recon_picture(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPUMcRefs *refs, const OVGPUMcBatch *mc,
                        const OVGPULmcs *lmcs, OVGPUTuBatch *tu, const OVGPUIntraBatch *intra)
{
    if (!ctx || !pic || (mc && !refs)) return OVGPU_EINVAL;
    int rc;
    if (mc && (rc = ovgpu_mc_picture(ctx, pic, refs, mc, lmcs)) != 0) return rc;
    if (!intra || !intra->nb_cmds) {
        // no intra blocks: one residual pass -- unless chroma scales have to be derived from the reconstructed luma first
        if (!tu) return 0;
        if (!ovgpu_tu_nb_anchors(tu)) return ovgpu_residual_picture(ctx, pic, tu);
        if ((rc = ovgpu_tu_check_picture(ctx, pic, tu)) != 0) return rc;
        if ((rc = ovgpu_tu_bind_levels(ctx, tu, 0, NULL, NULL, pic->width >> 2, pic->height >> 2, 0, 0)) != 0) return rc;
        if ((rc = ovgpu_tu_run_group(ctx, pic, tu, 0)) != 0 || (rc = recon_scales(ctx, pic, lmcs, tu)) != 0) return rc;
        return ovgpu_tu_run_group(ctx, pic, tu, 1);
    }
    if (pic->width != intra->pic_w || pic->height != intra->pic_h)
        return ovgpu_set_error(ctx, OVGPU_EINVAL, "intra batch was built for another picture size");
    if (tu) {
        if ((rc = ovgpu_tu_check_picture(ctx, pic, tu)) != 0) return rc;
        if ((rc = ovgpu_tu_bind_levels(ctx, tu, intra->serial, intra->map_l, intra->map_c, pic->width >> 2, pic->height >> 2,
                                       intra->lv.nb[0], intra->lv.nb[1])) != 0) return rc;
    }
    return recon_for_each_launch(&intra->lv,
        [&](int plane, int start, int n) {
            StageScope scope(ctx, OVGPU_STAGE_INTRA, 1);
            return launch_intra(ctx, pic, intra, plane, start, n);
        },
        [&](int group) { return tu ? ovgpu_tu_run_group(ctx, pic, tu, group) : 0; },
        [&]() { return recon_scales(ctx, pic, lmcs, tu); });
}

} // extern "C"
