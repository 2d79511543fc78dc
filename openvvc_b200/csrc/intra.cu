// intra.cu -- C-ABI entry points of luma + chroma intra prediction (kernel in intra_kernels.h, scheduling in intra_host.h).
// Replaces the intra_pred / intra_pred_c slots (vvc_intra_pred, libovvc/rcn_intra.c:484; vvc_intra_pred_chroma, :1015)
// for the intra blocks of a picture.
#include "common.cuh"
#define OVGPU_RECON_KERNEL_IMPL
#include "intra_host.h"
#include "recon_kernels.h"
#include <stdlib.h>

static_assert(sizeof(IntraCmdDev) == sizeof(OVGPUIntraCmd) && sizeof(IntraCmdDev) == 32, "intra command layout");

struct OVGPUIntraBatch {
    int max_cmds, nb_cmds, device, pic_w, pic_h, log2_ctu;
    IntraCmdDev *h_cmds, *d_cmds;
    OVGPUTuBatch *empty_tu;                         // residual batch without commands for prediction-only reconstruction calls
    IntraLevels lv;                                 // per plane: order (pinned, 2 * max_cmds) and level starts (host only)
    uint32_t *h_order, *d_order;                    // [plane][max_cmds]
    IspInfoDev *h_isp, *d_isp;                      // ISP table (grown on demand)
    int nb_isp, max_isp, has_isp;
    int map_w_u, map_cmds;              // row length of the position maps and how many of h_cmds are mapped into them
    uint32_t *pos_l, *pos_c;                        // host only: 1 + queue position of the command covering each 4x4 unit (luma / chroma plane)
    size_t map_units;
    // dataflow state of ovgpu_recon_picture (recon_kernels.h): pending word per unit and plane, queue heads, epoch
    uint32_t *d_pend;                               // [plane][dev_units]
    size_t dev_units;
    int *d_counter;                                 // [2]
    uint32_t epoch;
    uint64_t serial;
    StagingFence fence;
};

extern "C" {

int ovgpu_intra_batch_create(OVGPUCtx *ctx, int max_cmds, OVGPUIntraBatch **out)
{
    OVGPU_DEVICE_GUARD(ctx);
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
    b->lv.ws_lvl = (uint16_t *)malloc(sizeof(uint16_t) * 2 * (size_t)max_cmds);
    b->lv.ws_fill = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)max_cmds + 3));
    if (!b->lv.ws_lvl || !b->lv.ws_fill) { ovgpu_intra_batch_destroy(b); return ovgpu_set_error(ctx, OVGPU_ENOMEM, "intra level scratch"); }
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
    if (b->h_isp) cudaFreeHost(b->h_isp);
    if (b->d_isp) cudaFree(b->d_isp);
    if (b->empty_tu) ovgpu_tu_batch_destroy(b->empty_tu);
    free(b->pos_l); free(b->pos_c); free(b->lv.ws_map); free(b->lv.ws_lvl); free(b->lv.ws_fill);
    if (b->d_pend) cudaFree(b->d_pend);
    if (b->d_counter) cudaFree(b->d_counter);
    fence_destroy(&b->fence);
    free(b);
}

int ovgpu_intra_batch_upload(OVGPUCtx *ctx, OVGPUIntraBatch *b, const OVGPUIntraCmd *cmds, int nb_cmds,
                             int pic_w, int pic_h, int log2_ctu_s)
{
    OVGPU_DEVICE_GUARD(ctx);
    return ovgpu_intra_batch_upload2(ctx, b, cmds, nb_cmds, NULL, 0, pic_w, pic_h, log2_ctu_s);
}

int ovgpu_intra_batch_upload2(OVGPUCtx *ctx, OVGPUIntraBatch *b, const OVGPUIntraCmd *cmds, int nb_cmds,
                              const OVGPUIspInfo *isp, int nb_isp, int pic_w, int pic_h, int log2_ctu_s)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !b || nb_isp < 0 || nb_isp > 65536 || (nb_isp && !isp)) return OVGPU_EINVAL;
    if (!ctx || !b || nb_cmds < 0 || nb_cmds > b->max_cmds || (nb_cmds && !cmds) || log2_ctu_s < 5 || log2_ctu_s > 7 ||
        pic_w <= 0 || pic_h <= 0 || (pic_w & 7) || (pic_h & 7))
        return OVGPU_EINVAL;
    {
        HostSection hs(HT_INTRA_CHECK);
        for (int i = 0; i < nb_cmds; i++)
            if (intra_check_cmd(&cmds[i], pic_w, pic_h) ||
                ((cmds[i].flags & OVGPU_INTRA_ISP) && (cmds[i].pad[0] | (cmds[i].pad[1] << 8)) >= nb_isp))
                return ovgpu_set_error(ctx, OVGPU_EINVAL, "intra command %d malformed", i);
    }
    OVGPU_CUDA_CHECK(ctx, fence_wait(&b->fence));
    b->has_isp = 0;
    for (int i = 0; i < nb_cmds; i++) b->has_isp |= (cmds[i].flags & (OVGPU_INTRA_ISP | OVGPU_INTRA_IBC)) != 0;
    if (nb_isp > b->max_isp) {
        if (b->h_isp) cudaFreeHost(b->h_isp);
        if (b->d_isp) cudaFree(b->d_isp);
        b->h_isp = NULL; b->d_isp = NULL; b->max_isp = 0;
        if (cudaMallocHost(&b->h_isp, sizeof(IspInfoDev) * nb_isp) != cudaSuccess || cudaMalloc(&b->d_isp, sizeof(IspInfoDev) * nb_isp) != cudaSuccess)
            return ovgpu_set_error(ctx, OVGPU_ENOMEM, "ISP table");
        b->max_isp = nb_isp;
    }
    b->nb_isp = nb_isp;
    if (nb_isp) {
        memcpy(b->h_isp, isp, sizeof(IspInfoDev) * nb_isp);
        OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_isp, b->h_isp, sizeof(IspInfoDev) * nb_isp, cudaMemcpyHostToDevice, ctx->stream));
    }
    b->nb_cmds = nb_cmds; b->pic_w = pic_w; b->pic_h = pic_h; b->log2_ctu = log2_ctu_s; b->lv.nb[0] = b->lv.nb[1] = 0;
    {   // process-wide unique: the TU batch keys its cached reconstruction lists on it, and two intra batches must never share a value
        static uint64_t next_serial = 0;
        b->serial = __atomic_add_fetch(&next_serial, 1, __ATOMIC_RELAXED);
    }
    // the position maps stay all zero between pictures except for the blocks of the commands last mapped into them, which are
    // cleared here block by block (b->h_cmds still holds those commands) instead of clearing 2 x 2 MB per 4K picture
    const size_t units = (size_t)(pic_w >> 2) * (pic_h >> 2);
    if (units > b->map_units || (pic_w >> 2) != b->map_w_u) {
        free(b->pos_l); free(b->pos_c); free(b->lv.ws_map);
        b->pos_l = (uint32_t *)calloc(units, sizeof(uint32_t)); b->pos_c = (uint32_t *)calloc(units, sizeof(uint32_t));
        b->lv.ws_map = (uint16_t *)calloc(units, sizeof(uint16_t));
        b->map_units = (b->pos_l && b->pos_c && b->lv.ws_map) ? units : 0;
        b->map_w_u = pic_w >> 2; b->map_cmds = 0;
        if (!b->map_units) return ovgpu_set_error(ctx, OVGPU_ENOMEM, "intra position maps");
    } else if (b->map_cmds) {
        intra_clear_units(b->pos_l, b->map_w_u, (const OVGPUIntraCmd *)b->h_cmds, b->map_cmds, OVGPU_INTRA_LUMA);
        intra_clear_units(b->pos_c, b->map_w_u, (const OVGPUIntraCmd *)b->h_cmds, b->map_cmds, OVGPU_INTRA_CHROMA);
        b->map_cmds = 0;
    }
    b->lv.ws_clean = 1;
    b->lv.units_mask = ((1 << log2_ctu_s) >> 2) - 1;
    if (!nb_cmds) return 0;
    memcpy(b->h_cmds, cmds, sizeof(IntraCmdDev) * nb_cmds);
    b->map_cmds = nb_cmds;
    int lrc;
    { HostSection hs(HT_INTRA_LEVELS); lrc = intra_build_levels(&b->lv, b->max_cmds, cmds, nb_cmds, pic_w, pic_h, b->pos_l, b->pos_c); }
    if (lrc == -2) { b->nb_cmds = 0; return ovgpu_set_error(ctx, OVGPU_EINVAL, "a command waits on samples that are not reconstructed before it (block copy source, or an availability word that flags a unit of a later command)"); }
    if (lrc)
        return ovgpu_set_error(ctx, OVGPU_ENOMEM, "intra dependency levels");
    HostSection hs_copies(HT_INTRA_COPIES);
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
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !pic || !b) return OVGPU_EINVAL;
    if (!b->nb_cmds) return 0;
    if (b->has_isp) return ovgpu_set_error(ctx, OVGPU_EINVAL, "ISP and block copy commands are reconstructed by ovgpu_recon_picture only");
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

// the intra commands of plane p + the transform blocks they cover: one dataflow launch
static int recon_plane(OVGPUCtx *ctx, OVGPUPicture *pic, OVGPUIntraBatch *b, const OVGPUTuBatch *tu, int p, int nb_items)
{
    if (!nb_items) return 0;
    // per DEVICE (function attributes and occupancy are): slot = the context's device; a racing second thread computes the same values
    static int configured_dev[64], blocks_per_sm_dev[64];
    const int slot = ctx->device & 63;
    int blocks_per_sm = __atomic_load_n(&blocks_per_sm_dev[slot], __ATOMIC_RELAXED);
    if (!__atomic_load_n(&configured_dev[slot], __ATOMIC_ACQUIRE)) {
        cudaFuncSetAttribute(ovd_recon_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, ovd_recon_kernel, 128, RECON_SMEM_BYTES) != cudaSuccess || blocks_per_sm < 1)
            blocks_per_sm = 1;
        // tuning knob: resident workers per SM (default: what fits).  Fewer workers leave shared memory and issue slots to the
        // kernels of other pictures running next to this one; the queue order keeps any number >= 1 deadlock-free.
        const char *e = getenv("OVGPU_RECON_WORKERS_PER_SM");
        if (e && atoi(e) >= 1 && atoi(e) < blocks_per_sm) blocks_per_sm = atoi(e);
        __atomic_store_n(&blocks_per_sm_dev[slot], blocks_per_sm, __ATOMIC_RELAXED);
        __atomic_store_n(&configured_dev[slot], 1, __ATOMIC_RELEASE);
    }
    PlaneView v0 = plane_view(pic, 0), v1 = plane_view(pic, 1), v2 = plane_view(pic, 2);
    ReconArgs a;
    memset(&a, 0, sizeof(a));
    a.plane[0] = v0.p; a.plane[1] = v1.p; a.plane[2] = v2.p; a.pitch[0] = v0.pitch; a.pitch[1] = v1.pitch; a.pitch[2] = v2.pitch;
    a.log2_ctu = b->log2_ctu; a.bitdepth = pic->bitdepth; a.plane_id = p; a.cmds = b->d_cmds; a.nb_items = nb_items; a.isp = b->d_isp;
    ovgpu_tu_recon_args(tu, p, &a);
    a.pend = b->d_pend + (size_t)p * b->dev_units; a.serial = b->epoch; a.w_u = pic->width >> 2; a.h_u = pic->height >> 2;
    a.counter = b->d_counter + p;
    // Resident workers: what the dependency graph can keep busy, not what fits.  A plane of n commands on L dependency levels
    // has n / L commands runnable at a time on average; workers beyond a few times that only spin on their neighbours' flags and
    // hold the shared memory and issue slots that the kernels of other pictures could use.  Measured on an all-intra 4K picture
    // (12.8 k commands, 1311 levels per plane): the same 16-19 ms per picture with 1 or 8 workers per SM, but 158 against 45
    // pictures/s with eight pictures in flight.  (The queue order keeps ANY number of workers >= 1 deadlock-free.)
    int max_ctas = ctx->num_sms * blocks_per_sm;
    const int levels = b->lv.nb[p] > 0 ? b->lv.nb[p] : 1, want = 4 * (nb_items / levels + 1);
    if (want < max_ctas) max_ctas = want > ctx->num_sms ? want : ctx->num_sms;
    StageScope scope(ctx, OVGPU_STAGE_INTRA, 1);
    ovd_recon_kernel<<<nb_items < max_ctas ? nb_items : max_ctas, 128, RECON_SMEM_BYTES, ctx->stream>>>(a);
    OVGPU_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

/* Reconstruction of a picture in the reference's order: prediction of a CU, then its residual (coding_unit() ->
 * prediction -> rcn_transform_tree, vcl_coding_unit.c / rcn_transform_tree.c:1433), expressed as picture-level passes, plane
 * by plane (luma never reads chroma): motion compensation; luma residual of the blocks outside intra CUs; the luma intra
 * commands with their residual as one dataflow launch (recon_kernels.h); the LMCS chroma scales from the reconstructed luma;
 * then chroma the same way (CCLM reads finished luma). */
int ovgpu_recon_picture(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPUMcRefs *refs, const OVGPUMcBatch *mc,
                        const OVGPULmcs *lmcs, OVGPUTuBatch *tu, OVGPUIntraBatch *intra)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !pic || (mc && !refs)) return OVGPU_EINVAL;
    int rc;
    if (mc) { HostSection hs(HT_RECON_MC); if ((rc = ovgpu_mc_picture(ctx, pic, refs, mc, lmcs)) != 0) return rc; }
    const int has_intra = intra && intra->nb_cmds;
    if (!tu && !has_intra) return 0;
    if (has_intra && (pic->width != intra->pic_w || pic->height != intra->pic_h))
        return ovgpu_set_error(ctx, OVGPU_EINVAL, "intra batch was built for another picture size");
    if (!has_intra && !ovgpu_tu_nb_anchors(tu)) return ovgpu_residual_picture(ctx, pic, tu);      // one plain residual pass
    if (!tu) {          // prediction only: an empty residual batch keeps one code path; it lives with the intra batch (no allocation
        if (!intra->empty_tu) {                                    // or stream synchronisation per call)
            if ((rc = ovgpu_tu_batch_create(ctx, 1, 1, &intra->empty_tu)) != 0) return rc;
            OVGPUTuSlice sl = { 0, 0, 0, 0 };
            if ((rc = ovgpu_tu_batch_commit(ctx, intra->empty_tu, 0, 0, &sl)) != 0) { ovgpu_tu_batch_destroy(intra->empty_tu); intra->empty_tu = NULL; return rc; }
        }
        tu = intra->empty_tu;
    }
    const int w_u = pic->width >> 2, h_u = pic->height >> 2;
    const uint32_t *order[2] = { has_intra ? intra->lv.order[0] : NULL, has_intra ? intra->lv.order[1] : NULL };
    const int nb[2] = { has_intra ? (int)intra->lv.start[0][intra->lv.nb[0]] : 0, has_intra ? (int)intra->lv.start[1][intra->lv.nb[1]] : 0 };
    { HostSection hs(HT_TU_CHECK); rc = ovgpu_tu_check_picture(ctx, pic, tu); }
    if (!rc) rc = ovgpu_tu_bind_recon(ctx, tu, has_intra ? intra->serial : 0, has_intra ? intra->pos_l : NULL, has_intra ? intra->pos_c : NULL,
                                      w_u, h_u, order, nb);
    HostSection hs_launches(HT_RECON_LAUNCHES);
    if (!rc && has_intra) {
        // device state: pending words (never cleared: an epoch number marks the units of this run) and the queue heads
        const size_t units = (size_t)w_u * h_u;
        if (units > intra->dev_units) {
            if (intra->d_pend) cudaFree(intra->d_pend);
            intra->d_pend = NULL; intra->dev_units = 0;
            if (cudaMalloc(&intra->d_pend, 2 * units * sizeof(uint32_t)) != cudaSuccess) rc = ovgpu_set_error(ctx, OVGPU_ENOMEM, "recon pending maps");
            else { intra->dev_units = units; cudaMemsetAsync(intra->d_pend, 0, 2 * units * sizeof(uint32_t), ctx->stream); }
        }
        if (!rc && !intra->d_counter && cudaMalloc(&intra->d_counter, 2 * sizeof(int)) != cudaSuccess) rc = ovgpu_set_error(ctx, OVGPU_ENOMEM, "recon counters");
        if (!rc) {
            if (++intra->epoch == 0) intra->epoch = 1;
            ReconArgs t0, t1;
            ovgpu_tu_recon_args(tu, 0, &t0); ovgpu_tu_recon_args(tu, 1, &t1);
            ReconPrepK::Args pa = { intra->d_cmds, { t0.ritems, t1.ritems }, { nb[0], nb[1] }, { intra->d_pend, intra->d_pend + intra->dev_units },
                                    intra->d_counter, intra->epoch, w_u };
            const int mx = nb[0] > nb[1] ? nb[0] : nb[1];
            StageScope scope(ctx, OVGPU_STAGE_INTRA, 1);
            cudaError_t e = ovd_launch<ReconPrepK>(pa, dim3(OVGPU_CEIL_DIV(mx, ReconPrepK::WARPS), 2, 1), 0, ctx->stream);
            if (e != cudaSuccess) rc = ovgpu_set_error(ctx, OVGPU_ECUDA, "recon prep launch: %s", cudaGetErrorString(e));
        }
    }
    for (int p = 0; p < 2 && !rc; p++) {
        if (p == 1) rc = recon_scales(ctx, pic, lmcs, tu);
        if (!rc) rc = ovgpu_tu_run_inter(ctx, pic, tu, p);
        if (!rc && has_intra) rc = recon_plane(ctx, pic, intra, tu, p, nb[p]);
    }
    return rc;
}

} // extern "C"
