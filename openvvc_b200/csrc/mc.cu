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
    int nb_dmvr, nb_small_c, nb_big, nb_rpr;
    OVGPUMcScaling scaling;                   // ovgpu_mc_batch_set_scaling: the slice's scale factors (all 1 << 14 by default)
    uint32_t scaled[2];                       // ... and the reference indices whose factor is not 1, per list
    OVGPUMcScaling run_scaling;               // the table the LAST COMMIT was routed by: what ovgpu_mc_picture runs with, whatever
    uint32_t run_scaled[2];                   // set_scaling has been told since (the next picture's factors may already be in)
    DmvrMvDev *d_dmvr, *h_dmvr;               // refined MVs, one per DMVR item (device result + pinned read-back)
    McCmdDev *h_cmds, *d_cmds;
    McItemDev *d_items;                       // seven regions of max_items entries: plain | BDOF | PROF | DMVR | small chroma | big tiles | scaled,
                                              // built on the device by McPrep*K (mc_host.h)
    int32_t *d_block_sum;                     // scratch of the device-side list construction: [region][block of 256 commands]
    McPrepInfo *d_info, *h_info;              // its result (item counts, extents, references, first malformed command) + pinned copy
    cudaEvent_t ev_status;                    // context stream: the walk's verdict for the last run has been copied to h_info
    int status_pending;                       // ... and has not been looked at yet
    int nb_cmds_run;                          // commands of the commit that run belongs to
    int ran;                                  // ev_status has been recorded at least once
    // The upload and the walk run on the batch's own stream, so that they start as soon as the previous run's MC kernels are done
    // with the lists instead of queueing behind everything else the context's stream still holds for the previous picture.
    cudaStream_t prep_stream;
    cudaEvent_t ev_built;                     // prep stream: the lists of the last commit are built
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

// The verdict of the device-side walk for the last run of the batch (McGuardK, copied to pinned memory behind the MC kernels).
// wait = 0: only if the copy has already arrived.  Returns the error the host-side walk used to return from the commit.
static int mc_take_status(OVGPUCtx *ctx, OVGPUMcBatch *b, int wait)
{
    if (!b->status_pending) return 0;
    if (wait) OVGPU_CUDA_CHECK(ctx, cudaEventSynchronize(b->ev_status));
    else {
        const cudaError_t q = cudaEventQuery(b->ev_status);
        if (q == cudaErrorNotReady) return 0;
        OVGPU_CUDA_CHECK(ctx, q);
    }
    b->status_pending = 0;
    const McPrepInfo *f = b->h_info;
    switch (f->status) {
    case 0: return 0;
    case 1: return ovgpu_set_error(ctx, OVGPU_EINVAL, "MC command %d malformed", b->nb_cmds_run - f->status_arg);
    case 2: return ovgpu_set_error(ctx, OVGPU_EINVAL, "MC batch: too many work items");
    case 3: return ovgpu_set_error(ctx, OVGPU_EINVAL, "MC commands lie outside the picture");
    default: return ovgpu_set_error(ctx, OVGPU_EINVAL, "MC commands reference a missing picture (list %d, index %d)", f->status_arg >> 4, f->status_arg & 15);
    }
}

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
    for (int l = 0; l < 2; l++) for (int i = 0; i < 16; i++) b->scaling.scale[l][i][0] = b->scaling.scale[l][i][1] = 1 << 14;
    b->scaling.chroma_hor_col = 1;
    if (cudaMallocHost(&b->h_prof, sizeof(ProfInfoDev) * b->max_prof) != cudaSuccess ||
        cudaMalloc(&b->d_prof, sizeof(ProfInfoDev) * b->max_prof) != cudaSuccess ||
        cudaMallocHost(&b->h_cmds, sizeof(McCmdDev) * max_cmds) != cudaSuccess ||
        cudaMalloc(&b->d_block_sum, sizeof(int32_t) * MC_NB_REGIONS * OVGPU_CEIL_DIV(max_cmds, McPrepCountK::THREADS)) != cudaSuccess ||
        cudaMalloc(&b->d_info, sizeof(McPrepInfo)) != cudaSuccess || cudaMallocHost(&b->h_info, sizeof(McPrepInfo)) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_built, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&b->prep_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_status, cudaEventDisableTiming) != cudaSuccess ||
        cudaMallocHost(&b->h_dmvr, sizeof(DmvrMvDev) * max_items) != cudaSuccess ||
        cudaMalloc(&b->d_dmvr, sizeof(DmvrMvDev) * max_items) != cudaSuccess ||
        cudaMalloc(&b->d_cmds, sizeof(McCmdDev) * max_cmds) != cudaSuccess ||
        cudaMalloc(&b->d_items, sizeof(McItemDev) * MC_NB_REGIONS * max_items) != cudaSuccess) {
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
    if (b->d_block_sum) cudaFree(b->d_block_sum);
    if (b->d_info) cudaFree(b->d_info);
    if (b->h_info) cudaFreeHost(b->h_info);
    if (b->prep_stream) { cudaStreamSynchronize(b->prep_stream); cudaStreamDestroy(b->prep_stream); }
    if (b->ev_built) cudaEventDestroy(b->ev_built);
    if (b->ev_status) cudaEventDestroy(b->ev_status);
    if (b->d_cmds) cudaFree(b->d_cmds);
    if (b->d_items) cudaFree(b->d_items);
    if (b->h_dmvr) cudaFreeHost(b->h_dmvr);
    if (b->d_dmvr) cudaFree(b->d_dmvr);
    if (b->h_prof) cudaFreeHost(b->h_prof);
    if (b->d_prof) cudaFree(b->d_prof);
    fence_destroy(&b->fence);
    free(b);
}

int ovgpu_mc_batch_set_scaling(OVGPUCtx *ctx, OVGPUMcBatch *b, const OVGPUMcScaling *s)
{
    if (!ctx || !b) return OVGPU_EINVAL;
    // the standard bounds the ratio between 1/8 (a reference eight times smaller) and 2; a factor of 0 = entry not in use.
    // Checked before anything is changed: a refused table leaves the batch as it was.
    for (int l = 0; s && l < 2; l++)
        for (int i = 0; i < 16; i++)
            for (int k = 0; k < 2; k++) {
                const int f = s->scale[l][i][k];
                if (f && (f < 2048 || f > 32768)) return ovgpu_set_error(ctx, OVGPU_EINVAL, "scale factor %d of list %d, reference %d out of range", f, l, i);
            }
    b->scaled[0] = b->scaled[1] = 0;
    for (int l = 0; l < 2; l++)
        for (int i = 0; i < 16; i++) {
            for (int k = 0; k < 2; k++) {
                const int f = s ? s->scale[l][i][k] : 1 << 14;
                b->scaling.scale[l][i][k] = (uint16_t)(f ? f : 1 << 14);
            }
            if (b->scaling.scale[l][i][0] != (1 << 14) || b->scaling.scale[l][i][1] != (1 << 14)) b->scaled[l] |= 1u << i;
        }
    b->scaling.chroma_hor_col = s ? s->chroma_hor_col != 0 : 1;
    b->scaling.chroma_ver_col = s ? s->chroma_ver_col != 0 : 0;
    return 0;
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
    { const int ret = mc_take_status(ctx, b, 0); if (ret) return ret; }      // what the previous run of this batch found, if still unread
    // The commands go to the device as they are: validation, routing and the work-item lists are built there (mc_host.h).  The
    // host only counts the items per kernel to size the launches (no validation, no lists: a quarter of the old walk).
    b->nb_cmds = nb_cmds; b->nb_prof = nb_prof;
    b->run_scaling = b->scaling; b->run_scaled[0] = b->scaled[0]; b->run_scaled[1] = b->scaled[1];
    {
        HostSection hs_loop(HT_MC_LOOP);
        int cnt[MC_NB_REGIONS];
        mc_count_regions((const OVGPUMcCmd *)b->h_cmds, nb_cmds, b->max_items, cnt, b->scaled);
        b->nb_items = cnt[0]; b->nb_bdof = cnt[1]; b->nb_prof_items = cnt[2]; b->nb_dmvr = cnt[3]; b->nb_small_c = cnt[4]; b->nb_big = cnt[5]; b->nb_rpr = cnt[6];
    }
    HostSection hs_copies(HT_MC_COPIES);
    if (nb_cmds) {
        cudaStream_t ps = b->prep_stream;
        // (on the device) the MC kernels of the batch's previous run read the lists and the counters
        if (b->ran) OVGPU_CUDA_CHECK(ctx, cudaStreamWaitEvent(ps, b->ev_status, 0));
        // OVGPU_MC_ZERO_COPY=1 (measurement knob): McPrepCountK reads the pinned staging buffers through their device mapping and
        // writes the device copies, no copy engine involved (measured: 1379 against 1394 pictures/s end to end, no gain); by default
        // the commands take the DMA path and the kernels read the device copy
        static const int zero_copy = getenv("OVGPU_MC_ZERO_COPY") != NULL;
        McPrepArgs pa = { zero_copy ? b->h_cmds : b->d_cmds, b->d_cmds, zero_copy ? (const uint32_t *)b->h_prof : (const uint32_t *)b->d_prof, (uint32_t *)b->d_prof,
                          nb_cmds, nb_prof, OVGPU_CEIL_DIV(nb_cmds, McPrepCountK::THREADS), b->max_items, { b->scaled[0], b->scaled[1] }, b->d_block_sum, b->d_items, b->d_info };
        if (!zero_copy) {
            OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_cmds, b->h_cmds, sizeof(McCmdDev) * nb_cmds, cudaMemcpyHostToDevice, ps));
            if (nb_prof) OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->d_prof, b->h_prof, sizeof(ProfInfoDev) * nb_prof, cudaMemcpyHostToDevice, ps));
            OVGPU_CUDA_CHECK(ctx, fence_mark(&b->fence, ps));
        }
        OVGPU_CUDA_CHECK(ctx, cudaMemsetAsync(b->d_info, 0, sizeof(McPrepInfo), ps));
        OVGPU_CUDA_CHECK(ctx, ovd_launch<McPrepCountK>(pa, dim3(pa.nb_blocks, 1, 1), McPrepCountK::SMEM_BYTES, ps));
        if (zero_copy) OVGPU_CUDA_CHECK(ctx, fence_mark(&b->fence, ps));       // the staging buffers have been read
        OVGPU_CUDA_CHECK(ctx, ovd_launch<McPrepScanK>(pa, dim3(1, 1, 1), McPrepScanK::SMEM_BYTES, ps));
        OVGPU_CUDA_CHECK(ctx, ovd_launch<McPrepScatterK>(pa, dim3(pa.nb_blocks, 1, 1), McPrepScatterK::SMEM_BYTES, ps));
        OVGPU_CUDA_CHECK(ctx, cudaEventRecord(b->ev_built, ps));
        ctx->launches += 3;
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

int ovgpu_mc_picture(OVGPUCtx *ctx, OVGPUPicture *dst, const OVGPUMcRefs *refs, const OVGPUMcBatch *cb,
                     const OVGPULmcs *lmcs)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !dst || !refs || !cb) return OVGPU_EINVAL;
    OVGPUMcBatch *b = (OVGPUMcBatch *)cb;              // (the batch's CONTENT is not changed: launch bookkeeping only)
    { const int ret = mc_take_status(ctx, b, 0); if (ret) return ret; }        // a previous run's verdict, if it is in and unread (never waits)
    if (!b->nb_cmds || !(b->nb_items + b->nb_bdof + b->nb_prof_items + b->nb_dmvr + b->nb_small_c + b->nb_big + b->nb_rpr)) return 0;
    McK::Args a;
    memset(&a, 0, sizeof(a));
    for (int c = 0; c < 3; c++) { PlaneView v = plane_view(dst, c); a.dst[c] = v.p; a.dst_pitch[c] = v.pitch; }
    a.pic_w = dst->width; a.pic_h = dst->height; a.bitdepth = dst->bitdepth;
    McItemDev *const d_items = b->d_items;
    a.cmds = b->d_cmds; a.items = d_items; a.nb_items = b->nb_items;
    a.limit = b->d_info->limit;
    a.lmcs_fwd = lmcs ? lmcs->d_luts->fwd_lut : NULL;
    McGuardArgs ga = { b->d_info, dst->width, dst->height, { 0, 0 }, b->h_info };
    for (int l = 0; l < 2; l++)
        for (int i = 0; i < 16; i++) {
            const OVGPUPicture *p = refs->rpl[l][i];
            a.refs.poc[l][i] = refs->poc[l][i];
            if (!p) continue;
            // a reference of another size needs its scale factor in the batch's table (ovgpu_mc_batch_set_scaling)
            if (!((b->run_scaled[l] >> i) & 1) && (p->width != dst->width || p->height != dst->height))
                return ovgpu_set_error(ctx, OVGPU_EINVAL, "reference picture (list %d, index %d) differs in size and has no scale factor", l, i);
            if (p->bitdepth != dst->bitdepth) return ovgpu_set_error(ctx, OVGPU_EINVAL, "reference picture bit depth differs");
            ga.present[l] |= 1u << i;
            for (int c = 0; c < 3; c++) { a.refs.plane[l][i][c] = p->plane[c]; a.refs.pitch[l][i][c] = (int)(p->pitch[c] / 2); }
        }
    // McGuardK holds the walk's findings against this picture and these references: a malformed command, a command outside the
    // picture or a missing reference leaves every kernel below with nothing to do
    OVGPU_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, b->ev_built, 0));
    OVGPU_CUDA_CHECK(ctx, ovd_launch<McGuardK>(ga, dim3(1, 1, 1), 0, ctx->stream));
    ctx->launches += 1;
    StageScope scope(ctx, OVGPU_STAGE_MC, (b->nb_items > 0) + (b->nb_bdof > 0) + (b->nb_prof_items > 0) + (b->nb_dmvr > 0) + (b->nb_small_c > 0) + (b->nb_big > 0) + (b->nb_rpr > 0));
    // The kernels work on disjoint blocks of the picture: they run side by side on the context's side streams (fork / join with
    // events), the longest first, so that the tail of one fills with the others' CTAs instead of leaving SMs idle between launches.
    static const int lanes = getenv("OVGPU_MC_LANES") ? atoi(getenv("OVGPU_MC_LANES")) : 2;      // streams the kernels are dealt over (1: serial)
    const int par = lanes > 1 && ovgpu_ctx_aux(ctx) == 0;
    int lane = 0, nb_used = 0;
    cudaStream_t used[OVGPUCtx::NB_AUX];
    if (par) OVGPU_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    auto next_stream = [&]() -> cudaStream_t {
        const int l = par ? lane++ % (lanes < 1 + OVGPUCtx::NB_AUX ? lanes : 1 + OVGPUCtx::NB_AUX) : 0;
        if (l == 0) return ctx->stream;
        cudaStream_t s = ctx->aux[l - 1];
        if (l > nb_used) { used[nb_used++] = s; cudaStreamWaitEvent(s, ctx->ev_fork, 0); }
        return s;
    };
    if (b->nb_dmvr) {
        DmvrK::Args t;
        t.mc = a; t.mc.items = d_items + 3 * b->max_items; t.mc.nb_items = b->nb_dmvr; t.mc.limit = a.limit + 3; t.out = b->d_dmvr;
        OVGPU_CUDA_CHECK(ctx, ovd_launch<DmvrK>(t, dim3(OVGPU_CEIL_DIV(b->nb_dmvr, DmvrK::WARPS), 1, 1), DmvrK::SMEM_BYTES, next_stream()));
    }
    if (b->nb_bdof) {
        McK::Args t = a;
        t.items = d_items + b->max_items; t.nb_items = b->nb_bdof; t.limit = a.limit + 1;
        OVGPU_CUDA_CHECK(ctx, ovd_launch<BdofK>(t, dim3(OVGPU_CEIL_DIV(b->nb_bdof, BdofK::WARPS), 1, 1), BdofK::SMEM_BYTES, next_stream()));
    }
    if (b->nb_rpr) {
        McRprK::Args t;
        t.mc = a; t.mc.items = d_items + 6 * b->max_items; t.mc.nb_items = b->nb_rpr; t.mc.limit = a.limit + 6; t.prof = b->d_prof;
        for (int l = 0; l < 2; l++)
            for (int i = 0; i < 16; i++) {
                const OVGPUPicture *p = refs->rpl[l][i];
                t.ref_w[l][i] = p ? p->width : dst->width; t.ref_h[l][i] = p ? p->height : dst->height;
                t.scale[l][i][0] = b->run_scaling.scale[l][i][0]; t.scale[l][i][1] = b->run_scaling.scale[l][i][1];
            }
        t.hor_col = b->run_scaling.chroma_hor_col; t.ver_col = b->run_scaling.chroma_ver_col;
        OVGPU_CUDA_CHECK(ctx, ovd_launch<McRprK>(t, dim3(OVGPU_CEIL_DIV(b->nb_rpr, McRprK::WARPS), 1, 1), McRprK::SMEM_BYTES, next_stream()));
    }
    if (b->nb_big) {
        McK::Args t = a;
        t.items = d_items + 5 * b->max_items; t.nb_items = b->nb_big; t.limit = a.limit + 5;
        OVGPU_CUDA_CHECK(ctx, ovd_launch<McBigK>(t, dim3(OVGPU_CEIL_DIV(b->nb_big, McBigK::WARPS), 1, 1), McBigK::SMEM_BYTES, next_stream()));
    }
    if (b->nb_items)
        OVGPU_CUDA_CHECK(ctx, ovd_launch<McK>(a, dim3(OVGPU_CEIL_DIV(b->nb_items, McK::WARPS), 1, 1), McK::SMEM_BYTES, next_stream()));
    if (b->nb_prof_items) {
        ProfK::Args t;
        t.mc = a; t.mc.items = d_items + 2 * b->max_items; t.mc.nb_items = b->nb_prof_items; t.mc.limit = a.limit + 2; t.prof = b->d_prof;
        OVGPU_CUDA_CHECK(ctx, ovd_launch<ProfK>(t, dim3(OVGPU_CEIL_DIV(b->nb_prof_items, ProfK::WARPS * ProfK::SBS), 1, 1), ProfK::SMEM_BYTES, next_stream()));
    }
    if (b->nb_small_c) {
        McK::Args t = a;
        t.items = d_items + 4 * b->max_items; t.nb_items = b->nb_small_c; t.limit = a.limit + 4;
        OVGPU_CUDA_CHECK(ctx, ovd_launch<McSmallCK>(t, dim3(OVGPU_CEIL_DIV(b->nb_small_c, McSmallCK::WARPS * McSmallCK::SBS), 1, 1), McSmallCK::SMEM_BYTES, next_stream()));
    }
    for (int i = 0; i < nb_used; i++) {
        OVGPU_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev_join[i], used[i]));
        OVGPU_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[i], 0));
    }
    OVGPU_CUDA_CHECK(ctx, cudaEventRecord(b->ev_status, ctx->stream));       // (McGuardK has stored the verdict in h_info)
    b->status_pending = 1; b->nb_cmds_run = b->nb_cmds; b->ran = 1;
    return 0;
}

int ovgpu_mc_batch_read_dmvr(OVGPUCtx *ctx, OVGPUMcBatch *b, OVGPUDmvrMv *out, int max)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !b || (max > 0 && !out)) return OVGPU_EINVAL;
    { const int ret = mc_take_status(ctx, b, 1); if (ret) return ret; }        // (also gives the walk's own DMVR count, below)
    const int nb_dmvr = b->nb_cmds ? b->h_info->count[3] : 0;
    const int n = nb_dmvr < max ? nb_dmvr : max;
    if (n > 0) {
        OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(b->h_dmvr, b->d_dmvr, sizeof(DmvrMvDev) * n, cudaMemcpyDeviceToHost, ctx->stream));
        OVGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(out, b->h_dmvr, sizeof(DmvrMvDev) * n);
    }
    return nb_dmvr;
}

int ovgpu_mc_batch_status(OVGPUCtx *ctx, OVGPUMcBatch *b)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !b) return OVGPU_EINVAL;
    return mc_take_status(ctx, b, 1);
}

} // extern "C"

int ovgpu_lmcs_scale_params(const OVGPULmcs *l, const uint16_t **d_wnd_bnd, int *min_idx, int *max_idx_p1, int *crs_offset)
{
    if (!l || !l->has_data) return -1;
    *d_wnd_bnd = l->d_luts->wnd_bnd; *min_idx = l->min_idx; *max_idx_p1 = l->max_idx_p1; *crs_offset = l->crs_offset;
    return 0;
}
