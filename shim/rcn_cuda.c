/* shim/rcn_cuda.c -- binds the B200 reconstruction backend (include/ovvc_b200.h) into OpenVVC's decoder.
 *
 * Plain C, compiled against the reference's own headers and linked with the reference's own objects into build/libovvc.so
 * (shim/Makefile).  The reference's plug-in mechanism is rcn_init_functions() filling struct RCNFunctions and letting back ends
 * overwrite slots (libovvc/rcn.c:148-300, e.g. rcn_init_mc_functions_avx2 at :242-254).  This file adds one more back end:
 *
 *   rcn_init_functions()        the reference's own function is compiled as rcn_init_functions_stock (-D on rcn.c only, no source
 *                               edit); this wrapper calls it and then rcn_init_functions_cuda() when OVVC_RCN_BACKEND=cuda.
 *   rcn_init_functions_cuda()   re-points the COARSE slots (rcn_structures.h:595-693, :464-490, :408-413, :344-350, :333) at
 *                               recorders.  A recorder snapshots what the stock slot would have read from the OVCTUDec into the
 *                               PINNED staging buffers of the C ABI (ovgpu_*_staging) -- it never touches a sample.
 *   picture end                 alf.rcn_alf_filter_line of the last CTU line commits the lists and runs the picture on the GPU:
 *                               ovgpu_recon_picture -> LMCS^-1 -> DBF -> SAO -> ALF, downloads it into the OVFrame planes
 *                               (ovframe.h:84-124), patches the DMVR-refined MVs into the picture's TMVP plane and only then lets
 *                               the CTU lines be reported decoded (dpb.c:1309; slicedec.c is compiled with
 *                               ovdpb_report_decoded_ctu_line redirected here), which is what TMVP of later pictures waits on
 *                               (drv_mvp.c:292).
 *   pp_process_frame()          the output stage (post_proc.c:76, called from ovdec_receive_picture / _drain_picture): post_proc.c
 *                               is compiled with -Dpp_process_frame=pp_process_frame_stock and this file's wrapper runs the
 *                               up-scaling of resampled sequences (-u 1) and the film grain synthesis on the GPU from the
 *                               frame's device mirror; anything it does not take goes to the stock function.
 *
 * The decoder-side bookkeeping that later CPU stages depend on (deblocking edge / bs / QP maps, progress bit fields, MV maps)
 * stays the reference's: transform trees, TUs, ISP sub-trees and the deblocking pre-processing still run through the STOCK
 * control functions, with every sample-touching leaf slot replaced by a no-op or a recorder.
 *
 * 10 bit.  Pictures may change size inside a sequence (reference picture resampling): the per-picture objects follow the picture's
 * size and the slice's scale factors go with the MC commands.  Pictures of several tiles (rect entries; one slice per picture): an
 * entry decoder hands what it recorded to the picture's accumulator when its entry ends (entry threads move on to other pictures), and
 * the one that completes the picture issues it from the accumulated lists -- with the tile grid, since the reference's SAO and ALF
 * treat an entry's edges as picture borders.  Anything else is left to the stock table (the wrapper does not install the recorders). */
#include <pthread.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ovdefs.h"
#include "ovutils.h"
#include "ovlog.h"
#include "ovframe.h"
#include "ovdpb.h"
#include "nvcl_structures.h"
#include "dec_structures.h"
#include "ctudec.h"
#include "slicedec.h"
#include "drv.h"
#include "drv_utils.h"
#include "vcl.h"
#include "rcn.h"
#include "rcn_structures.h"
#include "rcn_lmcs.h"
#include "rcn_alf.h"
#include "dbf_utils.h"
#include "ovdec_internal.h"
#include "ovframepool.h"
#include "post_proc.h"
#include "overror.h"
#include "bitdepth.h"

#include "ovvc_b200.h"

void rcn_init_functions_stock(struct RCNFunctions *rcn_func, uint8_t ict_type, uint8_t lm_chroma_enabled,
                              uint8_t sps_chroma_vertical_collocated_flag, uint8_t lmcs_flag, uint8_t bitdepth);
void rcn_init_functions_cuda(struct RCNFunctions *rcn_func);
int8_t drv_lfnst_mode_l(uint8_t log2_tb_w, uint8_t log2_tb_h, int8_t intra_mode);

/* struct PROFInfo is private to rcn_inter.c (:1128-1134) and drv_affine_mvp.c (:3306); same layout as OVGPUProfInfo */
struct PROFInfo { int16_t dmv_scale_h_0[16], dmv_scale_v_0[16], dmv_scale_h_1[16], dmv_scale_v_1[16]; };
/* private to rcn_transform_tree.c (:51-78) and vcl_transform_unit.c */
struct TBInfo { uint16_t last_pos; uint64_t sig_sb_map; };
struct TUInfo { uint8_t is_sbt, cbf_mask; uint16_t pos_offset; uint8_t tr_skip_mask, cu_mts_flag, cu_mts_idx, lfnst_flag, lfnst_idx; struct TBInfo tb_info[3]; };
struct ISPTUInfo { uint8_t cbf_mask, tr_skip_mask, cu_mts_flag, cu_mts_idx, lfnst_flag, lfnst_idx; struct TBInfo tb_info[4]; };

#define SHIM_MAX_RECS   256
#define SHIM_MAX_FRAMES 96

enum { FRAME_IDLE = 0, FRAME_IN_FLIGHT = 1, FRAME_DONE = 2 };

typedef struct DmvrPatch { uint16_t ctb_x, ctb_y; uint8_t x0, y0, log2_w, log2_h; OVMV mv0, mv1; } DmvrPatch;

/* what the entries of a picture of several tiles have recorded so far (host memory; copied into the staging buffers of the entry
 * decoder that completes the picture) */
typedef struct ShimAcc {
    const OVFrame *frame;           /* the picture being accumulated, NULL when idle */
    int w, h, ctus_done, nb_col, nb_row, dbf_used, failed;
    uint16_t col[24], row[24];      /* where the entries begin (CTUs) */
    OVGPUMcCmd *mc; int nb_mc, cap_mc; OVGPUProfInfo *prof; int nb_prof, cap_prof; DmvrPatch *dmvr; int nb_dmvr, cap_dmvr;
    OVGPUTuCmd *tu; int nb_tu, cap_tu; int16_t *coeffs; int nb_coeffs, cap_coeffs; OVGPULmcsAnchor *anchors; int nb_anchors, cap_anchors;
    OVGPUIntraCmd *icmds; int nb_i, cap_i; OVGPUIspInfo *isp; int nb_isp, cap_isp;
    uint32_t *ver[OVGPU_DBF_NB_MAPS], *hor[OVGPU_DBF_NB_MAPS]; uint8_t *qp[3];
    OVGPUSaoCtu *sao_ctu; OVGPUAlfCtu *alf_ctu; uint8_t *cc_idx[2];
} ShimAcc;

typedef struct ShimFrame {
    void *key;                      /* frame->data[0]: stable for the life of the pool element */
    int w, h;
    OVGPUPicture *dev;              /* the picture as later pictures reference it (post-ALF), resident in HBM */
    int state;
    OVPicture *pic;                 /* from the first deferred line report */
    unsigned long stamp;
    ShimAcc acc;
} ShimFrame;

typedef struct ShimRec {
    OVCTUDec *ctudec;
    struct RCNFunctions stock;      /* the table rcn_init_functions_stock produced for this slice */
    OVGPUCtx *ctx;
    int w, h, log2_ctu, nb_ctu_w, nb_ctu_h;
    uint8_t ict_type, collocated, active, failed, multi;
    const OVGPUSaoCtu *use_sao; const OVGPUAlfCtu *use_alf; const uint8_t *use_cc[2];      /* per-CTU parameters the picture is issued with */
    OVGPUAlfCtu *alf_ctu; uint8_t *cc_idx[2];
    const OVFrame *frame;
    ShimFrame *sf;
    const struct OVLMCSData *lmcs_data;
    /* motion compensation */
    OVGPUMcBatch *mc; OVGPUMcCmd *mc_cmds; OVGPUProfInfo *prof; int nb_mc, cap_mc, nb_prof, cap_prof;
    DmvrPatch *dmvr; int nb_dmvr, cap_dmvr;
    /* residual */
    OVGPUTuBatch *tu; OVGPUTuCmd *tu_cmds; int16_t *coeffs; int nb_tu, cap_tu; size_t nb_coeffs, cap_coeffs;
    OVGPULmcsAnchor *anchors; int nb_anchors, cap_anchors, cur_anchor;
    /* intra */
    OVGPUIntraBatch *intra; OVGPUIntraCmd *icmds; int nb_i, cap_i; OVGPUIspInfo *isp; int nb_isp, cap_isp;
    /* in-loop filters */
    OVGPUDbfParams *dbf; OVGPUDbfMaps maps; int dbf_used;
    OVGPUSaoParams *sao; OVGPUSaoCtu *sao_ctu;
    OVGPUAlfParams *alf;
    OVGPULmcs *lmcs;
    OVGPUPicture *rec, *mid;
    unsigned long pictures;
} ShimRec;

static struct {
    pthread_mutex_t mtx, acc_mtx;
    pthread_cond_t cnd;
    ShimRec *recs[SHIM_MAX_RECS];
    int nb_recs;
    ShimFrame frames[SHIM_MAX_FRAMES];
    unsigned long stamp;
    int enabled, device, verbose;
} g = { PTHREAD_MUTEX_INITIALIZER, PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, { 0 }, 0, { { 0 } }, 0, -1, 0, 0 };

static __thread ShimRec *tls_rec;

/* ------------------------------------------------------------------------------------------------ selection / errors */
int ovshim_backend_enabled(void)
{
    if (g.enabled < 0) {
        const char *e = getenv("OVVC_RCN_BACKEND"), *d = getenv("OVVC_RCN_DEVICE"), *v = getenv("OVVC_RCN_VERBOSE");
        g.device = d ? atoi(d) : 0;
        g.verbose = v ? atoi(v) : 0;
        g.enabled = e && !strcmp(e, "cuda");
        if (g.enabled) ov_log(NULL, OVLOG_INFO, "rcn back end: B200 CUDA (libovvc_b200), device %d\n", g.device);
    }
    return g.enabled;
}

#define SHIM_CHECK(r, call)                                                                                          \
    do {                                                                                                             \
        int rc__ = (call);                                                                                           \
        if (rc__ < 0 && !(r)->failed) {                                                                              \
            (r)->failed = 1;                                                                                         \
            ov_log(NULL, OVLOG_ERROR, "rcn cuda back end: %s failed (%d): %s\n", #call, rc__, ovgpu_last_error((r)->ctx)); \
        }                                                                                                            \
    } while (0)

static void shim_fatal(const char *what)
{
    /* the slots are void: there is no error channel in the table; a picture the back end cannot take is fatal */
    fprintf(stderr, "rcn cuda back end: %s\n", what);
    abort();
}

static ShimRec *rec_of(OVCTUDec *c)
{
    ShimRec *r = tls_rec;
    if (r && r->ctudec == c) return r;
    pthread_mutex_lock(&g.mtx);
    r = NULL;
    for (int i = 0; i < g.nb_recs; i++) if (g.recs[i]->ctudec == c) r = g.recs[i];
    if (!r) {
        if (g.nb_recs == SHIM_MAX_RECS) shim_fatal("too many entry decoders");
        r = calloc(1, sizeof(*r));
        r->ctudec = c;
        r->cur_anchor = -1;
        g.recs[g.nb_recs++] = r;
    }
    pthread_mutex_unlock(&g.mtx);
    tls_rec = r;
    return r;
}

/* ------------------------------------------------------------------------------------------------ frames in HBM */
static ShimFrame *frame_lookup(const OVFrame *f, int create, OVGPUCtx *ctx)
{
    ShimFrame *sf = NULL, *lru = NULL;
    for (int i = 0; i < SHIM_MAX_FRAMES; i++) {
        ShimFrame *e = &g.frames[i];
        if (e->key == f->data[0] && e->w == (int)f->width && e->h == (int)f->height) { sf = e; break; }
        if (e->state != FRAME_IN_FLIGHT && (!lru || !e->key || (lru->key && e->stamp < lru->stamp))) lru = e;
    }
    if (!sf && create) {
        if (!lru) shim_fatal("frame table full");
        sf = lru;
        if (sf->dev && (sf->w != (int)f->width || sf->h != (int)f->height)) { ovgpu_picture_destroy(sf->dev); sf->dev = NULL; }
        if (!sf->dev && ovgpu_picture_create(ctx, (int)f->width, (int)f->height, 10, &sf->dev) < 0) shim_fatal("cannot allocate a device picture");
        sf->key = f->data[0]; sf->w = (int)f->width; sf->h = (int)f->height; sf->state = FRAME_IDLE; sf->pic = NULL;
        /* pin the pool element's planes so that the download is a real asynchronous DMA (ovframepool.c:194-231 allocates them
         * with ov_mallocz); registering twice is harmless, the registration lives as long as the pool element */
        for (int c = 0; c < 3; c++) ovgpu_host_register(ctx, f->data[c], f->size[c]);
    }
    if (sf) sf->stamp = ++g.stamp;
    return sf;
}

/* ------------------------------------------------------------------------------------------------ per-picture objects */
static void rec_grow_mc(ShimRec *r, int need_cmds, int need_prof)
{
    if (need_cmds <= r->cap_mc && need_prof <= r->cap_prof) return;
    int ncap = r->cap_mc, pcap = r->cap_prof;
    while (ncap < need_cmds) ncap = ncap ? 2 * ncap : 4096;
    while (pcap < need_prof) pcap = pcap ? 2 * pcap : 256;
    OVGPUMcBatch *nb = NULL; OVGPUMcCmd *cm; OVGPUProfInfo *pf;
    if (ovgpu_mc_batch_create(r->ctx, ncap, 4 * ncap + 64, &nb) < 0 || ovgpu_mc_batch_staging(r->ctx, nb, &cm, &pf) < 0) shim_fatal("MC batch allocation");
    if (r->mc) {
        memcpy(cm, r->mc_cmds, sizeof(*cm) * r->nb_mc);
        memcpy(pf, r->prof, sizeof(*pf) * r->nb_prof);
        ovgpu_ctx_sync(r->ctx);
        ovgpu_mc_batch_destroy(r->mc);
    }
    r->mc = nb; r->mc_cmds = cm; r->prof = pf; r->cap_mc = ncap; r->cap_prof = pcap;
}

static void rec_grow_tu(ShimRec *r, int need_cmds, size_t need_coeffs)
{
    if (need_cmds <= r->cap_tu && need_coeffs <= r->cap_coeffs) return;
    int ncap = r->cap_tu; size_t ccap = r->cap_coeffs;
    while (ncap < need_cmds) ncap = ncap ? 2 * ncap : 4096;
    while (ccap < need_coeffs) ccap = ccap ? 2 * ccap : ((size_t)1 << 20);
    OVGPUTuBatch *nb = NULL; OVGPUTuCmd *cm; int16_t *co;
    if (ovgpu_tu_batch_create(r->ctx, ncap, ccap, &nb) < 0 || ovgpu_tu_batch_staging(r->ctx, nb, &cm, &co) < 0) shim_fatal("TU batch allocation");
    if (r->tu) {
        memcpy(cm, r->tu_cmds, sizeof(*cm) * r->nb_tu);
        memcpy(co, r->coeffs, sizeof(*co) * r->nb_coeffs);
        ovgpu_ctx_sync(r->ctx);
        ovgpu_tu_batch_destroy(r->tu);
    }
    r->tu = nb; r->tu_cmds = cm; r->coeffs = co; r->cap_tu = ncap; r->cap_coeffs = ccap;
}

#define GROW(ptr, cap, need, type)                                                     \
    do {                                                                               \
        if ((need) > (cap)) {                                                          \
            int nc__ = (cap) ? (cap) : 256;                                            \
            while (nc__ < (need)) nc__ *= 2;                                           \
            (ptr) = (type *)realloc((ptr), sizeof(type) * (size_t)nc__);              \
            if (!(ptr)) shim_fatal("out of memory");                                   \
            (cap) = nc__;                                                              \
        }                                                                              \
    } while (0)

static void acc_begin(ShimAcc *a, const OVFrame *f, int w, int h, int ncw, int nch)
{
    const int w_u = w >> 2, h_u = h >> 2;
    if (a->w != w || a->h != h) {
        for (int k = 0; k < OVGPU_DBF_NB_MAPS; k++) {
            a->ver[k] = realloc(a->ver[k], sizeof(uint32_t) * nch * (w_u + 1)); a->hor[k] = realloc(a->hor[k], sizeof(uint32_t) * ncw * (h_u + 1));
            if (!a->ver[k] || !a->hor[k]) shim_fatal("out of memory");
        }
        for (int k = 0; k < 3; k++) { a->qp[k] = realloc(a->qp[k], (size_t)w_u * h_u); if (!a->qp[k]) shim_fatal("out of memory"); }
        a->sao_ctu = realloc(a->sao_ctu, sizeof(*a->sao_ctu) * ncw * nch); a->alf_ctu = realloc(a->alf_ctu, sizeof(*a->alf_ctu) * ncw * nch);
        a->cc_idx[0] = realloc(a->cc_idx[0], (size_t)ncw * nch); a->cc_idx[1] = realloc(a->cc_idx[1], (size_t)ncw * nch);
        if (!a->sao_ctu || !a->alf_ctu || !a->cc_idx[0] || !a->cc_idx[1]) shim_fatal("out of memory");
        a->w = w; a->h = h;
    }
    a->frame = f;
    a->ctus_done = a->nb_col = a->nb_row = a->dbf_used = a->failed = 0;
    a->nb_mc = a->nb_prof = a->nb_dmvr = a->nb_tu = a->nb_coeffs = a->nb_anchors = a->nb_i = a->nb_isp = 0;
}

static void rec_begin_picture(ShimRec *r, const OVFrame *f, const struct RectEntryInfo *e, int log2_ctu)
{
    OVCTUDec *c = r->ctudec;
    const int w = c->pic_w, h = c->pic_h, ctu = 1 << log2_ctu;
    const int ncw = (w + ctu - 1) >> log2_ctu, nch = (h + ctu - 1) >> log2_ctu;
    r->multi = e->ctb_x || e->ctb_y || e->nb_ctu_w != ncw || e->nb_ctu_h != nch;
    if ((int)f->width != w || (int)f->height != h) shim_fatal("frame size differs from the picture size");
    if (!r->ctx && ovgpu_ctx_create(&r->ctx, g.device, NULL) < 0) shim_fatal("no usable sm_100 device (ovgpu_ctx_create)");
    if (r->w != w || r->h != h || r->log2_ctu != log2_ctu) {
        if (r->rec) { ovgpu_ctx_sync(r->ctx); ovgpu_picture_destroy(r->rec); ovgpu_picture_destroy(r->mid); ovgpu_dbf_params_destroy(r->dbf);
                      ovgpu_sao_params_destroy(r->sao); ovgpu_alf_params_destroy(r->alf); ovgpu_intra_batch_destroy(r->intra); r->intra = NULL; r->cap_i = 0; free(r->icmds); r->icmds = NULL; }
        if (ovgpu_picture_create(r->ctx, w, h, 10, &r->rec) < 0 || ovgpu_picture_create(r->ctx, w, h, 10, &r->mid) < 0 ||
            ovgpu_dbf_params_create(r->ctx, w, h, log2_ctu, &r->dbf) < 0 || ovgpu_sao_params_create(r->ctx, w, h, log2_ctu, &r->sao) < 0 ||
            ovgpu_alf_params_create(r->ctx, w, h, log2_ctu, &r->alf) < 0) shim_fatal("device allocation for a picture failed");
        if (!r->lmcs && ovgpu_lmcs_create(r->ctx, &r->lmcs) < 0) shim_fatal("LMCS object");
        r->sao_ctu = realloc(r->sao_ctu, sizeof(*r->sao_ctu) * ncw * nch);
        r->alf_ctu = realloc(r->alf_ctu, sizeof(*r->alf_ctu) * ncw * nch);
        r->cc_idx[0] = realloc(r->cc_idx[0], (size_t)ncw * nch); r->cc_idx[1] = realloc(r->cc_idx[1], (size_t)ncw * nch);
        if (!r->sao_ctu || !r->alf_ctu || !r->cc_idx[0] || !r->cc_idx[1]) shim_fatal("out of memory");
        r->w = w; r->h = h; r->log2_ctu = log2_ctu;
    }
    r->nb_ctu_w = ncw; r->nb_ctu_h = nch;
    if (!r->mc) rec_grow_mc(r, (w >> 3) * (h >> 3) / 2 + 1024, 256);
    if (!r->tu) rec_grow_tu(r, (w >> 3) * (h >> 3) / 2 + 1024, (size_t)w * h / 2 + 65536);
    /* the staging calls wait for the H2D copies of the previous picture out of these pinned buffers */
    if (ovgpu_mc_batch_staging(r->ctx, r->mc, &r->mc_cmds, &r->prof) < 0 || ovgpu_tu_batch_staging(r->ctx, r->tu, &r->tu_cmds, &r->coeffs) < 0 ||
        ovgpu_dbf_params_staging(r->ctx, r->dbf, &r->maps) < 0) shim_fatal("staging buffers");
    r->nb_mc = r->nb_prof = r->nb_dmvr = r->nb_tu = r->nb_anchors = r->nb_i = r->nb_isp = 0;
    r->nb_coeffs = 0; r->cur_anchor = -1; r->dbf_used = 0; r->failed = 0;
    r->frame = f;
    pthread_mutex_lock(&g.mtx);
    r->sf = frame_lookup(f, 1, r->ctx);
    if (!(r->multi && r->sf->state == FRAME_IN_FLIGHT && r->sf->acc.frame == f)) {         /* else: a later entry of a picture under way */
        r->sf->state = FRAME_IN_FLIGHT;
        r->sf->pic = NULL;
        if (r->multi) acc_begin(&r->sf->acc, f, w, h, ncw, nch);
    }
    pthread_mutex_unlock(&g.mtx);
    r->active = 1;
}

/* ------------------------------------------------------------------------------------------------ motion compensation */
static OVGPUMcCmd *mc_new(ShimRec *r, int x0, int y0, int lw, int lh, int dir, OVMV mv0, OVMV mv1, int ref0, int ref1, int flags, int tool, uint32_t aux)
{
    OVCTUDec *c = r->ctudec;
    /* the fields of a list the block does not use are whatever the caller's stack held: never pass them on */
    if (dir == 1) { memset(&mv1, 0, sizeof(mv1)); ref1 = 0; mv0.bcw_idx_plus1 = 0; }
    if (dir == 2) { memset(&mv0, 0, sizeof(mv0)); ref0 = 0; }
    rec_grow_mc(r, r->nb_mc + 1, r->nb_prof);
    OVGPUMcCmd *m = &r->mc_cmds[r->nb_mc++];
    memset(m, 0, sizeof(*m));
    m->x0 = (int16_t)((c->ctb_x << r->log2_ctu) + x0); m->y0 = (int16_t)((c->ctb_y << r->log2_ctu) + y0);
    m->log2_w = (uint8_t)lw; m->log2_h = (uint8_t)lh; m->inter_dir = (uint8_t)dir;
    m->flags = (uint8_t)(flags | (c->drv_ctx.inter_ctx.prec_amvr == MV_PRECISION_HALF ? OVGPU_MC_HPEL_AMVR : 0));
    m->ref_idx0 = (int8_t)ref0; m->ref_idx1 = (int8_t)ref1;
    m->bcw_idx_plus1 = mv0.bcw_idx_plus1; m->tool = (uint8_t)tool;
    m->mv0_x = mv0.x; m->mv0_y = mv0.y; m->mv1_x = mv1.x; m->mv1_y = mv1.y;
    m->aux = aux;
    return m;
}

static void rec_mcp(OVCTUDec *const c, struct OVBuffInfo dst, int x0, int y0, int lw, int lh, OVMV mv, uint8_t type, uint8_t ref_idx)
{
    ShimRec *r = rec_of(c);
    OVMV z; memset(&z, 0, sizeof(z));
    (void)dst;
    if (type) mc_new(r, x0, y0, lw, lh, 2, z, mv, 0, ref_idx, OVGPU_MC_LUMA | OVGPU_MC_CHROMA, OVGPU_MC_TOOL_PLAIN, 0)->bcw_idx_plus1 = 0;
    else      mc_new(r, x0, y0, lw, lh, 1, mv, z, ref_idx, 0, OVGPU_MC_LUMA | OVGPU_MC_CHROMA, OVGPU_MC_TOOL_PLAIN, 0)->bcw_idx_plus1 = 0;
}

static void rec_mcp_b(OVCTUDec *const c, struct OVBuffInfo dst, struct InterDRVCtx *const ic, const OVPartInfo *const p, const OVMV mv0, const OVMV mv1,
                      unsigned int x0, unsigned int y0, unsigned int lw, unsigned int lh, uint8_t dir, uint8_t r0, uint8_t r1)
{
    (void)dst; (void)ic; (void)p;
    mc_new(rec_of(c), x0, y0, lw, lh, dir, mv0, mv1, r0, r1, OVGPU_MC_LUMA | OVGPU_MC_CHROMA, OVGPU_MC_TOOL_PLAIN, 0);
}
static void rec_mcp_b_l(OVCTUDec *const c, struct OVBuffInfo dst, struct InterDRVCtx *const ic, const OVPartInfo *const p, const OVMV mv0, const OVMV mv1,
                        unsigned int x0, unsigned int y0, unsigned int lw, unsigned int lh, uint8_t dir, uint8_t r0, uint8_t r1)
{
    (void)dst; (void)ic; (void)p;
    mc_new(rec_of(c), x0, y0, lw, lh, dir, mv0, mv1, r0, r1, OVGPU_MC_LUMA, OVGPU_MC_TOOL_PLAIN, 0);
}
static void rec_mcp_b_c(OVCTUDec *const c, struct OVBuffInfo dst, struct InterDRVCtx *const ic, const OVPartInfo *const p, const OVMV mv0, const OVMV mv1,
                        unsigned int x0, unsigned int y0, unsigned int lw, unsigned int lh, uint8_t dir, uint8_t r0, uint8_t r1)
{
    (void)dst; (void)ic; (void)p;
    mc_new(rec_of(c), x0, y0, lw, lh, dir, mv0, mv1, r0, r1, OVGPU_MC_CHROMA, OVGPU_MC_TOOL_PLAIN, 0);
}
static void rec_bdof_mcp_l(OVCTUDec *const c, struct OVBuffInfo dst, uint8_t x0, uint8_t y0, uint8_t lw, uint8_t lh, OVMV mv0, OVMV mv1, uint8_t r0, uint8_t r1)
{
    (void)dst;
    mc_new(rec_of(c), x0, y0, lw, lh, 3, mv0, mv1, r0, r1, OVGPU_MC_LUMA, OVGPU_MC_TOOL_BDOF, 0);
}
static void rec_prof_mcp_b_l(OVCTUDec *const c, struct OVBuffInfo dst, struct InterDRVCtx *const ic, const OVPartInfo *const p, const OVMV mv0, const OVMV mv1,
                             unsigned int x0, unsigned int y0, unsigned int lw, unsigned int lh, uint8_t dir, uint8_t r0, uint8_t r1,
                             uint8_t prof_dir, const struct PROFInfo *const pi)
{
    ShimRec *r = rec_of(c);
    (void)dst; (void)ic; (void)p;
    /* the CU's table is passed again with every 4x4 sub-block: store it once */
    if (!r->nb_prof || memcmp(&r->prof[r->nb_prof - 1], pi, sizeof(*pi))) {
        rec_grow_mc(r, r->nb_mc, r->nb_prof + 1);
        memcpy(&r->prof[r->nb_prof++], pi, sizeof(*pi));
    }
    mc_new(r, x0, y0, lw, lh, dir, mv0, mv1, r0, r1, OVGPU_MC_LUMA, OVGPU_MC_TOOL_PROF, (uint32_t)prof_dir | ((uint32_t)(r->nb_prof - 1) << 8));
}
static uint8_t rec_dmvr_mv_refine(OVCTUDec *const c, struct OVBuffInfo dst, uint8_t x0, uint8_t y0, uint8_t lw, uint8_t lh,
                                  OVMV *mv0, OVMV *mv1, uint8_t r0, uint8_t r1, uint8_t apply_bdof)
{
    ShimRec *r = rec_of(c);
    (void)dst;
    mc_new(r, x0, y0, lw, lh, 3, *mv0, *mv1, r0, r1, OVGPU_MC_LUMA | OVGPU_MC_CHROMA, OVGPU_MC_TOOL_DMVR, apply_bdof ? 1 : 0);
    /* the caller stores *mv0 / *mv1 into the CTU's TMVP array right after this call (vcl_coding_unit.c:2628-2645): the refined
     * vectors only exist once the GPU has run, so the same positions of the picture's MV plane are patched at picture end */
    GROW(r->dmvr, r->cap_dmvr, r->nb_dmvr + 1, DmvrPatch);
    DmvrPatch *d = &r->dmvr[r->nb_dmvr++];
    d->ctb_x = c->ctb_x; d->ctb_y = c->ctb_y; d->x0 = x0; d->y0 = y0; d->log2_w = lw; d->log2_h = lh; d->mv0 = *mv0; d->mv1 = *mv1;
    return 0;
}
static void rec_gpm_b(OVCTUDec *const c, struct VVCGPM *gp, int x0, int y0, int lw, int lh)
{
    OVGPUMcCmd *m = mc_new(rec_of(c), x0, y0, lw, lh, 3, gp->mv0, gp->mv1, gp->mv0.ref_idx, gp->mv1.ref_idx, OVGPU_MC_LUMA | OVGPU_MC_CHROMA,
                           OVGPU_MC_TOOL_GPM, (uint32_t)gp->split_dir | ((uint32_t)(gp->inter_dir0 - 1) << 8) | ((uint32_t)(gp->inter_dir1 - 1) << 9));
    m->bcw_idx_plus1 = 0;
}

/* ------------------------------------------------------------------------------------------------ intra */
static OVGPUIntraCmd *intra_new(ShimRec *r, int x0, int y0, int lw, int lh)
{
    OVCTUDec *c = r->ctudec;
    GROW(r->icmds, r->cap_i, r->nb_i + 1, OVGPUIntraCmd);
    OVGPUIntraCmd *m = &r->icmds[r->nb_i++];
    memset(m, 0, sizeof(*m));
    m->x0 = (int16_t)((c->ctb_x << r->log2_ctu) + x0); m->y0 = (int16_t)((c->ctb_y << r->log2_ctu) + y0);
    m->log2_w = (uint8_t)lw; m->log2_h = (uint8_t)lh;
    return m;
}

static void rec_intra_pred(const struct OVRCNCtx *const rc, const struct OVBuffInfo *buf, uint8_t mode, int x0, int y0, int lw, int lh, CUFlags fl)
{
    OVGPUIntraCmd *m = intra_new(rec_of(rc->ctudec), x0, y0, lw, lh);
    (void)buf;
    m->flags = OVGPU_INTRA_LUMA; m->mode = mode;
    m->avail_cols = rc->progress_field.vfield[x0 >> 2]; m->avail_rows = rc->progress_field.hfield[y0 >> 2];
    if (fl & flg_intra_bdpcm_luma_flag) m->pad[4] = 1 | ((fl & flg_intra_bdpcm_luma_dir) ? 2 : 0);
}
static void rec_intra_pred_mrl(const OVCTUDec *const c, OVSample *const src, ptrdiff_t stride, uint8_t mode, int x0, int y0, int lw, int lh, int mrl_idx)
{
    OVGPUIntraCmd *m = intra_new(rec_of((OVCTUDec *)c), x0, y0, lw, lh);
    (void)src; (void)stride;
    m->flags = OVGPU_INTRA_LUMA; m->mode = mode; m->mrl_idx = (uint8_t)mrl_idx;
    m->avail_cols = c->rcn_ctx.progress_field.vfield[x0 >> 2]; m->avail_rows = c->rcn_ctx.progress_field.hfield[y0 >> 2];
}
static void rec_intra_mip(const struct OVRCNCtx *const rc, uint8_t x0, uint8_t y0, uint8_t lw, uint8_t lh, uint8_t opaque)
{
    OVGPUIntraCmd *m = intra_new(rec_of(rc->ctudec), x0, y0, lw, lh);
    m->flags = OVGPU_INTRA_LUMA | OVGPU_INTRA_MIP; m->mode = opaque;
    m->avail_cols = rc->progress_field.vfield[x0 >> 2]; m->avail_rows = rc->progress_field.hfield[y0 >> 2];
}
/* x0 / y0 / sizes in CHROMA samples (vvc_intra_pred_chroma, rcn_intra.c:1015-1182) */
static void rec_intra_pred_c(const struct OVRCNCtx *const rc, uint8_t mode_c, int x0, int y0, int lw, int lh, CUFlags fl)
{
    ShimRec *r = rec_of(rc->ctudec);
    OVGPUIntraCmd *m = intra_new(r, x0 << 1, y0 << 1, lw + 1, lh + 1);
    m->flags = (uint8_t)(OVGPU_INTRA_CHROMA | (r->collocated ? OVGPU_INTRA_CCLM_COLLOCATED : 0)); m->mode_c = mode_c;
    m->avail_cols = rc->progress_field_c.vfield[x0 >> 1]; m->avail_rows = rc->progress_field_c.hfield[y0 >> 1];
    if (fl & flg_intra_bdpcm_chroma_flag) { m->pad[4] = 4 | ((fl & flg_intra_bdpcm_chroma_dir) ? 8 : 0); return; }
    fill_bs_map(&rc->ctudec->dbf_info.bs2_map_c, x0 << 1, y0 << 1, lw + 1, lh + 1);        /* the stock slot's bookkeeping (:1180) */
}
static void nop_intra_pred_isp(const OVCTUDec *const c, OVSample *const src, ptrdiff_t stride, uint8_t mode, int x0, int y0, int lw, int lh,
                               int lcw, int lch, int ox, int oy)
{
    (void)c; (void)src; (void)stride; (void)mode; (void)x0; (void)y0; (void)lw; (void)lh; (void)lcw; (void)lch; (void)ox; (void)oy;
}

/* rcn_ciip_b / rcn_ciip (rcn_inter.c:3012-3067): inter prediction, planar intra prediction, blend with a weight from the
 * neighbours' modes (rcn_ciip_weighted_sum :2968-2984) */
static void rec_ciip_cmd(ShimRec *r, int x0, int y0, int lw, int lh)
{
    OVCTUDec *c = r->ctudec;
    const int ms = c->part_ctx->log2_min_cb_s;
    const int abv = c->part_map.cu_mode_x[(x0 + (1 << lw) - 1) >> ms], lft = c->part_map.cu_mode_y[(y0 + (1 << lh) - 1) >> ms];
    const int wt = 1 + (abv == OV_INTRA || abv == OV_MIP) + (lft == OV_INTRA || lft == OV_MIP);
    const struct OVRCNCtx *rc = &c->rcn_ctx;
    const uint64_t lc = rc->progress_field.vfield[x0 >> 2], lr = rc->progress_field.hfield[y0 >> 2];
    const uint64_t cc = rc->progress_field_c.vfield[x0 >> 2], cr = rc->progress_field_c.hfield[y0 >> 2];
    OVGPUIntraCmd *m = intra_new(r, x0, y0, lw, lh);
    m->mode = OVINTRA_PLANAR; m->mode_c = OVINTRA_PLANAR; m->ciip_wt = (uint8_t)wt;
    m->avail_cols = lc; m->avail_rows = lr;
    if (lc == cc && lr == cr) { m->flags = OVGPU_INTRA_LUMA | OVGPU_INTRA_CHROMA | OVGPU_INTRA_CIIP; return; }
    m->flags = OVGPU_INTRA_LUMA | OVGPU_INTRA_CIIP;
    m = intra_new(r, x0, y0, lw, lh);
    m->mode_c = OVINTRA_PLANAR; m->ciip_wt = (uint8_t)wt; m->flags = OVGPU_INTRA_CHROMA | OVGPU_INTRA_CIIP;
    m->avail_cols = cc; m->avail_rows = cr;
}
static void rec_ciip_b(OVCTUDec *const c, const OVMV mv0, const OVMV mv1, unsigned int x0, unsigned int y0, unsigned int lw, unsigned int lh,
                       uint8_t dir, uint8_t r0, uint8_t r1)
{
    ShimRec *r = rec_of(c);
    mc_new(r, x0, y0, lw, lh, dir, mv0, mv1, r0, r1, OVGPU_MC_LUMA | OVGPU_MC_CHROMA, OVGPU_MC_TOOL_PLAIN, 0);
    rec_ciip_cmd(r, x0, y0, lw, lh);
}
static void rec_ciip(OVCTUDec *const c, int x0, int y0, int lw, int lh, OVMV mv, uint8_t ref_idx)
{
    ShimRec *r = rec_of(c);
    OVMV z; memset(&z, 0, sizeof(z));
    mc_new(r, x0, y0, lw, lh, 1, mv, z, ref_idx, 0, OVGPU_MC_LUMA | OVGPU_MC_CHROMA, OVGPU_MC_TOOL_PLAIN, 0)->bcw_idx_plus1 = 0;
    rec_ciip_cmd(r, x0, y0, lw, lh);
}

static void rec_ibc(OVCTUDec *const c, int16_t x0, int16_t y0, uint8_t lw, uint8_t lh, IBCMV mv, int flag)
{
    ShimRec *r = rec_of(c);
    OVGPUIntraCmd *m = NULL;
    /* rcn_ibc_c follows rcn_ibc_l for the same block (vcl_coding_unit.c:1060-1062): one command carries both */
    if (flag == OVGPU_INTRA_CHROMA && r->nb_i) {
        OVGPUIntraCmd *p = &r->icmds[r->nb_i - 1];
        if ((p->flags & OVGPU_INTRA_IBC) && p->x0 == (c->ctb_x << r->log2_ctu) + x0 && p->y0 == (c->ctb_y << r->log2_ctu) + y0 && p->log2_w == lw && p->log2_h == lh) m = p;
    }
    if (!m) m = intra_new(r, x0, y0, lw, lh);
    m->flags |= (uint8_t)(OVGPU_INTRA_IBC | flag);
    int16_t bv[2] = { (int16_t)mv.x, (int16_t)mv.y };
    memcpy(m->pad, bv, 4);
}
static void rec_ibc_l(OVCTUDec *const c, int16_t x0, int16_t y0, uint8_t lw, uint8_t lh, uint8_t log2_ctu, IBCMV mv) { (void)log2_ctu; rec_ibc(c, x0, y0, lw, lh, mv, OVGPU_INTRA_LUMA); }
static void rec_ibc_c(OVCTUDec *const c, int16_t x0, int16_t y0, uint8_t lw, uint8_t lh, uint8_t log2_ctu, IBCMV mv) { (void)log2_ctu; rec_ibc(c, x0, y0, lw, lh, mv, OVGPU_INTRA_CHROMA); }

/* ------------------------------------------------------------------------------------------------ residual */
static uint32_t tu_copy_coeffs(ShimRec *r, const int16_t *src, int lw, int lh)
{
    const int tw = 1 << lw, th = 1 << lh;
    const size_t n = (lw >= 2 && lh >= 2) ? (size_t)(tw < 32 ? tw : 32) * (th < 32 ? th : 32) : (size_t)tw * th;
    rec_grow_tu(r, r->nb_tu + 1, r->nb_coeffs + n);
    const size_t off = r->nb_coeffs;
    memcpy(r->coeffs + off, src, n * sizeof(int16_t));
    r->nb_coeffs += n;
    return (uint32_t)off;
}

static void tu_common(ShimRec *r, OVGPUTuCmd *m, CUFlags fl, const struct TUInfo *t)
{
    OVCTUDec *c = r->ctudec;
    m->tr_skip_mask = t->tr_skip_mask; m->cu_mts_flag = t->cu_mts_flag; m->cu_mts_idx = t->cu_mts_idx; m->lfnst_idx = t->lfnst_idx;
    m->qp[0] = c->dequant_luma.qp; m->qp[1] = c->dequant_cb.qp; m->qp[2] = c->dequant_cr.qp; m->qp[3] = c->dequant_joint_cb_cr.qp;
    m->qp_ts[0] = c->dequant_luma_skip.qp; m->qp_ts[1] = c->dequant_cb_skip.qp; m->qp_ts[2] = c->dequant_cr_skip.qp; m->qp_ts[3] = c->dequant_jcbcr_skip.qp;
    for (int k = 0; k < 3; k++) { m->last_pos[k] = t->tb_info[k].last_pos; m->sig_sb_map[k] = t->tb_info[k].sig_sb_map; }
    m->chroma_scale = 1 << 11;
    if (c->lmcs_info.scale_c_flag && r->cur_anchor >= 0) { m->flags |= OVGPU_TU_SCALE_ANCHOR; m->scale_anchor = (uint16_t)r->cur_anchor; }
    m->lfnst_mode_l = (fl & flg_mip_flag) ? OVINTRA_PLANAR : c->intra_mode;
    m->bdpcm = ((fl & flg_intra_bdpcm_luma_flag) ? 1 : 0) | ((fl & flg_intra_bdpcm_luma_dir) ? 2 : 0) |
               ((fl & flg_intra_bdpcm_chroma_flag) ? 4 : 0) | ((fl & flg_intra_bdpcm_chroma_dir) ? 8 : 0);
}

/* chroma mode the LFNST kernel choice follows: DM / CCLM resolved to the luma mode at the centre of the collocated block
 * (derive_lfnst_mode_c, drv_lfnst.c:87-104); x0 / y0 / sizes in chroma samples as the stock function takes them */
static uint8_t lfnst_mode_c(OVCTUDec *c, int x0, int y0, int lw, int lh)
{
    const int ms = c->part_ctx_c->log2_min_cb_s, x_pu = x0 >> ms, y_pu = y0 >> ms, nw = (1 << lw) >> ms, nh = (1 << lh) >> ms;
    const uint8_t luma_mode = c->drv_ctx.intra_info.luma_modes[(x_pu + ((y_pu + (nh >> 1)) << 5) + (nw >> 1))];
    const uint8_t mc = c->intra_mode_c;
    return (mc == OVINTRA_DM_CHROMA || (mc >= OVINTRA_LM_CHROMA && mc <= OVINTRA_MDLM_TOP)) ? luma_mode : mc;
}

static void rec_tu_chroma_part(ShimRec *r, OVGPUTuCmd *m, int cbf, const struct TUInfo *t, int lwc, int lhc)
{
    OVCTUDec *c = r->ctudec;
    if (cbf & 0x8) {
        m->coeff_off[0] = tu_copy_coeffs(r, c->residual_cb + t->pos_offset, lwc, lhc);
    } else {
        if (cbf & 0x2) m->coeff_off[0] = tu_copy_coeffs(r, c->residual_cb + t->pos_offset, lwc, lhc);
        if (cbf & 0x1) m->coeff_off[1] = tu_copy_coeffs(r, c->residual_cr + t->pos_offset, lwc, lhc);
    }
}

/* single tree (rcn_tu_st, rcn_transform_tree.c:1228-1303) and luma-only tree (rcn_tu_l :1305-1347): luma coordinates */
static void rec_tu_luma_tree(ShimRec *r, int x0, int y0, int lw, int lh, CUFlags fl, int cbf, const struct TUInfo *t, int with_chroma)
{
    OVCTUDec *c = r->ctudec;
    if (with_chroma) cbf &= 0x1b;
    else cbf = cbf ? 0x10 : 0;          /* transform_unit_l returns a plain flag (vcl_transform_unit.c:855), rcn_tu_l tests it as such (:1312) */
    if ((cbf & 0x8) && !(cbf & 0x3)) cbf &= ~0x8;
    if (!cbf) return;
    rec_grow_tu(r, r->nb_tu + 1, r->nb_coeffs);
    const int idx = r->nb_tu++;
    OVGPUTuCmd tmp;
    memset(&tmp, 0, sizeof(tmp));
    tmp.x0 = (int16_t)((c->ctb_x << r->log2_ctu) + x0); tmp.y0 = (int16_t)((c->ctb_y << r->log2_ctu) + y0);
    tmp.log2_w = (uint8_t)lw; tmp.log2_h = (uint8_t)lh; tmp.cbf_mask = (uint8_t)cbf;
    const int intra = (fl & flg_pred_mode_flag) != 0, mip = (fl & flg_mip_flag) != 0;
    if (intra && !mip && c->mts_implicit) tmp.flags |= OVGPU_TU_IMPLICIT_MTS;
    if (t->lfnst_flag) tmp.flags |= OVGPU_TU_LFNST_L;
    tu_common(r, &tmp, fl, t);
    if (cbf & 0x10) tmp.coeff_off[2] = tu_copy_coeffs(r, c->residual_y + t->pos_offset, lw, lh);
    rec_tu_chroma_part(r, &tmp, cbf, t, lw - 1, lh - 1);
    r->tu_cmds[idx] = tmp;          /* (the staging buffer may have moved while the coefficients were copied) */
}

/* chroma tree (rcn_tu_c, :1349-1390): chroma coordinates and sizes */
static void rec_tu_chroma_tree(ShimRec *r, int x0, int y0, int lw, int lh, CUFlags fl, int cbf, const struct TUInfo *t)
{
    OVCTUDec *c = r->ctudec;
    cbf &= 0xb;
    if ((cbf & 0x8) && !(cbf & 0x3)) cbf &= ~0x8;
    if (!cbf) return;
    rec_grow_tu(r, r->nb_tu + 1, r->nb_coeffs);
    const int idx = r->nb_tu++;
    OVGPUTuCmd tmp;
    memset(&tmp, 0, sizeof(tmp));
    tmp.x0 = (int16_t)((c->ctb_x << r->log2_ctu) + (x0 << 1)); tmp.y0 = (int16_t)((c->ctb_y << r->log2_ctu) + (y0 << 1));
    tmp.log2_w = (uint8_t)(lw + 1); tmp.log2_h = (uint8_t)(lh + 1); tmp.cbf_mask = (uint8_t)cbf;
    if (t->lfnst_flag) { tmp.flags |= OVGPU_TU_LFNST_C; tmp.lfnst_mode_c = lfnst_mode_c(c, x0, y0, lw, lh); }
    tu_common(r, &tmp, fl, t);
    rec_tu_chroma_part(r, &tmp, cbf, t, lw, lh);
    r->tu_cmds[idx] = tmp;
}

/* the leaves of rcn_transform_tree (:1454-1516): same implicit split of blocks wider / taller than the maximum transform size */
static void walk_transform_tree(ShimRec *r, int x0, int y0, int lw, int lh, int lmax, int depth, CUFlags fl, const struct TUInfo *t)
{
    OVCTUDec *c = r->ctudec;
    const int sv = lw > lmax, sh = lh > lmax, nsub = depth ? 1 : (1 << (sv + sh));
    if (lw > 6 && lh < 7) { walk_transform_tree(r, x0, y0, 6, lh, lmax, depth + 1, fl, &t[0]); walk_transform_tree(r, x0 + 64, y0, 6, lh, lmax, depth + 1, fl, &t[8]); return; }
    if (lh > 6 && lw < 7) { walk_transform_tree(r, x0, y0, lw, 6, lmax, depth + 1, fl, &t[0]); walk_transform_tree(r, x0, y0 + 64, lw, 6, lmax, depth + 1, fl, &t[8]); return; }
    if (sv || sh) {
        const int w1 = (1 << lw) >> sv, h1 = (1 << lh) >> sh, lw1 = lw - sv, lh1 = lh - sh;
        walk_transform_tree(r, x0, y0, lw1, lh1, lmax, depth + 1, fl, &t[0 * nsub]);
        if (sv) walk_transform_tree(r, x0 + w1, y0, lw1, lh1, lmax, depth + 1, fl, &t[1 * nsub]);
        if (sh) walk_transform_tree(r, x0, y0 + h1, lw1, lh1, lmax, depth + 1, fl, &t[2 * nsub]);
        if (sh && sv) walk_transform_tree(r, x0 + w1, y0 + h1, lw1, lh1, lmax, depth + 1, fl, &t[3 * nsub]);
        return;
    }
    if (c->transform_unit == &transform_unit_st)     rec_tu_luma_tree(r, x0, y0, lw, lh, fl, t->cbf_mask, t, 1);
    else if (c->transform_unit == &transform_unit_l) rec_tu_luma_tree(r, x0, y0, lw, lh, fl, t->cbf_mask, t, 0);
    else if (c->transform_unit == &transform_unit_c) rec_tu_chroma_tree(r, x0, y0, lw, lh, fl, t->cbf_mask, t);
}

static void rec_transform_tree(OVCTUDec *const c, uint8_t x0, uint8_t y0, uint8_t lw, uint8_t lh, uint8_t lmax, uint8_t depth, CUFlags fl, const struct TUInfo *const t)
{
    ShimRec *r = rec_of(c);
    walk_transform_tree(r, x0, y0, lw, lh, lmax, depth, fl, t);
    r->stock.tmp.rcn_transform_tree(c, x0, y0, lw, lh, lmax, depth, fl, t);     /* bookkeeping; predictions recorded by the leaf slots */
}
static void rec_tu_st(OVCTUDec *const c, uint8_t x0, uint8_t y0, uint8_t lw, uint8_t lh, CUFlags fl, uint8_t cbf, const struct TUInfo *const t)
{
    ShimRec *r = rec_of(c);
    rec_tu_luma_tree(r, x0, y0, lw, lh, fl, cbf, t, 1);
    r->stock.tmp.rcn_tu_st(c, x0, y0, lw, lh, fl, cbf, t);
}
static void rec_tu_c(OVCTUDec *const c, uint8_t x0, uint8_t y0, uint8_t lw, uint8_t lh, CUFlags fl, uint8_t cbf, const struct TUInfo *const t)
{
    ShimRec *r = rec_of(c);
    rec_tu_chroma_tree(r, x0, y0, lw, lh, fl, cbf, t);
    r->stock.tmp.rcn_tu_c(c, x0, y0, lw, lh, fl, cbf, t);
}

/* ISP coding unit (recon_isp_subtree_{v,h}, rcn_transform_tree.c:1087-1205): one intra command for the whole luma block plus one
 * transform unit per coded partition; geometry and transform types as the stock functions derive them */
static void rec_isp_subtree(OVCTUDec *const c, unsigned int x0, unsigned int y0, unsigned int lcw, unsigned int lch, uint8_t mode,
                            const struct ISPTUInfo *const t, int ver)
{
    ShimRec *r = rec_of(c);
    const struct CTUBitField *pf = &c->rcn_ctx.progress_field;
    int lp = (ver ? (int)lcw : (int)lch) - 2;
    const int other = ver ? (int)lch : (int)lcw;
    if (other < 4 && lp <= 4 - other) lp = 4 - other;
    const int nb = (1 << (ver ? lcw : lch)) >> lp;
    const int ltw = ver ? lp : (int)lcw, lth = ver ? (int)lch : lp;
    const int type_h = ver ? (c->mts_enabled && lp <= 4 && lp > 1) : (c->mts_enabled && lcw <= 4);
    const int type_v = ver ? (c->mts_enabled && lch <= 4) : (c->mts_enabled && lp <= 4 && lp > 1);
    GROW(r->isp, r->cap_isp, r->nb_isp + 1, OVGPUIspInfo);
    OVGPUIspInfo *info = &r->isp[r->nb_isp];
    memset(info, 0, sizeof(*info));
    info->common = ver ? pf->hfield[y0 >> 2] : pf->vfield[x0 >> 2];
    OVGPUIntraCmd *m = intra_new(r, x0, y0, lcw, lch);
    m->flags = (uint8_t)(OVGPU_INTRA_LUMA | OVGPU_INTRA_ISP | (ver ? OVGPU_INTRA_ISP_VER : 0)); m->mode = mode;
    m->pad[0] = (uint8_t)(r->nb_isp & 0xff); m->pad[1] = (uint8_t)(r->nb_isp >> 8);
    m->avail_cols = pf->vfield[x0 >> 2]; m->avail_rows = pf->hfield[y0 >> 2];
    r->nb_isp++;
    int q = 0;
    for (int i = 0; i < nb; i++) {
        const int off = i << lp;
        if (ver) { if (!(off & 3)) info->part[q++] = pf->vfield[(x0 + off) >> 2]; }
        else info->part[q++] = pf->hfield[((y0 + off) >> 2) + !!((y0 + off) % 4)];
        if (!((t->cbf_mask >> (nb - i - 1)) & 1)) continue;
        rec_grow_tu(r, r->nb_tu + 1, r->nb_coeffs);
        const int idx = r->nb_tu++;
        OVGPUTuCmd tmp;
        memset(&tmp, 0, sizeof(tmp));
        tmp.x0 = (int16_t)((c->ctb_x << r->log2_ctu) + x0 + (ver ? off : 0)); tmp.y0 = (int16_t)((c->ctb_y << r->log2_ctu) + y0 + (ver ? 0 : off));
        tmp.log2_w = (uint8_t)ltw; tmp.log2_h = (uint8_t)lth; tmp.cbf_mask = 0x10;
        tmp.flags = OVGPU_TU_ISP | (t->lfnst_flag ? OVGPU_TU_LFNST_L : 0);
        tmp.lfnst_idx = t->lfnst_idx; tmp.cu_mts_idx = (uint8_t)((type_h ? 1 : 0) | (type_v ? 2 : 0));
        tmp.qp[0] = c->dequant_luma.qp; tmp.chroma_scale = 1 << 11;
        tmp.last_pos[2] = t->tb_info[i].last_pos; tmp.sig_sb_map[2] = t->tb_info[i].sig_sb_map;
        if (t->lfnst_flag) tmp.lfnst_mode_l = (uint8_t)drv_lfnst_mode_l((uint8_t)lcw, (uint8_t)lch, (int8_t)mode);
        tmp.coeff_off[2] = tu_copy_coeffs(r, c->residual_y + (i << (ltw + lth)), ltw, lth);
        r->tu_cmds[idx] = tmp;
    }
    if (ver) r->stock.tmp.recon_isp_subtree_v(c, x0, y0, lcw, lch, mode, t);
    else     r->stock.tmp.recon_isp_subtree_h(c, x0, y0, lcw, lch, mode, t);
}
static void rec_isp_subtree_v(OVCTUDec *const c, unsigned int x0, unsigned int y0, unsigned int lcw, unsigned int lch, uint8_t mode, const struct ISPTUInfo *const t)
{ rec_isp_subtree(c, x0, y0, lcw, lch, mode, t, 1); }
static void rec_isp_subtree_h(OVCTUDec *const c, unsigned int x0, unsigned int y0, unsigned int lcw, unsigned int lch, uint8_t mode, const struct ISPTUInfo *const t)
{ rec_isp_subtree(c, x0, y0, lcw, lch, mode, t, 0); }

/* no-op leaves of the residual path: the stock TU functions still run for their bookkeeping */
static void nop_tr(const int16_t *src, int16_t *dst, ptrdiff_t s, int nl, int nc, int sh) { (void)src; (void)dst; (void)s; (void)nl; (void)nc; (void)sh; }
static void nop_dc(int16_t *const dst, int lw, int lh, int v) { (void)dst; (void)lw; (void)lh; (void)v; }
static void nop_lfnst(const int16_t *const src, int16_t *const dst, const int8_t *const m, int lw, int lh) { (void)src; (void)dst; (void)m; (void)lw; (void)lh; }
static void nop_add(const int16_t *src, OVSample *dst, int16_t st, int lw, int lh, int sc) { (void)src; (void)dst; (void)st; (void)lw; (void)lh; (void)sc; }
static void nop_dequant(int16_t *dst, const int16_t *src, int scale, int shift, uint8_t lw, uint8_t lh, uint64_t map)
{ (void)dst; (void)src; (void)scale; (void)shift; (void)lw; (void)lh; (void)map; }

/* ------------------------------------------------------------------------------------------------ LMCS */
static void rec_lmcs_chroma_scale(struct LMCSInfo *const li, int16_t stride, const struct CTUBitField *const pf, const OVSample *y, uint8_t x0, uint8_t y0)
{
    /* the value needs RECONSTRUCTED luma (rcn_lmcs.c:320-343): list the position and the availability masks the stock slot reads
     * there; the GPU derives the scale between its luma and chroma passes and every TU recorded from now on refers to it */
    OVCTUDec *c = (OVCTUDec *)((char *)li - offsetof(OVCTUDec, lmcs_info));
    ShimRec *r = rec_of(c);
    (void)stride; (void)y;
    GROW(r->anchors, r->cap_anchors, r->nb_anchors + 1, OVGPULmcsAnchor);
    OVGPULmcsAnchor *a = &r->anchors[r->nb_anchors];
    a->x0 = (int16_t)((c->ctb_x << r->log2_ctu) + x0); a->y0 = (int16_t)((c->ctb_y << r->log2_ctu) + y0);
    a->abv_mask = (uint16_t)((pf->hfield[y0 >> 2] >> ((x0 >> 2) + 1)) & 0xffff);
    a->lft_mask = (uint16_t)((pf->vfield[x0 >> 2] >> ((y0 >> 2) + 1)) & 0xffff);
    r->cur_anchor = r->nb_anchors++;
}
static void rec_init_lmcs(struct LMCSInfo *li, const struct OVLMCSData *const data)
{
    OVCTUDec *c = (OVCTUDec *)((char *)li - offsetof(OVCTUDec, lmcs_info));
    ShimRec *r = rec_of(c);
    r->stock.rcn_init_lmcs(li, data);          /* LUTs are built on the CPU as before (tiny), the GPU gets a copy at picture end */
    r->lmcs_data = data;
}
static void nop_lmcs_reshape(OVSample *dst, ptrdiff_t stride, const struct LMCSLUTs *const luts, int w, int h) { (void)dst; (void)stride; (void)luts; (void)w; (void)h; }

/* ------------------------------------------------------------------------------------------------ deblocking */
/* struct DBFInfo of one CTU -> the picture-level planes of OVGPUDbfMaps (layout in include/ovvc_b200.h) */
static void dbf_record_ctu(ShimRec *r, const struct DBFInfo *d, int cx, int cy, int nuw, int nuh)
{
    OVGPUDbfMaps *m = &r->maps;
    const int n_u = (1 << r->log2_ctu) >> 2, w_u = r->w >> 2, h_u = r->h >> 2;
    const uint64_t vmask = nuh >= 64 ? ~0ull : (1ull << nuh) - 1, hmask = nuw >= 64 ? ~0ull : (1ull << nuw) - 1;
    const uint64_t *vsrc[OVGPU_DBF_NB_MAPS] = { d->ctb_bound_ver + 8, d->aff_edg_ver + 8, d->bs1_map.ver, d->bs2_map.ver, d->affine_map.ver + 1,
                                                 d->ctb_bound_ver_c + 8, d->bs1_map_cb.ver, d->bs1_map_cr.ver, d->bs2_map_c.ver };
    const uint64_t *hsrc[OVGPU_DBF_NB_MAPS] = { d->ctb_bound_hor + 8, d->aff_edg_hor + 8, d->bs1_map.hor, d->bs2_map.hor, d->affine_map.hor + 1,
                                                 d->ctb_bound_hor_c + 8, d->bs1_map_cb.hor, d->bs1_map_cr.hor, d->bs2_map_c.hor };
    for (int k = 0; k < OVGPU_DBF_NB_MAPS; k++) {
        /* column nuw / row nuh (the far CTU border) belongs to the next CTU's record, except at the picture border */
        int last_c = (cx * n_u + nuw == w_u) ? nuw + 1 : nuw, last_r = (cy * n_u + nuh == h_u) ? nuh + 1 : nuh;
        if (k == OVGPU_DBF_AFFINE) { last_c = nuw; last_r = nuh; }
        uint32_t *ver = (uint32_t *)m->ver[k], *hor = (uint32_t *)m->hor[k];
        for (int i = 0; i < last_c; i++) ver[cy * (w_u + 1) + cx * n_u + i] = (uint32_t)(vsrc[k][i] & vmask);
        for (int i = 0; i < last_r; i++) hor[cx * (h_u + 1) + cy * n_u + i] = (uint32_t)((hsrc[k][i] >> 2) & hmask);
    }
    uint8_t *qy = (uint8_t *)m->qp[0], *qb = (uint8_t *)m->qp[1], *qr = (uint8_t *)m->qp[2];
    for (int y = 0; y < nuh; y++) {
        const size_t o = (size_t)(cy * n_u + y) * w_u + cx * n_u;
        memcpy(qy + o, &d->qp_map_y.hor[34 * (1 + y) + 2], nuw);
        memcpy(qb + o, &d->qp_map_cb.hor[34 * (1 + y) + 2], nuw);
        memcpy(qr + o, &d->qp_map_cr.hor[34 * (1 + y) + 2], nuw);
    }
}
static void rec_dbf_ctu(const struct OVRCNCtx *const rc, struct DBFInfo *const d, uint8_t log2_ctu, uint8_t last_x, uint8_t last_y)
{
    OVCTUDec *c = rc->ctudec;
    ShimRec *r = rec_of(c);
    d->disable_h = d->disable_v = 1;            /* the stock slot then only runs dbf_ctu_preproc_{v,h}: bs from motion (rcn_df.c:2180-2183) */
    r->stock.df.rcn_dbf_ctu(rc, d, log2_ctu, last_x, last_y);
    dbf_record_ctu(r, d, c->ctb_x, c->ctb_y, (1 << log2_ctu) >> 2, (1 << log2_ctu) >> 2);
    r->dbf_used = 1;
}
static void rec_dbf_truncated_ctu(const struct OVRCNCtx *const rc, struct DBFInfo *const d, uint8_t log2_ctu, uint8_t last_x, uint8_t last_y, uint8_t ctu_w, uint8_t ctu_h)
{
    OVCTUDec *c = rc->ctudec;
    ShimRec *r = rec_of(c);
    d->disable_h = d->disable_v = 1;
    r->stock.df.rcn_dbf_truncated_ctu(rc, d, log2_ctu, last_x, last_y, ctu_w, ctu_h);
    dbf_record_ctu(r, d, c->ctb_x, c->ctb_y, ctu_w >> 2, ctu_h >> 2);
    r->dbf_used = 1;
}

/* ------------------------------------------------------------------------------------------------ CTU plumbing -> no-ops */
static void nop_ctx_u8(const struct OVRCNCtx *const rc, uint8_t a) { (void)rc; (void)a; }
static void nop_ctx_u8_m(struct OVRCNCtx *rc, uint8_t a) { (void)rc; (void)a; }
static void nop_ctx_int_u8(const struct OVRCNCtx *const rc, int x, uint8_t a) { (void)rc; (void)x; (void)a; }
static void nop_ctx_int_int(const struct OVRCNCtx *const rc, int a, int b) { (void)rc; (void)a; (void)b; }
static void nop_filter_line(OVCTUDec *const c, const struct RectEntryInfo *const e, uint16_t y) { (void)c; (void)e; (void)y; }

static void rec_attach_frame_buff(struct OVRCNCtx *const rc, const OVFrame *const f, const struct RectEntryInfo *const e, uint8_t log2_ctu)
{
    ShimRec *r = rec_of(rc->ctudec);
    r->stock.rcn_attach_frame_buff(rc, f, e, log2_ctu);
    rec_begin_picture(r, f, e, log2_ctu);
}

/* ------------------------------------------------------------------------------------------------ picture end */
static void patch_dmvr(ShimRec *r)
{
    OVCTUDec *c = r->ctudec;
    struct InterDRVCtx *ic = &c->drv_ctx.inter_ctx;
    const struct MVPlane *pl[2] = { ic->tmvp_ctx.plane0, ic->tmvp_ctx.plane1 };
    if (!r->nb_dmvr) return;
    OVGPUDmvrMv *out = malloc(sizeof(*out) * r->nb_dmvr);
    const int n = ovgpu_mc_batch_read_dmvr(r->ctx, r->mc, out, r->nb_dmvr);
    if (n != r->nb_dmvr) { if (!r->failed) ov_log(NULL, OVLOG_ERROR, "rcn cuda back end: %d refined MVs for %d DMVR blocks\n", n, r->nb_dmvr); free(out); return; }
    /* same positions as the CTU-local writes of vcl_coding_unit.c:2633-2645, moved to the picture's plane the way tmvp_store_mv
     * lays it out (drv_lines.c:271-330) */
    const int nb_tmvp_unit = (1 << r->log2_ctu) >> 3, pln_stride = nb_tmvp_unit * c->nb_ctb_pic_w;
    for (int k = 0; k < r->nb_dmvr; k++) {
        const DmvrPatch *d = &r->dmvr[k];
        const int base = (d->ctb_x + d->ctb_y * pln_stride) * nb_tmvp_unit;
        const int tx = (d->x0 + 7) >> 3, ty = (d->y0 + 7) >> 3;
        OVMV m0 = d->mv0, m1 = d->mv1;
        m0.x = out[k].mv0_x; m0.y = out[k].mv0_y; m1.x = out[k].mv1_x; m1.y = out[k].mv1_y;
        for (int l = 0; l < 2; l++) {
            if (!pl[l] || !pl[l]->mvs) continue;
            OVMV *dst = pl[l]->mvs + base + ty * pln_stride + tx;
            const OVMV v = l ? m1 : m0;
            dst[0] = v;
            if (d->log2_w > 3) dst[1] = v;
            if (d->log2_h > 3) { dst[pln_stride] = v; if (d->log2_w > 3) dst[pln_stride + 1] = v; }
        }
    }
    free(out);
}

static void wait_reference(ShimRec *r, const OVPicture *p, const OVGPUPicture **dev)
{
    *dev = NULL;
    if (!p || !p->frame) return;
    pthread_mutex_lock(&g.mtx);
    ShimFrame *sf;
    /* picture-level form of rcn_inter_synchronization (rcn_inter.c:131-146): the reference must be complete in HBM (looked up again
     * after every wait: a table entry belongs to a frame only as long as its key says so) */
    while ((sf = frame_lookup(p->frame, 0, r->ctx)) && sf->state == FRAME_IN_FLIGHT) pthread_cond_wait(&g.cnd, &g.mtx);
    if (sf && sf->state == FRAME_DONE) *dev = sf->dev;
    pthread_mutex_unlock(&g.mtx);
    if (!*dev && !r->failed) { r->failed = 1; ov_log(NULL, OVLOG_ERROR, "rcn cuda back end: reference picture POC %d is not resident on the device\n", p->poc); }
}

/* debugging aid: OVVC_RCN_DUMP=<prefix> writes the recorded lists of every picture to <prefix>.<n>.{mc,tu,coef,intra,isp,anchor} */
static void rec_dump(ShimRec *r)
{
    const char *p = getenv("OVVC_RCN_DUMP");
    if (!p) return;
    const struct { const char *ext; const void *ptr; size_t bytes; } d[] = {
        { "mc", r->mc_cmds, sizeof(OVGPUMcCmd) * r->nb_mc }, { "tu", r->tu_cmds, sizeof(OVGPUTuCmd) * r->nb_tu }, { "coef", r->coeffs, 2 * r->nb_coeffs },
        { "intra", r->icmds, sizeof(OVGPUIntraCmd) * r->nb_i }, { "isp", r->isp, sizeof(OVGPUIspInfo) * r->nb_isp },
        { "anchor", r->anchors, sizeof(OVGPULmcsAnchor) * r->nb_anchors }, { "prof", r->prof, sizeof(OVGPUProfInfo) * r->nb_prof } };
    for (unsigned i = 0; i < sizeof(d) / sizeof(d[0]); i++) {
        char name[512];
        snprintf(name, sizeof(name), "%s.%d.%s", p, (int)r->frame->poc, d[i].ext);
        FILE *f = fopen(name, "wb");
        if (f) { if (d[i].bytes) fwrite(d[i].ptr, 1, d[i].bytes, f); fclose(f); }
    }
}

/* the entry's per-CTU SAO / ALF parameters (arrays of the entry decoder, indexed inside the entry: rcn_sao.c:221, rcn_alf.c:1319) ->
 * picture-level arrays in raster order */
static void entry_ctu_params(ShimRec *r, const struct RectEntryInfo *e, OVGPUSaoCtu *sao, OVGPUAlfCtu *alf, uint8_t *const cc[2])
{
    OVCTUDec *c = r->ctudec;
    const struct ALFInfo *ai = &c->alf_info;
    const int ncw = r->nb_ctu_w;
    const int sao_on = c->sao_info.sao_luma_flag || c->sao_info.sao_chroma_flag;
    const int alf_on = ai->alf_luma_enabled_flag || ai->alf_cb_enabled_flag || ai->alf_cr_enabled_flag;
    const uint8_t en = (uint8_t)((ai->alf_luma_enabled_flag ? 4 : 0) | (ai->alf_cb_enabled_flag ? 2 : 0) | (ai->alf_cr_enabled_flag ? 1 : 0));
    for (int y = 0; y < e->nb_ctu_h; y++) for (int x = 0; x < e->nb_ctu_w; x++) {
        const int i = y * e->nb_ctu_w + x, o = (e->ctb_y + y) * ncw + e->ctb_x + x;
        if (sao_on && c->sao_info.sao_params) {
            const SAOParamsCtu *s = &c->sao_info.sao_params[i];
            OVGPUSaoCtu *d = &sao[o];
            memcpy(d->offset_val, s->offset_val, sizeof(d->offset_val));
            for (int k = 0; k < 3; k++) {
                const int on = k ? c->sao_info.sao_chroma_flag : c->sao_info.sao_luma_flag;
                d->band_position[k] = s->band_position[k]; d->eo_class[k] = s->eo_class[k]; d->type_idx[k] = on ? s->type_idx[k] : 0;
            }
            d->pad = 0;
        }
        if (alf_on && ai->ctb_alf_params) {
            const ALFParamsCtu *s = &ai->ctb_alf_params[i];
            alf[o].ctb_alf_flag = s->ctb_alf_flag & en; alf[o].ctb_alf_idx = s->ctb_alf_idx;
            alf[o].cb_alternative = s->cb_alternative; alf[o].cr_alternative = s->cr_alternative;
            if (ai->cc_alf_cb_enabled_flag && ai->ctb_cc_alf_filter_idx[0]) cc[0][o] = ai->ctb_cc_alf_filter_idx[0][i];
            if (ai->cc_alf_cr_enabled_flag && ai->ctb_cc_alf_filter_idx[1]) cc[1][o] = ai->ctb_cc_alf_filter_idx[1][i];
        }
    }
}

#define ACC_APPEND(dst, nb, cap, src, n, type)                            \
    do {                                                                  \
        GROW(dst, cap, (nb) + (n) + 1, type);                             \
        if (n) memcpy((dst) + (nb), (src), sizeof(type) * (size_t)(n));   \
    } while (0)

/* an entry of a picture of several tiles is over: its lists go to the picture's accumulator (indices into the lists re-based), its
 * CTUs' deblocking maps and filter parameters to the picture-level planes.  Returns 1 when the picture is complete.  g.mtx held. */
static int acc_take_entry(ShimAcc *a, ShimRec *r, const struct RectEntryInfo *e)
{
    const int n_u = (1 << r->log2_ctu) >> 2, w_u = r->w >> 2, h_u = r->h >> 2, ncw = r->nb_ctu_w, nch = r->nb_ctu_h;
    ACC_APPEND(a->mc, a->nb_mc, a->cap_mc, r->mc_cmds, r->nb_mc, OVGPUMcCmd);
    for (int i = 0; i < r->nb_mc; i++) if (a->mc[a->nb_mc + i].tool == OVGPU_MC_TOOL_PROF) a->mc[a->nb_mc + i].aux += (uint32_t)a->nb_prof << 8;
    ACC_APPEND(a->prof, a->nb_prof, a->cap_prof, r->prof, r->nb_prof, OVGPUProfInfo);
    ACC_APPEND(a->dmvr, a->nb_dmvr, a->cap_dmvr, r->dmvr, r->nb_dmvr, DmvrPatch);
    ACC_APPEND(a->tu, a->nb_tu, a->cap_tu, r->tu_cmds, r->nb_tu, OVGPUTuCmd);
    for (int i = 0; i < r->nb_tu; i++) {
        OVGPUTuCmd *m = &a->tu[a->nb_tu + i];
        for (int k = 0; k < 3; k++) m->coeff_off[k] += (uint32_t)a->nb_coeffs;
        if (m->flags & OVGPU_TU_SCALE_ANCHOR) m->scale_anchor = (uint16_t)(m->scale_anchor + a->nb_anchors);
    }
    ACC_APPEND(a->coeffs, a->nb_coeffs, a->cap_coeffs, r->coeffs, (int)r->nb_coeffs, int16_t);
    ACC_APPEND(a->anchors, a->nb_anchors, a->cap_anchors, r->anchors, r->nb_anchors, OVGPULmcsAnchor);
    ACC_APPEND(a->icmds, a->nb_i, a->cap_i, r->icmds, r->nb_i, OVGPUIntraCmd);
    for (int i = 0; i < r->nb_i; i++) {
        OVGPUIntraCmd *m = &a->icmds[a->nb_i + i];
        if (m->flags & OVGPU_INTRA_ISP) { const int idx = (m->pad[0] | (m->pad[1] << 8)) + a->nb_isp; m->pad[0] = (uint8_t)(idx & 0xff); m->pad[1] = (uint8_t)(idx >> 8); }
    }
    ACC_APPEND(a->isp, a->nb_isp, a->cap_isp, r->isp, r->nb_isp, OVGPUIspInfo);
    a->nb_mc += r->nb_mc; a->nb_prof += r->nb_prof; a->nb_dmvr += r->nb_dmvr; a->nb_tu += r->nb_tu; a->nb_coeffs += (int)r->nb_coeffs;
    a->nb_anchors += r->nb_anchors; a->nb_i += r->nb_i; a->nb_isp += r->nb_isp;
    a->dbf_used |= r->dbf_used; a->failed |= r->failed;
    /* deblocking: the words of the entry's CTUs (dbf_record_ctu's layout; the far border column / row only at the picture's border) */
    const int cx0 = e->ctb_x, cx1 = e->ctb_x + e->nb_ctu_w, cy0 = e->ctb_y, cy1 = e->ctb_y + e->nb_ctu_h;
    const int nv = cx1 == ncw ? w_u + 1 - cx0 * n_u : (cx1 - cx0) * n_u, nh = cy1 == nch ? h_u + 1 - cy0 * n_u : (cy1 - cy0) * n_u;
    for (int k = 0; k < OVGPU_DBF_NB_MAPS; k++) {
        for (int cy = cy0; cy < cy1; cy++) { const size_t o = (size_t)cy * (w_u + 1) + cx0 * n_u; memcpy(a->ver[k] + o, r->maps.ver[k] + o, sizeof(uint32_t) * nv); }
        for (int cx = cx0; cx < cx1; cx++) { const size_t o = (size_t)cx * (h_u + 1) + cy0 * n_u; memcpy(a->hor[k] + o, r->maps.hor[k] + o, sizeof(uint32_t) * nh); }
    }
    /* the reference deblocks entry by entry with the neighbour flags of the entry (ctu_lft / ctu_abv in rcn_dbf_ctu, rcn_df.c:2176-
     * 2177): the edges that lie on an entry's left and upper border are never filtered, whatever the maps say.  Their boundary
     * strength goes to 0; the edge bits stay, the filter lengths of the neighbouring edges are derived from them */
    for (int k = 0; k < OVGPU_DBF_NB_MAPS; k++) {
        if (k != OVGPU_DBF_BS1 && k != OVGPU_DBF_BS2 && k != OVGPU_DBF_BS1_CB && k != OVGPU_DBF_BS1_CR && k != OVGPU_DBF_BS2_C) continue;
        if (cx0) for (int cy = cy0; cy < cy1; cy++) a->ver[k][(size_t)cy * (w_u + 1) + cx0 * n_u] = 0;
        if (cy0) for (int cx = cx0; cx < cx1; cx++) a->hor[k][(size_t)cx * (h_u + 1) + cy0 * n_u] = 0;
    }
    const int x_u0 = cx0 * n_u, x_u1 = cx1 * n_u < w_u ? cx1 * n_u : w_u, y_u1 = cy1 * n_u < h_u ? cy1 * n_u : h_u;
    for (int k = 0; k < 3; k++) for (int y = cy0 * n_u; y < y_u1; y++) memcpy(a->qp[k] + (size_t)y * w_u + x_u0, r->maps.qp[k] + (size_t)y * w_u + x_u0, x_u1 - x_u0);
    entry_ctu_params(r, e, a->sao_ctu, a->alf_ctu, a->cc_idx);
    /* the tile grid, from where the entries begin */
    int i;
    for (i = 0; i < a->nb_col && a->col[i] != cx0; i++) ;
    if (i == a->nb_col && a->nb_col < 21) a->col[a->nb_col++] = (uint16_t)cx0;
    for (i = 0; i < a->nb_row && a->row[i] != cy0; i++) ;
    if (i == a->nb_row && a->nb_row < 21) a->row[a->nb_row++] = (uint16_t)cy0;
    a->ctus_done += e->nb_ctu_w * e->nb_ctu_h;
    return a->ctus_done == ncw * nch;
}

static int cmp_u16(const void *a, const void *b) { return (int)*(const uint16_t *)a - (int)*(const uint16_t *)b; }

/* the accumulated picture -> the staging buffers of the entry decoder that completed it */
static void acc_load(ShimAcc *a, ShimRec *r, OVGPUTileGrid *grid)
{
    const int w_u = r->w >> 2, h_u = r->h >> 2, ncw = r->nb_ctu_w, nch = r->nb_ctu_h;
    r->nb_mc = r->nb_prof = r->nb_tu = 0; r->nb_coeffs = 0;
    rec_grow_mc(r, a->nb_mc + 1, a->nb_prof + 1);
    rec_grow_tu(r, a->nb_tu + 1, (size_t)a->nb_coeffs + 1);
    memcpy(r->mc_cmds, a->mc, sizeof(OVGPUMcCmd) * a->nb_mc); memcpy(r->prof, a->prof, sizeof(OVGPUProfInfo) * a->nb_prof);
    memcpy(r->tu_cmds, a->tu, sizeof(OVGPUTuCmd) * a->nb_tu); memcpy(r->coeffs, a->coeffs, sizeof(int16_t) * (size_t)a->nb_coeffs);
    GROW(r->dmvr, r->cap_dmvr, a->nb_dmvr + 1, DmvrPatch); memcpy(r->dmvr, a->dmvr, sizeof(DmvrPatch) * a->nb_dmvr);
    GROW(r->anchors, r->cap_anchors, a->nb_anchors + 1, OVGPULmcsAnchor); memcpy(r->anchors, a->anchors, sizeof(OVGPULmcsAnchor) * a->nb_anchors);
    GROW(r->icmds, r->cap_i, a->nb_i + 1, OVGPUIntraCmd); memcpy(r->icmds, a->icmds, sizeof(OVGPUIntraCmd) * a->nb_i);
    GROW(r->isp, r->cap_isp, a->nb_isp + 1, OVGPUIspInfo); memcpy(r->isp, a->isp, sizeof(OVGPUIspInfo) * a->nb_isp);
    r->nb_mc = a->nb_mc; r->nb_prof = a->nb_prof; r->nb_dmvr = a->nb_dmvr; r->nb_tu = a->nb_tu; r->nb_coeffs = (size_t)a->nb_coeffs;
    r->nb_anchors = a->nb_anchors; r->nb_i = a->nb_i; r->nb_isp = a->nb_isp; r->dbf_used = a->dbf_used; r->failed |= a->failed;
    for (int k = 0; k < OVGPU_DBF_NB_MAPS; k++) {
        memcpy((uint32_t *)r->maps.ver[k], a->ver[k], sizeof(uint32_t) * nch * (w_u + 1));
        memcpy((uint32_t *)r->maps.hor[k], a->hor[k], sizeof(uint32_t) * ncw * (h_u + 1));
    }
    for (int k = 0; k < 3; k++) memcpy((uint8_t *)r->maps.qp[k], a->qp[k], (size_t)w_u * h_u);
    r->use_sao = a->sao_ctu; r->use_alf = a->alf_ctu; r->use_cc[0] = a->cc_idx[0]; r->use_cc[1] = a->cc_idx[1];
    qsort(a->col, a->nb_col, sizeof(uint16_t), cmp_u16); qsort(a->row, a->nb_row, sizeof(uint16_t), cmp_u16);
    memset(grid, 0, sizeof(*grid));
    grid->nb_cols = a->nb_col; grid->nb_rows = a->nb_row;
    for (int i = 0; i < a->nb_col; i++) grid->col_bd[i] = a->col[i];
    for (int i = 0; i < a->nb_row; i++) grid->row_bd[i] = a->row[i];
    grid->col_bd[a->nb_col] = (uint16_t)ncw; grid->row_bd[a->nb_row] = (uint16_t)nch;
}

static void rec_end_picture(ShimRec *r, const OVGPUTileGrid *grid)
{
    OVCTUDec *c = r->ctudec;
    struct InterDRVCtx *ic = &c->drv_ctx.inter_ctx;
    const int ncw = r->nb_ctu_w, nch = r->nb_ctu_h;
    OVGPUCtx *x = r->ctx;
    OVGPUMcRefs refs;
    memset(&refs, 0, sizeof(refs));
    rec_dump(r);
    if (r->nb_mc) {
        for (int i = 0; i < ic->nb_active_ref0 && i < 16; i++) { wait_reference(r, ic->rpl0[i], &refs.rpl[0][i]); refs.poc[0][i] = ic->rpl0[i] ? ic->rpl0[i]->poc : 0; }
        for (int i = 0; i < ic->nb_active_ref1 && i < 16; i++) { wait_reference(r, ic->rpl1[i], &refs.rpl[1][i]); refs.poc[1][i] = ic->rpl1[i] ? ic->rpl1[i]->poc : 0; }
        for (int l = 0; l < 2; l++) for (int i = 0; i < 16; i++) if (!refs.rpl[l][i]) refs.rpl[l][i] = r->rec;      /* never dereference NULL on the device */
    }
    const int lmcs_on = c->lmcs_info.lmcs_enabled_flag && c->lmcs_info.luts;
    if (lmcs_on) {
        OVGPULmcsData ld;
        memset(&ld, 0, sizeof(ld));
        if (r->lmcs_data) {
            ld.lmcs_min_bin_idx = r->lmcs_data->lmcs_min_bin_idx; ld.lmcs_delta_max_bin_idx = r->lmcs_data->lmcs_delta_max_bin_idx;
            ld.lmcs_delta_cw_prec_minus1 = r->lmcs_data->lmcs_delta_cw_prec_minus1;
            memcpy(ld.lmcs_delta_abs_cw, r->lmcs_data->lmcs_delta_abs_cw, sizeof(ld.lmcs_delta_abs_cw));
            memcpy(ld.lmcs_delta_sign_cw_flag, r->lmcs_data->lmcs_delta_sign_cw_flag, sizeof(ld.lmcs_delta_sign_cw_flag));
            ld.lmcs_delta_abs_crs = r->lmcs_data->lmcs_delta_abs_crs; ld.lmcs_delta_sign_crs_flag = r->lmcs_data->lmcs_delta_sign_crs_flag;
        }
        /* struct LMCSLUTs (rcn_lmcs.c:77-82) at 10 bit == OVGPULmcsLuts */
        SHIM_CHECK(r, ovgpu_lmcs_upload(x, r->lmcs, (const OVGPULmcsLuts *)c->lmcs_info.luts));
        SHIM_CHECK(r, ovgpu_lmcs_set_data(r->lmcs, &ld));
    }
    OVGPUTuSlice sl;
    sl.dep_quant = c->residual_coding_l == &residual_coding_dpq; sl.sh_ts_disabled = c->sh_ts_disabled; sl.ict_type = r->ict_type; sl.pad = 0;
    if (r->nb_mc) {
        /* reference picture resampling: the slice's scale factors (ctudec_compute_refs_scaling, ctudec.c:42-84) route the commands
         * that name a scaled reference to the resampling kernel -- what the stock slots decide call by call (rcn_inter.c:2749-2966) */
        OVGPUMcScaling sc;
        memset(&sc, 0, sizeof(sc));
        sc.chroma_hor_col = 1; sc.chroma_ver_col = r->collocated;
        for (int i = 0; i < ic->nb_active_ref0 && i < 16; i++) {
            sc.scale[0][i][0] = ic->scale_fact_rpl0[i][0]; sc.scale[0][i][1] = ic->scale_fact_rpl0[i][1];
            if (ic->rpl0[i]) { sc.chroma_hor_col = ic->rpl0[i]->scale_info.chroma_hor_col_flag; sc.chroma_ver_col = ic->rpl0[i]->scale_info.chroma_ver_col_flag; }
        }
        for (int i = 0; i < ic->nb_active_ref1 && i < 16; i++) { sc.scale[1][i][0] = ic->scale_fact_rpl1[i][0]; sc.scale[1][i][1] = ic->scale_fact_rpl1[i][1]; }
        SHIM_CHECK(r, ovgpu_mc_batch_set_scaling(x, r->mc, &sc));
        SHIM_CHECK(r, ovgpu_mc_batch_commit(x, r->mc, r->nb_mc, r->nb_prof));
    }
    SHIM_CHECK(r, ovgpu_tu_batch_set_anchors(x, r->tu, r->anchors, c->lmcs_info.scale_c_flag ? r->nb_anchors : 0));
    SHIM_CHECK(r, ovgpu_tu_batch_commit(x, r->tu, r->nb_tu, r->nb_coeffs, &sl));
    if (!r->intra) SHIM_CHECK(r, ovgpu_intra_batch_create(x, (r->w >> 2) * (r->h >> 2) * 2 + 16, &r->intra));
    if (r->nb_i) SHIM_CHECK(r, ovgpu_intra_batch_upload2(x, r->intra, r->icmds, r->nb_i, r->isp, r->nb_isp, r->w, r->h, r->log2_ctu));
    /* the picture as later pictures will reference it lives in the frame's device mirror: the last stage writes there */
    OVGPUPicture *final = r->sf->dev;
    const int sao_on = c->sao_info.sao_luma_flag || c->sao_info.sao_chroma_flag;
    const struct ALFInfo *ai = &c->alf_info;
    const int alf_on = ai->alf_luma_enabled_flag || ai->alf_cb_enabled_flag || ai->alf_cr_enabled_flag;
    OVGPUPicture *rec = (sao_on || alf_on) ? r->rec : final;
    if (!r->failed) SHIM_CHECK(r, ovgpu_recon_picture(x, rec, r->nb_mc ? &refs : NULL, r->nb_mc ? r->mc : NULL, lmcs_on ? r->lmcs : NULL, r->tu, r->nb_i ? r->intra : NULL));
    if (lmcs_on && !r->failed) SHIM_CHECK(r, ovgpu_lmcs_reshape_picture(x, rec, r->lmcs, 0));            /* slicedec.c:745-750 */
    if (r->dbf_used && !r->failed) {
        SHIM_CHECK(r, ovgpu_dbf_params_commit(x, r->dbf, c->dbf_info.beta_offset, c->dbf_info.tc_offset));
        SHIM_CHECK(r, ovgpu_dbf_picture(x, rec, r->dbf));
    }
    OVGPUPicture *cur = rec;
    if (sao_on && !r->failed) {
        OVGPUPicture *dst = alf_on ? r->mid : final;
        SHIM_CHECK(r, ovgpu_sao_params_set_tiles(x, r->sao, grid));
        SHIM_CHECK(r, ovgpu_sao_params_upload(x, r->sao, r->use_sao));
        SHIM_CHECK(r, ovgpu_sao_picture(x, dst, cur, r->sao));
        cur = dst;
    }
    if (alf_on && !r->failed) {
        static __thread OVGPUAlfTables tab;
        memset(&tab, 0, sizeof(tab));
        /* RCNALF (rcn_alf.h:60-66) == the first four members of OVGPUAlfTables */
        memcpy(tab.luma_coeff, ai->rcn_alf.filter_coeff_dec, sizeof(tab.luma_coeff));
        memcpy(tab.luma_clip, ai->rcn_alf.filter_clip_dec, sizeof(tab.luma_clip));
        memcpy(tab.chroma_coeff, ai->rcn_alf.chroma_coeff_final, sizeof(tab.chroma_coeff));
        memcpy(tab.chroma_clip, ai->rcn_alf.chroma_clip_final, sizeof(tab.chroma_clip));
        if (ai->cc_alf_cb_enabled_flag && ai->aps_cc_alf_data_cb) memcpy(tab.cc_coeff[0], ai->aps_cc_alf_data_cb->alf_cc_mapped_coeff[0], sizeof(tab.cc_coeff[0]));
        if (ai->cc_alf_cr_enabled_flag && ai->aps_cc_alf_data_cr) memcpy(tab.cc_coeff[1], ai->aps_cc_alf_data_cr->alf_cc_mapped_coeff[1], sizeof(tab.cc_coeff[1]));
        const uint8_t *cc[2] = { ai->cc_alf_cb_enabled_flag && ai->ctb_cc_alf_filter_idx[0] ? r->use_cc[0] : NULL,
                                 ai->cc_alf_cr_enabled_flag && ai->ctb_cc_alf_filter_idx[1] ? r->use_cc[1] : NULL };
        SHIM_CHECK(r, ovgpu_alf_params_set_tiles(x, r->alf, grid));
        SHIM_CHECK(r, ovgpu_alf_params_upload(x, r->alf, &tab, r->use_alf, cc, ai->cc_alf_cb_enabled_flag && cc[0], ai->cc_alf_cr_enabled_flag && cc[1]));
        SHIM_CHECK(r, ovgpu_alf_picture(x, final, cur, r->alf));
        cur = final;
    }
    /* host-readable planes for ovdec_receive_picture (dectest.c:393-406 writes frame->data) */
    if (!r->failed) SHIM_CHECK(r, ovgpu_picture_download(x, final, (void *const *)r->frame->data, r->frame->linesize));
    SHIM_CHECK(r, ovgpu_ctx_sync(x));
    if (r->nb_mc) SHIM_CHECK(r, ovgpu_mc_batch_status(x, r->mc));       /* the command walk runs on the device: its verdict */
    patch_dmvr(r);
    OVPicture *pic;
    pthread_mutex_lock(&g.mtx);
    r->sf->acc.frame = NULL;
    r->sf->state = FRAME_DONE;
    pic = r->sf->pic;
    pthread_cond_broadcast(&g.cnd);
    pthread_mutex_unlock(&g.mtx);
    /* now the lines may be seen as decoded: TMVP of later pictures (drv_mvp.c:292) reads the patched MV plane */
    /* (several tiles: the reports of the last CTU line of the other bottom entries were deferred too) */
    if (pic) for (int y = 0; y < nch - (grid ? 0 : 1); y++) ovdpb_report_decoded_ctu_line(pic, y, 0, ncw - 1);
    r->active = 0;
    r->pictures++;
    if (g.verbose) fprintf(stderr, "rcn cuda: picture %lu of entry decoder %p: %d PU, %d TU (%zu coeffs), %d intra (%d ISP), %d anchors, %d DMVR, dbf %d sao %d alf %d lmcs %d%s\n",
                           r->pictures, (void *)c, r->nb_mc, r->nb_tu, r->nb_coeffs, r->nb_i, r->nb_isp, r->nb_anchors, r->nb_dmvr, r->dbf_used, sao_on, alf_on, lmcs_on,
                           r->failed ? "  FAILED" : "");
}

static void rec_alf_filter_line(OVCTUDec *const c, const struct RectEntryInfo *const e, uint16_t ctb_y)
{
    ShimRec *r = rec_of(c);
    if (!r->active || ctb_y != e->nb_ctu_h - 1) return;
    if (!r->multi) {
        entry_ctu_params(r, e, r->sao_ctu, r->alf_ctu, r->cc_idx);
        r->use_sao = r->sao_ctu; r->use_alf = r->alf_ctu; r->use_cc[0] = r->cc_idx[0]; r->use_cc[1] = r->cc_idx[1];
        rec_end_picture(r, NULL);
        return;
    }
    ShimAcc *a = &r->sf->acc;
    pthread_mutex_lock(&g.acc_mtx);
    if (g.verbose) fprintf(stderr, "rcn cuda: POC %d entry at CTU (%d, %d) recorded by %p: %d PU, %d TU, %d intra\n", (int)r->frame->poc, e->ctb_x, e->ctb_y, (void *)r->ctudec, r->nb_mc, r->nb_tu, r->nb_i);
    const int complete = acc_take_entry(a, r, e);
    OVGPUTileGrid grid;
    if (complete) acc_load(a, r, &grid);
    pthread_mutex_unlock(&g.acc_mtx);
    if (complete) {
        /* all the picture's entries have handed over, nobody else touches the accumulator until the next picture on this frame */
        rec_end_picture(r, &grid);
    } else r->active = 0;
}

/* slicedec.c is compiled with ovdpb_report_decoded_ctu_line redirected here: a CTU line of a picture the GPU has not produced yet
 * is not decoded, whatever the parser thinks -- the report is made by rec_end_picture once samples and refined MVs exist */
void ovshim_report_decoded_ctu_line(OVPicture *const pic, int y_ctu, int xmin_ctu, int xmax_ctu)
{
    if (ovshim_backend_enabled() && pic && pic->frame) {
        int deferred = 0;
        pthread_mutex_lock(&g.mtx);
        for (int i = 0; i < SHIM_MAX_FRAMES; i++) {
            ShimFrame *sf = &g.frames[i];
            if (sf->key == pic->frame->data[0] && sf->state == FRAME_IN_FLIGHT) { sf->pic = pic; deferred = 1; }
        }
        pthread_mutex_unlock(&g.mtx);
        if (deferred) return;
    }
    ovdpb_report_decoded_ctu_line(pic, y_ctu, xmin_ctu, xmax_ctu);
}

/* ------------------------------------------------------------------------------------------------ the table */
void rcn_init_functions_cuda(struct RCNFunctions *f)
{
    /* motion compensation and its tools (rcn_structures.h:628-693) */
    f->rcn_mcp = &rec_mcp; f->rcn_mcp_b = &rec_mcp_b; f->rcn_mcp_b_l = &rec_mcp_b_l; f->rcn_mcp_b_c = &rec_mcp_b_c;
    f->rcn_bdof_mcp_l = &rec_bdof_mcp_l; f->rcn_prof_mcp_b_l = &rec_prof_mcp_b_l; f->rcn_dmvr_mv_refine = &rec_dmvr_mv_refine;
    f->rcn_gpm_b = &rec_gpm_b; f->rcn_ciip_b = &rec_ciip_b; f->rcn_ciip = &rec_ciip; f->rcn_ibc_l = &rec_ibc_l; f->rcn_ibc_c = &rec_ibc_c;
    /* intra prediction leaves (:570-593, mip :321-324) */
    f->intra_pred = &rec_intra_pred; f->intra_pred_c = &rec_intra_pred_c; f->intra_pred_mrl = &rec_intra_pred_mrl; f->intra_pred_isp = &nop_intra_pred_isp;
    f->mip.rcn_intra_mip = &rec_intra_mip;
    /* transform trees (:464-490): recorded, then the stock control code runs over no-op arithmetic leaves */
    f->tmp.rcn_transform_tree = &rec_transform_tree; f->tmp.rcn_tu_st = &rec_tu_st; f->tmp.rcn_tu_c = &rec_tu_c;
    f->tmp.recon_isp_subtree_v = &rec_isp_subtree_v; f->tmp.recon_isp_subtree_h = &rec_isp_subtree_h;
    for (int t = 0; t < NB_TR_TYPES; t++) for (int s = 0; s < NB_TR_SIZES; s++) f->tr.func[t][s] = &nop_tr;
    f->tr.dc = &nop_dc;
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) f->lfnst.func[a][b] = &nop_lfnst;
    for (int i = 0; i < 7; i++) f->ict.add[i] = &nop_add;
    for (int i = 0; i < 6; i++) for (int k = 0; k < 3; k++) f->ict.ict[i][k] = &nop_add;
    f->tmp.dequant_tb_4x4 = &nop_dequant; f->tmp.dequant_tb_4x4_neg = &nop_dequant;
    /* LMCS (:534-538) */
    f->rcn_lmcs_compute_chroma_scale = &rec_lmcs_chroma_scale; f->rcn_init_lmcs = &rec_init_lmcs;
    f->lmcs_reshape_backward = &nop_lmcs_reshape; f->lmcs_reshape_forward = &nop_lmcs_reshape;
    /* in-loop filters (:408-413, :344-350, :333) */
    f->df.rcn_dbf_ctu = &rec_dbf_ctu; f->df.rcn_dbf_truncated_ctu = &rec_dbf_truncated_ctu;
    f->sao.rcn_sao_filter_line = &nop_filter_line; f->sao.rcn_sao_first_pix_rows = &nop_filter_line;
    f->alf.rcn_alf_filter_line = &rec_alf_filter_line;
    /* CTU-local buffers (:595-626): the picture is reconstructed in HBM, nothing is copied on the CPU */
    f->rcn_write_ctu_to_frame = &nop_ctx_u8; f->rcn_write_ctu_to_frame_border = &nop_ctx_int_int;
    f->rcn_ctu_to_intra_line = &nop_ctx_int_u8; f->rcn_intra_line_to_ctu = &nop_ctx_int_u8; f->rcn_update_ctu_border = &nop_ctx_u8_m;
    f->rcn_attach_frame_buff = &rec_attach_frame_buff;
}

/* ------------------------------------------------------------------------------------------------ output stage */
int pp_process_frame_stock(const OVSEI *sei, OVFrame **frame_p);

static struct {
    pthread_mutex_t mtx;
    OVGPUCtx *ctx;
    OVGPUFilmGrainSynth *fg;
    OVGPUPicture *out; int out_w, out_h;
} pp = { PTHREAD_MUTEX_INITIALIZER, NULL, NULL, NULL, 0, 0 };

/* Takes the place of pp_process_frame (post_proc.c:76-139).  What the stock function does to an output frame -- a second frame
 * from the same pool, film grain into it (fg_grain_apply_pic) and / or the decoded frame resampled into it at the pool's size
 * (pp_sample_rate_conv, which overwrites the grain: post_proc.c:100-128) -- is done from the frame's device mirror, and only the
 * result crosses PCIe.  Frames without a mirror, cancelled film grain, film grain on a frame whose rows are not contiguous (the
 * stock function walks such a frame with the wrong stride) or whose size it reads out of bounds: the stock function. */
int pp_process_frame(const OVSEI *sei, OVFrame **frame_p)
{
    if (!ovshim_backend_enabled() || !sei || !frame_p || !*frame_p || (!sei->sei_fg && !sei->upscale_flag)) return pp_process_frame_stock(sei, frame_p);
    OVFrame *frame = *frame_p;
    const int up = sei->upscale_flag != 0, w = (int)frame->width, h = (int)frame->height;
    if (!up) {
        int ok = !sei->sei_fg->fg_characteristics_cancel_flag && frame->linesize[0] == frame->width * sizeof(uint16_t) &&
                 frame->linesize[1] == (frame->width >> 1) * sizeof(uint16_t) && sei->sei_fg->fg_log2_scale_factor <= 15;
        for (int c = 0; c < 2 && ok; c++) {
            const int cw = w >> c, ch = h >> c;
            if ((cw % 16 && cw % 16 < 8) || (ch % 16 && ch % 16 < 8)) ok = 0;
        }
        if (!ok) return pp_process_frame_stock(sei, frame_p);
    }
    pthread_mutex_lock(&g.mtx);
    ShimFrame *sf = frame_lookup(frame, 0, NULL);
    OVGPUPicture *dev = sf && sf->state == FRAME_DONE ? sf->dev : NULL;
    pthread_mutex_unlock(&g.mtx);
    if (!dev) return pp_process_frame_stock(sei, frame_p);
    OVFrame *post = ovframepool_request_frame(frame->internal.frame_pool);
    if (!post) { ov_log(NULL, OVLOG_ERROR, "Could not get a writable picture for post processing\n"); ovframe_unref(frame_p); return OVVC_ENOMEM; }
    const int ow = up ? (int)post->width : w, oh = up ? (int)post->height : h;
    int rc = 0;
    pthread_mutex_lock(&pp.mtx);
    if (!pp.ctx && ovgpu_ctx_create(&pp.ctx, g.device, NULL) < 0) rc = -1;
    if (!rc && (!pp.out || pp.out_w != ow || pp.out_h != oh)) {
        if (pp.out) { ovgpu_ctx_sync(pp.ctx); ovgpu_picture_destroy(pp.out); pp.out = NULL; }
        if (ovgpu_picture_create(pp.ctx, ow, oh, 10, &pp.out) < 0) rc = -1;
        pp.out_w = ow; pp.out_h = oh;
    }
    if (!rc && up) {
        OVGPUScalingWindow win = { sei->scaling_info.scaling_win_left, sei->scaling_info.scaling_win_right, sei->scaling_info.scaling_win_top,
                                   sei->scaling_info.scaling_win_bottom, sei->scaling_info.chroma_hor_col_flag, sei->scaling_info.chroma_ver_col_flag };
        rc = ovgpu_picture_upscale(pp.ctx, pp.out, dev, &win);
    } else if (!rc) {
        const struct OVSEIFGrain *s = sei->sei_fg;
        OVGPUFilmGrain fg;
        memset(&fg, 0, sizeof(fg));
        fg.log2_scale_factor = s->fg_log2_scale_factor;
        for (int c = 0; c < 3; c++) {
            fg.comp_model_present_flag[c] = s->fg_comp_model_present_flag[c]; fg.num_model_values_minus1[c] = s->fg_num_model_values_minus1[c];
            for (int i = 0; i < 8; i++) {
                fg.intensity_interval_lower_bound[c][i] = s->fg_intensity_interval_lower_bound[c][i];
                fg.intensity_interval_upper_bound[c][i] = s->fg_intensity_interval_upper_bound[c][i];
                for (int k = 0; k < 3; k++) fg.comp_model_value[c][i][k] = s->fg_comp_model_value[c][i][k];
            }
        }
        if (!pp.fg && ovgpu_film_grain_create(pp.ctx, &pp.fg) < 0) rc = -1;
        if (!rc) rc = ovgpu_film_grain_picture(pp.ctx, pp.fg, pp.out, dev, &fg, (int32_t)frame->poc);
    }
    post->width = ow; post->height = oh;
    if (!rc) rc = ovgpu_picture_download(pp.ctx, pp.out, (void *const *)post->data, post->linesize);
    if (!rc) rc = ovgpu_ctx_sync(pp.ctx);
    if (rc < 0) ov_log(NULL, OVLOG_ERROR, "rcn cuda back end: output stage failed (%d): %s\n", rc, pp.ctx ? ovgpu_last_error(pp.ctx) : "no context");
    pthread_mutex_unlock(&pp.mtx);
    if (g.verbose) fprintf(stderr, "rcn cuda: output stage of POC %d on the GPU: %s -> %d x %d%s\n", (int)frame->poc, up ? "up-scaling" : "film grain", ow, oh, rc < 0 ? "  FAILED" : "");
    if (rc < 0) {                        /* the decoded frame is intact in host memory: let the stock function produce the output */
        ovframe_unref(&post);
        return pp_process_frame_stock(sei, frame_p);
    }
    ovframe_unref(frame_p);
    *frame_p = post;
    return 0;
}

/* Takes the place of the reference's rcn_init_functions (rcn.c:148; slicedec.c:1462 calls it per entry job). */
void rcn_init_functions(struct RCNFunctions *rcn_func, uint8_t ict_type, uint8_t lm_chroma_enabled,
                        uint8_t sps_chroma_vertical_collocated_flag, uint8_t lmcs_flag, uint8_t bitdepth)
{
    rcn_init_functions_stock(rcn_func, ict_type, lm_chroma_enabled, sps_chroma_vertical_collocated_flag, lmcs_flag, bitdepth);
    if (!ovshim_backend_enabled()) return;
    if (bitdepth != 10) { ov_log(NULL, OVLOG_WARNING, "rcn cuda back end: %d-bit stream, using the CPU table\n", bitdepth); return; }
    OVCTUDec *c = (OVCTUDec *)((char *)rcn_func - offsetof(OVCTUDec, rcn_funcs));
    ShimRec *r = rec_of(c);
    r->stock = *rcn_func;
    r->ict_type = ict_type; r->collocated = sps_chroma_vertical_collocated_flag;
    rcn_init_functions_cuda(rcn_func);
}
