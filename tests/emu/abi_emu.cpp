// tests/emu/abi_emu.cpp -- TEST HARNESS ONLY: the C ABI of include/ovvc_b200.h implemented over the CPU emulation of the product's
// kernel SOURCE (emu.cpp), with "device" pictures in host memory.  It exists so that the decoder-side recorders of shim/rcn_cuda.c
// (host logic) can be run and debugged against the stock decoder on a machine without a GPU: build/libovvc_emu.so links the same
// shim against this library instead of libovvc_b200.so.  Never shipped, never linked into the product; the product library has no
// CPU path (ovgpu_ctx_create fails without an sm_100 device).  Only the entry points the shim uses are implemented.
#include "emu.cpp"
#include <stdarg.h>
#include <stdio.h>

struct OVGPUCtx { char err[256]; };
struct OVGPUPicture { int w, h; uint16_t *plane[3]; int stride[3]; };
struct OVGPUMcBatch { std::vector<OVGPUMcCmd> cmds; std::vector<OVGPUProfInfo> prof; int nb, nb_prof; std::vector<OVGPUDmvrMv> dmvr; int nb_dmvr;
                      OVGPUMcScaling sc; uint32_t scaled[2]; };
struct OVGPUTuBatch { std::vector<OVGPUTuCmd> cmds; std::vector<int16_t> coeffs; int nb; size_t nb_coeffs; OVGPUTuSlice sl; std::vector<OVGPULmcsAnchor> anchors; };
struct OVGPUIntraBatch { std::vector<OVGPUIntraCmd> cmds; std::vector<OVGPUIspInfo> isp; int w, h, log2_ctu, levels; };
struct OVGPUDbfParams { int w, h, log2_ctu, beta, tc; std::vector<uint32_t> ver[OVGPU_DBF_NB_MAPS], hor[OVGPU_DBF_NB_MAPS]; std::vector<uint8_t> qp[3]; };
struct OVGPUSaoParams { int w, h, log2_ctu; std::vector<OVGPUSaoCtu> ctu; OVGPUTileGrid tiles; };
struct OVGPUAlfParams { int w, h, log2_ctu, en[2]; OVGPUAlfTables tab; std::vector<OVGPUAlfCtu> ctu; std::vector<uint8_t> cc[2]; OVGPUTileGrid tiles; };
struct OVGPULmcs { OVGPULmcsLuts luts; OVGPULmcsData data; int has_data; };

static int fail(OVGPUCtx *c, int code, const char *fmt, ...)
{
    if (c) { va_list ap; va_start(ap, fmt); vsnprintf(c->err, sizeof(c->err), fmt, ap); va_end(ap); }
    return code;
}
static int nctu(int w, int h, int l) { const int s = 1 << l; return ((w + s - 1) >> l) * ((h + s - 1) >> l); }

extern "C" {

int ovgpu_ctx_create(OVGPUCtx **ctx, int, void *) { *ctx = new OVGPUCtx(); (*ctx)->err[0] = 0; return 0; }
void ovgpu_ctx_destroy(OVGPUCtx *c) { delete c; }
int ovgpu_ctx_sync(OVGPUCtx *) { return 0; }
const char *ovgpu_last_error(const OVGPUCtx *c) { return c ? c->err : "no context"; }
int ovgpu_host_register(OVGPUCtx *, void *, size_t) { return 0; }
int ovgpu_host_unregister(OVGPUCtx *, void *) { return 0; }

int ovgpu_picture_create(OVGPUCtx *, int w, int h, int, OVGPUPicture **out)
{
    OVGPUPicture *p = new OVGPUPicture();
    p->w = w; p->h = h;
    for (int c = 0; c < 3; c++) {
        const int pw = c ? w >> 1 : w, ph = c ? h >> 1 : h;
        p->stride[c] = (pw + 127) & ~127;
        p->plane[c] = (uint16_t *)calloc((size_t)p->stride[c] * ph + 64, sizeof(uint16_t));
    }
    *out = p;
    return 0;
}
void ovgpu_picture_destroy(OVGPUPicture *p) { if (!p) return; for (int c = 0; c < 3; c++) free(p->plane[c]); delete p; }
int ovgpu_picture_download(OVGPUCtx *, const OVGPUPicture *p, void *const data[3], const size_t linesize[3])
{
    for (int c = 0; c < 3; c++) {
        const int pw = c ? p->w >> 1 : p->w, ph = c ? p->h >> 1 : p->h;
        for (int y = 0; y < ph; y++) memcpy((char *)data[c] + (size_t)y * linesize[c], p->plane[c] + (size_t)y * p->stride[c], (size_t)pw * 2);
    }
    return 0;
}

// output stage (shim: pp_process_frame)
int ovgpu_picture_upscale(OVGPUCtx *c, OVGPUPicture *dst, const OVGPUPicture *src, const OVGPUScalingWindow *win)
{
    const int w6[6] = { win ? win->left : 0, win ? win->right : 0, win ? win->top : 0, win ? win->bottom : 0, win ? win->chroma_hor_col != 0 : 1, win ? win->chroma_ver_col != 0 : 0 };
    if (emu_pp_upscale(dst->plane, dst->stride, dst->w, dst->h, src->plane, src->stride, src->w, src->h, w6)) return fail(c, OVGPU_EINVAL, "emulated up-scaling refused");
    return 0;
}
struct OVGPUFilmGrainSynth { int dummy; };
int ovgpu_film_grain_create(OVGPUCtx *, OVGPUFilmGrainSynth **s) { *s = new OVGPUFilmGrainSynth(); return 0; }
void ovgpu_film_grain_destroy(OVGPUFilmGrainSynth *s) { delete s; }
int ovgpu_film_grain_picture(OVGPUCtx *c, OVGPUFilmGrainSynth *, OVGPUPicture *dst, const OVGPUPicture *src, const OVGPUFilmGrain *fg, int32_t poc)
{
    if (dst->w != src->w || dst->h != src->h || emu_film_grain(dst->plane, dst->stride, src->plane, src->stride, src->w, src->h, fg, poc)) return fail(c, OVGPU_EINVAL, "emulated film grain refused");
    return 0;
}

/* same checks as the product's (sao.cu): boundaries ascending from 0 to the picture's size in CTUs, at most 20 x 20 tiles */
int ovgpu_tile_grid_check(const OVGPUTileGrid *t, int ncw, int nch)
{
    if (!t) return -1;
    if (t->nb_cols < 1 || t->nb_rows < 1 || t->nb_cols > 20 || t->nb_rows > 20 || t->col_bd[0] || t->row_bd[0] ||
        t->col_bd[t->nb_cols] != ncw || t->row_bd[t->nb_rows] != nch) return OVGPU_EINVAL;
    for (int i = 0; i < t->nb_cols; i++) if (t->col_bd[i + 1] <= t->col_bd[i]) return OVGPU_EINVAL;
    for (int i = 0; i < t->nb_rows; i++) if (t->row_bd[i + 1] <= t->row_bd[i]) return OVGPU_EINVAL;
    return 0;
}
static int set_tiles(OVGPUCtx *c, OVGPUTileGrid *dst, const OVGPUTileGrid *t, int w, int h, int l)
{
    const int s = 1 << l;
    if (!t || t->nb_cols * t->nb_rows <= 1) { memset(dst, 0, sizeof(*dst)); return 0; }
    if (ovgpu_tile_grid_check(t, (w + s - 1) >> l, (h + s - 1) >> l) < 0) return fail(c, OVGPU_EINVAL, "tile grid does not cover the picture");
    *dst = *t;
    return 0;
}
int ovgpu_alf_params_set_tiles(OVGPUCtx *c, OVGPUAlfParams *p, const OVGPUTileGrid *t) { return set_tiles(c, &p->tiles, t, p->w, p->h, p->log2_ctu); }
int ovgpu_sao_params_set_tiles(OVGPUCtx *c, OVGPUSaoParams *p, const OVGPUTileGrid *t) { return set_tiles(c, &p->tiles, t, p->w, p->h, p->log2_ctu); }

int ovgpu_mc_batch_create(OVGPUCtx *, int max_cmds, int, OVGPUMcBatch **b) { *b = new OVGPUMcBatch(); (*b)->cmds.resize(max_cmds + 1); (*b)->prof.resize(65536); (*b)->nb = (*b)->nb_prof = (*b)->nb_dmvr = 0; return ovgpu_mc_batch_set_scaling(NULL, *b, NULL); }
int ovgpu_mc_batch_set_scaling(OVGPUCtx *, OVGPUMcBatch *b, const OVGPUMcScaling *s)
{
    b->scaled[0] = b->scaled[1] = 0;
    for (int l = 0; l < 2; l++)
        for (int i = 0; i < 16; i++) {
            for (int k = 0; k < 2; k++) b->sc.scale[l][i][k] = s && s->scale[l][i][k] ? s->scale[l][i][k] : 1 << 14;
            if (b->sc.scale[l][i][0] != (1 << 14) || b->sc.scale[l][i][1] != (1 << 14)) b->scaled[l] |= 1u << i;
        }
    b->sc.chroma_hor_col = s ? s->chroma_hor_col : 1; b->sc.chroma_ver_col = s ? s->chroma_ver_col : 0;
    return 0;
}
void ovgpu_mc_batch_destroy(OVGPUMcBatch *b) { delete b; }
int ovgpu_mc_batch_staging(OVGPUCtx *, OVGPUMcBatch *b, OVGPUMcCmd **cmds, OVGPUProfInfo **prof) { *cmds = b->cmds.data(); *prof = b->prof.data(); return 0; }
int ovgpu_mc_batch_commit(OVGPUCtx *c, OVGPUMcBatch *b, int nb, int nb_prof)
{
    if (nb > (int)b->cmds.size() || nb_prof > (int)b->prof.size()) return fail(c, OVGPU_EINVAL, "MC batch overflow");
    for (int i = 0; i < nb; i++) if (mc_check_cmd(&b->cmds[i], nb_prof, b->scaled)) return fail(c, OVGPU_EINVAL, "MC command %d malformed", i);
    b->nb = nb; b->nb_prof = nb_prof;
    return 0;
}
int ovgpu_mc_batch_status(OVGPUCtx *, OVGPUMcBatch *) { return 0; }      // (this emulation validates in the commit)
int ovgpu_mc_batch_read_dmvr(OVGPUCtx *, OVGPUMcBatch *b, OVGPUDmvrMv *out, int max)
{
    for (int i = 0; i < b->nb_dmvr && i < max; i++) out[i] = b->dmvr[i];
    return b->nb_dmvr;
}

int ovgpu_tu_batch_create(OVGPUCtx *, int max_cmds, size_t max_coeffs, OVGPUTuBatch **b) { *b = new OVGPUTuBatch(); (*b)->cmds.resize(max_cmds + 1); (*b)->coeffs.resize(max_coeffs + 16); (*b)->nb = 0; (*b)->nb_coeffs = 0; return 0; }
void ovgpu_tu_batch_destroy(OVGPUTuBatch *b) { delete b; }
int ovgpu_tu_batch_staging(OVGPUCtx *, OVGPUTuBatch *b, OVGPUTuCmd **cmds, int16_t **coeffs) { *cmds = b->cmds.data(); *coeffs = b->coeffs.data(); return 0; }
int ovgpu_tu_batch_commit(OVGPUCtx *c, OVGPUTuBatch *b, int nb, size_t nb_coeffs, const OVGPUTuSlice *sl)
{
    if (nb > (int)b->cmds.size() || nb_coeffs > b->coeffs.size()) return fail(c, OVGPU_EINVAL, "TU batch overflow");
    for (int i = 0; i < nb; i++) if (tu_check_cmd(&b->cmds[i], nb_coeffs)) return fail(c, OVGPU_EINVAL, "TU command %d malformed", i);
    b->nb = nb; b->nb_coeffs = nb_coeffs; b->sl = *sl;
    return 0;
}
int ovgpu_tu_batch_set_anchors(OVGPUCtx *, OVGPUTuBatch *b, const OVGPULmcsAnchor *a, int nb) { b->anchors.assign(a, a + nb); return 0; }

int ovgpu_intra_batch_create(OVGPUCtx *, int, OVGPUIntraBatch **b) { *b = new OVGPUIntraBatch(); (*b)->levels = 0; return 0; }
void ovgpu_intra_batch_destroy(OVGPUIntraBatch *b) { delete b; }
int ovgpu_intra_batch_upload2(OVGPUCtx *, OVGPUIntraBatch *b, const OVGPUIntraCmd *cmds, int nb, const OVGPUIspInfo *isp, int nb_isp, int w, int h, int log2_ctu)
{
    b->cmds.assign(cmds, cmds + nb); b->isp.assign(isp, isp + nb_isp); b->w = w; b->h = h; b->log2_ctu = log2_ctu;
    return 0;
}
int ovgpu_intra_batch_levels(const OVGPUIntraBatch *b) { return b ? b->levels : OVGPU_EINVAL; }

int ovgpu_dbf_params_create(OVGPUCtx *, int w, int h, int log2_ctu, OVGPUDbfParams **out)
{
    OVGPUDbfParams *p = new OVGPUDbfParams();
    p->w = w; p->h = h; p->log2_ctu = log2_ctu; p->beta = p->tc = 0;
    const int ctu = 1 << log2_ctu, ncw = (w + ctu - 1) >> log2_ctu, nch = (h + ctu - 1) >> log2_ctu, w_u = w >> 2, h_u = h >> 2;
    for (int k = 0; k < OVGPU_DBF_NB_MAPS; k++) { p->ver[k].assign((size_t)nch * (w_u + 1) + 8, 0); p->hor[k].assign((size_t)ncw * (h_u + 1) + 8, 0); }
    for (int c = 0; c < 3; c++) p->qp[c].assign((size_t)w_u * h_u + 8, 0);
    *out = p;
    return 0;
}
void ovgpu_dbf_params_destroy(OVGPUDbfParams *p) { delete p; }
int ovgpu_dbf_params_staging(OVGPUCtx *, OVGPUDbfParams *p, OVGPUDbfMaps *m)
{
    m->width = p->w; m->height = p->h; m->log2_ctu_s = p->log2_ctu; m->beta_offset = m->tc_offset = 0;
    for (int k = 0; k < OVGPU_DBF_NB_MAPS; k++) { m->ver[k] = p->ver[k].data(); m->hor[k] = p->hor[k].data(); }
    for (int c = 0; c < 3; c++) m->qp[c] = p->qp[c].data();
    return 0;
}
int ovgpu_dbf_params_commit(OVGPUCtx *, OVGPUDbfParams *p, int beta, int tc) { p->beta = beta; p->tc = tc; return 0; }
int ovgpu_dbf_picture(OVGPUCtx *c, OVGPUPicture *pic, const OVGPUDbfParams *p)
{
    OVGPUDbfMaps m;
    ovgpu_dbf_params_staging(c, (OVGPUDbfParams *)p, &m);
    m.beta_offset = (int16_t)p->beta; m.tc_offset = (int16_t)p->tc;
    return emu_dbf_picture(pic->plane, pic->stride, 10, &m);
}

int ovgpu_sao_params_create(OVGPUCtx *, int w, int h, int l, OVGPUSaoParams **p) { *p = new OVGPUSaoParams(); (*p)->w = w; (*p)->h = h; (*p)->log2_ctu = l; return 0; }
void ovgpu_sao_params_destroy(OVGPUSaoParams *p) { delete p; }
int ovgpu_sao_params_upload(OVGPUCtx *, OVGPUSaoParams *p, const OVGPUSaoCtu *ctu) { p->ctu.assign(ctu, ctu + nctu(p->w, p->h, p->log2_ctu)); return 0; }
int ovgpu_sao_picture(OVGPUCtx *, OVGPUPicture *dst, const OVGPUPicture *src, const OVGPUSaoParams *p)
{ return emu_sao_picture_tiles(dst->plane, dst->stride, src->plane, src->stride, p->w, p->h, p->log2_ctu, 10, p->ctu.data(), p->tiles.nb_cols * p->tiles.nb_rows > 1 ? &p->tiles : NULL); }

int ovgpu_alf_params_create(OVGPUCtx *, int w, int h, int l, OVGPUAlfParams **p) { *p = new OVGPUAlfParams(); (*p)->w = w; (*p)->h = h; (*p)->log2_ctu = l; return 0; }
void ovgpu_alf_params_destroy(OVGPUAlfParams *p) { delete p; }
int ovgpu_alf_params_upload(OVGPUCtx *, OVGPUAlfParams *p, const OVGPUAlfTables *tab, const OVGPUAlfCtu *ctu, const uint8_t *const cc[2], int cb, int cr)
{
    const int n = nctu(p->w, p->h, p->log2_ctu);
    p->tab = *tab; p->ctu.assign(ctu, ctu + n); p->en[0] = cb; p->en[1] = cr;
    for (int k = 0; k < 2; k++) { if (cc && cc[k]) p->cc[k].assign(cc[k], cc[k] + n); else p->cc[k].assign(n, 0); }
    return 0;
}
int ovgpu_alf_picture(OVGPUCtx *, OVGPUPicture *dst, const OVGPUPicture *src, const OVGPUAlfParams *p)
{
    const uint8_t *cc[2] = { p->cc[0].data(), p->cc[1].data() };
    emu_alf_picture_tiles(dst->plane, dst->stride, src->plane, src->stride, p->w, p->h, p->log2_ctu, 10, &p->tab, p->ctu.data(), cc, p->en[0], p->en[1], p->tiles.nb_cols * p->tiles.nb_rows > 1 ? &p->tiles : NULL);
    return 0;
}

int ovgpu_lmcs_create(OVGPUCtx *, OVGPULmcs **l) { *l = new OVGPULmcs(); (*l)->has_data = 0; return 0; }
void ovgpu_lmcs_destroy(OVGPULmcs *l) { delete l; }
int ovgpu_lmcs_upload(OVGPUCtx *, OVGPULmcs *l, const OVGPULmcsLuts *luts) { l->luts = *luts; return 0; }
int ovgpu_lmcs_set_data(OVGPULmcs *l, const OVGPULmcsData *d) { l->data = *d; l->has_data = 1; return 0; }
int ovgpu_lmcs_reshape_picture(OVGPUCtx *, OVGPUPicture *pic, const OVGPULmcs *l, int forward)
{ return emu_lmcs_reshape(pic->plane[0], pic->stride[0], pic->w, pic->h, 10, forward ? l->luts.fwd_lut : l->luts.bwd_lut); }

int ovgpu_recon_picture(OVGPUCtx *c, OVGPUPicture *pic, const OVGPUMcRefs *refs, const OVGPUMcBatch *mcb, const OVGPULmcs *lmcs,
                        OVGPUTuBatch *tu, OVGPUIntraBatch *intra)
{
    OVGPUMcBatch *mc = (OVGPUMcBatch *)mcb;
    if (mc && mc->nb) {
        EmuMcRefs er;
        EmuMcScaling es;
        memset(&er, 0, sizeof(er));
        memset(&es, 0, sizeof(es));
        es.chroma_hor_col = mc->sc.chroma_hor_col; es.chroma_ver_col = mc->sc.chroma_ver_col;
        for (int l = 0; l < 2; l++) for (int i = 0; i < 16; i++) {
            const OVGPUPicture *rp = refs->rpl[l][i];
            er.poc[l][i] = refs->poc[l][i];
            es.scale[l][i][0] = mc->sc.scale[l][i][0]; es.scale[l][i][1] = mc->sc.scale[l][i][1];
            es.ref_w[l][i] = rp ? rp->w : pic->w; es.ref_h[l][i] = rp ? rp->h : pic->h;
            if (rp && !((mc->scaled[l] >> i) & 1) && (rp->w != pic->w || rp->h != pic->h)) return fail(c, OVGPU_EINVAL, "reference picture differs in size and has no scale factor");
            if (rp) for (int k = 0; k < 3; k++) { er.plane[l][i][k] = rp->plane[k]; er.stride[l][i][k] = rp->stride[k]; }
        }
        mc->nb_dmvr = mc_count_tool_items(mc->cmds.data(), mc->nb, OVGPU_MC_TOOL_DMVR);
        mc->dmvr.resize(mc->nb_dmvr + 1);
        int rc = emu_mc_picture_rpr(pic->plane, pic->stride, pic->w, pic->h, 10, &er, &es, mc->cmds.data(), mc->nb, lmcs ? lmcs->luts.fwd_lut : NULL,
                                mc->prof.data(), mc->nb_prof, mc->dmvr.data());
        if (rc < 0) return fail(c, OVGPU_EINVAL, "emulated MC: %d", rc);
    }
    static const OVGPULmcsLuts no_luts = { { 0 }, { 0 }, { 0 } };
    static const OVGPULmcsData no_data = { 0, 0, 0, { 0 }, { 0 }, 0, 0 };
    const int ni = intra ? (int)intra->cmds.size() : 0;
    int rc = emu_recon_picture(pic->plane, pic->stride, pic->w, pic->h, intra && ni ? intra->log2_ctu : 7, 10, ni ? intra->cmds.data() : NULL, ni,
                               tu ? tu->cmds.data() : NULL, tu ? tu->nb : 0, tu ? tu->coeffs.data() : NULL, tu ? tu->nb_coeffs : 0,
                               tu ? tu->sl.dep_quant : 0, tu ? tu->sl.sh_ts_disabled : 0, tu ? tu->sl.ict_type : 0,
                               tu && !tu->anchors.empty() ? tu->anchors.data() : NULL, tu ? (int)tu->anchors.size() : 0,
                               lmcs ? &lmcs->luts : &no_luts, lmcs && lmcs->has_data ? &lmcs->data : &no_data,
                               ni && !intra->isp.empty() ? intra->isp.data() : NULL, ni ? (int)intra->isp.size() : 0);
    if (rc < 0) return fail(c, OVGPU_EINVAL, "emulated reconstruction: %d", rc);
    if (intra) intra->levels = rc;
    return 0;
}

}
