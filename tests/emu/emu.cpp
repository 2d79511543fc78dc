// tests/emu/emu.cpp -- TEST HARNESS ONLY: runs the product's kernel source (openvvc_b200/csrc/*_kernels.h)
// on the CPU through the serial CTA emulation of csrc/dev.h (-DOVGPU_HOST_EMU), so kernel logic can be
// checked against the oracle on a machine without a GPU.  Never linked into libovvc_b200.so.
#include <vector>
#include <algorithm>
#include "../../openvvc_b200/csrc/alf_host.h"
#include "../../openvvc_b200/csrc/sao_kernels.h"
#include "../../openvvc_b200/csrc/mc_host.h"
#include "../../openvvc_b200/csrc/tr_host.h"
#include "../../openvvc_b200/csrc/dbf_kernels.h"
#include "../../openvvc_b200/csrc/intra_host.h"
#include "../../openvvc_b200/csrc/pp_kernels.h"
#include "../../openvvc_b200/csrc/fg_kernels.h"

static thread_local const int16_t *emu_tr_scales = NULL;     // LMCS chroma scales of the running emu_recon_picture call

template <int GROUP>
static void emu_tr_launch(uint16_t *const dst[3], const int dst_stride[3], int bitdepth, const OVGPUTuCmd *cmds,
                          const TbItemDev *items, int n, const int16_t *arena, const OVGPUTuSlice *sl)
{
    typedef TrK<GROUP> K;
    typename K::Args a;
    memset(&a, 0, sizeof(a));
    for (int c = 0; c < 3; c++) { a.dst[c] = dst[c]; a.dst_pitch[c] = dst_stride[c]; }
    a.bitdepth = bitdepth;
    a.cmds = (const TuCmdDev *)cmds; a.items = items; a.nb_items = n; a.arena = arena;
    a.dep_quant = sl->dep_quant; a.sh_ts_disabled = sl->sh_ts_disabled; a.ict_type = sl->ict_type;
    a.scales = emu_tr_scales;
    ovd_emulate<K>(a, ovd_dim3((n + K::GROUPS - 1) / K::GROUPS, 1, 1), K::SMEM_BYTES);
}

// one CTA of kernel K, phases p0 .. p1, on a caller-owned shared-memory buffer (the worker of ovd_recon_kernel interleaves the
// phases of two kernels on one buffer)
template <class K>
static void emu_phases(const typename K::Args &a, unsigned char *sm, int p0, int p1)
{
    std::vector<typename K::State> st(K::THREADS);
    if (p0 <= 0 && p1 >= 0) for (int t = 0; t < K::THREADS; t++) K::template phase<0>(a, sm, st[t], t, 0, 0, 0);
    if (p0 <= 1 && p1 >= 1) for (int t = 0; t < K::THREADS; t++) K::template phase<1>(a, sm, st[t], t, 0, 0, 0);
    if (p0 <= 2 && p1 >= 2) for (int t = 0; t < K::THREADS; t++) K::template phase<2>(a, sm, st[t], t, 0, 0, 0);
    if (p0 <= 3 && p1 >= 3) for (int t = 0; t < K::THREADS; t++) K::template phase<3>(a, sm, st[t], t, 0, 0, 0);
}
template <int GROUP>
static typename TrK<GROUP>::Args emu_tr_args(uint16_t *const dst[3], const int dst_stride[3], int bitdepth, const OVGPUTuCmd *cmds,
                                             const TbItemDev *items, int n, const int16_t *arena, const OVGPUTuSlice *sl)
{
    typename TrK<GROUP>::Args a;
    memset(&a, 0, sizeof(a));
    for (int c = 0; c < 3; c++) { a.dst[c] = dst[c]; a.dst_pitch[c] = dst_stride[c]; }
    a.bitdepth = bitdepth;
    a.cmds = (const TuCmdDev *)cmds; a.items = items; a.nb_items = n; a.arena = arena;
    a.dep_quant = sl->dep_quant; a.sh_ts_disabled = sl->sh_ts_disabled; a.ict_type = sl->ict_type;
    a.scales = emu_tr_scales;
    return a;
}

extern "C" {

void emu_set_reverse_block_order(int on) { ovd_emu_reverse_blocks = on; }

int emu_alf_picture_tiles(uint16_t *const dst[3], const int dst_stride[3],
                          const uint16_t *const src[3], const int src_stride[3],
                          int w, int h, int log2_ctu, int bitdepth,
                          const OVGPUAlfTables *tab, const OVGPUAlfCtu *ctu, const uint8_t *const cc_idx[2],
                          int cc_cb_enabled, int cc_cr_enabled, const OVGPUTileGrid *tiles);

int emu_alf_picture(uint16_t *const dst[3], const int dst_stride[3],
                    const uint16_t *const src[3], const int src_stride[3],
                    int w, int h, int log2_ctu, int bitdepth,
                    const OVGPUAlfTables *tab, const OVGPUAlfCtu *ctu, const uint8_t *const cc_idx[2],
                    int cc_cb_enabled, int cc_cr_enabled)
{
    return emu_alf_picture_tiles(dst, dst_stride, src, src_stride, w, h, log2_ctu, bitdepth, tab, ctu, cc_idx, cc_cb_enabled, cc_cr_enabled, NULL);
}

int emu_alf_picture_tiles(uint16_t *const dst[3], const int dst_stride[3],
                          const uint16_t *const src[3], const int src_stride[3],
                          int w, int h, int log2_ctu, int bitdepth,
                          const OVGPUAlfTables *tab, const OVGPUAlfCtu *ctu, const uint8_t *const cc_idx[2],
                          int cc_cb_enabled, int cc_cr_enabled, const OVGPUTileGrid *tiles)
{
    const TileGridDev *tgd = tiles && tiles->nb_cols * tiles->nb_rows > 1 ? (const TileGridDev *)tiles : NULL;
    AlfGeom g = alf_geom(w, h, log2_ctu);
    int n = g.nb_ctu_w * g.nb_ctu_h;
    std::vector<uint32_t> ltab((size_t)OVGPU_ALF_NUM_SETS * ALF_TAB_SET_WORDS + 4);
    // 16-byte align like cudaMalloc would
    uint32_t *lt = ltab.data();
    while ((uintptr_t)lt & 15) lt++;
    int wide = alf_pack_luma_tab(lt, tab);
    std::vector<AlfCtuDev> dctu(n);
    alf_pack_ctu(dctu.data(), n, ctu, cc_idx);
    ovd_dim3 grid(g.nb_ctu_w * g.l_tx, g.nb_ctu_h * g.l_ty, 1);
    if (wide) {
        typedef AlfLumaK<true> K;
        K::Args a = { src[0], dst[0], src_stride[0], dst_stride[0], w, h, log2_ctu, bitdepth,
                      g.l_tw, g.l_th, g.l_tx, g.l_ty, g.nb_ctu_w, lt, dctu.data(), tgd };
        ovd_emulate<K>(a, grid, K::SMEM_BYTES);
    } else {
        typedef AlfLumaK<false> K;
        K::Args a = { src[0], dst[0], src_stride[0], dst_stride[0], w, h, log2_ctu, bitdepth,
                      g.l_tw, g.l_th, g.l_tx, g.l_ty, g.nb_ctu_w, lt, dctu.data(), tgd };
        ovd_emulate<K>(a, grid, K::SMEM_BYTES);
    }
    {
        typedef AlfChromaK K;
        K::Args a;
        a.src[0] = src[1]; a.src[1] = src[2]; a.dst[0] = dst[1]; a.dst[1] = dst[2];
        a.luma = src[0];
        a.src_pitch = src_stride[1]; a.dst_pitch = dst_stride[1]; a.luma_pitch = src_stride[0];
        a.pic_w = w; a.pic_h = h; a.log2_ctu = log2_ctu; a.bitdepth = bitdepth;
        a.tw = g.c_tw; a.th = g.c_th; a.tiles_x_per_ctu = g.c_tx; a.tiles_y_per_ctu = g.c_ty; a.nb_ctu_w = g.nb_ctu_w;
        a.cc_enabled[0] = cc_cb_enabled && cc_idx && cc_idx[0];
        a.cc_enabled[1] = cc_cr_enabled && cc_idx && cc_idx[1];
        a.ctu = dctu.data();
        a.tiles = tgd;
        alf_pack_chroma_tab(&a.tab, tab);
        ovd_emulate<K>(a, ovd_dim3(g.nb_ctu_w * g.c_tx, g.nb_ctu_h * g.c_ty, 2), K::SMEM_BYTES);
    }
    return wide;
}

int emu_sao_picture_tiles(uint16_t *const dst[3], const int dst_stride[3],
                          const uint16_t *const src[3], const int src_stride[3],
                          int w, int h, int log2_ctu, int bitdepth, const OVGPUSaoCtu *ctu, const OVGPUTileGrid *tiles)
{
    SaoK::Args a;
    int ctu_s = 1 << log2_ctu;
    for (int c = 0; c < 3; c++) {
        a.src[c] = src[c]; a.dst[c] = dst[c]; a.src_pitch[c] = src_stride[c]; a.dst_pitch[c] = dst_stride[c];
    }
    a.pic_w = w; a.pic_h = h; a.log2_ctu = log2_ctu; a.bitdepth = bitdepth;
    a.nb_ctu_w = (w + ctu_s - 1) >> log2_ctu; a.nb_ctu_h = (h + ctu_s - 1) >> log2_ctu;
    a.ctu = (const SaoCtuDev *)ctu;
    a.tiles = tiles && tiles->nb_cols * tiles->nb_rows > 1 ? (const TileGridDev *)tiles : NULL;
    ovd_emulate<SaoK>(a, ovd_dim3((w + SaoK::BX * 8 - 1) / (SaoK::BX * 8), (h + SaoK::BY - 1) / SaoK::BY, 3), 0);
    return 0;
}

int emu_sao_picture(uint16_t *const dst[3], const int dst_stride[3],
                    const uint16_t *const src[3], const int src_stride[3],
                    int w, int h, int log2_ctu, int bitdepth, const OVGPUSaoCtu *ctu)
{
    return emu_sao_picture_tiles(dst, dst_stride, src, src_stride, w, h, log2_ctu, bitdepth, ctu, NULL);
}

// output stage: up-scaling (UpscaleK); win = left, right, top, bottom (chroma units), hor_col, ver_col
int emu_pp_upscale(uint16_t *const dst[3], const int dst_stride[3], int dst_w, int dst_h,
                   const uint16_t *const src[3], const int src_stride[3], int src_w, int src_h, const int win[6])
{
    UpscaleK::Args a;
    memset(&a, 0, sizeof(a));
    for (int c = 0; c < 3; c++) { a.dst[c] = dst[c]; a.dst_pitch[c] = dst_stride[c]; a.src[c] = src[c]; a.src_pitch[c] = src_stride[c]; }
    a.dst_w = dst_w; a.dst_h = dst_h; a.src_w = src_w; a.src_h = src_h; a.maxv = 1023;
    if (upscale_setup(&a, win[0], win[1], win[2], win[3], win[4], win[5])) return -1;
    ovd_emulate<UpscaleK>(a, ovd_dim3((dst_w + 31) / 32, (dst_h + 7) / 8, 3), 0);
    return 0;
}

// output stage: film grain synthesis (FgSeedK + FilmGrainK)
int emu_film_grain(uint16_t *const dst[3], const int dst_stride[3], const uint16_t *const src[3], const int src_stride[3], int w, int h,
                   const OVGPUFilmGrain *fg, int32_t poc)
{
    static std::vector<int8_t> db;
    if (db.empty()) { db.resize(FG_DB_BYTES); fg_build_database(db.data()); }
    if (!fg_size_ok(w, h)) return -1;
    static const uint8_t colour_offset[3] = { 0, 85, 170 };
    FilmGrainK::Args a;
    memset(&a, 0, sizeof(a));
    FgSeedK::Args sa;
    memset(&sa, 0, sizeof(sa));
    std::vector<uint32_t> seeds[3];
    for (int c = 0; c < 3; c++) {
        const int cw = w >> (c != 0), chh = h >> (c != 0);
        sa.nb16w[c] = (cw + 15) >> 4; sa.nb16h[c] = (chh + 15) >> 4;
        seeds[c].resize((size_t)sa.nb16w[c] * sa.nb16h[c]);
        sa.seeds[c] = seeds[c].data(); a.seeds[c] = seeds[c].data();
        sa.seed0[c] = fg_seeds[((uint32_t)poc + colour_offset[c]) % 256];
        a.dst[c] = dst[c]; a.dst_pitch[c] = dst_stride[c]; a.src[c] = src[c]; a.src_pitch[c] = src_stride[c];
    }
    a.w = w; a.h = h; a.db = db.data();
    fg_make_params(&a.p, fg);
    ovd_emulate<FgSeedK>(sa, ovd_dim3(1, 1, 3), 0);
    ovd_emulate<FilmGrainK>(a, ovd_dim3((w + FilmGrainK::COLS - 1) / FilmGrainK::COLS, (h + 15) / 16, 3), FilmGrainK::SMEM_BYTES);
    return 0;
}

// refs: plane pointers / strides (samples) / pocs laid out like OVOMcRefs ([list][ref][comp])
struct EmuMcRefs {
    const uint16_t *plane[2][16][3];
    int stride[2][16][3];
    int32_t poc[2][16];
};

// reference picture resampling: like OVOMcScaling (oracle.h)
struct EmuMcScaling {
    uint16_t scale[2][16][2];
    int32_t ref_w[2][16], ref_h[2][16];
    uint8_t chroma_hor_col, chroma_ver_col;
};

int emu_mc_picture_rpr(uint16_t *const dst[3], const int dst_stride[3], int w, int h, int bitdepth,
                       const EmuMcRefs *refs, const EmuMcScaling *sc, const OVGPUMcCmd *cmds, int nb_cmds, const uint16_t *lmcs_fwd,
                       const OVGPUProfInfo *prof, int nb_prof, OVGPUDmvrMv *dmvr_out);

int emu_mc_picture(uint16_t *const dst[3], const int dst_stride[3], int w, int h, int bitdepth,
                   const EmuMcRefs *refs, const OVGPUMcCmd *cmds, int nb_cmds, const uint16_t *lmcs_fwd,
                   const OVGPUProfInfo *prof, int nb_prof, OVGPUDmvrMv *dmvr_out)
{
    return emu_mc_picture_rpr(dst, dst_stride, w, h, bitdepth, refs, NULL, cmds, nb_cmds, lmcs_fwd, prof, nb_prof, dmvr_out);
}

int emu_mc_picture_rpr(uint16_t *const dst[3], const int dst_stride[3], int w, int h, int bitdepth,
                       const EmuMcRefs *refs, const EmuMcScaling *sc, const OVGPUMcCmd *cmds, int nb_cmds, const uint16_t *lmcs_fwd,
                       const OVGPUProfInfo *prof, int nb_prof, OVGPUDmvrMv *dmvr_out)
{
    uint32_t sm[2] = { 0, 0 };
    if (sc)
        for (int l = 0; l < 2; l++)
            for (int i = 0; i < 16; i++) if (sc->scale[l][i][0] != (1 << 14) || sc->scale[l][i][1] != (1 << 14)) sm[l] |= 1u << i;
    // the work lists come from the product's device-side construction (McPrep*K), checked here against the host restatement
    // (a region's capacity: the <=16x16 items of the plain and resampling routes, one item per tool command, 16 tiles of 32x32 for a
    // 128x128 command of the big route -- a picture of two such commands has no other commands to borrow room from)
    const int max_items = mc_count_items(cmds, nb_cmds) + mc_count_items(cmds, nb_cmds, sm, MC_ROUTE_RPR) + 17 * nb_cmds + 1;
    std::vector<McItemDev> items((size_t)MC_NB_REGIONS * max_items);
    McPrepInfo info;
    memset(&info, 0, sizeof(info));
    {
        const int nblk = (nb_cmds + McPrepCountK::THREADS - 1) / McPrepCountK::THREADS;
        std::vector<int32_t> bsum((size_t)MC_NB_REGIONS * (nblk + 1));
        McPrepArgs pa = { (const McCmdDev *)cmds, (McCmdDev *)cmds, NULL, NULL, nb_cmds, nb_prof, nblk, max_items, { sm[0], sm[1] }, bsum.data(), items.data(), &info };
        if (nb_cmds) {
            ovd_emulate<McPrepCountK>(pa, ovd_dim3(nblk, 1, 1), McPrepCountK::SMEM_BYTES);
            ovd_emulate<McPrepScanK>(pa, ovd_dim3(1, 1, 1), McPrepScanK::SMEM_BYTES);
            ovd_emulate<McPrepScatterK>(pa, ovd_dim3(nblk, 1, 1), McPrepScatterK::SMEM_BYTES);
        }
    }
    for (int i = 0; i < nb_cmds; i++)
        if (mc_check_cmd(&cmds[i], nb_prof, sm)) return info.bad_rev == nb_cmds - i ? -(i + 1) : -1000000;      // first malformed command, both ways
    if (info.bad_rev || info.overflow) return -1000001;
    {   // the host's cheap count sizes the launches; the guard kernel publishes the limits the MC kernels stop at
        int cnt[MC_NB_REGIONS];
        mc_count_regions(cmds, nb_cmds, max_items, cnt, sm);
        for (int r = 0; r < MC_NB_REGIONS; r++) if (cnt[r] != info.count[r]) return -1000020 - r;
        McGuardArgs ga = { &info, w, h, { 0, 0 }, NULL };
        for (int l = 0; l < 2; l++) for (int i = 0; i < 16; i++) if (refs->plane[l][i][0]) ga.present[l] |= 1u << i;
        ovd_emulate<McGuardK>(ga, ovd_dim3(1, 1, 1), 0);
        if (info.status) return -1000030 - info.status;
    }
    const int n_items = info.count[0], n_bdof = info.count[1], n_prof = info.count[2], n_dmvr = info.count[3], n_sc = info.count[4], n_big = info.count[5];
    {
        std::vector<McItemDev> want(max_items);
        const int routes[MC_NB_REGIONS] = { OVGPU_MC_TOOL_PLAIN, OVGPU_MC_TOOL_BDOF, OVGPU_MC_TOOL_PROF, OVGPU_MC_TOOL_DMVR, MC_ROUTE_SMALL_C, MC_ROUTE_BIG, MC_ROUTE_RPR };
        for (int r = 0; r < MC_NB_REGIONS; r++) {
            const int n = r == 0 || r == 6 ? mc_build_items(want.data(), cmds, nb_cmds, sm, routes[r]) : r == 5 ? mc_build_big_items(want.data(), cmds, nb_cmds, sm)
                                                                               : mc_build_tool_items(want.data(), cmds, nb_cmds, routes[r], sm);
            if (n != info.count[r] || (n && memcmp(want.data(), items.data() + (size_t)r * max_items, sizeof(McItemDev) * n))) return -1000002 - r;
        }
        int mx = 0, my = 0;
        for (int i = 0; i < nb_cmds; i++) { mx = std::max(mx, cmds[i].x0 + (1 << cmds[i].log2_w)); my = std::max(my, cmds[i].y0 + (1 << cmds[i].log2_h)); }
        if (mx != info.max_x || my != info.max_y) return -1000010;
    }
    std::vector<DmvrMvDev> dm(n_dmvr + 1);
    McItemDev *const it_plain = items.data(), *const it_bdof = items.data() + max_items, *const it_prof = items.data() + 2 * (size_t)max_items,
              *const it_dmvr = items.data() + 3 * (size_t)max_items, *const it_sc = items.data() + 4 * (size_t)max_items, *const it_big = items.data() + 5 * (size_t)max_items;
    McK::Args a;
    memset(&a, 0, sizeof(a));
    for (int c = 0; c < 3; c++) { a.dst[c] = dst[c]; a.dst_pitch[c] = dst_stride[c]; }
    a.pic_w = w; a.pic_h = h; a.bitdepth = bitdepth;
    a.cmds = (const McCmdDev *)cmds; a.items = it_plain; a.nb_items = n_items;
    a.lmcs_fwd = lmcs_fwd;
    a.limit = info.limit;
    for (int l = 0; l < 2; l++)
        for (int i = 0; i < 16; i++) {
            a.refs.poc[l][i] = refs->poc[l][i];
            for (int c = 0; c < 3; c++) { a.refs.plane[l][i][c] = refs->plane[l][i][c]; a.refs.pitch[l][i][c] = refs->stride[l][i][c]; }
        }
    if (n_items) ovd_emulate<McK>(a, ovd_dim3((n_items + McK::WARPS - 1) / McK::WARPS, 1, 1), McK::SMEM_BYTES);
    const int n_rpr = info.count[6];
    if (n_rpr) {
        McRprK::Args t;
        memset(&t, 0, sizeof(t));
        t.mc = a; t.mc.items = items.data() + 6 * (size_t)max_items; t.mc.nb_items = n_rpr; t.mc.limit = info.limit + 6; t.prof = (const ProfInfoDev *)prof;
        for (int l = 0; l < 2; l++)
            for (int i = 0; i < 16; i++) {
                t.ref_w[l][i] = sc->ref_w[l][i]; t.ref_h[l][i] = sc->ref_h[l][i];
                t.scale[l][i][0] = sc->scale[l][i][0]; t.scale[l][i][1] = sc->scale[l][i][1];
            }
        t.hor_col = sc->chroma_hor_col; t.ver_col = sc->chroma_ver_col;
        ovd_emulate<McRprK>(t, ovd_dim3((n_rpr + McRprK::WARPS - 1) / McRprK::WARPS, 1, 1), McRprK::SMEM_BYTES);
    }
    if (n_big) {
        McK::Args t = a;
        t.items = it_big; t.nb_items = n_big; t.limit = info.limit + 5;
        ovd_emulate<McBigK>(t, ovd_dim3((n_big + McBigK::WARPS - 1) / McBigK::WARPS, 1, 1), McBigK::SMEM_BYTES);
    }
    if (n_bdof) {
        McK::Args t = a;
        t.items = it_bdof; t.nb_items = n_bdof; t.limit = info.limit + 1;
        ovd_emulate<BdofK>(t, ovd_dim3((n_bdof + BdofK::WARPS - 1) / BdofK::WARPS, 1, 1), BdofK::SMEM_BYTES);
    }
    if (n_prof) {
        ProfK::Args t;
        t.mc = a; t.mc.items = it_prof; t.mc.nb_items = n_prof; t.mc.limit = info.limit + 2; t.prof = (const ProfInfoDev *)prof;
        ovd_emulate<ProfK>(t, ovd_dim3((n_prof + ProfK::WARPS * ProfK::SBS - 1) / (ProfK::WARPS * ProfK::SBS), 1, 1), ProfK::SMEM_BYTES);
    }
    if (n_sc) {
        McK::Args t = a;
        t.items = it_sc; t.nb_items = n_sc; t.limit = info.limit + 4;
        ovd_emulate<McSmallCK>(t, ovd_dim3((n_sc + McSmallCK::WARPS * McSmallCK::SBS - 1) / (McSmallCK::WARPS * McSmallCK::SBS), 1, 1), McSmallCK::SMEM_BYTES);
    }
    if (n_dmvr) {
        DmvrK::Args t;
        t.mc = a; t.mc.items = it_dmvr; t.mc.nb_items = n_dmvr; t.mc.limit = info.limit + 3; t.out = dm.data();
        ovd_emulate<DmvrK>(t, ovd_dim3((n_dmvr + DmvrK::WARPS - 1) / DmvrK::WARPS, 1, 1), DmvrK::SMEM_BYTES);
        if (dmvr_out) memcpy(dmvr_out, dm.data(), sizeof(DmvrMvDev) * n_dmvr);
    }
    return n_items + n_bdof + n_prof + n_dmvr + n_sc + n_big + n_rpr;
}

int emu_lmcs_build_luts(OVGPULmcsLuts *luts, const OVGPULmcsData *d, int bitdepth)
{
    lmcs_build_luts_host(luts, d, bitdepth);
    return 0;
}

int emu_lmcs_reshape(uint16_t *plane, int stride, int w, int h, int bitdepth, const uint16_t *lut)
{
    LmcsReshapeK::Args a = { plane, stride, w, h, bitdepth, lut };
    ovd_emulate<LmcsReshapeK>(a, ovd_dim3((w + 255) / 256, (h + 7) / 8, 1), 0);
    return 0;
}

int emu_residual_picture(uint16_t *const dst[3], const int dst_stride[3], int bitdepth, const OVGPUTuCmd *cmds, int nb_cmds,
                         const int16_t *arena, size_t nb_coeffs, int dep_quant, int sh_ts_disabled, int ict_type)
{
    for (int i = 0; i < nb_cmds; i++) if (tu_check_cmd(&cmds[i], nb_coeffs)) return -(i + 1);
    int ns, nl;
    tu_build_items(NULL, NULL, &ns, &nl, cmds, nb_cmds);
    std::vector<TbItemDev> items(ns + nl + 1);
    tu_build_items(items.data(), items.data() + ns, &ns, &nl, cmds, nb_cmds);
    OVGPUTuSlice sl = { (uint8_t)dep_quant, (uint8_t)sh_ts_disabled, (uint8_t)ict_type, 0 };
    if (ns) emu_tr_launch<32>(dst, dst_stride, bitdepth, cmds, items.data(), ns, arena, &sl);
    if (nl) emu_tr_launch<128>(dst, dst_stride, bitdepth, cmds, items.data() + ns, nl, arena, &sl);
    return ns + nl;
}

int emu_dbf_picture(uint16_t *const planes[3], const int stride[3], int bitdepth, const OVGPUDbfMaps *maps)
{
    DbfK<0>::Args a;
    memset(&a, 0, sizeof(a));
    for (int c = 0; c < 3; c++) { a.plane[c] = planes[c]; a.pitch[c] = stride[c]; }
    a.bitdepth = bitdepth;
    for (int k = 0; k < DBF_NB_MAPS; k++) { a.m.ver[k] = maps->ver[k]; a.m.hor[k] = maps->hor[k]; }
    for (int c = 0; c < 3; c++) a.m.qp[c] = maps->qp[c];
    a.m.n_u = (1 << maps->log2_ctu_s) >> 2; a.m.w_u = maps->width >> 2; a.m.h_u = maps->height >> 2;
    a.m.beta_offset = maps->beta_offset; a.m.tc_offset = maps->tc_offset;
    ovd_dim3 grid((a.m.w_u + 31) / 32, (a.m.h_u + 7) / 8, 3);
    ovd_emulate<DbfK<0> >(a, grid, 0);
    DbfK<1>::Args b;
    memcpy(&b, &a, sizeof(b));
    ovd_emulate<DbfK<1> >(b, grid, 0);
    return 0;
}

static void emu_intra_launch(uint16_t *const planes[3], const int stride[3], int log2_ctu, int bitdepth, const OVGPUIntraCmd *cmds,
                             const uint32_t *order, int cnt, int plane)
{
    IntraK::Args a = { { planes[0], planes[1], planes[2] }, { stride[0], stride[1], stride[2] }, log2_ctu, bitdepth,
                       (const IntraCmdDev *)cmds, order, cnt, plane };
    ovd_emulate<IntraK>(a, ovd_dim3((cnt + IntraK::WARPS - 1) / IntraK::WARPS, plane ? 2 : 1, 1), IntraK::SMEM_BYTES);
}

struct EmuLevels {
    std::vector<uint32_t> order[2], start[2];
    IntraLevels lv;
    EmuLevels(int n) { lv.ws_clean = 0; lv.units_mask = 0; lv.ws_map = NULL; lv.ws_lvl = NULL; lv.ws_fill = NULL; for (int p = 0; p < 2; p++) { order[p].resize(n + 1); start[p].resize(n + 3); lv.order[p] = order[p].data(); lv.start[p] = start[p].data(); } }
};

// The product keeps its level / position maps all zero between pictures and clears only the blocks it wrote (intra.cu,
// IntraLevels::ws_clean).  Replays that protocol over two command lists on the same maps and compares the second result with
// a build on fresh maps.  Returns 0 when identical and the maps are left as the protocol says.
int emu_levels_reuse_check(const OVGPUIntraCmd *a, int na, const OVGPUIntraCmd *b, int nb, int w, int h)
{
    const int w_u = w >> 2, h_u = h >> 2, nmax = na > nb ? na : nb;
    const size_t units = (size_t)w_u * h_u;
    std::vector<uint32_t> pos_l(units, 0), pos_c(units, 0), ref_l(units, 0), ref_c(units, 0);
    std::vector<uint16_t> map(units, 0), lvl(2 * nmax + 2);
    std::vector<uint32_t> fill(nmax + 4);
    EmuLevels P(nmax), F(nb);
    P.lv.ws_map = map.data(); P.lv.ws_lvl = lvl.data(); P.lv.ws_fill = fill.data(); P.lv.ws_clean = 1;
    if (intra_build_levels(&P.lv, nmax + 1, a, na, w, h, pos_l.data(), pos_c.data())) return -1;
    for (size_t i = 0; i < units; i++) if (map[i]) return -2;
    intra_clear_units(pos_l.data(), w_u, a, na, OVGPU_INTRA_LUMA);
    intra_clear_units(pos_c.data(), w_u, a, na, OVGPU_INTRA_CHROMA);
    for (size_t i = 0; i < units; i++) if (pos_l[i] || pos_c[i]) return -3;
    if (intra_build_levels(&P.lv, nmax + 1, b, nb, w, h, pos_l.data(), pos_c.data())) return -4;
    if (intra_build_levels(&F.lv, nb + 1, b, nb, w, h, ref_l.data(), ref_c.data())) return -5;
    if (pos_l != ref_l || pos_c != ref_c) return -6;
    for (int p = 0; p < 2; p++) {
        if (P.lv.nb[p] != F.lv.nb[p]) return -7;
        for (int l = 0; l <= F.lv.nb[p]; l++) if (P.lv.start[p][l] != F.lv.start[p][l]) return -8;
        for (uint32_t q = 0; q < F.lv.start[p][F.lv.nb[p]]; q++) if (P.lv.order[p][q] != F.lv.order[p][q]) return -9;
    }
    for (size_t i = 0; i < units; i++) if (map[i]) return -10;
    return 0;
}

// the product's plain transform-block lists (tr_host.h): items[0 .. ns) small, items[ns .. ns + nl) large
int emu_tu_items(const OVGPUTuCmd *cmds, int n, TbItemDev *items, int *ns, int *nl)
{
    tu_build_items(NULL, NULL, ns, nl, cmds, n);
    tu_build_items(items, items + *ns, ns, nl, cmds, n);
    return *ns + *nl;
}

int emu_intra_picture(uint16_t *const planes[3], const int stride[3], int w, int h, int log2_ctu, int bitdepth, const OVGPUIntraCmd *cmds, int n)
{
    for (int i = 0; i < n; i++) if (intra_check_cmd(&cmds[i], w, h)) return -(i + 1);
    EmuLevels L(n);
    if (intra_build_levels(&L.lv, n + 1, cmds, n, w, h)) return -1000000;
    intra_for_each_launch(&L.lv, [&](int plane, int start, int cnt) {
        emu_intra_launch(planes, stride, log2_ctu, bitdepth, cmds, L.lv.order[plane] + start, cnt, plane);
        return 0;
    });
    return L.lv.nb[0] + L.lv.nb[1];
}

// ovgpu_recon_picture's schedule on the emulated kernels: the same host lists (intra_host.h, tr_host.h), the same per-item
// sequence as ovd_recon_kernel (recon_kernels.h) -- prediction phases, then the residual phases of the blocks the command
// covers -- executed serially in queue order, where every dependency is already satisfied (no waiting to emulate).
int emu_recon_picture(uint16_t *const planes[3], const int stride[3], int w, int h, int log2_ctu, int bitdepth,
                      const OVGPUIntraCmd *cmds, int n, const OVGPUTuCmd *tus, int nb_tus, const int16_t *arena, size_t nb_coeffs,
                      int dep_quant, int sh_ts_disabled, int ict_type,
                      const OVGPULmcsAnchor *anchors, int nb_anchors, const OVGPULmcsLuts *luts, const OVGPULmcsData *lmcs,
                      const OVGPUIspInfo *isp, int nb_isp)
{
    for (int i = 0; i < n; i++)
        if (intra_check_cmd(&cmds[i], w, h) || ((cmds[i].flags & OVGPU_INTRA_ISP) && (cmds[i].pad[0] | (cmds[i].pad[1] << 8)) >= nb_isp)) return -(i + 1);
    for (int i = 0; i < nb_tus; i++) if (tu_check_cmd(&tus[i], nb_coeffs)) return -(100000 + i);
    const int w_u = w >> 2, h_u = h >> 2;
    EmuLevels L(n);
    std::vector<uint32_t> pos_l((size_t)w_u * h_u), pos_c((size_t)w_u * h_u);
    if (intra_build_levels(&L.lv, n + 1, cmds, n, w, h, pos_l.data(), pos_c.data())) return -1000000;
    const int nb[2] = { (int)L.lv.start[0][L.lv.nb[0]], (int)L.lv.start[1][L.lv.nb[1]] };
    std::vector<TbItemDev> items(3 * (size_t)nb_tus + 1);
    std::vector<ReconItemDev> ri[2] = { std::vector<ReconItemDev>(nb[0] + 1), std::vector<ReconItemDev>(nb[1] + 1) };
    for (int p = 0; p < 2; p++) for (int q = 0; q < nb[p]; q++) ri[p][q].cmd = L.lv.order[p][q];
    ReconItemDev *rip[2] = { ri[0].data(), ri[1].data() };
    TuReconLists g;
    tu_build_recon_lists(items.data(), &g, rip, nb, tus, nb_tus, pos_l.data(), pos_c.data(), w_u);
    std::vector<int16_t> scales(nb_anchors + 1);
    OVGPUTuSlice sl = { (uint8_t)dep_quant, (uint8_t)sh_ts_disabled, (uint8_t)ict_type, 0 };
    emu_tr_scales = nb_anchors ? scales.data() : NULL;
    // the pending words as ReconPrepK leaves them (checked below: every unit is released exactly by its command)
    std::vector<uint32_t> pend(2 * (size_t)w_u * h_u, 0);
    int counter[2] = { -1, -1 };
    {
        ReconPrepK::Args pa = { (const IntraCmdDev *)cmds, { rip[0], rip[1] }, { nb[0], nb[1] }, { pend.data(), pend.data() + (size_t)w_u * h_u }, counter, 7u, w_u };
        const int mx = nb[0] > nb[1] ? nb[0] : nb[1];
        if (mx) ovd_emulate<ReconPrepK>(pa, ovd_dim3((mx + ReconPrepK::WARPS - 1) / ReconPrepK::WARPS, 2, 1), 0);
    }
    for (int p = 0; p < 2; p++) {
        if (p == 1 && nb_anchors) {
            LmcsScaleK::Args a = { planes[0], stride[0], bitdepth, (const LmcsAnchorDev *)anchors, nb_anchors, scales.data(), luts->wnd_bnd,
                                   lmcs->lmcs_min_bin_idx, 16 - lmcs->lmcs_delta_max_bin_idx,
                                   lmcs->lmcs_delta_sign_crs_flag ? -(int)lmcs->lmcs_delta_abs_crs : (int)lmcs->lmcs_delta_abs_crs };
            ovd_emulate<LmcsScaleK>(a, ovd_dim3((nb_anchors + LmcsScaleK::WARPS - 1) / LmcsScaleK::WARPS, 1, 1), LmcsScaleK::SMEM_BYTES);
        }
        if (g.n_small[p]) emu_tr_launch<32>(planes, stride, bitdepth, tus, items.data() + g.start_small[p], g.n_small[p], arena, &sl);
        if (g.n_large[p]) emu_tr_launch<128>(planes, stride, bitdepth, tus, items.data() + g.start_large[p], g.n_large[p], arena, &sl);
        for (int q = 0; q < nb[p]; q++) {
            const ReconItemDev &it = ri[p][q];
            const OVGPUIntraCmd &c = cmds[it.cmd];
            const int xu = c.x0 >> 2, yu = c.y0 >> 2, nw = (1 << c.log2_w) >> 2, nh = (1 << c.log2_h) >> 2;
            uint32_t *pd = pend.data() + (size_t)p * w_u * h_u;
            // what the kernel would wait for must already be released in queue order
            if (c.flags & OVGPU_INTRA_IBC) {
                int ux0, uy0, ux1, uy1;
                ibc_source_units(*(const IntraCmdDev *)&c, &ux0, &uy0, &ux1, &uy1);
                for (int y = uy0; y <= uy1; y++) for (int x = ux0; x <= ux1; x++) if (pd[(size_t)y * w_u + x] == 7u) return -2000000 - q;
            } else {
                const int um = ((1 << log2_ctu) >> 2) - 1, xl = xu & um, yl = yu & um;
                for (int k = 0; k < 2 * nh + 1 + 2 * nw; k++) {
                    int ux, uy, bit;
                    if (k < 2 * nh + 1) { ux = xu - 1; uy = yu - 1 + k; bit = yl + k; bit = bit < 64 ? (int)((c.avail_cols >> bit) & 1) : 0; }
                    else { const int col = k - (2 * nh + 1); ux = xu + col; uy = yu - 1; bit = xl + col + 1; bit = bit < 64 ? (int)((c.avail_rows >> bit) & 1) : 0; }
                    if (bit && ux >= 0 && uy >= 0 && ux < w_u && uy < h_u && pd[(size_t)uy * w_u + ux] == 7u) return -2000000 - q;
                }
            }
            IntraK::Args ia;
            memset(&ia, 0, sizeof(ia));
            for (int k = 0; k < 3; k++) { ia.plane[k] = planes[k]; ia.pitch[k] = stride[k]; }
            ia.log2_ctu = log2_ctu; ia.bitdepth = bitdepth; ia.cmds = (const IntraCmdDev *)cmds; ia.order = &it.cmd; ia.nb = 1;
            ia.comp0 = p; ia.warp_comps = p ? 2 : 1;
            const int ibc = (c.flags & OVGPU_INTRA_IBC) != 0;
            if (p == 0 && (c.flags & OVGPU_INTRA_ISP)) {
                // the partition loop of ovd_recon_kernel
                if (it.nb_tb > 4) return -5000000 - q;
                const IspInfoDev *wd = (const IspInfoDev *)&isp[c.pad[0] | (c.pad[1] << 8)];
                const int ver = (c.flags & OVGPU_INTRA_ISP_VER) != 0;
                int lp, nbp, qq = 0;
                isp_geometry(c.log2_w, c.log2_h, ver, &lp, &nbp);
                for (int i = 0; i < nbp; i++) {
                    const int off = i << lp;
                    if (!ver || !(off & 3)) {
                        ia.isp_ox = ver ? off : 0; ia.isp_oy = ver ? 0 : off;
                        ia.isp_lpw = ver ? (lp > 2 ? lp : 2) : c.log2_w; ia.isp_lph = ver ? c.log2_h : lp;
                        ia.isp_cols = ver ? wd->part[qq] : wd->common; ia.isp_rows = ver ? wd->common : wd->part[qq];
                        qq++;
                        ovd_emulate<IntraK>(ia, ovd_dim3(1, 1, 1), IntraK::SMEM_BYTES);
                    }
                    int kb = -1;
                    for (int k = 0; k < it.nb_tb; k++) {
                        const OVGPUTuCmd &t = tus[items[it.tb_start + k].cmd];
                        if ((ver ? t.x0 - c.x0 : t.y0 - c.y0) == off) kb = k;
                    }
                    if (kb < 0) continue;
                    if (it.large) emu_tr_launch<128>(planes, stride, bitdepth, tus, items.data() + it.tb_start + kb, 1, arena, &sl);
                    else emu_tr_launch<32>(planes, stride, bitdepth, tus, items.data() + it.tb_start + kb, 1, arena, &sl);
                }
            } else {
                // as in ovd_recon_kernel: the residual of the first blocks is transformed into shared memory first (phases 0-2),
                // held there while the prediction runs behind it in the same buffer, and added afterwards (phase 3)
                std::vector<uint64_t> smv(RECON_SMEM_BYTES / 8 + 2);
                unsigned char *sm = (unsigned char *)smv.data();
                memset(sm, 0xcd, RECON_SMEM_BYTES);
                const int nk0 = it.nb_tb < 4 ? it.nb_tb : 4;
                TrK<128>::Args tl = emu_tr_args<128>(planes, stride, bitdepth, tus, items.data() + it.tb_start, 1, arena, &sl);
                TrK<32>::Args ts = emu_tr_args<32>(planes, stride, bitdepth, tus, items.data() + it.tb_start, nk0, arena, &sl);
                if (nk0 && it.large) emu_phases<TrK<128> >(tl, sm, 0, 2);
                else if (nk0) emu_phases<TrK<32> >(ts, sm, 0, 2);
                if (ibc) {
                    uint16_t *pl[3] = { planes[0], planes[1], planes[2] };
                    ibc_copy(pl, stride, p, *(const IntraCmdDev *)&c, 0, 1);
                } else {
                    // the parameters are derived ahead of the phases, as the kernel does before it waits for the neighbours
                    for (int t = 0; t < IntraK::THREADS; t++) IntraK::predecode(ia, sm + RECON_TR_SMEM_BYTES, t, 0, 0);
                    ia.predecoded = 1;
                    emu_phases<IntraK>(ia, sm + RECON_TR_SMEM_BYTES, 0, 3);
                }
                if (nk0 && it.large) emu_phases<TrK<128> >(tl, sm, 3, 3);
                else if (nk0) emu_phases<TrK<32> >(ts, sm, 3, 3);
                if (it.large) for (int k = 1; k < it.nb_tb; k++) emu_tr_launch<128>(planes, stride, bitdepth, tus, items.data() + it.tb_start + k, 1, arena, &sl);
                else for (int k = 4; k < it.nb_tb; k += TrK<32>::GROUPS)
                    emu_tr_launch<32>(planes, stride, bitdepth, tus, items.data() + it.tb_start + k, it.nb_tb - k < TrK<32>::GROUPS ? it.nb_tb - k : TrK<32>::GROUPS, arena, &sl);
            }
            for (int y = yu; y < yu + nh; y++) for (int x = xu; x < xu + nw; x++) {
                if (pd[(size_t)y * w_u + x] != 7u) return -3000000 - q;
                pd[(size_t)y * w_u + x] = 0;
            }
        }
    }
    emu_tr_scales = NULL;
    for (size_t i = 0; i < pend.size(); i++) if (pend[i]) return -4000000;
    return L.lv.nb[0] + L.lv.nb[1];
}

}
