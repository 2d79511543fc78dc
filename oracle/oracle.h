/* oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement ("port") of OpenVVC's rcn_* reconstruction hot path, written from scratch at
 * whole-picture granularity (the reference works CTU by CTU with saved rows/columns; the results are
 * the same because both implement the VVC decoding order DBF-V -> DBF-H -> SAO -> ALF).
 * Every function cites the reference file:line it restates.  The port is pinned by differential tests
 * against the compiled, unmodified reference (oracle/_ref, built by oracle/Makefile) and by the golden
 * fixtures under tests/golden/ that were generated from that reference build.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * liboracle.so.  The product path (openvvc_b200/) never does.
 *
 * All samples are 10-bit (or 8-bit) values stored in uint16_t, planar 4:2:0, strides in SAMPLES.
 */
#ifndef OVO_ORACLE_H
#define OVO_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- ALF ------------------------------------------------------------------------------------------- */
#define OVO_ALF_NUM_SETS      24   /* 16 fixed + 8 APS   (rcn_alf.h:48-51)                  */
#define OVO_ALF_SET_SIZE      (4 * 25 * 13) /* transposes x classes x coeffs (rcn_alf.h:62)  */

/* Same content as the reference's RCNALF (rcn_alf.h:60-66) plus the CC-ALF coefficients that the reference
 * reads straight from the APS (rcn_alf.c:1414).  The same struct is the product's table format
 * (include/ovvc_b200.h, OVGPUAlfTables) -- keep them identical. */
typedef struct OVOAlfTables {
    int16_t luma_coeff[OVO_ALF_NUM_SETS][OVO_ALF_SET_SIZE];
    int16_t luma_clip [OVO_ALF_NUM_SETS][OVO_ALF_SET_SIZE];   /* clip VALUES (1024/32/8/2 at 10 bit) */
    int16_t chroma_coeff[8][7];
    int16_t chroma_clip [8][7];
    int16_t cc_coeff[2][4][8];                                /* [cb,cr][filt_idx-1][7 used]        */
} OVOAlfTables;

/* Per-CTU parameters, raster order, one byte each (ALFParamsCtu dec_structures.h:277-283 and
 * ALFInfo.ctb_cc_alf_filter_idx ctudec.h:207). */
typedef struct OVOAlfCtuParams {
    const uint8_t *ctb_alf_flag;   /* bit2 luma, bit1 cb, bit0 cr */
    const uint8_t *ctb_alf_idx;    /* 0..23 luma filter set       */
    const uint8_t *cb_alternative;
    const uint8_t *cr_alternative;
    const uint8_t *cc_idx_cb;      /* 0 = off, 1..4               */
    const uint8_t *cc_idx_cr;
    uint8_t cc_cb_enabled, cc_cr_enabled;
} OVOAlfCtuParams;

/* APS description used to (re)build tables: mirrors the fields of OVALFData that
 * rcn_alf_reconstruct_coeff_APS reads (nvcl_structures.h:649-669, rcn_alf.c:196-281). */
typedef struct OVOAlfAps {
    uint8_t luma_clip_flag;
    uint8_t num_luma_filters;                 /* alf_luma_num_filters_signalled_minus1 + 1 */
    uint8_t luma_coeff_delta_idx[25];
    int16_t luma_coeff[25][13];
    int16_t luma_clip_idx[25][13];
    uint8_t chroma_clip_flag;
    uint8_t num_chroma_alts;                  /* alf_chroma_num_alt_filters_minus1 + 1     */
    int16_t chroma_coeff[8][7];
    int16_t chroma_clip_idx[8][7];
} OVOAlfAps;

void ovo_alf_build_tables(OVOAlfTables *t, const OVOAlfAps *luma_aps, int nb_luma_aps,
                          const OVOAlfAps *chroma_aps, int bitdepth);

/* Whole-picture ALF (+CC-ALF). src = post-SAO picture, dst = output (must not alias src). */
void ovo_alf_picture(uint16_t *const dst[3], const int dst_stride[3],
                     const uint16_t *const src[3], const int src_stride[3],
                     int pic_w, int pic_h, int log2_ctu_s, int bitdepth,
                     const OVOAlfTables *tab, const OVOAlfCtuParams *ctu);
/* same with a tile grid (OVOTileGrid below; NULL = one tile): tile edges are padded like picture borders */
struct OVOTileGrid;
void ovo_alf_picture_tiles(uint16_t *const dst[3], const int dst_stride[3], const uint16_t *const src[3], const int src_stride[3],
                           int pic_w, int pic_h, int log2_ctu_s, int bitdepth, const OVOAlfTables *tab, const OVOAlfCtuParams *ctu,
                           const struct OVOTileGrid *tg);

/* Classification only (class | transpose<<5 per 4x4 block, raster over the picture's 4x4 grid). */
void ovo_alf_classify_picture(uint8_t *cls, const uint16_t *src, int stride,
                              int pic_w, int pic_h, int log2_ctu_s, int bitdepth);

/* ---- SAO ------------------------------------------------------------------------------------------- */
/* Compact per-CTU SAO parameters (SAOParamsCtu dec_structures.h:263-274). */
typedef struct OVOSaoCtu {
    int16_t offset_val[3][5];
    uint8_t band_position[3];
    uint8_t eo_class[3];
    uint8_t type_idx[3];       /* 0 off, 1 band, 2 edge */
    uint8_t pad;
} OVOSaoCtu;

/* Tiles (rect entries): the reference runs its in-loop filters entry by entry and treats the edges of an entry like picture
 * borders -- SAO leaves their samples alone for the edge classes that would look across (rcn_sao.c:164-177, 211-214), ALF pads by
 * replication (rcn_extend_filter_region, rcn_ctu.c:316-510).  Boundaries in CTUs: col_bd[0] = 0 ... col_bd[nb_cols] = picture width in
 * CTUs, rows likewise.  NULL = one tile. */
typedef struct OVOTileGrid { int32_t nb_cols, nb_rows; uint16_t col_bd[22], row_bd[22]; } OVOTileGrid;
void ovo_sao_picture_tiles(uint16_t *const dst[3], const int dst_stride[3], const uint16_t *const src[3], const int src_stride[3],
                           int pic_w, int pic_h, int log2_ctu_s, int bitdepth, const OVOSaoCtu *ctu, const OVOTileGrid *tiles);
void ovo_sao_picture(uint16_t *const dst[3], const int dst_stride[3],
                     const uint16_t *const src[3], const int src_stride[3],
                     int pic_w, int pic_h, int log2_ctu_s, int bitdepth,
                     const OVOSaoCtu *ctu);

/* ---- Motion compensation ---------------------------------------------------------------------------- */
/* One command = one call of rcn_mcp_b / rcn_mcp_b_l / rcn_mcp_b_c / rcn_mcp (rcn_inter.c:2750-2966) with
 * picture-absolute luma coordinates.  Identical to OVGPUMcCmd (include/ovvc_b200.h). */
#define OVO_MC_LUMA      1
#define OVO_MC_CHROMA    2
#define OVO_MC_HPEL_AMVR 4      /* inter_ctx->prec_amvr == MV_PRECISION_HALF (rcn_inter.c:572) */
#define OVO_MC_TOOL_PLAIN 0     /* rcn_mcp_b / rcn_mcp_b_l / rcn_mcp_b_c                                             */
#define OVO_MC_TOOL_BDOF  1     /* rcn_bdof_mcp_l (rcn_inter.c:1137): one <= 16x16 luma sub-block, bi-prediction      */
#define OVO_MC_TOOL_PROF  2     /* rcn_prof_mcp_b_l (rcn_inter.c:2865): one 4x4 affine luma sub-block;                 *
                                 * aux = prof_dir (bits 0-1) | index into the PROF table << 8                          */
#define OVO_MC_TOOL_DMVR  3     /* rcn_dmvr_mv_refine (rcn_inter.c:873): one <= 16x16 sub-block, luma AND chroma;          *
                                 * aux bit 0 = apply_bdof.  The refined MVs are an OUTPUT (fed back to TMVP on the CPU).   */
#define OVO_MC_TOOL_GPM   4     /* rcn_gpm_b (rcn_inter.c:3118): aux = split_dir | list of mv0 << 8 | list of mv1 << 9      */
typedef struct OVODmvrMv { int32_t mv0_x, mv0_y, mv1_x, mv1_y; } OVODmvrMv;
typedef struct OVOProfInfo {    /* struct PROFInfo, rcn_inter.c:1128-1134 */
    int16_t dmv_scale_h_0[16], dmv_scale_v_0[16], dmv_scale_h_1[16], dmv_scale_v_1[16];
} OVOProfInfo;
typedef struct OVOMcCmd {
    int16_t x0, y0;
    uint8_t log2_w, log2_h;
    uint8_t inter_dir;          /* 1 L0, 2 L1, 3 bi */
    uint8_t flags;
    int8_t  ref_idx0, ref_idx1;
    uint8_t bcw_idx_plus1;
    uint8_t tool;               /* OVO_MC_TOOL_*                                                          */
    int32_t mv0_x, mv0_y, mv1_x, mv1_y;   /* 1/16 luma sample */
    uint32_t aux;
} OVOMcCmd;

typedef struct OVOMcRefs {
    const uint16_t *plane[2][16][3];   /* [list][ref_idx][component] */
    int stride[2][16][3];              /* samples */
    int32_t poc[2][16];
} OVOMcRefs;

/* lmcs_fwd_lut: 1<<bitdepth entries or NULL (rcn_lmcs_reshape_forward, rcn_lmcs.c:287) */
void ovo_mc_picture(uint16_t *const dst[3], const int dst_stride[3], int pic_w, int pic_h, int bitdepth,
                    const OVOMcRefs *refs, const OVOMcCmd *cmds, int nb_cmds, const uint16_t *lmcs_fwd_lut);
/* same with the PROF table that tool-2 commands index */
void ovo_mc_picture2(uint16_t *const dst[3], const int dst_stride[3], int pic_w, int pic_h, int bitdepth,
                     const OVOMcRefs *refs, const OVOMcCmd *cmds, int nb_cmds, const uint16_t *lmcs_fwd_lut,
                     const OVOProfInfo *prof_tab);
/* same, and the refined MVs of the DMVR commands (in command order) are written to dmvr_out (may be NULL) */
void ovo_mc_picture3(uint16_t *const dst[3], const int dst_stride[3], int pic_w, int pic_h, int bitdepth,
                     const OVOMcRefs *refs, const OVOMcCmd *cmds, int nb_cmds, const uint16_t *lmcs_fwd_lut,
                     const OVOProfInfo *prof_tab, OVODmvrMv *dmvr_out);

/* Reference picture resampling (rcn_inter.c:1990-2747): per reference the scale factors the slice derived
 * (inter_ctx->scale_fact_rpl0/1[ref_idx][hor, ver], ctudec.c:40-84; 1 << 14 = same size) and the reference's own size;
 * the two chroma siting flags are the SPS's (dpb.c:1167-1168).  PLAIN, GPM and PROF commands may name scaled references
 * (BDOF / DMVR are switched off by the decoder when a reference is scaled). */
typedef struct OVOMcScaling {
    uint16_t scale[2][16][2];
    int32_t ref_w[2][16], ref_h[2][16];
    uint8_t chroma_hor_col, chroma_ver_col;
} OVOMcScaling;
void ovo_mc_picture_rpr(uint16_t *const dst[3], const int dst_stride[3], int pic_w, int pic_h, int bitdepth,
                        const OVOMcRefs *refs, const OVOMcScaling *sc, const OVOMcCmd *cmds, int nb_cmds,
                        const uint16_t *lmcs_fwd_lut, const OVOProfInfo *prof_tab);

/* ---- Output stage: up-scaling of a decoded picture to the sequence's maximum size (pp_sample_rate_conv,
 * pp_pic_scale.c:252-377; struct ScalingInfo, ovdpb.h:85-93: scaling window offsets in chroma sample units) -------- */
typedef struct OVOScalingWindow { uint16_t left, right, top, bottom; uint8_t chroma_hor_col, chroma_ver_col; } OVOScalingWindow;
int ovo_pp_upscale_plane(uint16_t *dst, int dst_stride, int dst_w, int dst_h, const uint16_t *src, int src_stride, int src_w, int src_h,
                         const OVOScalingWindow *win, int luma, int bitdepth);

/* ---- Output stage: film grain synthesis (fg_grain_apply_pic, pp_film_grain.c:814-977).  The fields of struct OVSEIFGrain
 * (nvcl_structures.h:115-140) the synthesis reads, as the SEI carries them (the chroma conversion of the model values is part
 * of the synthesis). ------------------------------------------------------------------------------------------------- */
typedef struct OVOFilmGrain {
    uint8_t log2_scale_factor;
    uint8_t comp_model_present_flag[3];
    uint8_t num_model_values_minus1[3];
    uint8_t intensity_interval_lower_bound[3][8], intensity_interval_upper_bound[3][8];
    int16_t comp_model_value[3][8][3];
} OVOFilmGrain;
void ovo_fg_database(int8_t *db /* [13][13][64][64] */, int deblock);
int  ovo_fg_size_ok(int w, int h);
int  ovo_fg_apply(uint16_t *const dst[3], const uint16_t *const src[3], int w, int h, const OVOFilmGrain *fg, int32_t poc, const int8_t *db);

/* ---- LMCS ------------------------------------------------------------------------------------------- */
typedef struct OVOLmcsData {          /* OVLMCSData, nvcl_structures.h:671-680 */
    uint8_t lmcs_min_bin_idx;
    uint8_t lmcs_delta_max_bin_idx;
    uint8_t lmcs_delta_cw_prec_minus1;
    uint16_t lmcs_delta_abs_cw[16];
    uint8_t lmcs_delta_sign_cw_flag[16];
    uint8_t lmcs_delta_abs_crs;
    uint8_t lmcs_delta_sign_crs_flag;
} OVOLmcsData;

typedef struct OVOLmcsLuts {          /* struct LMCSLUTs, rcn_lmcs.c:77-82 (10-bit: 1024 entries) */
    uint16_t fwd_lut[1024];
    uint16_t bwd_lut[1024];
    uint16_t wnd_bnd[17];
} OVOLmcsLuts;

void ovo_lmcs_build_luts(OVOLmcsLuts *luts, const OVOLmcsData *data, int bitdepth);
int ovo_lmcs_chroma_scale(const uint16_t *src, int stride, unsigned abv_mask, unsigned lft_mask,
                          const OVOLmcsLuts *luts, const OVOLmcsData *d, int bitdepth);
void ovo_lmcs_reshape_plane(uint16_t *plane, int stride, int w, int h, const uint16_t *lut, int bitdepth);

/* ---- Residual: dequantisation, LFNST, inverse transform, residual add ------------------------------------- */
/* One command = one tmp.rcn_tu_st (rcn_transform_tree.c:1228) or tmp.rcn_tu_c (:1349, dual-tree chroma) call;
 * picture-absolute luma coordinates, luma-sized log2_w/h.  Identical to OVGPUTuCmd (include/ovvc_b200.h).
 * Index order of the per-TB arrays follows TUInfo.tb_info: [0] Cb (or the joint Cb-Cr block), [1] Cr, [2] luma. */
#define OVO_TU_IMPLICIT_MTS 1   /* luma: intra, non-MIP CU in a slice with mts_implicit (rcn_transform_tree.c:435) */
#define OVO_TU_LFNST_L      2   /* TUInfo.lfnst_flag applies to the luma TB  (rcn_tu_st / rcn_tu_l)              */
#define OVO_TU_LFNST_C      4   /* TUInfo.lfnst_flag applies to the chroma TBs (rcn_tu_c only)                    */
#define OVO_TU_ISP         16   /* luma transform block of an ISP partition (rcn_tu_isp_{v,h}, rcn_transform_tree.c:970-1085): log2_w / log2_h
                                 * may be 0 or 1, x0 / y0 any; cu_mts_idx bit 0 / bit 1 = DST-VII (else DCT-II) horizontally / vertically;
                                 * lfnst_mode_l = drv_lfnst_mode_l(coding block size, intra mode), already wide-angle mapped */
typedef struct OVOTuCmd {
    int16_t x0, y0;
    uint8_t log2_w, log2_h;      /* luma TU size */
    uint8_t cbf_mask;            /* 0x10 luma, 0x08 joint CbCr, 0x02 Cb, 0x01 Cr */
    uint8_t tr_skip_mask;        /* same bits */
    uint8_t cu_mts_flag, cu_mts_idx;
    uint8_t flags;               /* OVO_TU_* */
    uint8_t lfnst_idx;           /* TUInfo.lfnst_idx: 0/1 selects the kernel of the set */
    uint8_t qp[4];               /* dequant_luma / cb / cr / joint_cb_cr .qp (ctudec.h:829-833) */
    uint8_t qp_ts[4];            /* the *_skip variants */
    uint16_t last_pos[3];
    int16_t chroma_scale;        /* lmcs_chroma_scale (2048 = neutral) */
    uint8_t lfnst_mode_l;        /* intra mode LFNST derives its set from: ctudec->intra_mode, PLANAR for MIP   */
    uint8_t lfnst_mode_c;        /* chroma: resolved mode (DM / CCLM replaced by the collocated luma mode)      */
    uint16_t scale_anchor;       /* GPU side only (OVGPU_TU_SCALE_ANCHOR); the oracle takes chroma_scale        */
    uint32_t coeff_off[3];       /* int16 offsets into the coefficient arena */
    uint32_t bdpcm;              /* bit 0 luma BDPCM, bit 1 its direction, bit 2 chroma BDPCM, bit 3 its direction (rcn_bdpcm_tb) */
    uint64_t sig_sb_map[3];
} OVOTuCmd;

void ovo_residual_picture(uint16_t *const dst[3], const int dst_stride[3], int pic_w, int pic_h, int bitdepth,
                          const OVOTuCmd *cmds, int nb_cmds, const int16_t *arena,
                          int dep_quant, int sh_ts_disabled, int ict_type);

/* ---- Deblocking filter ---------------------------------------------------------------------------------- */
/* One block of the synthetic decoder state (openvvc_b200/synth_dbf.py: DBF_BLK) */
typedef struct OVODbfBlk {
    int16_t x0, y0;
    uint8_t log2_w, log2_h;
    uint8_t flags;               /* 1 intra, 2 cbf luma, 4 cbf cb, 8 cbf cr, 16 MV/ref mismatch on its edges, 32 affine */
    uint8_t qp_y, qp_cb, qp_cr;
    uint16_t pad;
} OVODbfBlk;

/* Picture-level maps; identical to OVGPUDbfMaps (include/ovvc_b200.h). n_u = CTU size / 4, w_u = width / 4,
 * h_u = height / 4.  ver[m][cy * (w_u + 1) + X]: bit j <=> unit row cy * n_u + j of vertical edge column X;
 * hor[m][cx * (h_u + 1) + Y]: bit j <=> unit column cx * n_u + j of horizontal edge row Y; qp[c][Y * w_u + X]. */
enum { OVO_DBF_EDGE_TU = 0, OVO_DBF_EDGE_SB, OVO_DBF_BS1, OVO_DBF_BS2, OVO_DBF_AFFINE, OVO_DBF_EDGE_C,
       OVO_DBF_BS1_CB, OVO_DBF_BS1_CR, OVO_DBF_BS2_C, OVO_DBF_NB_MAPS };
typedef struct OVODbfMaps {
    int32_t width, height, log2_ctu_s;
    int16_t beta_offset, tc_offset;
    uint32_t *ver[OVO_DBF_NB_MAPS];
    uint32_t *hor[OVO_DBF_NB_MAPS];
    uint8_t *qp[3];
} OVODbfMaps;

/* in place on the picture: all vertical edges, then all horizontal edges, luma and chroma (rcn_df.c) */
void ovo_dbf_picture(uint16_t *const planes[3], const int stride[3], int bitdepth, const OVODbfMaps *maps);

/* ---- Intra prediction ------------------------------------------------------------------------------------ */
/* One rcn_funcs.intra_pred call (vvc_intra_pred, rcn_intra.c:484-560): luma block, no ISP / MRL / MIP / BDPCM.
 * avail_cols / avail_rows are progress_field.vfield[x0 >> 2] / .hfield[y0 >> 2] at call time (CTU-local, bit 0 = the
 * unit of the neighbouring CTU line, bit 1 + u = 4x4 unit u): syntax-order information recorded on the CPU. */
/* flags == 0: placeholder for a coding unit that is not intra (oracle only): its samples stay, the reference harness
 * marks it in the progress field like the decoder does for every reconstructed CU */
#define OVO_INTRA_LUMA   1    /* flags: the command predicts a luma block with `mode` */
#define OVO_INTRA_CHROMA 2    /* flags: Cb and Cr blocks of half the size are predicted with `mode_c` (single tree: same availability) */
#define OVO_INTRA_CIIP   4    /* flags: combined inter / intra: planar prediction blended over the inter prediction already in the
                                * picture, (wt * intra + (4 - wt) * inter + 2) >> 2 (rcn_ciip_weighted_sum, rcn_inter.c:2968-3010;
                                * put_weighted_ciip_pixels, rcn_mc.c:1611-1627); chroma keeps the inter prediction when log2_w <= 2 */
#define OVO_INTRA_CCLM_COLLOCATED 8   /* flags: sps_chroma_vertical_collocated_flag (5-tap cross luma down-sampling for CCLM) */
#define OVO_INTRA_MIP    16   /* flags: matrix-based intra prediction of the luma block; `mode` = ctudec->cu_opaque: bit 7 transposed,
                                * bits 0-5 the matrix index (rcn_intra_mip, rcn_intra_mip.c:386-403) */
#define OVO_INTRA_ISP      32  /* flags: intra sub-partitions (recon_isp_subtree_{h,v}, rcn_transform_tree.c:1087-1205): the command is the whole
                                * luma coding block; its partitions are predicted AND reconstructed one after the other, so the
                                * command carries its transform blocks (isp index in pad[0] | pad[1] << 8) */
#define OVO_INTRA_IBC      128 /* flags: intra block copy (rcn_ibc_l / rcn_ibc_c, rcn_ibc.c:8-140): the block is a copy of already reconstructed
                                * samples of the same picture at the block vector pad[0..1] (x) / pad[2..3] (y), int16 little endian, luma samples */
#define OVO_INTRA_ISP_VER  64  /* flags: vertical split (cu_opaque == 2), else horizontal */
typedef struct OVOIntraCmd {
    int16_t x0, y0;               /* picture-absolute luma samples */
    uint8_t log2_w, log2_h;       /* luma size */
    uint8_t mode;                 /* luma: 0 planar, 1 DC, 2..66 angular (before wide-angle remapping) */
    uint8_t mode_c;               /* chroma: same range (DM already resolved), 67 CCLM, 68 MDLM left, 69 MDLM top
                                   * (rcn_intra_cclm.c; read the luma block of the same command, so luma goes first) */
    uint8_t flags;
    uint8_t mrl_idx;              /* luma reference line 0..2 (ctudec->cu_opaque with flg_mrl_flag; vvc_intra_pred_multi_ref,
                                   * rcn_intra.c:645-786); needs y0 >= 4 inside the CTU */
    uint8_t ciip_wt;              /* CIIP: 1 + above CU is intra + left CU is intra (1..3) */
    uint8_t pad[5];
    uint64_t avail_cols, avail_rows;
} OVOIntraCmd;
void ovo_intra_picture(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h, int log2_ctu_s, int bitdepth,
                       const OVOIntraCmd *cmds, int nb_cmds);
/* Per ISP coding unit: the availability words vvc_intra_pred_isp reads (rcn_intra.c:586-596; the progress field does not
 * change during the CU -- the whole block is marked before the sub-tree, vcl_transform_unit.c:1878) and the transform blocks
 * of the (at most 4) partitions in order, cbf_mask == 0 where a partition has none.  TU flags carry OVO_TU_ISP. */
typedef struct OVOIspInfo {
    uint64_t common;             /* vertical split: hfield[y0 >> 2];  horizontal split: vfield[x0 >> 2]                         */
    uint64_t part[4];            /* per prediction call q: vertical: vfield[x_q >> 2]; horizontal: hfield[(y_q >> 2) + !!(y_q % 4)] */
    OVOTuCmd tu[4];
} OVOIspInfo;
typedef struct OVOIspSlice { uint8_t dep_quant, mts_enabled, pad[2]; } OVOIspSlice;
/* like ovo_intra_picture; commands flagged OVO_INTRA_ISP are reconstructed (prediction + residual per partition) */
void ovo_intra_picture2(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h, int log2_ctu_s, int bitdepth,
                        const OVOIntraCmd *cmds, int nb_cmds, const OVOIspInfo *isp, const int16_t *arena, int dep_quant);
/* one ISP coding unit, luma: partitions predicted and reconstructed in order; cm = CTU size - 1 */
void ovo_isp_cu(uint16_t *plane, int stride, int bitdepth, int cm, const OVOIntraCmd *c, const OVOIspInfo *info,
                const int16_t *arena, int dep_quant);
/* residual of one transform block flagged OVO_TU_ISP into res (raster, w x h) */
void ovo_isp_residual(int16_t *res, const int16_t *arena, const OVOTuCmd *c, int dep_quant, int bitdepth);

/* ---- Reconstruction in CU order (port_recon.c): prediction, then residual, coding unit by coding unit -------- */
typedef struct OVOLmcsAnchor { int16_t x0, y0; uint16_t abv_mask, lft_mask; } OVOLmcsAnchor;   /* == OVGPULmcsAnchor */
/* anchors (may be NULL): 64-aligned CU origins where the reference derives lmcs_chroma_scale; the value stays in force for
 * the following transform units flagged with bit 3 (OVGPU_TU_SCALE_ANCHOR), the others take their chroma_scale field */
void ovo_recon_picture(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h, int log2_ctu_s, int bitdepth,
                       const OVOIntraCmd *cmds, const OVOTuCmd *tus, int nb_cu, const int16_t *arena,
                       int dep_quant, int sh_ts_disabled, int ict_type,
                       const OVOLmcsAnchor *anchors, int nb_anchors, const OVOLmcsLuts *luts, const OVOLmcsData *lmcs);
/* same; commands flagged OVO_INTRA_ISP take their luma from the ISP table (isp_subtree_{v,h}: the luma sub-tree, then the CU's
 * chroma blocks through rcn_tu_c, vcl_transform_unit.c:1536-1556): tus[i] then only contributes its chroma blocks */
void ovo_recon_picture2(uint16_t *const planes[3], const int stride[3], int pic_w, int pic_h, int log2_ctu_s, int bitdepth,
                        const OVOIntraCmd *cmds, const OVOTuCmd *tus, int nb_cu, const int16_t *arena,
                        int dep_quant, int sh_ts_disabled, int ict_type,
                        const OVOLmcsAnchor *anchors, int nb_anchors, const OVOLmcsLuts *luts, const OVOLmcsData *lmcs,
                        const OVOIspInfo *isp);

#ifdef __cplusplus
}
#endif
#endif
