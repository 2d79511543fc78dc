/* ovvc_b200.h -- C ABI of the B200-native reconstruction backend for OpenVVC.
 *
 * This is the drop-in boundary for OpenVVC's rcn_* hot path (struct RCNFunctions,
 * libovvc/rcn_structures.h:499-694).  Plain C, plain pointers and sizes: it is what a
 * `rcn_init_functions_cuda()` added next to rcn_init_*_functions_avx2 (libovvc/rcn.c:242-254) binds to;
 * INTEGRATION.md shows that binding.  The CPU keeps CABAC / coding-tree parsing; each entry point below
 * replaces one coarse reference slot (named in its comment) and works on whole pictures that live in
 * HBM.  Every function returns 0 or a negative OVGPU_E* code; there is NO CPU fallback -- if the CUDA
 * device or kernel image is unusable the call fails.
 *
 * Pictures are planar 4:2:0, one uint16_t per sample (OVSample at BITDEPTH=10, libovvc/bitdepth.h:38).
 * Host-side planes use the OVFrame convention: data[3] + linesize[3] in BYTES (libovvc/ovframe.h:84-93).
 */
#ifndef OVVC_B200_H
#define OVVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OVGPU_EINVAL   (-22)   /* bad argument                                             */
#define OVGPU_ENOMEM   (-12)   /* host or device allocation failed                         */
#define OVGPU_ECUDA    (-5)    /* CUDA runtime / launch error (see ovgpu_last_error)       */
#define OVGPU_ENODEV   (-19)   /* no usable sm_100 device                                  */

typedef struct OVGPUCtx OVGPUCtx;          /* one per GPU / decoder thread: stream + scratch   */
typedef struct OVGPUPicture OVGPUPicture;  /* device-resident picture (3 planes)               */

/* ---- context ---------------------------------------------------------------------------------------- */
/* cuda_stream: a cudaStream_t to launch on (e.g. the caller's stream), or NULL to create a private one. */
int  ovgpu_ctx_create(OVGPUCtx **ctx, int device, void *cuda_stream);
void ovgpu_ctx_destroy(OVGPUCtx *ctx);
int  ovgpu_ctx_sync(OVGPUCtx *ctx);
const char *ovgpu_last_error(const OVGPUCtx *ctx);
/* number of kernels this context launched since creation (bench.py's "gpu_launches") */
uint64_t ovgpu_launch_count(const OVGPUCtx *ctx);
/* debugging aid: with OVGPU_HOST_TIMING=1 in the environment the library accumulates the host time of its commit / launch
 * sections (all contexts); this prints the table to stderr */
void ovgpu_host_timing_dump(void);
/* per-stage device timing with CUDA events on the context's stream */
enum OVGPUStage { OVGPU_STAGE_ALF_LUMA = 0, OVGPU_STAGE_ALF_CHROMA, OVGPU_STAGE_SAO, OVGPU_STAGE_DBF_V,
                  OVGPU_STAGE_DBF_H, OVGPU_STAGE_MC, OVGPU_STAGE_RESIDUAL, OVGPU_STAGE_INTRA,
                  OVGPU_STAGE_LMCS, OVGPU_NB_STAGES };
int  ovgpu_profile_enable(OVGPUCtx *ctx, int on);
/* accumulated milliseconds and launch count for a stage since the last reset (syncs the stream) */
int  ovgpu_profile_get(OVGPUCtx *ctx, int stage, double *ms, uint64_t *launches);
int  ovgpu_profile_reset(OVGPUCtx *ctx);
/* write `bytes` of device scratch (> L2) so the next timed launch starts cold; bench.py uses it */
int  ovgpu_flush_l2(OVGPUCtx *ctx);

/* Page-locks caller memory (cudaHostRegister) so that copies to / from it are asynchronous DMA: what the reference's frame pool
 * would do at allocation time (ovframepool_request_planes -> ovmempool_popelem -> ov_mallocz, ovframepool.c:194-231, mempool.c:73-86).
 * Registering a range twice is not an error.  _unregister before the memory is freed. */
int  ovgpu_host_register(OVGPUCtx *ctx, void *ptr, size_t bytes);
int  ovgpu_host_unregister(OVGPUCtx *ctx, void *ptr);

/* ---- pictures  (replaces the ov_mallocz'ed planes of ovframepool_request_planes, ovframepool.c:194) --- */
int  ovgpu_picture_create(OVGPUCtx *ctx, int width, int height, int bitdepth, OVGPUPicture **pic);
/* wrap caller-owned device memory (e.g. a torch tensor that NCCL broadcasts) */
int  ovgpu_picture_wrap(OVGPUCtx *ctx, int width, int height, int bitdepth,
                        void *const dev_planes[3], const size_t pitch_bytes[3], OVGPUPicture **pic);
void ovgpu_picture_destroy(OVGPUPicture *pic);
int  ovgpu_picture_upload(OVGPUCtx *ctx, OVGPUPicture *pic, const void *const data[3], const size_t linesize[3]);
int  ovgpu_picture_download(OVGPUCtx *ctx, const OVGPUPicture *pic, void *const data[3], const size_t linesize[3]);
int  ovgpu_picture_planes(const OVGPUPicture *pic, void *dev_planes[3], size_t pitch_bytes[3]);
int  ovgpu_picture_copy(OVGPUCtx *ctx, OVGPUPicture *dst, const OVGPUPicture *src);

/* ---- Output stage  (examples/dectest.c:376-406 write_decoded_frame_to_file: the output window of a frame as Y, Cb, Cr rows of
 *      16-bit samples; CI/ovnreg:82-85 hashes that file) --------------------------------------------------------------------- */
typedef struct OVGPUWindow {        /* struct Window (ovframe.h): offsets in CHROMA sample units, luma = offset << 1 */
    int32_t offset_lft, offset_rgt, offset_abv, offset_blw;
} OVGPUWindow;
/* Up-scaling of an output picture to the sequence's maximum size (replaces pp_sample_rate_conv, pp_pic_scale.c:252-377, as
 * pp_process_frame applies it under sei->upscale_flag, post_proc.c:117-128): struct ScalingInfo (ovdpb.h:85-93) of the picture --
 * scaling window offsets in chroma sample units, chroma sample siting.  win == NULL: no offsets, horizontally collocated chroma.
 * dst must not be smaller than src's scaling window (the reference's down-sampling branch is not reached from post_proc.c). */
typedef struct OVGPUScalingWindow { uint16_t left, right, top, bottom; uint8_t chroma_hor_col, chroma_ver_col; } OVGPUScalingWindow;
int  ovgpu_picture_upscale(OVGPUCtx *ctx, OVGPUPicture *dst, const OVGPUPicture *src, const OVGPUScalingWindow *win);
/* Film grain synthesis on an output picture (replaces fg_grain_apply_pic, pp_film_grain.c:814-977, as pp_process_frame runs it when
 * the stream carries a film grain characteristics SEI, post_proc.c:50-57, 100-103: de-blocking on, 10 bit): the fields of struct
 * OVSEIFGrain (nvcl_structures.h:115-140) the synthesis reads, as the SEI carries them.  Pictures whose luma or chroma size leaves a
 * block row / column of fewer than 8 samples (size % 16 in 1..7) are refused: the reference reads out of its planes there. */
typedef struct OVGPUFilmGrain {
    uint8_t log2_scale_factor;
    uint8_t comp_model_present_flag[3];
    uint8_t num_model_values_minus1[3];
    uint8_t intensity_interval_lower_bound[3][8], intensity_interval_upper_bound[3][8];
    int16_t comp_model_value[3][8][3];
} OVGPUFilmGrain;
typedef struct OVGPUFilmGrainSynth OVGPUFilmGrainSynth;   /* the 13 x 13 grain patterns in HBM + the per-picture generator states */
int  ovgpu_film_grain_create(OVGPUCtx *ctx, OVGPUFilmGrainSynth **synth);
void ovgpu_film_grain_destroy(OVGPUFilmGrainSynth *synth);
/* dst = src + grain; poc = the frame's POC (the generator's seed, pp_film_grain.c:869-872).  Asynchronous. */
int  ovgpu_film_grain_picture(OVGPUCtx *ctx, OVGPUFilmGrainSynth *synth, OVGPUPicture *dst, const OVGPUPicture *src,
                              const OVGPUFilmGrain *fg, int32_t poc);
/* the rows dectest would write, copied to host planes (win == NULL: the whole picture) */
int  ovgpu_picture_download_window(OVGPUCtx *ctx, const OVGPUPicture *pic, const OVGPUWindow *win, void *const data[3], const size_t linesize[3]);
/* running MD5 (RFC 1321) of those rows, computed on the device: ovgpu_md5_picture appends one picture (asynchronous, one GPU
 * thread: a serial hash, ~0.2 s per 4K picture -- a checking aid, not part of the timed path); ovgpu_md5_final returns the digest
 * of everything appended so far (blocking) without ending the hash */
typedef struct OVGPUMd5 OVGPUMd5;
int  ovgpu_md5_create(OVGPUCtx *ctx, OVGPUMd5 **md5);
void ovgpu_md5_destroy(OVGPUMd5 *md5);
int  ovgpu_md5_reset(OVGPUCtx *ctx, OVGPUMd5 *md5);
int  ovgpu_md5_picture(OVGPUCtx *ctx, OVGPUMd5 *md5, const OVGPUPicture *pic, const OVGPUWindow *win);
int  ovgpu_md5_final(OVGPUCtx *ctx, OVGPUMd5 *md5, uint8_t digest[16]);

/* ---- Reference hand-off between the GPUs of one node  (the device-side form of ovdpb_synchro_ref_decoded_ctus /
 *      rcn_inter_synchronization, dpb.c:1242-1269, rcn_inter.c:131-146: a picture that later pictures reference must
 *      be complete where they are reconstructed).  Pictures are sharded over the ranks (one process per GPU, picture k
 *      of the decoding order on rank k mod G, like the reference's frame threads, ovdec.c:189-248); when a reference
 *      picture is finished its owner PUBLISHES it to the ranks whose reference picture lists name it, and those ranks
 *      CONSUME it into their copy: grouped ncclSend / ncclRecv of the three planes over NVLink, queued on the calling
 *      context's stream (no host wait).  Every rank must issue its publish / consume calls in the same picture order.
 *      NCCL is loaded with dlopen on the first ovgpu_dist_* call. --------------------------------------------------- */
typedef struct OVGPUDist OVGPUDist;
#define OVGPU_DIST_ID_BYTES 128
/* rank 0: a fresh communicator id, to be handed to every rank by whatever means the host has */
int  ovgpu_dist_unique_id(void *id, size_t bytes);
/* collective over the nb_ranks processes (ncclCommInitRank) on the context's device */
int  ovgpu_dist_create(OVGPUCtx *ctx, int nb_ranks, int rank, const void *id, OVGPUDist **dist);
void ovgpu_dist_destroy(OVGPUDist *dist);
int  ovgpu_picture_publish(OVGPUCtx *ctx, OVGPUDist *dist, const OVGPUPicture *pic, const int *consumers, int nb_consumers);
int  ovgpu_picture_consume(OVGPUCtx *ctx, OVGPUDist *dist, OVGPUPicture *pic, int owner);
/* bytes this rank has queued for sending / receiving so far */
int  ovgpu_dist_traffic(const OVGPUDist *dist, uint64_t *bytes_sent, uint64_t *bytes_received);

/* ---- ALF / CC-ALF  (replaces alf.rcn_alf_filter_line, rcn_alf.c:1285, for all CTU lines of a picture,
 *                     and alf.rcn_alf_reconstruct_coeff_APS, rcn_alf.c:264) ------------------------------ */
#define OVGPU_ALF_NUM_SETS 24
#define OVGPU_ALF_SET_SIZE (4 * 25 * 13)

/* Byte-identical to the reference's RCNALF (rcn_alf.h:60-66) followed by the CC-ALF coefficients the
 * reference reads from the APS (OVALFData.alf_cc_mapped_coeff, nvcl_structures.h:668). */
typedef struct OVGPUAlfTables {
    int16_t luma_coeff[OVGPU_ALF_NUM_SETS][OVGPU_ALF_SET_SIZE];
    int16_t luma_clip [OVGPU_ALF_NUM_SETS][OVGPU_ALF_SET_SIZE];
    int16_t chroma_coeff[8][7];
    int16_t chroma_clip [8][7];
    int16_t cc_coeff[2][4][8];
} OVGPUAlfTables;

/* Byte-identical to ALFParamsCtu (dec_structures.h:277-283). */
typedef struct OVGPUAlfCtu {
    uint8_t ctb_alf_flag;     /* bit2 luma, bit1 cb, bit0 cr */
    uint8_t ctb_alf_idx;
    uint8_t cb_alternative;
    uint8_t cr_alternative;
} OVGPUAlfCtu;

/* APS fields rcn_alf_reconstruct_coeff_APS reads (OVALFData, nvcl_structures.h:649-669) */
typedef struct OVGPUAlfAps {
    uint8_t luma_clip_flag;
    uint8_t num_luma_filters;
    uint8_t luma_coeff_delta_idx[25];
    int16_t luma_coeff[25][13];
    int16_t luma_clip_idx[25][13];
    uint8_t chroma_clip_flag;
    uint8_t num_chroma_alts;
    int16_t chroma_coeff[8][7];
    int16_t chroma_clip_idx[8][7];
} OVGPUAlfAps;

/* Host-side table construction (tiny, runs once per slice on the CPU like the reference's). */
int ovgpu_alf_reconstruct_coeff_aps(OVGPUAlfTables *tab, const OVGPUAlfAps *luma_aps, int nb_luma_aps,
                                    const OVGPUAlfAps *chroma_aps, int bitdepth);

typedef struct OVGPUAlfParams OVGPUAlfParams;   /* pinned staging + device copy of one picture's ALF state */
int  ovgpu_alf_params_create(OVGPUCtx *ctx, int width, int height, int log2_ctu_s, OVGPUAlfParams **p);
void ovgpu_alf_params_destroy(OVGPUAlfParams *p);
/* cc_idx[c]: ALFInfo.ctb_cc_alf_filter_idx[c] (ctudec.h:207), one byte per CTU, may be NULL when disabled */
int  ovgpu_alf_params_upload(OVGPUCtx *ctx, OVGPUAlfParams *p, const OVGPUAlfTables *tab,
                             const OVGPUAlfCtu *ctu, const uint8_t *const cc_idx[2],
                             int cc_cb_enabled, int cc_cr_enabled);
/* dst = ALF(src): src is the post-SAO picture; dst must be a different picture. Asynchronous. */
/* tile (rect entry) edges padded like picture borders (the is_border flags of rcn_alf_filter_line -> rcn_extend_filter_region,
 * rcn_alf.c:1311-1330, rcn_ctu.c:316-510): classification, luma / chroma filters and CC-ALF fetch their samples clamped to the
 * CTU's tile; OVGPUTileGrid is declared with the SAO entry points below */
struct OVGPUTileGrid;
int  ovgpu_alf_params_set_tiles(OVGPUCtx *ctx, OVGPUAlfParams *p, const struct OVGPUTileGrid *tiles);
int  ovgpu_alf_picture(OVGPUCtx *ctx, OVGPUPicture *dst, const OVGPUPicture *src, const OVGPUAlfParams *p);
/* Whole call with HOST buffers: H2D of planes + params, ALF, D2H into dst planes, then stream sync. */
int  ovgpu_alf_frame_host(OVGPUCtx *ctx, void *const dst_data[3], const size_t dst_linesize[3],
                          const void *const src_data[3], const size_t src_linesize[3],
                          int width, int height, int log2_ctu_s, int bitdepth,
                          const OVGPUAlfTables *tab, const OVGPUAlfCtu *ctu, const uint8_t *const cc_idx[2],
                          int cc_cb_enabled, int cc_cr_enabled);

/* ---- SAO  (replaces sao.rcn_sao_first_pix_rows + sao.rcn_sao_filter_line, rcn_sao.c:190-293) ---------- */
/* Compact form of SAOParamsCtu (dec_structures.h:263-274): the fields rcn_sao.c actually reads. */
typedef struct OVGPUSaoCtu {
    int16_t offset_val[3][5];
    uint8_t band_position[3];
    uint8_t eo_class[3];
    uint8_t type_idx[3];      /* 0 off, 1 band, 2 edge */
    uint8_t pad;
} OVGPUSaoCtu;

typedef struct OVGPUSaoParams OVGPUSaoParams;
int  ovgpu_sao_params_create(OVGPUCtx *ctx, int width, int height, int log2_ctu_s, OVGPUSaoParams **p);
void ovgpu_sao_params_destroy(OVGPUSaoParams *p);
int  ovgpu_sao_params_upload(OVGPUCtx *ctx, OVGPUSaoParams *p, const OVGPUSaoCtu *ctu);
/* dst = SAO(src): src is the deblocked picture; dst must be a different picture. Asynchronous. */
/* Tiles (rect entries, struct RectEntryInfo): the reference runs its in-loop filters entry by entry and treats an entry's edges like
 * picture borders -- SAO leaves their samples alone for the edge classes that would look across them (rcn_sao.c:164-177, 211-214,
 * 241-245; the single-CTU-row quirk of rcn_sao_first_pix_rows applies per entry).  Boundaries in CTUs: col_bd[0] = 0 ...
 * col_bd[nb_cols] = the picture's width in CTUs, rows likewise; at most 20 x 20 tiles.  NULL or one tile: none (the default). */
typedef struct OVGPUTileGrid { int32_t nb_cols, nb_rows; uint16_t col_bd[22], row_bd[22]; } OVGPUTileGrid;
int  ovgpu_tile_grid_check(const OVGPUTileGrid *tiles, int nb_ctu_w, int nb_ctu_h);
int  ovgpu_sao_params_set_tiles(OVGPUCtx *ctx, OVGPUSaoParams *p, const OVGPUTileGrid *tiles);
int  ovgpu_sao_picture(OVGPUCtx *ctx, OVGPUPicture *dst, const OVGPUPicture *src, const OVGPUSaoParams *p);
int  ovgpu_sao_frame_host(OVGPUCtx *ctx, void *const dst_data[3], const size_t dst_linesize[3],
                          const void *const src_data[3], const size_t src_linesize[3],
                          int width, int height, int log2_ctu_s, int bitdepth, const OVGPUSaoCtu *ctu);

/* ---- LMCS  (replaces rcn_init_lmcs, lmcs_reshape_forward/backward; rcn_lmcs.c:274-361) ------------------ */
typedef struct OVGPULmcsData {        /* OVLMCSData (nvcl_structures.h:671-680) */
    uint8_t lmcs_min_bin_idx;
    uint8_t lmcs_delta_max_bin_idx;
    uint8_t lmcs_delta_cw_prec_minus1;
    uint16_t lmcs_delta_abs_cw[16];
    uint8_t lmcs_delta_sign_cw_flag[16];
    uint8_t lmcs_delta_abs_crs;
    uint8_t lmcs_delta_sign_crs_flag;
} OVGPULmcsData;

typedef struct OVGPULmcsLuts {        /* struct LMCSLUTs (rcn_lmcs.c:77-82) at 10 bit */
    uint16_t fwd_lut[1024];
    uint16_t bwd_lut[1024];
    uint16_t wnd_bnd[17];
} OVGPULmcsLuts;

/* host-side LUT construction (rcn_init_lmcs, rcn_lmcs.c:345) */
int  ovgpu_lmcs_build_luts(OVGPULmcsLuts *luts, const OVGPULmcsData *data, int bitdepth);
typedef struct OVGPULmcs OVGPULmcs;   /* device copy of the LUTs of one slice */
int  ovgpu_lmcs_create(OVGPUCtx *ctx, OVGPULmcs **l);
void ovgpu_lmcs_destroy(OVGPULmcs *l);
int  ovgpu_lmcs_upload(OVGPUCtx *ctx, OVGPULmcs *l, const OVGPULmcsLuts *luts);
/* the slice's LMCS syntax: window range and chroma scaling offset for the on-device chroma scale derivation */
int  ovgpu_lmcs_set_data(OVGPULmcs *l, const OVGPULmcsData *data);
/* luma plane of pic <- bwd_lut[luma] (forward != 0: fwd_lut); the reference does this per CTU after
 * rcn_write_ctu_to_frame (slicedec.c:745-750). In place. */
int  ovgpu_lmcs_reshape_picture(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPULmcs *l, int forward);

/* ---- Motion compensation  (replaces rcn_mcp, rcn_mcp_b, rcn_mcp_b_l, rcn_mcp_b_c; rcn_inter.c:2750-2966) - */
#define OVGPU_MC_LUMA      1   /* rcn_mcp_b_l (or rcn_mcp_b with both bits)            */
#define OVGPU_MC_CHROMA    2   /* rcn_mcp_b_c                                           */
#define OVGPU_MC_HPEL_AMVR 4   /* inter_ctx->prec_amvr == MV_PRECISION_HALF             */

#define OVGPU_MC_TOOL_PLAIN 0  /* rcn_mcp_b / rcn_mcp_b_l / rcn_mcp_b_c                                             */
#define OVGPU_MC_TOOL_BDOF  1  /* rcn_bdof_mcp_l (rcn_inter.c:1137): one <=16x16 luma sub-block of a BDOF CU, as called
                                * from vcl_coding_unit.c:2648-2672; bi-prediction, luma only                          */
#define OVGPU_MC_TOOL_PROF  2  /* rcn_prof_mcp_b_l (rcn_inter.c:2865): one 4x4 luma sub-block of an affine CU, as called
                                * from drv_affine_mvp.c:3355; aux = prof_dir (bits 0-1) | PROF table index << 8       */
#define OVGPU_MC_TOOL_DMVR  3  /* rcn_dmvr_mv_refine (rcn_inter.c:873): one <=16x16 sub-block of a DMVR CU, luma AND chroma, as
                                * called from vcl_coding_unit.c:2621; aux bit 0 = apply_bdof.  The refined MVs come back
                                * through ovgpu_mc_batch_read_dmvr (the caller stores them for TMVP, :2628-2645).        */
#define OVGPU_MC_TOOL_GPM   4  /* rcn_gpm_b (rcn_inter.c:3118): geometric partitioning of a CU of 8..64 samples a side, luma AND chroma:
                                * two uni-predictions blended sample by sample with the weights of split `aux & 63`
                                * (rcn_gpm_weights_and_steps :3080, rcn_init_gpm_params rcn_gpm.c:150, put_weighted_gpm_bi_pixels
                                * rcn_mc.c:1631).  mv0 / ref_idx0 address list (aux >> 8) & 1 (gpm_ctx.inter_dir0 - 1), mv1 / ref_idx1
                                * list (aux >> 9) & 1; inter_dir = 3, default BCW                                           */
typedef struct OVGPUDmvrMv { int32_t mv0_x, mv0_y, mv1_x, mv1_y; } OVGPUDmvrMv;
/* struct PROFInfo (rcn_inter.c:1128-1134): the 16 per-sample delta-MV scales of an affine CU and list */
typedef struct OVGPUProfInfo {
    int16_t dmv_scale_h_0[16], dmv_scale_v_0[16], dmv_scale_h_1[16], dmv_scale_v_1[16];
} OVGPUProfInfo;

/* One recorded call; x0/y0 are picture-absolute luma samples ((ctb_x << log2_ctu_s) + x0 in the reference). */
typedef struct OVGPUMcCmd {
    int16_t x0, y0;
    uint8_t log2_w, log2_h;
    uint8_t inter_dir;            /* 1 = L0, 2 = L1, 3 = bi                              */
    uint8_t flags;
    int8_t  ref_idx0, ref_idx1;
    uint8_t bcw_idx_plus1;        /* OVMV.bcw_idx_plus1 of mv0 (ctudec.h:281-289)        */
    uint8_t tool;                 /* OVGPU_MC_TOOL_*                                     */
    int32_t mv0_x, mv0_y, mv1_x, mv1_y;   /* 1/16 luma sample                            */
    uint32_t aux;
} OVGPUMcCmd;

/* Reference picture lists of the current slice: inter_ctx->rpl0/rpl1 (ctudec.h:349-350) and their POCs. */
typedef struct OVGPUMcRefs {
    const OVGPUPicture *rpl[2][16];
    int32_t poc[2][16];
} OVGPUMcRefs;

/* Reference picture resampling (replaces the rcn_mcp_rpr_* / rcn_mc_rpr_b_* / rcn_mc_rpr_prof_b_l paths the stock slots take
 * when a scale factor differs from 1, rcn_inter.c:1990-2747, 2749-2966): the slice's scale factors
 * inter_ctx->scale_fact_rpl0 / 1 [ref_idx][hor, ver] (ctudec.h:351-352, derived in ctudec_compute_refs_scaling, ctudec.c:42-84;
 * 1 << 14 = same size, 0 = entry not in use) and the SPS's chroma sample siting (OVPicture.scale_info, dpb.c:1167-1168). */
typedef struct OVGPUMcScaling {
    uint16_t scale[2][16][2];
    uint8_t chroma_hor_col, chroma_ver_col, pad[2];
} OVGPUMcScaling;

typedef struct OVGPUMcBatch OVGPUMcBatch;   /* pinned staging + device copy of one picture's PU list */
int  ovgpu_mc_batch_create(OVGPUCtx *ctx, int max_cmds, int max_items, OVGPUMcBatch **b);
void ovgpu_mc_batch_destroy(OVGPUMcBatch *b);
/* splits the PUs into <=16x16 work items on the host and copies both lists to the device (async) */
int  ovgpu_mc_batch_upload(OVGPUCtx *ctx, OVGPUMcBatch *b, const OVGPUMcCmd *cmds, int nb_cmds);
/* same, with the PROF table that OVGPU_MC_TOOL_PROF commands index (may be NULL when nb_prof == 0) */
int  ovgpu_mc_batch_upload2(OVGPUCtx *ctx, OVGPUMcBatch *b, const OVGPUMcCmd *cmds, int nb_cmds,
                            const OVGPUProfInfo *prof, int nb_prof);
/* The scale factors the batch's NEXT upload / commit is routed by (NULL: none scaled, the default): PLAIN, PROF and GPM commands
 * that name a scaled reference go to the resampling kernel, BDOF / DMVR commands that do are malformed (the decoder switches
 * both tools off for such CUs).  ovgpu_mc_picture then accepts references whose size differs from the picture's for exactly
 * those entries.  Host-side state only; call it before the commit of every picture whose factors differ from the last one's. */
int  ovgpu_mc_batch_set_scaling(OVGPUCtx *ctx, OVGPUMcBatch *b, const OVGPUMcScaling *scaling);
/* Zero-copy staging: _staging returns the batch's PINNED command (and PROF table) buffers for the recorder to fill,
 * _commit validates, builds the work lists in one pass and issues the asynchronous H2D copies. */
int  ovgpu_mc_batch_staging(OVGPUCtx *ctx, OVGPUMcBatch *b, OVGPUMcCmd **cmds, OVGPUProfInfo **prof);
int  ovgpu_mc_batch_commit(OVGPUCtx *ctx, OVGPUMcBatch *b, int nb_cmds, int nb_prof);
/* Refined MVs of the DMVR commands of the last ovgpu_mc_picture call on this batch, in command order: D2H copy into
 * `out` (up to max entries) + stream sync; returns the number of DMVR commands or a negative error. */
int  ovgpu_mc_batch_read_dmvr(OVGPUCtx *ctx, OVGPUMcBatch *b, OVGPUDmvrMv *out, int max);
/* The commands are validated, routed and split into work items ON THE DEVICE (three small kernels behind the commit; the host
 * only counts items to size the launches and never waits).  What that walk refuses -- a malformed command, more work items
 * than the batch holds, commands outside the picture, a reference the caller did not supply -- makes the batch's kernels do
 * nothing and is reported, as OVGPU_EINVAL with the message the old host-side walk gave, by this call (blocking: it waits
 * for the batch's last run) and by the next ovgpu_mc_picture / ovgpu_mc_batch_read_dmvr on the batch. */
int  ovgpu_mc_batch_status(OVGPUCtx *ctx, OVGPUMcBatch *b);
/* writes the prediction of every command into dst; lmcs may be NULL (no forward reshape). Asynchronous. */
int  ovgpu_mc_picture(OVGPUCtx *ctx, OVGPUPicture *dst, const OVGPUMcRefs *refs, const OVGPUMcBatch *b,
                      const OVGPULmcs *lmcs);

/* ---- Residual  (replaces tmp.rcn_tu_st / tmp.rcn_tu_c, rcn_transform_tree.c:1228-1390, and below them
 *                 dequant_tb_4x4 (rcn_dequant.c:160), process_lfnst(_luma) (drv_lfnst.c:120), tr.func / tr.dc
 *                 (rcn_transform.c:45-598) and ict.add / ict.ict (rcn_residuals.c:46-331)) ------------------------- */
#define OVGPU_TU_IMPLICIT_MTS 1   /* luma TB of an intra, non-MIP CU while ctudec->mts_implicit (rcn_transform_tree.c:435) */
#define OVGPU_TU_LFNST_L      2   /* TUInfo.lfnst_flag applies to the luma TB                                       */
#define OVGPU_TU_LFNST_C      4   /* TUInfo.lfnst_flag applies to the chroma TBs (dual-tree chroma, rcn_tu_c)       */
#define OVGPU_TU_SCALE_ANCHOR 8   /* the LMCS chroma scale is the one the GPU derives for anchor `scale_anchor` of the batch
                                   * (ovgpu_tu_batch_set_anchors) instead of `chroma_scale`                         */

#define OVGPU_TU_ISP         16   /* luma transform block of an ISP partition (rcn_tu_isp_{v,h}, rcn_transform_tree.c:970-1085): luma only, log2_w /
                                   * log2_h from 0 (1xN, Nx1 blocks) to 6, x0 / y0 any; cu_mts_idx bit 0 / bit 1 = DST-VII (else DCT-II) for the
                                   * horizontal / vertical pass (recon_isp_subtree_{v,h} :1110-1111, 1178-1179); lfnst_mode_l = the value of
                                   * drv_lfnst_mode_l(coding block size, intra mode), i.e. already wide-angle mapped; coefficients of blocks
                                   * thinner than 4 in raster order with stride 1 << log2_w.  Only through ovgpu_recon_picture. */
/* One recorded transform unit; x0/y0 picture-absolute luma samples, log2_w/h the luma size (chroma TBs are half).
 * Per-TB arrays are indexed like TUInfo.tb_info (rcn_transform_tree.c:56-66): [0] Cb or joint CbCr, [1] Cr, [2] luma. */
typedef struct OVGPUTuCmd {
    int16_t x0, y0;
    uint8_t log2_w, log2_h;
    uint8_t cbf_mask;             /* 0x10 luma, 0x08 joint CbCr, 0x02 Cb, 0x01 Cr                  */
    uint8_t tr_skip_mask;
    uint8_t cu_mts_flag, cu_mts_idx;
    uint8_t flags;                /* OVGPU_TU_*                                                    */
    uint8_t lfnst_idx;
    uint8_t qp[4];                /* dequant_luma / cb / cr / joint_cb_cr .qp (ctudec.h:829-833)   */
    uint8_t qp_ts[4];             /* dequant_*_skip.qp                                             */
    uint16_t last_pos[3];
    int16_t chroma_scale;         /* lmcs_info.lmcs_chroma_scale at call time (2048 = neutral)     */
    uint8_t lfnst_mode_l;         /* ctudec->intra_mode (PLANAR for MIP CUs)                       */
    uint8_t lfnst_mode_c;         /* chroma mode with DM / CCLM resolved to the collocated luma mode (drv_lfnst.c:97-104) */
    uint16_t scale_anchor;        /* with OVGPU_TU_SCALE_ANCHOR: index into the batch's anchor list */
    uint32_t coeff_off[3];        /* int16 offsets into the coefficient arena; data in the reference's buffer order:
                                   * 4x4 sub-blocks of 16 contiguous values, min(32,w)/4 per row (rcn_dequant.c:243-253) */
    uint32_t bdpcm;               /* BDPCM of the transform-skip blocks (rcn_bdpcm_tb, rcn_transform_tree.c:630-688; cu_flags
                                   * flg_intra_bdpcm_*): bit 0 luma, bit 1 luma direction (1 = vertical accumulation), bit 2 chroma, bit 3
                                   * chroma direction; the blocks must carry their tr_skip_mask bit */
    uint64_t sig_sb_map[3];
} OVGPUTuCmd;

/* slice-level switches of the residual path */
typedef struct OVGPUTuSlice {
    uint8_t dep_quant;            /* ctudec->residual_coding_l == residual_coding_dpq              */
    uint8_t sh_ts_disabled;       /* ctudec->sh_ts_disabled                                        */
    uint8_t ict_type;             /* (ph_joint_cbcr_sign << 1) | chroma residual scaling (slicedec.c:1354-1359) */
    uint8_t pad;
} OVGPUTuSlice;

typedef struct OVGPUTuBatch OVGPUTuBatch;   /* pinned staging + device copy of one picture's TU list and coefficients */
int  ovgpu_tu_batch_create(OVGPUCtx *ctx, int max_cmds, size_t max_coeffs, OVGPUTuBatch **b);
void ovgpu_tu_batch_destroy(OVGPUTuBatch *b);
int  ovgpu_tu_batch_upload(OVGPUCtx *ctx, OVGPUTuBatch *b, const OVGPUTuCmd *cmds, int nb_cmds,
                           const int16_t *coeffs, size_t nb_coeffs, const OVGPUTuSlice *slice);
/* Zero-copy staging (what the north-star recorder uses): _staging returns the batch's PINNED command / coefficient
 * buffers (after the previous H2D copy out of them has completed) for the parser to write into directly; _commit
 * validates the commands, builds the work lists and issues the asynchronous H2D copies.  upload = staging + memcpy +
 * commit. */
int  ovgpu_tu_batch_staging(OVGPUCtx *ctx, OVGPUTuBatch *b, OVGPUTuCmd **cmds, int16_t **coeffs);
int  ovgpu_tu_batch_commit(OVGPUCtx *ctx, OVGPUTuBatch *b, int nb_cmds, size_t nb_coeffs, const OVGPUTuSlice *slice);
/* LMCS chroma residual scaling needs lmcs_info.lmcs_chroma_scale, which the reference derives from RECONSTRUCTED luma
 * (rcn_lmcs_compute_chroma_scale, rcn_lmcs.c:320-343) whenever a coding unit starts on a 64-aligned position
 * (coding_unit, vcl_coding_unit.c:725-730) -- the CPU parser cannot know it.  The recorder instead lists those positions
 * ("anchors") with the two 16-bit availability masks the reference reads there ((hfield[y0 >> 2] >> ((x0 >> 2) + 1)) & 0xffff
 * and the vfield counterpart, CTU-local x0 / y0) and tags every transform unit with the anchor in force
 * (OVGPU_TU_SCALE_ANCHOR + scale_anchor); ovgpu_recon_picture derives the scales on the GPU once the luma plane is
 * reconstructed and before any chroma residual is added.  Needs the LMCS object given to ovgpu_recon_picture to carry the
 * slice's OVGPULmcsData (ovgpu_lmcs_set_data). */
typedef struct OVGPULmcsAnchor {
    int16_t x0, y0;               /* picture-absolute luma position (multiples of 64 inside the CTU) */
    uint16_t abv_mask, lft_mask;
} OVGPULmcsAnchor;
int  ovgpu_tu_batch_set_anchors(OVGPUCtx *ctx, OVGPUTuBatch *b, const OVGPULmcsAnchor *anchors, int nb_anchors);
/* pic += residual of every TU (pic holds the prediction). Asynchronous. */
int  ovgpu_residual_picture(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPUTuBatch *b);

/* ---- Deblocking filter  (replaces df.rcn_dbf_ctu / df.rcn_dbf_truncated_ctu, rcn_df.c:2170-2228, for all CTUs) ---- */
/* Picture-level form of what struct DBFInfo (ctudec.h:130-170) holds per CTU at rcn_dbf_ctu time, i.e. after the CPU
 * side has run dbf_ctu_preproc_{h,v} (MV / reference mismatch -> bs1, rcn_df.c:1822-1874).  With n_u = CTU size / 4,
 * w_u = width / 4, h_u = height / 4 (4x4 luma units):
 *   ver[m][cy * (w_u + 1) + X]  bit j  <=>  vertical edge at unit column X (0..w_u), unit row cy * n_u + j
 *   hor[m][cx * (h_u + 1) + Y]  bit j  <=>  horizontal edge at unit row Y (0..h_u), unit column cx * n_u + j
 *   qp[c][Y * w_u + X]                      QP of the unit (qp_map_y / qp_map_cb / qp_map_cr)
 * The recorder copies, for CTU (cx, cy) and local index i: ver[m][..cx * n_u + i] = <map>.ver[i] and
 * hor[m][..cy * n_u + i] = <map>.hor[i] >> 2; map m is one of: */
enum OVGPUDbfMap {
    OVGPU_DBF_EDGE_TU = 0,   /* ctb_bound_ver / ctb_bound_hor [8 + i]        transform / coding block edges      */
    OVGPU_DBF_EDGE_SB,       /* aff_edg_ver / aff_edg_hor [8 + i]            affine 8x8 sub-block edges          */
    OVGPU_DBF_BS1,           /* bs1_map                                      coded residual or MV/ref mismatch   */
    OVGPU_DBF_BS2,           /* bs2_map                                      intra / CIIP                        */
    OVGPU_DBF_AFFINE,        /* affine_map [1 + i]: the unit at index i belongs to an affine CU                  */
    OVGPU_DBF_EDGE_C,        /* ctb_bound_ver_c / ctb_bound_hor_c [8 + i]                                         */
    OVGPU_DBF_BS1_CB, OVGPU_DBF_BS1_CR, OVGPU_DBF_BS2_C,
    OVGPU_DBF_NB_MAPS
};
typedef struct OVGPUDbfMaps {
    int32_t width, height, log2_ctu_s;
    int16_t beta_offset, tc_offset;       /* DBFInfo.beta_offset / tc_offset */
    const uint32_t *ver[OVGPU_DBF_NB_MAPS];
    const uint32_t *hor[OVGPU_DBF_NB_MAPS];
    const uint8_t *qp[3];
} OVGPUDbfMaps;

typedef struct OVGPUDbfParams OVGPUDbfParams;   /* pinned staging + device copy of one picture's deblocking maps */
int  ovgpu_dbf_params_create(OVGPUCtx *ctx, int width, int height, int log2_ctu_s, OVGPUDbfParams **p);
void ovgpu_dbf_params_destroy(OVGPUDbfParams *p);
int  ovgpu_dbf_params_upload(OVGPUCtx *ctx, OVGPUDbfParams *p, const OVGPUDbfMaps *maps);
/* zero-copy variant: _staging fills `maps` with pointers into the object's pinned buffer (writable by the caller),
 * _commit issues the asynchronous H2D copy */
int  ovgpu_dbf_params_staging(OVGPUCtx *ctx, OVGPUDbfParams *p, OVGPUDbfMaps *maps);
int  ovgpu_dbf_params_commit(OVGPUCtx *ctx, OVGPUDbfParams *p, int beta_offset, int tc_offset);
/* deblocks pic in place: all vertical edges, then all horizontal edges (luma and chroma). Asynchronous. */
int  ovgpu_dbf_picture(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPUDbfParams *p);

/* ---- Intra prediction  (replaces the intra_pred slot: vvc_intra_pred, rcn_intra.c:484-560, and the intra_pred_c slot:
 *      vvc_intra_pred_chroma, rcn_intra.c:1015-1182, planar / DC / angular modes) -------------------------------------- */
/* One intra coding unit: a luma block and / or its 4:2:0 Cb + Cr blocks of half the size (no ISP / BDPCM in this revision).  avail_cols / avail_rows are the two words the
 * reference reads from its progress bit-field at call time: rcn_ctx.progress_field.vfield[x0 >> 2] and .hfield[y0 >> 2]
 * (CTU-local x0 / y0; bit 0 = the unit in the neighbouring CTU line, bit 1 + u = 4x4 unit u, rcn_structures.h:81-85).
 * In a single tree progress_field_c holds the same words (rcn_tu_st marks both with the same rectangles), so the chroma
 * blocks use them too.  They are decode-order information, so the recorder copies them; the samples come from the
 * picture on the GPU. */
#define OVGPU_INTRA_LUMA   1    /* flags: the command predicts a luma block with `mode` */
#define OVGPU_INTRA_CHROMA 2    /* flags: Cb and Cr blocks of half the size are predicted with `mode_c` (single tree: same availability) */
#define OVGPU_INTRA_CIIP   4    /* flags: combined inter / intra: planar prediction blended over the inter prediction already in the
                                * picture, (wt * intra + (4 - wt) * inter + 2) >> 2 (rcn_ciip_weighted_sum, rcn_inter.c:2968-3010;
                                * put_weighted_ciip_pixels, rcn_mc.c:1611-1627); chroma keeps the inter prediction when log2_w <= 2 */
#define OVGPU_INTRA_CCLM_COLLOCATED 8   /* flags: sps_chroma_vertical_collocated_flag (5-tap cross luma down-sampling for CCLM) */
#define OVGPU_INTRA_MIP    16   /* flags: matrix-based intra prediction of the luma block (mip.rcn_intra_mip slot, rcn_intra_mip.c:386-403;
                                * called from rcn_intra_tu, rcn_transform_tree.c:1403-1406); `mode` = ctudec->cu_opaque: bit 7 = transposed,
                                * bits 0-5 = matrix index (< 16 for 4x4, < 8 for the 8x8 class, < 6 otherwise); no MRL / CIIP */
#define OVGPU_INTRA_ISP      32  /* flags: intra sub-partitions (tmp.recon_isp_subtree_{h,v}, rcn_transform_tree.c:1087-1205; intra_pred_isp,
                                  * rcn_intra.c:567-643): the command is the WHOLE luma coding block; its 2 or 4 partitions are predicted and
                                  * reconstructed one after the other inside ovgpu_recon_picture, with the transform units flagged
                                  * OVGPU_TU_ISP that lie inside the block.  pad[0] | pad[1] << 8 = index into the batch's OVGPUIspInfo table.
                                  * Not accepted by ovgpu_intra_picture (prediction alone is meaningless for ISP). */
#define OVGPU_INTRA_IBC     128  /* flags: intra block copy (rcn_ibc_l / rcn_ibc_c slots, rcn_ibc.c:8-140): the block (luma and / or both chroma
                                  * blocks, per the LUMA / CHROMA flags) is a copy of samples of the same picture reconstructed before it;
                                  * block vector in luma samples, int16 little endian: x in pad[0..1], y in pad[2..3] (the reference's
                                  * position in its CTU ring buffer = picture position (x0 + bv.x, y0 + bv.y) for every vector a stream
                                  * may carry); chroma reads ((x0 + bv.x) >> 1, (y0 + bv.y) >> 1).  Only through ovgpu_recon_picture. */
#define OVGPU_INTRA_ISP_VER  64  /* flags: vertical split (cu_opaque == 2), else horizontal */
/* The availability words vvc_intra_pred_isp reads for an ISP coding unit at CTU-local (x0, y0): the progress field does not
 * change during the CU (the whole block is marked before the sub-tree, vcl_transform_unit.c:1878), so with q counting the
 * prediction calls (one per partition; vertical split of 1- or 2-wide partitions: one per 4 columns):
 *   vertical:   common = hfield[y0 >> 2],  part[q] = vfield[x_q >> 2]
 *   horizontal: common = vfield[x0 >> 2],  part[q] = hfield[(y_q >> 2) + !!(y_q % 4)] */
typedef struct OVGPUIspInfo { uint64_t common, part[4]; } OVGPUIspInfo;
typedef struct OVGPUIntraCmd {
    int16_t x0, y0;               /* picture-absolute luma samples */
    uint8_t log2_w, log2_h;       /* luma size */
    uint8_t mode;                 /* luma: 0 planar, 1 DC, 2..66 angular (before wide-angle remapping) */
    uint8_t mode_c;               /* chroma: same range (DM already resolved), 67 CCLM, 68 MDLM left, 69 MDLM top
                                   * (rcn_intra_cclm.c; read the luma block of the same command, so luma goes first) */
    uint8_t flags;
    uint8_t mrl_idx;              /* luma reference line 0..2 (ctudec->cu_opaque with flg_mrl_flag; vvc_intra_pred_multi_ref,
                                   * rcn_intra.c:645-786); needs y0 >= 4 inside the CTU */
    uint8_t ciip_wt;              /* CIIP: 1 + above CU is intra + left CU is intra (1..3) */
    uint8_t pad[5];               /* pad[0..3]: ISP table index / block vector (see the flags); pad[4]: BDPCM prediction of the block
                                   * (cu_flags flg_intra_bdpcm_*; vvc_intra_pred rcn_intra.c:511-522, vvc_intra_pred_chroma :1043-1074: the pure
                                   * horizontal / vertical kernel on the unfiltered references, no PDPC) -- bit 0 luma, bit 1 luma direction
                                   * (1 = vertical), bit 2 chroma, bit 3 chroma direction, the layout of OVGPUTuCmd.bdpcm */
    uint64_t avail_cols, avail_rows;
} OVGPUIntraCmd;

typedef struct OVGPUIntraBatch OVGPUIntraBatch;
int  ovgpu_intra_batch_create(OVGPUCtx *ctx, int max_cmds, OVGPUIntraBatch **b);
void ovgpu_intra_batch_destroy(OVGPUIntraBatch *b);
/* cmds in decode order; the host sorts them into dependency levels (blocks of a level only read blocks of lower levels) */
int  ovgpu_intra_batch_upload(OVGPUCtx *ctx, OVGPUIntraBatch *b, const OVGPUIntraCmd *cmds, int nb_cmds,
                              int pic_w, int pic_h, int log2_ctu_s);
/* same, with the ISP table that OVGPU_INTRA_ISP commands index (may be NULL when nb_isp == 0) */
int  ovgpu_intra_batch_upload2(OVGPUCtx *ctx, OVGPUIntraBatch *b, const OVGPUIntraCmd *cmds, int nb_cmds,
                               const OVGPUIspInfo *isp, int nb_isp, int pic_w, int pic_h, int log2_ctu_s);
int  ovgpu_intra_batch_levels(const OVGPUIntraBatch *b);
/* predicts every block into pic (one launch per dependency level); neighbours are read from pic. Asynchronous. */
int  ovgpu_intra_picture(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPUIntraBatch *b);

/* ---- Reconstruction order  (replaces the CU loop of the reference: every coding unit is predicted and then gets its
 *      residual through rcn_transform_tree before the next one is predicted, vcl_coding_unit.c -> rcn_res_wrap,
 *      rcn_transform_tree.c:1433; intra blocks read neighbours holding prediction + residual) --------------------------
 * One call per picture: motion compensation of `mc` (may be NULL); then plane by plane (luma, then Cb + Cr -- luma never
 * reads chroma): the residual of the transform blocks that lie outside intra commands, then ONE dataflow launch in which
 * every intra command of `intra` (may be NULL) is predicted (CIIP: blended) as soon as the units its reference samples come
 * from are reconstructed, followed at once by the residual of the transform blocks it covers.  Between the planes the LMCS
 * chroma scales of the batch's anchors are derived from the reconstructed luma.  Transform units are matched to intra
 * commands by position.  The result equals predicting and adding the residual CU by CU in decode order.  Asynchronous.
 * A TU batch used here is re-ordered for this schedule; commit it again before handing it to ovgpu_residual_picture. */
int  ovgpu_recon_picture(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPUMcRefs *refs, const OVGPUMcBatch *mc,
                         const OVGPULmcs *lmcs, OVGPUTuBatch *tu, OVGPUIntraBatch *intra);

#ifdef __cplusplus
}
#endif
#endif /* OVVC_B200_H */
