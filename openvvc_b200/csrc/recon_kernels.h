// recon_kernels.h -- dependency-driven reconstruction of a picture's intra blocks (sm_100a).
//
// Reference behaviour: the CU loop -- every coding unit is predicted and gets its residual before the next one is predicted
// (coding_unit -> prediction -> rcn_transform_tree, libovvc/vcl_coding_unit.c, libovvc/rcn_transform_tree.c:1433), because
// intra prediction reads reconstructed neighbours.  On the CPU that is a serial chain per entry thread; here it is a
// dataflow over the whole picture: the host sorts the intra commands of a plane topologically (by dependency level,
// intra_host.h) and ONE launch per plane executes them -- every CTA repeatedly takes the next command from a device queue
// (atomic counter, in sorted order), waits until the 4x4 units its reference samples come from are reconstructed (a
// "pending" word per unit, acquire loads), runs the prediction phases of IntraK and the residual phases of TrK for that
// block, and publishes its units (fence + release).  A command only ever waits on commands that precede it in the queue
// order, and those have been taken by CTAs that are running, so the scheme cannot deadlock however few CTAs are resident.
//
// The kernel body re-uses the phase functions of intra_kernels.h / tr_kernels.h unchanged; tests/emu runs the same item
// sequence serially (one "CTA", no waiting needed in queue order).
#pragma once
#include "intra_kernels.h"
#include "tr_kernels.h"

struct ReconItemDev {           // one intra command of one plane, in queue order
    uint32_t cmd;               // index into the intra command array
    uint32_t tb_start;          // its transform blocks: tb[tb_start .. tb_start + nb_tb)
    uint16_t nb_tb;
    uint16_t large;             // they are > 256 samples each (TrK<128>: one CTA per block) -- all blocks of a CU plane have one size
};

struct IspInfoDev { uint64_t common, part[4]; };      // == OVGPUIspInfo: availability words of an ISP coding unit

// recon_isp_subtree_{v,h} (rcn_transform_tree.c:1097-1108, 1166-1176): log2 of the partition size along the split, partition count
OVD void isp_geometry(int log2_w, int log2_h, int ver, int *lp, int *nb)
{
    int p = (ver ? log2_w : log2_h) - 2;
    const int o = ver ? log2_h : log2_w;
    if (o < 4 && p <= 4 - o) p = 4 - o;
    *lp = p; *nb = (1 << (ver ? log2_w : log2_h)) >> p;
}

// Intra block copy (rcn_ibc_l / rcn_ibc_c, rcn_ibc.c:8-140): block vector in pad[0..3]; the 4x4 units the source block touches
// (the luma range also bounds the chroma one) and the copy itself, strided over `nt` workers
OVD void ibc_vector(const IntraCmdDev &c, int *bvx, int *bvy)
{
    *bvx = (int16_t)(c.pad[0] | (c.pad[1] << 8)); *bvy = (int16_t)(c.pad[2] | (c.pad[3] << 8));
}
OVD void ibc_source_units(const IntraCmdDev &c, int *ux0, int *uy0, int *ux1, int *uy1)
{
    int bvx, bvy;
    ibc_vector(c, &bvx, &bvy);
    *ux0 = (c.x0 + bvx) >> 2; *uy0 = (c.y0 + bvy) >> 2;
    *ux1 = (c.x0 + bvx + (1 << c.log2_w) - 1) >> 2; *uy1 = (c.y0 + bvy + (1 << c.log2_h) - 1) >> 2;
}
OVD void ibc_copy(uint16_t *const plane[3], const int pitch[3], int plane_id, const IntraCmdDev &c, int tid, int nt)
{
    int bvx, bvy;
    ibc_vector(c, &bvx, &bvy);
    const int sh = plane_id ? 1 : 0, lw = c.log2_w - sh, n = 1 << (lw + c.log2_h - sh);
    const int dx = c.x0 >> sh, dy = c.y0 >> sh, sx = (c.x0 + bvx) >> sh, sy = (c.y0 + bvy) >> sh;
    for (int k = plane_id ? 1 : 0; k < (plane_id ? 3 : 1); k++)
        for (int o = tid; o < n; o += nt) {
            const int y = o >> lw, x = o & ((1 << lw) - 1);
            plane[k][(size_t)(dy + y) * pitch[k] + dx + x] = plane[k][(size_t)(sy + y) * pitch[k] + sx + x];
        }
}

struct ReconArgs {
    uint16_t *plane[3];
    int pitch[3], log2_ctu, bitdepth;
    int plane_id;                       // 0: luma commands, 1: chroma commands (Cb and Cr)
    const IntraCmdDev *cmds;
    const IspInfoDev *isp;              // table the ISP commands index through pad[0..1]
    const ReconItemDev *ritems;
    int nb_items;
    // residual
    const TuCmdDev *tu_cmds;
    const TbItemDev *tb;
    const int16_t *arena;
    int dep_quant, sh_ts_disabled, ict_type;
    const int16_t *scales;
    // dataflow state
    uint32_t *pend;                     // per 4x4 unit of this plane: == serial while an intra command covering it is not reconstructed
    uint32_t serial;
    int w_u, h_u;
    int *counter;                       // queue head
};

OVD IntraK::Args recon_intra_args(const ReconArgs &a, const uint32_t *one_cmd)
{
    IntraK::Args ia;
    for (int c = 0; c < 3; c++) { ia.plane[c] = a.plane[c]; ia.pitch[c] = a.pitch[c]; }
    ia.log2_ctu = a.log2_ctu; ia.bitdepth = a.bitdepth; ia.cmds = a.cmds; ia.order = one_cmd; ia.nb = 1;
    ia.comp0 = a.plane_id; ia.warp_comps = a.plane_id ? 2 : 1;
    ia.isp_ox = ia.isp_oy = ia.isp_lpw = ia.isp_lph = 0; ia.isp_cols = ia.isp_rows = 0;
    ia.predecoded = 0;
    return ia;
}

template <int GROUP>
OVD typename TrK<GROUP>::Args recon_tr_args(const ReconArgs &a, const TbItemDev *items, int n)
{
    typename TrK<GROUP>::Args ta;
    for (int c = 0; c < 3; c++) { ta.dst[c] = a.plane[c]; ta.dst_pitch[c] = a.pitch[c]; }
    ta.bitdepth = a.bitdepth; ta.cmds = a.tu_cmds; ta.items = items; ta.nb_items = n; ta.arena = a.arena;
    ta.dep_quant = a.dep_quant; ta.sh_ts_disabled = a.sh_ts_disabled; ta.ict_type = a.ict_type; ta.scales = a.scales;
    return ta;
}

// marks the units of every intra command of a plane as pending and resets the queue head (grid.y = plane)
struct ReconPrepK {
    static const int THREADS = 128, MIN_BLOCKS = 8, PHASES = 1;
    static const int WARPS = 4;
    struct Args {
        const IntraCmdDev *cmds;
        const ReconItemDev *ritems[2];
        int nb_items[2];
        uint32_t *pend[2];
        int *counter;                   // [2]
        uint32_t serial;
        int w_u;
    };
    struct State { int dummy; };
    template <int P>
    OVD void phase(const Args &a, unsigned char *smem, State &st, int tid, int bx, int by, int bz)
    {
        (void)smem; (void)st; (void)bz;
        const int warp = tid >> 5, lane = tid & 31, idx = bx * WARPS + warp, p = by;
        if (bx == 0 && tid == 0) a.counter[p] = 0;
        if (idx >= a.nb_items[p]) return;
        const IntraCmdDev c = a.cmds[a.ritems[p][idx].cmd];
        const int xu = c.x0 >> 2, yu = c.y0 >> 2, nw = (1 << c.log2_w) >> 2, nh = (1 << c.log2_h) >> 2;
        for (int o = lane; o < nw * nh; o += 32) a.pend[p][(size_t)(yu + o / nw) * a.w_u + xu + o % nw] = a.serial;
    }
};

// Shared memory of a worker: the residual blocks (TrK's layout) and, behind them, the prediction's reference arrays (IntraK's):
// the residual of a command is transformed BEFORE the command waits for its neighbours and is held while the prediction runs.
static const int RECON_TR_SMEM_BYTES = ((TrK<128>::SMEM_BYTES > TrK<32>::SMEM_BYTES ? TrK<128>::SMEM_BYTES : TrK<32>::SMEM_BYTES) + 15) & ~15;

#if !defined(OVGPU_HOST_EMU) && defined(OVGPU_RECON_KERNEL_IMPL)      // the one translation unit that launches it (intra.cu)
// polling load: relaxed (no cache maintenance per iteration); the acquire fence follows once the flag has flipped
__device__ __forceinline__ uint32_t recon_ld_relaxed(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void recon_st_release(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// One CTA = one worker.  Phases of IntraK / TrK are separated by CTA barriers exactly as ovd_kernel does.  The command, its
// transform-block items and their TU commands are staged in shared memory once per item, so the phase functions (which
// re-read their command in every phase) do not go back to global memory.
struct ReconStage {
    IntraCmdDev cmd;
    uint32_t zero;
    TbItemDev tb[4];
    TuCmdDev tu[4];
    IspInfoDev isp;
};

__global__ void __launch_bounds__(128, 8) ovd_recon_kernel(const ReconArgs a)
{
    extern __shared__ __align__(16) unsigned char ovd_smem[];
    __shared__ int s_item;
    __shared__ __align__(16) ReconStage sg;
    const int tid = threadIdx.x;
    IntraK::State ist; TrK<32>::State ts; TrK<128>::State tl;
    int next = 0;
    if (tid == 0) next = atomicAdd(a.counter, 1);
    for (;;) {
        if (tid == 0) s_item = next;
        __syncthreads();
        const int i = s_item;
        if (i >= a.nb_items) return;
        if (tid == 0) next = atomicAdd(a.counter, 1);      // the next command is claimed now: its round trip overlaps this one's work
        const ReconItemDev it = a.ritems[i];
        if (tid < (int)(sizeof(IntraCmdDev) / 4)) ((uint32_t *)&sg.cmd)[tid] = ((const uint32_t *)&a.cmds[it.cmd])[tid];
        if (tid == 32) sg.zero = 0;
        __syncthreads();
        const IntraCmdDev c = sg.cmd;
        const int xu = c.x0 >> 2, yu = c.y0 >> 2, nw = (1 << c.log2_w) >> 2, nh = (1 << c.log2_h) >> 2;
        const int ibc = (c.flags & 128) != 0;
        const int isp = a.plane_id == 0 && (c.flags & 32);
        unsigned char *const sm_tr = ovd_smem, *const sm_intra = ovd_smem + RECON_TR_SMEM_BYTES;
#define RECON_TR_PHASES(G, st, ta, P0, P1)                                                                                      \
        do {                                                                                                                    \
            if (P0 <= 0 && P1 >= 0) { TrK<G>::phase<0>(ta, sm_tr, st, tid, 0, 0, 0); __syncthreads(); }                         \
            if (P0 <= 1 && P1 >= 1) { TrK<G>::phase<1>(ta, sm_tr, st, tid, 0, 0, 0); __syncthreads(); }                         \
            if (P0 <= 2 && P1 >= 2) { TrK<G>::phase<2>(ta, sm_tr, st, tid, 0, 0, 0); __syncthreads(); }                         \
            if (P0 <= 3 && P1 >= 3) { TrK<G>::phase<3>(ta, sm_tr, st, tid, 0, 0, 0); __syncthreads(); }                         \
        } while (0)
#define RECON_INTRA_PHASES()                                                                                                    \
        do {                                                                                                                    \
            IntraK::phase<0>(ia, sm_intra, ist, tid, 0, 0, 0); __syncthreads();                                                 \
            IntraK::phase<1>(ia, sm_intra, ist, tid, 0, 0, 0); __syncthreads();                                                 \
            IntraK::phase<2>(ia, sm_intra, ist, tid, 0, 0, 0); __syncthreads();                                                 \
            IntraK::phase<3>(ia, sm_intra, ist, tid, 0, 0, 0); __syncthreads();                                                 \
        } while (0)
        // stages the transform-block items k0 .. k0 + nk of the command and their TU commands (+ the ISP words) in shared memory
        auto stage_tbs = [&](int k0, int nk) {
            if (tid < nk) { TbItemDev t = a.tb[it.tb_start + k0 + tid]; sg.tb[tid].cmd = tid; sg.tb[tid].kind = t.kind;
                            /* the TU command index travels through shared memory to the copy below */ ((uint32_t *)sm_tr)[tid] = t.cmd; }
            if (isp && tid >= 32 && tid < 32 + (int)(sizeof(IspInfoDev) / 4))
                ((uint32_t *)&sg.isp)[tid - 32] = ((const uint32_t *)&a.isp[c.pad[0] | (c.pad[1] << 8)])[tid - 32];
            __syncthreads();
            for (int o = tid; o < nk * (int)(sizeof(TuCmdDev) / 4); o += 128) {
                const int k = o / (int)(sizeof(TuCmdDev) / 4), wd = o - k * (int)(sizeof(TuCmdDev) / 4);
                ((uint32_t *)&sg.tu[k])[wd] = ((const uint32_t *)&a.tu_cmds[((const uint32_t *)sm_tr)[k]])[wd];
            }
            __syncthreads();
        };
        // ---- before waiting for anything: the residual of the first blocks does not depend on the picture (dequantisation and
        // both transform passes read coefficients only), so it is computed into shared memory now, off the dependency chain;
        // only its addition to the prediction (phase 3) has to follow the prediction
        const int nk0 = ovd_min(4, (int)it.nb_tb);
        if (nk0 || isp) stage_tbs(0, nk0);                   // (an ISP command without coded partitions still needs its words)
        if (nk0 && !isp) {
            if (it.large) { TrK<128>::Args ta = recon_tr_args<128>(a, sg.tb, 1); ta.cmds = sg.tu; RECON_TR_PHASES(128, tl, ta, 0, 2); }
            else          { TrK<32>::Args ta = recon_tr_args<32>(a, sg.tb, nk0); ta.cmds = sg.tu; RECON_TR_PHASES(32, ts, ta, 0, 2); }
        }
        // ... and neither do the parameters the prediction derives from the command (plain blocks: an ISP command derives them per partition)
        IntraK::Args ia = recon_intra_args(a, &sg.zero);
        ia.cmds = &sg.cmd;
        if (!ibc && !isp) { IntraK::predecode(ia, sm_intra, tid, 0, 0); ia.predecoded = 1; }
        if (ibc) {  // intra block copy: the units of the source block
            int ux0, uy0, ux1, uy1;
            ibc_source_units(c, &ux0, &uy0, &ux1, &uy1);
            const int sw = ux1 - ux0 + 1;
            for (int k = tid; k < sw * (uy1 - uy0 + 1); k += 128) {
                const uint32_t *f = a.pend + (size_t)(uy0 + k / sw) * a.w_u + ux0 + k % sw;
                while (recon_ld_relaxed(f) == a.serial) __nanosleep(32);
            }
        } else {
            // wait for the available neighbour units this block's reference samples come from: the left column from the
            // corner down to 2h below the top (avail_cols: bit 1 + row inside the CTU), the above row over 2w (avail_rows)
            const int um = ((1 << a.log2_ctu) >> 2) - 1, xl = xu & um, yl = yu & um;
            for (int k = tid; k < 2 * nh + 1 + 2 * nw; k += 128) {
                int ux, uy, bit;
                if (k < 2 * nh + 1) { ux = xu - 1; uy = yu - 1 + k; bit = yl + k; bit = bit < 64 ? (int)((c.avail_cols >> bit) & 1) : 0; }
                else { const int col = k - (2 * nh + 1); ux = xu + col; uy = yu - 1; bit = xl + col + 1; bit = bit < 64 ? (int)((c.avail_rows >> bit) & 1) : 0; }
                if (bit && ux >= 0 && uy >= 0 && ux < a.w_u && uy < a.h_u) {
                    const uint32_t *f = a.pend + (size_t)uy * a.w_u + ux;
                    while (recon_ld_relaxed(f) == a.serial) __nanosleep(32);
                }
            }
        }
        __threadfence();                                    // acquire side of the units' release stores (fence.acq_rel.gpu)
        __syncthreads();
        if (ibc) {
            ibc_copy(a.plane, a.pitch, a.plane_id, c, tid, 128);
            __syncthreads();
        } else if (!isp) {
            RECON_INTRA_PHASES();                           // ---- prediction
        }
        // ---- residual of the blocks this command covers, up to four at a time (an ISP command has at most four)
        for (int k0 = 0; k0 < (isp ? 1 : it.nb_tb); k0 += 4) {
            const int nk = ovd_min(4, it.nb_tb - k0);
            if (k0) stage_tbs(k0, nk);
            if (isp) {
                // recon_isp_subtree_{v,h} (rcn_transform_tree.c:1087-1205): the partitions one after the other, each predicted
                // (vertical split: in groups of at least four columns) and then reconstructed with its own transform block
                const int ver = (c.flags & 64) != 0;
                int lp, nb, q = 0;
                isp_geometry(c.log2_w, c.log2_h, ver, &lp, &nb);
                for (int i = 0; i < nb; i++) {
                    const int off = i << lp;
                    if (!ver || !(off & 3)) {
                        ia.isp_ox = ver ? off : 0; ia.isp_oy = ver ? 0 : off;
                        ia.isp_lpw = ver ? ovd_max(lp, 2) : c.log2_w; ia.isp_lph = ver ? c.log2_h : lp;
                        ia.isp_cols = ver ? sg.isp.part[q] : sg.isp.common; ia.isp_rows = ver ? sg.isp.common : sg.isp.part[q];
                        q++;
                        RECON_INTRA_PHASES();
                    }
                    int kb = -1;                                    // the transform block of this partition, if it has one
                    for (int k = 0; k < nk; k++)
                        if ((ver ? (sg.tu[k].x0 - c.x0) : (sg.tu[k].y0 - c.y0)) == off) kb = k;
                    if (kb < 0) continue;
                    if (it.large) { TrK<128>::Args ta = recon_tr_args<128>(a, sg.tb + kb, 1); ta.cmds = sg.tu; RECON_TR_PHASES(128, tl, ta, 0, 3); }
                    else          { TrK<32>::Args ta = recon_tr_args<32>(a, sg.tb + kb, 1); ta.cmds = sg.tu; RECON_TR_PHASES(32, ts, ta, 0, 3); }
                }
            } else if (it.large) {
                for (int k = 0; k < nk; k++) {
                    TrK<128>::Args ta = recon_tr_args<128>(a, sg.tb + k, 1);
                    ta.cmds = sg.tu;
                    if (k0 == 0 && k == 0) RECON_TR_PHASES(128, tl, ta, 3, 3);      // computed before the wait
                    else RECON_TR_PHASES(128, tl, ta, 0, 3);
                }
            } else {
                TrK<32>::Args ta = recon_tr_args<32>(a, sg.tb, nk);
                ta.cmds = sg.tu;
                if (k0 == 0) RECON_TR_PHASES(32, ts, ta, 3, 3);                      // computed before the wait
                else RECON_TR_PHASES(32, ts, ta, 0, 3);
            }
        }
#undef RECON_TR_PHASES
#undef RECON_INTRA_PHASES
        // ---- publish: every thread's stores are ordered before the barrier above; the release stores are cumulative over it
        for (int o = tid; o < nw * nh; o += 128) recon_st_release(a.pend + (size_t)(yu + o / nw) * a.w_u + xu + o % nw, 0u);
    }
}
#endif

static const int RECON_SMEM_BYTES = RECON_TR_SMEM_BYTES + IntraK::SMEM_BYTES;
static_assert(sizeof(IntraK::Blk) <= IntraK::SM_BLK, "IntraK::Blk must fit its shared-memory slot");
static_assert(sizeof(TrK<32>::Tb) <= TrK<32>::SM_TB && sizeof(TrK<128>::Tb) <= TrK<128>::SM_TB, "TrK::Tb must fit its shared-memory slot");
static_assert((((TrK<32>::SM_CT + TrK<32>::SM_T2 + TrK<32>::SM_RES) * 2) & 7) == 0 && (((TrK<128>::SM_CT + TrK<128>::SM_T2 + TrK<128>::SM_RES) * 2) & 7) == 0, "Tb slot alignment");
