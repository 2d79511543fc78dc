"""Seeded synthetic intra coding units (SURVEY.md §8(d) "Intra", seed 1006): luma blocks in decode order (CTU by CTU,
z-order inside a CTU) with modes U[0,66].  The availability words of OVGPUIntraCmd are filled by the recorder
(the reference harness in tests; `availability()` below reproduces them for the bench)."""
import numpy as np

from . import synth

INTRA_CMD = np.dtype([("x0", np.int16), ("y0", np.int16), ("log2_w", np.uint8), ("log2_h", np.uint8), ("mode", np.uint8),
                      ("mode_c", np.uint8), ("flags", np.uint8), ("mrl_idx", np.uint8), ("ciip_wt", np.uint8), ("pad", np.uint8, 5),
                      ("avail_cols", np.uint64), ("avail_rows", np.uint64)])
assert INTRA_CMD.itemsize == 32

INTRA_LUMA, INTRA_CHROMA, INTRA_CIIP, INTRA_CCLM_COLLOCATED, INTRA_MIP, INTRA_ISP, INTRA_ISP_VER = 1, 2, 4, 8, 16, 32, 64
INTRA_IBC = 128


def intra_commands(w, h, seed=1006, log2_ctu=7, min_log2=2, max_log2=6, chroma=False, p_ciip=0.0, p_mrl=0.0, p_cclm=0.0,
                   collocated=False, p_inter=0.0, blocks=None, p_mip=0.0, p_isp=0.0, p_ibc=0.0):
    """p_inter: fraction of coding units that are NOT intra (flags = 0 placeholders): they keep the picture's samples -- the
    already reconstructed inter CUs of a real picture -- and only take part in the availability replay.  Without them an
    all-intra picture without residual degenerates to the mid value everywhere.  The placeholders are for the oracle
    harness (which has to mark them in the reference's progress field); drop them before handing commands to the GPU."""
    rng = np.random.default_rng(seed)
    part = blocks if blocks is not None else synth.partition(w, h, rng, log2_ctu, min_log2=min_log2, max_log2=min(max_log2, log2_ctu), p_split=0.5)
    cmds = np.zeros(len(part), dtype=INTRA_CMD)
    done = np.zeros((h >> 2, w >> 2), bool) if p_ibc else None      # 4x4 units of the coding units decoded so far
    ctu = 1 << log2_ctu
    for i, (x, y, lw, lh) in enumerate(part):
        c = cmds[i]
        c["x0"], c["y0"], c["log2_w"], c["log2_h"] = x, y, lw, lh
        if p_ibc:
            if i:
                px, py, plw, plh = part[i - 1]
                done[py >> 2:(py + (1 << plh)) >> 2, px >> 2:(px + (1 << plw)) >> 2] = True
            if rng.random() < p_ibc:
                # intra block copy (screen content): a block vector to samples decoded before, in this CTU or the one to its left,
                # rows inside the CTU row -- what the reference's CTU ring buffer holds (rcn_ibc.c:14-27); any sample alignment
                bw, bh, cx, cy = 1 << lw, 1 << lh, x >> log2_ctu, y >> log2_ctu
                for _ in range(16):
                    sx = int(rng.integers(max(0, (cx - 1) * ctu), min(w - bw, x + bw) + 1))
                    sy = int(rng.integers(cy * ctu, min(h, (cy + 1) * ctu) - bh + 1))
                    if rng.random() < 0.3:
                        sx &= ~3; sy &= ~3
                    if done[sy >> 2:((sy + bh - 1) >> 2) + 1, sx >> 2:((sx + bw - 1) >> 2) + 1].all():
                        c["flags"] = INTRA_IBC | INTRA_LUMA | (INTRA_CHROMA if chroma else 0)
                        c["pad"][:4] = np.array([sx - x, sy - y], np.int16).view(np.uint8)
                        break
                if c["flags"]:
                    continue
        r = rng.random()
        if p_inter and rng.random() < p_inter:
            continue
        c["flags"] = INTRA_LUMA
        c["mode"] = 0 if r < 0.15 else (1 if r < 0.3 else (int(rng.choice([18, 50, 2, 34, 66])) if r < 0.4 else int(rng.integers(2, 67))))
        if chroma:
            # single tree: Cb / Cr blocks of half the size, mode = DM (the luma mode) or an explicit planar / DC / ver / hor / random
            r = rng.random()
            mc = int(c["mode"]) if r < 0.4 else (0 if r < 0.5 else (1 if r < 0.6 else (int(rng.choice([18, 50, 2, 34, 66])) if r < 0.7 else int(rng.integers(2, 67)))))
            c["flags"] |= INTRA_CHROMA
            c["mode_c"] = mc
            if p_cclm and rng.random() < p_cclm:                   # CCLM / MDLM left / MDLM top (30 % of chroma)
                c["mode_c"] = int(rng.integers(67, 70))
                if collocated:
                    c["flags"] |= INTRA_CCLM_COLLOCATED
        if p_ciip and lw + lh >= 6 and rng.random() < p_ciip:       # CIIP: CUs of >= 64 samples (4x16 has the narrow-chroma rule)
            c["flags"] |= INTRA_CIIP
            c["ciip_wt"] = int(rng.integers(1, 4))
        if p_mrl and (y & ((1 << log2_ctu) - 1)) and not (c["flags"] & INTRA_CIIP) and rng.random() < p_mrl:
            c["mrl_idx"] = int(rng.integers(1, 3))                 # not on the first row of a CTU (vcl_coding_unit.c:1377)
        if p_mip and not (c["flags"] & INTRA_CIIP) and not c["mrl_idx"] and rng.random() < p_mip:
            # MIP (10 % of intra CUs): `mode` carries cu_opaque = transposed << 7 | matrix index
            nm = 16 if (lw == 2 and lh == 2) else (8 if (lw == 2 or lh == 2 or (lw <= 3 and lh <= 3)) else 6)
            c["flags"] |= INTRA_MIP
            c["mode"] = (int(rng.integers(0, 2)) << 7) | int(rng.integers(0, nm))
        if p_isp and not (c["flags"] & (INTRA_CIIP | INTRA_MIP)) and not c["mrl_idx"] and lw + lh >= 5 and rng.random() < p_isp:
            # ISP (10 % of intra CUs): whole luma coding block, split direction in the flags; the per-partition
            # transform blocks and availability words live in the ISP table (isp_info), indexed through pad[0..1]
            c["flags"] |= INTRA_ISP | (INTRA_ISP_VER if rng.random() < 0.5 else 0)
    return cmds


def availability(w, h, cmds, log2_ctu=7, isp_words=None):
    """avail_cols / avail_rows of every command = progress_field.vfield[x0 >> 2] / .hfield[y0 >> 2] at call time
    (init_ctu_bitfield(_border), drv_utils.h:110-260; ctu_field_set_rect_bitfield :46-72), replayed from the decode order.
    isp_words (ISP_WORDS array, filled in place): the words vvc_intra_pred_isp reads for the partitions of ISP commands, taken
    after the whole coding block has been marked (vcl_transform_unit.c:1878)."""
    ctu = 1 << log2_ctu
    n_u = ctu >> 2
    ncw, nch = -(-w // ctu), -(-h // ctu)
    out = cmds.copy()
    cur = None
    for c in out:
        cx, cy = int(c["x0"]) >> log2_ctu, int(c["y0"]) >> log2_ctu
        if cur != (cx, cy):
            cur = (cx, cy)
            cw_u, ch_u = min(ctu, w - cx * ctu) >> 2, min(ctu, h - cy * ctu) >> 2
            lft, abv, tr = int(cx > 0), int(cy > 0), int(cy > 0 and cx + 1 < ncw)
            hf = [0] * (n_u + 2)
            vf = [0] * (n_u + 2)
            full = cw_u == n_u and ch_u == n_u
            if full:
                for i in range(1, n_u + 1):
                    hf[i] = lft
                    vf[i] = abv
                inner = ((1 << n_u) - 1) << 1
                hf[0] = inner if abv else 0
                vf[0] = inner if lft else 0
            else:
                for i in range(1, ch_u + 1):
                    hf[i] = lft
                for i in range(1, cw_u + 1):
                    vf[i] = abv
                hf[0] = (((1 << cw_u) - 1) << 1) if abv else 0
                vf[0] = (((1 << ch_u) - 1) << 1) if lft else 0
            hf[0] |= abv & lft
            vf[0] |= abv & lft
            if tr:
                hf[0] |= ((1 << n_u) - 1) << (n_u + 1)
            if cx + 2 == ncw and (w & (ctu - 1)):
                hf[0] &= (1 << ((((w & (ctu - 1)) + ctu) >> 2) + 1)) - 1
        x0, y0 = int(c["x0"]) & (ctu - 1), int(c["y0"]) & (ctu - 1)
        c["avail_cols"] = np.uint64(vf[x0 >> 2] & 0xFFFFFFFFFFFFFFFF)
        c["avail_rows"] = np.uint64(hf[y0 >> 2] & 0xFFFFFFFFFFFFFFFF)
        xu, yu, nw, nh = x0 >> 2, y0 >> 2, (1 << int(c["log2_w"])) >> 2, (1 << int(c["log2_h"])) >> 2
        for i in range(1, nh + 1):
            hf[yu + i] |= ((1 << nw) - 1) << (1 + xu)
        for i in range(1, nw + 1):
            vf[xu + i] |= ((1 << nh) - 1) << (1 + yu)
        if isp_words is not None and (int(c["flags"]) & INTRA_ISP):
            wd = isp_words[int(c["pad"][0]) | (int(c["pad"][1]) << 8)]
            ver = bool(int(c["flags"]) & INTRA_ISP_VER)
            lp, nb = isp_partitions(int(c["log2_w"]), int(c["log2_h"]), ver)
            wd["common"] = np.uint64((hf[y0 >> 2] if ver else vf[x0 >> 2]) & 0xFFFFFFFFFFFFFFFF)
            q = 0
            for i in range(nb):
                off = i << lp
                if ver and (off & 3):
                    continue
                v = vf[(x0 + off) >> 2] if ver else hf[((y0 + off) >> 2) + (1 if (y0 + off) % 4 else 0)]
                wd["part"][q] = np.uint64(v & 0xFFFFFFFFFFFFFFFF)
                q += 1
    return out


# == OVGPUIspInfo / the head of OVOIspInfo: availability words of an ISP coding unit
ISP_WORDS = np.dtype([("common", np.uint64), ("part", np.uint64, 4)])


def isp_partitions(lw, lh, ver):
    """(log2 partition size along the split, number of partitions), recon_isp_subtree_{v,h} (rcn_transform_tree.c:1097-1108, 1166-1176)"""
    lp, lo = (lw if ver else lh) - 2, (lh if ver else lw)
    if lo < 4 and lp <= 4 - lo:
        lp = 4 - lo
    return lp, (1 << (lw if ver else lh)) >> lp


def isp_tables(cmds, seed, dep_quant=1, mts_enabled=1, p_cbf=0.7):
    """Transform blocks of the ISP commands: returns (cmds with the table index in pad[0..1], words[ISP_WORDS] zeroed -- the
    harness / availability replay fills them, tus[n_isp][4] of TU_CMD with the OVGPU_TU_ISP flag, arena)."""
    from . import synth_tu
    rng = np.random.default_rng(seed)
    cmds = cmds.copy()
    idx = np.nonzero(cmds["flags"] & INTRA_ISP)[0]
    words = np.zeros(len(idx), dtype=ISP_WORDS)
    tus = np.zeros((len(idx), 4), dtype=synth_tu.TU_CMD)
    arena, off = [], 0
    for k, i in enumerate(idx):
        c = cmds[i]
        c["pad"][0], c["pad"][1] = k & 255, k >> 8
        lw, lh, ver = int(c["log2_w"]), int(c["log2_h"]), bool(c["flags"] & INTRA_ISP_VER)
        lp, nb = isp_partitions(lw, lh, ver)
        tlw, tlh = (lp, lh) if ver else (lw, lp)
        qp = int(rng.integers(22, 38)) + 12
        lfnst = tlw >= 2 and tlh >= 2 and rng.random() < 0.3
        lfnst_idx = int(rng.integers(0, 2))
        # drv_lfnst_mode_l(coding block, intra mode): the wide-angle mapping of drv_lfnst.c:50-76
        mode = int(c["mode"])
        if mode > 1:
            ms = (0, 6, 10, 12, 14, 15)[abs(lw - lh)]
            if lw > lh and mode < 2 + ms:
                mode += 65
            elif lh > lw and mode > 66 - ms:
                mode -= 67
        lmode = mode + 14 + 67 if mode < 0 else (mode + 14 if mode >= 67 else mode)
        for p in range(nb):
            if rng.random() >= p_cbf:
                continue
            t = tus[k, p]
            t["x0"], t["y0"] = int(c["x0"]) + (p << lp if ver else 0), int(c["y0"]) + (0 if ver else p << lp)
            t["log2_w"], t["log2_h"], t["cbf_mask"] = tlw, tlh, 0x10
            t["flags"] = synth_tu.TU_ISP | (synth_tu.TU_LFNST_L if lfnst else 0)
            t["qp"] = [qp, qp, qp, qp]; t["qp_ts"] = [max(qp, 16)] * 4
            t["lfnst_idx"], t["lfnst_mode_l"] = lfnst_idx, lmode
            dst_h = mts_enabled and 1 < tlw <= 4 if ver else mts_enabled and lw <= 4
            dst_v = mts_enabled and lh <= 4 if ver else mts_enabled and 1 < tlh <= 4
            t["cu_mts_idx"] = (1 if dst_h else 0) | (2 if dst_v else 0)
            cmax = synth_tu._cmax(qp, tlw, tlh, dep_quant, True)
            if tlw < 2 or tlh < 2:
                # 1xN / 2xN / Nx1 / Nx2: raster coefficients (dequant_non_4x4_sb), nothing beyond the 32nd row / column
                # (64-point zero-out); last significant position as the parser reports it
                w, n = 1 << tlw, 1 << (tlw + tlh)
                co = np.rint(rng.laplace(0, 10, size=n)).astype(np.int64)
                co[rng.random(n) < 0.5] = 0
                c2 = co.reshape(1 << tlh, w)
                c2[32:, :] = 0; c2[:, 32:] = 0
                co = np.clip(co, -min(2000, cmax), min(2000, cmax)).astype(np.int16)
                nz = np.nonzero(co)[0]
                if not len(nz):
                    co[0] = 3; nz = np.array([0])
                last = ((int(nz[-1]) // w) << 8) | (int(nz[-1]) % w)
                # significance map by coefficient group (8x2 / 2x8 for the two-sample-thin shapes, rcn_transform_tree.c:927-928,
                # 1008-1009: the transforms skip what it does not flag); one-sample-thin blocks use last_pos instead
                sig = 1
                gw, gh = (8, 2) if tlh == 1 else ((2, 8) if tlw == 1 else (w, 1 << tlh))
                c3 = co.reshape(1 << tlh, w)
                for sy in range(max(1, (1 << tlh) // gh)):
                    for sx in range(max(1, w // gw)):
                        if sy < 8 and sx < 8 and c3[sy * gh:(sy + 1) * gh, sx * gw:(sx + 1) * gw].any():
                            sig |= 1 << (sy * 8 + sx)
                sig = np.uint64(sig)
            else:
                co, last, sig = synth_tu._tb_coeffs(rng, tlw, tlh, "lfnst" if lfnst else "tr", False, cmax)
            t["coeff_off"][2], t["last_pos"][2], t["sig_sb_map"][2] = off, last, sig
            arena.append(co); off += len(co)
    arena = np.concatenate(arena + [np.zeros(64, np.int16)]) if arena else np.zeros(64, np.int16)
    return cmds, words, tus, np.ascontiguousarray(arena, dtype=np.int16)
