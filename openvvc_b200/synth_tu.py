"""Seeded synthetic transform-unit lists (SURVEY.md §8(d) "Residual", seed 1005): what vcl_transform_unit.c
would hand to tmp.rcn_tu_st for the TUs of non-intra CUs -- cbf / transform-skip / MTS flags, QPs, last
position, significant sub-block maps and the coefficient buffers in the reference's 4x4-sub-block order."""
import numpy as np

from . import synth

CLIP_P = 0.02   # share of sub-blocks with out-of-range coefficients (exercise the dequant / transform clips)
TU_IMPLICIT_MTS, TU_LFNST_L, TU_LFNST_C, TU_SCALE_ANCHOR, TU_ISP = 1, 2, 4, 8, 16
# == OVGPUTuCmd (include/ovvc_b200.h)
TU_CMD = np.dtype({"names": ["x0", "y0", "log2_w", "log2_h", "cbf_mask", "tr_skip_mask", "cu_mts_flag", "cu_mts_idx",
                             "flags", "lfnst_idx", "qp", "qp_ts", "last_pos", "chroma_scale", "lfnst_mode_l", "lfnst_mode_c",
                             "scale_anchor", "coeff_off", "bdpcm", "sig_sb_map"],
                   "formats": [np.int16, np.int16, np.uint8, np.uint8, np.uint8, np.uint8, np.uint8, np.uint8,
                               np.uint8, np.uint8, (np.uint8, 4), (np.uint8, 4), (np.uint16, 3), np.int16, np.uint8, np.uint8,
                               np.uint16, (np.uint32, 3), np.uint32, (np.uint64, 3)],
                   "offsets": [0, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 16, 20, 26, 28, 29, 30, 32, 44, 48], "itemsize": 72})


def _tb_coeffs(rng, lw, lh, mode, mts, cmax=32000):
    """Returns (coeff array in the reference's buffer order, last_pos, sig_sb_map).
    mode: 'tr' (transform), 'dc', 'ts_raw' (transform skip, plain residual), 'ts_deq' (sh_ts_disabled)."""
    w, h = 1 << lw, 1 << lh
    cw, ch = min(32, w), min(32, h)
    if lw < 2 or lh < 2 or mode == "ts_raw":
        # raster layout (dequant_non_4x4_sb / memcpy paths)
        n = w * h if mode == "ts_raw" else cw * ch
        c = np.rint(rng.laplace(0, 8, size=n)).astype(np.int64)
        c[rng.random(n) < 0.5] = 0
        sig = 1
        if mode == "dc":
            c[1:] = 0
            c[0] = int(rng.integers(-300, 300)) or 7
        elif lw < 2 or lh < 2:
            # thin chroma blocks are coded in 2x2 (or 2x8 / 8x2) coefficient groups (rcn_transform_tree.c:592-603):
            # keep coefficients only inside groups flagged in sig_sb_map
            lsw = 3 if lw > 2 else 1
            lsh = 3 if lh > 2 else 1
            c2 = c.reshape(h, w)
            sig = 0
            for sy in range(max(1, h >> lsh)):
                for sx in range(max(1, w >> lsw)):
                    if (sx == 0 and sy == 0) or rng.random() < 0.5:
                        sig |= 1 << (sy * 8 + sx)
                    else:
                        c2[sy << lsh:(sy + 1) << lsh, sx << lsw:(sx + 1) << lsw] = 0
            c = c2.reshape(-1)
        c = np.clip(c, -min(2000, cmax), min(2000, cmax)).astype(np.int16)
        return c, (0 if mode == "dc" else 1), np.uint64(sig)
    sbw, sbh = cw // 4, ch // 4
    if mode == "lfnst" and cmax < 32000:
        cmax = max(1, cmax // 16)                   # keep the LFNST output inside int16 too (legal levels)
    lim_w, lim_h = (min(sbw, 4), min(sbh, 4)) if mts else (sbw, sbh)        # MTS zero-out: 16x16 coefficients
    sig = np.zeros((8, 8), dtype=bool)
    sig[0, 0] = True
    if mode not in ("dc", "lfnst"):                 # LFNST: only the first sub-block carries coefficients
        for sy in range(lim_h):
            for sx in range(lim_w):
                if rng.random() < 0.9 / (1 + sx + sy):
                    sig[sy, sx] = True
    buf = np.zeros((sbh, sbw, 16), dtype=np.int16)
    last = 0
    for sy in range(sbh):
        for sx in range(sbw):
            if sig[sy, sx]:
                c = np.rint(rng.laplace(0, 12.0 / (1 + 0.5 * (sx + sy)), size=16)).astype(np.int64)
                if rng.random() < CLIP_P:
                    c = c * 400                       # exercise the dequantisation / transform clips
                c = np.clip(c, -cmax, cmax).astype(np.int16)
                if mode == "dc":
                    c[1:] = 0
                    c[0] = int(rng.integers(-300, 300)) or 5
                buf[sy, sx] = c
                nz = np.nonzero(c)[0]
                if len(nz):
                    k = int(nz[-1])
                    last = ((4 * sy + (k >> 2)) << 8) | (4 * sx + (k & 3))
    if mode == "dc":
        last = 0
    elif last == 0:
        buf[0, 0, 1] = 3
        last = 1
    m = 0
    for sy in range(8):
        for sx in range(8):
            if sig[sy, sx]:
                m |= 1 << (sy * 8 + sx)
    return buf.reshape(-1), last, np.uint64(m)


def _cmax(qp, lw, lh, dep_quant, legal):
    """Largest |coefficient| whose dequantised value stays inside int16 (rcn_dequant.c:92-157). The reference's
    SIMD dequantisation saturates to -32768 where its C code clips to -32767, so SIMD-vs-C comparisons use
    `legal` coefficient levels -- which is also all a conforming bitstream can carry."""
    if not legal:
        return 32000
    s = lw + lh
    q = qp + 1 if dep_quant else qp
    shift = 6 + (1 if dep_quant else 0) - 5 - q // 6 + (s >> 1) + (s & 1)
    scale = ((40, 45, 51, 57, 64, 72), (57, 64, 72, 80, 90, 102))[s & 1][q % 6]
    gain = max(scale * 2.0 ** (-shift), 72 * 2.0 ** (max(qp, 16) // 6 - 6))      # transform and transform-skip scaling
    return max(1, int(32000 / gain))


def tu_commands(w, h, seed=1005, log2_ctu=7, ict_type=0, sh_ts_disabled=0, min_log2=2, p_cbf=0.6, dep_quant=1, legal=False, p_intra=0.0,
                blocks=None, intra_mask=None, p_bdpcm=0.0):
    """Returns (cmds[TU_CMD], arena int16). One TU per partition block (<= 64x64 luma).
    blocks / intra_mask: use this partition (and these intra CUs) instead of drawing them (synth_recon)."""
    rng = np.random.default_rng(seed)
    if blocks is None:
        blocks = synth.partition(w, h, rng, log2_ctu, min_log2=min_log2, max_log2=min(6, log2_ctu))
    cmds = np.zeros(len(blocks), dtype=TU_CMD)
    arena = []
    off = 0

    def push(c):
        nonlocal off
        arena.append(c)
        o = off
        off += len(c)
        return o

    for i, (x, y, lw, lh) in enumerate(blocks):
        c = cmds[i]
        c["x0"], c["y0"], c["log2_w"], c["log2_h"] = x, y, lw, lh
        qp = int(rng.integers(22, 38)) + 12
        c["qp"] = [qp, max(12, qp - int(rng.integers(0, 4))), max(12, qp - int(rng.integers(0, 4))), max(12, qp - int(rng.integers(0, 4)))]
        c["qp_ts"] = np.maximum(c["qp"], 16)
        c["chroma_scale"] = int(rng.integers(1500, 2700)) if (ict_type & 1) else 2048
        cbf, ts = 0, 0
        # intra CUs (LFNST idx on intra): implicit MTS / LFNST on luma (rcn_tu_st), or a dual-tree
        # chroma TU with LFNST (rcn_tu_c)
        intra = rng.random() < p_intra
        if intra_mask is not None:
            intra = bool(intra_mask[i])
        dual_c = intra and lw >= 3 and lh >= 3 and rng.random() < 0.3
        lfnst_l = lfnst_c = False
        if intra:
            c["lfnst_mode_l"] = int(rng.integers(0, 67))
            c["lfnst_mode_c"] = int(rng.integers(0, 67))
            c["lfnst_idx"] = int(rng.integers(0, 2))
            if not dual_c and rng.random() < 0.7:
                c["flags"] |= TU_IMPLICIT_MTS
        if (intra and not dual_c and rng.random() < 0.85) or (not intra and rng.random() < p_cbf):   # luma
            cbf |= 0x10
            can_ts = lw <= 5 and lh <= 5
            is_ts = can_ts and rng.random() < 0.08
            mts = (not is_ts) and lw <= 5 and lh <= 5 and rng.random() < 0.15
            lfnst_l = intra and not is_ts and not mts and rng.random() < 0.45
            if lfnst_l:
                c["flags"] |= TU_LFNST_L
            if is_ts:
                ts |= 0x10
                mode = "ts_deq" if sh_ts_disabled else "ts_raw"
            else:
                mode = "lfnst" if lfnst_l else ("dc" if (not mts and rng.random() < 0.15) else "tr")
            if mts:
                c["cu_mts_flag"] = 1
                c["cu_mts_idx"] = int(rng.integers(0, 4))
            co, last, sig = _tb_coeffs(rng, lw, lh, mode, mts, _cmax(qp, lw, lh, dep_quant, legal))
            c["coeff_off"][2] = push(co); c["last_pos"][2] = last; c["sig_sb_map"][2] = sig
        lwc, lhc = lw - 1, lh - 1
        r = rng.random()
        if dual_c:
            r *= 0.55
            lfnst_c = rng.random() < 0.6
            if lfnst_c:
                c["flags"] |= TU_LFNST_C
        if r < 0.15:                                               # joint Cb-Cr
            cbf |= 0x8 | int(rng.integers(1, 4))
            is_ts = lwc <= 5 and lhc <= 5 and rng.random() < 0.08
            if is_ts:
                ts |= 0x1
            mode = ("ts_deq" if sh_ts_disabled else "ts_raw") if is_ts else ("lfnst" if lfnst_c else ("dc" if rng.random() < 0.15 else "tr"))
            co, last, sig = _tb_coeffs(rng, lwc, lhc, mode, False, _cmax(qp, lwc, lhc, dep_quant, legal))
            c["coeff_off"][0] = push(co); c["last_pos"][0] = last; c["sig_sb_map"][0] = sig
        elif r < 0.55:
            for k, bit in ((0, 0x2), (1, 0x1)):
                if rng.random() < 0.6:
                    cbf |= bit
                    is_ts = lwc <= 5 and lhc <= 5 and rng.random() < 0.08
                    if is_ts:
                        ts |= bit
                    mode = ("ts_deq" if sh_ts_disabled else "ts_raw") if is_ts else ("lfnst" if lfnst_c else ("dc" if rng.random() < 0.15 else "tr"))
                    co, last, sig = _tb_coeffs(rng, lwc, lhc, mode, False, _cmax(qp, lwc, lhc, dep_quant, legal))
                    c["coeff_off"][k] = push(co); c["last_pos"][k] = last; c["sig_sb_map"][k] = sig
        if p_bdpcm and intra and ts and rng.random() < p_bdpcm:
            # BDPCM CU (screen content): its transform-skip blocks accumulate along rows or columns; luma and chroma flagged apart
            b = 0
            if ts & 0x10:
                b |= 1 | (int(rng.integers(0, 2)) << 1)
            if (ts & 0x3) and rng.random() < 0.7:
                b |= 4 | (int(rng.integers(0, 2)) << 3)
            c["bdpcm"] = b
        c["cbf_mask"], c["tr_skip_mask"] = cbf, ts
    arena = np.concatenate(arena + [np.zeros(64, np.int16)]) if arena else np.zeros(64, np.int16)
    return cmds, np.ascontiguousarray(arena, dtype=np.int16)
