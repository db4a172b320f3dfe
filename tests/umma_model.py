"""Numpy model of the convolution kernels' data movement and MMA issue logic (CPU, no GPU needed).

It mirrors csrc/conv_kernels.cu step for step -- the activation layout in HBM, the bulk copies into the
shared-memory stage, the UMMA shared-memory descriptors (no-swizzle K-major canonical layout), the
input-stationary N = 3*NB issue sequence over the 8-slot TMEM ring (zeroed slots, always-accumulate MMAs,
ring-wrap splits), and the epilogue -- but executes sequentially.  Because the operand images are the REAL bytes
produced by the C ABI's host packer (mlsim_cnn_pack_host) and the result is compared with the CPU oracle,
this pins the index arithmetic of the kernels before any GPU time is spent.  It does not model
synchronisation.
"""
from __future__ import annotations

import numpy as np

CM_NPX = 130
CM_PLANE = CM_NPX * 16
CM_STAGE_BYTES = 8 * CM_PLANE
CM_SLOTS = 8
CM_RING = 6


def bf16_round(x: np.ndarray) -> np.ndarray:
    """float32 -> bf16 (RNE) -> float32."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    lsb = (b >> 16) & 1
    b = (b + 0x7FFF + lsb) & 0xFFFF0000
    return b.astype(np.uint32).view(np.float32)


def bf16_bits(x: np.ndarray) -> np.ndarray:
    return (bf16_round(x).view(np.uint32) >> 16).astype(np.uint16)


def bits_to_f32(h: np.ndarray) -> np.ndarray:
    return (h.astype(np.uint32) << 16).view(np.float32)


def make_desc(start: int, lbo: int, sbo: int) -> int:
    assert start % 16 == 0 and lbo % 16 == 0 and sbo % 16 == 0
    return (start >> 4) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46)


def gather_operand(smem: np.ndarray, desc: int, rows: int) -> np.ndarray:
    """Decode a no-swizzle K-major descriptor: element (r, k), k < 16, lives at
    start + (r//8)*SBO + (r%8)*16 + (k//8)*LBO + (k%8)*2."""
    start = (desc & 0x3FFF) << 4
    lbo = ((desc >> 16) & 0x3FFF) << 4
    sbo = ((desc >> 32) & 0x3FFF) << 4
    r = np.arange(rows)[:, None]
    k = np.arange(16)[None, :]
    addr = start + (r // 8) * sbo + (r % 8) * 16 + (k // 8) * lbo + (k % 8) * 2
    h = smem[addr].astype(np.uint16) | (smem[addr + 1].astype(np.uint16) << 8)
    return bits_to_f32(h)


def umma(tmem: np.ndarray, dcol: int, desc_a: int, desc_b: int, n: int, smem: np.ndarray, accumulate: bool):
    assert n % 16 == 0 and 16 <= n <= 256 and dcol + n <= tmem.shape[1]
    a = gather_operand(smem, desc_a, 128).astype(np.float64)
    b = gather_operand(smem, desc_b, n).astype(np.float64)
    d = a @ b.T
    if accumulate:
        tmem[:, dcol:dcol + n] += d
    else:
        tmem[:, dcol:dcol + n] = d


def act_offset(b, x, c8, yp, nx, ny):
    """byte offset of the 16-byte granule (b, x, c8, yp) in act[b][x][8][ny+2][8] (bf16)."""
    return ((((b * nx + x) * 8 + c8) * (ny + 2)) + yp) * 16


def pack_activation(t: np.ndarray) -> np.ndarray:
    """t: [B, C<=64, nx, ny] float -> bytes of the device activation layout (zero halo columns)."""
    B, Cc, nx, ny = t.shape
    buf = np.zeros(B * nx * 8 * (ny + 2) * 8, dtype=np.uint16)
    v = buf.reshape(B, nx, 8, ny + 2, 8)
    full = np.zeros((B, 64, nx, ny), dtype=np.float32)
    full[:, :Cc] = t
    # [B, c8, e, nx, ny] -> [B, nx, c8, ny, e]
    v[:, :, :, 1:ny + 1, :] = bf16_bits(full).reshape(B, 8, 8, nx, ny).transpose(0, 3, 1, 4, 2)
    return buf.view(np.uint8)


def unpack_activation(raw: np.ndarray, B, nx, ny) -> np.ndarray:
    v = raw.view(np.uint16).reshape(B, nx, 8, ny + 2, 8)[:, :, :, 1:ny + 1, :]
    return bits_to_f32(np.ascontiguousarray(v.transpose(0, 2, 4, 1, 3))).reshape(B, 64, nx, ny)


def conv_first_model(u, v, wimg: np.ndarray, scale: float):
    """k_conv_first: returns activation bytes."""
    B, nx, ny = u.shape
    out = np.zeros(B * nx * 8 * (ny + 2) * 16, dtype=np.uint8)
    LBO_A, LBO_B = 128 * 16, 64 * 16
    WB = 4 * LBO_B
    bias = wimg[WB:WB + 256].view(np.float32)
    smem = np.zeros(4 * LBO_A + WB, dtype=np.uint8)
    sA, sW = 0, 4 * LBO_A
    smem[sW:sW + WB] = wimg[:WB]
    nyt = (ny + 127) // 128
    tmem = np.zeros((128, 64), dtype=np.float64)
    for b in range(B):
        for x in range(nx):
            for yt in range(nyt):
                a16 = np.zeros((128, 32), dtype=np.uint16)
                for m in range(128):
                    y = yt * 128 + m
                    val = np.zeros(32, dtype=np.float32)
                    for i in range(3):
                        for j in range(3):
                            xx, yy = x + i - 1, y + j - 1
                            if 0 <= xx < nx and 0 <= yy < ny:
                                val[(i * 3 + j) * 2 + 0] = np.float32(u[b, xx, yy]) * np.float32(scale)
                                val[(i * 3 + j) * 2 + 1] = np.float32(v[b, xx, yy]) * np.float32(scale)
                    val[18] = val[19] = 1.0                      # bias columns (bias_hi, bias_lo in the packed weights)
                    a16[m] = bf16_bits(val)
                for c in range(4):
                    blk = np.ascontiguousarray(a16[:, c * 8:(c + 1) * 8]).view(np.uint8).reshape(-1)
                    smem[sA + c * LBO_A: sA + c * LBO_A + 128 * 16] = blk
                for k16 in range(2):
                    da = make_desc(sA + k16 * 2 * LBO_A, LBO_A, 128)
                    db = make_desc(sW + k16 * 2 * LBO_B, LBO_B, 128)
                    umma(tmem, 0, da, db, 64, smem, k16 > 0)
                for m in range(128):
                    y = yt * 128 + m
                    if y >= ny:
                        continue
                    f = np.maximum(tmem[m].astype(np.float32), 0).astype(np.float32)     # bias came through the GEMM
                    h = bf16_bits(f)
                    for c in range(8):
                        o = act_offset(b, x, c, y + 1, nx, ny)
                        out[o:o + 16] = h[c * 8:(c + 1) * 8].view(np.uint8)
    return out


def conv_mid_model(act_in: np.ndarray, wimg: np.ndarray, B, nx, ny, rs, NB, relu, grid=3, stages=6):
    """k_conv_mid<NB>: returns (activation bytes, or [B, NB, nx, ny] float32 pre-activation for NB == 16)."""
    LBO_B = 3 * NB * 16
    WDY = 8 * LBO_B
    WB = 3 * WDY
    bias = wimg[WB:WB + NB * 4].view(np.float32)
    stage_off = ((WB + NB * 4 + 127) // 128) * 128
    smem = np.zeros(stage_off + stages * CM_STAGE_BYTES, dtype=np.uint8)
    smem[:WB] = wimg[:WB]
    out_act = np.zeros(B * nx * 8 * (ny + 2) * 16, dtype=np.uint8) if NB == 64 else None
    out_y = np.zeros((B, NB, nx, ny), dtype=np.float32) if NB != 64 else None
    nyt = (ny + 127) // 128
    nstrips = (nx + rs - 1) // rs
    nitems = B * nstrips * nyt
    plane = ny + 2
    rng = np.random.default_rng(0)
    for cta in range(grid):
        tmem = np.zeros((128, NB * CM_SLOTS))             # the epilogue warps zero every slot at kernel start
        smem[stage_off:] = rng.integers(0, 255, smem.size - stage_off, dtype=np.uint8)   # stale stage bytes
        m = 0
        nb = 0
        for item in range(cta, nitems, grid):
            yt = item % nyt
            strip = (item // nyt) % nstrips
            b = item // (nyt * nstrips)
            x0 = strip * rs
            x1 = min(x0 + rs, nx)
            xlo, xhi = max(x0 - 1, 0), min(x1, nx - 1)
            y0 = yt * 128
            npx = min(CM_NPX, ny + 2 - y0)
            done_rows = []
            for X in range(xlo, xhi + 1):
                stage = m % stages
                sbase = stage_off + stage * CM_STAGE_BYTES
                # producer: 8 bulk copies
                for c in range(8):
                    src = (((b * nx + X) * 8 + c) * plane + y0) * 16
                    smem[sbase + c * CM_PLANE: sbase + c * CM_PLANE + npx * 16] = act_in[src: src + npx * 16]
                lo, hi = max(X - 1, x0), min(X + 1, x1 - 1)

                def is_new(x):
                    return (x == X + 1) or (x == X and X == 0)

                # accumulator window, exactly as the kernel's MMA issuer computes it: ring period 6, slots 6 and 7
                # mirror ring positions 0 and 1, so the window (X-1, X, X+1) is always slots q0..q0+2 (one segment)
                q0 = (nb + CM_RING + X - 1 - x0) % CM_RING
                pos_lo = lo - (X - 1)
                nrows = hi - lo + 1
                assert q0 + pos_lo + nrows <= CM_SLOTS
                for t in range(12):
                    dyi, k16 = t >> 2, t & 3
                    da = make_desc(sbase + dyi * 16 + k16 * 2 * CM_PLANE, CM_PLANE, 128)
                    db = make_desc(dyi * WDY + k16 * 2 * LBO_B + pos_lo * NB * 16, LBO_B, 128)
                    umma(tmem, (q0 + pos_lo) * NB, da, db, nrows * NB, smem, True)
                if X - 1 >= x0:
                    done_rows.append(X - 1)
                if X == xhi and xhi == x1 - 1:
                    done_rows.append(X)
                # epilogue for rows completed by this input row
                for x in done_rows:
                    slot = (nb + x - x0) % CM_RING
                    if slot < 2:        # partial sums parked in the mirror slot
                        tmem[:, slot * NB:(slot + 1) * NB] += tmem[:, (CM_RING + slot) * NB:(CM_RING + slot + 1) * NB]
                        tmem[:, (CM_RING + slot) * NB:(CM_RING + slot + 1) * NB] = 0.0
                    for mm in range(128):
                        y = y0 + mm
                        if y >= ny:
                            continue
                        acc = tmem[mm, slot * NB:(slot + 1) * NB].astype(np.float32) + bias
                        if NB == 64:
                            if relu:
                                acc = np.maximum(acc, 0)
                            h = bf16_bits(acc.astype(np.float32))
                            for c in range(8):
                                o = act_offset(b, x, c, y + 1, nx, ny)
                                out_act[o:o + 16] = h[c * 8:(c + 1) * 8].view(np.uint8)
                        else:
                            out_y[b, :, x, y] = acc
                    tmem[:, slot * NB:(slot + 1) * NB] = 0.0          # slot re-zeroed right after it is drained
                done_rows = []
                m += 1
            nb += x1 - x0
    return out_act if NB == 64 else out_y
