"""Index arithmetic of the tcgen05 window-attention kernels against the oracle's index maps (CPU).

Restates, formula by formula, what an element-wise thread of fiber_b200/csrc/window_attn_tc.cu computes for
(query row i, key column j) — fill_tables / WinGeo (window_common.cuh), tc_bj / tc_code and the madd[] select
(window_attn_tc.cu) — and checks it against oracle.fiber_oracle's relative_position_index, shift_attn_mask and
window_token_source, which are pinned to the reference (swin_transformer.py:165-176, :327-350, :366-384).
The compile-time constants of the kernels (12 x 12 windows, shift 6) are what is being checked here.
"""
import pytest
import torch

from oracle import fiber_oracle as O

WS, TW2 = 12, 23


def tc_bj(j):       # window_attn_tc.cu: tc_bj
    return (j // WS) * TW2 + j % WS


def tc_code(j):     # window_attn_tc.cu: tc_code (key side, compile time: shift == ws / 2)
    return (1 if (j // WS) >= WS // 2 else 0) | (2 if (j % WS) >= WS // 2 else 0)


def fill_tables(shift):  # window_common.cuh: fill_tables, per-token part
    aq4, code, tok = [], [], []
    for i in range(WS * WS):
        th, tw = i // WS, i % WS
        bidx = th * TW2 + tw
        aq4.append(4 * (bidx + (WS - 1) * (TW2 + 1)))
        code.append((1 if th >= WS - shift else 0) | (2 if tw >= WS - shift else 0))
        tok.append(th | (tw << 8))
    return aq4, code, tok


def decode(g, H, W, shift):  # WinGeo::decode
    nWw, nWh = W // WS, H // WS
    nW = nWw * nWh
    b, w = divmod(g, nW)
    wh, ww = divmod(w, nWw)
    emask = ((1 if wh == nWh - 1 else 0) | (2 if ww == nWw - 1 else 0)) if shift > 0 else 0
    return b * H * W, wh * WS + shift, ww * WS + shift, emask


def geo_row(img_base, h0, w0, th, tw, H, W):  # WinGeo::row
    hp, wp = h0 + th, w0 + tw
    hp -= H if hp >= H else 0
    wp -= W if wp >= W else 0
    return img_base + hp * W + wp


@pytest.mark.parametrize("H,W,shift", [(24, 24, 6), (24, 24, 0), (36, 24, 6), (12, 12, 0), (48, 96, 6)])
def test_tc_index_maps(H, W, shift):
    aq4, code, tok = fill_tables(shift)
    N = WS * WS
    rpi = O.relative_position_index(WS)
    src = O.window_token_source(H, W, WS, shift)
    mask = O.shift_attn_mask(H, W, WS, shift)
    nW = (H // WS) * (W // WS)
    for g in range(2 * nW):  # two images
        img_base, h0, w0, emask = decode(g, H, W, shift)
        w = g % nW
        for i in range(N):
            th, tw = tok[i] & 255, tok[i] >> 8
            assert geo_row(img_base, h0, w0, th, tw, H, W) == (g // nW) * H * W + int(src[w, i])
        if g >= nW:
            continue
        for i in range(0, N, 5):
            madd = [bool((code[i] ^ c) & emask) for c in range(4)]  # element-wise warps: madd[c] != 0
            for j in range(N):
                assert (aq4[i] - 4 * tc_bj(j)) // 4 == int(rpi[i, j])
                masked = madd[tc_code(j)]
                assert masked == (mask is not None and float(mask[w, i, j]) != 0.0), (g, i, j)


def test_tc_column_ownership():
    """Every (query row < 128, key column) has exactly one element-wise owner and every 16-byte P piece one writer."""
    owners = {}
    for warp in range(8):
        q, hf = warp & 3, warp >> 2
        for lane in range(32):
            row = q * 32 + lane
            for c in range(72):
                owners.setdefault((row, hf * 72 + c), []).append(warp)
            for c8 in range(9):
                p8 = hf * 9 + c8
                assert p8 * 8 == hf * 72 + c8 * 8  # piece p8 holds exactly the thread's columns c8*8 .. c8*8+7
    assert len(owners) == 128 * 144 and all(len(v) == 1 for v in owners.values())
