"""Helpers that drive the reference's own CPU code (oracle/_ref/libgpujpeg_refcpu.so) from numpy."""
import numpy as np

import _oracle as o


def ref_encode_coef(coef, w, h, quality, rst, interleaved=0, comps=3, sampling=(1, 1)):
    out = np.empty(4096 + coef.size * 8, np.uint8)
    n = o.ref.ref_encode_from_coef_ss(np.ascontiguousarray(coef).reshape(-1), w, h, comps, quality, rst, interleaved,
                                      sampling[0], sampling[1], out, out.size)
    assert n > 0
    return out[:n].copy()


def split_segments(jpeg):
    """Marker walk equivalent to the reference reader's scan splitter: returns (scan, index, offset, size)
    arrays over the stuffed entropy-coded bytes (markers excluded) and the DHT tables by [class][id]."""
    j = bytes(jpeg)
    i = 2
    bits = np.zeros((2, 2, 17), np.uint8)
    vals = np.zeros((2, 2, 256), np.uint8)
    segs = []
    scan = 0
    while i < len(j):
        assert j[i] == 0xFF
        m = j[i + 1]
        if m == 0xD9:
            break
        ln = (j[i + 2] << 8) | j[i + 3]
        if m == 0xC4:
            p = i + 4
            while p < i + 2 + ln:
                tc, th = j[p] >> 4, j[p] & 15
                b = np.frombuffer(j[p + 1:p + 17], np.uint8)
                n = int(b.sum())
                bits[tc, th, 1:] = b
                vals[tc, th, :n] = np.frombuffer(j[p + 17:p + 17 + n], np.uint8)
                p += 17 + n
        if m == 0xDA:
            p = i + 2 + ln
            start = p
            idx = 0
            while True:
                if j[p] == 0xFF and j[p + 1] != 0:
                    mm = j[p + 1]
                    segs.append((scan, idx, start, p - start))
                    if 0xD0 <= mm <= 0xD7:
                        p += 2
                        start = p
                        idx += 1
                        continue
                    break
                p += 2 if j[p] == 0xFF else 1
            scan += 1
            i = p
            continue
        i += 2 + ln
    a = np.array(segs, np.int64)
    return a, bits, vals


def ref_decode_coef(jpeg, w, h, rst, interleaved=0, comps=3, sampling=(1, 1)):
    segs, bits, vals = split_segments(jpeg)
    if tuple(sampling) == (1, 1):
        dw, dh = (w + 7) // 8 * 8, (h + 7) // 8 * 8
        coef = np.zeros((comps, dw * dh), np.int16)
    else:
        coef = np.zeros(o.coef_count(w, h, sampling, interleaved, comps), np.int16)
    data = np.ascontiguousarray(jpeg, np.uint8)
    rc = o.ref.ref_huff_decode_ss(data, data.size, w, h, comps, rst, interleaved, sampling[0], sampling[1], len(segs),
                               segs[:, 0].astype(np.int32).copy(), segs[:, 1].astype(np.int32).copy(),
                               segs[:, 2].astype(np.uint64).copy(), segs[:, 3].astype(np.uint64).copy(),
                               np.ascontiguousarray(bits).reshape(-1), np.ascontiguousarray(vals).reshape(-1),
                               coef.reshape(-1))
    assert rc == 0, rc
    return coef


def ref_idct_planes(coef, w, h, quality):
    """reference gpujpeg_idct_cpu_perform per block, then +128 / clamp / de-block as gpujpeg_idct_cpu does."""
    raw = np.zeros((2, 64), np.uint8)
    fwd = np.zeros((2, 64), np.float32)
    inv = np.zeros((2, 64), np.uint16)
    assert o.ref.ref_quant_tables(quality, raw, fwd, inv) == 0
    dw, dh = (w + 7) // 8 * 8, (h + 7) // 8 * 8
    comps = coef.shape[0]
    planes = np.zeros((comps, dh, dw), np.uint8)
    for c in range(comps):
        blocks = coef[c].reshape(-1, 64).copy()
        for b in range(blocks.shape[0]):
            o.ref.ref_idct_block(blocks[b], inv[0 if c == 0 else 1])
        px = np.clip(blocks.astype(np.int32) + 128, 0, 255).astype(np.uint8)
        px = px.reshape(dh // 8, dw // 8, 8, 8).transpose(0, 2, 1, 3).reshape(dh, dw)
        planes[c] = px
    return planes
