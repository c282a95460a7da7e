"""Pins the CPU oracle against the REFERENCE's own device functions executed on the GPU
(oracle/_ref/libcrref_cuda.so, per-thread harness at the end of oracle/ref_kernels/ref_driver.cu).
TEST INFRASTRUCTURE -- never imported by the product.

Unlike the reference's bin / coarse / fine KERNELS (implicitly warp-synchronous, mis-execute on
sm_100a), these functions are plain per-thread code and run correctly on Blackwell:

  coverage   trianglePixelCoverage<0/1>   FineRaster.inl:245-282 -> cover8x8_exact_fast / _conservative_fast +
                                          cover8x8_setupLUT / cover8x8_lookupMask, Util.inl:148-271   (the LUT path)
  samples    triangleSampleCoverage<S>    FineRaster.inl:286-309 -> coverMSAA_fast, Util.inl:361-383
  shading    runFragmentShader + GouraudShader   FineRaster.inl:49-119, PixelPipe.inl:43-69
  blending   runBlendShader + Blend*      FineRaster.inl:123-142, PixelPipe.inl:73-85, Util.inl:42-60

Prints one JSON line with the number of compared values and mismatches per family.

    python oracle/run_ref_devfuncs.py [--width 2048 --height 2048 --tris 40000 --seed 99]
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def load():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcrref_cuda.so"))
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.crref_last_error.restype = ctypes.c_char_p
    lib.crref_cover_tiles.argtypes = [vp, vp, ci, ci, ci, vp, vp, vp]
    lib.crref_cover_samples.argtypes = [vp, vp, ci, ci, ci, ci, vp]
    lib.crref_shade_gouraud.argtypes = [vp, vp, vp, ci, ci, vp, vp]
    lib.crref_blend.argtypes = [ci, vp, vp, ci, vp, vp]
    return lib


def ck(lib, rc):
    if rc != 0:
        raise RuntimeError("reference harness: " + lib.crref_last_error().decode())


def tile_pairs(hdr, w, h, rng, per_tri=6):
    """(header index, tileX, tileY) for tiles in and around every sub-triangle's bounding box."""
    x = np.stack([(hdr[:, k] & 0xFFFF).astype(np.int16).astype(np.int64) for k in range(3)], 1)
    y = np.stack([(hdr[:, k] >> 16).astype(np.int16).astype(np.int64) for k in range(3)], 1)
    # viewport-centred subpixels -> pixels
    lox = np.floor((x.min(1) + w * 8) / 16).astype(np.int64) >> 3
    hix = np.floor((x.max(1) + w * 8) / 16).astype(np.int64) >> 3
    loy = np.floor((y.min(1) + h * 8) / 16).astype(np.int64) >> 3
    hiy = np.floor((y.max(1) + h * 8) / 16).astype(np.int64) >> 3
    tw, th = (w + 7) >> 3, (h + 7) >> 3
    out = []
    n = hdr.shape[0]
    for _ in range(per_tri):
        tx = lox - 1 + (rng.random(n) * (hix - lox + 3)).astype(np.int64)
        ty = loy - 1 + (rng.random(n) * (hiy - loy + 3)).astype(np.int64)
        ok = (tx >= 0) & (tx < tw) & (ty >= 0) & (ty < th)
        out.append(np.stack([np.arange(n)[ok], tx[ok], ty[ok]], 1))
    return np.ascontiguousarray(np.concatenate(out, 0), np.int32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--height", type=int, default=2048)
    ap.add_argument("--tris", type=int, default=40000)
    ap.add_argument("--seed", type=int, default=99)
    args = ap.parse_args()
    import torch
    import cudaraster_linux_b200 as crb
    from oracle import binding as G
    from tests import util
    lib, L = load(), G.lib()
    L.gold_run_shader.argtypes = [ctypes.POINTER(G.Config), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
                                  ctypes.c_void_p, ctypes.c_void_p]
    w, h = args.width, args.height
    rng = np.random.default_rng(args.seed)
    out = {"status": "ok", "viewport": [w, h]}

    # ---- scene: mixed sizes (sub-pixel .. hundreds of px), both windings, frustum crossers -------------
    parts = [crb.scenes.random_soup(args.tris // 4, seed=args.seed + k, stride_floats=8, size=s) for k, s in enumerate((0.004, 0.02, 0.2, 1.5))]
    verts = np.concatenate([p[0] for p in parts], 0)
    idx = np.concatenate([p[1] + sum(q[0].shape[0] for q in parts[:k]) for k, p in enumerate(parts)], 0).astype(np.int32)

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).cuda()

    for s_log2 in (0, 1, 2, 3):
        gs = util.gold_setup(verts, idx, w, h, "gouraud", 3, s_log2)
        hdr, dat = np.ascontiguousarray(gs["triHeader"]), np.ascontiguousarray(gs["triData"])
        # live sub-triangle slots: singles in their own slot, clipped runs behind misc
        sub = gs["triSubtris"]
        live = list(np.nonzero(sub == 1)[0])
        for t in np.nonzero(sub > 1)[0]:
            live += list(range(int(hdr[t, 3]), int(hdr[t, 3]) + int(sub[t])))
        live = np.array(live, np.int64)
        d_hdr = dev(hdr.view(np.int32))
        key = "s%d" % s_log2

        if s_log2 == 0:
            # ---- 8x8 LUT coverage vs the oracle's exact rule ------------------------------------------------
            pairs = tile_pairs(hdr[live], w, h, rng)
            pairs[:, 0] = live[pairs[:, 0]]
            d_pairs = dev(pairs)
            d_ex = torch.zeros(pairs.shape[0], dtype=torch.int64, device="cuda")
            d_co = torch.zeros_like(d_ex)
            d_lut = torch.zeros(768, dtype=torch.int64, device="cuda")
            ck(lib, lib.crref_cover_tiles(d_hdr.data_ptr(), d_pairs.data_ptr(), pairs.shape[0], w, h, d_ex.data_ptr(), d_co.data_ptr(), d_lut.data_ptr()))
            ex = d_ex.cpu().numpy().view(np.uint64)
            co = d_co.cpu().numpy().view(np.uint64)
            gold = np.array([L.gold_cover_tile(hdr[p[0]].ctypes.data, w, h, int(p[1]), int(p[2])) for p in pairs], np.uint64)
            bad = np.nonzero(ex != gold)[0]
            out["cover8x8_exact_fast"] = {"pairs": int(pairs.shape[0]), "nonempty": int((gold != 0).sum()), "full": int((gold == np.uint64(0xFFFFFFFFFFFFFFFF)).sum()),
                                          "mismatch": int(bad.size)}
            if bad.size:
                out["cover8x8_exact_fast"]["first"] = [[int(v) for v in pairs[b]] + [hex(int(ex[b])), hex(int(gold[b]))] for b in bad[:4]]
                out["status"] = "mismatch"
            out["cover8x8_conservative_fast"] = {"pairs": int(pairs.shape[0]), "not_superset_of_exact": int(((gold & ~co) != 0).sum())}
            if out["cover8x8_conservative_fast"]["not_superset_of_exact"]:
                out["status"] = "mismatch"
        else:
            # ---- MSAA sample coverage of single pixels vs the oracle ------------------------------------------
            sel = live[rng.integers(0, live.size, size=min(live.size, 60000))]
            x = np.stack([(hdr[sel, k] & 0xFFFF).astype(np.int16).astype(np.int64) for k in range(3)], 1)
            y = np.stack([(hdr[sel, k] >> 16).astype(np.int16).astype(np.int64) for k in range(3)], 1)
            lox, hix = (x.min(1) + w * 8) >> 4, (x.max(1) + w * 8) >> 4
            loy, hiy = (y.min(1) + h * 8) >> 4, (y.max(1) + h * 8) >> 4
            px = lox - 1 + (rng.random(sel.size) * (hix - lox + 3)).astype(np.int64)
            py = loy - 1 + (rng.random(sel.size) * (hiy - loy + 3)).astype(np.int64)
            ok = (px >= 0) & (px < w) & (py >= 0) & (py < h)
            pairs = np.ascontiguousarray(np.stack([sel[ok], px[ok], py[ok]], 1), np.int32)
            d_pairs = dev(pairs)
            d_out = torch.zeros(pairs.shape[0], dtype=torch.int32, device="cuda")
            ck(lib, lib.crref_cover_samples(d_hdr.data_ptr(), d_pairs.data_ptr(), pairs.shape[0], w, h, s_log2, d_out.data_ptr()))
            got = d_out.cpu().numpy().view(np.uint32)
            gold = np.array([L.gold_cover_samples(hdr[p[0]].ctypes.data, w, h, s_log2, int(p[1]), int(p[2])) for p in pairs], np.uint32)
            bad = np.nonzero(got != gold)[0]
            out["coverMSAA_fast_" + key] = {"pairs": int(pairs.shape[0]), "nonempty": int((gold != 0).sum()), "partial": int(((gold != 0) & (gold != (1 << (1 << s_log2)) - 1)).sum()),
                                            "mismatch": int(bad.size)}
            if bad.size:
                out["status"] = "mismatch"

        # ---- fragment shader front end: barycentrics + Gouraud colour ------------------------------------
        cfg = G.make_config(w, h, s_log2, 3, 32, "gouraud", "BlendReplace")
        sel = live[rng.integers(0, live.size, size=min(live.size, 50000))]
        x = np.stack([(hdr[sel, k] & 0xFFFF).astype(np.int16).astype(np.int64) for k in range(3)], 1)
        y = np.stack([(hdr[sel, k] >> 16).astype(np.int16).astype(np.int64) for k in range(3)], 1)
        px = np.clip((x.sum(1) // 3 + w * 8) >> 4, 0, w - 1)
        py = np.clip((y.sum(1) // 3 + h * 8) >> 4, 0, h - 1)
        n_s = 1 << s_log2
        masks = rng.integers(0, 1 << n_s, size=sel.size)
        cen = np.array([L.gold_centroid_code(s_log2, int(m)) for m in masks], np.int64)
        frags = np.ascontiguousarray(np.stack([sel, px, py, cen], 1), np.int32)
        d_frags, d_dat, d_verts = dev(frags), dev(dat.view(np.int32)), dev(np.ascontiguousarray(verts, np.float32))
        d_col = torch.zeros(frags.shape[0], dtype=torch.int32, device="cuda")
        d_bary = torch.zeros(frags.shape[0] * 6, dtype=torch.float32, device="cuda")
        ck(lib, lib.crref_shade_gouraud(d_dat.data_ptr(), d_verts.data_ptr(), d_frags.data_ptr(), frags.shape[0], s_log2, d_col.data_ptr(), d_bary.data_ptr()))
        col = d_col.cpu().numpy().view(np.uint32)
        bary = d_bary.cpu().numpy().reshape(-1, 6)
        vv = np.ascontiguousarray(verts, np.float32)
        gcol = np.zeros(frags.shape[0], np.uint32)
        gbary = np.zeros((frags.shape[0], 6), np.float32)
        c1, b6 = ctypes.c_uint32(0), (ctypes.c_float * 6)()
        for i, q in enumerate(frags):
            L.gold_run_shader(ctypes.byref(cfg), vv.ctypes.data, dat.ctypes.data, int(q[0]), int(q[1]), int(q[2]), int(q[3]), ctypes.byref(c1), b6)
            gcol[i] = c1.value
            gbary[i] = b6[:]
        same_bary = (bary.view(np.uint32) == gbary.view(np.uint32)) | (np.isnan(bary) & np.isnan(gbary))
        out["shade_gouraud_" + key] = {"fragments": int(frags.shape[0]), "bary_bit_mismatch": int((~same_bary).any(1).sum()), "color_mismatch": int((col != gcol).sum()),
                                       "color_max_lsb": util.color_max_diff(col, gcol)}
        if out["shade_gouraud_" + key]["bary_bit_mismatch"] or out["shade_gouraud_" + key]["color_max_lsb"] > 1:
            out["status"] = "mismatch"

    # ---- blend shaders --------------------------------------------------------------------------------
    n = 200000
    src = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    dst = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    src[:256] = np.arange(256, dtype=np.uint32) * 0x01010101
    dst[:256] = 0xFFFFFFFF
    d_src, d_dst = dev(src.view(np.int32)), dev(dst.view(np.int32))
    for name, bid in G.BLEND.items():
        d_out = torch.zeros(n, dtype=torch.int32, device="cuda")
        d_wr = torch.zeros(n, dtype=torch.int32, device="cuda")
        ck(lib, lib.crref_blend(bid, d_src.data_ptr(), d_dst.data_ptr(), n, d_out.data_ptr(), d_wr.data_ptr()))
        got, wr = d_out.cpu().numpy().view(np.uint32), d_wr.cpu().numpy()
        res = np.where(wr != 0, got, dst)   # what ends up in the colour buffer
        gold = np.array([L.gold_blend(bid, int(s), int(d)) for s, d in zip(src[:20000], dst[:20000])], np.uint32)
        out[name] = {"pairs": 20000, "mismatch": int((res[:20000] != gold).sum())}
        if out[name]["mismatch"]:
            out["status"] = "mismatch"
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
