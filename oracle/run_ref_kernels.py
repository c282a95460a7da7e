"""Runs the REFERENCE's own CUDA kernels, rebuilt for sm_100a (oracle/_ref/libcrref_cuda.so, built by
oracle/Makefile from the sources under /root/reference), on one of the benchmark workloads.
TEST / BASELINE INFRASTRUCTURE -- never imported by the product.

Meant to be run as a subprocess with a timeout: the Fermi code is implicitly warp-synchronous
 and can hang or fault on Blackwell.  Prints one JSON line:
  {"status": "ok"|"mismatch"|..., "stage_ms": {...}, "frame_ms": ..., "Mtris/s": ..., "depth_mismatch_texels": n, ...}

    python oracle/run_ref_kernels.py --workload c2 [--frames 7] [--check]
"""
import argparse
import ctypes
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def load(path=None):
    lib = ctypes.CDLL(path or os.environ.get("CRREF_LIBRARY") or os.path.join(ROOT, "oracle", "_ref", "libcrref_cuda.so"))
    vp = ctypes.c_void_p
    lib.crref_last_error.restype = ctypes.c_char_p
    lib.crref_pipe_name.restype = ctypes.c_char_p
    lib.crref_draw.argtypes = [ctypes.c_char_p, vp, ctypes.c_size_t, vp, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                               ctypes.c_uint32, ctypes.c_uint32, vp, vp]
    lib.crref_get_setup_output.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, vp]
    return lib


def draw(lib, pipe, verts, idx, w, h, samples_log2, clear=(0xFFCC6633, 0xFFFFBB3F), frames=1):
    """Renders `frames` times; returns (color, depth, [stage seconds per frame], atomics)."""
    import torch
    n = 1 << samples_log2
    rw, rh = (w + 7) & ~7, (h + 7) & ~7
    vb = torch.from_numpy(np.ascontiguousarray(verts, np.float32)).cuda()
    ib = torch.from_numpy(np.ascontiguousarray(idx, np.int32)).cuda()
    color = torch.zeros((rh, rw * n), dtype=torch.int32, device="cuda")
    depth = torch.zeros((rh, rw * n), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    times, atomics = [], (ctypes.c_int * 7)()
    for _ in range(frames):
        st = (ctypes.c_float * 4)()
        rc = lib.crref_draw(pipe.encode(), vb.data_ptr(), vb.numel() * 4, ib.data_ptr(), idx.shape[0], color.data_ptr(), depth.data_ptr(), w, h, n,
                            1, clear[0], clear[1], st, atomics)
        if rc != 0:
            raise RuntimeError("crref_draw: " + lib.crref_last_error().decode())
        times.append(list(st))
    torch.cuda.synchronize()
    return color.cpu().numpy().view(np.uint32), depth.cpu().numpy().view(np.uint32), times, list(atomics)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--frames", type=int, default=7)
    ap.add_argument("--check", action="store_true", help="compare depth/colour with the CPU oracle")
    ap.add_argument("--check-product", action="store_true", help="compare depth/colour with the B200 pipeline (libcrb200.so)")
    ap.add_argument("--check-setup", action="store_true", help="compare triSubtris/triHeader/triData with the CPU oracle, bit for bit")
    args = ap.parse_args()
    import bench
    blend = "BlendReplace"
    if args.workload.startswith("soup"):
        # parity scene: mixed sizes, frustum-crossing and w <= 0 triangles (exercises the clipper)
        import cudaraster_linux_b200 as crb
        shader, flags, s_log2 = {"soup": ("gouraud", 3, 0), "soup_msaa": ("gouraud", 3, 2), "soup_pass": ("passthrough", 1, 0), "soup_blend": ("gouraud", 3, 0), "soup_msaa_front": ("gouraud", 3, 2)}[args.workload]
        if args.workload == "soup_blend":
            blend = "BlendSrcOver"   # reads dst: every fragment of a pixel must reach the ROP in submission order
        w, h, desc = 640, 360, args.workload
        # soup_msaa_front: no triangle with a vertex at w <= 0 (their clipped remains have ill-conditioned depth planes, DESIGN.md "known divergences")
        verts, idx = crb.scenes.random_soup(20000, seed=1237, stride_floats=8 if shader == "gouraud" else 4, **({"behind_fraction": 0.0} if args.workload == "soup_msaa_front" else {}))
        if blend != "BlendReplace":
            verts[:, 7] = np.random.default_rng(3).uniform(0.2, 1.0, verts.shape[0]).astype(np.float32)   # alpha
    elif args.workload == "c1":
        import cudaraster_linux_b200 as crb
        w, h, desc, shader, flags, s_log2 = 1024, 768, "C1 cube", "passthrough", 1, 0
        verts, idx = crb.scenes.cube(w, h)
    elif args.workload == "ties":
        # the same mesh submitted twice with other colours: every fragment of the second copy ties in depth with the first and must lose
        import cudaraster_linux_b200 as crb
        w, h, desc, shader, flags, s_log2 = 512, 384, "duplicated grid (depth ties)", "gouraud", 3, 0
        v, i = crb.scenes.grid_gouraud(160, 100)
        verts = np.concatenate([v, v])
        verts[v.shape[0]:, 4:8] = np.random.default_rng(9).uniform(0, 1, (v.shape[0], 4)).astype(np.float32)
        idx = np.concatenate([i, i + v.shape[0]])
    else:
        desc, verts, idx, w, h, shader, s_log2, flags, _ = bench.make_scene(args.workload)
    pipe = "ref_%s_s%d_f%d_%s" % (shader, s_log2, flags, blend)
    lib = load()
    names = [lib.crref_pipe_name(i).decode() for i in range(lib.crref_num_pipes())]
    if pipe not in names:
        print(json.dumps({"status": "no reference pipe %s (the reference ships no such shader)" % pipe}))
        return 0
    color, depth, times, atomics = draw(lib, pipe, verts, idx, w, h, s_log2, frames=args.frames)
    steady = times[2:] if len(times) > 3 else times
    med = [statistics.median(t[i] for t in steady) * 1e3 for i in range(4)]
    libname = os.path.basename(os.environ.get("CRREF_LIBRARY") or "libcrref_cuda.so")
    out = {"status": "ok", "library": libname,
           "what": ("reference kernels with the synchronisation patch (oracle/ref_kernels/b200_sync_patch.py: lock-step assumptions made explicit, nothing else changed)"
                    if "sync" in libname else "reference kernels, unmodified sources") + ", rebuilt for sm_100a behind oracle/ref_kernels/shim.h, launch shapes of CudaRaster.cpp:593-655",
           "stage_ms": dict(zip(("triangleSetup", "binRaster", "coarseRaster", "fineRaster"), med)), "frame_ms": sum(med),
           "Mtris/s": idx.shape[0] / (sum(med) * 1e-3) / 1e6, "atomics": atomics}
    if args.check_setup:
        from tests import util
        n = idx.shape[0]
        nsub = atomics[0]
        sub = np.zeros(n, np.uint8)
        hdr = np.zeros((nsub, 4), np.uint32)
        dat = np.zeros((nsub, 16), np.uint32)
        if lib.crref_get_setup_output(n, nsub, sub.ctypes.data, hdr.ctypes.data, dat.ctypes.data) != 0:
            raise RuntimeError(lib.crref_last_error().decode())
        gs = util.gold_setup(verts, idx, w, h, shader, flags, s_log2)
        try:
            n_single, n_multi = util.compare_setup({"triSubtris": sub, "triHeader": hdr, "triData": dat, "counters": {"numSubtris": nsub}}, gs, n, flags)
            out["setup"] = {"status": "bit-exact", "single": n_single, "clipped": n_multi, "numSubtris": int(nsub)}
        except AssertionError as e:
            out["setup"] = {"status": "mismatch: %s" % str(e)[:200]}
    if args.check:
        from tests import util
        g = util.draw_gold(verts, idx, w, h, shader, flags, s_log2, blend)
        out["depth_mismatch_texels"] = int((depth != g["depth"]).sum())
        out["color_max_lsb"] = util.color_max_diff(color, g["color"])
        out["color_mismatch_texels"] = int((color != g["color"]).sum())
        if 0 < out["depth_mismatch_texels"] <= 8:   # few enough to list: (row, texel column, reference depth, oracle depth)
            ys, xs = np.nonzero(depth != g["depth"])
            out["depth_mismatches"] = [[int(y), int(x), int(depth[y, x]), int(g["depth"][y, x])] for y, x in zip(ys, xs)]
        if out["depth_mismatch_texels"]:
            out["status"] = "mismatch"
    if args.check_product:
        # the B200 pipeline (through its C ABI) against the reference's kernels, frame against frame
        import cudaraster_linux_b200 as crb
        from tests import util
        r = crb.CudaRaster(0)
        cc, cd = util.draw_cuda(r, crb, verts, idx, w, h, shader, flags, s_log2, blend, pipe="PixelPipe_passthrough" if args.workload == "c1" else None)
        out["product_depth_mismatch_texels"] = int((cd != depth).sum())
        out["product_color_max_lsb"] = util.color_max_diff(cc, color)
        r.close()
        if out["product_depth_mismatch_texels"] and out["status"] == "ok":
            out["status"] = "mismatch"
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
