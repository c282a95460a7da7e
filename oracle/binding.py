"""ctypes binding of the CPU golden model (oracle/libcrgolden.so).  TEST INFRASTRUCTURE."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FLAG_DEPTH, FLAG_LERP, FLAG_QUADS = 1, 2, 4
SHADER = {"passthrough": 0, "gouraud": 1, "texPhong": 2, "gouraudDiscard": 3, "gouraudQuads": 4}
BLEND = {"BlendReplace": 0, "BlendSrcOver": 1, "BlendAdditive": 2, "BlendDepthOnly": 3}


class Config(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("samplesLog2", ctypes.c_int32), ("flags", ctypes.c_uint32),
                ("vertexStride", ctypes.c_int32), ("shader", ctypes.c_int32), ("blend", ctypes.c_int32), ("deferredClear", ctypes.c_int32),
                ("clearColor", ctypes.c_uint32), ("clearDepth", ctypes.c_uint32), ("numThreads", ctypes.c_int32),
                ("fullWidth", ctypes.c_int32), ("fullHeight", ctypes.c_int32), ("centerOfsX", ctypes.c_int32), ("centerOfsY", ctypes.c_int32),
                ("clipLoX", ctypes.c_float), ("clipHiX", ctypes.c_float), ("clipLoY", ctypes.c_float), ("clipHiY", ctypes.c_float),
                ("subX0", ctypes.c_int32), ("subY0", ctypes.c_int32), ("vpWidth", ctypes.c_int32), ("vpHeight", ctypes.c_int32),
                ("cullLoX", ctypes.c_float), ("cullHiX", ctypes.c_float), ("cullLoY", ctypes.c_float), ("cullHiY", ctypes.c_float),
                ("windowed", ctypes.c_int32)]


class Counts(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in ("numTris", "numSubtris", "numVisibleTris", "eBin", "eTile", "eCov", "eShade", "fragments",
                                              "fragmentsWritten", "vertsReferenced")]


def build(with_reference=True):
    """Compiles the oracle (and, when /root/reference is present, oracle/_ref from the reference's own sources)."""
    targets = ["libcrgolden.so"] + (["ref"] if with_reference else [])
    r = subprocess.run(["make", "-C", _HERE] + targets, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the oracle failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libcrgolden.so")
        if not os.path.exists(path):
            build(with_reference=False)
        _LIB = ctypes.CDLL(path)
        L = _LIB
        vp = ctypes.c_void_p
        L.gold_select_flips.restype = ctypes.c_uint32
        L.gold_centroid_code.restype = ctypes.c_uint32
        L.gold_encode_depth.restype = ctypes.c_uint32
        L.gold_encode_depth.argtypes = [ctypes.c_uint32]
        L.gold_clear_depth.restype = ctypes.c_uint32
        L.gold_clear_depth.argtypes = [ctypes.c_float]
        L.gold_to_abgr.restype = ctypes.c_uint32
        L.gold_to_abgr.argtypes = [ctypes.c_float] * 4
        L.gold_blend.restype = ctypes.c_uint32
        L.gold_blend.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32]
        L.gold_clip_triangle.argtypes = [vp, vp, vp, vp]
        L.gold_setup_pleq.argtypes = [vp, vp, vp, vp, ctypes.c_float, ctypes.c_int, vp]
        L.gold_cover_tile.restype = ctypes.c_uint64
        L.gold_cover_tile.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.gold_cover_samples.restype = ctypes.c_uint32
        L.gold_cover_samples.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.gold_triangle_setup.argtypes = [ctypes.POINTER(Config), vp, vp, ctypes.c_int, vp, vp, vp, ctypes.c_int]
        L.gold_render.argtypes = [ctypes.POINTER(Config), vp, vp, ctypes.c_int, vp, vp, ctypes.POINTER(Counts)]
        L.gold_time_render.restype = ctypes.c_double
        L.gold_time_render.argtypes = [ctypes.POINTER(Config), vp, vp, ctypes.c_int, vp, vp, ctypes.c_int]
        L.gold_resolve.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int]
        L.gold_vertex_shader.argtypes = [vp, vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int]
        assert L.gold_sizeof_config() == ctypes.sizeof(Config), "Config layout mismatch"
        assert L.gold_sizeof_counts() == ctypes.sizeof(Counts), "Counts layout mismatch"
    return _LIB


def parent_cell(full, x0, size, limit=2048):
    """(origin, extent) of the parent viewport of a window [x0, x0+size) along one axis of a `full` px frame."""
    n = -(-full // limit)
    cell = (-(-full // n) + 7) & ~7
    p0 = (x0 // cell) * cell
    pw = min(cell, full - p0)
    assert x0 + size <= p0 + pw, "a sort-first window must lie inside one parent cell"
    return p0, pw


def make_config(width, height, samples_log2=0, flags=FLAG_DEPTH, vertex_stride=16, shader="passthrough", blend="BlendReplace",
                clear=None, threads=1, sub=None):
    """clear = (abgr, encodedDepth) or None; sub = (fullW, fullH, x0, y0) for a sort-first window."""
    c = Config()
    c.width, c.height, c.samplesLog2, c.flags = width, height, samples_log2, flags
    c.vertexStride, c.shader, c.blend = vertex_stride, SHADER[shader], BLEND[blend]
    c.deferredClear = 0 if clear is None else 1
    if clear is not None:
        c.clearColor, c.clearDepth = clear
    c.numThreads = threads
    if sub is None:
        c.fullWidth, c.fullHeight, c.centerOfsX, c.centerOfsY = width, height, 0, 0
        c.vpWidth, c.vpHeight = width, height
        c.clipLoX = c.clipLoY = c.cullLoX = c.cullLoY = -1.0
        c.clipHiX = c.clipHiY = c.cullHiX = c.cullHiY = 1.0
        c.subX0 = c.subY0 = 0
        c.windowed = 0
    else:
        # sort-first window: parent viewport = the cell of an even grid of <= 2048 px cells that holds the window
        # (the whole frame when it fits); same rule as cudaraster-linux_b200/csrc/Context.cu prepareFrame()
        fw, fh, x0, y0 = sub
        (px0, pw), (py0, ph) = parent_cell(fw, x0, width), parent_cell(fh, y0, height)
        c.fullWidth, c.fullHeight = fw, fh
        c.vpWidth, c.vpHeight = pw, ph
        c.subX0, c.subY0 = x0 - px0, y0 - py0
        c.centerOfsX = px0 * 16 + pw * 8 - fw * 8
        c.centerOfsY = py0 * 16 + ph * 8 - fh * 8
        c.clipLoX = np.float32(2.0 * px0 / fw - 1.0)
        c.clipHiX = np.float32(2.0 * (px0 + pw) / fw - 1.0)
        c.clipLoY = np.float32(2.0 * py0 / fh - 1.0)
        c.clipHiY = np.float32(2.0 * (py0 + ph) / fh - 1.0)
        c.cullLoX = np.float32(2.0 * x0 / fw - 1.0)
        c.cullHiX = np.float32(2.0 * (x0 + width) / fw - 1.0)
        c.cullLoY = np.float32(2.0 * y0 / fh - 1.0)
        c.cullHiY = np.float32(2.0 * (y0 + height) / fh - 1.0)
        c.windowed = 1
    return c


def clear_values(rgba=(0.2, 0.4, 0.8, 1.0), depth=1.0):
    L = lib()
    return L.gold_to_abgr(*[float(v) for v in rgba]), L.gold_clear_depth(float(depth))


def triangle_setup(cfg, verts, idx, max_subtris=None):
    L = lib()
    verts = np.ascontiguousarray(verts, np.float32)
    idx = np.ascontiguousarray(idx, np.int32)
    n = idx.shape[0]
    cap = n * 7 + 16 if max_subtris is None else max_subtris
    sub = np.zeros(max(n, 1), np.uint8)
    hdr = np.zeros((cap, 4), np.uint32)
    dat = np.zeros((cap, 16), np.uint32)
    num = L.gold_triangle_setup(ctypes.byref(cfg), verts.ctypes.data, idx.ctypes.data, n, sub.ctypes.data, hdr.ctypes.data, dat.ctypes.data, cap)
    return {"numSubtris": num, "triSubtris": sub[:n], "triHeader": hdr[:min(num, cap)], "triData": dat[:min(num, cap)]}


def render(cfg, verts, idx, color=None, depth=None, want_counts=False):
    """Renders one frame; color/depth are U32 arrays [roundedH, roundedW*N] (in/out)."""
    L = lib()
    verts = np.ascontiguousarray(verts, np.float32)
    idx = np.ascontiguousarray(idx, np.int32)
    n_s = 1 << cfg.samplesLog2
    rw, rh = (cfg.width + 7) & ~7, (cfg.height + 7) & ~7
    if color is None:
        color = np.zeros((rh, rw * n_s), np.uint32)
    if depth is None:
        depth = np.zeros((rh, rw * n_s), np.uint32)
    counts = Counts()
    num = L.gold_render(ctypes.byref(cfg), verts.ctypes.data, idx.ctypes.data, idx.shape[0], color.ctypes.data, depth.ctypes.data,
                        ctypes.byref(counts) if want_counts else None)
    out = {"color": color, "depth": depth, "numSubtris": num}
    if want_counts:
        out["counts"] = {k: getattr(counts, k) for k, _ in Counts._fields_}
    return out


def time_render(cfg, verts, idx, reps=3):
    L = lib()
    verts = np.ascontiguousarray(verts, np.float32)
    idx = np.ascontiguousarray(idx, np.int32)
    n_s = 1 << cfg.samplesLog2
    rw, rh = (cfg.width + 7) & ~7, (cfg.height + 7) & ~7
    color = np.zeros((rh, rw * n_s), np.uint32)
    depth = np.zeros((rh, rw * n_s), np.uint32)
    return L.gold_time_render(ctypes.byref(cfg), verts.ctypes.data, idx.ctypes.data, idx.shape[0], color.ctypes.data, depth.ctypes.data, reps)


def resolve(surface, width, height, num_samples, flip_y=False):
    """Box-filter resolve of a [roundedH, roundedW*N] U32 surface into a [height, width] image."""
    src = np.ascontiguousarray(surface, np.uint32)
    dst = np.zeros((height, width), np.uint32)
    lib().gold_resolve(src.ctypes.data, width, height, num_samples, dst.ctypes.data, width, 1 if flip_y else 0)
    return dst


def vertex_shader(matrix_colmajor, in_verts, out_stride_floats):
    """clipPos = M * (modelPos, 1); input floats 3.. are carried through behind clipPos."""
    m = np.ascontiguousarray(matrix_colmajor, np.float32).reshape(16)
    vin = np.ascontiguousarray(in_verts, np.float32)
    out = np.zeros((vin.shape[0], out_stride_floats), np.float32)
    lib().gold_vertex_shader(m.ctypes.data, vin.ctypes.data, vin.shape[1], out.ctypes.data, out_stride_floats, vin.shape[0])
    return out


def hardware_threads():
    return lib().gold_hardware_threads()


def algorithmic_bytes(counts, num_tris, k_varyings, pixels, samples, lerp, deferred_clear=True):
    """B_alg of SURVEY.md 8(d), from golden-model counts."""
    c = counts
    t, tsub, v = num_tris, c["numSubtris"], c["vertsReferenced"]
    b = 12 * t + 16 * v + 16 * k_varyings * v      # indices, clip positions, varyings touched
    b += 1 * t + 80 * tsub                          # setup out
    b += 1 * t + 16 * tsub                          # bin in
    b += (8 + 16) * c["eBin"]                       # bin queue write+read, header re-read
    b += (8 + 16) * c["eTile"]                      # tile queue write+read, header re-read
    b += 16 * c["eCov"]                             # depth plane
    if lerp:
        b += 48 * c["eShade"]                       # w/u/v planes
    b += 8 * pixels * samples * (1 if deferred_clear else 2)
    return int(b)
