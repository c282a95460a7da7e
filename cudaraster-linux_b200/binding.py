"""ctypes binding of include/crb200.h and the Python mirror of the reference host API."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RenderModeFlag_EnableDepth = 1
RenderModeFlag_EnableLerp = 2
RenderModeFlag_EnableQuads = 4


class CrbError(RuntimeError):
    """Raised where the reference would call fail() (print + exit)."""


class _PipeSpec(ctypes.Structure):
    _fields_ = [("samplesLog2", ctypes.c_int32), ("vertexStructSize", ctypes.c_int32), ("renderModeFlags", ctypes.c_uint32),
                ("profilingMode", ctypes.c_int32), ("blendShaderName", ctypes.c_char * 128)]


class Atomics(ctypes.Structure):
    _fields_ = [("numSubtris", ctypes.c_int32), ("numBinEntries", ctypes.c_int32), ("numCoarseItems", ctypes.c_int32),
                ("numTileEntries", ctypes.c_int32), ("numActiveTiles", ctypes.c_int32), ("overflow", ctypes.c_int32),
                ("numLargeTris", ctypes.c_int32), ("numQueuedCtas", ctypes.c_int32), ("allocBarrier", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class BatchFrame(ctypes.Structure):
    """crb_batch_frame (include/crb200.h)."""
    _fields_ = [("color", ctypes.c_void_p), ("depth", ctypes.c_void_p), ("width", ctypes.c_int32), ("height", ctypes.c_int32), ("numSamples", ctypes.c_int32),
                ("vertices", ctypes.c_void_p), ("vertexBytes", ctypes.c_size_t), ("indices", ctypes.c_void_p), ("numTris", ctypes.c_int32),
                ("clear", ctypes.c_int32), ("clearColor", ctypes.c_uint32), ("clearDepth", ctypes.c_uint32),
                ("pushDst", ctypes.c_void_p), ("pushBytes", ctypes.c_size_t), ("signalWord", ctypes.c_void_p), ("signalValue", ctypes.c_uint32), ("surfaceSlot", ctypes.c_int32)]


class WorkBuffers(ctypes.Structure):
    _fields_ = [("triSubtris", ctypes.c_void_p), ("triHeader", ctypes.c_void_p), ("triData", ctypes.c_void_p), ("maxSubtris", ctypes.c_int32),
                ("binQueue", ctypes.c_void_p), ("binStart", ctypes.c_void_p), ("binTotal", ctypes.c_void_p), ("numBins", ctypes.c_int32),
                ("tileQueue", ctypes.c_void_p), ("tileStart", ctypes.c_void_p), ("tileCount", ctypes.c_void_p), ("numTiles", ctypes.c_int32),
                ("activeTiles", ctypes.c_void_p)]


def library_path():
    # CRB200_LIBRARY: an experiment build of the same library (tools/build_variant.sh); never a different implementation
    return os.environ.get("CRB200_LIBRARY") or os.path.join(_HERE, "libcrb200.so")


def build_library(verbose=False):
    """Compiles every CUDA source for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", _HERE, "-j8", "libcrb200.so"], capture_output=True, text=True)
    if r.returncode != 0:
        raise CrbError("building libcrb200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return library_path()


def load_library():
    """Loads libcrb200.so.  Fails loudly when it is missing: there is no fallback path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise CrbError("libcrb200.so is not built (%s); run __graft_entry__.build() or make -C cudaraster-linux_b200" % path)
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    vp, i32, u32, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_float
    sigs = {
        "crb_abi_version": (i32, []),
        "crb_create": (i32, [i32, ctypes.POINTER(vp)]),
        "crb_destroy": (i32, [vp]),
        "crb_last_error": (ctypes.c_char_p, [vp]),
        "crb_set_surfaces": (i32, [vp, vp, vp, i32, i32, i32]),
        "crb_deferred_clear": (i32, [vp, u32, u32]),
        "crb_pack_abgr": (u32, [f32, f32, f32, f32]),
        "crb_encode_clear_depth": (u32, [f32]),
        "crb_set_pixel_pipe": (i32, [vp, vp]),
        "crb_set_pixel_pipe_by_name": (i32, [vp, vp, ctypes.c_char_p]),
        "crb_set_vertex_buffer": (i32, [vp, vp, ctypes.c_size_t]),
        "crb_set_index_buffer": (i32, [vp, vp, i32]),
        "crb_set_subviewport": (i32, [vp, i32, i32, i32, i32]),
        "crb_draw_triangles": (i32, [vp, vp]),
        "crb_draw_triangles_async": (i32, [vp, vp]),
        "crb_finish": (i32, [vp, vp]),
        "crb_draw_triangles_host": (i32, [vp, vp, ctypes.c_size_t, vp, i32, vp, vp, vp]),
        "crb_draw_triangles_host_async": (i32, [vp, vp, ctypes.c_size_t, vp, i32, vp, vp, vp]),
        "crb_get_stats": (i32, [vp, ctypes.POINTER(f32 * 4)]),
        "crb_set_stage_timing": (i32, [vp, i32]),
        "crb_get_stage_timing": (i32, [vp, ctypes.POINTER(ctypes.c_double * 4), ctypes.POINTER(i32)]),
        "crb_get_stage_timing_frames": (i32, [vp, vp, i32]),
        "crb_draw_batch_async": (i32, [vp, vp, i32, vp]),
        "crb_batch_join": (i32, [vp, vp]),
        "crb_compute_chunk_bounds": (i32, [vp, i32, vp, i32, vp, vp]),
        "crb_set_chunk_bounds": (i32, [vp, vp]),
        "crb_split_frame": (i32, [i32, i32, i32, vp, i32]),
        "crb_get_counters": (i32, [vp, ctypes.POINTER(Atomics)]),
        "crb_get_profiling_info": (i32, [vp, ctypes.c_char_p, ctypes.c_size_t]),
        "crb_get_launch_count": (i32, [vp]),
        "crb_get_work_buffers": (i32, [vp, ctypes.POINTER(WorkBuffers)]),
        "crb_download": (i32, [vp, vp, vp, ctypes.c_size_t]),
        "crb_set_binning_mode": (i32, [vp, i32]),
        "crb_get_last_frame_direct": (i32, [vp]),
        "crb_set_color_layout": (i32, [vp, i32]),
        "crb_set_color_pitch": (i32, [vp, i32]),
        "crb_ipc_alloc": (i32, [ctypes.c_size_t, ctypes.POINTER(vp), ctypes.c_char_p]),
        "crb_ipc_free": (i32, [vp]),
        "crb_ipc_open": (i32, [ctypes.c_char_p, ctypes.POINTER(vp)]),
        "crb_ipc_close": (i32, [vp]),
        "crb_ipc_signal": (i32, [vp, u32, vp]),
        "crb_ipc_copy": (i32, [vp, vp, ctypes.c_size_t, vp]),
        "crb_ipc_delay": (i32, [vp, u32]),
        "crb_ipc_copy_2d": (i32, [vp, ctypes.c_size_t, vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, vp]),
        "crb_resolve_surface": (i32, [vp, i32, i32, i32, vp, i32, i32, vp]),
        "crb_write_ppm": (i32, [ctypes.c_char_p, vp, i32, i32, i32]),
        "crb_launch_vertex_shader": (i32, [vp, ctypes.c_char_p, vp, vp, i32, vp, ctypes.c_size_t, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


EXPORTED_SYMBOLS = ["crb_abi_version", "crb_create", "crb_destroy", "crb_last_error", "crb_set_surfaces", "crb_deferred_clear", "crb_pack_abgr",
                    "crb_encode_clear_depth", "crb_set_pixel_pipe", "crb_set_pixel_pipe_by_name", "crb_set_vertex_buffer", "crb_set_index_buffer",
                    "crb_set_subviewport", "crb_draw_triangles", "crb_draw_triangles_async", "crb_finish", "crb_draw_triangles_host", "crb_draw_triangles_host_async", "crb_get_stats", "crb_get_counters",
                    "crb_set_stage_timing", "crb_get_stage_timing", "crb_get_stage_timing_frames", "crb_draw_batch_async", "crb_batch_join", "crb_compute_chunk_bounds", "crb_set_chunk_bounds", "crb_split_frame", "crb_get_profiling_info", "crb_get_launch_count", "crb_get_work_buffers", "crb_download",
                    "crb_set_binning_mode", "crb_get_last_frame_direct", "crb_set_color_layout", "crb_set_color_pitch", "crb_ipc_alloc", "crb_ipc_free", "crb_ipc_open", "crb_ipc_close", "crb_ipc_signal", "crb_ipc_copy", "crb_ipc_delay", "crb_ipc_copy_2d", "crb_resolve_surface", "crb_write_ppm", "crb_launch_vertex_shader"]


def pipe_name(base, samples_log2, flags, blend="BlendReplace"):
    """Name of a precompiled pipe variant (csrc/BuiltinPipes.cu)."""
    return "PixelPipe_%s_s%d_f%d_%s" % (base, samples_log2, flags, blend)


class _ExternalMemory:
    """Device memory owned by someone else: just enough of the tensor interface for CudaRaster.setSurfaces."""

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = ptr, nbytes

    def data_ptr(self):
        return self.ptr


class CudaSurface:
    """Render target (reference: CudaSurface.hpp:36-78): LINEAR device memory, U32 texels,
    rows = rounded height, row pitch = rounded width * numSamples, MSAA samples of an 8x8 tile
    stored as numSamples horizontally adjacent 8x8 blocks."""
    FORMAT_RGBA8 = 0
    FORMAT_DEPTH32 = 1

    def __init__(self, size, fmt, num_samples=1, device="cuda:0"):
        import torch
        w, h = int(size[0]), int(size[1])
        if min(w, h) <= 0:
            raise CrbError("CudaSurface: Size must be positive!")
        if max(w, h) > 2048:
            raise CrbError("CudaSurface: CR_MAXVIEWPORT_SIZE exceeded!")
        if fmt not in (0, 1):
            raise CrbError("CudaSurface: Invalid format!")
        if num_samples > 8:
            raise CrbError("CudaSurface: numSamples cannot exceed 8!")
        if num_samples < 1 or (num_samples & (num_samples - 1)):
            raise CrbError("CudaSurface: numSamples must be a power of two!")
        self.size = (w, h)
        self.rounded_size = ((w + 7) & ~7, (h + 7) & ~7)
        self.texture_size = (self.rounded_size[0] * num_samples, self.rounded_size[1])
        self.format = fmt
        self.num_samples = num_samples
        self.tensor = torch.zeros((self.texture_size[1], self.texture_size[0]), dtype=torch.int32, device=device)

    @classmethod
    def from_pointer(cls, ptr, size, fmt, num_samples=1):
        """A surface over device memory the caller owns (e.g. a slot of a peer GPU's frame buffer mapped with CUDA IPC,
        multigpu.PeerFrameSink): same layout, nothing is allocated."""
        self = cls.__new__(cls)
        w, h = int(size[0]), int(size[1])
        self.size = (w, h)
        self.rounded_size = ((w + 7) & ~7, (h + 7) & ~7)
        self.texture_size = (self.rounded_size[0] * num_samples, self.rounded_size[1])
        self.format, self.num_samples = fmt, num_samples
        self.tensor = _ExternalMemory(int(ptr), self.texture_size[0] * self.texture_size[1] * 4)
        return self

    def getSize(self):
        return self.size

    def getRoundedSize(self):
        return self.rounded_size

    def getTextureSize(self):
        return self.texture_size

    def getFormat(self):
        return self.format

    def getNumSamples(self):
        return self.num_samples

    def getSamplesLog2(self):
        return self.num_samples.bit_length() - 1

    def numpy(self):
        return self.tensor.cpu().numpy().view(np.uint32)

    def resolve(self, flip_y=False, stream=None):
        """MSAA resolve into a linear [height, width] int32 CUDA tensor (reference: CudaSurface::resolveToScreen,
        CudaSurface.hpp:73): box filter over the samples of every pixel (crb_resolve_surface)."""
        import torch
        lib = load_library()
        out = torch.empty((self.size[1], self.size[0]), dtype=torch.int32, device=self.tensor.device)
        s = torch.cuda.current_stream(self.tensor.device).cuda_stream if stream is None else stream
        rc = lib.crb_resolve_surface(self.tensor.data_ptr(), self.size[0], self.size[1], self.num_samples, out.data_ptr(), self.size[0], 1 if flip_y else 0,
                                     ctypes.c_void_p(s))
        if rc != 0:
            raise CrbError("CudaSurface: resolve failed with status %d" % rc)
        return out

    def writePPM(self, path):
        """Resolves and writes the colour image as a binary PPM, top scanline first."""
        img = np.ascontiguousarray(self.resolve(flip_y=True).cpu().numpy().view(np.uint32))
        rc = load_library().crb_write_ppm(path.encode(), img.ctypes.data, self.size[0], self.size[1], self.size[0])
        if rc != 0:
            raise CrbError("CudaSurface: cannot write %s" % path)


class CudaRaster:
    """Python mirror of FW::CudaRaster (reference: CudaRaster.hpp:42-190) over the C ABI."""

    def __init__(self, device=0):
        import torch
        if not torch.cuda.is_available():
            raise CrbError("CudaRaster: No CUDA-capable devices found!")
        self.lib = load_library()
        self.device = device
        self.torch = torch
        ctx = ctypes.c_void_p()
        rc = self.lib.crb_create(device, ctypes.byref(ctx))
        if rc != 0:
            raise CrbError("CudaRaster: crb_create failed with status %d (no CPU fallback exists)" % rc)
        self.ctx = ctx
        self._keep = {}

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.crb_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise CrbError(self.lib.crb_last_error(self.ctx).decode() or ("status %d" % rc))

    # -- reference API -----------------------------------------------------------------------
    def setSurfaces(self, color, depth):
        if color is None and depth is None:
            self._check(self.lib.crb_set_surfaces(self.ctx, None, None, 0, 0, 1))
            return
        if color is None:
            raise CrbError("CudaRaster: No color buffer specified!")
        if depth is None:
            raise CrbError("CudaRaster: No depth buffer specified!")
        if color.getFormat() != CudaSurface.FORMAT_RGBA8:
            raise CrbError("CudaRaster: Unsupported color buffer format!")
        if depth.getFormat() != CudaSurface.FORMAT_DEPTH32:
            raise CrbError("CudaRaster: Unsupported depth buffer format!")
        if color.getSize() != depth.getSize():
            raise CrbError("CudaRaster: Mismatch in size between surfaces!")
        if color.getNumSamples() != depth.getNumSamples():
            raise CrbError("CudaRaster: Mismatch in multisampling between surfaces!")
        self._keep["color"], self._keep["depth"] = color, depth
        self._check(self.lib.crb_set_surfaces(self.ctx, color.tensor.data_ptr(), depth.tensor.data_ptr(), color.size[0], color.size[1], color.num_samples))

    def deferredClear(self, color=(0.0, 0.0, 0.0, 0.0), depth=1.0):
        abgr = self.lib.crb_pack_abgr(*[float(c) for c in color])
        self._check(self.lib.crb_deferred_clear(self.ctx, abgr, self.lib.crb_encode_clear_depth(float(depth))))

    def setPixelPipe(self, module, name):
        """module: a ctypes.CDLL of a pixel-pipe shared object, or None for the built-in pipes."""
        handle = None if module is None else ctypes.c_void_p(module._handle)
        self._check(self.lib.crb_set_pixel_pipe_by_name(self.ctx, handle, name.encode()))

    def setVertexBuffer(self, buf, ofs=0):
        self._keep["vb"] = buf
        self._check(self.lib.crb_set_vertex_buffer(self.ctx, buf.data_ptr() + ofs, buf.numel() * buf.element_size() - ofs))

    def setIndexBuffer(self, buf, ofs, num_tris):
        self._keep["ib"] = buf
        self._check(self.lib.crb_set_index_buffer(self.ctx, buf.data_ptr() + ofs, int(num_tris)))

    def setColorLayout(self, tile_major):
        """False = row-major colour surface (reference layout), True = tile-major (64 contiguous texels per 8x8 tile)."""
        self._check(self.lib.crb_set_color_layout(self.ctx, 1 if tile_major else 0))

    def setColorPitch(self, pitch_texels):
        """Row pitch of the colour surface (0 = its own): the surface is a window of a larger image (crb_set_color_pitch)."""
        self._check(self.lib.crb_set_color_pitch(self.ctx, int(pitch_texels)))

    def setBinningMode(self, mode):
        """0 = general path only, 1 = automatic (default), 2 = direct tile path on every eligible frame, 3 = like 2 without the
        micro-triangle visibility buffer."""
        self._check(self.lib.crb_set_binning_mode(self.ctx, int(mode)))

    def lastFrameDirect(self):
        return bool(self.lib.crb_get_last_frame_direct(self.ctx))

    def setSubViewport(self, full_w, full_h, x0, y0):
        self._check(self.lib.crb_set_subviewport(self.ctx, full_w, full_h, x0, y0))

    def drawTriangles(self, stream=None, asynchronous=False):
        """asynchronous=True enqueues the frame without the blocking counter read-back; call finish()."""
        s = self.torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        fn = self.lib.crb_draw_triangles_async if asynchronous else self.lib.crb_draw_triangles
        self._check(fn(self.ctx, ctypes.c_void_p(s)))

    def finish(self, stream=None):
        """Synchronizes and checks the counters of all asynchronous frames (raises if one overflowed)."""
        s = self.torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        self._check(self.lib.crb_finish(self.ctx, ctypes.c_void_p(s)))

    def drawTrianglesHost(self, h_verts, h_idx, num_tris, h_color, h_depth=None, stream=None):
        """Host-buffer entry: pinned torch CPU tensors in, surfaces out (crb_draw_triangles_host)."""
        s = self.torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        self._check(self.lib.crb_draw_triangles_host(
            self.ctx, h_verts.data_ptr(), h_verts.numel() * h_verts.element_size(), h_idx.data_ptr(), int(num_tris), h_color.data_ptr(),
            None if h_depth is None else h_depth.data_ptr(), ctypes.c_void_p(s)))

    def drawTrianglesHostAsync(self, h_verts, h_idx, num_tris, h_color, h_depth=None, stream=None):
        """Pipelined host-buffer entry (crb_draw_triangles_host_async): returns at once; call finish()."""
        s = self.torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        self._check(self.lib.crb_draw_triangles_host_async(
            self.ctx, h_verts.data_ptr(), h_verts.numel() * h_verts.element_size(), h_idx.data_ptr(), int(num_tris), h_color.data_ptr(),
            None if h_depth is None else h_depth.data_ptr(), ctypes.c_void_p(s)))

    def getStats(self):
        out = (ctypes.c_float * 4)()
        self._check(self.lib.crb_get_stats(self.ctx, ctypes.byref(out)))
        return {"setupTime": out[0], "binTime": out[1], "coarseTime": out[2], "fineTime": out[3]}

    def setStageTiming(self, enable):
        """Asynchronous frames also record the five stage events (off by default: it splits the kernel chain)."""
        self._check(self.lib.crb_set_stage_timing(self.ctx, 1 if enable else 0))

    def getStageTiming(self):
        """Mean ms per stage over the asynchronous frames finished since setStageTiming(True)."""
        out, n = (ctypes.c_double * 4)(), ctypes.c_int(0)
        self._check(self.lib.crb_get_stage_timing(self.ctx, ctypes.byref(out), ctypes.byref(n)))
        return {"triangleSetup": out[0], "binRaster": out[1], "coarseRaster": out[2], "fineRaster": out[3], "frames": n.value}

    def getStageTimingFrames(self, max_frames=4096):
        """[frames][5] float32 array: the four stage intervals (ms) and the duration of the composite copy (0 without one) of every
        asynchronous frame finished since setStageTiming(True)."""
        out = np.zeros((max_frames, 5), np.float32)
        n = self.lib.crb_get_stage_timing_frames(self.ctx, out.ctypes.data, max_frames)
        return out[:n]

    def makeBatch(self, frames):
        """frames: list of dicts {color, depth (CudaSurface), vb, ib (tensors), num_tris, clear=((r,g,b,a), depth) or None} ->
        a crb_batch_frame array for drawBatch (the tensors / surfaces must outlive it)."""
        arr = (BatchFrame * len(frames))()
        for b, fr in zip(arr, frames):
            color, depth = fr.get("color"), fr.get("depth")
            if color is not None:
                b.color, b.depth = color.tensor.data_ptr(), depth.tensor.data_ptr()
                b.width, b.height, b.numSamples = color.size[0], color.size[1], color.num_samples
            vb, ib = fr.get("vb"), fr.get("ib")
            if vb is not None:
                b.vertices, b.vertexBytes = vb.data_ptr(), vb.numel() * vb.element_size()
            if ib is not None:
                b.indices, b.numTris = ib.data_ptr(), int(fr["num_tris"])
            clear = fr.get("clear")
            if clear is not None:
                b.clear = 1
                b.clearColor = self.lib.crb_pack_abgr(*[float(c) for c in clear[0]])
                b.clearDepth = self.lib.crb_encode_clear_depth(float(clear[1]))
            if fr.get("push_dst"):      # composite: copy the finished colour surface into a peer frame slot (side stream)
                b.pushDst, b.pushBytes, b.surfaceSlot = int(fr["push_dst"]), int(fr["push_bytes"]), int(fr.get("slot", 0))
            if fr.get("signal_word"):   # frame mark for the consumer
                b.signalWord, b.signalValue = int(fr["signal_word"]), int(fr["signal_value"])
        return arr

    def drawBatch(self, batch, stream=None):
        """crb_draw_batch_async: every frame of `batch` (makeBatch) enqueued by ONE C call; call finish()."""
        s = self.torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        self._check(self.lib.crb_draw_batch_async(self.ctx, ctypes.cast(batch, ctypes.c_void_p), len(batch), ctypes.c_void_p(s)))

    def computeChunkBounds(self, vb, ib, num_tris, vertex_stride, stream=None):
        """Per-chunk clip-space bounds of a mesh for the sort-first geometry cull (crb_compute_chunk_bounds): a float32 CUDA tensor
        [ceil(num_tris / 256), 4]; make it once per mesh, pass it to setChunkBounds after setVertexBuffer / setIndexBuffer."""
        out = self.torch.empty(((int(num_tris) + 255) // 256, 4), dtype=self.torch.float32, device=vb.device)
        s = self.torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        rc = self.lib.crb_compute_chunk_bounds(vb.data_ptr(), int(vertex_stride), ib.data_ptr(), int(num_tris), out.data_ptr(), ctypes.c_void_p(s))
        if rc != 0:
            raise CrbError("CudaRaster: crb_compute_chunk_bounds failed with status %d" % rc)
        return out

    def setChunkBounds(self, bounds):
        self._keep["bounds"] = bounds
        self._check(self.lib.crb_set_chunk_bounds(self.ctx, None if bounds is None else bounds.data_ptr()))

    def batchJoin(self, stream=None):
        """Makes the stream wait for the composite copies of drawBatch (crb_batch_join)."""
        s = self.torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        self._check(self.lib.crb_batch_join(self.ctx, ctypes.c_void_p(s)))

    def getProfilingInfo(self):
        buf = ctypes.create_string_buffer(8192)
        self._check(self.lib.crb_get_profiling_info(self.ctx, buf, 8192))
        return buf.value.decode()

    # -- extras ------------------------------------------------------------------------------
    def launchVertexShader(self, module, name, in_vertices, out_vertices, num_vertices, constants, stream=None):
        """The demo's vertex-shader launch (test/SceneCR.cpp:263-282): `name`_launch of a pixel-pipe module (None = built in);
        constants = bytes-like block passed by value (the reference's c_constants)."""
        s = self.torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        handle = None if module is None else ctypes.c_void_p(module._handle)
        blob = bytes(constants)
        rc = self.lib.crb_launch_vertex_shader(handle, name.encode(), in_vertices.data_ptr(), out_vertices.data_ptr(), int(num_vertices), blob, len(blob), ctypes.c_void_p(s))
        if rc != 0:
            raise CrbError("CudaRaster: vertex shader %s failed with status %d" % (name, rc))

    def getCounters(self):
        a = Atomics()
        self._check(self.lib.crb_get_counters(self.ctx, ctypes.byref(a)))
        return {k: getattr(a, k) for k, _ in Atomics._fields_}

    def getLaunchCount(self):
        return self.lib.crb_get_launch_count(self.ctx)

    def _dev_to_numpy(self, ptr, nbytes, dtype):
        out = np.empty(nbytes, np.uint8)
        self._check(self.lib.crb_download(self.ctx, ptr, out.ctypes.data, nbytes))
        return out.view(dtype)

    def getWorkBuffers(self, num_tris):
        """Downloads setup output and the bin/tile queues (parity tests)."""
        wb = WorkBuffers()
        self._check(self.lib.crb_get_work_buffers(self.ctx, ctypes.byref(wb)))
        c = self.getCounters()
        n_sub = min(c["numSubtris"], wb.maxSubtris)
        res = {"counters": c}
        res["triSubtris"] = self._dev_to_numpy(wb.triSubtris, max(num_tris, 1), np.uint8)[:num_tris]
        res["triHeader"] = self._dev_to_numpy(wb.triHeader, max(n_sub, 1) * 16, np.uint32).reshape(-1, 4)[:n_sub]
        res["triData"] = self._dev_to_numpy(wb.triData, max(n_sub, 1) * 64, np.uint32).reshape(-1, 16)[:n_sub]
        res["binStart"] = self._dev_to_numpy(wb.binStart, 256 * 4, np.int32)[:wb.numBins]
        res["binTotal"] = self._dev_to_numpy(wb.binTotal, 256 * 4, np.int32)[:wb.numBins]
        res["binQueue"] = self._dev_to_numpy(wb.binQueue, max(c["numBinEntries"], 1) * 4, np.int32)[:c["numBinEntries"]]
        res["tileStart"] = self._dev_to_numpy(wb.tileStart, wb.numTiles * 4, np.int32)
        res["tileCount"] = self._dev_to_numpy(wb.tileCount, wb.numTiles * 4, np.int32)
        res["tileQueue"] = self._dev_to_numpy(wb.tileQueue, max(c["numTileEntries"], 1) * 4, np.int32)[:c["numTileEntries"]]
        res["activeTiles"] = self._dev_to_numpy(wb.activeTiles, max(c["numActiveTiles"], 1) * 4, np.int32)[:c["numActiveTiles"]]
        return res
