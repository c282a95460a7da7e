// FW::CudaRaster / FW::CudaSurface / FW::Buffer / FW::CudaModule -- the reference's HOST API kept as
// a thin, header-only C++ layer over the C ABI of libcrb200.so (include/crb200.h).
//
// A caller of the reference (test/SceneCR.cpp:67-90, :149-195, :243-317) keeps its code:
//
//     FW::CudaRaster cr;  cr.init();
//     FW::CudaSurface color(FW::Vec2i(w, h), FW::CudaSurface::FORMAT_RGBA8), depth(..., FORMAT_DEPTH32);
//     FW::CudaModule  pipes("libmypipes.so");              // was: CudaCompiler::compile() -> cubin
//     cr.setSurfaces(&color, &depth);
//     cr.setPixelPipe(&pipes, "PixelPipe_passthrough");    // finds <name>_triangleSetup ... by name
//     cr.deferredClear(FW::Vec4f(0.2f, 0.4f, 0.8f, 1.0f));
//     cr.setVertexBuffer(&vb, 0);  cr.setIndexBuffer(&ib, 0, numTris);
//     cr.drawTriangles();
//     FW::CudaRaster::Stats st = cr.getStats();            // seconds per stage
//
// Same names, argument meaning and error behaviour (fail(): print + exit, base/Defs.hpp:107-117) as
//   src/cudaraster/CudaRaster.hpp:42-190, src/cudaraster/CudaSurface.hpp:36-78,
//   src/framework/gpu/Buffer.hpp:39-158 (the subset CudaRaster touches), gpu/CudaModule.hpp:41-178.
// Differences, all forced by dropping the GL / driver-API plumbing:
//   * CudaSurface owns LINEAR device memory (same tile-replicated MSAA layout) instead of a GL
//     texture registered as a CUarray; getCudaPtr() replaces getCudaArray(), download() reads it back.
//   * CudaModule wraps a dlopen() handle of a pixel-pipe shared object built with nvcc from
//     <cudaraster/cuda/PixelPipe.inl> + CR_DEFINE_PIXEL_PIPE; NULL / "" = pipes built into libcrb200.so.
//   * DebugParams is accepted and ignored: the product has no host emulation path (the CPU
//     restatement lives in oracle/ and is test infrastructure only).
// Additions: setSubViewport() (sort-first windows), drawTrianglesAsync()/finish(), stream selection, setBinningMode(),
// setColorLayout()/setColorPitch()/setSurfacePointers() (multi-GPU composites over peer memory), CudaSurface::resolve*().
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <string>

#include "../crb200.h"
#include "base/Defs.hpp"
#include "base/Math.hpp"

namespace FW {

// Device buffer with the slice of the reference's Buffer interface that CudaRaster and its callers
// use (gpu/Buffer.hpp:87-119): resizeDiscard / set / getSize / getCudaPtr.
class Buffer {
public:
    Buffer(void) : m_ptr(NULL), m_size(0) {}
    Buffer(const void* ptr, S64 size) : m_ptr(NULL), m_size(0) { set(ptr, size); }
    ~Buffer(void) { if (m_ptr) cudaFree(m_ptr); }

    S64 getSize(void) const { return m_size; }
    void resizeDiscard(S64 size) {
        if (size == m_size) return;
        if (m_ptr) cudaFree(m_ptr);
        m_ptr = NULL;
        m_size = 0;
        if (size > 0 && cudaMalloc(&m_ptr, (size_t)size) != cudaSuccess) fail("Buffer: cudaMalloc(%lld) failed!", (long long)size);
        m_size = size;
    }
    void set(const void* ptr, S64 size) {
        resizeDiscard(size);
        if (size > 0 && cudaMemcpy(m_ptr, ptr, (size_t)size, cudaMemcpyHostToDevice) != cudaSuccess) fail("Buffer: cudaMemcpy failed!");
    }
    void getRange(void* dst, S64 ofs, S64 size) const {
        if (size > 0 && cudaMemcpy(dst, (const char*)m_ptr + ofs, (size_t)size, cudaMemcpyDeviceToHost) != cudaSuccess) fail("Buffer: cudaMemcpy failed!");
    }
    void* getCudaPtr(S64 ofs = 0) { return (char*)m_ptr + ofs; }

private:
    Buffer(const Buffer&);             // forbidden
    Buffer& operator=(const Buffer&);  // forbidden
    void* m_ptr;
    S64 m_size;
};

// Render target (CudaSurface.hpp:36-78, CudaSurface.cpp:39-110): same size rounding, formats,
// sample limits and messages; the storage is linear device memory of getTextureSize() U32 texels.
class CudaSurface {
public:
    enum Format { FORMAT_RGBA8 = 0, FORMAT_DEPTH32, NUM_FORMAT };

    CudaSurface(const Vec2i& size, Format format, int numSamples = 1) : m_size(size), m_format(format), m_numSamples(numSamples), m_ptr(NULL) {
        if (size.x <= 0 || size.y <= 0) fail("CudaSurface: Size must be positive!");
        if (size.x > CRB_MAX_VIEWPORT || size.y > CRB_MAX_VIEWPORT) fail("CudaSurface: CR_MAXVIEWPORT_SIZE exceeded!");
        if (format < 0 || format >= NUM_FORMAT) fail("CudaSurface: Invalid format!");
        if (numSamples > CRB_MAX_SAMPLES) fail("CudaSurface: numSamples cannot exceed 8!");
        if (numSamples < 1 || (numSamples & (numSamples - 1)) != 0) fail("CudaSurface: numSamples must be a power of two!");
        m_roundedSize = Vec2i((size.x + CRB_TILE_SIZE - 1) & -CRB_TILE_SIZE, (size.y + CRB_TILE_SIZE - 1) & -CRB_TILE_SIZE);
        m_textureSize = Vec2i(m_roundedSize.x * numSamples, m_roundedSize.y);
        if (cudaMalloc(&m_ptr, getSizeBytes()) != cudaSuccess) fail("CudaSurface: cudaMalloc failed!");
        cudaMemset(m_ptr, 0, getSizeBytes());
    }
    ~CudaSurface(void) { if (m_ptr) cudaFree(m_ptr); }

    const Vec2i& getSize(void) const { return m_size; }                // original size
    const Vec2i& getRoundedSize(void) const { return m_roundedSize; }  // rounded to full 8x8 tiles
    const Vec2i& getTextureSize(void) const { return m_textureSize; }  // 8x8 tiles replicated horizontally for MSAA
    Format getFormat(void) const { return m_format; }
    int getNumSamples(void) const { return m_numSamples; }
    int getSamplesLog2(void) const { int l = 0; while ((1 << l) < m_numSamples) l++; return l; }

    void* getCudaPtr(void) { return m_ptr; }  // replaces getCudaArray(): [textureSize.y][textureSize.x] U32, row 0 = bottom scanline
    size_t getSizeBytes(void) const { return (size_t)m_textureSize.x * (size_t)m_textureSize.y * sizeof(U32); }
    void download(U32* dst) const {
        if (cudaMemcpy(dst, m_ptr, getSizeBytes(), cudaMemcpyDeviceToHost) != cudaSuccess) fail("CudaSurface: cudaMemcpy failed!");
    }
    // Stand-ins for resolveToScreen (CudaSurface.hpp:73: "Resolves MSAA and writes pixels into the current GL render
    // target"): box-filter resolve into linear device memory of getSize().x * getSize().y U32, or straight to a PPM file.
    void resolve(void* d_dst, bool flipY = false, cudaStream_t stream = NULL) const {
        if (crb_resolve_surface(m_ptr, m_size.x, m_size.y, m_numSamples, d_dst, m_size.x, flipY ? 1 : 0, stream) != CRB_OK) fail("CudaSurface: resolve failed!");
    }
    void resolveToFile(const std::string& ppmPath) const {
        void* d = NULL;
        const size_t bytes = (size_t)m_size.x * (size_t)m_size.y * sizeof(U32);
        if (cudaMalloc(&d, bytes) != cudaSuccess) fail("CudaSurface: cudaMalloc failed!");
        resolve(d, true);
        U32* h = (U32*)malloc(bytes);
        if (!h || cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) fail("CudaSurface: cudaMemcpy failed!");
        cudaFree(d);
        const int rc = crb_write_ppm(ppmPath.c_str(), h, m_size.x, m_size.y, m_size.x);
        free(h);
        if (rc != CRB_OK) fail("CudaSurface: cannot write '%s'!", ppmPath.c_str());
    }

private:
    CudaSurface(const CudaSurface&);             // forbidden
    CudaSurface& operator=(const CudaSurface&);  // forbidden
    Vec2i m_size, m_roundedSize, m_textureSize;
    Format m_format;
    S32 m_numSamples;
    void* m_ptr;
};

// Pixel-pipe module: a shared object that contains CR_DEFINE_PIXEL_PIPE instantiations.  Takes the
// place of the cubin CudaModule that CudaCompiler produced (gpu/CudaModule.cpp:54-58).
class CudaModule {
public:
    explicit CudaModule(const std::string& sharedObject = "") : m_handle(NULL) {
        if (!sharedObject.empty()) {
            m_handle = dlopen(sharedObject.c_str(), RTLD_NOW | RTLD_GLOBAL);
            if (!m_handle) fail("CudaModule: cannot load '%s': %s", sharedObject.c_str(), dlerror());
        }
    }
    ~CudaModule(void) { if (m_handle) dlclose(m_handle); }
    void* getHandle(void) const { return m_handle; }  // NULL = the pipes built into libcrb200.so

    // The demo's vertex-shader launch (test/SceneCR.cpp:263-282: getGlobal("c_constants") + setParam* + launchKernel):
    // runs `<name>_launch` (CR_DEFINE_VERTEX_SHADER) over numVertices vertices; the constants block travels by value.
    template <class Constants>
    void launchVertexShader(const std::string& name, Buffer& inVertices, Buffer& outVertices, int numVertices, const Constants& constants, cudaStream_t stream = NULL) {
        if (crb_launch_vertex_shader(m_handle, name.c_str(), inVertices.getCudaPtr(), outVertices.getCudaPtr(), numVertices, &constants, sizeof(Constants), stream) != CRB_OK)
            fail("CudaModule: vertex shader '%s' not found or failed to launch!", name.c_str());
    }

private:
    CudaModule(const CudaModule&);             // forbidden
    CudaModule& operator=(const CudaModule&);  // forbidden
    void* m_handle;
};

class CudaRaster {
public:
    struct Stats {       // statistics for the previous call to drawTriangles(), in SECONDS
        F32 setupTime;   // TriangleSetup
        F32 binTime;     // BinRaster
        F32 coarseTime;  // CoarseRaster
        F32 fineTime;    // FineRaster
    };
    struct DebugParams {  // accepted for source compatibility; there is no host emulation in the product
        bool emulateTriangleSetup, emulateBinRaster, emulateCoarseRaster, emulateFineRaster;
        DebugParams(void) : emulateTriangleSetup(false), emulateBinRaster(false), emulateCoarseRaster(false), emulateFineRaster(false) {}
    };

    explicit CudaRaster(int device = 0) : m_ctx(NULL), m_device(device), m_stream(NULL) {}
    ~CudaRaster(void) { if (m_ctx) crb_destroy(m_ctx); }

    void init(void) {  // CudaRaster.cpp:98-128
        if (m_ctx) return;
        const int rc = crb_create(m_device, &m_ctx);
        if (rc == CRB_ERR_NO_DEVICE) fail("CudaRaster: No CUDA-capable (sm_100) devices found! The B200 pipeline has no CPU path.");
        if (rc != CRB_OK) fail("CudaRaster: crb_create failed (%d)!", rc);
    }

    void setSurfaces(CudaSurface* color, CudaSurface* depth) {  // CudaRaster.cpp:132-170, same checks in the same order
        init();
        if (!color && !depth) { check(crb_set_surfaces(m_ctx, NULL, NULL, 0, 0, 1)); return; }
        if (!color) fail("CudaRaster: No color buffer specified!");
        if (!depth) fail("CudaRaster: No depth buffer specified!");
        if (color->getFormat() != CudaSurface::FORMAT_RGBA8) fail("CudaRaster: Unsupported color buffer format!");
        if (depth->getFormat() != CudaSurface::FORMAT_DEPTH32) fail("CudaRaster: Unsupported depth buffer format!");
        if (color->getSize().x != depth->getSize().x || color->getSize().y != depth->getSize().y) fail("CudaRaster: Mismatch in size between surfaces!");
        if (color->getNumSamples() != depth->getNumSamples()) fail("CudaRaster: Mismatch in multisampling between surfaces!");
        check(crb_set_surfaces(m_ctx, color->getCudaPtr(), depth->getCudaPtr(), color->getSize().x, color->getSize().y, color->getNumSamples()));
    }

    void deferredClear(const Vec4f& color = Vec4f(0.0f, 0.0f, 0.0f, 0.0f), F32 depth = 1.0f) {  // CudaRaster.cpp:174-179
        init();
        check(crb_deferred_clear(m_ctx, crb_pack_abgr(color.x, color.y, color.z, color.w), crb_encode_clear_depth(depth)));
    }

    void setPixelPipe(CudaModule* module, const std::string& name) {  // CudaRaster.cpp:183-216
        init();
        check(crb_set_pixel_pipe_by_name(m_ctx, module ? module->getHandle() : NULL, name.c_str()));
    }

    void setVertexBuffer(Buffer* buf, S64 ofs) {  // CudaRaster.cpp:220-224
        init();
        check(crb_set_vertex_buffer(m_ctx, buf ? buf->getCudaPtr(ofs) : NULL, buf ? (size_t)(buf->getSize() - ofs) : 0));
    }
    void setIndexBuffer(Buffer* buf, S64 ofs, int numTris) {  // CudaRaster.cpp:226-233
        init();
        check(crb_set_index_buffer(m_ctx, buf ? buf->getCudaPtr(ofs) : NULL, numTris));
    }

    void drawTriangles(void) { init(); check(crb_draw_triangles(m_ctx, m_stream)); }  // CudaRaster.cpp:237-342

    Stats getStats(void) {  // CudaRaster.cpp:346-363
        init();
        float s[4];
        check(crb_get_stats(m_ctx, s));
        Stats st;
        st.setupTime = s[0]; st.binTime = s[1]; st.coarseTime = s[2]; st.fineTime = s[3];
        return st;
    }
    std::string getProfilingInfo(void) {  // CudaRaster.cpp:367-497 (ProfilingMode_Default report)
        init();
        char buf[8192];
        check(crb_get_profiling_info(m_ctx, buf, sizeof(buf)));
        return buf;
    }
    void setDebugParams(const DebugParams&) {}

    // ---- additions ------------------------------------------------------------------------------
    void setStream(cudaStream_t stream) { m_stream = stream; }
    void setSubViewport(int fullWidth, int fullHeight, int x0, int y0) { init(); check(crb_set_subviewport(m_ctx, fullWidth, fullHeight, x0, y0)); }
    void drawTrianglesAsync(void) { init(); check(crb_draw_triangles_async(m_ctx, m_stream)); }
    bool finish(void) {  // false = some asynchronous frame overflowed a work buffer: redraw it
        init();
        const int rc = crb_finish(m_ctx, m_stream);
        if (rc == CRB_ERR_OVERFLOW) return false;
        check(rc);
        return true;
    }
    crb_atomics getCounters(void) { init(); crb_atomics a; check(crb_get_counters(m_ctx, &a)); return a; }
    // binning strategy (crb200.h): 0 = ordered two-level sort only, 1 = automatic, 2 = direct tile path whenever the pipe allows, 3 = 2 without micro-triangles
    void setBinningMode(int mode) { init(); check(crb_set_binning_mode(m_ctx, mode)); }
    bool lastFrameDirect(void) { init(); return crb_get_last_frame_direct(m_ctx) != 0; }
    // colour surface addressing for multi-GPU composites: tile-major layout (peer-memory frame slots), row pitch of a larger image (sort-first windows in place)
    void setColorLayout(bool tileMajor) { init(); check(crb_set_color_layout(m_ctx, tileMajor ? 1 : 0)); }
    void setColorPitch(int pitchTexels) { init(); check(crb_set_color_pitch(m_ctx, pitchTexels)); }
    // surfaces over memory the caller owns (e.g. a frame slot of another GPU mapped with crb_ipc_open); the reference's checks are the caller's business here
    void setSurfacePointers(void* d_color, void* d_depth, const Vec2i& size, int numSamples = 1) { init(); check(crb_set_surfaces(m_ctx, d_color, d_depth, size.x, size.y, numSamples)); }
    // sort-first geometry cull (crb200.h): per-chunk clip-space bounds of the CURRENT vertex / index buffers, or NULL
    void setChunkBounds(const float* d_bounds) { init(); check(crb_set_chunk_bounds(m_ctx, d_bounds)); }
    crb_ctx* getContext(void) { init(); return m_ctx; }

private:
    CudaRaster(const CudaRaster&);             // forbidden
    CudaRaster& operator=(const CudaRaster&);  // forbidden
    void check(int rc) { if (rc != CRB_OK) fail("%s", crb_last_error(m_ctx)); }

    crb_ctx* m_ctx;
    int m_device;
    cudaStream_t m_stream;
};

}  // namespace FW
