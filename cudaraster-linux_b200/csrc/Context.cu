// Host driver behind the C ABI (include/crb200.h).  B200-native replacement for the reference's
// CudaRaster host class (src/cudaraster/CudaRaster.cpp:53-363, :508-665): state setters, work-buffer
// sizing with the same slack / overflow-retry policy, the stage launches bracketed by five CUDA
// events, counter read-back.  Differences: runtime API on an explicit stream, frame parameters as
// kernel arguments, linear-memory surfaces, CSR queues instead of segment pools, optional
// asynchronous operation (the reference blocks on the atomics read-back every frame).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cudaraster/cuda/PrivateDefs.hpp"
#include "../../include/cudaraster/cuda/Util.cuh"

extern "C" int crb_ipc_signal(void* d_word, uint32_t value, void* stream);
extern "C" int crb_bin_launches(const crb_frame* f);
extern "C" int crb_coarse_launches(const crb_frame* f);
extern "C" int crb_launch_direct_alloc(const crb_frame* f, void* stream);
extern "C" int crb_launch_direct_scatter(const crb_frame* f, void* stream);

namespace {

struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    // Grow-only "resizeDiscard" (reference: gpu/Buffer.cpp resizeDiscard): contents are not kept.
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

}  // namespace

struct crb_ctx {
    int device = 0;
    int numSMs = 1;
    std::string err;

    // state (CudaRaster.hpp:69-108)
    uint32_t* color = nullptr;
    uint32_t* depth = nullptr;
    int width = 0, height = 0, numSamples = 1, samplesLog2 = 0;
    int fullWidth = 0, fullHeight = 0, subX0 = 0, subY0 = 0;
    bool deferredClear = false;
    bool colorTiled = false;             // crb_set_color_layout
    int colorPitch = 0;                  // crb_set_color_pitch (0 = the surface's own pitch)
    uint32_t clearColor = 0, clearDepth = 0;
    const void* vertices = nullptr;
    size_t vertexBytes = 0;
    const int32_t* indices = nullptr;
    int numTris = 0;
    const float* chunkBounds = nullptr;  // crb_set_chunk_bounds: belongs to the current vertex / index buffers
    bool verticesSet = false, indicesSet = false;  // an EMPTY buffer is still a buffer (CudaRaster.cpp:220-233)
    bool hasPipe = false;
    crb_pipe_desc pipe{};
    crb_pipe_spec spec{};
    std::string pipeName;

    // work buffers
    int maxSubtris = 1, maxBinEntries = 1, maxTileEntries = 1, maxItems = 1, maxLarge = 16384;
    DevBuf triSubtris, triHeader, triData;
    DevBuf binCountMat, binStart, binTotal, binQueue;
    DevBuf items, binItemBase, binItemCount, tileCountMat;
    DevBuf tileQueue, tileStart, tileCount, activeTiles, activeRecs;
    DevBuf tileCounter;                  // direct tile path: per-tile counters, zero between frames
    DevBuf profCounters;                 // ProfilingMode_Counters: CRB_PROF_NUM numerator / denominator pairs
    unsigned long long hostProf[CRB_PROF_WORDS] = {};
    DevBuf visBuffer;                    // micro-triangle visibility buffer (8 B / pixel), all ones between frames
    size_t visBytes = 0;                 // extent the current surface uses
    DevBuf tileCursor;                   // direct tile path: per-tile queue cursors (alloc -> scatter)
    DevBuf triTileCode;                  // direct tile path: one word per input triangle (setup -> scatter)
    DevBuf largeList;                    // direct tile path: {entry, slot} of the large sub-triangles (setup -> alloc, scatter)
    DevBuf batchQueued;                  // direct tile path: one byte per batch of 32 triangles (did setup leave their words in triTileCode?)
    // Binning strategy (crb_set_binning_mode): the direct path runs whenever the pipe is order independent.
    int binningMode = 1;                 // 0 never, 1 automatic (= 2), 2 on every eligible frame
    struct PendingFrame {                // what crb_finish needs to know about an asynchronous frame
        int numTris = 0;
        cudaStream_t stream = nullptr;
        bool hadClear = false;           // the frame consumed a deferred clear (restored if it must be redrawn)
        uint32_t clearColor = 0, clearDepth = 0;
    } pendingFrame[64];
    bool lastFrameDirect = false;
    bool microOff = false;               // crb_set_binning_mode(3)
    bool microEnabled = true;            // CRB_MICRO=0 keeps small triangles on the tile queues (A/B measurements)
    DevBuf atomics;
    crb_atomics* hostAtomics = nullptr;  // pinned + mapped; slot 0 = synchronous draws, slots 1.. = ring of asynchronous frames
    crb_atomics* hostAtomicsDev = nullptr;   // the same memory as the device sees it: the fine raster kernel stores the frame's counters there itself
    int counterSlot = 0;                 // slot the next frame reports into
    int pending = 0;                     // asynchronous frames not yet checked by crb_finish
    DevBuf hostVerts, hostIdx;           // device staging for crb_draw_triangles_host

    // Between frames the count matrices are all zero and the counter block of the NEXT frame is zeroed by
    // the fine raster kernel (double-buffered), so a steady-state frame enqueues kernels only.  needReset
    // forces the memsets: first frame, reallocation, layout change, overflow, any failed launch.
    bool needReset = true;
    int atomicsParity = 0;
    int lastNumBins = -1, lastMatPitch = -1, lastNumChunks = -1, lastCtasPerChunk = -1;
    int debugFlags = 0;
    bool chainLaunches = true;           // programmatic dependent launch between the kernels of a frame (CRB_NO_PDL=1 turns it off)
    bool stageTiming = false;            // asynchronous frames: record the five stage events too (crb_set_stage_timing)
    static constexpr int kTimingRing = 64;
    cudaEvent_t ringEv[kTimingRing][7] = {};   // five stage events + start / end of the frame's composite copy (side stream)
    bool ringComp[kTimingRing] = {};
    double stageSumMs[4] = {0, 0, 0, 0};
    int stageFrames = 0;
    std::vector<float> stageFrameMs;     // 5 intervals per finished frame since crb_set_stage_timing: four stages + composite (crb_get_stage_timing_frames)

    // crb_draw_triangles_host_async: upload / render / download of consecutive frames overlap
    struct HostPipeline {
        bool init = false;
        cudaStream_t up = nullptr, up2 = nullptr, down = nullptr;   // vertices and indices upload concurrently (two DMA queues in flight)
        cudaEvent_t uploaded2[2] = {};
        DevBuf verts[2], idx[2];
        cudaEvent_t uploaded[2] = {}, rendered[2] = {}, downloaded = nullptr;
        long long frames = 0;
    } hp;

    // crb_draw_batch_async: composite copies (frames pushed into a peer GPU's frame slots) on a side stream
    struct Composite {
        bool init = false, any = false;
        cudaStream_t side = nullptr;
        cudaEvent_t rendered[4] = {}, pushed[4] = {}, last = nullptr;
        bool pushedValid[4] = {false, false, false, false};
    } comp;

    cudaEvent_t ev[5] = {};
    crb_frame frame{};
    crb_atomics lastAtomics{};
    int launchCount = 0;
    bool drawn = false;
    bool evRecorded = false;             // ev[] hold a synchronous frame
};

namespace {

int setError(crb_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

int cudaFail(crb_ctx* c, const char* what, cudaError_t e) { return setError(c, CRB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e)); }

#define CRB_CUDA(ctx, call)                                        \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cudaFail((ctx), #call, e_);  \
    } while (0)

constexpr int kAsyncRing = 64;

int popc8(int v) { return __builtin_popcount((unsigned)v & 0xFF); }

// The direct tile path serves every frame of an order-independent pipe: triangles of any size are exact there (large ones are
// counted and scattered by whole CTAs, row spans found by bisection -- Overlap.cuh forEachCellStrided), so the decision needs
// no knowledge of the scene: the first frame of a shape and frames with a changing triangle count take it as well.
bool wantDirect(const crb_ctx* c) {
    if (c->binningMode == 0 || !c->hasPipe || !c->pipe.orderIndependent) return false;
    if (c->spec.profilingMode != ProfilingMode_Default) return false;   // the counters describe the ordered two-level path
    return true;
}

// Fills the frame block and (re)allocates the work buffers for the current capacities.
int prepareFrame(crb_ctx* c) {
    crb_frame& f = c->frame;
    std::memset(&f, 0, sizeof(f));
    f.numTris = c->numTris;
    f.vertexStride = c->spec.vertexStructSize;
    f.vertexBuffer = c->vertices;
    f.indexBuffer = c->indices;

    f.viewportWidth = c->width;
    f.viewportHeight = c->height;
    f.widthPixels = (c->width + CR_TILE_SIZE - 1) & -CR_TILE_SIZE;
    f.heightPixels = (c->height + CR_TILE_SIZE - 1) & -CR_TILE_SIZE;
    f.widthTiles = f.widthPixels >> CR_TILE_LOG2;
    f.heightTiles = f.heightPixels >> CR_TILE_LOG2;
    f.numTiles = f.widthTiles * f.heightTiles;
    f.widthBins = (f.widthTiles + CR_BIN_SIZE - 1) >> CR_BIN_LOG2;
    f.heightBins = (f.heightTiles + CR_BIN_SIZE - 1) >> CR_BIN_LOG2;
    f.numBins = f.widthBins * f.heightBins;
    f.samplesLog2 = c->samplesLog2;

    if (c->fullWidth > 0) {
        // Sort-first window.  Triangles are set up in a PARENT viewport of at most
        // 2048^2 px (the whole frame when it fits; otherwise the cell of an even grid over the frame
        // that holds the window) exactly as an unsplit render of that viewport would set them up --
        // snapped once in the subpixel grid of the full frame -- and the surface is a scissor
        // rectangle inside it.  All windows of one parent therefore produce identical pixels.
        const int ncx = (c->fullWidth + CR_MAXVIEWPORT_SIZE - 1) / CR_MAXVIEWPORT_SIZE, ncy = (c->fullHeight + CR_MAXVIEWPORT_SIZE - 1) / CR_MAXVIEWPORT_SIZE;
        const int cellW = (((c->fullWidth + ncx - 1) / ncx) + 7) & ~7, cellH = (((c->fullHeight + ncy - 1) / ncy) + 7) & ~7;
        const int px0 = (c->subX0 / cellW) * cellW, py0 = (c->subY0 / cellH) * cellH;
        const int pw = std::min(cellW, c->fullWidth - px0), ph = std::min(cellH, c->fullHeight - py0);
        if (c->subX0 + c->width > px0 + pw || c->subY0 + c->height > py0 + ph)
            return setError(c, CRB_ERR_INVALID, "CudaRaster: a sort-first window must lie inside one %dx%d cell of the frame!", cellW, cellH);
        f.windowed = 1;
        f.chunkBounds = (const float4*)c->chunkBounds;
        f.fullWidth = c->fullWidth;
        f.fullHeight = c->fullHeight;
        f.viewportWidth = pw;
        f.viewportHeight = ph;
        f.centerOfsX = px0 * CR_SUBPIXEL_SIZE + pw * (CR_SUBPIXEL_SIZE / 2) - c->fullWidth * (CR_SUBPIXEL_SIZE / 2);
        f.centerOfsY = py0 * CR_SUBPIXEL_SIZE + ph * (CR_SUBPIXEL_SIZE / 2) - c->fullHeight * (CR_SUBPIXEL_SIZE / 2);
        f.subX0 = c->subX0 - px0;
        f.subY0 = c->subY0 - py0;
        f.clipLoX = (float)(2.0 * px0 / c->fullWidth - 1.0);
        f.clipHiX = (float)(2.0 * (px0 + pw) / c->fullWidth - 1.0);
        f.clipLoY = (float)(2.0 * py0 / c->fullHeight - 1.0);
        f.clipHiY = (float)(2.0 * (py0 + ph) / c->fullHeight - 1.0);
        f.cullLoX = (float)(2.0 * c->subX0 / c->fullWidth - 1.0);
        f.cullHiX = (float)(2.0 * (c->subX0 + c->width) / c->fullWidth - 1.0);
        f.cullLoY = (float)(2.0 * c->subY0 / c->fullHeight - 1.0);
        f.cullHiY = (float)(2.0 * (c->subY0 + c->height) / c->fullHeight - 1.0);
    } else {
        f.windowed = 0;
        f.fullWidth = c->width;
        f.fullHeight = c->height;
        f.centerOfsX = f.centerOfsY = 0;
        f.subX0 = f.subY0 = 0;
        f.clipLoX = f.clipLoY = f.cullLoX = f.cullLoY = -1.0f;
        f.clipHiX = f.clipHiY = f.cullHiX = f.cullHiY = 1.0f;
    }
    f.originX = f.viewportWidth * (CR_SUBPIXEL_SIZE / 2) - f.subX0 * CR_SUBPIXEL_SIZE;
    f.originY = f.viewportHeight * (CR_SUBPIXEL_SIZE / 2) - f.subY0 * CR_SUBPIXEL_SIZE;

    f.deferredClear = c->deferredClear ? 1 : 0;
    f.clearColor = c->clearColor;
    f.clearDepth = c->clearDepth;
    f.colorBuffer = c->color;
    f.depthBuffer = c->depth;
    f.surfacePitch = f.widthPixels << c->samplesLog2;
    if (c->colorTiled && c->samplesLog2 != 0) return setError(c, CRB_ERR_INVALID, "CudaRaster: the tile-major colour layout is single-sample only!");
    f.colorTiled = c->colorTiled ? 1 : 0;
    if (c->colorPitch != 0 && (c->samplesLog2 != 0 || c->colorTiled || c->colorPitch < f.widthPixels))
        return setError(c, CRB_ERR_INVALID, "CudaRaster: a colour pitch needs a single-sample row-major surface and must cover the rounded width!");
    f.colorPitch = c->colorPitch != 0 ? c->colorPitch : f.surfacePitch;

    f.ctasPerChunk = std::max(1, CRB_MIN_CHUNK_TRIS / CRB_SETUP_THREADS);
    while (((int64_t)c->numTris + CRB_SETUP_THREADS * f.ctasPerChunk - 1) / (CRB_SETUP_THREADS * f.ctasPerChunk) > CRB_MAX_CHUNKS) f.ctasPerChunk *= 2;
    f.chunkTris = CRB_SETUP_THREADS * f.ctasPerChunk;
    f.numChunks = (c->numTris + f.chunkTris - 1) / f.chunkTris;
    f.matPitch = (f.numChunks + 3) & ~3;
    f.maxSubtris = c->maxSubtris;
    f.maxBinEntries = c->maxBinEntries;
    f.maxTileEntries = c->maxTileEntries;
    f.maxItems = c->maxItems;
    f.directMode = wantDirect(c) ? 1 : 0;
    f.microMode = (f.directMode && c->samplesLog2 == 0 && c->microEnabled && !c->microOff) ? 1 : 0;
    f.numSMs = c->numSMs;
    f.chainLaunches = c->chainLaunches ? 1 : 0;
    f.debugFlags = c->debugFlags;

    const void* oldBinMat = c->binCountMat.ptr;
    const void* oldTileMat = c->tileCountMat.ptr;

    CRB_CUDA(c, c->triSubtris.reserve((size_t)c->maxSubtris));
    CRB_CUDA(c, c->triHeader.reserve((size_t)c->maxSubtris * 16));
    CRB_CUDA(c, c->triData.reserve((size_t)c->maxSubtris * 64));
    CRB_CUDA(c, c->binCountMat.reserve((size_t)std::max(4, f.numBins * f.matPitch) * 4));
    CRB_CUDA(c, c->binStart.reserve(CR_MAXBINS_SQR * 4));
    CRB_CUDA(c, c->binTotal.reserve(CR_MAXBINS_SQR * 4));
    CRB_CUDA(c, c->binItemBase.reserve(CR_MAXBINS_SQR * 4));
    CRB_CUDA(c, c->binItemCount.reserve(CR_MAXBINS_SQR * 4));
    CRB_CUDA(c, c->binQueue.reserve((size_t)c->maxBinEntries * 4));
    CRB_CUDA(c, c->items.reserve((size_t)c->maxItems * sizeof(crb_item)));
    CRB_CUDA(c, c->tileCountMat.reserve((size_t)c->maxItems * CR_BIN_SQR * 4));
    CRB_CUDA(c, c->tileQueue.reserve((size_t)c->maxTileEntries * 4));
    CRB_CUDA(c, c->tileStart.reserve(CR_MAXTILES_SQR * 4));
    CRB_CUDA(c, c->tileCount.reserve(CR_MAXTILES_SQR * 4));
    CRB_CUDA(c, c->activeTiles.reserve(CR_MAXTILES_SQR * 4));
    CRB_CUDA(c, c->activeRecs.reserve(CR_MAXTILES_SQR * 16));
    const void* oldTileCounter = c->tileCounter.ptr;
    CRB_CUDA(c, c->tileCounter.reserve(CR_MAXTILES_SQR * 4));
    if (oldTileCounter != c->tileCounter.ptr) c->needReset = true;
    CRB_CUDA(c, c->tileCursor.reserve(CR_MAXTILES_SQR * 4));
    CRB_CUDA(c, c->profCounters.reserve(CRB_PROF_WORDS * sizeof(unsigned long long)));
    f.profCounters = (unsigned long long*)c->profCounters.ptr;
    f.profilingMode = c->spec.profilingMode;
    if (f.microMode) {
        const void* oldVis = c->visBuffer.ptr;
        const size_t need = (size_t)f.widthPixels * f.heightPixels * 8;
        CRB_CUDA(c, c->visBuffer.reserve(need));
        if (oldVis != c->visBuffer.ptr) c->needReset = true;
        c->visBytes = std::max(c->visBytes, need);
    }
    if (f.directMode) CRB_CUDA(c, c->triTileCode.reserve(((size_t)std::max(c->numTris, 1) + 4) * 4));
    if (f.directMode) CRB_CUDA(c, c->largeList.reserve((size_t)c->maxLarge * 8));
    f.maxLarge = c->maxLarge;
    if (f.directMode) CRB_CUDA(c, c->batchQueued.reserve((size_t)std::max(c->numTris, 1) / 32 + 16));

    f.triSubtris = (uint8_t*)c->triSubtris.ptr;
    f.triHeader = (uint4*)c->triHeader.ptr;
    f.triData = (uint4*)c->triData.ptr;
    f.binCountMat = (int32_t*)c->binCountMat.ptr;
    f.binStart = (int32_t*)c->binStart.ptr;
    f.binTotal = (int32_t*)c->binTotal.ptr;
    f.binQueue = (int32_t*)c->binQueue.ptr;
    f.items = (crb_item*)c->items.ptr;
    f.binItemBase = (int32_t*)c->binItemBase.ptr;
    f.binItemCount = (int32_t*)c->binItemCount.ptr;
    f.tileCountMat = (int32_t*)c->tileCountMat.ptr;
    f.tileQueue = (int32_t*)c->tileQueue.ptr;
    f.tileStart = (int32_t*)c->tileStart.ptr;
    f.tileCount = (int32_t*)c->tileCount.ptr;
    f.activeTiles = (int32_t*)c->activeTiles.ptr;
    f.activeRecs = (int4*)c->activeRecs.ptr;
    f.tileCounter = (int32_t*)c->tileCounter.ptr;
    f.tileCursor = (int32_t*)c->tileCursor.ptr;
    f.visBuffer = (unsigned long long*)c->visBuffer.ptr;
    f.triTileCode = (uint32_t*)c->triTileCode.ptr;
    f.batchQueued = (uint8_t*)c->batchQueued.ptr;
    f.largeList = (int2*)c->largeList.ptr;
    f.atomics = (crb_atomics*)c->atomics.ptr + c->atomicsParity;
    f.nextAtomics = (crb_atomics*)c->atomics.ptr + (c->atomicsParity ^ 1);
    f.hostCounters = c->hostAtomicsDev + c->counterSlot;
    if (oldBinMat != c->binCountMat.ptr || oldTileMat != c->tileCountMat.ptr || c->lastNumBins != f.numBins || c->lastMatPitch != f.matPitch ||
        c->lastNumChunks != f.numChunks || c->lastCtasPerChunk != f.ctasPerChunk)
        c->needReset = true;
    c->lastNumBins = f.numBins; c->lastMatPitch = f.matPitch; c->lastNumChunks = f.numChunks; c->lastCtasPerChunk = f.ctasPerChunk;
    return CRB_OK;
}

// Enqueues one frame (reference: CudaRaster::launchStages, CudaRaster.cpp:508-665).  ev = the five
// stage events to record (nullptr: none -- asynchronous frames run as an unbroken chain of kernels).
int launchStages(crb_ctx* c, cudaStream_t s, cudaEvent_t* ev) {
    const crb_frame* f = &c->frame;
    if (c->needReset) {
        CRB_CUDA(c, cudaMemsetAsync(c->atomics.ptr, 0, 2 * sizeof(crb_atomics), s));
        CRB_CUDA(c, cudaMemsetAsync(c->binCountMat.ptr, 0, c->binCountMat.cap, s));
        CRB_CUDA(c, cudaMemsetAsync(c->tileCountMat.ptr, 0, c->tileCountMat.cap, s));
        CRB_CUDA(c, cudaMemsetAsync(c->tileCounter.ptr, 0, c->tileCounter.cap, s));
        if (c->visBuffer.ptr) CRB_CUDA(c, cudaMemsetAsync(c->visBuffer.ptr, 0xFF, c->visBuffer.cap, s));
        c->needReset = false;
    }
    // several setup CTAs per chunk ADD their bin counts into one column: that (large-scene) layout needs a zeroed matrix.
    // (Letting the bin scatter zero the cells it reads was measured 3x slower than this memset: 466 vs 169 us on C4.)
    if (f->numTris > 0 && f->ctasPerChunk > 1 && !f->directMode) CRB_CUDA(c, cudaMemsetAsync(c->binCountMat.ptr, 0, (size_t)f->matPitch * f->numBins * 4, s));
    if (c->spec.profilingMode != ProfilingMode_Default) CRB_CUDA(c, cudaMemsetAsync(c->profCounters.ptr, 0, CRB_PROF_WORDS * sizeof(unsigned long long), s));
    c->needReset = true;   // until every launch of this frame went through
    if (ev) CRB_CUDA(c, cudaEventRecord(ev[0], s));
    int rc = c->pipe.triangleSetup(f, s);
    if (rc != CRB_OK) return setError(c, rc, "CudaRaster: triangleSetup launch failed (%s)", cudaGetErrorString(cudaGetLastError()));
    if (ev) CRB_CUDA(c, cudaEventRecord(ev[1], s));
    // direct tile path: queue allocation stands where the bin stage stands, the unordered scatter where the coarse stage does
    rc = f->directMode ? crb_launch_direct_alloc(f, s) : c->pipe.binRaster(f, s);
    if (rc != CRB_OK) return setError(c, rc, "CudaRaster: binRaster launch failed (%s)", cudaGetErrorString(cudaGetLastError()));
    if (ev) CRB_CUDA(c, cudaEventRecord(ev[2], s));
    rc = f->directMode ? crb_launch_direct_scatter(f, s) : c->pipe.coarseRaster(f, s);
    if (rc != CRB_OK) return setError(c, rc, "CudaRaster: coarseRaster launch failed (%s)", cudaGetErrorString(cudaGetLastError()));
    if (ev) CRB_CUDA(c, cudaEventRecord(ev[3], s));
    rc = c->pipe.fineRaster(f, s);
    if (rc != CRB_OK) return setError(c, rc, "CudaRaster: fineRaster launch failed (%s)", cudaGetErrorString(cudaGetLastError()));
    if (ev) CRB_CUDA(c, cudaEventRecord(ev[4], s));
    if (ev == c->ev) c->evRecorded = true;
    c->launchCount += (f->numTris > 0 ? 1 : 0) + crb_bin_launches(f) + crb_coarse_launches(f) + 1;
    c->needReset = false;
    c->atomicsParity ^= 1;   // the fine raster kernel zeroed the other block for the next frame
    c->lastFrameDirect = f->directMode != 0;
    return CRB_OK;
}

int validateDraw(crb_ctx* c) {
    // same checks, same order, same messages as CudaRaster::drawTriangles (CudaRaster.cpp:243-260)
    if (!c->color) return setError(c, CRB_ERR_INVALID, "CudaRaster: Surfaces not set!");
    if (!c->hasPipe) return setError(c, CRB_ERR_INVALID, "CudaRaster: Pixel pipe not set!");
    if (!c->verticesSet || (!c->vertices && c->numTris > 0)) return setError(c, CRB_ERR_INVALID, "CudaRaster: Vertex buffer not set!");
    if (!c->indicesSet || (!c->indices && c->numTris > 0)) return setError(c, CRB_ERR_INVALID, "CudaRaster: Index buffer not set!");
    if (c->spec.samplesLog2 != c->samplesLog2) return setError(c, CRB_ERR_INVALID, "CudaRaster: Mismatch in multisampling between pixel pipe and surface!");
    return CRB_OK;
}

}  // namespace

extern "C" {

int crb_abi_version(void) { return CRB_ABI_VERSION; }

int crb_create(int device, crb_ctx** out) {
    if (!out) return CRB_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return CRB_ERR_NO_DEVICE;  // "CudaRaster: No CUDA-capable devices found!" -- there is no CPU path
    }
    if (device < 0 || device >= count) return CRB_ERR_INVALID;
    crb_ctx* c = new crb_ctx();
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete c; return CRB_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return CRB_ERR_CUDA; }
    if (prop.major < 10) {  // sm_100a code only
        delete c;
        return CRB_ERR_NO_DEVICE;
    }
    c->numSMs = prop.multiProcessorCount;
    for (int i = 0; i < 5; i++)
        if (cudaEventCreate(&c->ev[i]) != cudaSuccess) { delete c; return CRB_ERR_CUDA; }
    for (int i = 0; i < crb_ctx::kTimingRing; i++)
        for (int k = 0; k < 7; k++)
            if (cudaEventCreate(&c->ringEv[i][k]) != cudaSuccess) { delete c; return CRB_ERR_CUDA; }
    if (c->atomics.reserve(2 * sizeof(crb_atomics)) != cudaSuccess) { delete c; return CRB_ERR_CUDA; }
    const char* noPdl = getenv("CRB_NO_PDL");
    c->chainLaunches = !(noPdl && noPdl[0] == '1');
    const char* direct = getenv("CRB_DIRECT");   // 0 / 1 / 2, see crb_set_binning_mode
    if (direct && direct[0] >= '0' && direct[0] <= '2') c->binningMode = direct[0] - '0';
    const char* micro = getenv("CRB_MICRO");
    if (micro && micro[0] == '0') c->microEnabled = false;
    const char* dbg = getenv("CRB_DEBUG_FLAGS");
    c->debugFlags = dbg ? atoi(dbg) : 0;
    if (cudaHostAlloc((void**)&c->hostAtomics, sizeof(crb_atomics) * (1 + kAsyncRing), cudaHostAllocMapped) != cudaSuccess) { delete c; return CRB_ERR_CUDA; }
    if (cudaHostGetDevicePointer((void**)&c->hostAtomicsDev, c->hostAtomics, 0) != cudaSuccess) { delete c; return CRB_ERR_CUDA; }
    std::memset(c->hostAtomics, 0, sizeof(crb_atomics) * (1 + kAsyncRing));
    *out = c;
    return CRB_OK;
}

int crb_destroy(crb_ctx* c) {
    if (!c) return CRB_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    DevBuf* bufs[] = {&c->triSubtris, &c->triHeader, &c->triData, &c->binCountMat, &c->binStart, &c->binTotal, &c->binQueue, &c->items, &c->binItemBase,
                      &c->binItemCount, &c->tileCountMat, &c->tileQueue, &c->tileStart, &c->tileCount, &c->activeTiles, &c->activeRecs, &c->atomics, &c->hostVerts, &c->hostIdx, &c->tileCounter, &c->tileCursor, &c->triTileCode, &c->batchQueued, &c->largeList, &c->visBuffer, &c->profCounters};
    for (DevBuf* b : bufs) b->release();
    if (c->hp.init) {
        cudaStreamDestroy(c->hp.up);
        cudaStreamDestroy(c->hp.up2);
        cudaStreamDestroy(c->hp.down);
        for (int i = 0; i < 2; i++) {
            cudaEventDestroy(c->hp.uploaded2[i]);
            cudaEventDestroy(c->hp.uploaded[i]);
            cudaEventDestroy(c->hp.rendered[i]);
            c->hp.verts[i].release();
            c->hp.idx[i].release();
        }
        cudaEventDestroy(c->hp.downloaded);
    }
    if (c->comp.init) {
        cudaStreamDestroy(c->comp.side);
        for (int k = 0; k < 4; k++) { cudaEventDestroy(c->comp.rendered[k]); cudaEventDestroy(c->comp.pushed[k]); }
        cudaEventDestroy(c->comp.last);
    }
    if (c->hostAtomics) cudaFreeHost(c->hostAtomics);
    for (int i = 0; i < 5; i++)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < crb_ctx::kTimingRing; i++)
        for (int k = 0; k < 7; k++)
            if (c->ringEv[i][k]) cudaEventDestroy(c->ringEv[i][k]);
    delete c;
    return CRB_OK;
}

const char* crb_last_error(const crb_ctx* c) { return c ? c->err.c_str() : "null context"; }

int crb_set_surfaces(crb_ctx* c, void* d_color, void* d_depth, int width, int height, int numSamples) {
    if (!c) return CRB_ERR_INVALID;
    if (!d_color && !d_depth) {
        c->color = c->depth = nullptr;
        return CRB_OK;
    }
    // CudaRaster::setSurfaces / CudaSurface::CudaSurface checks and messages.  Nothing is committed before every check passed: a
    // failed call leaves the context WITHOUT surfaces (a later draw fails with "Surfaces not set!") instead of new pointers
    // with the old dimensions.
    c->color = c->depth = nullptr;
    if (!d_color) return setError(c, CRB_ERR_INVALID, "CudaRaster: No color buffer specified!");
    if (!d_depth) return setError(c, CRB_ERR_INVALID, "CudaRaster: No depth buffer specified!");
    if (std::min(width, height) <= 0) return setError(c, CRB_ERR_INVALID, "CudaSurface: Size must be positive!");
    if (std::max(width, height) > CR_MAXVIEWPORT_SIZE) return setError(c, CRB_ERR_LIMIT, "CudaSurface: CR_MAXVIEWPORT_SIZE exceeded!");
    if (numSamples > 8) return setError(c, CRB_ERR_LIMIT, "CudaSurface: numSamples cannot exceed 8!");
    if (numSamples < 1 || popc8(numSamples) != 1) return setError(c, CRB_ERR_INVALID, "CudaSurface: numSamples must be a power of two!");
    c->color = (uint32_t*)d_color;
    c->depth = (uint32_t*)d_depth;
    c->width = width;
    c->height = height;
    c->numSamples = numSamples;
    c->samplesLog2 = popc8(numSamples - 1);
    return CRB_OK;
}

int crb_deferred_clear(crb_ctx* c, uint32_t abgr, uint32_t encodedDepth) {
    if (!c) return CRB_ERR_INVALID;
    c->deferredClear = true;
    c->clearColor = abgr;
    c->clearDepth = encodedDepth;
    return CRB_OK;
}

uint32_t crb_pack_abgr(float r, float g, float b, float a) { return FW::Vec4f(r, g, b, a).toABGR(); }

uint32_t crb_encode_clear_depth(float depth) {
    double d = (double)depth * 4294967296.0;
    unsigned long long q = d <= 0.0 ? 0ull : (d >= 18446744073709551615.0 ? ~0ull : (unsigned long long)d);
    return FW::encodeDepth((uint32_t)std::min<unsigned long long>(q, 0xFFFFFFFFull));
}

int crb_set_pixel_pipe(crb_ctx* c, const crb_pipe_desc* pipe) {
    if (!c) return CRB_ERR_INVALID;
    c->hasPipe = false;
    if (!pipe) return CRB_OK;
    if (!pipe->spec || !pipe->triangleSetup || !pipe->binRaster || !pipe->coarseRaster || !pipe->fineRaster)
        return setError(c, CRB_ERR_INVALID, "CudaRaster: Invalid pixel pipe!");
    c->pipe = *pipe;
    c->spec = *pipe->spec;
    c->pipe.spec = &c->spec;
    c->pipeName = pipe->name ? pipe->name : "";
    c->pipe.name = c->pipeName.c_str();
    c->hasPipe = true;
    return CRB_OK;
}

int crb_set_pixel_pipe_by_name(crb_ctx* c, void* module, const char* name) {
    if (!c || !name) return CRB_ERR_INVALID;
    void* handle = module;
    if (!handle) {
        // the pipes compiled into this library
        Dl_info info;
        if (dladdr((const void*)&crb_abi_version, &info) && info.dli_fname) handle = dlopen(info.dli_fname, RTLD_NOW | RTLD_NOLOAD);
        if (!handle) handle = dlopen(nullptr, RTLD_NOW);
    }
    const std::string n(name);
    crb_pipe_desc d{};
    d.name = name;
    d.spec = (const crb_pipe_spec*)dlsym(handle, (n + "_spec").c_str());
    d.triangleSetup = (crb_stage_fn)dlsym(handle, (n + "_triangleSetup").c_str());
    d.binRaster = (crb_stage_fn)dlsym(handle, (n + "_binRaster").c_str());
    d.coarseRaster = (crb_stage_fn)dlsym(handle, (n + "_coarseRaster").c_str());
    d.fineRaster = (crb_stage_fn)dlsym(handle, (n + "_fineRaster").c_str());
    typedef int (*ProbeFn)(void);
    ProbeFn frameBytes = (ProbeFn)dlsym(handle, (n + "_frameBytes").c_str());
    if (frameBytes && frameBytes() != (int)sizeof(crb_frame)) {
        c->hasPipe = false;
        return setError(c, CRB_ERR_INVALID, "CudaRaster: pixel pipe module '%s' was built against other headers than this library (frame block %d vs %d bytes): rebuild it!", name,
                        frameBytes(), (int)sizeof(crb_frame));
    }
    ProbeFn probe = (ProbeFn)dlsym(handle, (n + "_orderIndependent").c_str());   // optional (pipes built before ABI 2 lack it)
    d.orderIndependent = 0;
    if (probe) {
        cudaSetDevice(c->device);
        d.orderIndependent = probe();
    }
    if (!d.spec || !d.triangleSetup || !d.binRaster || !d.coarseRaster || !d.fineRaster) {
        c->hasPipe = false;
        return setError(c, CRB_ERR_INVALID, "CudaRaster: Invalid pixel pipe!");
    }
    return crb_set_pixel_pipe(c, &d);
}

int crb_set_vertex_buffer(crb_ctx* c, const void* d_vertices, size_t bytes) {
    if (!c) return CRB_ERR_INVALID;
    if (c->vertices != d_vertices) c->chunkBounds = nullptr;
    c->vertices = d_vertices;
    c->vertexBytes = bytes;
    c->verticesSet = d_vertices != nullptr || bytes == 0;
    return CRB_OK;
}

int crb_set_index_buffer(crb_ctx* c, const void* d_indices, int numTris) {
    if (!c) return CRB_ERR_INVALID;
    if (numTris < 0) return setError(c, CRB_ERR_INVALID, "CudaRaster: negative triangle count!");
    if (c->indices != (const int32_t*)d_indices || c->numTris != numTris) c->chunkBounds = nullptr;
    c->indices = (const int32_t*)d_indices;
    c->numTris = numTris;
    c->indicesSet = d_indices != nullptr || numTris == 0;
    return CRB_OK;
}

int crb_split_frame(int fullWidth, int fullHeight, int parts, int* outRects, int maxRects) {
    // the parent cells of prepareFrame() first, then every cell into the same power-of-two grid (columns before rows)
    if (fullWidth <= 0 || fullHeight <= 0 || parts < 1 || (maxRects > 0 && !outRects)) return -1;
    const int ncx = (fullWidth + CR_MAXVIEWPORT_SIZE - 1) / CR_MAXVIEWPORT_SIZE, ncy = (fullHeight + CR_MAXVIEWPORT_SIZE - 1) / CR_MAXVIEWPORT_SIZE;
    const int cellW = (((fullWidth + ncx - 1) / ncx) + 7) & ~7, cellH = (((fullHeight + ncy - 1) / ncy) + 7) & ~7;
    int sub = 1, lg = 0;
    while (ncx * ncy * sub < parts) { sub *= 2; lg++; }
    const int cols = 1 << ((lg + 1) / 2), rows = 1 << (lg / 2);
    struct R { int x, y, w, h; };
    std::vector<R> rects;
    for (int cy = 0; cy < ncy; cy++)
        for (int cx = 0; cx < ncx; cx++) {
            const int px0 = cx * cellW, py0 = cy * cellH, pw = std::min(cellW, fullWidth - px0), ph = std::min(cellH, fullHeight - py0);
            const int sw = (((pw + cols - 1) / cols) + 7) & ~7, sh = (((ph + rows - 1) / rows) + 7) & ~7;
            for (int r = 0; r < rows; r++)
                for (int q = 0; q < cols; q++) {
                    const int x0 = px0 + q * sw, y0 = py0 + r * sh, w = std::min(sw, px0 + pw - x0), h = std::min(sh, py0 + ph - y0);
                    if (w > 0 && h > 0) rects.push_back({x0, y0, w, h});
                }
        }
    std::sort(rects.begin(), rects.end(), [](const R& a, const R& b) { return a.y != b.y ? a.y < b.y : a.x < b.x; });
    for (int i = 0; i < (int)rects.size() && i < maxRects; i++) {
        outRects[4 * i + 0] = rects[i].x; outRects[4 * i + 1] = rects[i].y; outRects[4 * i + 2] = rects[i].w; outRects[4 * i + 3] = rects[i].h;
    }
    return (int)rects.size();
}

int crb_set_chunk_bounds(crb_ctx* c, const float* d_bounds) {
    if (!c) return CRB_ERR_INVALID;
    static_assert(CRB_CHUNK_BOUNDS_TRIS == CRB_SETUP_THREADS, "one bounds record per setup CTA");
    c->chunkBounds = d_bounds;
    return CRB_OK;
}

int crb_set_color_pitch(crb_ctx* c, int pitchTexels) {
    if (!c || pitchTexels < 0) return CRB_ERR_INVALID;
    c->colorPitch = pitchTexels;
    return CRB_OK;
}

int crb_set_color_layout(crb_ctx* c, int tileMajor) {
    if (!c) return CRB_ERR_INVALID;
    c->colorTiled = tileMajor != 0;
    return CRB_OK;
}

int crb_set_binning_mode(crb_ctx* c, int mode) {
    if (!c || mode < 0 || mode > 3) return CRB_ERR_INVALID;
    c->binningMode = mode == 3 ? 2 : mode;
    c->microOff = mode == 3;
    return CRB_OK;
}

int crb_get_last_frame_direct(crb_ctx* c) { return c && c->lastFrameDirect ? 1 : 0; }

int crb_set_subviewport(crb_ctx* c, int fullWidth, int fullHeight, int x0, int y0) {
    if (!c) return CRB_ERR_INVALID;
    if (fullWidth <= 0) {
        c->fullWidth = c->fullHeight = c->subX0 = c->subY0 = 0;
        return CRB_OK;
    }
    if (fullHeight <= 0 || x0 < 0 || y0 < 0 || (x0 & 7) || (y0 & 7)) return setError(c, CRB_ERR_INVALID, "CudaRaster: sub-viewport origin must be a non-negative multiple of 8!");
    c->fullWidth = fullWidth;
    c->fullHeight = fullHeight;
    c->subX0 = x0;
    c->subY0 = y0;
    return CRB_OK;
}

int crb_draw_triangles(crb_ctx* c, void* stream) {
    if (!c) return CRB_ERR_INVALID;
    int rc = validateDraw(c);
    if (rc != CRB_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CUDA(c, cudaSetDevice(c->device));
    if (c->fullWidth > 0 && (c->subX0 + c->width > c->fullWidth || c->subY0 + c->height > c->fullHeight))
        return setError(c, CRB_ERR_INVALID, "CudaRaster: sub-viewport exceeds the full frame!");

    // Initial capacities: the reference's slack for sub-triangles (CudaRaster.cpp:239-241, :264-275);
    // queue capacities are estimates that the retry loop corrects once per scene.
    const int numTris = c->numTris;
    const int numTilesEst = (((c->width + 7) >> 3) * ((c->height + 7) >> 3));
    c->maxSubtris = std::max(c->maxSubtris, numTris + 4096);
    c->maxBinEntries = std::max(c->maxBinEntries, numTris + numTris / 4 + 16384);
    c->maxTileEntries = std::max(c->maxTileEntries, std::max(numTilesEst, numTris * 2) + 65536);
    c->launchCount = 0;

    for (int attempt = 0;; attempt++) {
        if (c->maxSubtris > CR_MAXSUBTRIS_SIZE) return setError(c, CRB_ERR_LIMIT, "CudaRaster: CR_MAXSUBTRIS_SIZE exceeded!");
        c->maxItems = c->maxBinEntries / CRB_ITEM_ENTRIES + CR_MAXBINS_SQR + 1;
        c->counterSlot = 0;
        rc = prepareFrame(c);
        if (rc != CRB_OK) return rc;
        rc = launchStages(c, s, c->ev);
        if (rc != CRB_OK) return rc;
        // counters back on the host (reference: CudaRaster.cpp:326 -- one blocking round trip per frame): the fine raster kernel
        // stored them into the mapped block itself, no copy is enqueued
        CRB_CUDA(c, cudaStreamSynchronize(s));
        crb_atomics a = *c->hostAtomics;
        a.numSubtris += numTris;
        c->lastAtomics = a;
        if (a.overflow == 0) break;
        c->needReset = true;   // the kernels of an overflowed frame return early and leave the count matrices dirty
        if (attempt > 8) return setError(c, CRB_ERR_LIMIT, "CudaRaster: work buffers keep overflowing (flags %d)", a.overflow);
        // grow and rerun ALL stages (CudaRaster.cpp:328-338)
        if (a.overflow & 1) c->maxSubtris = std::max(c->maxSubtris, a.numSubtris + 4096);
        if (a.overflow & (2 | 8)) c->maxBinEntries = std::max(c->maxBinEntries, a.numBinEntries + a.numBinEntries / 16 + 16384);
        if (a.overflow & 4) c->maxTileEntries = std::max(c->maxTileEntries, a.numTileEntries + a.numTileEntries / 16 + 65536);
        if (a.overflow & 32) c->maxLarge = std::max(c->maxLarge, a.numLargeTris + a.numLargeTris / 4 + 4096);
    }
    c->deferredClear = false;
    c->drawn = true;
    return CRB_OK;
}

int crb_finish(crb_ctx* c, void* stream) {
    if (!c) return CRB_ERR_INVALID;
    CRB_CUDA(c, cudaSetDevice(c->device));
    CRB_CUDA(c, cudaStreamSynchronize((cudaStream_t)stream));
    // frames may have been enqueued on other streams than the one passed in: their counter copies must have landed too
    for (int i = 0; i < c->pending; i++) {
        bool seen = c->pendingFrame[i].stream == (cudaStream_t)stream;
        for (int k = 0; k < i && !seen; k++) seen = c->pendingFrame[k].stream == c->pendingFrame[i].stream;
        if (!seen) CRB_CUDA(c, cudaStreamSynchronize(c->pendingFrame[i].stream));
    }
    if (c->comp.init && c->comp.any) CRB_CUDA(c, cudaStreamSynchronize(c->comp.side));
    int overflowed = 0, firstBad = -1;
    for (int i = 0; i < c->pending; i++) {
        if (c->stageTiming) {
            for (int k = 0; k < 4; k++) {
                float ms = 0.0f;
                if (cudaEventElapsedTime(&ms, c->ringEv[i][k], c->ringEv[i][k + 1]) == cudaSuccess) c->stageSumMs[k] += ms;
                if (c->stageFrameMs.size() < (size_t)5 * 65536) c->stageFrameMs.push_back(ms);
            }
            float compMs = 0.0f;   // the frame's composite copy (crb_draw_batch_async, pushDst), 0 when it has none
            if (c->ringComp[i] && cudaEventElapsedTime(&compMs, c->ringEv[i][5], c->ringEv[i][6]) != cudaSuccess) compMs = 0.0f;
            if (c->stageFrameMs.size() < (size_t)5 * 65536) c->stageFrameMs.push_back(compMs);
            c->stageFrames++;
        }
        crb_atomics a = c->hostAtomics[1 + i];
        a.numSubtris += c->pendingFrame[i].numTris;
        if (overflowed && a.overflow != 0 && (a.overflow & 16) != 0) continue;   // a frame skipped because an EARLIER frame of the batch overflowed (sticky flag, bit 4): its counters mean nothing
        c->lastAtomics = a;
        if (a.overflow == 0) continue;
        overflowed++;
        if (firstBad < 0) firstBad = i;
        if (a.overflow & 1) c->maxSubtris = std::max(c->maxSubtris, a.numSubtris + 4096);
        if (a.overflow & (2 | 8)) c->maxBinEntries = std::max(c->maxBinEntries, a.numBinEntries + a.numBinEntries / 16 + 16384);
        if (a.overflow & 4) c->maxTileEntries = std::max(c->maxTileEntries, a.numTileEntries + a.numTileEntries / 16 + 65536);
        if (a.overflow & 32) c->maxLarge = std::max(c->maxLarge, a.numLargeTris + a.numLargeTris / 4 + 4096);
    }
    const int n = c->pending;
    c->pending = 0;
    if (overflowed) {
        c->needReset = true;
        // the frames from firstBad on must be redrawn: give the first of them its deferred clear back
        const crb_ctx::PendingFrame& pf = c->pendingFrame[firstBad];
        if (pf.hadClear) { c->deferredClear = true; c->clearColor = pf.clearColor; c->clearDepth = pf.clearDepth; }
    }
    // An overflowed frame returns early and leaves the self-cleaning scratch state (count matrices, tile counters, visibility
    // buffer) dirty, so every frame enqueued AFTER it in this batch is suspect as well: all of them must be redrawn.
    if (overflowed)
        return setError(c, CRB_ERR_OVERFLOW, "CudaRaster: %d of %d asynchronous frames overflowed a work buffer (first: frame %d of the batch); capacities grown, redraw that frame and all later ones",
                        overflowed, n, firstBad);
    return CRB_OK;
}

int crb_draw_triangles_async(crb_ctx* c, void* stream) {
    if (!c) return CRB_ERR_INVALID;
    int rc = validateDraw(c);
    if (rc != CRB_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CUDA(c, cudaSetDevice(c->device));
    if (c->fullWidth > 0 && (c->subX0 + c->width > c->fullWidth || c->subY0 + c->height > c->fullHeight))
        return setError(c, CRB_ERR_INVALID, "CudaRaster: sub-viewport exceeds the full frame!");
    if (c->pending == kAsyncRing) {
        rc = crb_finish(c, stream);
        if (rc != CRB_OK) return rc;
    }
    const int numTris = c->numTris;
    const int numTilesEst = (((c->width + 7) >> 3) * ((c->height + 7) >> 3));
    c->maxSubtris = std::max(c->maxSubtris, numTris + 4096);
    c->maxBinEntries = std::max(c->maxBinEntries, numTris + numTris / 4 + 16384);
    c->maxTileEntries = std::max(c->maxTileEntries, std::max(numTilesEst, numTris * 2) + 65536);
    if (c->maxSubtris > CR_MAXSUBTRIS_SIZE) return setError(c, CRB_ERR_LIMIT, "CudaRaster: CR_MAXSUBTRIS_SIZE exceeded!");
    c->maxItems = c->maxBinEntries / CRB_ITEM_ENTRIES + CR_MAXBINS_SQR + 1;
    c->launchCount = 0;
    c->ringComp[c->pending] = false;
    c->counterSlot = 1 + c->pending;   // the frame's fine raster kernel stores its counters there (mapped host memory): no copy between frames
    rc = prepareFrame(c);
    if (rc != CRB_OK) return rc;
    rc = launchStages(c, s, c->stageTiming ? c->ringEv[c->pending] : nullptr);
    if (rc != CRB_OK) return rc;
    crb_ctx::PendingFrame& pf = c->pendingFrame[c->pending];
    pf.numTris = numTris; pf.stream = s; pf.hadClear = c->deferredClear; pf.clearColor = c->clearColor; pf.clearDepth = c->clearDepth;
    c->pending++;
    c->deferredClear = false;
    c->drawn = true;
    return CRB_OK;
}

int crb_draw_triangles_host(crb_ctx* c, const void* h_vertices, size_t vertexBytes, const int32_t* h_indices, int numTris, uint32_t* h_color, uint32_t* h_depth,
                            void* stream) {
    if (!c) return CRB_ERR_INVALID;
    if (!c->color) return setError(c, CRB_ERR_INVALID, "CudaRaster: Surfaces not set!");
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CUDA(c, cudaSetDevice(c->device));
    CRB_CUDA(c, c->hostVerts.reserve(std::max<size_t>(vertexBytes, 16)));
    CRB_CUDA(c, c->hostIdx.reserve(std::max<size_t>((size_t)numTris * 12, 16)));
    CRB_CUDA(c, cudaMemcpyAsync(c->hostVerts.ptr, h_vertices, vertexBytes, cudaMemcpyHostToDevice, s));
    CRB_CUDA(c, cudaMemcpyAsync(c->hostIdx.ptr, h_indices, (size_t)numTris * 12, cudaMemcpyHostToDevice, s));
    c->vertices = c->hostVerts.ptr;
    c->vertexBytes = vertexBytes;
    c->indices = (const int32_t*)c->hostIdx.ptr;
    c->numTris = numTris;
    c->verticesSet = c->indicesSet = true;
    int rc = crb_draw_triangles(c, stream);
    if (rc != CRB_OK) return rc;
    const size_t surfBytes = (size_t)c->frame.surfacePitch * c->frame.heightPixels * 4;
    if (h_color) CRB_CUDA(c, cudaMemcpyAsync(h_color, c->color, surfBytes, cudaMemcpyDeviceToHost, s));
    if (h_depth) CRB_CUDA(c, cudaMemcpyAsync(h_depth, c->depth, surfBytes, cudaMemcpyDeviceToHost, s));
    CRB_CUDA(c, cudaStreamSynchronize(s));
    return CRB_OK;
}

int crb_draw_triangles_host_async(crb_ctx* c, const void* h_vertices, size_t vertexBytes, const int32_t* h_indices, int numTris, uint32_t* h_color,
                                  uint32_t* h_depth, void* stream) {
    if (!c) return CRB_ERR_INVALID;
    if (!c->color) return setError(c, CRB_ERR_INVALID, "CudaRaster: Surfaces not set!");
    cudaStream_t s = (cudaStream_t)stream;
    CRB_CUDA(c, cudaSetDevice(c->device));
    crb_ctx::HostPipeline& hp = c->hp;
    if (!hp.init) {
        CRB_CUDA(c, cudaStreamCreateWithFlags(&hp.up, cudaStreamNonBlocking));
        CRB_CUDA(c, cudaStreamCreateWithFlags(&hp.up2, cudaStreamNonBlocking));
        CRB_CUDA(c, cudaStreamCreateWithFlags(&hp.down, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CRB_CUDA(c, cudaEventCreateWithFlags(&hp.uploaded2[i], cudaEventDisableTiming));
            CRB_CUDA(c, cudaEventCreateWithFlags(&hp.uploaded[i], cudaEventDisableTiming));
            CRB_CUDA(c, cudaEventCreateWithFlags(&hp.rendered[i], cudaEventDisableTiming));
        }
        CRB_CUDA(c, cudaEventCreateWithFlags(&hp.downloaded, cudaEventDisableTiming));
        hp.init = true;
    }
    const int slot = (int)(hp.frames & 1);
    const size_t idxBytes = (size_t)numTris * 12;
    if (std::max<size_t>(vertexBytes, 16) > hp.verts[slot].cap || std::max<size_t>(idxBytes, 16) > hp.idx[slot].cap) {
        CRB_CUDA(c, cudaDeviceSynchronize());   // growing a staging buffer frees memory frames in flight may read
        CRB_CUDA(c, hp.verts[slot].reserve(std::max<size_t>(vertexBytes, 16)));
        CRB_CUDA(c, hp.idx[slot].reserve(std::max<size_t>(idxBytes, 16)));
    }
    // upload: the slot was last read by the frame before the previous one
    if (hp.frames >= 2) {
        CRB_CUDA(c, cudaStreamWaitEvent(hp.up, hp.rendered[slot], 0));
        CRB_CUDA(c, cudaStreamWaitEvent(hp.up2, hp.rendered[slot], 0));
    }
    CRB_CUDA(c, cudaMemcpyAsync(hp.verts[slot].ptr, h_vertices, vertexBytes, cudaMemcpyHostToDevice, hp.up));
    CRB_CUDA(c, cudaMemcpyAsync(hp.idx[slot].ptr, h_indices, idxBytes, cudaMemcpyHostToDevice, hp.up2));
    CRB_CUDA(c, cudaEventRecord(hp.uploaded[slot], hp.up));
    CRB_CUDA(c, cudaEventRecord(hp.uploaded2[slot], hp.up2));
    // render on the caller's stream: needs this frame's upload, and the previous frame's surfaces downloaded
    // (the previous frame's download is already ordered before it, see the end of this function)
    CRB_CUDA(c, cudaStreamWaitEvent(s, hp.uploaded[slot], 0));
    CRB_CUDA(c, cudaStreamWaitEvent(s, hp.uploaded2[slot], 0));
    c->vertices = hp.verts[slot].ptr;
    c->vertexBytes = vertexBytes;
    c->indices = (const int32_t*)hp.idx[slot].ptr;
    c->numTris = numTris;
    c->verticesSet = c->indicesSet = true;
    int rc = crb_draw_triangles_async(c, stream);
    if (rc != CRB_OK) return rc;
    CRB_CUDA(c, cudaEventRecord(hp.rendered[slot], s));
    // download
    CRB_CUDA(c, cudaStreamWaitEvent(hp.down, hp.rendered[slot], 0));
    const size_t surfBytes = (size_t)c->frame.surfacePitch * c->frame.heightPixels * 4;
    if (h_color) CRB_CUDA(c, cudaMemcpyAsync(h_color, c->color, surfBytes, cudaMemcpyDeviceToHost, hp.down));
    if (h_depth) CRB_CUDA(c, cudaMemcpyAsync(h_depth, c->depth, surfBytes, cudaMemcpyDeviceToHost, hp.down));
    CRB_CUDA(c, cudaEventRecord(hp.downloaded, hp.down));
    // later work on the caller's stream (and crb_finish) sees the downloaded frame
    CRB_CUDA(c, cudaStreamWaitEvent(s, hp.downloaded, 0));
    hp.frames++;
    return CRB_OK;
}

int crb_get_stats(crb_ctx* c, float out[4]) {
    if (!c || !out) return CRB_ERR_INVALID;
    out[0] = out[1] = out[2] = out[3] = 0.0f;
    if (!c->drawn || !c->evRecorded) return CRB_OK;
    CRB_CUDA(c, cudaEventSynchronize(c->ev[4]));
    for (int i = 0; i < 4; i++) {
        float ms = 0.0f;
        CRB_CUDA(c, cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]));
        out[i] = ms * 1.0e-3f;  // seconds, like CudaRaster::Stats
    }
    return CRB_OK;
}

int crb_set_stage_timing(crb_ctx* c, int enable) {
    if (!c) return CRB_ERR_INVALID;
    c->stageTiming = enable != 0;
    c->stageSumMs[0] = c->stageSumMs[1] = c->stageSumMs[2] = c->stageSumMs[3] = 0.0;
    c->stageFrames = 0;
    c->stageFrameMs.clear();
    return CRB_OK;
}

int crb_get_stage_timing_frames(crb_ctx* c, float* outMs, int maxFrames) {
    if (!c || (maxFrames > 0 && !outMs)) return 0;
    const int n = std::min((int)(c->stageFrameMs.size() / 5), std::max(maxFrames, 0));
    if (n > 0) std::memcpy(outMs, c->stageFrameMs.data(), (size_t)n * 5 * sizeof(float));
    return n;
}

int crb_draw_batch_async(crb_ctx* c, const crb_batch_frame* frames, int numFrames, void* stream) {
    if (!c || numFrames < 0 || (numFrames > 0 && !frames)) return CRB_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    crb_ctx::Composite& cp = c->comp;
    for (int i = 0; i < numFrames; i++) {
        const crb_batch_frame& b = frames[i];
        int rc = CRB_OK;
        const int slot = b.surfaceSlot & 3;
        if (b.pushDst) {
            CRB_CUDA(c, cudaSetDevice(c->device));
            if (!cp.init) {
                CRB_CUDA(c, cudaStreamCreateWithFlags(&cp.side, cudaStreamNonBlocking));
                for (int k = 0; k < 4; k++) {
                    CRB_CUDA(c, cudaEventCreateWithFlags(&cp.rendered[k], cudaEventDisableTiming));
                    CRB_CUDA(c, cudaEventCreateWithFlags(&cp.pushed[k], cudaEventDisableTiming));
                }
                CRB_CUDA(c, cudaEventCreateWithFlags(&cp.last, cudaEventDisableTiming));
                cp.init = true;
            }
            if (cp.pushedValid[slot]) CRB_CUDA(c, cudaStreamWaitEvent(s, cp.pushed[slot], 0));   // the copy that last read this local surface
        }
        if (b.color || b.depth) rc = crb_set_surfaces(c, b.color, b.depth, b.width, b.height, b.numSamples);
        if (rc == CRB_OK && b.vertices) rc = crb_set_vertex_buffer(c, b.vertices, b.vertexBytes);
        if (rc == CRB_OK && b.indices) rc = crb_set_index_buffer(c, b.indices, b.numTris);
        if (rc == CRB_OK && b.clear) rc = crb_deferred_clear(c, b.clearColor, b.clearDepth);
        if (rc == CRB_OK) rc = crb_draw_triangles_async(c, stream);
        if (rc != CRB_OK) return rc;
        const int ringIdx = c->pending - 1;   // the frame just enqueued
        if (c->stageTiming && ringIdx >= 0) c->ringComp[ringIdx] = b.pushDst != nullptr;
        if (b.pushDst) {
            CRB_CUDA(c, cudaEventRecord(cp.rendered[slot], s));
            CRB_CUDA(c, cudaStreamWaitEvent(cp.side, cp.rendered[slot], 0));
            if (c->stageTiming && ringIdx >= 0) CRB_CUDA(c, cudaEventRecord(c->ringEv[ringIdx][5], cp.side));
            CRB_CUDA(c, cudaMemcpyAsync(b.pushDst, c->color, b.pushBytes, cudaMemcpyDeviceToDevice, cp.side));
            if (b.signalWord && crb_ipc_signal(b.signalWord, b.signalValue, cp.side) != CRB_OK) return setError(c, CRB_ERR_CUDA, "CudaRaster: frame mark failed");
            if (c->stageTiming && ringIdx >= 0) CRB_CUDA(c, cudaEventRecord(c->ringEv[ringIdx][6], cp.side));
            CRB_CUDA(c, cudaEventRecord(cp.pushed[slot], cp.side));
            cp.pushedValid[slot] = true;
            cp.any = true;
        } else if (b.signalWord) {
            if (crb_ipc_signal(b.signalWord, b.signalValue, s) != CRB_OK) return setError(c, CRB_ERR_CUDA, "CudaRaster: frame mark failed");
        }
    }
    return CRB_OK;
}

int crb_batch_join(crb_ctx* c, void* stream) {
    if (!c) return CRB_ERR_INVALID;
    crb_ctx::Composite& cp = c->comp;
    if (!cp.init || !cp.any) return CRB_OK;
    CRB_CUDA(c, cudaSetDevice(c->device));
    CRB_CUDA(c, cudaEventRecord(cp.last, cp.side));
    CRB_CUDA(c, cudaStreamWaitEvent((cudaStream_t)stream, cp.last, 0));
    return CRB_OK;
}

int crb_get_stage_timing(crb_ctx* c, double outMeanMs[4], int* outFrames) {
    if (!c || !outMeanMs) return CRB_ERR_INVALID;
    for (int k = 0; k < 4; k++) outMeanMs[k] = c->stageFrames > 0 ? c->stageSumMs[k] / c->stageFrames : 0.0;
    if (outFrames) *outFrames = c->stageFrames;
    return CRB_OK;
}

int crb_get_counters(crb_ctx* c, crb_atomics* out) {
    if (!c || !out) return CRB_ERR_INVALID;
    *out = c->lastAtomics;
    return CRB_OK;
}

int crb_get_launch_count(crb_ctx* c) { return c ? c->launchCount : 0; }

int crb_get_profiling_info(crb_ctx* c, char* buf, size_t bufSize) {
    if (!c || !buf || bufSize == 0) return CRB_ERR_INVALID;
    std::string s("\n");
    if (!c->hasPipe) s += "Pixel pipe not set!\n";
    float st[4];
    int rc = crb_get_stats(c, st);
    if (rc != CRB_OK) return rc;
    const crb_atomics& a = c->lastAtomics;
    const float total = st[0] + st[1] + st[2] + st[3];
    const float pct = total > 0.0f ? 100.0f / total : 0.0f;
    char line[256];
    if (c->spec.profilingMode == ProfilingMode_Timers && c->hasPipe && c->drawn) {
        // ProfilingMode_Timers report (CudaRaster.cpp:452-487): each timer as a percentage of its stage's total, with the
        // reference's format strings for the regions that exist in these kernels (lane-0 clock64() brackets).
        CRB_CUDA(c, cudaDeviceSynchronize());
        CRB_CUDA(c, cudaMemcpy(c->hostProf, c->profCounters.ptr, sizeof(c->hostProf), cudaMemcpyDeviceToHost));
        auto pct = [&](int t, int parent) {
            return 100.0 * (double)c->hostProf[2 * CRB_PROF_NUM + t] / std::max((double)c->hostProf[2 * CRB_PROF_NUM + parent], 1.0);
        };
        s += "ProfilingMode_Timers\n--------------------\n\n";
        s += "TriangleSetup:\n- Compute\n";
        snprintf(line, sizeof(line), "  - Cull & snap      %4.1f%%\n", pct(CRB_TIMER_SetupCullSnap, CRB_TIMER_SetupTotal)); s += line;
        snprintf(line, sizeof(line), "  - Pleq setup       %4.1f%%\n", pct(CRB_TIMER_SetupPleq, CRB_TIMER_SetupTotal)); s += line;
        snprintf(line, sizeof(line), "  - Clip             %4.1f%%\n", pct(CRB_TIMER_SetupClip, CRB_TIMER_SetupTotal)); s += line;
        s += "- Memory\n";
        snprintf(line, sizeof(line), "  - Vertex read      %4.1f%%\n", pct(CRB_TIMER_SetupVertexRead, CRB_TIMER_SetupTotal)); s += line;
        s += "- Marshal\n";
        snprintf(line, sizeof(line), "  - Bin histogram    %4.1f%%\n\n", pct(CRB_TIMER_SetupBinning, CRB_TIMER_SetupTotal)); s += line;
        // bin / coarse stages (reference: cuda/PrivateDefs.hpp:221-248): the regions that exist in the count / scan / scatter kernels
        s += "BinRaster:\n- Compute\n";
        snprintf(line, sizeof(line), "  - Rasterize        %4.1f%%\n", pct(CRB_TIMER_BinRasterize, CRB_TIMER_BinTotal) - pct(CRB_TIMER_BinCountTiles, CRB_TIMER_BinTotal)); s += line;
        s += "- Memory\n";
        snprintf(line, sizeof(line), "  - Read tri header  %4.1f%%\n", pct(CRB_TIMER_BinReadTriHeader, CRB_TIMER_BinTotal)); s += line;
        s += "- Marshal\n";
        snprintf(line, sizeof(line), "  - Count emit       %4.1f%%\n", pct(CRB_TIMER_BinCountTiles, CRB_TIMER_BinTotal)); s += line;
        snprintf(line, sizeof(line), "  - Allocate segs    %4.1f%%\n\n", pct(CRB_TIMER_BinScan, CRB_TIMER_BinTotal)); s += line;
        s += "CoarseRaster:\n- Compute\n";
        snprintf(line, sizeof(line), "  - Rasterize        %4.1f%%\n", pct(CRB_TIMER_CoarseRasterize, CRB_TIMER_CoarseTotal)); s += line;
        s += "- Memory\n";
        snprintf(line, sizeof(line), "  - Stream read      %4.1f%%\n", pct(CRB_TIMER_CoarseStreamRead, CRB_TIMER_CoarseTotal)); s += line;
        s += "- Marshal\n";
        snprintf(line, sizeof(line), "  - Count prefsum    %4.1f%%\n\n", pct(CRB_TIMER_CoarseScan, CRB_TIMER_CoarseTotal)); s += line;
        s += "FineRaster:\n";
        snprintf(line, sizeof(line), "- Shader             %4.1f%%\n", pct(CRB_TIMER_FineShade, CRB_TIMER_FineTotal)); s += line;
        s += "- Compute\n";
        snprintf(line, sizeof(line), "  - Pixel coverage   %4.1f%%\n", pct(CRB_TIMER_FinePixelCoverage, CRB_TIMER_FineTotal)); s += line;
        snprintf(line, sizeof(line), "  - Z kill           %4.1f%%\n", pct(CRB_TIMER_FineZKill, CRB_TIMER_FineTotal)); s += line;
        s += "- Memory\n";
        snprintf(line, sizeof(line), "  - Read tile        %4.1f%%\n", pct(CRB_TIMER_FineReadTile, CRB_TIMER_FineTotal)); s += line;
        snprintf(line, sizeof(line), "  - Write tile       %4.1f%%\n", pct(CRB_TIMER_FineWriteTile, CRB_TIMER_FineTotal)); s += line;
        s += "\n";
        snprintf(buf, bufSize, "%s", s.c_str());
        return CRB_OK;
    }
    if (c->spec.profilingMode == ProfilingMode_Counters && c->hasPipe && c->drawn) {
        // ProfilingMode_Counters report (CudaRaster.cpp:424-450): the reference's format strings for the counters that exist
        // in this pipeline; bin / coarse lines come from the frame counters (their reference counters describe its own
        // round / segment / merge loops).
        CRB_CUDA(c, cudaDeviceSynchronize());
        CRB_CUDA(c, cudaMemcpy(c->hostProf, c->profCounters.ptr, sizeof(c->hostProf), cudaMemcpyDeviceToHost));
        auto ratio = [&](int k) { return (double)c->hostProf[2 * k] / std::max((double)c->hostProf[2 * k + 1], 1.0); };
        s += "ProfilingMode_Counters\n----------------------\n\n";
        s += "TriangleSetup:\n";
        snprintf(line, sizeof(line), "- Viewport cull        %.1f%%\n", ratio(CRB_PROF_SetupViewportCull)); s += line;
        snprintf(line, sizeof(line), "- Backface cull        %.1f%%\n", ratio(CRB_PROF_SetupBackfaceCull)); s += line;
        snprintf(line, sizeof(line), "- Between pixels cull  %.1f%%\n", ratio(CRB_PROF_SetupBetweenPixelsCull)); s += line;
        snprintf(line, sizeof(line), "- Clipped              %.1f%%\n", ratio(CRB_PROF_SetupClipped)); s += line;
        snprintf(line, sizeof(line), "- Avg. samples / tri   %.2f\n\n", ratio(CRB_PROF_SetupSamplesPerTri)); s += line;
        // bin / coarse stages: the reference's lines (cuda/PrivateDefs.hpp:168-187).  A "round" is one batch of 32 queue entries of a
        // warp; the three coverage paths are the footprint classes of the scatter (one cell / at most 2x2 cells / edge-refined);
        // there is no input ring buffer, no segment pool and no stream merge here, so those lines read 0.
        s += "BinRaster:\n";
        snprintf(line, sizeof(line), "- Input overflows      %.0f\n", 0.0); s += line;
        snprintf(line, sizeof(line), "- Avg. triangles/round %.1f\n", ratio(CRB_PROF_BinTrisPerRound)); s += line;
        snprintf(line, sizeof(line), "- Avg. tri bb size     %.1f\n", ratio(CRB_PROF_BinTriBBArea)); s += line;
        snprintf(line, sizeof(line), "- Coverage single path %.1f%%\n", ratio(CRB_PROF_BinTriSinglePath)); s += line;
        snprintf(line, sizeof(line), "- Coverage fast path   %.1f%%\n", ratio(CRB_PROF_BinTriFastPath)); s += line;
        snprintf(line, sizeof(line), "- Coverage slow path   %.1f%%\n", ratio(CRB_PROF_BinTriSlowPath)); s += line;
        snprintf(line, sizeof(line), "- Segment allocs/round %.1f\n", 0.0); s += line;
        snprintf(line, sizeof(line), "- Bin queue entries    %d\n\n", a.numBinEntries); s += line;
        s += "CoarseRaster:\n";
        snprintf(line, sizeof(line), "- Bins                 %.0f\n", (double)c->hostProf[2 * CRB_PROF_CoarseBins]); s += line;
        snprintf(line, sizeof(line), "- Rounds / Bin         %.1f\n", ratio(CRB_PROF_CoarseRoundsPerBin)); s += line;
        snprintf(line, sizeof(line), "- Merge / Round        %.1f\n", 0.0); s += line;
        snprintf(line, sizeof(line), "- Triangles / Round    %.1f\n", ratio(CRB_PROF_CoarseTrisPerRound)); s += line;
        snprintf(line, sizeof(line), "- Tiles / Round        %.1f\n", ratio(CRB_PROF_CoarseTilesPerRound)); s += line;
        snprintf(line, sizeof(line), "- Emits / Round        %.1f\n", ratio(CRB_PROF_CoarseEmitsPerRound)); s += line;
        snprintf(line, sizeof(line), "- Allocs / Round       %.1f\n", 0.0); s += line;
        snprintf(line, sizeof(line), "- Emits / Triangle     %.2f\n", ratio(CRB_PROF_CoarseEmitsPerTri)); s += line;
        snprintf(line, sizeof(line), "- Case A               %.0f%%\n", ratio(CRB_PROF_CoarseCaseA)); s += line;
        snprintf(line, sizeof(line), "- Case B               %.0f%%\n", 0.0); s += line;
        snprintf(line, sizeof(line), "- Case C               %.0f%%\n\n", ratio(CRB_PROF_CoarseCaseC)); s += line;
        s += "FineRaster:\n- Triangles culled\n";
        snprintf(line, sizeof(line), "  - Early Z kill       %.1f%%\n", ratio(CRB_PROF_FineEarlyZCull)); s += line;
        snprintf(line, sizeof(line), "  - Empty coverage     %.1f%%\n", ratio(CRB_PROF_FineEmptyCull)); s += line;
        snprintf(line, sizeof(line), "- Z kills              %.1f%%\n", ratio(CRB_PROF_FineZKill)); s += line;
        snprintf(line, sizeof(line), "- MSAA kills           %.1f%%\n", ratio(CRB_PROF_FineMSAAKill)); s += line;
        snprintf(line, sizeof(line), "- Avg. tri/tile        %.0f\n", ratio(CRB_PROF_FineTriPerTile)); s += line;
        snprintf(line, sizeof(line), "- Avg. frag/tri        %.1f\n", ratio(CRB_PROF_FineFragPerTri)); s += line;
        snprintf(line, sizeof(line), "- Avg. frag/tile       %.0f\n", ratio(CRB_PROF_FineFragPerTile)); s += line;
        s += "\n";
        snprintf(buf, bufSize, "%s", s.c_str());
        return CRB_OK;
    }
    s += "ProfilingMode_Default\n---------------------\n\n";
    const char* names[4] = {"triangleSetup", "binRaster", "coarseRaster", "fineRaster"};
    for (int i = 0; i < 4; i++) {
        snprintf(line, sizeof(line), "%-16s%.3f ms (%.0f%%)\n", names[i], st[i] * 1.0e3f, st[i] * pct);
        s += line;
    }
    s += "\n";
    snprintf(line, sizeof(line), "%-16s%-10d(%.1f MB)\n", "numSubtris", a.numSubtris, (float)a.numSubtris * 81.0f / 1048576.0f);
    s += line;
    snprintf(line, sizeof(line), "%-16s%-10d(%.1f MB)\n", "numBinEntries", a.numBinEntries, (float)a.numBinEntries * 4.0f / 1048576.0f);
    s += line;
    snprintf(line, sizeof(line), "%-16s%-10d(%.1f MB)\n", "numTileEntries", a.numTileEntries, (float)a.numTileEntries * 4.0f / 1048576.0f);
    s += line;
    snprintf(line, sizeof(line), "%-16s%-10d\n", "numActiveTiles", a.numActiveTiles);
    s += line;
    s += "\n";
    snprintf(buf, bufSize, "%s", s.c_str());
    return CRB_OK;
}

int crb_get_work_buffers(crb_ctx* c, crb_work_buffers* out) {
    if (!c || !out) return CRB_ERR_INVALID;
    const crb_frame& f = c->frame;
    out->triSubtris = f.triSubtris;
    out->triHeader = f.triHeader;
    out->triData = f.triData;
    out->maxSubtris = f.maxSubtris;
    out->binQueue = f.binQueue;
    out->binStart = f.binStart;
    out->binTotal = f.binTotal;
    out->numBins = f.numBins;
    out->tileQueue = f.tileQueue;
    out->tileStart = f.tileStart;
    out->tileCount = f.tileCount;
    out->numTiles = f.numTiles;
    out->activeTiles = f.activeTiles;
    return CRB_OK;
}

int crb_launch_vertex_shader(void* module, const char* name, const void* d_in, void* d_out, int numVertices, const void* h_constants, size_t constantsBytes, void* stream) {
    if (!name) return CRB_ERR_INVALID;
    void* handle = module;
    if (!handle) {
        Dl_info info;
        if (dladdr((const void*)&crb_abi_version, &info) && info.dli_fname) handle = dlopen(info.dli_fname, RTLD_NOW | RTLD_NOLOAD);
        if (!handle) handle = dlopen(nullptr, RTLD_NOW);
    }
    crb_vertex_shader_fn fn = (crb_vertex_shader_fn)dlsym(handle, (std::string(name) + "_launch").c_str());
    if (!fn) return CRB_ERR_INVALID;
    return fn(d_in, d_out, numVertices, h_constants, constantsBytes, stream);
}

int crb_write_ppm(const char* path, const uint32_t* px, int width, int height, int pitch) {
    if (!path || !px || width <= 0 || height <= 0 || pitch < width) return CRB_ERR_INVALID;
    FILE* fp = fopen(path, "wb");
    if (!fp) return CRB_ERR_INVALID;
    fprintf(fp, "P6\n%d %d\n255\n", width, height);
    std::vector<unsigned char> row((size_t)width * 3);
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < width; x++) {
            const uint32_t t = px[(size_t)y * pitch + x];
            row[3 * x + 0] = (unsigned char)(t & 0xFF);
            row[3 * x + 1] = (unsigned char)((t >> 8) & 0xFF);
            row[3 * x + 2] = (unsigned char)((t >> 16) & 0xFF);
        }
        if (fwrite(row.data(), 1, row.size(), fp) != row.size()) { fclose(fp); return CRB_ERR_INVALID; }
    }
    fclose(fp);
    return CRB_OK;
}

int crb_download(crb_ctx* c, const void* d_src, void* h_dst, size_t bytes) {
    if (!c || (bytes && (!d_src || !h_dst))) return CRB_ERR_INVALID;
    CRB_CUDA(c, cudaSetDevice(c->device));
    CRB_CUDA(c, cudaDeviceSynchronize());
    CRB_CUDA(c, cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return CRB_OK;
}

}  // extern "C"
