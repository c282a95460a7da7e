// Work-buffer formats shared by the library kernels and pixel-pipe translation units.
// CRTriangleHeader / CRTriangleData / PixelPipeSpec keep the reference's layout bit for bit
// (src/cudaraster/cuda/PrivateDefs.hpp:26-62, :147-154) because the parity tests compare them
// against the golden model and the rebuilt reference kernels.  crb_frame replaces CRParams
// (:68-120): it is passed BY VALUE as the kernel argument (no __constant__ upload per launch).
#pragma once
#include "../../crb200.h"
#include "../base/Math.hpp"
#include "Constants.hpp"

namespace FW {

struct CRTriangleHeader {  // 16 B
    S16 v0x, v0y;          // subpixels relative to the viewport centre; valid if triSubtris == 1
    S16 v1x, v1y;
    S16 v2x, v2y;
    U32 misc;              // triSubtris == 1: zmin:20 | f01:4 | f12:4 | f20:4;  >= 2: sub-triangle base
};

struct CRTriangleData {  // 64 B
    U32 zx, zy, zb, zslope;  // depth = zx*sampleX + zy*sampleY + zb (U32 wrap)
    S32 wx, wy, wb;          // evaluated at (sampleX*2+1, sampleY*2+1): minW/w * CR_BARY_MAX
    S32 ux, uy, ub;
    S32 vx, vy, vb;
    U32 vi0, vi1, vi2;
};

typedef crb_pipe_spec PixelPipeSpec;
typedef crb_atomics CRAtomics;

}  // namespace FW

#define CRB_TILECODE_GENERAL 0x40000000u

// Profiling modes of a pixel pipe (reference: cuda/PrivateDefs.hpp:161-270, CR_PROFILING_MODE).  Counters: the kernels of a
// pipe compiled with -DCR_PROFILING_MODE=ProfilingMode_Counters add to crb_frame::profCounters -- the counters of the
// reference's report that have a meaning in this pipeline, each a numerator / denominator pair (CRProfCounter).
#define ProfilingMode_Default 0
#define ProfilingMode_Counters 1
#define ProfilingMode_Timers 2   // clock64() brackets around the code regions of the reference's timers that exist here (lane 0 of every warp)
enum {
    CRB_PROF_SetupViewportCull = 0, CRB_PROF_SetupBackfaceCull, CRB_PROF_SetupBetweenPixelsCull, CRB_PROF_SetupClipped, CRB_PROF_SetupSamplesPerTri,
    CRB_PROF_FineEarlyZCull, CRB_PROF_FineEmptyCull, CRB_PROF_FineZKill, CRB_PROF_FineMSAAKill, CRB_PROF_FineTriPerTile, CRB_PROF_FineFragPerTri,
    CRB_PROF_FineFragPerTile,
    // bin / coarse stages (reference: cuda/PrivateDefs.hpp:168-187).  A "round" is one batch of 32 queue entries of a warp; the
    // reference's three coverage paths map onto the footprint classes of scatterBatch: one cell / at most 2x2 cells / refined.
    CRB_PROF_BinTrisPerRound, CRB_PROF_BinTriBBArea, CRB_PROF_BinTriSinglePath, CRB_PROF_BinTriFastPath, CRB_PROF_BinTriSlowPath,
    CRB_PROF_CoarseBins, CRB_PROF_CoarseRoundsPerBin, CRB_PROF_CoarseTrisPerRound, CRB_PROF_CoarseTilesPerRound, CRB_PROF_CoarseEmitsPerRound, CRB_PROF_CoarseEmitsPerTri,
    CRB_PROF_CoarseCaseA, CRB_PROF_CoarseCaseC,
    CRB_PROF_NUM
};
// ProfilingMode_Timers (reference: CR_PROFILING_TIMERS, cuda/PrivateDefs.hpp:207-270): clock totals, stored behind the counters
enum {
    CRB_TIMER_SetupTotal = 0, CRB_TIMER_SetupVertexRead, CRB_TIMER_SetupCullSnap, CRB_TIMER_SetupPleq, CRB_TIMER_SetupClip, CRB_TIMER_SetupBinning,
    CRB_TIMER_FineTotal, CRB_TIMER_FineReadTile, CRB_TIMER_FinePixelCoverage, CRB_TIMER_FineZKill, CRB_TIMER_FineShade, CRB_TIMER_FineWriteTile,
    // bin stage: scan kernel (one CTA per bin) + scatter kernel (one warp per chunk); coarse stage likewise
    CRB_TIMER_BinTotal, CRB_TIMER_BinScan, CRB_TIMER_BinReadTriHeader, CRB_TIMER_BinRasterize, CRB_TIMER_BinCountTiles,
    CRB_TIMER_CoarseTotal, CRB_TIMER_CoarseScan, CRB_TIMER_CoarseStreamRead, CRB_TIMER_CoarseRasterize,
    CRB_TIMER_NUM
};
#define CRB_PROF_WORDS (2 * CRB_PROF_NUM + CRB_TIMER_NUM)

// One coarse work item: `count` consecutive entries of one bin's queue.
struct crb_item {
    int32_t bin;
    int32_t first;   // index into binQueue
    int32_t count;
    int32_t slot;    // k-th item of its bin (row of tileCountMat = itemBase[bin] + slot)
};

struct crb_frame {
    // ---- input (CRParams: numTris, vertexBuffer, indexBuffer)
    int32_t numTris;
    int32_t vertexStride;         // bytes per vertex (multiple of 16)
    const void* vertexBuffer;
    const int32_t* indexBuffer;   // numTris x int3

    // ---- viewport (CRParams: viewportWidth .. numTiles)
    int32_t viewportWidth, viewportHeight;   // size of the viewport the triangles are set up in: the surface itself, or the PARENT
                                             // frame (<= 2048^2) of a sort-first window; header coordinates are relative to its centre
    int32_t widthPixels, heightPixels;       // rounded to tiles
    int32_t widthBins, heightBins, numBins;
    int32_t widthTiles, heightTiles, numTiles;
    int32_t samplesLog2;
    // sort-first window: vertices are snapped in the grid of the FULL frame and shifted by the integer
    // offset of the parent viewport's centre; the surface is a scissor rectangle inside the parent viewport.
    int32_t fullWidth, fullHeight, centerOfsX, centerOfsY;
    int32_t windowed;                        // 1 = sort-first window (surface != parent viewport)
    int32_t subX0, subY0;                    // pixel origin of the surface (scissor) inside the parent viewport, multiples of 8
    int32_t originX, originY;                // header subpixel coordinate + origin = subpixel position relative to the surface corner
    float clipLoX, clipHiX, clipLoY, clipHiY;   // parent viewport in full-frame NDC: the window triangles are clipped against
    float cullLoX, cullHiX, cullLoY, cullHiY;   // surface (scissor) in full-frame NDC: triangles wholly outside are culled early
    const float4* chunkBounds;               // sort-first window: {min x/w, min y/w, max x/w, max y/w} per CRB_SETUP_THREADS input triangles, or nullptr

    int32_t deferredClear;
    uint32_t clearColor, clearDepth;
    uint32_t* colorBuffer;        // linear surfaces, pitch in texels = widthPixels << samplesLog2
    uint32_t* depthBuffer;
    int32_t surfacePitch;
    int32_t colorPitch;           // texels per row of the colour surface: surfacePitch, or the pitch of a larger image the surface is a
                                  // window of (crb_set_color_pitch: sort-first windows rendered straight into the full frame)
    int32_t colorTiled;           // 1 = the colour surface is tile-major (single sample only; crb_set_color_layout)

    // ---- setup output
    int32_t maxSubtris;
    uint8_t* triSubtris;
    uint4* triHeader;
    uint4* triData;               // 4 x uint4 per sub-triangle

    // ---- bin stage: count matrix -> exclusive offsets, CSR queue
    int32_t chunkTris;            // consecutive triangles per chunk = CRB_SETUP_THREADS * ctasPerChunk
    int32_t ctasPerChunk;         // setup CTAs that share one chunk (power of two; 1 for <= 4M triangles)
    int32_t numChunks;            // ceil(numTris / chunkTris)
    int32_t matPitch;             // numChunks rounded up to a multiple of 4
    int32_t* binCountMat;         // [numBins][matPitch]; counts, then exclusive prefix over chunks
    int32_t* binStart;            // [CR_MAXBINS_SQR]
    int32_t* binTotal;            // [CR_MAXBINS_SQR]
    int32_t maxBinEntries;
    int32_t* binQueue;            // [maxBinEntries] triIdx*8 + (single ? 7 : sub)

    // ---- coarse stage: work items, per-item tile counts, CSR queue
    int32_t maxItems;
    crb_item* items;              // [maxItems]
    int32_t* binItemBase;         // [CR_MAXBINS_SQR] first item row of each bin
    int32_t* binItemCount;        // [CR_MAXBINS_SQR]
    int32_t* tileCountMat;        // [maxItems][CR_BIN_SQR]; counts, then exclusive prefix over items
    int32_t maxTileEntries;
    int32_t* tileQueue;           // [maxTileEntries]
    int32_t* tileStart;           // [CR_MAXTILES_SQR] indexed by global tile index
    int32_t* tileCount;           // [CR_MAXTILES_SQR]
    int32_t* activeTiles;         // [CR_MAXTILES_SQR]
    int4* activeRecs;             // [CR_MAXTILES_SQR] {tile index, queue start, queue count, 0}: one load per fine warp

    // ---- direct tile path (crb_set_binning_mode); tileStart / tileCount / activeRecs / tileQueue as above
    int32_t directMode;           // 1 = this frame runs setup -> directAlloc -> directScatter -> fine, queues unordered
    int32_t* tileCounter;         // [CR_MAXTILES_SQR] entries per tile, counted by setup, zeroed again by directAllocKernel
    int32_t* tileCursor;          // [CR_MAXTILES_SQR] end of the tile's queue extent, counted down by the scatter pass
    // Micro-triangle visibility (single-sample direct frames): setup rasterizes triangles whose pixel footprint is at
    // most 4x4 straight into this per-pixel buffer with 64-bit atomicMin on (depth << 32 | entry + 1) -- the same
    // (depth, index) minimum the fine raster keeps -- and never queues them; the fine raster merges the buffer into
    // its tile state and writes the neutral value (all ones) back, so the buffer is clean for the next frame.
    int32_t microMode;            // 1 = on (implies directMode)
    unsigned long long* visBuffer; // [heightPixels][widthPixels]
    int2* largeList;              // [maxLarge] {queue entry, record slot} of the LARGE sub-triangles (> CRB_DIRECT_MAX_TILES tiles on an axis), appended
    int32_t maxLarge;             // by setup; the allocation and scatter kernels walk it with whole CTAs (one large triangle per CTA at a time)
    uint8_t* batchQueued;         // [ceil(numTris / 32)] 1 = the 32 triangles of the batch have their words in triTileCode (one of them was queued)
    uint32_t* triTileCode;        // [numTris] what the scatter pass needs to know about a triangle in ONE word: 0 = nothing to place,
                                  // CRB_TILECODE_GENERAL = go through triSubtris / the headers (clipped, refined or large), else
                                  // tile x0 | y0 << 8 | (nx-1) << 16 | (ny-1) << 17 | 1 << 31 of a footprint of at most 2x2 tiles

    unsigned long long* profCounters;   // [CRB_PROF_WORDS]: counter pairs, then timers; zeroed before every frame of a profiling pipe
    int32_t profilingMode;              // the pipe's CR_PROFILING_MODE: picks the instrumented instances of the bin / coarse kernels (they live in the library)

    crb_atomics* atomics;         // counters of THIS frame (zero when the frame starts)
    crb_atomics* nextAtomics;     // counters of the next frame: zeroed by this frame's fine raster kernel
    crb_atomics* hostCounters;    // mapped host memory: the fine raster kernel leaves a copy of this frame's counters there (no D2H copy on the stream)
    int32_t numSMs;
    int32_t debugFlags;           // CRB_DEBUG_FLAGS environment variable; bit 0: fine raster always takes the general coverage path
    int32_t chainLaunches;        // 1 = kernels are launched with programmatic stream serialization (see Util.cuh)
};
