/* crb200.h -- C ABI of the B200-native CudaRaster pipeline (libcrb200.so).
 *
 * Drop-in boundary for the hot path of tcoppex/cudaraster-linux:
 *   FW::CudaRaster::drawTriangles() -> triangleSetup -> binRaster -> coarseRaster -> fineRaster
 *   (reference: src/cudaraster/CudaRaster.cpp:237-342, :508-665).
 * Every entry point cites the reference interface it replaces.  Plain pointers and sizes only;
 * all functions return 0 on success or a non-zero crb_status (the C++ shim in
 * include/cudaraster/CudaRaster.hpp turns non-zero into the reference's fail() convention,
 * src/framework/base/Defs.hpp:107-117).  There is no CPU fallback: without a CUDA device
 * crb_create() fails.
 */
#ifndef CRB200_H_
#define CRB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRB_ABI_VERSION 3

typedef enum crb_status {
    CRB_OK = 0,
    CRB_ERR_INVALID = 1,      /* bad argument / state (message in crb_last_error) */
    CRB_ERR_CUDA = 2,         /* a CUDA runtime call failed                         */
    CRB_ERR_NO_DEVICE = 3,    /* no CUDA device: the product has no CPU path        */
    CRB_ERR_LIMIT = 4,        /* a format limit was exceeded (CR_MAXSUBTRIS_SIZE...) */
    CRB_ERR_OVERFLOW = 5      /* crb_finish: an asynchronous frame overflowed a work buffer and must be redrawn */
} crb_status;

/* Render-mode flags: cuda/PixelPipe.hpp:30-35. */
enum { CRB_FLAG_DEPTH = 1, CRB_FLAG_LERP = 2, CRB_FLAG_QUADS = 4 };

/* Format limits: cuda/Constants.hpp:21-39. */
enum {
    CRB_MAX_VIEWPORT = 2048,
    CRB_MAX_SAMPLES = 8,
    CRB_MAX_SUBTRIS = 1 << 24,
    CRB_TILE_SIZE = 8,
    CRB_BIN_TILES = 16
};

typedef struct crb_ctx crb_ctx;

/* == PixelPipeSpec, cuda/PrivateDefs.hpp:147-154 (same field order and sizes). */
typedef struct crb_pipe_spec {
    int32_t samplesLog2;
    int32_t vertexStructSize;
    uint32_t renderModeFlags;
    int32_t profilingMode;
    char blendShaderName[128];
} crb_pipe_spec;

/* == CRAtomics, cuda/PrivateDefs.hpp:123-143, extended with the counters of the new
 * count/scan/scatter binning (the segment counters of the reference do not exist here). */
typedef struct crb_atomics {
    int32_t numSubtris;        /* starts at numTris; += n for every triangle clipped into n>=2  */
    int32_t numBinEntries;     /* triangle-bin pairs written to the bin queue                  */
    int32_t numCoarseItems;    /* work items the coarse stage processed                        */
    int32_t numTileEntries;    /* triangle-tile pairs written to the tile queue                */
    int32_t numActiveTiles;    /* tiles the fine stage touches                                 */
    int32_t overflow;          /* bit0 subtris, bit1 bin queue, bit2 tile queue, bit3 items, bit4 an earlier frame of the batch overflowed, bit5 large list */
    int32_t numLargeTris;      /* direct path: sub-triangles spanning > CRB_DIRECT_MAX_TILES tiles on an axis (counted and scattered by whole CTAs from a global list) */
    int32_t numQueuedCtas;     /* direct path: setup CTAs that put at least one sub-triangle on a tile queue (0 = everything went the micro way) */
    int32_t allocBarrier;      /* direct path: grid barrier of the queue-allocation kernel (internal) */
    int32_t reserved;
} crb_atomics;

/* Everything a stage launcher needs; filled by crb_draw_triangles().  Opaque to C callers,
 * defined in include/cudaraster/cuda/PrivateDefs.hpp for pixel-pipe translation units. */
typedef struct crb_frame crb_frame;

/* A stage launcher enqueues one stage on `stream` (a cudaStream_t) and returns a crb_status.
 * The four launchers of a pipe replace the four kernels CR_DEFINE_PIXEL_PIPE emits
 * (cuda/PixelPipe.inl:241-275) and are found by the same string names
 * (<pipe>_triangleSetup ... <pipe>_fineRaster, <pipe>_spec; CudaRaster.cpp:190-201). */
typedef int (*crb_stage_fn)(const crb_frame* frame, void* stream);

typedef struct crb_pipe_desc {
    const char* name;
    const crb_pipe_spec* spec;
    crb_stage_fn triangleSetup;
    crb_stage_fn binRaster;
    crb_stage_fn coarseRaster;
    crb_stage_fn fineRaster;
    int32_t orderIndependent;   /* 1 = depth test on, blend ignores dst, shader never discards: the surviving fragment of a sample is
                                 * the (depth, submission index) minimum whatever the processing order (<pipe>_orderIndependent,
                                 * emitted by CR_DEFINE_PIXEL_PIPE).  Lets crb_draw_triangles use the direct tile path. */
} crb_pipe_desc;

/* ---- lifetime: CudaRaster::CudaRaster/init/~CudaRaster, CudaRaster.cpp:53-128 ---------------- */
int crb_abi_version(void);
int crb_create(int device, crb_ctx** out);
int crb_destroy(crb_ctx* ctx);
const char* crb_last_error(const crb_ctx* ctx);

/* ---- state setters ---------------------------------------------------------------------------
 * crb_set_surfaces: CudaRaster::setSurfaces (CudaRaster.cpp:132-170) + CudaSurface
 * (CudaSurface.cpp:39-91).  Surfaces are LINEAR device memory, U32 texels,
 * row pitch = roundedWidth * numSamples texels, rows = roundedHeight (rounded = up to a
 * multiple of 8).  Sample i of pixel (x,y) lives at texel ((x>>3)*8*N + i*8 + (x&7), y), the
 * reference's horizontally tile-replicated MSAA layout (FineRaster.inl:909, :1088-1100).
 * Colour texel = 0xAABBGGRR, depth texel = raw encoded U32.  Row 0 is the bottom scanline. */
int crb_set_surfaces(crb_ctx* ctx, void* d_color, void* d_depth, int width, int height, int numSamples);

/* CudaRaster::deferredClear (CudaRaster.cpp:174-179): clear on the next draw. */
int crb_deferred_clear(crb_ctx* ctx, uint32_t abgr, uint32_t encodedDepth);
/* Vec4f::toABGR (base/Math.cpp:41-48) and the depth encoding of CudaRaster.cpp:178. */
uint32_t crb_pack_abgr(float r, float g, float b, float a);
uint32_t crb_encode_clear_depth(float depth);

/* CudaRaster::setPixelPipe (CudaRaster.cpp:183-216).  The _by_name form resolves
 * "<name>_triangleSetup" ... "<name>_spec" with dlsym in `module` (a dlopen handle of a pixel-pipe
 * shared object; NULL = the pipes built into libcrb200.so), exactly like
 * CudaModule::getKernel(name + "_triangleSetup"). */
int crb_set_pixel_pipe(crb_ctx* ctx, const crb_pipe_desc* pipe);
int crb_set_pixel_pipe_by_name(crb_ctx* ctx, void* module, const char* name);

/* CudaRaster::setVertexBuffer / setIndexBuffer (CudaRaster.cpp:220-233); device pointers.
 * Vertices: vertexStructSize bytes each, clipPos (4 x F32) first, then 4 x F32 varyings.
 * Indices: numTris x int32[3]. */
int crb_set_vertex_buffer(crb_ctx* ctx, const void* d_vertices, size_t bytes);
int crb_set_index_buffer(crb_ctx* ctx, const void* d_indices, int numTris);

/* Sort-first sub-viewport (new): this context renders the rectangle
 * [x0,x0+width) x [y0,y0+height) (multiples of 8) of a fullWidth x fullHeight frame whose
 * clip space the vertices are expressed in.  width/height must match crb_set_surfaces.
 * Vertices are snapped once in full-frame subpixels so seams are watertight for any split.
 * Passing fullWidth = 0 restores the plain single viewport. */
int crb_set_subviewport(crb_ctx* ctx, int fullWidth, int fullHeight, int x0, int y0);

/* Sort-first partition of a frame (new): cuts a fullWidth x fullHeight frame into AT LEAST `parts` rectangles
 * {x0, y0, w, h} (row-major order, origins multiples of 8) none of which straddles a parent cell of crb_set_subviewport --
 * first the ceil(full / 2048) x ceil(full / 2048) cells, then every cell into the same power-of-two grid.  Writes up to
 * maxRects rectangles (4 ints each) and returns their number (call with maxRects = 0 to size the array); < 0 on bad arguments.
 * Rank r of `world` renders rectangles r, r + world, ... */
int crb_split_frame(int fullWidth, int fullHeight, int parts, int* outRects, int maxRects);

/* Sort-first geometry cull (new).  Every rank of a sort-first split sets up ALL triangles of the mesh and drops those outside
 * its window one by one; with per-chunk bounds a rank skips whole chunks -- CRB_CHUNK_BOUNDS_TRIS consecutive input
 * triangles -- whose clip-space bounding box lies outside the window, without touching their vertices.  The bounds depend
 * on the mesh only (not on the window): crb_compute_chunk_bounds makes them once per mesh (device, stream ordered) --
 * d_bounds[chunk] = {min x/w, min y/w, max x/w, max y/w}, the whole plane for a chunk with a vertex at w <= 0 --, and
 * crb_set_chunk_bounds hands them to a context for the CURRENT vertex / index buffers (setting either buffer forgets them).
 * A skipped chunk is exactly a chunk all of whose triangles the per-triangle window test would have culled: frames and
 * setup records are unchanged. */
#define CRB_CHUNK_BOUNDS_TRIS 256
int crb_compute_chunk_bounds(const void* d_vertices, int vertexStride, const int32_t* d_indices, int numTris, float* d_bounds, void* stream);
int crb_set_chunk_bounds(crb_ctx* ctx, const float* d_bounds);

/* Binning strategy (new).  The general path is the stable two-level sort (bin raster + coarse raster) that keeps
 * every tile queue in submission order, like the reference.  The DIRECT path skips both levels: triangle setup counts
 * the tiles of every triangle, one kernel allocates the tile queues and one scatters the entries with atomics, in
 * arbitrary order -- valid only for order-independent pipes (crb_pipe_desc.orderIndependent); it is built for frames
 * of small triangles (a triangle spanning more than CRB_DIRECT_MAX_TILES tiles on an axis is handled by a whole CTA
 * at a time: rows of tiles are shared out over its threads and the covered span of a row is found by bisection).  mode 0 =
 * never, 1 = automatic (default) = 2 = direct on every frame whose pipe allows it -- the decision needs no knowledge of the
 * scene, so the first frame and frames with a changing triangle count take the direct path too --, 3 = like 2 but without
 * the micro-triangle visibility buffer (below; a testing aid: setup then writes every record in full).  Surfaces are
 * bit-identical on all paths.
 * Micro-triangles: on single-sample direct frames, triangle setup rasterizes every triangle whose pixel-centre
 * footprint is at most 4x4 pixels itself -- exact coverage, plane depth, one 64-bit atomicMin of (depth << 32 | index)
 * per covered pixel into a per-pixel visibility buffer -- and never queues it; the fine raster merges the buffer
 * into its tile state under the same (depth, index) rule and restores it. */
#define CRB_DIRECT_MAX_TILES 4
int crb_set_binning_mode(crb_ctx* ctx, int mode);
/* 1 when the last frame ran on the direct path. */
int crb_get_last_frame_direct(crb_ctx* ctx);

/* ---- the hot path ----------------------------------------------------------------------------
 * CudaRaster::drawTriangles (CudaRaster.cpp:237-342): sizes the work buffers, launches the four
 * stages on `stream` (NULL = default stream), reads the counters back and, if a queue
 * overflowed, grows the buffers and re-runs the frame, like the reference's retry loop. */
int crb_draw_triangles(crb_ctx* ctx, void* stream);

/* Asynchronous variant (new; the reference blocks on the counter read-back every frame,
 * CudaRaster.cpp:326): enqueues one frame with the CURRENT work-buffer capacities and returns without
 * synchronizing, so consecutive frames run back to back on the GPU.  The counters of every frame
 * are copied to pinned host memory; crb_finish() synchronizes `stream` and checks them: CRB_OK, or
 * CRB_ERR_OVERFLOW if some frame overflowed a queue (its output is incomplete and so may be that of
 * every frame enqueued after it in the same batch -- the message names the first one; the capacities
 * have been grown, redraw them -- a synchronous crb_draw_triangles() of the same scene first makes
 * that impossible).  At most 64 frames may be pending; the 65th call finishes implicitly. */
int crb_draw_triangles_async(crb_ctx* ctx, void* stream);
int crb_finish(crb_ctx* ctx, void* stream);

/* The same call for HOST buffers (the reference's Buffer class mirrors host memory to the device
 * on demand, gpu/Buffer.cpp:235-346): uploads vertices + indices, draws with a deferred clear if
 * one is pending, downloads colour (and depth when h_depth != NULL).  Host pointers should be
 * pinned for full PCIe rate.  This is the end-to-end entry the benchmark times. */
int crb_draw_triangles_host(crb_ctx* ctx, const void* h_vertices, size_t vertexBytes, const int32_t* h_indices, int numTris,
                            uint32_t* h_color, uint32_t* h_depth, void* stream);

/* Pipelined variant of crb_draw_triangles_host for streams of frames: the upload of frame k+1 (own
 * copy stream, double-buffered staging), the rendering of frame k (crb_draw_triangles_async on `stream`)
 * and the download of frame k-1 (own copy stream) overlap.  Returns without blocking; the host output
 * buffers are valid after crb_finish(ctx, stream) (or after synchronising `stream`), which also reports
 * work-buffer overflow like crb_draw_triangles_async.  Host buffers must be pinned and must stay
 * untouched until then.  The deferred clear, pipe and surfaces are those set when the call is made. */
int crb_draw_triangles_host_async(crb_ctx* ctx, const void* h_vertices, size_t vertexBytes, const int32_t* h_indices, int numTris,
                                  uint32_t* h_color, uint32_t* h_depth, void* stream);

/* CudaRaster::getStats (CudaRaster.cpp:346-363): seconds per stage of the last draw
 * {setup, bin, coarse, fine}.  Synchronizes. */
int crb_get_stats(crb_ctx* ctx, float outSeconds[4]);
/* Stage timing of ASYNCHRONOUS frames (the profiling hooks of CudaRaster.cpp:593-661 for the
 * non-blocking entry): when enabled, crb_draw_triangles_async records the five stage events as well
 * (this splits the kernel chain, so it is off by default) and crb_finish accumulates them;
 * crb_get_stage_timing returns the mean milliseconds per stage over the frames finished since
 * crb_set_stage_timing was last called. */
int crb_set_stage_timing(crb_ctx* ctx, int enable);
int crb_get_stage_timing(crb_ctx* ctx, double outMeanMs[4], int* outFrames);
/* The same per frame: FIVE floats for each of up to maxFrames frames finished since crb_set_stage_timing, oldest first -- the
 * four stage intervals (ms) and, fifth, the duration of the frame's composite copy (crb_draw_batch_async with pushDst; 0
 * without one); returns the number of frames written.  Mtris/s as SURVEY.md 8(d) defines it is numTris / median over frames
 * of the sum of a frame's four intervals. */
int crb_get_stage_timing_frames(crb_ctx* ctx, float* outMs, int maxFrames);

/* A stream of frames in ONE call (new): for every element, the state setters that are non-NULL (surfaces, vertex buffer, index
 * buffer, deferred clear) followed by crb_draw_triangles_async on `stream` -- the frame loop of an application, or of the
 * benchmark, without a host-language round trip per frame.  Stops at the first error. */
typedef struct crb_batch_frame {
    void* color;                 /* NULL with depth NULL: keep the current surfaces */
    void* depth;
    int32_t width, height, numSamples;
    const void* vertices;        /* NULL: keep the current vertex buffer */
    size_t vertexBytes;
    const void* indices;         /* NULL: keep the current index buffer */
    int32_t numTris;
    int32_t clear;               /* != 0: crb_deferred_clear(clearColor, clearDepth) before the draw */
    uint32_t clearColor, clearDepth;
    /* Composite step of a multi-GPU frame, enqueued with the frame so that a frame costs ONE host call:
     *   pushDst != NULL   the finished colour surface (pushBytes bytes) is copied into pushDst -- a frame slot in the display
     *                     GPU's memory (crb_ipc_open) -- by the DMA engines on an internal side stream, overlapped with the next
     *                     frame; `surfaceSlot` (0..3) names the local colour surface the frame renders into: the render waits
     *                     until the previous copy out of that surface has finished;
     *   signalWord != NULL  signalValue is stored there after the frame (after its copy, if any): the consumer's frame mark. */
    void* pushDst;
    size_t pushBytes;
    void* signalWord;
    uint32_t signalValue;
    int32_t surfaceSlot;
} crb_batch_frame;
int crb_draw_batch_async(crb_ctx* ctx, const crb_batch_frame* frames, int numFrames, void* stream);
/* Makes `stream` wait for the copies crb_draw_batch_async put on the side stream (so that an event recorded on `stream`
 * afterwards covers the composite); crb_finish waits for them as well. */
int crb_batch_join(crb_ctx* ctx, void* stream);
/* g_crAtomics read-back (CudaRaster.cpp:326). */
int crb_get_counters(crb_ctx* ctx, crb_atomics* out);
/* CudaRaster::getProfilingInfo (CudaRaster.cpp:367-497): the ProfilingMode_Default report, or -- for a pipe compiled with
 * -DCR_PROFILING_MODE=ProfilingMode_Counters -- the counters report (the reference's counters that exist in this pipeline). */
int crb_get_profiling_info(crb_ctx* ctx, char* buf, size_t bufSize);
/* Number of kernels the last crb_draw_triangles enqueued (all retries included). */
int crb_get_launch_count(crb_ctx* ctx);

/* ---- inspection of the work buffers (parity tests; the reference exposes the same data through
 * its Buffer members and DebugParams, CudaRaster.hpp:53-67, :109-138) ------------------------- */
typedef struct crb_work_buffers {
    const uint8_t* triSubtris;   /* [maxSubtris] U8                                    */
    const void* triHeader;       /* [maxSubtris] 16 B  (cuda/PrivateDefs.hpp:26-36)   */
    const void* triData;         /* [maxSubtris] 64 B  (cuda/PrivateDefs.hpp:40-62)   */
    int32_t maxSubtris;
    const int32_t* binQueue;     /* triIdx*8+sub, grouped by bin, submission order     */
    const int32_t* binStart;     /* [numBins] first entry of each bin                  */
    const int32_t* binTotal;     /* [numBins] entries per bin                          */
    int32_t numBins;
    const int32_t* tileQueue;    /* triIdx*8+sub, grouped by tile, submission order    */
    const int32_t* tileStart;    /* [numTiles]                                         */
    const int32_t* tileCount;    /* [numTiles]                                         */
    int32_t numTiles;
    const int32_t* activeTiles;  /* [numActiveTiles]                                   */
} crb_work_buffers;
int crb_get_work_buffers(crb_ctx* ctx, crb_work_buffers* out); /* device pointers, valid until the next draw */
/* Buffer::getPtr() of the reference (gpu/Buffer.cpp:235-346: device -> host mirror): blocking copy of
 * `bytes` bytes of device memory (a work buffer or a surface) into host memory. */
int crb_download(crb_ctx* ctx, const void* d_src, void* h_dst, size_t bytes);

/* ---- around the hot path -------------------------------------------------------
 * MSAA resolve: CudaSurface::resolveToScreen (CudaSurface.hpp:73, body dropped by the Linux port; the demo calls
 * its own after drawTriangles, test/SceneCR.cpp:297).  Box filter of the numSamples samples of every pixel of a
 * surface in the layout of crb_set_surfaces into a LINEAR image of dstPitch texels per row: per 8-bit channel
 * (sum + N/2) >> log2 N.  flipY != 0 writes the top scanline first (image order).  Enqueued on `stream`. */
int crb_resolve_surface(const void* d_src, int width, int height, int numSamples, void* d_dst, int dstPitch, int flipY, void* stream);
/* Image writer for golden images / visual diffing: binary PPM (P6) of a HOST 0xAABBGGRR image, rows as given. */
int crb_write_ppm(const char* path, const uint32_t* h_pixels, int width, int height, int pitch);

/* Vertex-shader stage: the user kernel the demo launches before drawTriangles (test/shader/PassThrough.cu:16-35,
 * test/shader/Shaders.cu:56-112, launch at test/SceneCR.cpp:263-282).  CR_DEFINE_VERTEX_SHADER (cuda/PixelPipe.inl)
 * emits `<name>_launch` with this signature; constants (the reference's c_constants block) travel by value as a
 * kernel argument (<= CRB_MAX_VS_CONSTANTS bytes).  The kernel is enqueued on `stream` with programmatic stream
 * serialization, so a following crb_draw_triangles_async on the same stream starts its prologue while it drains. */
#define CRB_MAX_VS_CONSTANTS 1024
typedef int (*crb_vertex_shader_fn)(const void* d_inVertices, void* d_outVertices, int numVertices, const void* h_constants, size_t constantsBytes, void* stream);
/* Resolves "<name>_launch" in `module` (NULL = libcrb200.so) and calls it. */
int crb_launch_vertex_shader(void* module, const char* name, const void* d_inVertices, void* d_outVertices, int numVertices, const void* h_constants,
                             size_t constantsBytes, void* stream);

/* ---- multi-GPU composite over NVLink peer memory -----------------------------------
 * One process per GPU.  The display process allocates the frame slots of ALL ranks in its own memory and exports them
 * (CUDA IPC); every other process maps them and passes its slot to crb_set_surfaces as the colour surface: the fine
 * raster's colour stores then travel over NVLink / NVSwitch as the tiles are finished -- the composite is the render
 * itself, there is no gather step.  crb_ipc_signal leaves a stream-ordered 32-bit mark (e.g. the frame number) in
 * shared memory for the consumer.  Handles are cudaIpcMemHandle_t (64 bytes). */
/* Layout of the COLOUR surface (single sample): 0 = the reference's row-major layout (crb_set_surfaces), 1 = tile-major --
 * tile t = tx + ty * ceil(width / 8) occupies texels [64 t, 64 t + 64), pixel (x, y) of the tile at y * 8 + x; same size.
 * A warp then writes its tile as two full 128-byte lines; meant for frame slots in a peer GPU's memory. */
int crb_set_color_layout(crb_ctx* ctx, int tileMajor);
/* Row pitch of the COLOUR surface in texels (single sample, row-major; 0 = the surface's own rounded width): lets the colour
 * pointer of crb_set_surfaces address a window inside a larger image, e.g. a sort-first rectangle inside the full frame in
 * the display GPU's memory -- every rank then renders its rectangles in place and the composite needs no paste.  The window
 * should be a multiple of 8 px wide and high (the kernels write whole tiles).  The host-buffer entries ignore it. */
int crb_set_color_pitch(crb_ctx* ctx, int pitchTexels);
#define CRB_IPC_HANDLE_BYTES 64
int crb_ipc_alloc(size_t bytes, void** d_ptr, unsigned char handle[CRB_IPC_HANDLE_BYTES]);
int crb_ipc_free(void* d_ptr);
int crb_ipc_open(const unsigned char handle[CRB_IPC_HANDLE_BYTES], void** d_ptr);
int crb_ipc_close(void* d_ptr);
int crb_ipc_signal(void* d_word, uint32_t value, void* stream);
/* Stream-ordered device-to-device copy (either side may be mapped peer memory): the DMA engines move a finished frame
 * into its slot with full-size NVLink packets. */
int crb_ipc_copy(void* d_dst, const void* d_src, size_t bytes, void* stream);
/* The same for a rectangle (rows of widthBytes bytes): a sort-first window rendered locally, pasted into the full frame in the
 * display GPU's memory by the DMA engines. */
int crb_ipc_copy_2d(void* d_dst, size_t dstPitchBytes, const void* d_src, size_t srcPitchBytes, size_t widthBytes, size_t height, void* stream);
/* Stream-ordered pause of `nanoseconds` (<= 0.1 s): rank r of a job that composites into one display GPU starts its frame loop
 * r / world of a frame time late, so that the ranks' frame pushes interleave at the display GPU's NVLink ingress. */
int crb_ipc_delay(void* stream, unsigned int nanoseconds);

#ifdef __cplusplus
}
#endif
#endif /* CRB200_H_ */
