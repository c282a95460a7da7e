// Headless driver for the REFERENCE's own CUDA kernels, rebuilt for sm_100a.  TEST INFRASTRUCTURE.
//
// The reference device sources are compiled from where they lie under /root/reference
// (-I /root/reference/src; nothing is copied into this repository) behind ref_kernels/shim.h.  This
// file replaces what src/framework/gpu/CudaModule.cpp + src/cudaraster/CudaRaster.cpp do on the host:
// fills c_crParams / g_crAtomics, points the texture/surface shims at linear memory, launches the four
// kernels with the reference's launch shapes (CudaRaster.cpp:593-655) and its buffer sizing +
// overflow-retry policy (:264-339), and times the stages with the reference's five events.
//
// Used (a) as a second oracle for the parity tests and (b) as the "reference kernels rebuilt for
// B200" timing next to the new pipeline.  Best effort: the Fermi code is implicitly
// warp-synchronous and may mis-execute on sm_100a; callers run it in a
// separate process with a timeout and report "does not run correctly" instead of a number.
#include "shim.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

// The reference chains inline-asm statements through the CARRY FLAG (add.cc / addc in cuda/Util.hpp:85-87,
// used by cover8x8_lookupMask, Util.inl:249-271) without marking them volatile.  nvcc 12.9 treats such
// statements as pure: it DELETES the add.cc whose register result is dead and reorders the rest, so the LUT
// coverage path of the reference is miscompiled by today's toolchain (seen in the PTX: the seven add.cc of
// cover8x8_lookupMask are gone, the addc read a stale flag; on the GPU 8 315 of 60 599 triangle/tile
// masks come out wrong).  Making every asm statement of the reference volatile restores the order the
// source spells out.  This is a toolchain shim like the texture/vote shims in shim.h, not an algorithm change.
#define asm asm volatile
#include <cudaraster/cuda/PixelPipe.inl>
#undef asm

// ---- the reference's own test shaders (test/shader/PassThrough.cu:44-56 is re-stated here because
// that file also drags in the GLUT demo's constants; GouraudShader comes from PixelPipe.inl) -----
namespace FW {
class RefFragmentShader_passthrough : public FragmentShaderBase {
public:
    __device__ __inline__ void run(void) { m_color = toABGR(make_float4(1.0f, 0.0f, 0.0f, 1.0f)); }
};
}  // namespace FW

#define REF_PIPES(X)                                                             \
    X(ref_passthrough_s0_f1_BlendReplace, FW::ShadedVertexBase, FW::RefFragmentShader_passthrough, FW::BlendReplace, 0, 1) \
    X(ref_passthrough_s0_f0_BlendReplace, FW::ShadedVertexBase, FW::RefFragmentShader_passthrough, FW::BlendReplace, 0, 0) \
    X(ref_passthrough_s2_f1_BlendReplace, FW::ShadedVertexBase, FW::RefFragmentShader_passthrough, FW::BlendReplace, 2, 1) \
    X(ref_gouraud_s0_f3_BlendReplace, FW::GouraudVertex, FW::GouraudShader, FW::BlendReplace, 0, 3)                        \
    X(ref_gouraud_s0_f1_BlendReplace, FW::GouraudVertex, FW::GouraudShader, FW::BlendReplace, 0, 1)                        \
    X(ref_gouraud_s0_f3_BlendSrcOver, FW::GouraudVertex, FW::GouraudShader, FW::BlendSrcOver, 0, 3)                        \
    X(ref_gouraud_s1_f3_BlendReplace, FW::GouraudVertex, FW::GouraudShader, FW::BlendReplace, 1, 3)                        \
    X(ref_gouraud_s2_f3_BlendReplace, FW::GouraudVertex, FW::GouraudShader, FW::BlendReplace, 2, 3)                        \
    X(ref_gouraud_s3_f3_BlendReplace, FW::GouraudVertex, FW::GouraudShader, FW::BlendReplace, 3, 3)

using FW::PixelPipeSpec;  // the reference's macro names it unqualified
// nvcc's host stub registers the extern "C" module globals at global scope
// and PixelPipe.inl:27-37 only DECLARES them (extern "C" without initializer): define them here.
namespace FW {
extern "C" {
__constant__ FW::CRParams c_crParams;
__device__ FW::CRAtomics g_crAtomics;
__constant__ FW::S32 c_profLaunchIdx;
__constant__ CUdeviceptr c_profData;
__device__ cr_texture<float4, 1> t_vertexBuffer;
__device__ cr_texture<uint4, 1> t_triHeader;
__device__ cr_texture<uint4, 1> t_triData;
__device__ cr_surface<void, 2> s_colorBuffer;
__device__ cr_surface<void, 2> s_depthBuffer;
}
}  // namespace FW
using FW::c_crParams;
using FW::g_crAtomics;
using FW::c_profLaunchIdx;
using FW::c_profData;
using FW::t_vertexBuffer;
using FW::t_triHeader;
using FW::t_triData;
using FW::s_colorBuffer;
using FW::s_depthBuffer;
#define X(NAME, V, F, B, S, M) CR_DEFINE_PIXEL_PIPE(NAME, V, F, B, S, M)
REF_PIPES(X)
#undef X

typedef void (*KernelFn)(void);
struct RefPipe {
    const char* name;
    KernelFn setup, bin, coarse, fine;
    int samplesLog2, vertexBytes;
};
static const RefPipe g_pipes[] = {
#define X(NAME, V, F, B, S, M) {#NAME, NAME##_triangleSetup, NAME##_binRaster, NAME##_coarseRaster, NAME##_fineRaster, S, (int)sizeof(V)},
    REF_PIPES(X)
#undef X
};

static char g_err[512] = "";
#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            snprintf(g_err, sizeof(g_err), "%s: %s", #call, cudaGetErrorString(e_));                  \
            return 2;                                                                                 \
        }                                                                                             \
    } while (0)

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) cap = n;
        return e;
    }
};

static Buf b_triSubtris, b_triHeader, b_triData, b_binFirstSeg, b_binTotal, b_binSegData, b_binSegNext, b_binSegCount, b_activeTiles, b_tileFirstSeg,
    b_tileSegData, b_tileSegNext, b_tileSegCount;
static int g_maxSubtris = 1, g_maxBinSegs = 1, g_maxTileSegs = 1;
static cudaEvent_t g_ev[5];
static bool g_evInit = false;

extern "C" const char* crref_last_error(void) { return g_err; }
extern "C" int crref_num_pipes(void) { return (int)(sizeof(g_pipes) / sizeof(g_pipes[0])); }
extern "C" const char* crref_pipe_name(int i) { return g_pipes[i].name; }

// Renders one frame with the reference kernels.  Surfaces are linear U32 device memory with
// pitch = roundedWidth * numSamples texels (the layout the reference's CUarray surfaces have).
// stageSeconds[4] = setup, bin, coarse, fine (median is the caller's business); atomicsOut[7] = CRAtomics.
extern "C" int crref_draw(const char* pipeName, const void* d_verts, size_t vertBytes, const int* d_indices, int numTris, void* d_color, void* d_depth,
                          int width, int height, int numSamples, int deferredClear, unsigned clearColor, unsigned clearDepth, float* stageSeconds,
                          int* atomicsOut) {
    const RefPipe* pipe = nullptr;
    for (const RefPipe& p : g_pipes)
        if (std::strcmp(p.name, pipeName) == 0) pipe = &p;
    if (!pipe) {
        snprintf(g_err, sizeof(g_err), "unknown reference pipe %s", pipeName);
        return 1;
    }
    if ((1 << pipe->samplesLog2) != numSamples) {
        snprintf(g_err, sizeof(g_err), "sample count mismatch");
        return 1;
    }
    if (!g_evInit) {
        for (int i = 0; i < 5; i++) CK(cudaEventCreate(&g_ev[i]));
        g_evInit = true;
    }
    int dev = 0, numSMs = 1;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev));
    const int numFineWarps = CR_FINE_MAX_WARPS;

    // CudaRaster::setSurfaces (CudaRaster.cpp:161-169)
    const int wPix = (width + CR_TILE_SIZE - 1) & -CR_TILE_SIZE, hPix = (height + CR_TILE_SIZE - 1) & -CR_TILE_SIZE;
    const int wTiles = wPix >> CR_TILE_LOG2, hTiles = hPix >> CR_TILE_LOG2, numTiles = wTiles * hTiles;
    const int wBins = (wTiles + CR_BIN_SIZE - 1) >> CR_BIN_LOG2, hBins = (hTiles + CR_BIN_SIZE - 1) >> CR_BIN_LOG2, numBins = wBins * hBins;

    // CudaRaster::drawTriangles sizing (CudaRaster.cpp:239-275)
    const int roundSize = CR_BIN_WARPS * 32, minBatches = CR_BIN_STREAMS_SIZE * 2, maxRounds = 32;
    int batches = numTris / (roundSize * minBatches);
    batches = batches < 1 ? 1 : (batches > maxRounds ? maxRounds : batches);
    const int binBatchSize = batches * roundSize;
    g_maxSubtris = std::max(g_maxSubtris, numTris + 4096);
    g_maxBinSegs = std::max(g_maxBinSegs, std::max(numBins * CR_BIN_STREAMS_SIZE, (numTris - 1) / CR_BIN_SEG_SIZE + 1) + 256);
    g_maxTileSegs = std::max(g_maxTileSegs, std::max(numTiles, (numTris - 1) / CR_TILE_SEG_SIZE + 1) + 4096);

    CK(b_binFirstSeg.reserve(CR_MAXBINS_SQR * CR_BIN_STREAMS_SIZE * 4));
    CK(b_binTotal.reserve(CR_MAXBINS_SQR * CR_BIN_STREAMS_SIZE * 4));
    CK(b_activeTiles.reserve(CR_MAXTILES_SQR * 4));
    CK(b_tileFirstSeg.reserve(CR_MAXTILES_SQR * 4));

    FW::CRAtomics atomics;
    for (int attempt = 0; attempt < 8; attempt++) {
        if (g_maxSubtris > CR_MAXSUBTRIS_SIZE) {
            snprintf(g_err, sizeof(g_err), "CR_MAXSUBTRIS_SIZE exceeded");
            return 1;
        }
        CK(b_triSubtris.reserve((size_t)g_maxSubtris));
        CK(b_triHeader.reserve((size_t)g_maxSubtris * sizeof(FW::CRTriangleHeader)));
        CK(b_triData.reserve((size_t)g_maxSubtris * sizeof(FW::CRTriangleData)));
        CK(b_binSegData.reserve((size_t)g_maxBinSegs * CR_BIN_SEG_SIZE * 4));
        CK(b_binSegNext.reserve((size_t)g_maxBinSegs * 4));
        CK(b_binSegCount.reserve((size_t)g_maxBinSegs * 4));
        CK(b_tileSegData.reserve((size_t)g_maxTileSegs * CR_TILE_SEG_SIZE * 4));
        CK(b_tileSegNext.reserve((size_t)g_maxTileSegs * 4));
        CK(b_tileSegCount.reserve((size_t)g_maxTileSegs * 4));

        // CudaRaster::launchStages (CudaRaster.cpp:513-586)
        FW::CRParams p;
        std::memset(&p, 0, sizeof(p));
        p.numTris = numTris;
        p.vertexBuffer = (CUdeviceptr)d_verts;
        p.indexBuffer = (CUdeviceptr)d_indices;
        p.viewportWidth = width;   p.viewportHeight = height;
        p.widthPixels = wPix;      p.heightPixels = hPix;
        p.widthBins = wBins;       p.heightBins = hBins;   p.numBins = numBins;
        p.widthTiles = wTiles;     p.heightTiles = hTiles; p.numTiles = numTiles;
        p.binBatchSize = binBatchSize;
        p.deferredClear = deferredClear ? 1 : 0;
        p.clearColor = clearColor; p.clearDepth = clearDepth;
        p.maxSubtris = g_maxSubtris;
        p.triSubtris = (CUdeviceptr)b_triSubtris.p; p.triHeader = (CUdeviceptr)b_triHeader.p; p.triData = (CUdeviceptr)b_triData.p;
        p.maxBinSegs = g_maxBinSegs;
        p.binFirstSeg = (CUdeviceptr)b_binFirstSeg.p; p.binTotal = (CUdeviceptr)b_binTotal.p; p.binSegData = (CUdeviceptr)b_binSegData.p;
        p.binSegNext = (CUdeviceptr)b_binSegNext.p;   p.binSegCount = (CUdeviceptr)b_binSegCount.p;
        p.maxTileSegs = g_maxTileSegs;
        p.activeTiles = (CUdeviceptr)b_activeTiles.p; p.tileFirstSeg = (CUdeviceptr)b_tileFirstSeg.p; p.tileSegData = (CUdeviceptr)b_tileSegData.p;
        p.tileSegNext = (CUdeviceptr)b_tileSegNext.p; p.tileSegCount = (CUdeviceptr)b_tileSegCount.p;
        CK(cudaMemcpyToSymbol(c_crParams, &p, sizeof(p)));

        std::memset(&atomics, 0, sizeof(atomics));
        atomics.numSubtris = numTris;
        CK(cudaMemcpyToSymbol(g_crAtomics, &atomics, sizeof(atomics)));

        cr_texture<float4, 1> tv = {(const float4*)d_verts};
        cr_texture<uint4, 1> th = {(const uint4*)b_triHeader.p}, td = {(const uint4*)b_triData.p};
        cr_surface<void, 2> sc = {(unsigned char*)d_color, (unsigned)(wPix * numSamples * 4)}, sd = {(unsigned char*)d_depth, (unsigned)(wPix * numSamples * 4)};
        CK(cudaMemcpyToSymbol(t_vertexBuffer, &tv, sizeof(tv)));
        CK(cudaMemcpyToSymbol(t_triHeader, &th, sizeof(th)));
        CK(cudaMemcpyToSymbol(t_triData, &td, sizeof(td)));
        CK(cudaMemcpyToSymbol(s_colorBuffer, &sc, sizeof(sc)));
        CK(cudaMemcpyToSymbol(s_depthBuffer, &sd, sizeof(sd)));

        CK(cudaEventRecord(g_ev[0], 0));
        if (numTris > 0) {
            // grid folded to 2-D above 65535 blocks like CudaModule::selectGridSize (gpu/CudaModule.cpp:676-691)
            int blocks = (numTris - 1) / (CR_SETUP_WARPS * 32) + 1;
            dim3 grid(blocks, 1);
            if (blocks > 65535) {
                int gx = 65535;
                while (blocks % gx != 0 && gx > 32768) gx--;
                if (blocks % gx != 0) gx = 65535;
                grid = dim3(gx, (blocks + gx - 1) / gx);
            }
            ((void (*)(void))pipe->setup)<<<grid, dim3(32, CR_SETUP_WARPS)>>>();
        }
        CK(cudaEventRecord(g_ev[1], 0));
        ((void (*)(void))pipe->bin)<<<dim3(CR_BIN_STREAMS_SIZE, 1), dim3(32, CR_BIN_WARPS)>>>();
        CK(cudaEventRecord(g_ev[2], 0));
        ((void (*)(void))pipe->coarse)<<<dim3(numSMs, 1), dim3(32, CR_COARSE_WARPS)>>>();
        CK(cudaEventRecord(g_ev[3], 0));
        ((void (*)(void))pipe->fine)<<<dim3(numSMs, 1), dim3(32, numFineWarps)>>>();
        CK(cudaEventRecord(g_ev[4], 0));
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpyFromSymbol(&atomics, g_crAtomics, sizeof(atomics)));
        if (atomics.numSubtris <= g_maxSubtris && atomics.numBinSegs <= g_maxBinSegs && atomics.numTileSegs <= g_maxTileSegs) break;
        g_maxSubtris = std::max(g_maxSubtris, atomics.numSubtris + 4096);
        g_maxBinSegs = std::max(g_maxBinSegs, atomics.numBinSegs + 256);
        g_maxTileSegs = std::max(g_maxTileSegs, atomics.numTileSegs + 4096);
    }
    if (stageSeconds)
        for (int i = 0; i < 4; i++) {
            float ms = 0.0f;
            CK(cudaEventElapsedTime(&ms, g_ev[i], g_ev[i + 1]));
            stageSeconds[i] = ms * 1.0e-3f;
        }
    if (atomicsOut) std::memcpy(atomicsOut, &atomics, sizeof(atomics));
    return 0;
}

// Copies the reference's setup output to host arrays (numSubtris entries of header/data).
extern "C" int crref_get_setup_output(int numTris, int numSubtris, unsigned char* subtris, void* header, void* data) {
    CK(cudaMemcpy(subtris, b_triSubtris.p, (size_t)numTris, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(header, b_triHeader.p, (size_t)numSubtris * sizeof(FW::CRTriangleHeader), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(data, b_triData.p, (size_t)numSubtris * sizeof(FW::CRTriangleData), cudaMemcpyDeviceToHost));
    return 0;
}

//------------------------------------------------------------------------------------------------
// Per-thread harness around the reference's own DEVICE FUNCTIONS (no warp-synchronous code
// involved, so these do run correctly on sm_100a): the LUT coverage path the reference's fine raster
// really uses (FineRaster.inl:245-282 -> Util.inl:148-271), its MSAA sample coverage
// (FineRaster.inl:286-309 -> Util.inl:361-383), its fragment-shader front end (FineRaster.inl:49-119,
// PixelPipe.inl:43-69) and its blend shaders (PixelPipe.inl:73-85, Util.inl:42-60).  The parity tests
// compare the CPU oracle with these, value for value (tests/test_gpu_ref_kernels.py).
//------------------------------------------------------------------------------------------------

namespace {

// pairs[i] = {header index, tileX, tileY}; one thread per pair.  lutDump (optional) receives the 768 LUT words.
__global__ void refCoverTilesKernel(const uint4* headers, const int3* pairs, int n, FW::U64* exact, FW::U64* conservative, FW::U64* lutDump) {
    __shared__ volatile FW::U64 s_lut[CR_COVER8X8_LUT_SIZE];
    FW::cover8x8_setupLUT(s_lut);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (lutDump && blockIdx.x == 0)
        for (int k = threadIdx.x; k < CR_COVER8X8_LUT_SIZE; k += blockDim.x) lutDump[k] = s_lut[k];
    if (i >= n) return;
    const int3 p = pairs[i];
    const uint4 h = headers[p.x];
    exact[i] = FW::trianglePixelCoverage<0>(h, p.y, p.z, s_lut);
    conservative[i] = FW::trianglePixelCoverage<1>(h, p.y, p.z, s_lut);
}

// pairs[i] = {header index, pixelX, pixelY}
template <int S>
__global__ void refCoverSamplesKernel(const uint4* headers, const int3* pairs, int n, FW::U32* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int3 p = pairs[i];
    out[i] = FW::triangleSampleCoverage<S>(headers[p.x], p.y, p.z);
}

// frags[i] = {dataIdx, pixelX, pixelY, centroid code}; triData / vertices come through the texture shims.
template <class V, class FS, int S, FW::U32 Flags>
__global__ void refShadeKernel(const int4* frags, int n, FW::U32* colorOut, float* baryOut) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 q = frags[i];
    FS fs;
    FW::runFragmentShader<V, FS, S, Flags>(fs, q.x, q.x, q.y, q.z, (FW::U32)q.w, nullptr);
    colorOut[i] = fs.m_color;
    if (baryOut) {
        baryOut[i * 6 + 0] = fs.m_center.x;   baryOut[i * 6 + 1] = fs.m_center.y;   baryOut[i * 6 + 2] = fs.m_center.z;
        baryOut[i * 6 + 3] = fs.m_centroid.x; baryOut[i * 6 + 4] = fs.m_centroid.y; baryOut[i * 6 + 5] = fs.m_centroid.z;
    }
}

template <class B>
__global__ void refBlendKernel(const FW::U32* src, const FW::U32* dst, int n, FW::U32* out, int* write) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    B bs;
    FW::runBlendShader<B>(bs, i, i & 2047, i >> 11, i & 7, src[i], dst[i]);
    out[i] = bs.m_color;
    write[i] = bs.m_writeColor ? 1 : 0;
}

int setViewport(int width, int height) {
    FW::CRParams p;
    std::memset(&p, 0, sizeof(p));
    p.viewportWidth = width;
    p.viewportHeight = height;
    CK(cudaMemcpyToSymbol(c_crParams, &p, sizeof(p)));
    return 0;
}

}  // namespace

extern "C" int crref_cover_tiles(const void* d_headers, const int* d_pairs, int n, int width, int height, unsigned long long* d_exact,
                                 unsigned long long* d_conservative, unsigned long long* d_lutDump) {
    if (setViewport(width, height)) return 2;
    refCoverTilesKernel<<<(n + 255) / 256, 256>>>((const uint4*)d_headers, (const int3*)d_pairs, n, (FW::U64*)d_exact, (FW::U64*)d_conservative, (FW::U64*)d_lutDump);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}

extern "C" int crref_cover_samples(const void* d_headers, const int* d_pairs, int n, int width, int height, int samplesLog2, unsigned* d_out) {
    if (setViewport(width, height)) return 2;
    const int g = (n + 255) / 256;
    switch (samplesLog2) {
        case 1: refCoverSamplesKernel<1><<<g, 256>>>((const uint4*)d_headers, (const int3*)d_pairs, n, d_out); break;
        case 2: refCoverSamplesKernel<2><<<g, 256>>>((const uint4*)d_headers, (const int3*)d_pairs, n, d_out); break;
        case 3: refCoverSamplesKernel<3><<<g, 256>>>((const uint4*)d_headers, (const int3*)d_pairs, n, d_out); break;
        default: snprintf(g_err, sizeof(g_err), "samplesLog2 must be 1..3"); return 1;
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}

// Gouraud front end: the reference's runFragmentShader + GouraudShader for (samplesLog2, Depth|Lerp).
extern "C" int crref_shade_gouraud(const void* d_triData, const void* d_verts, const int* d_frags, int n, int samplesLog2, unsigned* d_color, float* d_bary) {
    cr_texture<float4, 1> tv = {(const float4*)d_verts};
    cr_texture<uint4, 1> td = {(const uint4*)d_triData};
    CK(cudaMemcpyToSymbol(t_vertexBuffer, &tv, sizeof(tv)));
    CK(cudaMemcpyToSymbol(t_triData, &td, sizeof(td)));
    FW::CRParams p;
    std::memset(&p, 0, sizeof(p));
    p.vertexBuffer = (CUdeviceptr)d_verts;
    CK(cudaMemcpyToSymbol(c_crParams, &p, sizeof(p)));
    const int g = (n + 127) / 128;
    switch (samplesLog2) {
        case 0: refShadeKernel<FW::GouraudVertex, FW::GouraudShader, 0, 3><<<g, 128>>>((const int4*)d_frags, n, d_color, d_bary); break;
        case 1: refShadeKernel<FW::GouraudVertex, FW::GouraudShader, 1, 3><<<g, 128>>>((const int4*)d_frags, n, d_color, d_bary); break;
        case 2: refShadeKernel<FW::GouraudVertex, FW::GouraudShader, 2, 3><<<g, 128>>>((const int4*)d_frags, n, d_color, d_bary); break;
        case 3: refShadeKernel<FW::GouraudVertex, FW::GouraudShader, 3, 3><<<g, 128>>>((const int4*)d_frags, n, d_color, d_bary); break;
        default: snprintf(g_err, sizeof(g_err), "samplesLog2 must be 0..3"); return 1;
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}

// blend: 0 Replace, 1 SrcOver, 2 Additive, 3 DepthOnly (the order of the oracle's blend ids)
extern "C" int crref_blend(int blend, const unsigned* d_src, const unsigned* d_dst, int n, unsigned* d_out, int* d_write) {
    const int g = (n + 255) / 256;
    switch (blend) {
        case 0: refBlendKernel<FW::BlendReplace><<<g, 256>>>(d_src, d_dst, n, d_out, d_write); break;
        case 1: refBlendKernel<FW::BlendSrcOver><<<g, 256>>>(d_src, d_dst, n, d_out, d_write); break;
        case 2: refBlendKernel<FW::BlendAdditive><<<g, 256>>>(d_src, d_dst, n, d_out, d_write); break;
        case 3: refBlendKernel<FW::BlendDepthOnly><<<g, 256>>>(d_src, d_dst, n, d_out, d_write); break;
        default: snprintf(g_err, sizeof(g_err), "unknown blend"); return 1;
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}
