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

#include <cudaraster/cuda/PixelPipe.inl>

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
