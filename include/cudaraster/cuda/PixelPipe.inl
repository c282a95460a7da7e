// Include this file from a pixel-pipe translation unit (compiled by nvcc for sm_100a), define
// the shader classes and instantiate CR_DEFINE_PIXEL_PIPE -- the same recipe as the reference
// (src/cudaraster/cuda/PixelPipe.inl:241-275, usage in test/shader/PassThrough.cu:16-67).
//
// The reference's macro emits four Fermi kernels that the host finds by name with
// cuModuleGetFunction.  Here it emits four extern "C" *stage launchers* with the same names
// (crb_stage_fn: they enqueue the sm_100a kernels of that stage on a stream) plus the
// <name>_spec record, so CudaRaster::setPixelPipe(module, name) / crb_set_pixel_pipe_by_name
// can still resolve a pipe from its string name with dlsym.  triangleSetup and fineRaster are
// instantiated for the pipe's vertex / shader classes; binRaster and coarseRaster do not depend
// on the pipe and forward to the kernels inside libcrb200.so.
#pragma once
#include <cstring>
#include "PixelPipe.hpp"
#include "TriangleSetup.cuh"
#include "FineRaster.cuh"
#include "FineRasterMSAA.cuh"

#ifndef CR_PROFILING_MODE
#define CR_PROFILING_MODE ProfilingMode_Default
#endif

extern "C" int crb_launch_bin_raster(const crb_frame* frame, void* stream);
extern "C" int crb_launch_coarse_raster(const crb_frame* frame, void* stream);

// Is the pipe ORDER INDEPENDENT (crb_pipe_desc.orderIndependent)?  needsDst() is device code (the reference's
// contract: "must be a constant", cuda/PixelPipe.hpp:112), so the answer is computed by a one-thread kernel, once.
namespace FW {
template <class FragmentShaderClass, class BlendShaderClass, int SamplesLog2, U32 RenderModeFlags>
static __global__ void pipeProbeKernel(int* out) {
    BlendShaderClass b;
    const bool quadsMsaa = (RenderModeFlags & RenderModeFlag_EnableQuads) != 0 && SamplesLog2 > 0;   // shades in order (FineRasterMSAA.cuh)
    *out = ((RenderModeFlags & RenderModeFlag_EnableDepth) != 0 && !b.needsDst() && FragmentShaderClass::CanDiscard == 0 && !quadsMsaa) ? 1 : 0;
}
template <class FragmentShaderClass, class BlendShaderClass, int SamplesLog2, U32 RenderModeFlags>
inline int probeOrderIndependent() {
    int* d = nullptr;
    int v = 0;
    if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) return 0;
    pipeProbeKernel<FragmentShaderClass, BlendShaderClass, SamplesLog2, RenderModeFlags><<<1, 1>>>(d);
    if (cudaMemcpy(&v, d, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) v = 0;
    cudaFree(d);
    return v;
}
}  // namespace FW

#define CR_DEFINE_PIXEL_PIPE(PIPE_NAME, VERTEX_STRUCT, FRAGMENT_SHADER, BLEND_SHADER, SAMPLES_LOG2, RENDER_MODE_FLAGS)                      \
    extern "C" int PIPE_NAME##_triangleSetup(const crb_frame* frame, void* stream) {                                                        \
        return FW::launchTriangleSetup<VERTEX_STRUCT, SAMPLES_LOG2, RENDER_MODE_FLAGS, CR_PROFILING_MODE>(frame, stream);                   \
    }                                                                                                                                       \
    extern "C" int PIPE_NAME##_binRaster(const crb_frame* frame, void* stream) { return crb_launch_bin_raster(frame, stream); }             \
    extern "C" int PIPE_NAME##_coarseRaster(const crb_frame* frame, void* stream) { return crb_launch_coarse_raster(frame, stream); }       \
    extern "C" int PIPE_NAME##_fineRaster(const crb_frame* frame, void* stream) {                                                           \
        return FW::FineRasterLauncher<VERTEX_STRUCT, FRAGMENT_SHADER, BLEND_SHADER, SAMPLES_LOG2, RENDER_MODE_FLAGS, CR_PROFILING_MODE>::launch(frame, stream); \
    }                                                                                                                                       \
    extern "C" int PIPE_NAME##_orderIndependent(void) {                                                                                     \
        static int cached = -1;                                                                                                             \
        if (cached < 0) cached = FW::probeOrderIndependent<FRAGMENT_SHADER, BLEND_SHADER, SAMPLES_LOG2, RENDER_MODE_FLAGS>();               \
        return cached;                                                                                                                      \
    }                                                                                                                                       \
    /* layout guard: the module and the library must have been built from the same headers (crb_frame travels by value) */                 \
    extern "C" int PIPE_NAME##_frameBytes(void) { return (int)sizeof(crb_frame); }                                                         \
    extern "C" const crb_pipe_spec PIPE_NAME##_spec = {                                                                                     \
        /* samplesLog2 */ SAMPLES_LOG2,                                                                                                     \
        /* vertexStructSize */ (int)sizeof(VERTEX_STRUCT),                                                                                  \
        /* renderModeFlags */ RENDER_MODE_FLAGS,                                                                                            \
        /* profilingMode */ CR_PROFILING_MODE,                                                                                              \
        /* blendShaderName */ #BLEND_SHADER,                                                                                                \
    };

// Vertex-shader stage (reference: the user kernel of test/shader/PassThrough.cu:16-35 and its launch,
// test/SceneCR.cpp:263-282).  SHADER is a functor type with
//     __device__ void operator()(const INPUT_VERTEX& in, SHADED_VERTEX& out, const CONSTANTS& c, int vertexIdx) const;
// The macro emits the kernel (one thread per vertex, 128-thread CTAs like the reference's 32x4 blocks) and
// `NAME_launch` (crb_vertex_shader_fn, include/crb200.h), found by name like the pipe's stage launchers.
namespace FW {
template <class InputVertex, class ShadedVertex, class Constants, class Shader>
static __global__ void __launch_bounds__(128) vertexShaderKernel(const InputVertex* __restrict__ in, ShadedVertex* __restrict__ out, int numVertices, const __grid_constant__ Constants c) {
    gridDepLaunchDependents();
    gridDepWait();   // frames in flight may still read the output buffer
    const int v = blockIdx.x * 128 + threadIdx.x;
    if (v >= numVertices) return;
    ShadedVertex o;
    Shader()(in[v], o, c, v);
    out[v] = o;
}

template <class InputVertex, class ShadedVertex, class Constants, class Shader>
inline int launchVertexShader(const void* in, void* out, int numVertices, const void* constants, size_t constantsBytes, void* stream) {
    static_assert(sizeof(Constants) <= CRB_MAX_VS_CONSTANTS, "vertex shader constants must fit a kernel argument");
    if (numVertices <= 0) return CRB_OK;
    if (!in || !out || !constants || constantsBytes != sizeof(Constants)) return CRB_ERR_INVALID;
    Constants c;
    memcpy(&c, constants, sizeof(Constants));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((numVertices + 127) / 128));
    cfg.blockDim = dim3(128);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, vertexShaderKernel<InputVertex, ShadedVertex, Constants, Shader>, (const InputVertex*)in, (ShadedVertex*)out, numVertices, c) == cudaSuccess ? CRB_OK
                                                                                                                                                                                 : CRB_ERR_CUDA;
}
}  // namespace FW

#define CR_DEFINE_VERTEX_SHADER(NAME, INPUT_VERTEX, SHADED_VERTEX, CONSTANTS, SHADER)                                                                    \
    extern "C" int NAME##_launch(const void* in, void* out, int numVertices, const void* constants, size_t constantsBytes, void* stream) {               \
        return FW::launchVertexShader<INPUT_VERTEX, SHADED_VERTEX, CONSTANTS, SHADER>(in, out, numVertices, constants, constantsBytes, stream);          \
    }
