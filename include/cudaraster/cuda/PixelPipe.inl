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
#include "PixelPipe.hpp"
#include "TriangleSetup.cuh"
#include "FineRaster.cuh"
#include "FineRasterMSAA.cuh"

#ifndef CR_PROFILING_MODE
#define CR_PROFILING_MODE ProfilingMode_Default
#endif

extern "C" int crb_launch_bin_raster(const crb_frame* frame, void* stream);
extern "C" int crb_launch_coarse_raster(const crb_frame* frame, void* stream);

#define CR_DEFINE_PIXEL_PIPE(PIPE_NAME, VERTEX_STRUCT, FRAGMENT_SHADER, BLEND_SHADER, SAMPLES_LOG2, RENDER_MODE_FLAGS)                      \
    extern "C" int PIPE_NAME##_triangleSetup(const crb_frame* frame, void* stream) {                                                        \
        return FW::launchTriangleSetup<VERTEX_STRUCT, SAMPLES_LOG2, RENDER_MODE_FLAGS>(frame, stream);                                      \
    }                                                                                                                                       \
    extern "C" int PIPE_NAME##_binRaster(const crb_frame* frame, void* stream) { return crb_launch_bin_raster(frame, stream); }             \
    extern "C" int PIPE_NAME##_coarseRaster(const crb_frame* frame, void* stream) { return crb_launch_coarse_raster(frame, stream); }       \
    extern "C" int PIPE_NAME##_fineRaster(const crb_frame* frame, void* stream) {                                                           \
        return FW::FineRasterLauncher<VERTEX_STRUCT, FRAGMENT_SHADER, BLEND_SHADER, SAMPLES_LOG2, RENDER_MODE_FLAGS>::launch(frame, stream); \
    }                                                                                                                                       \
    extern "C" const crb_pipe_spec PIPE_NAME##_spec = {                                                                                     \
        /* samplesLog2 */ SAMPLES_LOG2,                                                                                                     \
        /* vertexStructSize */ (int)sizeof(VERTEX_STRUCT),                                                                                  \
        /* renderModeFlags */ RENDER_MODE_FLAGS,                                                                                            \
        /* profilingMode */ CR_PROFILING_MODE,                                                                                              \
        /* blendShaderName */ #BLEND_SHADER,                                                                                                \
    };
