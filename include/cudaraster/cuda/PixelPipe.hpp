// CudaRaster device-side public interface ("pixel pipe" plugin API) for the B200 pipeline.
//
// Same contract as the reference (src/cudaraster/cuda/PixelPipe.hpp:30-196): a user translation
// unit includes PixelPipe.inl, derives a vertex struct from ShadedVertexBase, a fragment shader
// from FragmentShaderBase and (optionally) a blend shader from BlendShaderBase, and instantiates
//   CR_DEFINE_PIXEL_PIPE(PipeName, VertexStruct, FragmentShader, BlendShader, SamplesLog2, RenderModeFlags)
// which yields the symbols PipeName_triangleSetup / _binRaster / _coarseRaster / _fineRaster and
// PipeName_spec that CudaRaster::setPixelPipe() looks up by name.  Here those four symbols are
// host-callable stage launchers (crb_stage_fn) instead of raw Fermi kernels.
//
// Additions (opt-in, safe defaults):
//   FragmentShader::CanDiscard   enum, 1 by default.  A shader that never sets m_discard may
//                                declare `enum { CanDiscard = 0 };` which lets the fine raster
//                                resolve visibility first and shade only the surviving fragment
//                                of every sample (identical output, far less shading).
// Not supported yet: RenderModeFlag_EnableQuads (dFdx/dFdy); launching such a pipe fails.
#pragma once
#include "Util.cuh"

namespace FW {

enum {
    RenderModeFlag_EnableDepth = 1 << 0,  // depth test + depth write
    RenderModeFlag_EnableLerp = 1 << 1,   // varying interpolation
    RenderModeFlag_EnableQuads = 1 << 2,  // numerical derivatives (unsupported here)
};

// Vertex layout: clipPos first, then one Vec4f per varying and nothing else.
struct ShadedVertexBase {
    Vec4f clipPos;
};

class FragmentShaderBase {
public:
    enum { CanDiscard = 1 };

#ifdef __CUDACC__
    __device__ __forceinline__ Vec4f getVaryingAtVertex(int varyingIdx, int vertIdx) const {
        const float4 t = __ldg(reinterpret_cast<const float4*>(m_vertexBuffer) + (size_t)vertIdx * (m_vertexBytes / (int)sizeof(Vec4f)) + varyingIdx + 1);
        return Vec4f(t.x, t.y, t.z, t.w);
    }
    // v0*b.x + v1*b.y + v2*b.z with the reference's rounding order: fma(v2, b.z, fma(v0, b.x, v1*b.y)).
    __device__ __forceinline__ Vec4f interpolateVarying(int varyingIdx, const Vec3f& bary) const {
        const Vec4f a = getVaryingAtVertex(varyingIdx, m_vertIdx.x);
        const Vec4f b = getVaryingAtVertex(varyingIdx, m_vertIdx.y);
        const Vec4f c = getVaryingAtVertex(varyingIdx, m_vertIdx.z);
        return Vec4f(__fmaf_rn(c.x, bary.z, __fmaf_rn(a.x, bary.x, __fmul_rn(b.x, bary.y))), __fmaf_rn(c.y, bary.z, __fmaf_rn(a.y, bary.x, __fmul_rn(b.y, bary.y))),
                     __fmaf_rn(c.z, bary.z, __fmaf_rn(a.z, bary.x, __fmul_rn(b.z, bary.y))), __fmaf_rn(c.w, bary.z, __fmaf_rn(a.w, bary.x, __fmul_rn(b.w, bary.y))));
    }
    __device__ __forceinline__ void run(void) {}
#endif

public:
    // Inputs.
    S32 m_triIdx;        // input triangle index
    Vec3i m_vertIdx;     // its three vertex indices
    Vec2i m_pixelPos;    // integer pixel position
    S32 m_vertexBytes;   // sizeof(vertex struct)
    const void* m_vertexBuffer;

    Vec3f m_center;      // barycentrics at the pixel centre, and their screen-space derivatives
    Vec3f m_centerDX;
    Vec3f m_centerDY;
    Vec3f m_centroid;    // barycentrics at the covered sample nearest the centre (MSAA)
    Vec3f m_centroidDX;
    Vec3f m_centroidDY;

    // Outputs.
    U32 m_color;         // ABGR_8888
    bool m_discard;      // true culls the fragment
};

class BlendShaderBase {
public:
#ifdef __CUDACC__
    __device__ __forceinline__ bool needsDst(void) { return true; }  // must be a constant
    __device__ __forceinline__ void run(void) {}
#endif
public:
    S32 m_triIdx;
    Vec2i m_pixelPos;
    S32 m_sampleIdx;
    U32 m_src;          // colour from the fragment shader
    U32 m_dst;          // colour in the framebuffer
    U32 m_color;        // out: blended colour
    bool m_writeColor;  // out: false disables the colour write
};

// ---- stock shaders (reference: cuda/PixelPipe.hpp:130-175, cuda/PixelPipe.inl:61-85) -----------
struct GouraudVertex : ShadedVertexBase {
    Vec4f color;  // varying 0
};

class GouraudShader : public FragmentShaderBase {
public:
    enum { CanDiscard = 0 };
#ifdef __CUDACC__
    __device__ __forceinline__ void run(void) { m_color = toABGR(interpolateVarying(0, m_centroid)); }
#endif
};

class BlendReplace : public BlendShaderBase {  // dst = src
public:
#ifdef __CUDACC__
    __device__ __forceinline__ bool needsDst(void) { return false; }
    __device__ __forceinline__ void run(void) { m_color = m_src; }
#endif
};

class BlendSrcOver : public BlendShaderBase {  // dst = lerp(dst, src, src.a)
public:
#ifdef __CUDACC__
    __device__ __forceinline__ void run(void) { m_color = blendABGR(m_src, m_dst, m_src, ~m_src, m_src, ~m_src); }
#endif
};

class BlendAdditive : public BlendShaderBase {  // dst += src
public:
#ifdef __CUDACC__
    __device__ __forceinline__ void run(void) { m_color = blendABGRClamp(m_src, m_dst, ~0u, ~0u, ~0u, ~0u); }
#endif
};

class BlendDepthOnly : public BlendShaderBase {  // dst = dst
public:
#ifdef __CUDACC__
    __device__ __forceinline__ bool needsDst(void) { return false; }
    __device__ __forceinline__ void run(void) { m_writeColor = false; }
#endif
};

#define ProfilingMode_Default 0
#define ProfilingMode_Counters 1   // accepted for source compatibility; reports like Default
#define ProfilingMode_Timers 2

}  // namespace FW
