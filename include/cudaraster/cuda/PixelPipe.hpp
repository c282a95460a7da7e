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
//   RenderModeFlag_EnableQuads   dFdx / dFdy are warp shuffles between the four lanes that own the pixels of
//                                a 2x2 quad (the reference goes through a shared-memory scratch row,
//                                cuda/PixelPipe.hpp:59-69); like there, they are only valid in quads mode.
#pragma once
#include "Util.cuh"

namespace FW {

enum {
    RenderModeFlag_EnableDepth = 1 << 0,  // depth test + depth write
    RenderModeFlag_EnableLerp = 1 << 1,   // varying interpolation
    RenderModeFlag_EnableQuads = 1 << 2,  // numerical derivatives in the fragment shader (dFdx / dFdy); degrades performance
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
    // Numerical derivatives (only valid when RenderModeFlag_EnableQuads is set; reference: cuda/PixelPipe.hpp:59-69).
    // In quads mode the shader runs converged on the four lanes that own the pixels of a 2x2 quad -- pixel
    // (x, y) of the tile row block sits on lane x + 8*(y & 3) -- so the neighbours are lanes ^1 and ^8.
    __device__ __forceinline__ F32 dFdx(F32 v) const {
        const int l = (int)(threadIdx.x & 31);
        const F32 hi = __shfl_sync(m_quadMask, v, l | 1), lo = __shfl_sync(m_quadMask, v, l & ~1);
        return __fsub_rn(hi, lo);
    }
    __device__ __forceinline__ F32 dFdy(F32 v) const {
        const int l = (int)(threadIdx.x & 31);
        const F32 hi = __shfl_sync(m_quadMask, v, l | 8), lo = __shfl_sync(m_quadMask, v, l & ~8);
        return __fsub_rn(hi, lo);
    }
    __device__ __forceinline__ Vec2f dFdx(const Vec2f& v) const { return Vec2f(dFdx(v.x), dFdx(v.y)); }
    __device__ __forceinline__ Vec2f dFdy(const Vec2f& v) const { return Vec2f(dFdy(v.x), dFdy(v.y)); }
    __device__ __forceinline__ Vec3f dFdx(const Vec3f& v) const { return Vec3f(dFdx(v.x), dFdx(v.y), dFdx(v.z)); }
    __device__ __forceinline__ Vec3f dFdy(const Vec3f& v) const { return Vec3f(dFdy(v.x), dFdy(v.y), dFdy(v.z)); }
    __device__ __forceinline__ Vec4f dFdx(const Vec4f& v) const { return Vec4f(dFdx(v.x), dFdx(v.y), dFdx(v.z), dFdx(v.w)); }
    __device__ __forceinline__ Vec4f dFdy(const Vec4f& v) const { return Vec4f(dFdy(v.x), dFdy(v.y), dFdy(v.z), dFdy(v.w)); }
    __device__ __forceinline__ void run(void) {}
#endif

public:
    // Inputs.
    U32 m_quadMask;      // lanes of this fragment's 2x2 quad (quads mode), else this lane only
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

// ProfilingMode_Default / _Counters / _Timers: PrivateDefs.hpp

}  // namespace FW
