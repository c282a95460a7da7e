// Pixel pipes compiled into libcrb200.so: the plugin instances the benchmark configurations need
// (SURVEY.md 2 #17, 8d).  Each is an ordinary CR_DEFINE_PIXEL_PIPE instantiation, exactly what a
// user translation unit would write (reference: test/shader/PassThrough.cu:44-67,
// test/shader/Shaders.cu:120-205).  The reference compiles one variant at run time from -D
// defines (test/SceneCR.cpp:170-177); here the variants are precompiled and the variant is part
// of the name:  <base>_s<samplesLog2>_f<renderModeFlags>_<blend>.
#include "../../include/cudaraster/cuda/PixelPipe.inl"

using namespace FW;

// ---- vertex formats ------------------------------------------------------------------------------
typedef ShadedVertexBase ShadedVertex_passthrough;  // test/shader/PassThrough.hpp:37
typedef GouraudVertex ShadedVertex_gouraud;         // test/shader/Shaders.hpp:80
struct ShadedVertex_texPhong : ShadedVertexBase {   // test/shader/Shaders.hpp:86-91
    Vec4f cameraPos;     // varying 0
    Vec4f cameraNormal;  // varying 1
    Vec4f texCoord;      // varying 2
};

// ---- fragment shaders ----------------------------------------------------------------------------
// Constant red (test/shader/PassThrough.cu:44-56).
class FragmentShader_passthrough : public FragmentShaderBase {
public:
    enum { CanDiscard = 0 };
    __device__ __forceinline__ void run(void) { m_color = toABGR(Vec4f(1.0f, 0.0f, 0.0f, 1.0f)); }
};

typedef GouraudShader FragmentShader_gouraud;

// Gouraud with an alpha test: exercises m_discard and therefore the in-order shading path.
class FragmentShader_gouraudDiscard : public FragmentShaderBase {
public:
    __device__ __forceinline__ void run(void) {
        const Vec4f c = interpolateVarying(0, m_centroid);
        if (c.w < 0.5f) { m_discard = true; return; }
        m_color = toABGR(c);
    }
};

// Screen-space derivatives of the interpolated colour (RenderModeFlag_EnableQuads; exercises dFdx / dFdy):
// colour = (8|dFdx c.r| + 8|dFdy c.r|, 8|dFdx c.g| + 8|dFdy c.g|, c.b, c.a).  Oracle: runShaderQuads.
class FragmentShader_gouraudQuads : public FragmentShaderBase {
public:
    enum { CanDiscard = 0 };
    __device__ __forceinline__ void run(void) {
        const Vec4f c = interpolateVarying(0, m_centroid);
        const F32 dxr = dFdx(c.x), dyr = dFdy(c.x), dxg = dFdx(c.y), dyg = dFdy(c.y);
        m_color = toABGR(Vec4f(__fmaf_rn(fabsf(dyr), 8.0f, __fmul_rn(fabsf(dxr), 8.0f)), __fmaf_rn(fabsf(dyg), 8.0f, __fmul_rn(fabsf(dxg), 8.0f)), c.z, c.w));
    }
};

// Phong lighting of test/shader/Shaders.cu:37-51 with a PROCEDURAL checker texture: the
// reference's textured variant samples a texture atlas asset that is not in the tree.  All
// arithmetic is spelled out in IEEE single operations so that the CPU oracle
// (oracle/golden.hpp: phongProc) reproduces it bit for bit.
class FragmentShader_texPhong : public FragmentShaderBase {
public:
    enum { CanDiscard = 0 };
    static __device__ __forceinline__ F32 dot3(const Vec4f& a, const Vec4f& b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, __fmul_rn(a.x, b.x))); }
    __device__ __forceinline__ void run(void) {
        const Vec4f P = interpolateVarying(0, m_centroid);
        const Vec4f Nn = interpolateVarying(1, m_centroid);
        const Vec4f T = interpolateVarying(2, m_centroid);
        const F32 il = __frcp_rn(__fsqrt_rn(dot3(P, P)));
        const F32 nl = __frcp_rn(__fsqrt_rn(dot3(Nn, Nn)));
        const Vec4f I(__fmul_rn(P.x, il), __fmul_rn(P.y, il), __fmul_rn(P.z, il), 0.0f);
        const Vec4f N(__fmul_rn(Nn.x, nl), __fmul_rn(Nn.y, nl), __fmul_rn(Nn.z, nl), 0.0f);
        const F32 dIN = dot3(I, N);
        const F32 k = __fmul_rn(dIN, 2.0f);
        const Vec4f R(__fmaf_rn(-N.x, k, I.x), __fmaf_rn(-N.y, k, I.y), __fmaf_rn(-N.z, k, I.z), 0.0f);
        const F32 diffuse = __fmaf_rn(fmaxf(-dIN, 0.0f), 0.75f, 0.25f);
        F32 sp = fmaxf(-dot3(I, R), 0.0f);
        sp = __fmul_rn(sp, sp); sp = __fmul_rn(sp, sp); sp = __fmul_rn(sp, sp); sp = __fmul_rn(sp, sp);  // glossiness 16
        const int cu = (int)floorf(__fmul_rn(T.x, 16.0f)), cv = (int)floorf(__fmul_rn(T.y, 16.0f));
        const bool odd = ((cu ^ cv) & 1) != 0;
        const F32 r = odd ? 0.9f : 0.2f, g = odd ? 0.6f : 0.5f, b = odd ? 0.3f : 0.8f;
        m_color = toABGR(Vec4f(__fmaf_rn(sp, 0.5f, __fmul_rn(diffuse, r)), __fmaf_rn(sp, 0.5f, __fmul_rn(diffuse, g)), __fmaf_rn(sp, 0.5f, __fmul_rn(diffuse, b)), 1.0f));
    }
};

// ---- pipes ---------------------------------------------------------------------------------------
#define CRB_PIPE(BASE, VTX, FS, BLEND, S, F) CR_DEFINE_PIXEL_PIPE(PixelPipe_##BASE##_s##S##_f##F##_##BLEND, VTX, FS, BLEND, S, F)

// The demo's default pipe (test/App.hpp:48-55: depth on, lerp off, no blend, 1 sample).
CR_DEFINE_PIXEL_PIPE(PixelPipe_passthrough, ShadedVertex_passthrough, FragmentShader_passthrough, BlendReplace, 0, 1)

#define CRB_PASSTHROUGH(S, F, BLEND) CRB_PIPE(passthrough, ShadedVertex_passthrough, FragmentShader_passthrough, BLEND, S, F)
#define CRB_GOURAUD(S, F, BLEND) CRB_PIPE(gouraud, ShadedVertex_gouraud, FragmentShader_gouraud, BLEND, S, F)
#define CRB_GOURAUD_DISCARD(S, F, BLEND) CRB_PIPE(gouraudDiscard, ShadedVertex_gouraud, FragmentShader_gouraudDiscard, BLEND, S, F)
#define CRB_GOURAUD_QUADS(S, F, BLEND) CRB_PIPE(gouraudQuads, ShadedVertex_gouraud, FragmentShader_gouraudQuads, BLEND, S, F)
#define CRB_TEXPHONG(S, F, BLEND) CRB_PIPE(texPhong, ShadedVertex_texPhong, FragmentShader_texPhong, BLEND, S, F)

CRB_PASSTHROUGH(0, 0, BlendReplace)
CRB_PASSTHROUGH(0, 1, BlendReplace)
CRB_PASSTHROUGH(0, 1, BlendDepthOnly)
CRB_PASSTHROUGH(1, 1, BlendReplace)
CRB_PASSTHROUGH(2, 1, BlendReplace)
CRB_PASSTHROUGH(3, 1, BlendReplace)

CRB_GOURAUD(0, 0, BlendReplace)
CRB_GOURAUD(0, 1, BlendReplace)
CRB_GOURAUD(0, 2, BlendReplace)
CRB_GOURAUD(0, 3, BlendReplace)
CRB_GOURAUD(0, 2, BlendSrcOver)
CRB_GOURAUD(0, 3, BlendSrcOver)
CRB_GOURAUD(0, 3, BlendAdditive)
CRB_GOURAUD(1, 3, BlendReplace)
CRB_GOURAUD(2, 3, BlendReplace)
CRB_GOURAUD(3, 3, BlendReplace)
CRB_GOURAUD(2, 3, BlendSrcOver)
CRB_GOURAUD(2, 2, BlendSrcOver)

CRB_GOURAUD_DISCARD(0, 3, BlendReplace)
CRB_GOURAUD_DISCARD(2, 3, BlendReplace)

// RenderModeFlag_EnableQuads (flags bit 2): visibility-first and in-order single-sample paths, MSAA, and a
// discarding shader whose helper lanes must keep running
CRB_GOURAUD_QUADS(0, 7, BlendReplace)
CRB_GOURAUD_QUADS(0, 7, BlendSrcOver)
CRB_GOURAUD_QUADS(0, 6, BlendSrcOver)
CRB_GOURAUD_QUADS(2, 7, BlendReplace)
CRB_GOURAUD_QUADS(1, 7, BlendSrcOver)
CRB_GOURAUD_DISCARD(0, 7, BlendReplace)

CRB_TEXPHONG(0, 3, BlendReplace)
CRB_TEXPHONG(2, 3, BlendReplace)

// ---- ProfilingMode_Counters variants (reference: -DCR_PROFILING_MODE=ProfilingMode_Counters, test/SceneCR.cpp:170-177) ----
#undef CR_PROFILING_MODE
#define CR_PROFILING_MODE ProfilingMode_Counters
CRB_PIPE(gouraudCounters, ShadedVertex_gouraud, FragmentShader_gouraud, BlendReplace, 0, 3)
CRB_PIPE(gouraudCounters, ShadedVertex_gouraud, FragmentShader_gouraud, BlendReplace, 2, 3)
#undef CR_PROFILING_MODE
#define CR_PROFILING_MODE ProfilingMode_Timers
CRB_PIPE(gouraudTimers, ShadedVertex_gouraud, FragmentShader_gouraud, BlendReplace, 0, 3)
#undef CR_PROFILING_MODE
#define CR_PROFILING_MODE ProfilingMode_Default

// ---- vertex shaders ----------------------------------------------------------------
// The demo's pass-through vertex shader (test/shader/PassThrough.cu:16-35): clipPos = posToClip * (modelPos, 1).
// The matrix-vector product is spelled as the fma chain nvcc contracts the reference's generic
// Matrix::operator* into (framework/base/Math.hpp: r[i] += m(i,j) * v[j], j ascending), so the CPU oracle
// (oracle/golden.hpp: transformPoint) reproduces it bit for bit.
struct VsConstants_passthrough { Mat4f posToClip; };                  // test/shader/PassThrough.hpp:15-18
struct VsInputVertex_passthrough { Vec3f modelPos; };                 // test/shader/PassThrough.hpp:26-29
struct VsInputVertex_color { Vec3f modelPos; Vec4f color; };

static __device__ __forceinline__ Vec4f transformPoint(const Mat4f& m, const Vec3f& p) {
    Vec4f r;
    r.x = __fmaf_rn(m.m[3][0], 1.0f, __fmaf_rn(m.m[2][0], p.z, __fmaf_rn(m.m[1][0], p.y, __fmul_rn(m.m[0][0], p.x))));
    r.y = __fmaf_rn(m.m[3][1], 1.0f, __fmaf_rn(m.m[2][1], p.z, __fmaf_rn(m.m[1][1], p.y, __fmul_rn(m.m[0][1], p.x))));
    r.z = __fmaf_rn(m.m[3][2], 1.0f, __fmaf_rn(m.m[2][2], p.z, __fmaf_rn(m.m[1][2], p.y, __fmul_rn(m.m[0][2], p.x))));
    r.w = __fmaf_rn(m.m[3][3], 1.0f, __fmaf_rn(m.m[2][3], p.z, __fmaf_rn(m.m[1][3], p.y, __fmul_rn(m.m[0][3], p.x))));
    return r;
}

struct VertexShader_passthrough {
    __device__ __forceinline__ void operator()(const VsInputVertex_passthrough& in, ShadedVertex_passthrough& out, const VsConstants_passthrough& c, int) const {
        out.clipPos = transformPoint(c.posToClip, in.modelPos);
    }
};
// Transform + per-vertex colour carried through (the Gouraud pipe's vertex format).
struct VertexShader_color {
    __device__ __forceinline__ void operator()(const VsInputVertex_color& in, ShadedVertex_gouraud& out, const VsConstants_passthrough& c, int) const {
        out.clipPos = transformPoint(c.posToClip, in.modelPos);
        out.color = in.color;
    }
};

CR_DEFINE_VERTEX_SHADER(vertexShader_passthrough, VsInputVertex_passthrough, ShadedVertex_passthrough, VsConstants_passthrough, VertexShader_passthrough)
CR_DEFINE_VERTEX_SHADER(vertexShader_color, VsInputVertex_color, ShadedVertex_gouraud, VsConstants_passthrough, VertexShader_color)
