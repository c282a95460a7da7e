// A USER pixel-pipe translation unit, written exactly the way the reference's demo writes one
// (test/shader/PassThrough.cu:16-67): include PixelPipe.inl, define the vertex struct, a vertex
// shader kernel, a fragment shader class, and instantiate CR_DEFINE_PIXEL_PIPE.  Built into
// libuserpipes.so by examples/cpp/Makefile; FW::CudaRaster::setPixelPipe(&module, "PixelPipe_user")
// finds the stage entry points by name.
#include <cudaraster/cuda/PixelPipe.inl>

using namespace FW;

struct UserConstants { Mat4f posToClip; };   // the reference's c_constants block (test/shader/PassThrough.hpp:15-18); passed by value

struct InputVertex { Vec3f modelPos; };
typedef GouraudVertex ShadedVertex_user;   // clipPos + one varying (colour)

// clipPos = posToClip * (modelPos, 1); colour from the position (test/shader/PassThrough.cu:16-35).
struct VertexShader_user {
    __device__ __forceinline__ void operator()(const InputVertex& in, ShadedVertex_user& out, const UserConstants& c, int) const {
        const Vec3f p = in.modelPos;
        out.clipPos = c.posToClip * Vec4f(p, 1.0f);
        out.color = Vec4f(p.x * 0.5f + 0.5f, p.y * 0.5f + 0.5f, p.z * 0.5f + 0.5f, 1.0f);
    }
};

// Emits vertexShader_user_launch, which CudaModule::launchVertexShader("vertexShader_user", ...) finds by name
// (the reference finds the kernel by name and launches it with cuLaunchGrid, test/SceneCR.cpp:81, :275-282).
CR_DEFINE_VERTEX_SHADER(vertexShader_user, InputVertex, ShadedVertex_user, UserConstants, VertexShader_user)

#ifndef USER_STRIPE_SHIFT
#define USER_STRIPE_SHIFT 3   // a -D define of the run-time compiler (FW::CudaCompiler::define) changes the checker size
#endif

// Interpolated colour with a screen-space stripe pattern: exercises m_pixelPos and varyings.
class FragmentShader_user : public FragmentShaderBase {
public:
    enum { CanDiscard = 0 };
    __device__ __forceinline__ void run(void) {
        Vec4f c = interpolateVarying(0, m_centroid);
        if (((m_pixelPos.x >> USER_STRIPE_SHIFT) ^ (m_pixelPos.y >> USER_STRIPE_SHIFT)) & 1) c = Vec4f(c.x * 0.5f, c.y * 0.5f, c.z * 0.5f, 1.0f);
        m_color = toABGR(c);
    }
};

CR_DEFINE_PIXEL_PIPE(PixelPipe_user, ShadedVertex_user, FragmentShader_user, BlendReplace, 0, RenderModeFlag_EnableDepth | RenderModeFlag_EnableLerp)
