// A USER pixel-pipe translation unit, written exactly the way the reference's demo writes one
// (test/shader/PassThrough.cu:16-67): include PixelPipe.inl, define the vertex struct, a vertex
// shader kernel, a fragment shader class, and instantiate CR_DEFINE_PIXEL_PIPE.  Built into
// libuserpipes.so by examples/cpp/Makefile; FW::CudaRaster::setPixelPipe(&module, "PixelPipe_user")
// finds the stage entry points by name.
#include <cudaraster/cuda/PixelPipe.inl>

using namespace FW;

struct UserConstants { Mat4f posToClip; };
__constant__ UserConstants c_user;

struct InputVertex { Vec3f modelPos; };
typedef GouraudVertex ShadedVertex_user;   // clipPos + one varying (colour)

// clipPos = posToClip * (modelPos, 1); colour from the position (test/shader/PassThrough.cu:16-35).
extern "C" __global__ void vertexShader_user(const InputVertex* in, ShadedVertex_user* out, int numVertices) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numVertices) return;
    const Vec3f p = in[i].modelPos;
    out[i].clipPos = c_user.posToClip * Vec4f(p, 1.0f);
    out[i].color = Vec4f(p.x * 0.5f + 0.5f, p.y * 0.5f + 0.5f, p.z * 0.5f + 0.5f, 1.0f);
}

// Host entry the demo calls instead of cuLaunchGrid on a kernel found by name (test/SceneCR.cpp:275-282).
extern "C" int userLaunchVertexShader(const float* posToClipColumnMajor, const void* d_in, void* d_out, int numVertices, void* stream) {
    UserConstants c;
    for (int col = 0; col < 4; col++)
        for (int row = 0; row < 4; row++) c.posToClip.m[col][row] = posToClipColumnMajor[col * 4 + row];
    if (cudaMemcpyToSymbolAsync(c_user, &c, sizeof(c), 0, cudaMemcpyHostToDevice, (cudaStream_t)stream) != cudaSuccess) return 1;
    vertexShader_user<<<(numVertices + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const InputVertex*)d_in, (ShadedVertex_user*)d_out, numVertices);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

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
