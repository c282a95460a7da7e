// The reference's demo frame (test/App.cpp:208-241 cube, test/SceneCR.cpp:243-317 render) written
// against the kept host API (include/cudaraster/CudaRaster.hpp) -- no GL, no GLUT:
//   vertex shader kernel -> deferredClear -> setVertexBuffer/setIndexBuffer -> drawTriangles -> getStats.
// Usage: cube <libuserpipes.so | UserPipes.cu> <out.raw> [width height [stripeShift]]
// A .cu argument is compiled at run time with FW::CudaCompiler (cache next to out.raw), the way
// test/SceneCR.cpp:67-90, :170-179 compiles its shader file, with -DUSER_STRIPE_SHIFT=<stripeShift>.
// Writes width, height (int32) then the colour and depth surfaces (U32 each) to out.raw.
#include <cudaraster/CudaCompiler.hpp>
#include <cudaraster/CudaRaster.hpp>

#include <cmath>
#include <cstdio>
#include <vector>

using namespace FW;

struct UserConstants { Mat4f posToClip; };   // must match UserPipes.cu

static void mul(float* o, const float* a, const float* b) {  // column-major o = a * b
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) {
            float s = 0.0f;
            for (int k = 0; k < 4; k++) s += a[k * 4 + r] * b[c * 4 + k];
            o[c * 4 + r] = s;
        }
}

int main(int argc, char** argv) {
    if (argc < 3) { printf("usage: cube <libuserpipes.so> <out.raw> [w h]\n"); return 2; }
    const int w = argc > 4 ? atoi(argv[3]) : 1024, h = argc > 4 ? atoi(argv[4]) : 768;

    static const float pos[8][3] = {{-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1}, {1, -1, -1}, {1, -1, 1}, {1, 1, -1}, {1, 1, 1}};
    static const int idx[12][3] = {{7, 3, 1}, {7, 1, 5}, {7, 5, 6}, {6, 5, 4}, {6, 4, 2}, {2, 4, 0}, {2, 0, 3}, {3, 0, 1}, {3, 7, 6}, {3, 6, 2}, {5, 1, 0}, {5, 0, 4}};

    // perspective(60 deg, w/h, 0.1, 100) * lookAt((2,2,4) -> origin)
    const float f = 1.0f / tanf(30.0f * 3.14159265f / 180.0f), zn = 0.1f, zf = 100.0f, asp = (float)w / (float)h;
    const float proj[16] = {f / asp, 0, 0, 0, 0, f, 0, 0, 0, 0, -(zf + zn) / (zf - zn), -1, 0, 0, -2 * zf * zn / (zf - zn), 0};
    const float eye[3] = {2, 2, 4};
    float fw[3] = {-eye[0], -eye[1], -eye[2]};
    const float fl = sqrtf(fw[0] * fw[0] + fw[1] * fw[1] + fw[2] * fw[2]);
    for (float& v : fw) v /= fl;
    float s[3] = {fw[1] * 0 - fw[2] * 1, fw[2] * 0 - fw[0] * 0, fw[0] * 1 - fw[1] * 0};  // f x up(0,1,0)
    const float sl = sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    for (float& v : s) v /= sl;
    const float u[3] = {s[1] * fw[2] - s[2] * fw[1], s[2] * fw[0] - s[0] * fw[2], s[0] * fw[1] - s[1] * fw[0]};
    const float view[16] = {s[0], u[0], -fw[0], 0, s[1], u[1], -fw[1], 0, s[2], u[2], -fw[2], 0,
                            -(s[0] * eye[0] + s[1] * eye[1] + s[2] * eye[2]), -(u[0] * eye[0] + u[1] * eye[1] + u[2] * eye[2]), fw[0] * eye[0] + fw[1] * eye[1] + fw[2] * eye[2], 1};
    float mvp[16];
    mul(mvp, proj, view);

    CudaRaster cr;
    cr.init();
    CudaSurface color(Vec2i(w, h), CudaSurface::FORMAT_RGBA8), depth(Vec2i(w, h), CudaSurface::FORMAT_DEPTH32);
    const std::string arg1(argv[1]);
    CudaCompiler compiler;
    CudaModule* compiled = NULL;
    if (arg1.size() > 3 && arg1.substr(arg1.size() - 3) == ".cu") {
        const std::string outPath(argv[2]);
        const size_t slash = outPath.find_last_of('/');
        compiler.setCachePath((slash == std::string::npos ? std::string(".") : outPath.substr(0, slash)) + "/cudacache");
        compiler.setSourceFile(arg1);
        compiler.define("USER_STRIPE_SHIFT", argc > 5 ? atoi(argv[5]) : 3);
        compiled = compiler.compile();
    }
    CudaModule preBuilt(compiled ? "" : arg1);
    CudaModule& module = compiled ? *compiled : preBuilt;
    Buffer inVerts(pos, sizeof(pos)), indices(idx, sizeof(idx)), shadedVerts;
    shadedVerts.resizeDiscard(8 * 32);  // sizeof(GouraudVertex)
    UserConstants constants;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) constants.posToClip.m[c][r] = mvp[c * 4 + r];
    module.launchVertexShader("vertexShader_user", inVerts, shadedVerts, 8, constants);   // test/SceneCR.cpp:263-282

    cr.setSurfaces(&color, &depth);
    cr.setPixelPipe(&module, "PixelPipe_user");
    cr.deferredClear(Vec4f(0.2f, 0.4f, 0.8f, 1.0f));
    cr.setVertexBuffer(&shadedVerts, 0);
    cr.setIndexBuffer(&indices, 0, 12);
    cr.drawTriangles();
    const CudaRaster::Stats st = cr.getStats();
    printf("CudaRaster: setup = %.3f ms, bin = %.3f ms, coarse = %.3f ms, fine = %.3f ms\n", st.setupTime * 1e3f, st.binTime * 1e3f, st.coarseTime * 1e3f, st.fineTime * 1e3f);
    printf("%s", cr.getProfilingInfo().c_str());

    std::vector<U32> hc(color.getSizeBytes() / 4), hd(depth.getSizeBytes() / 4);
    color.download(hc.data());
    depth.download(hd.data());
    std::vector<float> hv(8 * 8);
    shadedVerts.getRange(hv.data(), 0, 8 * 32);
    FILE* fp = fopen(argv[2], "wb");
    if (!fp) fail("cube: cannot write %s", argv[2]);
    const int dims[2] = {color.getTextureSize().x, color.getTextureSize().y};
    fwrite(dims, 4, 2, fp);
    fwrite(hc.data(), 4, hc.size(), fp);
    fwrite(hd.data(), 4, hd.size(), fp);
    fwrite(hv.data(), 4, hv.size(), fp);
    fclose(fp);
    color.resolveToFile(std::string(argv[2]) + ".ppm");   // stands in for resolveToScreen (test/SceneCR.cpp:297)
    return 0;
}
