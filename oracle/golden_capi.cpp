// C ABI over golden.hpp for ctypes (tests, smoke, bench cpu_baseline).  TEST INFRASTRUCTURE.
#include "golden.hpp"

#include <chrono>

using namespace gold;

extern "C" {

int gold_abi_version(void) { return 6; }
int gold_sizeof_config(void) { return (int)sizeof(Config); }
int gold_sizeof_counts(void) { return (int)sizeof(Counts); }

// ---- helper KATs ---------------------------------------------------------------------------------
unsigned gold_select_flips(int dx, int dy) { return selectFlips(dx, dy); }
int gold_msaa_centroid(int samplesLog2, unsigned mask) { return msaaCentroid(samplesLog2, mask); }
unsigned gold_centroid_code(int samplesLog2, unsigned mask) { return centroidCode(samplesLog2, mask); }
unsigned gold_encode_depth(unsigned d) { return encodeDepth(d); }
unsigned gold_clear_depth(float d) { return clearDepthFromFloat(d); }
unsigned gold_to_abgr(float r, float g, float b, float a) { return toABGR(r, g, b, a); }
unsigned gold_blend(int blend, unsigned src, unsigned dst) {
    U32 out = dst;
    if (!runBlend(blend, src, dst, out)) return dst;
    return out;
}
int gold_msaa_x(int samplesLog2, int i) { return kMsaaX[samplesLog2][i]; }

// clip one triangle against the reference frustum; bary out = 9 x (u,v); returns vertex count
int gold_clip_triangle(const float* v0, const float* v1, const float* v2, float* baryOut) {
    F32 d1[4], d2[4];
    for (int k = 0; k < 4; k++) d1[k] = v1[k] - v0[k], d2[k] = v2[k] - v0[k];
    const F32 lo[3] = {-1.0f, -1.0f, -1.0f}, hi[3] = {1.0f, 1.0f, 1.0f};
    B2 b[9];
    int n = clipTriangle(b, v0, v1, v2, d1, d2, lo, hi);
    for (int i = 0; i < n; i++) baryOut[2 * i] = b[i].u, baryOut[2 * i + 1] = b[i].v;
    return n;
}

void gold_setup_pleq(const float* values, const int* v0, const int* d1, const int* d2, float areaRcp, int samplesLog2, unsigned* out) {
    U3 p = setupPleq(values[0], values[1], values[2], I2{v0[0], v0[1]}, I2{d1[0], d1[1]}, I2{d2[0], d2[1]}, areaRcp, samplesLog2);
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}

// 8x8 pixel-centre coverage of tile (tileX,tileY) for a header in a width x height viewport
unsigned long long gold_cover_tile(const void* header, int width, int height, int tileX, int tileY) {
    Config c{};
    c.width = c.vpWidth = width; c.height = c.vpHeight = height;
    return coverTile(c, *(const TriHeader*)header, tileX, tileY);
}
unsigned gold_cover_samples(const void* header, int width, int height, int samplesLog2, int px, int py) {
    Config c{};
    c.width = c.vpWidth = width; c.height = c.vpHeight = height; c.samplesLog2 = samplesLog2;
    Edges e;
    edgesFromHeader(*(const TriHeader*)header, e);
    return coverPixelSamples(c, e, px, py);
}

// One fragment-shader run (FineRaster.inl:49-119): colour + centre / centroid barycentrics.  Returns 1 when discarded.
int gold_run_shader(const Config* c, const void* verts, const void* triData, int dataIdx, int px, int py, unsigned centroid, unsigned* color, float* bary6) {
    const TriData& d = ((const TriData*)triData)[dataIdx];
    const int S = c->samplesLog2;
    if (bary6) {
        Bary a = computeBary(d, (px * 2 + 1) << S, (py * 2 + 1) << S);
        Bary b = S == 0 ? a : computeBary(d, (px << (S + 1)) + (S32)(centroid & 0xF), (py << (S + 1)) + (S32)(centroid >> 4));
        bary6[0] = a.b0; bary6[1] = a.b1; bary6[2] = a.b2; bary6[3] = b.b0; bary6[4] = b.b1; bary6[5] = b.b2;
    }
    U32 col = 0;
    bool keep = runShader(*c, verts, d, dataIdx, px, py, centroid, col);
    *color = col;
    return keep ? 0 : 1;
}

// ---- stages -------------------------------------------------------------------------------------
int gold_triangle_setup(const Config* c, const void* verts, const int* indices, int numTris, unsigned char* triSubtris, void* triHeader,
                        void* triData, int maxSubtris) {
    return triangleSetup(*c, verts, indices, numTris, triSubtris, (TriHeader*)triHeader, (TriData*)triData, maxSubtris);
}

int gold_render(const Config* c, const void* verts, const int* indices, int numTris, unsigned* color, unsigned* depth, Counts* counts) {
    return renderFrame(*c, verts, indices, numTris, color, depth, counts);
}

// Times `reps` frames; returns the median seconds per frame.
double gold_time_render(const Config* c, const void* verts, const int* indices, int numTris, unsigned* color, unsigned* depth, int reps) {
    std::vector<double> t;
    for (int i = 0; i < reps; i++) {
        auto a = std::chrono::steady_clock::now();
        renderFrame(*c, verts, indices, numTris, color, depth, nullptr);
        auto b = std::chrono::steady_clock::now();
        t.push_back(std::chrono::duration<double>(b - a).count());
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

void gold_resolve(const unsigned* src, int width, int height, int numSamples, unsigned* dst, int dstPitch, int flipY) {
    resolveSurface(src, width, height, numSamples, dst, dstPitch, flipY != 0);
}

// vertex shader of the demo: in = numVertices x inStrideFloats floats (modelPos first), out = numVertices x outStrideFloats floats
// (clipPos first); floats 3.. of the input are copied behind clipPos (the "colour carried through" variant)
void gold_vertex_shader(const float* matrix, const float* in, int inStrideFloats, float* out, int outStrideFloats, int numVertices) {
    for (int v = 0; v < numVertices; v++) {
        transformPoint(matrix, in + (size_t)v * inStrideFloats, out + (size_t)v * outStrideFloats);
        for (int k = 4; k < outStrideFloats; k++) out[(size_t)v * outStrideFloats + k] = (k - 1 < inStrideFloats) ? in[(size_t)v * inStrideFloats + k - 1] : 0.0f;
    }
}

int gold_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
