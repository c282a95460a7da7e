// Thin C wrapper that compiles the REFERENCE's own host-compilable helpers from where they lie
// under /root/reference (cuda/Util.hpp, base/Math.cpp) into oracle/_ref/libcrref_host.so.
// No reference source is copied into this repository.  TEST INFRASTRUCTURE: used by
// oracle/gen_golden_vectors.py to produce tests/golden/*.json and by the CPU tests (when
// /root/reference is present) to pin oracle/golden.hpp.
#include <cudaraster/cuda/Util.hpp>
#include <cudaraster/cuda/PrivateDefs.hpp>
#include <base/Math.hpp>

using namespace FW;

extern "C" {
unsigned ref_select_flips(int dx, int dy) { return cover8x8_selectFlips(dx, dy); }
int ref_msaa_centroid(int samplesLog2, unsigned mask) { return selectMSAACentroid(samplesLog2, mask); }
unsigned ref_encode_depth(unsigned d) { return encodeDepth(d); }
unsigned ref_decode_depth(unsigned d) { return decodeDepth(d); }
int ref_msaa_x(int samplesLog2, int i) { return c_msaaPatterns[samplesLog2][i]; }
unsigned ref_to_abgr(float r, float g, float b, float a) { return Vec4f(r, g, b, a).toABGR(); }
unsigned ref_clear_depth(float depth) { return encodeDepth((U32)min((U64)(depth * exp2(32)), (U64)FW_U32_MAX)); }
int ref_clip_triangle(const float* v0, const float* v1, const float* v2, float* baryOut) {
    float d1[4], d2[4];
    for (int k = 0; k < 4; k++) d1[k] = v1[k] - v0[k], d2[k] = v2[k] - v0[k];
    return clipTriangleWithFrustum(baryOut, v0, v1, v2, d1, d2);
}
int ref_sizeof_header(void) { return (int)sizeof(CRTriangleHeader); }
int ref_sizeof_data(void) { return (int)sizeof(CRTriangleData); }
int ref_sizeof_params(void) { return (int)sizeof(CRParams); }
unsigned ref_depth_min(void) { return CR_DEPTH_MIN; }
unsigned ref_depth_max(void) { return CR_DEPTH_MAX; }
int ref_bary_max(void) { return CR_BARY_MAX; }
}
