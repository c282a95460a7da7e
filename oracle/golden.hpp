// golden.hpp -- scalar CPU restatement of the CudaRaster pipeline (TEST INFRASTRUCTURE).
//
// This is the ORACLE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build, link or call it.  The product (cudaraster-linux_b200/csrc)
// never includes this file and has no CPU fallback.
//
// It restates, rule for rule, the DEVICE semantics of the reference pipeline (all paths are
// relative to /root/reference/src/cudaraster):
//   triangle setup    cuda/TriangleSetup.inl:19-417, cuda/Util.inl:65-91, cuda/Util.hpp:207-286
//   coverage          cuda/Util.inl:95-134 (8x8 exact), :340-357 (MSAA)
//   depth/shade/ROP   cuda/FineRaster.inl:21-119, :433-495, :686-726, :758-851, :1034-1111
//   blend arithmetic  cuda/Util.inl:30-60, CudaRaster.cpp:1358-1375
//   bin/tile overlap  CudaRaster.cpp:978-988, :1129-1143 (used only for the B_alg counters)
// Where the reference's host emulators (CudaRaster.cpp:669-1403) and its device code disagree
//  the DEVICE wins.  Float expressions are written with explicit fmaf()/single
// operations exactly where nvcc 12.9 contracts the reference's expressions (verified in the PTX of
// the reference kernels rebuilt with oracle/ref_kernels/shim.h); build with -ffp-contract=off.
//
// Parity pin: the reference ships no golden vectors.  The helpers below are pinned
// against the reference's own host-compilable functions (oracle/ref_host, fixtures in
// tests/golden/), and the whole model against the reference CUDA kernels rebuilt for sm_100a
// (oracle/ref_kernels) on the GPU box.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace gold {

typedef uint8_t U8;
typedef int16_t S16;
typedef uint32_t U32;
typedef int32_t S32;
typedef uint64_t U64;
typedef int64_t S64;
typedef float F32;

// ---- formats and limits (cuda/Constants.hpp:21-84, cuda/PrivateDefs.hpp:26-62) -------------
enum {
    kSubpixelLog2 = 4,
    kTileLog2 = 3,
    kBinLog2 = 4,
    kMaxViewportLog2 = 11,
    kTileSize = 8,
    kTilePixels = 64,
};
static const U32 kDepthMin = 2200u << 3;
static const U32 kDepthMax = 0xFFFFFFFFu - (2200u << 3);
static const S32 kBaryMax = (1 << (30 - kSubpixelLog2)) - 1;

enum { kFlagDepth = 1, kFlagLerp = 2, kFlagQuads = 4 };
enum { kShaderConstant = 0, kShaderGouraud = 1, kShaderPhongProc = 2, kShaderGouraudDiscard = 3, kShaderGouraudQuads = 4 };
enum { kBlendReplace = 0, kBlendSrcOver = 1, kBlendAdditive = 2, kBlendDepthOnly = 3 };

struct TriHeader {  // 16 B
    S16 v0x, v0y, v1x, v1y, v2x, v2y;
    U32 misc;
};
struct TriData {  // 64 B
    U32 zx, zy, zb, zslope;
    S32 wx, wy, wb;
    S32 ux, uy, ub;
    S32 vx, vy, vb;
    U32 vi0, vi1, vi2;
};
static_assert(sizeof(TriHeader) == 16 && sizeof(TriData) == 64, "format");

// c_msaaPatterns[log2 N][sampleY] = sampleX (cuda/Util.hpp:27-35)
static const int kMsaaX[4][8] = {
    {0, 0, 0, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0, 0, 0}, {1, 3, 0, 2, 0, 0, 0, 0}, {7, 2, 4, 0, 6, 3, 1, 5}};

// ---- PTX conversion semantics ----------------------------------------------------------------
static inline F32 asF32(U32 u) { F32 f; std::memcpy(&f, &u, 4); return f; }
static inline U32 asU32(F32 f) { U32 u; std::memcpy(&u, &f, 4); return u; }

// cvt.rni.sat.s32.f32: round to nearest even, clamp, NaN -> 0.
static inline S32 cvtRniSatS32(F32 a) {
    if (a != a) return 0;
    if (a >= 2147483648.0f) return 0x7FFFFFFF;
    if (a <= -2147483648.0f) return (S32)0x80000000;
    return (S32)std::nearbyintf(a);
}
// cvt.rni.sat.u32.f32
static inline U32 cvtRniSatU32(F32 a) {
    if (!(a > 0.0f)) return 0;  // NaN, negatives, zero
    if (a >= 4294967296.0f) return 0xFFFFFFFFu;
    return (U32)(S64)std::nearbyintf(a);
}
// cvt.rzi.u32.f32 (what nvcc emits for a C cast (U32)float; clamps, NaN -> 0)
static inline U32 cvtRziU32(F32 a) {
    if (!(a > 0.0f)) return 0;
    if (a >= 4294967296.0f) return 0xFFFFFFFFu;
    return (U32)(S64)a;
}
// PTX shr.s64 / shl with a (possibly out of range) 32-bit shift amount: clamps at the width.
static inline S64 shrS64(S64 v, S32 sh) {
    U32 s = (U32)sh;
    if (s > 63) s = 63;
    return v >> s;
}
static inline S32 wrapS32(S64 v) { return (S32)(U32)(U64)v; }
static inline S32 mulS32(S32 a, S32 b) { return wrapS32((S64)a * (S64)b); }

// ---- small helpers (cuda/Util.hpp:182-217, :290-298; cuda/Util.inl:30-60) --------------------
static inline int msaaCentroid(int samplesLog2, U32 mask) {
    int n = 1 << samplesLog2;
    if (mask == 0 || mask == (1u << n) - 1) return -1;
    int best = -1, bestDist = 0x7FFFFFFF;
    for (int i = 0; i < n; i++) {
        if (!(mask >> i & 1)) continue;
        int ax = kMsaaX[samplesLog2][i] * 2 + 1 - n, ay = i * 2 + 1 - n;
        int dist = ax * ax + ay * ay;
        if (dist < bestDist) best = i, bestDist = dist;
    }
    return best;
}

static inline U32 selectFlips(S32 dx, S32 dy) {
    const U32 FY = 1 << 2, FX = 1 << 3, SW = 1 << 4, CO = 1 << 5;
    U32 f = 0;
    if (dy > 0 || (dy == 0 && dx <= 0)) f ^= FX ^ FY ^ CO;
    if (dx > 0) f ^= FX ^ FY;
    S32 ax = dx >= 0 ? dx : -dx, ay = dy >= 0 ? dy : -dy;
    if (ax < ay) f ^= SW ^ FY;
    return f;
}

static inline U32 encodeDepth(U32 depth) {
    double v = (double)depth / (65536.0 * 65536.0 - 1.0);
    v = std::min(std::max(v, 0.0), 1.0);
    v = v * (double)(kDepthMax - kDepthMin) + (double)kDepthMin;
    return (U32)v;
}
static inline U32 clearDepthFromFloat(F32 depth) {  // CudaRaster.cpp:174-179
    double d = (double)depth * 4294967296.0;
    U64 q = d <= 0.0 ? 0 : (d >= 18446744073709551615.0 ? ~(U64)0 : (U64)d);
    return encodeDepth((U32)std::min<U64>(q, 0xFFFFFFFFull));
}

// fma.rm + cvt.rmi.sat.u32, top byte: floor(c*255 + 0.5) clamped to [0,255] (exact in F64).
static inline U32 packChannel(F32 c) {
    if (c != c) return 0;
    double v = std::floor((double)c * 4278190080.0 + 8388608.0);
    if (v <= 0.0) return 0;
    if (v >= 4294967295.0) return 255;
    return (U32)((U64)v >> 24);
}
static inline U32 toABGR(F32 r, F32 g, F32 b, F32 a) {
    return packChannel(r) | packChannel(g) << 8 | packChannel(b) << 16 | packChannel(a) << 24;
}

static inline U32 blendChannel(U32 s, U32 d, U32 fs, U32 fd, bool clamp) {
    U32 t = s * fs + d * fd;
    if (clamp) t = std::min(t, 255u * 255u);
    return ((t * 0x010101u + 0x800000u) >> 24) & 0xFF;
}
static inline U32 blendABGR(U32 src, U32 dst, U32 fsC, U32 fdC, U32 fsA, U32 fdA, bool clamp) {
    U32 r = 0;
    for (int c = 0; c < 4; c++) {
        U32 fs = (c == 3 ? fsA : fsC) >> 24, fd = (c == 3 ? fdA : fdC) >> 24;
        r |= blendChannel((src >> (8 * c)) & 0xFF, (dst >> (8 * c)) & 0xFF, fs, fd, clamp) << (8 * c);
    }
    return r;
}
// returns false when the blend shader disables the colour write.
static inline bool runBlend(int blend, U32 src, U32 dst, U32& out) {
    switch (blend) {
        case kBlendReplace: out = src; return true;
        case kBlendSrcOver: out = blendABGR(src, dst, src, ~src, src, ~src, false); return true;
        case kBlendAdditive: out = blendABGR(src, dst, ~0u, ~0u, ~0u, ~0u, true); return true;
        default: return false;
    }
}
static inline bool blendNeedsDst(int blend) { return blend == kBlendSrcOver || blend == kBlendAdditive; }

// ---- clipper (cuda/Util.hpp:221-286); barycentric Sutherland-Hodgman -------------------------
struct B2 { F32 u, v; };

static inline F32 planeDist(F32 p0, F32 p1, F32 p2, B2 b) { return fmaf(p2, b.v, fmaf(p1, b.u, p0)); }

static inline int clipPolyPlane(B2* out, const B2* in, int n, F32 p0, F32 p1, F32 p2) {
    int m = 0;
    if (n < 3) return 0;
    B2 a = in[n - 1];
    F32 da = planeDist(p0, p1, p2, a);
    for (int i = 0; i < n; i++) {
        B2 b = in[i];
        F32 db = planeDist(p0, p1, p2, b);
        if (da * db < 0.0f) {
            F32 tb = da / (da - db);
            F32 ta = 1.0f - tb;
            out[m].u = fmaf(a.u, ta, b.u * tb);
            out[m].v = fmaf(a.v, ta, b.v * tb);
            m++;
        }
        if (db >= 0.0f) out[m++] = b;
        a = b;
        da = db;
    }
    return m;
}

// v0,v1,v2: clip-space xyzw; d1 = v1-v0, d2 = v2-v0.  bary receives up to 9 vertices.
// lo[a],hi[a]: clip window per axis (x,y,z); the reference is lo=-1, hi=+1 everywhere.
static inline int clipTriangle(B2* bary, const F32* v0, const F32* v1, const F32* v2, const F32* d1, const F32* d2, const F32* lo, const F32* hi) {
    int n = 3;
    bary[0] = {0.0f, 0.0f};
    bary[1] = {1.0f, 0.0f};
    bary[2] = {0.0f, 1.0f};
    for (int a = 0; a < 3; a++) {
        bool any = (v0[3] * hi[a] < v0[a]) | (v0[3] * lo[a] > v0[a]) | (v1[3] * hi[a] < v1[a]) | (v1[3] * lo[a] > v1[a]) |
                   (v2[3] * hi[a] < v2[a]) | (v2[3] * lo[a] > v2[a]);
        if (!any) continue;
        B2 tmp[9];
        // plane "lo": x - lo*w >= 0  (reference: w + x);  plane "hi": hi*w - x >= 0  (reference: w - x)
        n = clipPolyPlane(tmp, bary, n, fmaf(-lo[a], v0[3], v0[a]), fmaf(-lo[a], d1[3], d1[a]), fmaf(-lo[a], d2[3], d2[a]));
        n = clipPolyPlane(bary, tmp, n, fmaf(hi[a], v0[3], -v0[a]), fmaf(hi[a], d1[3], -d1[a]), fmaf(hi[a], d2[3], -d2[a]));
    }
    return n;
}

// ---- plane equation in fixed point (cuda/Util.inl:65-91) -------------------------------------
struct I2 { S32 x, y; };
struct U3 { U32 x, y, z; };

static inline U3 setupPleq(F32 val0, F32 val1, F32 val2, I2 v0, I2 d1, I2 d2, F32 areaRcp, int samplesLog2) {
    F32 mx = std::fmax(std::fmax(val0, val1), val2);
    int sh = std::min(std::max(((S32)asU32(mx) >> 23) - (127 + 22), 0), 8);
    S32 t0 = (S32)(cvtRziU32(val0) >> sh);
    S32 t1 = (S32)((cvtRziU32(val1) >> sh) - (U32)t0);
    S32 t2 = (S32)((cvtRziU32(val2) >> sh) - (U32)t0);

    U32 rcpMant = (asU32(areaRcp) & 0x007FFFFFu) | 0x00800000u;
    int rcpShift = (23 + 127) - ((S32)asU32(areaRcp) >> 23);

    S64 xc = (S64)((U64)((S64)t1 * d2.y - (S64)t2 * d1.y) * (U64)rcpMant);
    S64 yc = (S64)((U64)((S64)t2 * d1.x - (S64)t1 * d2.x) * (U64)rcpMant);
    int sub = kSubpixelLog2 - samplesLog2;
    U3 p;
    p.x = (U32)(U64)shrS64(xc, rcpShift - (sh + sub));
    p.y = (U32)(U64)shrS64(yc, rcpShift - (sh + sub));

    S32 cx = (v0.x * 2 + std::min(std::min(d1.x, d2.x), 0) + std::max(std::max(d1.x, d2.x), 0)) >> (sub + 1);
    S32 cy = (v0.y * 2 + std::min(std::min(d1.y, d2.y), 0) + std::max(std::max(d1.y, d2.y), 0)) >> (sub + 1);
    S32 vcx = v0.x - (S32)((U32)cx << sub);
    S32 vcy = v0.y - (S32)((U32)cy << sub);

    p.z = (U32)t0 << sh;
    S64 corr = (S64)((U64)(shrS64(xc, 13) * (S64)vcx) + (U64)(shrS64(yc, 13) * (S64)vcy));
    p.z -= (U32)(U64)shrS64(corr, rcpShift - (sh + 13));
    p.z -= p.x * (U32)cx + p.y * (U32)cy;
    return p;
}

// ---- per-frame configuration -------------------------------------------------------------------
struct Config {
    S32 width, height;   // viewport = surface size before rounding to tiles
    S32 samplesLog2;
    U32 flags;           // kFlag*
    S32 vertexStride;    // bytes, multiple of 16
    S32 shader;          // kShader*
    S32 blend;           // kBlend*
    S32 deferredClear;
    U32 clearColor, clearDepth;
    S32 numThreads;      // fine-stage worker threads (band parallel); <=1 = scalar
    // Sort-first window support.  Triangles are set up in a PARENT viewport
    // (vpWidth x vpHeight <= 2048^2: the whole frame when it fits) exactly as the reference sets
    // them up for a viewport of that size, except that vertices are snapped once in the grid of
    // the FULL frame (fullWidth x fullHeight) and shifted by the integer offset of the parent's
    // centre from the full-frame centre (centerOfs, subpixels).  The surface (width x height) is a
    // scissor rectangle at pixel (subX0, subY0) of the parent.  Plain single viewport:
    // vp = full = surface size, offsets = 0.
    S32 fullWidth, fullHeight, centerOfsX, centerOfsY;
    // Clip window = the parent viewport in full-frame NDC (x in [clipLoX,clipHiX] * w, same for y).
    // (-1,+1) for a plain viewport, which makes every expression below bit-identical to the
    // reference's  w < |x|  /  w + x  /  w - x  forms (multiplying by +-1.0f is exact).
    F32 clipLoX, clipHiX, clipLoY, clipHiY;
    S32 subX0, subY0;    // pixel origin of the surface inside the parent viewport (multiples of 8)
    S32 vpWidth, vpHeight;
    // Cull window = the surface in full-frame NDC: triangles wholly outside it are dropped early.
    F32 cullLoX, cullHiX, cullLoY, cullHiY;
    S32 windowed;
};
// header subpixel coordinate + origin = subpixel position relative to the surface corner
static inline S64 originX(const Config& c) { return (S64)c.vpWidth * 8 - (S64)c.subX0 * 16; }
static inline S64 originY(const Config& c) { return (S64)c.vpHeight * 8 - (S64)c.subY0 * 16; }

struct Counts {  // golden counts that define the algorithmic bytes
    S64 numTris, numSubtris, numVisibleTris;
    S64 eBin, eTile, eCov, eShade;
    S64 fragments, fragmentsWritten;
    S64 vertsReferenced;
};

static inline const F32* vertexAt(const void* verts, S32 stride, U32 idx, int slot) {
    return (const F32*)((const U8*)verts + (size_t)idx * stride) + slot * 4;
}

// ---- triangle setup (cuda/TriangleSetup.inl) -------------------------------------------------
struct Snapped { I2 p0, p1, p2, lo, hi; F32 rcpW[3]; };

static inline void snapTriangle(const Config& c, const F32* v0, const F32* v1, const F32* v2, Snapped& s) {
    F32 sx = (F32)(c.fullWidth << (kSubpixelLog2 - 1));
    F32 sy = (F32)(c.fullHeight << (kSubpixelLog2 - 1));
    s.rcpW[0] = 1.0f / v0[3];
    s.rcpW[1] = 1.0f / v1[3];
    s.rcpW[2] = 1.0f / v2[3];
    // Plain single viewport: centerOfs == 0, so this is exactly TriangleSetup.inl:23-28.
    // Wrapping subtraction mirrors the device (a saturated snap minus a non-zero offset).
    s.p0 = {wrapS32((S64)cvtRniSatS32(v0[0] * s.rcpW[0] * sx) - c.centerOfsX), wrapS32((S64)cvtRniSatS32(v0[1] * s.rcpW[0] * sy) - c.centerOfsY)};
    s.p1 = {wrapS32((S64)cvtRniSatS32(v1[0] * s.rcpW[1] * sx) - c.centerOfsX), wrapS32((S64)cvtRniSatS32(v1[1] * s.rcpW[1] * sy) - c.centerOfsY)};
    s.p2 = {wrapS32((S64)cvtRniSatS32(v2[0] * s.rcpW[2] * sx) - c.centerOfsX), wrapS32((S64)cvtRniSatS32(v2[1] * s.rcpW[2] * sy) - c.centerOfsY)};
    s.lo = {std::min(std::min(s.p0.x, s.p1.x), s.p2.x), std::min(std::min(s.p0.y, s.p1.y), s.p2.y)};
    s.hi = {std::max(std::max(s.p0.x, s.p1.x), s.p2.x), std::max(std::max(s.p0.y, s.p1.y), s.p2.y)};
}

// 0 visible, 1 backfacing/degenerate, 2 falls between samples (TriangleSetup.inl:37-99)
static inline int prepareTriangle(const Config& c, const Snapped& s, I2& d1, I2& d2, S32& area) {
    d1 = {wrapS32((S64)s.p1.x - s.p0.x), wrapS32((S64)s.p1.y - s.p0.y)};
    d2 = {wrapS32((S64)s.p2.x - s.p0.x), wrapS32((S64)s.p2.y - s.p0.y)};
    area = wrapS32((S64)mulS32(d1.x, d2.y) - (S64)mulS32(d1.y, d2.x));
    if (area <= 0) return 1;

    int sampleSize = 1 << (kSubpixelLog2 - c.samplesLog2);
    S32 biasX = (c.vpWidth << (kSubpixelLog2 - 1)) - (sampleSize >> 1);
    S32 biasY = (c.vpHeight << (kSubpixelLog2 - 1)) - (sampleSize >> 1);
    S32 lox = wrapS32((S64)s.lo.x + (sampleSize - 1) + biasX) & -sampleSize;
    S32 loy = wrapS32((S64)s.lo.y + (sampleSize - 1) + biasY) & -sampleSize;
    S32 hix = wrapS32((S64)s.hi.x + biasX) & -sampleSize;
    S32 hiy = wrapS32((S64)s.hi.y + biasY) & -sampleSize;
    if (lox > hix || loy > hiy) return 2;

    S32 diff = wrapS32((S64)hix + hiy - lox - loy);
    if (diff <= sampleSize) {
        for (int pass = 0; pass < 2; pass++) {
            S32 qx = pass == 0 ? lox : hix, qy = pass == 0 ? loy : hiy;
            I2 t0 = {wrapS32((S64)s.p0.x + biasX - qx), wrapS32((S64)s.p0.y + biasY - qy)};
            I2 t1 = {wrapS32((S64)s.p1.x + biasX - qx), wrapS32((S64)s.p1.y + biasY - qy)};
            I2 t2 = {wrapS32((S64)s.p2.x + biasX - qx), wrapS32((S64)s.p2.y + biasY - qy)};
            S32 e0 = wrapS32((S64)mulS32(t0.x, t1.y) - (S64)mulS32(t0.y, t1.x));
            S32 e1 = wrapS32((S64)mulS32(t1.x, t2.y) - (S64)mulS32(t1.y, t2.x));
            S32 e2 = wrapS32((S64)mulS32(t2.x, t0.y) - (S64)mulS32(t2.y, t0.x));
            if (!(e0 < 0 || e1 < 0 || e2 < 0)) break;  // this sample is covered
            if (pass == 1 || diff == 0) return 2;
        }
    }
    return 0;
}

static inline void setupTriangle(const Config& c, TriHeader* th, TriData* td, const U32* vidx,
                                 const F32* v0, const F32* v1, const F32* v2, B2 b0, B2 b1, B2 b2,
                                 const Snapped& s, I2 d1, I2 d2, S32 area) {
    const int S = c.samplesLog2;
    F32 areaRcp = 0.0f;
    I2 wv0 = {0, 0};
    if (c.flags & (kFlagDepth | kFlagLerp)) {
        areaRcp = 1.0f / (F32)area;
        // plane equations are set up in viewport-corner coordinates (TriangleSetup.inl:127-128) and,
        // for a sort-first window, translated to the surface afterwards (exact integer shift), so
        // every window of one parent viewport renders the very same depth / barycentric values
        // as the unsplit viewport.
        wv0 = {s.p0.x + (c.vpWidth << (kSubpixelLog2 - 1)), s.p0.y + (c.vpHeight << (kSubpixelLog2 - 1))};
    }
    U3 zp = {0, 0, 0};
    U32 zmin = 0, zslope = 0;
    if (c.flags & kFlagDepth) {
        const F32 zcoef = (F32)(kDepthMax - kDepthMin) * 0.5f;
        const F32 zbias = (F32)(kDepthMax + kDepthMin) * 0.5f;  // U32 wrap of the sum is intended (== 0xFFFFFFFF)
        F32 z0 = fmaf(v0[2] * zcoef, s.rcpW[0], zbias);
        F32 z1 = fmaf(v1[2] * zcoef, s.rcpW[1], zbias);
        F32 z2 = fmaf(v2[2] * zcoef, s.rcpW[2], zbias);
        I2 zv0 = {wv0.x - (1 << (kSubpixelLog2 - S - 1)), wv0.y - (1 << (kSubpixelLog2 - S - 1))};
        zp = setupPleq(z0, z1, z2, zv0, d1, d2, areaRcp, S);
        zmin = cvtRniSatU32(std::fmin(std::fmin(z0, z1), z2) - (F32)(2200u << S));
        if (S != 0) {
            S32 ax = (S32)zp.x; ax = ax >= 0 ? ax : wrapS32(-(S64)ax);
            S32 ay = std::max((S32)zp.y, -0x7FFFFFFF); ay = ay >= 0 ? ay : -ay;
            U32 tmp = (U32)ax + (U32)ay;
            int k = std::max(S - 1, 0);
            zslope = tmp << k;
            if ((zslope >> k) != tmp) zslope = 0xFFFFFFFFu;
        }
    }
    U3 wp = {0, 0, 0}, up = {0, 0, 0}, vp = {0, 0, 0};
    if (c.flags & kFlagLerp) {
        F32 wcoef = std::fmin(std::fmin(v0[3], v1[3]), v2[3]) * (F32)kBaryMax;
        F32 w0 = wcoef * s.rcpW[0], w1 = wcoef * s.rcpW[1], w2 = wcoef * s.rcpW[2];
        wp = setupPleq(w0, w1, w2, wv0, d1, d2, areaRcp, S + 1);
        up = setupPleq(b0.u * w0, b1.u * w1, b2.u * w2, wv0, d1, d2, areaRcp, S + 1);
        vp = setupPleq(b0.v * w0, b1.v * w1, b2.v * w2, wv0, d1, d2, areaRcp, S + 1);
    }
    {   // translate the constants from full-frame to viewport-local sample coordinates
        const U32 ox = (U32)c.subX0 << S, oy = (U32)c.subY0 << S;
        zp.z += zp.x * ox + zp.y * oy;
        wp.z += wp.x * (2 * ox) + wp.y * (2 * oy);
        up.z += up.x * (2 * ox) + up.y * (2 * oy);
        vp.z += vp.x * (2 * ox) + vp.y * (2 * oy);
    }
    if (c.flags & kFlagDepth) { td->zx = zp.x; td->zy = zp.y; td->zb = zp.z; td->zslope = zslope; }
    if (c.flags & kFlagLerp) {
        td->wx = (S32)wp.x; td->wy = (S32)wp.y; td->wb = (S32)wp.z;
        td->ux = (S32)up.x; td->uy = (S32)up.y; td->ub = (S32)up.z;
        td->vx = (S32)vp.x; td->vy = (S32)vp.y; td->vb = (S32)vp.z;
    } else {
        td->vb = 0;
    }
    td->vi0 = vidx[0]; td->vi1 = vidx[1]; td->vi2 = vidx[2];

    U32 f01 = selectFlips(d1.x, d1.y);
    U32 f12 = selectFlips(d2.x - d1.x, d2.y - d1.y);
    U32 f20 = selectFlips(-d2.x, -d2.y);
    th->v0x = (S16)s.p0.x; th->v0y = (S16)s.p0.y;
    th->v1x = (S16)s.p1.x; th->v1y = (S16)s.p1.y;
    th->v2x = (S16)s.p2.x; th->v2y = (S16)s.p2.y;
    th->misc = (zmin & 0xFFFFF000u) | (f01 << 6) | (f12 << 2) | (f20 >> 2);
}

// One input triangle.  Returns the number of surviving sub-triangles (0..7).
//  * 0/1 survivors are written to slot `tri`.
//  * >=2 survivors need a run of slots starting at `base`; pass base < 0 to only count
//    (nothing is written), then call again with the allocated base.
// When base + n > maxSubtris nothing is written for that triangle (TriangleSetup.inl:372-378).
static inline int setupOneTriangle(const Config& c, const void* verts, const S32* indices, S32 tri, S32 base,
                                   TriHeader* triHeader, TriData* triData, S32 maxSubtris) {
    const S32 aabbLimit = (1 << (kMaxViewportLog2 + kSubpixelLog2)) - 1;
    const F32 lo[3] = {c.clipLoX, c.clipLoY, -1.0f}, hi[3] = {c.clipHiX, c.clipHiY, 1.0f};
    U32 vidx[3] = {(U32)indices[tri * 3 + 0], (U32)indices[tri * 3 + 1], (U32)indices[tri * 3 + 2]};
    const F32* v0 = vertexAt(verts, c.vertexStride, vidx[0], 0);
    const F32* v1 = vertexAt(verts, c.vertexStride, vidx[1], 0);
    const F32* v2 = vertexAt(verts, c.vertexStride, vidx[2], 0);

    // (1) all three vertices outside one clip plane -> culled (TriangleSetup.inl:262-281; the
    //     reference's "v0 outside anything" pre-test is only a shortcut for the same predicate)
    for (int a = 0; a < 3; a++) {
        if ((v0[3] * hi[a] < v0[a]) & (v1[3] * hi[a] < v1[a]) & (v2[3] * hi[a] < v2[a])) return 0;
        if ((v0[3] * lo[a] > v0[a]) & (v1[3] * lo[a] > v1[a]) & (v2[3] * lo[a] > v2[a])) return 0;
    }
    if (c.windowed) {   // sort-first window: wholly outside the surface rectangle -> culled (pure cull)
        const F32 clo[2] = {c.cullLoX, c.cullLoY}, chi[2] = {c.cullHiX, c.cullHiY};
        for (int a = 0; a < 2; a++) {
            if ((v0[3] * chi[a] < v0[a]) & (v1[3] * chi[a] < v1[a]) & (v2[3] * chi[a] < v2[a])) return 0;
            if ((v0[3] * clo[a] > v0[a]) & (v1[3] * clo[a] > v1[a]) & (v2[3] * clo[a] > v2[a])) return 0;
        }
    }
    // (2) inside the depth range and inside the S16 guard band -> fast path (:285-321)
    Snapped s;
    I2 d1, d2;
    S32 area;
    if ((v0[3] >= std::fabs(v0[2])) & (v1[3] >= std::fabs(v1[2])) & (v2[3] >= std::fabs(v2[2]))) {
        snapTriangle(c, v0, v1, v2, s);
        S32 loxy = std::min(s.lo.x, s.lo.y), hixy = std::max(s.hi.x, s.hi.y);
        if (loxy >= -32768 && hixy <= 32767 && wrapS32((S64)hixy - loxy) <= aabbLimit) {
            int res = prepareTriangle(c, s, d1, d2, area);
            if (res == 0)
                setupTriangle(c, &triHeader[tri], &triData[tri], vidx, v0, v1, v2, {0.0f, 0.0f}, {1.0f, 0.0f}, {0.0f, 1.0f}, s, d1, d2, area);
            return res == 0 ? 1 : 0;
        }
    }
    // (3) clip (:326-412)
    F32 e1[4], e2[4];
    for (int k = 0; k < 4; k++) e1[k] = v1[k] - v0[k], e2[k] = v2[k] - v0[k];
    B2 bary[9];
    int numVerts = clipTriangle(bary, v0, v1, v2, e1, e2, lo, hi);
    F32 cv[9][4];
    for (int i = 0; i < numVerts; i++)
        for (int k = 0; k < 4; k++) cv[i][k] = fmaf(e2[k], bary[i].v, fmaf(e1[k], bary[i].u, v0[k]));

    int numSub = 0;
    for (int i = 2; i < numVerts; i++) {
        snapTriangle(c, cv[0], cv[i - 1], cv[i], s);
        if (prepareTriangle(c, s, d1, d2, area) == 0) numSub++;
    }
    S32 slot = tri;
    if (numSub > 1) {
        if (base < 0) return numSub;
        triHeader[tri].misc = (U32)base;
        if (base + numSub > maxSubtris) return numSub;
        slot = base;
    }
    for (int i = 2; i < numVerts && numSub > 0; i++) {
        snapTriangle(c, cv[0], cv[i - 1], cv[i], s);
        if (prepareTriangle(c, s, d1, d2, area) == 0) {
            setupTriangle(c, &triHeader[slot], &triData[slot], vidx, cv[0], cv[i - 1], cv[i], bary[0], bary[i - 1], bary[i], s, d1, d2, area);
            slot++;
        }
    }
    return numSub;
}

// Whole setup stage.  Sub-triangle runs (>=2 survivors of a clipped triangle) are allocated
// sequentially in triangle order starting at numTris (the device allocates with an atomic, so
// its base slots differ; compare through the triHeader[tri].misc indirection).
// Returns the final sub-triangle cursor (CRAtomics.numSubtris); writes are clamped to maxSubtris.
static inline S32 triangleSetup(const Config& c, const void* verts, const S32* indices, S32 numTris,
                                U8* triSubtris, TriHeader* triHeader, TriData* triData, S32 maxSubtris) {
    int nthreads = std::max(1, (int)c.numThreads);
    auto passA = [&](int t) {
        S32 lo = (S32)((S64)numTris * t / nthreads), hi = (S32)((S64)numTris * (t + 1) / nthreads);
        for (S32 tri = lo; tri < hi; tri++)
            triSubtris[tri] = (U8)setupOneTriangle(c, verts, indices, tri, -1, triHeader, triData, maxSubtris);
    };
    if (nthreads == 1) passA(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(passA, t);
        for (auto& x : th) x.join();
    }
    S32 cursor = numTris;
    for (S32 tri = 0; tri < numTris; tri++) {
        if (triSubtris[tri] > 1) {
            setupOneTriangle(c, verts, indices, tri, cursor, triHeader, triData, maxSubtris);
            cursor += triSubtris[tri];
        }
    }
    return cursor;
}

// ---- coverage (cuda/Util.inl:95-134, :340-357) -----------------------------------------------
struct Edges {
    // E_i(sample) = (ox - sx) * dy - (oy - sy) * dx - tie >= 0, origin/sample in viewport-centred subpixels
    S64 ox[3], oy[3], dx[3], dy[3], tie[3];
};
static inline void edgesFromHeader(const TriHeader& h, Edges& e) {
    S64 x0 = h.v0x, y0 = h.v0y, x1 = h.v1x, y1 = h.v1y, x2 = h.v2x, y2 = h.v2y;
    e.ox[0] = x0; e.oy[0] = y0; e.dx[0] = x1 - x0; e.dy[0] = y1 - y0;
    e.ox[1] = x1; e.oy[1] = y1; e.dx[1] = x2 - x1; e.dy[1] = y2 - y1;
    e.ox[2] = x0; e.oy[2] = y0; e.dx[2] = x0 - x2; e.dy[2] = y0 - y2;
    for (int i = 0; i < 3; i++) e.tie[i] = (e.dy[i] > 0 || (e.dy[i] == 0 && e.dx[i] <= 0)) ? 1 : 0;
}
static inline bool sampleInside(const Edges& e, S64 sx, S64 sy) {
    for (int i = 0; i < 3; i++)
        if ((e.ox[i] - sx) * e.dy[i] - (e.oy[i] - sy) * e.dx[i] - e.tie[i] < 0) return false;
    return true;
}
// 64-bit pixel-centre coverage of one 8x8 tile, bit = x + 8*y (single-sample rule).
static inline U64 coverTile(const Config& c, const TriHeader& h, int tileX, int tileY) {
    Edges e;
    edgesFromHeader(h, e);
    U64 m = 0;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            S64 sx = (S64)(tileX * 8 + x) * 16 + 8 - originX(c);
            S64 sy = (S64)(tileY * 8 + y) * 16 + 8 - originY(c);
            if (sampleInside(e, sx, sy)) m |= (U64)1 << (x + 8 * y);
        }
    return m;
}
// sample mask of one pixel (bit i = sample i), cuda/Util.inl:340-383
static inline U32 coverPixelSamples(const Config& c, const Edges& e, int px, int py) {
    const int S = c.samplesLog2, N = 1 << S;
    S64 cx = (S64)px * 16 + 8 - originX(c), cy = (S64)py * 16 + 8 - originY(c);
    U32 m = 0;
    for (int i = 0; i < N; i++) {
        S64 offx = S == 0 ? 0 : (S64)(kMsaaX[S][i] * 2 + 1 - N) * (1 << (kSubpixelLog2 - S - 1));
        S64 offy = S == 0 ? 0 : (S64)(i * 2 + 1 - N) * (1 << (kSubpixelLog2 - S - 1));
        if (sampleInside(e, cx + offx, cy + offy)) m |= 1u << i;
    }
    return m;
}

// conservative triangle-vs-square overlap (CudaRaster.cpp:978-988 / :1129-1143); coordinates in
// viewport-corner subpixels, square centre (cx,cy), half extent `half`.
static inline bool overlapsSquare(S64 v0x, S64 v0y, S64 d01x, S64 d01y, S64 d02x, S64 d02y, S64 lox, S64 loy, S64 hix, S64 hiy, S64 cx, S64 cy, S64 half) {
    if (lox >= cx + half || loy >= cy + half || hix <= cx - half || hiy <= cy - half) return false;
    S64 p0x = cx - v0x, p0y = cy - v0y, p1x = p0x - d01x, p1y = p0y - d01y;
    S64 d12x = d02x - d01x, d12y = d02y - d01y;
    if (p0x * d01y - p0y * d01x >= (std::llabs(d01x) + std::llabs(d01y)) * half) return false;
    if (p0y * d02x - p0x * d02y >= (std::llabs(d02x) + std::llabs(d02y)) * half) return false;
    if (p1x * d12y - p1y * d12x >= (std::llabs(d12x) + std::llabs(d12y)) * half) return false;
    return true;
}

// ---- shading (cuda/FineRaster.inl:21-119, cuda/PixelPipe.inl:43-65) ---------------------------
struct Bary { F32 b0, b1, b2; };
static inline Bary computeBary(const TriData& d, S32 sx, S32 sy) {
    F32 w = 1.0f / (F32)wrapS32((S64)mulS32(d.wx, sx) + mulS32(d.wy, sy) + d.wb);
    F32 u = w * (F32)wrapS32((S64)mulS32(d.ux, sx) + mulS32(d.uy, sy) + d.ub);
    F32 v = w * (F32)wrapS32((S64)mulS32(d.vx, sx) + mulS32(d.vy, sy) + d.vb);
    return {1.0f - u - v, u, v};
}
static inline void lerpVarying(F32* out, const void* verts, S32 stride, const TriData& d, int varying, Bary b) {
    const F32* a0 = vertexAt(verts, stride, d.vi0, varying + 1);
    const F32* a1 = vertexAt(verts, stride, d.vi1, varying + 1);
    const F32* a2 = vertexAt(verts, stride, d.vi2, varying + 1);
    for (int k = 0; k < 4; k++) out[k] = fmaf(a2[k], b.b2, fmaf(a0[k], b.b0, a1[k] * b.b1));
}

// Procedural "texture + Phong" shader used by BASELINE config 3.  The reference's texPhong
// (test/shader/Shaders.cu:37-51, :124-190) needs a texture atlas asset that is not in the tree,
// so the texture is replaced by a procedural checker; every operation is IEEE single (sqrt, div,
// fma written out) so that the CUDA and CPU versions agree bit for bit.
static inline F32 dot3(const F32* a, const F32* b) { return fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0])); }
static inline void phongProc(const F32* camPos, const F32* camNrm, const F32* tex, F32* rgba) {
    F32 il = 1.0f / std::sqrt(dot3(camPos, camPos));
    F32 nl = 1.0f / std::sqrt(dot3(camNrm, camNrm));
    F32 I[3] = {camPos[0] * il, camPos[1] * il, camPos[2] * il};
    F32 N[3] = {camNrm[0] * nl, camNrm[1] * nl, camNrm[2] * nl};
    F32 dIN = dot3(I, N);
    F32 k = dIN * 2.0f;
    F32 R[3] = {fmaf(-N[0], k, I[0]), fmaf(-N[1], k, I[1]), fmaf(-N[2], k, I[2])};
    F32 diffuse = fmaf(std::fmax(-dIN, 0.0f), 0.75f, 0.25f);
    F32 sp = std::fmax(-dot3(I, R), 0.0f);
    sp = sp * sp; sp = sp * sp; sp = sp * sp; sp = sp * sp;  // glossiness 16
    // procedural texture: 16x16 checker in (u,v), two albedos
    F32 fu = tex[0] * 16.0f, fv = tex[1] * 16.0f;
    int cu = (int)std::floor(fu), cvv = (int)std::floor(fv);
    bool odd = ((cu ^ cvv) & 1) != 0;
    F32 alb[3] = {odd ? 0.9f : 0.2f, odd ? 0.6f : 0.5f, odd ? 0.3f : 0.8f};
    rgba[0] = fmaf(sp, 0.5f, diffuse * alb[0]);
    rgba[1] = fmaf(sp, 0.5f, diffuse * alb[1]);
    rgba[2] = fmaf(sp, 0.5f, diffuse * alb[2]);
    rgba[3] = 1.0f;
}

// Returns false when the fragment is discarded.
static inline bool runShader(const Config& c, const void* verts, const TriData& d, S32 triIdx, int px, int py, U32 centroidCode, U32& color) {
    (void)triIdx;
    Bary b = {0.0f, 0.0f, 1.0f};
    if (c.flags & kFlagLerp) {
        const int S = c.samplesLog2;
        if (S == 0) b = computeBary(d, px * 2 + 1, py * 2 + 1);
        else b = computeBary(d, (px << (S + 1)) + (S32)(centroidCode & 0xF), (py << (S + 1)) + (S32)(centroidCode >> 4));
    }
    switch (c.shader) {
        case kShaderConstant: color = toABGR(1.0f, 0.0f, 0.0f, 1.0f); return true;
        case kShaderGouraud: {
            F32 v[4];
            lerpVarying(v, verts, c.vertexStride, d, 0, b);
            color = toABGR(v[0], v[1], v[2], v[3]);
            return true;
        }
        case kShaderGouraudDiscard: {  // alpha-test variant (exercises m_discard)
            F32 v[4];
            lerpVarying(v, verts, c.vertexStride, d, 0, b);
            if (v[3] < 0.5f) return false;
            color = toABGR(v[0], v[1], v[2], v[3]);
            return true;
        }
        case kShaderPhongProc: {
            F32 cp[4], cn[4], tx[4], rgba[4];
            lerpVarying(cp, verts, c.vertexStride, d, 0, b);
            lerpVarying(cn, verts, c.vertexStride, d, 1, b);
            lerpVarying(tx, verts, c.vertexStride, d, 2, b);
            phongProc(cp, cn, tx, rgba);
            color = toABGR(rgba[0], rgba[1], rgba[2], rgba[3]);
            return true;
        }
    }
    color = 0xFF0000FFu;
    return true;
}

// RenderModeFlag_EnableQuads (cuda/PixelPipe.hpp:34, :59-69; cuda/FineRaster.inl:396-430, :705-724,
// :1055-1090): the shader runs on all four pixels of every 2x2 quad that holds a fragment, and
// dFdx(v) = v(x|1) - v(x&~1), dFdy(v) = v(y|1) - v(y&~1) inside the quad.  Each quad pixel is shaded at
// ITS OWN shading point (codes[k], k = (x&1) + 2*(y&1)): the pixel centre when single-sampled, the
// centroid of its own sample mask under MSAA (centre when that mask is empty).
// Test shader "gouraudQuads": colour = (8|dFdx c.r| + 8|dFdy c.r|, 8|dFdx c.g| + 8|dFdy c.g|, c.b, c.a).
static inline bool runShaderQuads(const Config& c, const void* verts, const TriData& d, int px, int py, const U32* codes, U32& color) {
    const int S = c.samplesLog2;
    F32 v[4][4];
    for (int k = 0; k < 4; k++) {
        const int qx = (px & ~1) + (k & 1), qy = (py & ~1) + (k >> 1);
        Bary b = {0.0f, 0.0f, 1.0f};
        if (c.flags & kFlagLerp) {
            if (S == 0) b = computeBary(d, qx * 2 + 1, qy * 2 + 1);
            else b = computeBary(d, (qx << (S + 1)) + (S32)(codes[k] & 0xF), (qy << (S + 1)) + (S32)(codes[k] >> 4));
        }
        lerpVarying(v[k], verts, c.vertexStride, d, 0, b);
    }
    const int own = (px & 1) + 2 * (py & 1), row = own & 2, col = own & 1;
    F32 dx[2], dy[2];
    for (int i = 0; i < 2; i++) {
        dx[i] = v[row + 1][i] - v[row][i];
        dy[i] = v[2 + col][i] - v[col][i];
    }
    color = toABGR(fmaf(std::fabs(dy[0]), 8.0f, std::fabs(dx[0]) * 8.0f), fmaf(std::fabs(dy[1]), 8.0f, std::fabs(dx[1]) * 8.0f), v[own][2], v[own][3]);
    return true;
}

// The multi-sample kernel's per-pixel conservative depth kill (cuda/FineRaster.inl:1034-1068), needed only
// because in quads mode a killed HELPER pixel is shaded at its centre instead of its centroid.
// pixZMax = the kernel's tileDepth[pixel]: CR_DEPTH_MAX until the pixel's first ROP of this draw, then the
// maximum over its samples (FineRaster.inl:934-935, :1101-1108).
static inline bool msaaPixelZKill(const Config& c, const TriHeader& h, const TriData& d, int px, int py, U32 pixZMax) {
    if (!(c.flags & kFlagDepth)) return false;
    const int S = c.samplesLog2;
    const U32 zbase = ((d.zx * (U32)px + d.zy * (U32)py) << S) + d.zb;
    const U32 zmin = ((d.zx + d.zy) << std::max(S - 1, 0)) + zbase - d.zslope;
    if (zmin >= pixZMax && zmin < zmin + d.zslope * 2u) return true;
    return (h.misc & 0xFFFFF000u) >= pixZMax;
}

static inline U32 centroidCode(int S, U32 sampleMask) {  // FineRaster.inl:151-164
    int y = msaaCentroid(S, sampleMask);
    if (y < 0) return 0x11u << S;
    return (U32)(kMsaaX[S][y] * 0x02 + y * 0x20 + 0x11);
}

// ---- fine raster: the serial rule -----------------------------------------------
// Surfaces: U32 [roundedH][roundedW * N]; sample i of pixel (x,y) at column (x>>3)*8*N + i*8 + (x&7).
struct Surface { U32* color; U32* depth; S32 roundedW, roundedH, pitch; U32* pixZMax; };   // pixZMax: [roundedH][roundedW], MSAA + quads only

static inline size_t texelIndex(const Surface& s, int N, int x, int y, int sample) {
    return (size_t)y * s.pitch + (size_t)(x >> 3) * 8 * N + (size_t)sample * 8 + (x & 7);
}

// Rasterizes one sub-triangle restricted to tile rows [tileRowLo,tileRowHi) under the serial rule.
static inline void rasterTriangle(const Config& c, const void* verts, const TriHeader& h, const TriData& d, S32 triIdx,
                                  const Surface& s, int tileRowLo, int tileRowHi, Counts* cnt) {
    const int S = c.samplesLog2, N = 1 << S;
    Edges e;
    edgesFromHeader(h, e);
    // tile bbox from the snapped vertices (viewport-corner subpixels), then exact tests per sample
    S64 minx = std::min(std::min(h.v0x, h.v1x), h.v2x) + originX(c);
    S64 maxx = std::max(std::max(h.v0x, h.v1x), h.v2x) + originX(c);
    S64 miny = std::min(std::min(h.v0y, h.v1y), h.v2y) + originY(c);
    S64 maxy = std::max(std::max(h.v0y, h.v1y), h.v2y) + originY(c);
    if (maxx < 0 || maxy < 0) return;
    int tx0 = (int)std::max<S64>(minx >> 7, 0), tx1 = (int)std::min<S64>(maxx >> 7, s.roundedW / 8 - 1);
    int ty0 = (int)std::max<S64>(miny >> 7, tileRowLo), ty1 = (int)std::min<S64>(maxy >> 7, tileRowHi - 1);
    for (int ty = ty0; ty <= ty1; ty++)
        for (int tx = tx0; tx <= tx1; tx++) {
            bool anyCov = false, anyWritten = false;
            // MSAA + quads: the reference's per-pixel conservative kill decides whether the shader/ROP pair
            // runs at all, and the ROP refreshes the pixel's bound even when no sample survives
            const bool quadsMsaa = (c.flags & kFlagQuads) != 0 && S > 0;
            for (int q = 0; q < 16; q++) {   // the 2x2 quads of the tile; pixels are independent of each other
                const int qx0 = tx * 8 + (q & 3) * 2, qy0 = ty * 8 + (q >> 2) * 2;
                U32 masks[4], codes[4];
                bool kill[4] = {false, false, false, false};
                bool any = false;
                for (int k = 0; k < 4; k++) {
                    masks[k] = coverPixelSamples(c, e, qx0 + (k & 1), qy0 + (k >> 1));
                    any |= masks[k] != 0;
                }
                if (!any) continue;
                // quads mode: the four pixels of a quad are shaded TOGETHER, each at its own shading point, from
                // the state the quad had before this triangle (the four lanes run in lock step in the reference)
                for (int k = 0; k < 4; k++) {
                    if (quadsMsaa && masks[k] != 0)
                        kill[k] = msaaPixelZKill(c, h, d, qx0 + (k & 1), qy0 + (k >> 1), s.pixZMax[(size_t)(qy0 + (k >> 1)) * s.roundedW + qx0 + (k & 1)]);
                    codes[k] = centroidCode(S, kill[k] ? 0u : masks[k]);
                }
                for (int k = 0; k < 4; k++) {
                    const int px = qx0 + (k & 1), py = qy0 + (k >> 1);
                    const U32 mask = masks[k];
                    if (!mask) continue;
                    anyCov = true;
                    if (cnt) cnt->fragments++;
                    // depth test per sample (strictly less) against the state left by earlier triangles
                    U32 pass = 0;
                    U32 depth[8];
                    for (int i = 0; i < N; i++) {
                        if (!(mask >> i & 1)) continue;
                        if (c.flags & kFlagDepth) {
                            U32 sx = (U32)(px * N + kMsaaX[S][i]), sy = (U32)(py * N + i);
                            depth[i] = d.zx * sx + d.zy * sy + d.zb;
                            if (depth[i] >= s.depth[texelIndex(s, N, px, py, i)]) continue;
                        }
                        pass |= 1u << i;
                    }
                    // (the reference shades whenever sampleMask != 0; shading has no side effects,
                    //  so skipping it when no sample survives is unobservable)
                    if (!pass && !quadsMsaa) continue;
                    if (kill[k]) continue;
                    U32 color;
                    if (c.flags & kFlagQuads) {
                        if (c.shader == kShaderGouraudQuads) runShaderQuads(c, verts, d, px, py, codes, color);
                        else if (!runShader(c, verts, d, triIdx, px, py, codes[k], color)) continue;
                    } else if (!runShader(c, verts, d, triIdx, px, py, centroidCode(S, mask), color)) continue;
                    if (pass) {
                        anyWritten = true;
                        if (cnt) cnt->fragmentsWritten++;
                    }
                    for (int i = 0; i < N; i++) {
                        if (!(pass >> i & 1)) continue;
                        size_t t = texelIndex(s, N, px, py, i);
                        if (c.flags & kFlagDepth) s.depth[t] = depth[i];
                        U32 out;
                        if (runBlend(c.blend, color, s.color[t], out)) s.color[t] = out;
                    }
                    if (quadsMsaa && (c.flags & kFlagDepth)) {   // the ROP ran on this pixel: its conservative bound becomes exact
                        U32 m = 0;
                        for (int i = 0; i < N; i++) m = std::max(m, s.depth[texelIndex(s, N, px, py, i)]);
                        U32& z = s.pixZMax[(size_t)py * s.roundedW + px];
                        if (m < z) z = m;
                    }
                }
            }
            if (cnt) cnt->eCov += anyCov, cnt->eShade += anyWritten;
        }
}

// ---- around the hot path ---------------------------------------------------------
// MSAA resolve (CudaSurface::resolveToScreen, CudaSurface.hpp:73 -- the Linux port dropped its body, so this
// row is "parity unpinned": box filter, per 8-bit channel (sum of the N samples + N/2) >> log2 N).
static inline void resolveSurface(const U32* src, int width, int height, int N, U32* dst, int dstPitch, bool flipY) {
    const int roundedW = (width + 7) & ~7;
    int L = 0;
    while ((1 << L) < N) L++;
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            U32 sum[4] = {0, 0, 0, 0};
            for (int i = 0; i < N; i++) {
                const U32 t = src[(size_t)y * roundedW * N + (size_t)(x >> 3) * 8 * N + (size_t)i * 8 + (x & 7)];
                for (int ch = 0; ch < 4; ch++) sum[ch] += (t >> (8 * ch)) & 0xFF;
            }
            U32 out = 0;
            for (int ch = 0; ch < 4; ch++) out |= (((sum[ch] + (U32)(N >> 1)) >> L) & 0xFF) << (8 * ch);
            dst[(size_t)(flipY ? height - 1 - y : y) * dstPitch + x] = out;
        }
}

// Vertex-shader stage: clipPos = posToClip * (modelPos, 1) (test/shader/PassThrough.cu:31-34) with the product
// evaluated as the fma chain nvcc contracts framework/base/Math.hpp's generic Matrix::operator* into
// (r[i] += m(i,j) * v[j], j ascending).  m is column-major: m[col*4 + row].
static inline void transformPoint(const F32* m, const F32* p, F32* out) {
    for (int r = 0; r < 4; r++) out[r] = fmaf(m[12 + r], 1.0f, fmaf(m[8 + r], p[2], fmaf(m[4 + r], p[1], m[r] * p[0])));
}

// Sub-triangle index list in submission order (what the bin/tile queues carry: tri*8 + sub|7).
// Full frame: setup + serial-rule rasterization, band-parallel over tile rows when numThreads > 1.
// Returns CRAtomics.numSubtris.
static inline S32 renderFrame(const Config& c, const void* verts, const S32* indices, S32 numTris, U32* color, U32* depth, Counts* counts) {
    const int N = 1 << c.samplesLog2;
    Surface s;
    s.color = color; s.depth = depth;
    s.roundedW = (c.width + 7) & ~7; s.roundedH = (c.height + 7) & ~7;
    s.pitch = s.roundedW * N;
    std::vector<U32> pixZMax;
    s.pixZMax = nullptr;
    if ((c.flags & kFlagQuads) && c.samplesLog2 > 0) {
        pixZMax.assign((size_t)s.roundedW * s.roundedH, kDepthMax);
        s.pixZMax = pixZMax.data();
    }

    std::vector<U8> subtris((size_t)std::max(numTris, 1));
    std::vector<TriHeader> hdr;
    std::vector<TriData> data;
    S32 numSubtris = numTris;
    {   // size with the host driver's slack and retry on overflow (CudaRaster.cpp:264-339)
        S32 cap = numTris + 4096;
        for (;;) {
            hdr.assign((size_t)cap, TriHeader{});
            data.assign((size_t)cap, TriData{});
            numSubtris = triangleSetup(c, verts, indices, numTris, subtris.data(), hdr.data(), data.data(), cap);
            if (numSubtris <= cap) break;
            cap = numSubtris;
        }
    }
    if (counts) {
        std::memset(counts, 0, sizeof(*counts));
        counts->numTris = numTris;
    }
    if (c.deferredClear)
        for (size_t i = 0, n = (size_t)s.pitch * s.roundedH; i < n; i++) color[i] = c.clearColor, depth[i] = c.clearDepth;

    int tileRows = s.roundedH / 8;
    int nthreads = std::min(std::max(1, (int)c.numThreads), std::max(1, tileRows));
    std::vector<Counts> tc((size_t)nthreads);
    auto worker = [&](int t) {
        Counts& k = tc[(size_t)t];
        std::memset(&k, 0, sizeof(k));
        int r0 = (int)((S64)tileRows * t / nthreads), r1 = (int)((S64)tileRows * (t + 1) / nthreads);
        for (S32 tri = 0; tri < numTris; tri++) {
            int n = subtris[(size_t)tri];
            for (int sub = 0; sub < n; sub++) {
                S32 di = n == 1 ? tri : (S32)hdr[(size_t)tri].misc + sub;
                rasterTriangle(c, verts, hdr[(size_t)di], data[(size_t)di], tri, s, r0, r1, counts ? &k : nullptr);
            }
        }
    };
    if (nthreads == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(worker, t);
        for (auto& x : th) x.join();
    }
    if (counts) {
        for (auto& k : tc) {
            counts->fragments += k.fragments; counts->fragmentsWritten += k.fragmentsWritten;
            counts->eCov += k.eCov; counts->eShade += k.eShade;
        }
        // overlap counters for the algorithmic-byte formula
        S32 maxV = 0;
        for (S64 i = 0; i < (S64)numTris * 3; i++) maxV = std::max(maxV, indices[i]);
        std::vector<U8> vertSeen((size_t)maxV + 1, 0);
        S64 vref = 0;
        for (S64 i = 0; i < (S64)numTris * 3; i++)
            if (!vertSeen[(size_t)indices[i]]) vertSeen[(size_t)indices[i]] = 1, vref++;
        int tilesX = s.roundedW / 8, tilesY = s.roundedH / 8;
        int binsX = (tilesX + 15) / 16, binsY = (tilesY + 15) / 16;
        S64 nsub = 0, nvis = 0, eBin = 0, eTile = 0;
        for (S32 tri = 0; tri < numTris; tri++) {
            int n = subtris[(size_t)tri];
            if (n) nvis++;
            for (int sub = 0; sub < n; sub++) {
                nsub++;
                S32 di = n == 1 ? tri : (S32)hdr[(size_t)tri].misc + sub;
                const TriHeader& h = hdr[(size_t)di];
                S64 v0x = h.v0x + originX(c), v0y = h.v0y + originY(c);
                S64 d01x = h.v1x - h.v0x, d01y = h.v1y - h.v0y, d02x = h.v2x - h.v0x, d02y = h.v2y - h.v0y;
                S64 lox = v0x + std::min<S64>(0, std::min(d01x, d02x)), hix = v0x + std::max<S64>(0, std::max(d01x, d02x));
                S64 loy = v0y + std::min<S64>(0, std::min(d01y, d02y)), hiy = v0y + std::max<S64>(0, std::max(d01y, d02y));
                for (int level = 0; level < 2; level++) {
                    const S64 half = level == 0 ? 16 * 8 * 16 / 2 : 8 * 16 / 2;
                    const int nx = level == 0 ? binsX : tilesX, ny = level == 0 ? binsY : tilesY;
                    auto fl = [&](S64 v) { return (v >= 0 ? v : v - (2 * half - 1)) / (2 * half); };
                    int x0 = (int)std::max<S64>(fl(lox) - 1, 0), x1 = (int)std::min<S64>(fl(hix) + 1, nx - 1);
                    int y0 = (int)std::max<S64>(fl(loy) - 1, 0), y1 = (int)std::min<S64>(fl(hiy) + 1, ny - 1);
                    S64 n2 = 0;
                    for (int y = y0; y <= y1; y++)
                        for (int x = x0; x <= x1; x++)
                            n2 += overlapsSquare(v0x, v0y, d01x, d01y, d02x, d02y, lox, loy, hix, hiy, (2 * x + 1) * half, (2 * y + 1) * half, half);
                    (level == 0 ? eBin : eTile) += n2;
                }
            }
        }
        counts->numSubtris = nsub;
        counts->numVisibleTris = nvis;
        counts->eBin = eBin;
        counts->eTile = eTile;
        counts->vertsReferenced = vref;
    }
    return numSubtris;
}

}  // namespace gold
