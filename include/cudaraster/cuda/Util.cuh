// Device helpers of the B200 pipeline: saturating conversions, colour pack / blend, flip-bit
// selection, the barycentric clipper and the fixed-point plane-equation setup.
//
// These follow the reference's NUMERICAL contract (reference files
// src/cudaraster/cuda/Util.hpp:182-298 and Util.inl:30-91) so that triHeader / triData are
// bit-identical, but are written against today's intrinsics: no Fermi video-SIMD PTX, no
// texture references, no implicit warp-synchronous code.  Every float expression spells out its
// roundings (__fmul_rn / __fmaf_rn ...) at exactly the places where nvcc contracts the
// reference's source, so results do not depend on -fmad.
#pragma once
#include <cuda_runtime.h>
#include "PrivateDefs.hpp"

namespace FW {

// c_msaaPatterns[log2 N][sampleY] = sampleX  (reference: cuda/Util.hpp:27-35)
#define CRB_MSAA_X_TABLE {{0, 0, 0, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0, 0, 0}, {1, 3, 0, 2, 0, 0, 0, 0}, {7, 2, 4, 0, 6, 3, 1, 5}}

__host__ __device__ __forceinline__ int msaaSampleX(int samplesLog2, int i) {
    // packed 3-bit nibbles, sample i in bits [4i, 4i+3]
    const unsigned pat[4] = {0x0u, 0x10u, 0x2031u, 0x51360427u};
    return (int)((pat[samplesLog2] >> (4 * i)) & 0xFu);
}

#ifdef __CUDACC__

// ---- 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256, PTX ISA 8.8 .v8.b32) ----------------------
// A 64-byte triData record is two 32-byte halves {row 0, row 1} / {row 2, row 3}.  One 256-bit access per half instead of
// one 128-bit access per row: half the L1 wavefronts of the record gathers (the fine raster is bound by them: ncu
// l1tex__data_pipe_lsu_wavefronts) and whole 32-byte sectors on the way out (no partial-sector fills in L2).
// The address must be 32-byte aligned and global.
#ifndef CRB_WIDE_LDST
#define CRB_WIDE_LDST 1
#endif
#ifndef CRB_WIDE_LD
#define CRB_WIDE_LD CRB_WIDE_LDST
#endif
#ifndef CRB_WIDE_ST
#define CRB_WIDE_ST CRB_WIDE_LDST
#endif
#ifndef CRB_FINE_STREAM
#define CRB_FINE_STREAM 3   // bit 0 = visibility buffer with streaming (evict-first) accesses, bit 1 = record gathers without L1 allocation: both are read once, the vertices they would displace from L1 are reused (C2 fine raster 34.9 -> 34.1 us)
#endif
__device__ __forceinline__ void ldg256(const uint4* p, uint4& a, uint4& b) {
#if CRB_FINE_STREAM & 2
    asm("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
        : "l"(__cvta_generic_to_global(p)));
}
__device__ __forceinline__ uint4 ldg128Record(const uint4* p) {
#if CRB_FINE_STREAM & 2
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(__cvta_generic_to_global(p)));
    return r;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ void stg256(uint4* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(__cvta_generic_to_global(p)), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
                 "r"(b.z), "r"(b.w)
                 : "memory");
}

// ---- conversions with PTX semantics -----------------------------------------------------------
__device__ __forceinline__ S32 f32ToS32SatRni(F32 a) { return __float2int_rn(a); }     // cvt.rni.sat.s32 (CUDA's float->int intrinsics saturate, NaN -> 0)
__device__ __forceinline__ U32 f32ToU32SatRni(F32 a) { return __float2uint_rn(a); }    // cvt.rni.sat.u32
__device__ __forceinline__ U32 f32ToU32SatRmi(F32 a) { return __float2uint_rd(a); }    // cvt.rmi.sat.u32
__device__ __forceinline__ U32 f32ToU32Rzi(F32 a) { return __float2uint_rz(a); }       // (U32)float

// ---- colour -------------------------------------------------------------------------------------
// Round-half-up of c*255 with clamping, evaluated as floor(c*255*2^24 + 2^23) >> 24 with the
// multiply-add and the conversion both rounding down (reference: cuda/Util.inl:30-38).
__device__ __forceinline__ U32 packUnorm8x4(F32 r, F32 g, F32 b, F32 a) {
    const F32 scale = 16777216.0f * 255.0f, half = 8388608.0f;
    U32 x = f32ToU32SatRmi(__fmaf_rd(r, scale, half));
    U32 y = f32ToU32SatRmi(__fmaf_rd(g, scale, half));
    U32 z = f32ToU32SatRmi(__fmaf_rd(b, scale, half));
    U32 w = f32ToU32SatRmi(__fmaf_rd(a, scale, half));
    return (x >> 24) | ((y >> 24) << 8) | ((z >> 24) << 16) | (w & 0xFF000000u);
}
__device__ __forceinline__ U32 toABGR(const Vec4f& c) { return packUnorm8x4(c.x, c.y, c.z, c.w); }
__device__ __forceinline__ U32 toABGR(float4 c) { return packUnorm8x4(c.x, c.y, c.z, c.w); }

// 8-bit fixed point blend: ((s*fs + d*fd) * 0x010101 + 0x800000) >> 24 per channel; the factors
// are taken from the top byte of the factor words (reference: cuda/Util.inl:42-60).
template <bool Clamp>
__device__ __forceinline__ U32 blendChannels(U32 src, U32 dst, U32 fsC, U32 fdC, U32 fsA, U32 fdA) {
    U32 out = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        U32 fs = (c == 3 ? fsA : fsC) >> 24, fd = (c == 3 ? fdA : fdC) >> 24;
        U32 t = ((src >> (8 * c)) & 0xFF) * fs + ((dst >> (8 * c)) & 0xFF) * fd;
        if (Clamp) t = min(t, 255u * 255u);
        out |= (((t * 0x010101u + 0x800000u) >> 24) & 0xFF) << (8 * c);
    }
    return out;
}
__device__ __forceinline__ U32 blendABGR(U32 src, U32 dst, U32 fsC, U32 fdC, U32 fsA, U32 fdA) { return blendChannels<false>(src, dst, fsC, fdC, fsA, fdA); }
__device__ __forceinline__ U32 blendABGRClamp(U32 src, U32 dst, U32 fsC, U32 fdC, U32 fsA, U32 fdA) { return blendChannels<true>(src, dst, fsC, fdC, fsA, fdA); }

#endif  // __CUDACC__

// ---- edge orientation code stored in triHeader.misc (reference: cuda/Util.hpp:207-217) ---------
__host__ __device__ __forceinline__ U32 cover8x8_selectFlips(S32 dx, S32 dy) {
    const U32 FY = 1u << CR_FLIPBIT_FLIP_Y, FX = 1u << CR_FLIPBIT_FLIP_X, SW = 1u << CR_FLIPBIT_SWAP_XY, CO = 1u << CR_FLIPBIT_COMPL;
    U32 f = 0;
    if (dy > 0 || (dy == 0 && dx <= 0)) f ^= FX ^ FY ^ CO;
    if (dx > 0) f ^= FX ^ FY;
    if (abs(dx) < abs(dy)) f ^= SW ^ FY;
    return f;
}

// Sample nearest to the pixel centre among the covered ones; -1 when all or none are covered
// (reference: cuda/Util.hpp:182-203).
__host__ __device__ __forceinline__ int selectMSAACentroid(int samplesLog2, U32 sampleMask) {
    int n = 1 << samplesLog2;
    if (sampleMask == 0 || sampleMask == (1u << n) - 1) return -1;
    int best = -1, bestDist = 0x7FFFFFFF;
    for (int i = 0; i < n; i++) {
        if (!((sampleMask >> i) & 1)) continue;
        int ax = msaaSampleX(samplesLog2, i) * 2 + 1 - n, ay = i * 2 + 1 - n;
        int d = ax * ax + ay * ay;
        if (d < bestDist) best = i, bestDist = d;
    }
    return best;
}

__host__ __device__ inline U32 encodeDepth(U32 depth) {  // reference: cuda/Util.hpp:290-298
    double v = (double)depth / (65536.0 * 65536.0 - 1.0);
    v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
    v = v * (double)(CR_DEPTH_MAX - CR_DEPTH_MIN) + (double)CR_DEPTH_MIN;
    return (U32)v;
}

#ifdef __CUDACC__

// ---- clipper: Sutherland-Hodgman in barycentric space (reference: cuda/Util.hpp:221-286) -------
// dist(b) = p0 + p1*b.x + p2*b.y evaluated as fma(p2, b.y, fma(p1, b.x, p0)).
__device__ __forceinline__ F32 clipPlaneDist(F32 p0, F32 p1, F32 p2, float2 b) { return __fmaf_rn(p2, b.y, __fmaf_rn(p1, b.x, p0)); }

static __device__ __noinline__ int clipPolygonWithPlane(float2* out, const float2* in, int n, F32 p0, F32 p1, F32 p2) {
    if (n < 3) return 0;
    int m = 0;
    float2 a = in[n - 1];
    F32 da = clipPlaneDist(p0, p1, p2, a);
    for (int i = 0; i < n; i++) {
        float2 b = in[i];
        F32 db = clipPlaneDist(p0, p1, p2, b);
        if (__fmul_rn(da, db) < 0.0f) {
            F32 tb = __fdiv_rn(da, __fsub_rn(da, db));
            F32 ta = __fsub_rn(1.0f, tb);
            out[m].x = __fmaf_rn(a.x, ta, __fmul_rn(b.x, tb));
            out[m].y = __fmaf_rn(a.y, ta, __fmul_rn(b.y, tb));
            m++;
        }
        if (db >= 0.0f) out[m++] = b;
        a = b;
        da = db;
    }
    return m;
}

// Clips against lo[a]*w <= v[a] <= hi[a]*w for a = x,y,z.  With lo = -1, hi = +1 every expression
// reduces exactly to the reference's  w < |x|,  w + x,  w - x.
static __device__ __noinline__ int clipTriangleWithWindow(float2* bary, float4 v0, float4 v1, float4 v2, float4 d1, float4 d2, const F32* lo, const F32* hi) {
    int n = 3;
    bary[0] = make_float2(0.0f, 0.0f);
    bary[1] = make_float2(1.0f, 0.0f);
    bary[2] = make_float2(0.0f, 1.0f);
    const F32 c0[3] = {v0.x, v0.y, v0.z}, c1[3] = {v1.x, v1.y, v1.z}, c2[3] = {v2.x, v2.y, v2.z};
    const F32 e1[3] = {d1.x, d1.y, d1.z}, e2[3] = {d2.x, d2.y, d2.z};
    float2 tmp[9];
    for (int a = 0; a < 3; a++) {
        bool any = (__fmul_rn(v0.w, hi[a]) < c0[a]) | (__fmul_rn(v0.w, lo[a]) > c0[a]) | (__fmul_rn(v1.w, hi[a]) < c1[a]) | (__fmul_rn(v1.w, lo[a]) > c1[a]) |
                   (__fmul_rn(v2.w, hi[a]) < c2[a]) | (__fmul_rn(v2.w, lo[a]) > c2[a]);
        if (!any) continue;
        n = clipPolygonWithPlane(tmp, bary, n, __fmaf_rn(-lo[a], v0.w, c0[a]), __fmaf_rn(-lo[a], d1.w, e1[a]), __fmaf_rn(-lo[a], d2.w, e2[a]));
        n = clipPolygonWithPlane(bary, tmp, n, __fmaf_rn(hi[a], v0.w, -c0[a]), __fmaf_rn(hi[a], d1.w, -e1[a]), __fmaf_rn(hi[a], d2.w, -e2[a]));
    }
    return n;
}

// ---- plane equation in fixed point (reference: cuda/Util.inl:65-91) ----------------------------
// values at the three vertices -> (x slope, y slope, constant) such that
// plane(sx, sy) = x*sx + y*sy + z in sample units (samplesLog2 selects the unit).  All integer
// after the initial float -> U32 truncation; S64 intermediates; shifts are arithmetic.
// Integer part shared by the two entry points below: t0 = value at vertex 0, t1 / t2 = differences to it (all >> sh).
__device__ __forceinline__ uint3 setupPleqFixed(S32 t0, S32 t1, S32 t2, int sh, int2 v0, int2 d1, int2 d2, F32 areaRcp, int samplesLog2) {
    U32 rcpMant = ((U32)__float_as_int(areaRcp) & 0x007FFFFFu) | 0x00800000u;
    int rcpShift = (23 + 127) - (__float_as_int(areaRcp) >> 23);

    S64 xc = ((S64)t1 * d2.y - (S64)t2 * d1.y) * (S64)rcpMant;
    S64 yc = ((S64)t2 * d1.x - (S64)t1 * d2.x) * (S64)rcpMant;
    const int sub = CR_SUBPIXEL_LOG2 - samplesLog2;
    uint3 p;
    p.x = (U32)(xc >> (rcpShift - (sh + sub)));
    p.y = (U32)(yc >> (rcpShift - (sh + sub)));

    S32 cx = (v0.x * 2 + min(min(d1.x, d2.x), 0) + max(max(d1.x, d2.x), 0)) >> (sub + 1);
    S32 cy = (v0.y * 2 + min(min(d1.y, d2.y), 0) + max(max(d1.y, d2.y), 0)) >> (sub + 1);
    S32 vcx = v0.x - (cx << sub);
    S32 vcy = v0.y - (cy << sub);

    p.z = (U32)t0 << sh;
    p.z -= (U32)(((xc >> 13) * vcx + (yc >> 13) * vcy) >> (rcpShift - (sh + 13)));
    p.z -= p.x * (U32)cx + p.y * (U32)cy;
    return p;
}

__device__ __forceinline__ uint3 setupPleq(float3 values, int2 v0, int2 d1, int2 d2, F32 areaRcp, int samplesLog2) {
    F32 mx = fmaxf(fmaxf(values.x, values.y), values.z);
    int sh = min(max((__float_as_int(mx) >> 23) - (127 + 22), 0), 8);
    S32 t0 = (S32)(f32ToU32Rzi(values.x) >> sh);
    S32 t1 = (S32)((f32ToU32Rzi(values.y) >> sh) - (U32)t0);
    S32 t2 = (S32)((f32ToU32Rzi(values.z) >> sh) - (U32)t0);
    return setupPleqFixed(t0, t1, t2, sh, v0, d1, d2, areaRcp, samplesLog2);
}

// The plane through (0, value, 0) (Which = 1) or (0, 0, value) (Which = 2): the u / v planes of an UNCLIPPED triangle, whose
// barycentrics at the vertices are (0,0), (1,0), (0,1) -- the reference multiplies them out (TriangleSetup.inl:160-166:
// 0 * w0', 1 * w1', 0 * w2') and gets exactly these values: 0 * x is +-0 (or NaN for a non-finite x), which the float -> U32
// conversion turns into 0 and fmaxf ignores next to a value >= 0; an all-NaN triple gives the zero plane on both routes.
template <int Which>
__device__ __forceinline__ uint3 setupPleqUnit(F32 value, int2 v0, int2 d1, int2 d2, F32 areaRcp, int samplesLog2) {
    F32 mx = fmaxf(0.0f, value);
    int sh = min(max((__float_as_int(mx) >> 23) - (127 + 22), 0), 8);
    const S32 t = (S32)(f32ToU32Rzi(value) >> sh);
    return setupPleqFixed(0, Which == 1 ? t : 0, Which == 2 ? t : 0, sh, v0, d1, d2, areaRcp, samplesLog2);
}

// ---- programmatic dependent launch (sm_90+) --------------------------------------------------------
// Every kernel of a frame is launched with cudaLaunchAttributeProgrammaticStreamSerialization: the
// next kernel's CTAs become resident (launch latency, prologue, shared-memory initialisation) while
// the previous kernel drains, and block in gridDepWait() until the previous grid has completed and its
// memory is visible.  Nothing a kernel does BEFORE gridDepWait() may touch frame memory.
__device__ __forceinline__ void gridDepLaunchDependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void gridDepWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// (static: every shared object must launch through ITS OWN copy -- nvcc links the CUDA runtime statically
// into each of them, and a kernel can only be launched by the runtime instance it is registered with)
// clusterSize > 1: the grid is launched as thread-block clusters of that many CTAs (distributed shared memory between them).
template <class Kernel>
static inline cudaError_t launchChained(Kernel kernel, int grid, int block, cudaStream_t stream, const crb_frame& f, int clusterSize = 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (f.chainLaunches) {   // 0 = plain stream order (CRB_NO_PDL=1, a debugging aid)
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        n++;
    }
    if (clusterSize > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)clusterSize;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        n++;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, f);
}

// ---- ProfilingMode_Counters (reference: CR_COUNT, cuda/Util.hpp) -------------------------------------------------
// num / den are added to the counter's numerator / denominator; compiles to nothing in the other modes.
template <int ProfMode>
__device__ __forceinline__ void profCount(const crb_frame& f, int counter, unsigned long long num, unsigned long long den) {
    if (ProfMode == ProfilingMode_Counters) {
        if (num) atomicAdd(&f.profCounters[2 * counter], num);
        if (den) atomicAdd(&f.profCounters[2 * counter + 1], den);
    }
}
// The same for a predicate evaluated by every lane of a CONVERGED warp: one pair of atomics per warp.
template <int ProfMode>
__device__ __forceinline__ void profCountWarp(const crb_frame& f, int counter, bool numPred, bool denPred) {
    if (ProfMode == ProfilingMode_Counters) {
        const unsigned n = __popc(__ballot_sync(0xFFFFFFFFu, numPred)), d = __popc(__ballot_sync(0xFFFFFFFFu, denPred));
        if ((threadIdx.x & 31) == 0) profCount<ProfMode>(f, counter, 100ull * n, d);
    }
}

// ---- ProfilingMode_Timers (reference: CR_TIMER_IN / CR_TIMER_OUT) -----------------------------------------------
// Lane 0 of a warp brackets a code region with clock64() and adds the difference to the timer's total.
template <int ProfMode>
struct ProfTimer {
    long long t0;
    __device__ __forceinline__ void start() {
        if (ProfMode == ProfilingMode_Timers) t0 = clock64();
    }
    __device__ __forceinline__ void stop(const crb_frame& f, int timer) {
        if (ProfMode == ProfilingMode_Timers) {
            const long long d = clock64() - t0;
            if ((threadIdx.x & 31) == 0 && d > 0) atomicAdd(&f.profCounters[2 * CRB_PROF_NUM + timer], (unsigned long long)d);
        }
    }
};

// ---- warp helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ U32 laneId() { return threadIdx.x & 31; }
__device__ __forceinline__ U32 laneMaskLt() { U32 r; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(r)); return r; }

#endif  // __CUDACC__

}  // namespace FW
