// Stage 1 -- triangle setup + per-chunk bin histogram, hand-written for sm_100a.
//
// Computes what the reference's triangleSetupImpl computes (src/cudaraster/cuda/
// TriangleSetup.inl:19-417): frustum cull, subpixel snap (cvt.rni.sat), backface and
// between-samples cull, guard-band test, frustum clipping into <= 7 sub-triangles, integer
// plane equations, and the triSubtris / triHeader / triData records -- bit for bit.
//
// What is different (B200 design, DESIGN.md "setup"):
//  * one thread per input triangle; a CTA (= 256 consecutive triangles, one CHUNK or a slice of
//    one) also histograms the bins its sub-triangles touch in shared memory and publishes one
//    column of the [bin][chunk] count matrix.  That column is all the bin stage needs to place the
//    chunk's queue entries with a plain scan -- no 16-CTA bin kernel, no segment lists;
//  * the frame parameters travel as a __grid_constant__ kernel argument, not through
//    __constant__ uploads; vertices are read through the read-only path with 128-bit loads;
//  * the clipper lives in a __noinline__ cold path so the common path needs no local stack.
#pragma once
#include "Overlap.cuh"

#ifndef CRB_PLEQ_UNCLIPPED
#define CRB_PLEQ_UNCLIPPED 1  // u / v planes of an unclipped triangle: two of the three vertex values are exactly zero
#endif

namespace FW {

struct SnappedTri {
    int2 p0, p1, p2, lo, hi;
    float3 rcpW;
};

__device__ __forceinline__ void snapTriangle(const crb_frame& f, float4 v0, float4 v1, float4 v2, SnappedTri& s) {
    const F32 sx = (F32)(f.fullWidth << (CR_SUBPIXEL_LOG2 - 1));
    const F32 sy = (F32)(f.fullHeight << (CR_SUBPIXEL_LOG2 - 1));
    s.rcpW = make_float3(__frcp_rn(v0.w), __frcp_rn(v1.w), __frcp_rn(v2.w));
    s.p0 = make_int2(f32ToS32SatRni(__fmul_rn(__fmul_rn(v0.x, s.rcpW.x), sx)) - f.centerOfsX, f32ToS32SatRni(__fmul_rn(__fmul_rn(v0.y, s.rcpW.x), sy)) - f.centerOfsY);
    s.p1 = make_int2(f32ToS32SatRni(__fmul_rn(__fmul_rn(v1.x, s.rcpW.y), sx)) - f.centerOfsX, f32ToS32SatRni(__fmul_rn(__fmul_rn(v1.y, s.rcpW.y), sy)) - f.centerOfsY);
    s.p2 = make_int2(f32ToS32SatRni(__fmul_rn(__fmul_rn(v2.x, s.rcpW.z), sx)) - f.centerOfsX, f32ToS32SatRni(__fmul_rn(__fmul_rn(v2.y, s.rcpW.z), sy)) - f.centerOfsY);
    s.lo = make_int2(min(min(s.p0.x, s.p1.x), s.p2.x), min(min(s.p0.y, s.p1.y), s.p2.y));
    s.hi = make_int2(max(max(s.p0.x, s.p1.x), s.p2.x), max(max(s.p0.y, s.p1.y), s.p2.y));
}

// 0 = visible, 1 = backfacing / degenerate, 2 = no sample inside the AABB.
template <int SamplesLog2>
__device__ __forceinline__ int prepareTriangle(const crb_frame& f, const SnappedTri& s, int2& d1, int2& d2, S32& area) {
    d1 = make_int2(s.p1.x - s.p0.x, s.p1.y - s.p0.y);
    d2 = make_int2(s.p2.x - s.p0.x, s.p2.y - s.p0.y);
    area = d1.x * d2.y - d1.y * d2.x;
    if (area <= 0) return 1;

    const int sampleSize = 1 << (CR_SUBPIXEL_LOG2 - SamplesLog2);
    const S32 biasX = (f.viewportWidth << (CR_SUBPIXEL_LOG2 - 1)) - (sampleSize >> 1);
    const S32 biasY = (f.viewportHeight << (CR_SUBPIXEL_LOG2 - 1)) - (sampleSize >> 1);
    const S32 lox = (s.lo.x + (sampleSize - 1) + biasX) & -sampleSize;
    const S32 loy = (s.lo.y + (sampleSize - 1) + biasY) & -sampleSize;
    const S32 hix = (s.hi.x + biasX) & -sampleSize;
    const S32 hiy = (s.hi.y + biasY) & -sampleSize;
    if (lox > hix || loy > hiy) return 2;

    const S32 diff = hix + hiy - lox - loy;
    if (diff <= sampleSize) {
        // the AABB holds one or two sample points: test them exactly (no fill-rule bias here)
        S32 qx = lox, qy = loy;
        for (int pass = 0;; pass++) {
            int2 t0 = make_int2(s.p0.x + biasX - qx, s.p0.y + biasY - qy);
            int2 t1 = make_int2(s.p1.x + biasX - qx, s.p1.y + biasY - qy);
            int2 t2 = make_int2(s.p2.x + biasX - qx, s.p2.y + biasY - qy);
            S32 e0 = t0.x * t1.y - t0.y * t1.x;
            S32 e1 = t1.x * t2.y - t1.y * t2.x;
            S32 e2 = t2.x * t0.y - t2.y * t0.x;
            if (!(e0 < 0 || e1 < 0 || e2 < 0)) break;
            if (pass == 1 || diff == 0) return 2;
            qx = hix, qy = hiy;
        }
    }
    return 0;
}

// Writes one sub-triangle record and returns its packed header (for the bin histogram).
template <int SamplesLog2, U32 RenderModeFlags, bool Unclipped = false>
__device__ __forceinline__ uint4 setupTriangle(const crb_frame& f, uint4* th, uint4* td, int3 vidx, float4 v0, float4 v1, float4 v2,
                                               float2 b0, float2 b1, float2 b2, const SnappedTri& s, int2 d1, int2 d2, S32 area, uint3* zpOut = nullptr,
                                               bool microOnly = false) {
    // microOnly: the triangle is rasterized by setup itself (micro mode) and never queued, so nothing will read its header
    // or its depth-plane row: only the shading rows (w/u/v planes, vertex ids) are produced.
    F32 areaRcp = 0.0f;
    int2 wv0 = make_int2(0, 0);
    uint4 row0 = make_uint4(0, 0, 0, 0), row1 = row0, row2 = row0, row3 = row0;   // the triData record
    if ((RenderModeFlags & (CRB_FLAG_DEPTH | CRB_FLAG_LERP)) != 0) {
        areaRcp = __frcp_rn((F32)area);
        // Plane equations are set up in viewport-corner coordinates (the reference's wv0) and, for
        // a sort-first window, translated to the surface below: every window of one parent
        // viewport therefore renders the very same depth / barycentrics as the unsplit viewport.
        wv0.x = s.p0.x + (f.viewportWidth << (CR_SUBPIXEL_LOG2 - 1));
        wv0.y = s.p0.y + (f.viewportHeight << (CR_SUBPIXEL_LOG2 - 1));
    }

    U32 zmin = 0;
    if ((RenderModeFlags & CRB_FLAG_DEPTH) != 0) {
        const F32 zcoef = (F32)(CR_DEPTH_MAX - CR_DEPTH_MIN) * 0.5f;
        const F32 zbias = (F32)(CR_DEPTH_MAX + CR_DEPTH_MIN) * 0.5f;
        float3 zvert;
        zvert.x = __fmaf_rn(__fmul_rn(v0.z, zcoef), s.rcpW.x, zbias);
        zvert.y = __fmaf_rn(__fmul_rn(v1.z, zcoef), s.rcpW.y, zbias);
        zvert.z = __fmaf_rn(__fmul_rn(v2.z, zcoef), s.rcpW.z, zbias);
        int2 zv0 = make_int2(wv0.x - (1 << (CR_SUBPIXEL_LOG2 - SamplesLog2 - 1)), wv0.y - (1 << (CR_SUBPIXEL_LOG2 - SamplesLog2 - 1)));
        uint3 zp = setupPleq(zvert, zv0, d1, d2, areaRcp, SamplesLog2);
        if (!microOnly) zmin = f32ToU32SatRni(__fsub_rn(fminf(fminf(zvert.x, zvert.y), zvert.z), (F32)CR_LERP_ERROR(SamplesLog2)));
        U32 zslope = 0;
        if (SamplesLog2 != 0) {
            S32 ax = (S32)zp.x; ax = ax >= 0 ? ax : -ax;
            S32 ay = max((S32)zp.y, -FW_S32_MAX); ay = ay >= 0 ? ay : -ay;
            U32 tmp = (U32)ax + (U32)ay;
            const int k = SamplesLog2 - 1 > 0 ? SamplesLog2 - 1 : 0;
            zslope = tmp << k;
            if ((zslope >> k) != tmp) zslope = FW_U32_MAX;
        }
        zp.z += zp.x * ((U32)f.subX0 << SamplesLog2) + zp.y * ((U32)f.subY0 << SamplesLog2);
        row0 = make_uint4(zp.x, zp.y, zp.z, zslope);
        if (zpOut) *zpOut = zp;
    }

    if ((RenderModeFlags & CRB_FLAG_LERP) != 0) {
        F32 wcoef = __fmul_rn(fminf(fminf(v0.w, v1.w), v2.w), (F32)CR_BARY_MAX);
        float3 wvert = make_float3(__fmul_rn(wcoef, s.rcpW.x), __fmul_rn(wcoef, s.rcpW.y), __fmul_rn(wcoef, s.rcpW.z));
        float3 uvert = make_float3(__fmul_rn(b0.x, wvert.x), __fmul_rn(b1.x, wvert.y), __fmul_rn(b2.x, wvert.z));
        float3 vvert = make_float3(__fmul_rn(b0.y, wvert.x), __fmul_rn(b1.y, wvert.y), __fmul_rn(b2.y, wvert.z));
        uint3 wp = setupPleq(wvert, wv0, d1, d2, areaRcp, SamplesLog2 + 1);
        uint3 up, vp;
        if (CRB_PLEQ_UNCLIPPED && Unclipped) {   // b = (0,0), (1,0), (0,1): uvert = (0, wvert.y, 0), vvert = (0, 0, wvert.z)
            up = setupPleqUnit<1>(wvert.y, wv0, d1, d2, areaRcp, SamplesLog2 + 1);
            vp = setupPleqUnit<2>(wvert.z, wv0, d1, d2, areaRcp, SamplesLog2 + 1);
        } else {
            up = setupPleq(uvert, wv0, d1, d2, areaRcp, SamplesLog2 + 1);
            vp = setupPleq(vvert, wv0, d1, d2, areaRcp, SamplesLog2 + 1);
        }
        const U32 ox2 = (U32)f.subX0 << (SamplesLog2 + 1), oy2 = (U32)f.subY0 << (SamplesLog2 + 1);
        wp.z += wp.x * ox2 + wp.y * oy2;
        up.z += up.x * ox2 + up.y * oy2;
        vp.z += vp.x * ox2 + vp.y * oy2;
        row1 = make_uint4(wp.x, wp.y, wp.z, up.x);
        row2 = make_uint4(up.y, up.z, vp.x, vp.y);
        row3 = make_uint4(vp.z, (U32)vidx.x, (U32)vidx.y, (U32)vidx.z);
    } else {
        row3 = make_uint4(0u, (U32)vidx.x, (U32)vidx.y, (U32)vidx.z);
    }
    if (CRB_WIDE_ST && Unclipped && (CRB_WIDE_ST > 1 || microOnly)) {
        // A micro triangle's record leaves as two whole 32-byte sectors, depth row included although nothing reads it: a
        // half-written sector costs a DRAM fill when L2 evicts it (C2: fine raster 39.1 -> 37.4 us, setup unchanged).  Records
        // that carry all four rows anyway stay on four 128-bit stores: as 256-bit stores C3's setup took 252 instead of 225 us.
        // Fast path only: inside a non-inlined function (the clipper's cold path) ptxas 12.9 turns st.global.v8.b32 into a
        // 32-bit store of the first element (found by the MSAA soup: words 1-7 of every clipped record came out zero).
        stg256(td, row0, row1);
        stg256(td + 2, row2, row3);
    } else {
        if ((RenderModeFlags & CRB_FLAG_DEPTH) != 0 && !microOnly) td[0] = row0;
        if ((RenderModeFlags & CRB_FLAG_LERP) != 0) { td[1] = row1; td[2] = row2; }
        td[3] = row3;
    }

    if (microOnly) return make_uint4(0, 0, 0, 0);
    const U32 f01 = cover8x8_selectFlips(d1.x, d1.y);
    const U32 f12 = cover8x8_selectFlips(d2.x - d1.x, d2.y - d1.y);
    const U32 f20 = cover8x8_selectFlips(-d2.x, -d2.y);
    uint4 h = make_uint4(((U32)s.p0.x & 0xFFFFu) | ((U32)s.p0.y << 16), ((U32)s.p1.x & 0xFFFFu) | ((U32)s.p1.y << 16),
                         ((U32)s.p2.x & 0xFFFFu) | ((U32)s.p2.y << 16), (zmin & 0xFFFFF000u) | (f01 << 6) | (f12 << 2) | (f20 >> 2));
    *th = h;
    return h;
}

// Fire-and-forget 64-bit minimum on a GLOBAL address, under a predicate.  (atomicMin through a pointer read from the frame
// block inside a non-inlined function is a GENERIC atomic: nvcc emits ATOM.E.MIN.64 with a predicate result, waits for it, and
// branches into a shared-memory CAS fallback -- one L2 round trip per covered pixel on the thread's critical path.  red.global
// has no result and no window check: REDG.E.MIN.64.)  v = hi << 32 | lo.
__device__ __forceinline__ void redMinGlobalU64If(bool cond, unsigned long long* gptr, U32 hi, U32 lo) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 v;\n\tsetp.ne.b32 p, %0, 0;\n\tmov.b64 v, {%3, %2};\n\t@p red.global.min.u64 [%1], v;\n\t}" ::"r"((U32)cond), "l"(gptr),
        "r"(hi), "r"(lo));
}

// Micro-triangle path (crb_frame::microMode): a sub-triangle whose pixel-centre footprint is at most 4x4 pixels is
// rasterized right here -- exact coverage of its <= 16 candidate pixels (the fine raster's own small-triangle
// evaluation, FineRaster.cuh coverSmall4x4), plane depth per covered pixel, 64-bit minimum of
// (depth << 32 | entry + 1) into the visibility buffer -- and is never queued.  Inlined into the setup kernel: out of line it
// re-read the frame fields through generic loads and moved the global-memory descriptor into uniform registers before every
// reduction (C2 setup 39.8 -> 37.7 us inlined, at the same 48 registers).  (x*, y*) = snapped vertices, viewport-centred
// subpixels (what the header would hold); (pxLo*, n*) = its pixel rectangle in surface pixels.
static __device__ __forceinline__ void microRaster(const crb_frame& f, S32 x0, S32 y0, S32 x1, S32 y1, S32 x2, S32 y2, U32 zx, U32 zy, U32 zb, S32 entry, S32 pxLoX,
                                                   S32 pxLoY, int nx, int ny) {
    // centre of pixel (pxLoX, pxLoY) in viewport-centred subpixels
    const S32 px = (pxLoX << CR_SUBPIXEL_LOG2) + (CR_SUBPIXEL_SIZE >> 1) - f.originX, py = (pxLoY << CR_SUBPIXEL_LOG2) + (CR_SUBPIXEL_SIZE >> 1) - f.originY;
    const S32 dx0 = x1 - x0, dy0 = y1 - y0, dx1 = x2 - x1, dy1 = y2 - y1, dx2 = x0 - x2, dy2 = y0 - y2;
    S32 e0 = (x0 - px) * dy0 - (y0 - py) * dx0 - ((dy0 > 0 || (dy0 == 0 && dx0 <= 0)) ? 1 : 0);
    S32 e1 = (x1 - px) * dy1 - (y1 - py) * dx1 - ((dy1 > 0 || (dy1 == 0 && dx1 <= 0)) ? 1 : 0);
    S32 e2 = (x0 - px) * dy2 - (y0 - py) * dx2 - ((dy2 > 0 || (dy2 == 0 && dx2 <= 0)) ? 1 : 0);
    const S32 a0 = -(dy0 << CR_SUBPIXEL_LOG2), a1 = -(dy1 << CR_SUBPIXEL_LOG2), a2 = -(dy2 << CR_SUBPIXEL_LOG2);
    const S32 b0 = dx0 << CR_SUBPIXEL_LOG2, b1 = dx1 << CR_SUBPIXEL_LOG2, b2 = dx2 << CR_SUBPIXEL_LOG2;
    const size_t pitch = (size_t)f.widthPixels;
    unsigned long long* row = reinterpret_cast<unsigned long long*>(__cvta_generic_to_global(f.visBuffer)) + (size_t)pxLoY * pitch + pxLoX;
    U32 zrow = zb + zx * (U32)pxLoX + zy * (U32)pxLoY;
    const U32 id = (U32)(entry + 1);
#pragma unroll 1
    for (int r = 0; r < ny; r++) {
        // column c of the row: e_i + c * a_i; the sign of the OR of the three is the coverage test
        const S32 m0 = e0 | e1 | e2;
        const S32 m1 = (e0 + a0) | (e1 + a1) | (e2 + a2);
        const S32 m2 = (e0 + 2 * a0) | (e1 + 2 * a1) | (e2 + 2 * a2);
        const S32 m3 = (e0 + 3 * a0) | (e1 + 3 * a1) | (e2 + 3 * a2);
        redMinGlobalU64If(m0 >= 0, row, zrow, id);
        redMinGlobalU64If((m1 >= 0) & (nx > 1), row + 1, zrow + zx, id);
        redMinGlobalU64If((m2 >= 0) & (nx > 2), row + 2, zrow + 2 * zx, id);
        redMinGlobalU64If((m3 >= 0) & (nx > 3), row + 3, zrow + 3 * zx, id);
        e0 += b0; e1 += b1; e2 += b2;
        zrow += zy;
        row += pitch;
    }
}

// What binning needs from a finished sub-triangle (header h, queue entry `entry`, stored in record `slot`).  General path: the
// bins it touches go into the CTA's bin histogram.  Direct tile path (crb_frame::directMode): every tile it touches is counted
// straight into the per-tile counters with fire-and-forget global reductions; sub-triangles that span more than
// CRB_DIRECT_MAX_TILES tiles on an axis go on a GLOBAL list (crb_frame::largeList) instead: the queue-allocation kernel counts
// them and the scatter kernel places them with one whole CTA per triangle, all SMs sharing the list -- a thread walking
// thousands of tiles alone, or the one CTA that happens to hold the scene's big triangles, would be a long tail.
struct SetupCtaShared {   // per CTA
    int binCount[CR_MAXBINS_SQR];   // bin histogram of the CTA's chunk (general path)
};
struct SetupShared {      // a VIEW of the binning scratch in shared memory (passed by value: two pointers)
    int* queuedAny;        // direct path: a sub-triangle of this CTA went to the tile counters (or onto the large list)
    SetupCtaShared* cta;
};

// Returns the triangle's tile code (crb_frame::triTileCode) on the direct path, 0 on the general path.
// DeferSmall: the caller counts a footprint of at most 2x2 tiles itself from the returned code.  (Unused: counting
// warp-aggregated with __match_any_sync at the end of the kernel measured 37.6 vs 38.5 us on C2 but 225 vs 215 us on C4.)
template <int SamplesLog2, bool DeferSmall>
__device__ __forceinline__ U32 histogramBins(const crb_frame& f, uint4 h, S32 entry, int slot, const SetupShared sh) {
    int* s_binCount = sh.cta->binCount;
    TriFootprint fp = triFootprint<SamplesLog2>(h.x, h.y, h.z, f);
    if (f.directMode) {
        const CellRange t = cellRange<CR_TILE_LOG2>(fp, 0, 0, f.widthTiles - 1, f.heightTiles - 1);
        *sh.queuedAny = 1;
        if ((t.nx > CRB_DIRECT_MAX_TILES) | (t.ny > CRB_DIRECT_MAX_TILES)) {
            const int k = atomicAdd(&f.atomics->numLargeTris, 1);
            if (k < f.maxLarge) f.largeList[k] = make_int2(entry, slot);
            else atomicOr(&f.atomics->overflow, 32);   // counted, not listed: the host grows the list and reruns the frame
            return CRB_TILECODE_GENERAL;
        }
        if (!t.refine) {   // at most 2x2 tiles, never refined: the rectangle IS the tile set
            if (t.nx > 0) {
                if (!DeferSmall) {
                    int* p = &f.tileCounter[t.x0 + t.y0 * f.widthTiles];
                    atomicAdd(p, 1);
                    if (t.nx > 1) atomicAdd(p + 1, 1);
                    if (t.ny > 1) {
                        atomicAdd(p + f.widthTiles, 1);
                        if (t.nx > 1) atomicAdd(p + f.widthTiles + 1, 1);
                    }
                }
                return 0x80000000u | (U32)t.x0 | ((U32)t.y0 << 8) | ((U32)(t.nx - 1) << 16) | ((U32)(t.ny - 1) << 17);
            }
            return 0;
        }
        forEachCell<SamplesLog2, CR_TILE_LOG2>(fp, 0, 0, f.widthTiles - 1, f.heightTiles - 1, [&](S32 tx, S32 ty) { atomicAdd(&f.tileCounter[tx + ty * f.widthTiles], 1); });
        return CRB_TILECODE_GENERAL;
    }
    // footprints of at most 2x2 bins are never refined (Overlap.cuh): the rectangle IS the cell set
    const CellRange r = cellRange<CR_BIN_LOG2 + CR_TILE_LOG2>(fp, 0, 0, f.widthBins - 1, f.heightBins - 1);
    if (!r.refine) {
        if (r.nx > 0) {
            int* p = &s_binCount[r.x0 + r.y0 * f.widthBins];
            atomicAdd(p, 1);
            if (r.nx > 1) atomicAdd(p + 1, 1);
            if (r.ny > 1) {
                atomicAdd(p + f.widthBins, 1);
                if (r.nx > 1) atomicAdd(p + f.widthBins + 1, 1);
            }
        }
        return 0;
    }
    forEachCell<SamplesLog2, CR_BIN_LOG2 + CR_TILE_LOG2>(fp, 0, 0, f.widthBins - 1, f.heightBins - 1,
                                                          [&](S32 bx, S32 by) { atomicAdd(&s_binCount[bx + by * f.widthBins], 1); });
    return 0;
}

// Cold path: clip against the view window, fan the polygon, re-snap / re-cull every sub-triangle.
template <int SamplesLog2, U32 RenderModeFlags>
__device__ __noinline__ int setupClippedTriangle(const crb_frame& f, int tri, int3 vidx, float4 v0, float4 v1, float4 v2, const SetupShared sh) {
    const F32 lo[3] = {f.clipLoX, f.clipLoY, -1.0f}, hi[3] = {f.clipHiX, f.clipHiY, 1.0f};
    const float4 d1 = make_float4(__fsub_rn(v1.x, v0.x), __fsub_rn(v1.y, v0.y), __fsub_rn(v1.z, v0.z), __fsub_rn(v1.w, v0.w));
    const float4 d2 = make_float4(__fsub_rn(v2.x, v0.x), __fsub_rn(v2.y, v0.y), __fsub_rn(v2.z, v0.z), __fsub_rn(v2.w, v0.w));
    float2 bary[9];
    int numVerts = clipTriangleWithWindow(bary, v0, v1, v2, d1, d2, lo, hi);

    auto vertexAt = [&](int i) {
        float2 b = bary[i];
        return make_float4(__fmaf_rn(d2.x, b.y, __fmaf_rn(d1.x, b.x, v0.x)), __fmaf_rn(d2.y, b.y, __fmaf_rn(d1.y, b.x, v0.y)),
                           __fmaf_rn(d2.z, b.y, __fmaf_rn(d1.z, b.x, v0.z)), __fmaf_rn(d2.w, b.y, __fmaf_rn(d1.w, b.x, v0.w)));
    };

    SnappedTri s;
    int2 e1, e2;
    S32 area;
    int numSub = 0;
    if (numVerts >= 3) {
        const float4 c0 = vertexAt(0);
        float4 cPrev = vertexAt(1);
        for (int i = 2; i < numVerts; i++) {
            float4 cCur = vertexAt(i);
            snapTriangle(f, c0, cPrev, cCur, s);
            if (prepareTriangle<SamplesLog2>(f, s, e1, e2, area) == 0) numSub++;
            cPrev = cCur;
        }
    }
    f.triSubtris[tri] = (U8)numSub;
    if (numSub == 0) return 0;

    // 0/1 survivors live in slot `tri`; more take a contiguous run from the global cursor
    // (allocation order is non-deterministic, consumers go through triHeader[tri].misc).
    int slot = tri;
    if (numSub > 1) {
        // the device counter holds the EXTRA slots; slot numbering starts at numTris like CRAtomics.numSubtris
        slot = f.numTris + atomicAdd(&f.atomics->numSubtris, numSub);
        reinterpret_cast<U32*>(&f.triHeader[tri])[3] = (U32)slot;
        if (slot + numSub > f.maxSubtris) {   // overflow: counted, not written; the host grows the buffers and reruns
            atomicOr(&f.atomics->overflow, 1);
            return numSub;
        }
    }
    const float4 c0 = vertexAt(0);
    float4 cPrev = vertexAt(1);
    int emitted = 0;   // index of the sub-triangle among the survivors = the `sub` of its queue entry (reference: BinRaster.inl:148-157)
    for (int i = 2; i < numVerts; i++) {
        float4 cCur = vertexAt(i);
        snapTriangle(f, c0, cPrev, cCur, s);
        if (prepareTriangle<SamplesLog2>(f, s, e1, e2, area) == 0) {
            uint4 h = setupTriangle<SamplesLog2, RenderModeFlags>(f, &f.triHeader[slot], &f.triData[(size_t)slot * 4], vidx, c0, cPrev, cCur, bary[0], bary[i - 1], bary[i], s, e1, e2, area);
            histogramBins<SamplesLog2, false>(f, h, numSub > 1 ? tri * 8 + emitted : tri * 8 + 7, slot, sh);
            slot++;
            emitted++;
        }
        cPrev = cCur;
    }
    return numSub;
}

// One input triangle, vertices in registers: cull / snap / cull / plane equations / records / binning.  Returns the triangle's
// tile code (direct path).  prof: ProfilingMode_Counters bits (1 viewport cull, 2 backface cull, 3 between-pixels cull, 4 clipped, 5 survived).
template <int SamplesLog2, U32 RenderModeFlags, int ProfMode>
__device__ __forceinline__ U32 setupOneTriangle(const crb_frame& f, int tri, int3 vidx, float4 v0, float4 v1, float4 v2, const SetupShared sh, U32& prof, ProfTimer<ProfMode>& tm) {
    const S32 aabbLimit = (1 << (CR_MAXVIEWPORT_LOG2 + CR_SUBPIXEL_LOG2)) - 1;
    U32 tileCode = 0;
    // all three vertices outside one plane of the view window -> culled
    const F32 wx0h = __fmul_rn(v0.w, f.clipHiX), wx1h = __fmul_rn(v1.w, f.clipHiX), wx2h = __fmul_rn(v2.w, f.clipHiX);
    const F32 wx0l = __fmul_rn(v0.w, f.clipLoX), wx1l = __fmul_rn(v1.w, f.clipLoX), wx2l = __fmul_rn(v2.w, f.clipLoX);
    const F32 wy0h = __fmul_rn(v0.w, f.clipHiY), wy1h = __fmul_rn(v1.w, f.clipHiY), wy2h = __fmul_rn(v2.w, f.clipHiY);
    const F32 wy0l = __fmul_rn(v0.w, f.clipLoY), wy1l = __fmul_rn(v1.w, f.clipLoY), wy2l = __fmul_rn(v2.w, f.clipLoY);
    bool outside = ((wx0h < v0.x) & (wx1h < v1.x) & (wx2h < v2.x)) | ((wx0l > v0.x) & (wx1l > v1.x) & (wx2l > v2.x)) |
                   ((wy0h < v0.y) & (wy1h < v1.y) & (wy2h < v2.y)) | ((wy0l > v0.y) & (wy1l > v1.y) & (wy2l > v2.y)) |
                   ((v0.w < v0.z) & (v1.w < v1.z) & (v2.w < v2.z)) | ((v0.w < -v0.z) & (v1.w < -v1.z) & (v2.w < -v2.z));
    if (f.windowed) {
        // sort-first window: also cull what lies wholly outside the surface rectangle (a pure cull:
        // a half-space that holds all three vertices holds the triangle, whatever the sign of w)
        outside |= ((__fmul_rn(v0.w, f.cullHiX) < v0.x) & (__fmul_rn(v1.w, f.cullHiX) < v1.x) & (__fmul_rn(v2.w, f.cullHiX) < v2.x)) |
                   ((__fmul_rn(v0.w, f.cullLoX) > v0.x) & (__fmul_rn(v1.w, f.cullLoX) > v1.x) & (__fmul_rn(v2.w, f.cullLoX) > v2.x)) |
                   ((__fmul_rn(v0.w, f.cullHiY) < v0.y) & (__fmul_rn(v1.w, f.cullHiY) < v1.y) & (__fmul_rn(v2.w, f.cullHiY) < v2.y)) |
                   ((__fmul_rn(v0.w, f.cullLoY) > v0.y) & (__fmul_rn(v1.w, f.cullLoY) > v1.y) & (__fmul_rn(v2.w, f.cullLoY) > v2.y));
    }
    tm.stop(f, CRB_TIMER_SetupVertexRead);   // loads + the cull test that first consumes them
    if (outside) {
        f.triSubtris[tri] = 0;
        prof |= 2;
        return 0;
    }
    // inside the depth range: snap; inside the S16 guard band and small enough -> fast path
    bool done = false;
    if ((v0.w >= fabsf(v0.z)) & (v1.w >= fabsf(v1.z)) & (v2.w >= fabsf(v2.z))) {
        SnappedTri s;
        tm.start();
        snapTriangle(f, v0, v1, v2, s);
        const S32 loxy = min(s.lo.x, s.lo.y), hixy = max(s.hi.x, s.hi.y);
        if (loxy >= -32768 && hixy <= 32767 && hixy - loxy <= aabbLimit) {
            int2 d1, d2;
            S32 area;
            const int res = prepareTriangle<SamplesLog2>(f, s, d1, d2, area);
            tm.stop(f, CRB_TIMER_SetupCullSnap);
            f.triSubtris[tri] = (res == 0) ? 1 : 0;
            prof |= res == 1 ? 4u : res == 2 ? 8u : 32u;
            if (res == 0) {
                // Micro mode: a footprint of at most 4x4 pixel centres (and an extent that keeps the S32 edge functions of
                // microRaster exact) is rasterized right here and never queued; one without any pixel centre inside the
                // surface is dropped.  Same pixel range as triFootprint (Overlap.cuh).
                bool micro = false, nothing = false;
                S32 pxLoX = 0, pxLoY = 0, pxHiX = 0, pxHiY = 0;
                if (SamplesLog2 == 0 && (RenderModeFlags & CRB_FLAG_DEPTH) != 0 && f.microMode) {
                    pxLoX = max((s.lo.x + f.originX + 7) >> CR_SUBPIXEL_LOG2, 0);
                    pxLoY = max((s.lo.y + f.originY + 7) >> CR_SUBPIXEL_LOG2, 0);
                    pxHiX = min((s.hi.x + f.originX - 8) >> CR_SUBPIXEL_LOG2, f.widthPixels - 1);
                    pxHiY = min((s.hi.y + f.originY - 8) >> CR_SUBPIXEL_LOG2, f.heightPixels - 1);
                    nothing = (pxLoX > pxHiX) | (pxLoY > pxHiY);
                    micro = !nothing & (pxHiX - pxLoX < 4) & (pxHiY - pxLoY < 4) & (s.hi.x - s.lo.x < (64 << CR_SUBPIXEL_LOG2)) & (s.hi.y - s.lo.y < (64 << CR_SUBPIXEL_LOG2));
                }
                if (!nothing) {
                    uint3 zp = make_uint3(0, 0, 0);
                    tm.start();
                    uint4 h = setupTriangle<SamplesLog2, RenderModeFlags, true>(f, &f.triHeader[tri], &f.triData[(size_t)tri * 4], vidx, v0, v1, v2, make_float2(0.0f, 0.0f),
                                                                                make_float2(1.0f, 0.0f), make_float2(0.0f, 1.0f), s, d1, d2, area, &zp, micro);
                    tm.stop(f, CRB_TIMER_SetupPleq);
                    tm.start();
                    if (micro) microRaster(f, s.p0.x, s.p0.y, s.p1.x, s.p1.y, s.p2.x, s.p2.y, zp.x, zp.y, zp.z, tri * 8 + 7, pxLoX, pxLoY, pxHiX - pxLoX + 1, pxHiY - pxLoY + 1);
                    else tileCode = histogramBins<SamplesLog2, false>(f, h, tri * 8 + 7, tri, sh);
                    tm.stop(f, CRB_TIMER_SetupBinning);
                }
            }
            done = true;
        }
    }
    if (!done) {
        prof |= 16;
        tm.start();
        if (setupClippedTriangle<SamplesLog2, RenderModeFlags>(f, tri, vidx, v0, v1, v2, sh) > 0) { tileCode = CRB_TILECODE_GENERAL; prof |= 32; }
        tm.stop(f, CRB_TIMER_SetupClip);
    }
    return tileCode;
}

// One thread per input triangle, one CTA per chunk (or slice of a chunk).
template <class VertexClass, int SamplesLog2, U32 RenderModeFlags, int ProfMode = ProfilingMode_Default>
static __global__ void __launch_bounds__(CRB_SETUP_THREADS, CRB_SETUP_MIN_BLOCKS) triangleSetupKernel(const __grid_constant__ crb_frame f) {
    __shared__ SetupCtaShared scratch;
    __shared__ int s_queuedAny[CRB_SETUP_THREADS / 32];   // per WARP = per batch of 32 triangles (crb_frame::batchQueued)
    __shared__ int s_ctaQueued;                            // the first warp of the CTA that queued something counts the CTA
    int* const s_binCount = scratch.binCount;
    const SetupShared sh = {&s_queuedAny[threadIdx.x >> 5], &scratch};
    gridDepLaunchDependents();
    for (int i = threadIdx.x; i < CR_MAXBINS_SQR; i += CRB_SETUP_THREADS) s_binCount[i] = 0;
    if ((threadIdx.x & 31) == 0) s_queuedAny[threadIdx.x >> 5] = 0;
    if (threadIdx.x == 0) s_ctaQueued = 0;
    __syncthreads();
    gridDepWait();   // the previous frame's kernels still read the work buffers written below

    const int stride4 = (int)(sizeof(VertexClass) / sizeof(float4));
    const float4* __restrict__ verts = reinterpret_cast<const float4*>(f.vertexBuffer);

    const int tri = blockIdx.x * CRB_SETUP_THREADS + threadIdx.x;   // one thread per input triangle
    ProfTimer<ProfMode> tmTotal, tm;
    tmTotal.start();
    U32 tileCode = 0;
    U32 prof = 0;   // ProfilingMode_Counters: bit 0 triangle, 1 viewport cull, 2 backface cull, 3 between-pixels cull, 4 clipped, 5 survived
    // Sort-first window with per-chunk bounds (crb_set_chunk_bounds): a chunk whose clip-space box lies outside the window --
    // by more than the rounding of the per-triangle test below, so that every one of its triangles WOULD be culled there --
    // is dropped without reading a vertex.
    bool chunkCulled = false;
    if (f.windowed && f.chunkBounds != nullptr) {
        const float4 b = __ldg(&f.chunkBounds[blockIdx.x]);
        const F32 eps = 1.0e-5f;
        chunkCulled = (b.x > f.cullHiX + eps * (1.0f + fabsf(f.cullHiX))) | (b.z < f.cullLoX - eps * (1.0f + fabsf(f.cullLoX))) |
                      (b.y > f.cullHiY + eps * (1.0f + fabsf(f.cullHiY))) | (b.w < f.cullLoY - eps * (1.0f + fabsf(f.cullLoY)));
    }
    if (tri < f.numTris && chunkCulled) {
        f.triSubtris[tri] = 0;
        prof = 1 | 2;
    } else if (tri < f.numTris) {
        prof = 1;
        tm.start();
        const int3 vidx = make_int3(__ldg(&f.indexBuffer[tri * 3 + 0]), __ldg(&f.indexBuffer[tri * 3 + 1]), __ldg(&f.indexBuffer[tri * 3 + 2]));
        const float4 v0 = __ldg(&verts[(size_t)vidx.x * stride4]);
        const float4 v1 = __ldg(&verts[(size_t)vidx.y * stride4]);
        const float4 v2 = __ldg(&verts[(size_t)vidx.z * stride4]);
        tileCode = setupOneTriangle<SamplesLog2, RenderModeFlags, ProfMode>(f, tri, vidx, v0, v1, v2, sh, prof, tm);
    }

    tmTotal.stop(f, CRB_TIMER_SetupTotal);
    if (ProfMode == ProfilingMode_Counters) {   // reference: TriangleSetup.inl:255-258, :271, :304-305, :329
        __syncwarp();
        profCountWarp<ProfMode>(f, CRB_PROF_SetupViewportCull, (prof & 2) != 0, (prof & 1) != 0);
        profCountWarp<ProfMode>(f, CRB_PROF_SetupBackfaceCull, (prof & 4) != 0, (prof & 1) != 0);
        profCountWarp<ProfMode>(f, CRB_PROF_SetupBetweenPixelsCull, (prof & 8) != 0, (prof & 1) != 0);
        profCountWarp<ProfMode>(f, CRB_PROF_SetupClipped, (prof & 16) != 0, (prof & 1) != 0);
        profCountWarp<ProfMode>(f, CRB_PROF_SetupSamplesPerTri, false, (prof & 32) != 0);   // denominator: triangles that survive setup (the fine raster adds the samples)
    }
    if (f.directMode) {
        // One word per triangle for the scatter pass -- but only where something was queued: a batch of 32 triangles (= this warp)
        // that were all culled or rasterized right here (every batch of a micro-triangle frame) leaves ONE byte instead of 128 B
        // of zeros.  Everything here is warp-local: no CTA barrier at the end of the kernel on this path.
        __syncwarp();
        const int queued = *sh.queuedAny;
        if (queued != 0 && tri < f.numTris) f.triTileCode[tri] = tileCode;
        if ((threadIdx.x & 31) == 0 && tri < f.numTris) {
            f.batchQueued[tri >> 5] = (uint8_t)(queued != 0);
            // one global atomic per CTA, not per warp: 156 000 atomics on ONE address cost C3's setup 45 us
            if (queued != 0 && atomicExch(&s_ctaQueued, 1) == 0) atomicAdd(&f.atomics->numQueuedCtas, 1);
        }
        return;
    }
    // publish this CTA's bin histogram: one column of binCountMat[bin][chunk]
    __syncthreads();
    int* col = f.binCountMat + blockIdx.x / f.ctasPerChunk;
    if (f.ctasPerChunk == 1) {
        for (int b = threadIdx.x; b < f.numBins; b += CRB_SETUP_THREADS) col[(size_t)b * f.matPitch] = s_binCount[b];
    } else {
        for (int b = threadIdx.x; b < f.numBins; b += CRB_SETUP_THREADS)
            if (s_binCount[b] != 0) atomicAdd(&col[(size_t)b * f.matPitch], s_binCount[b]);
    }
}

template <class VertexClass, int SamplesLog2, U32 RenderModeFlags, int ProfMode = ProfilingMode_Default>
inline int launchTriangleSetup(const crb_frame* f, void* stream) {
    if (f->numTris <= 0) return CRB_OK;
    const int grid = (f->numTris + CRB_SETUP_THREADS - 1) / CRB_SETUP_THREADS;
    return launchChained(triangleSetupKernel<VertexClass, SamplesLog2, RenderModeFlags, ProfMode>, grid, CRB_SETUP_THREADS, (cudaStream_t)stream, *f) == cudaSuccess ? CRB_OK : CRB_ERR_CUDA;
}

}  // namespace FW
