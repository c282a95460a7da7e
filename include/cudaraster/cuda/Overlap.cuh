// Conservative triangle -> cell (bin or tile) overlap enumeration shared by the setup, bin and
// coarse stages.  New design (the reference walks edge functions in BinRaster.inl:263-299 and
// CoarseRaster.inl:297-435); contract kept from SURVEY.md A.8: the produced set is a SUPERSET of
// the cells that contain a covered sample, and every stage that enumerates the cells of the same
// triangle gets exactly the same set (count and scatter passes must agree).
//
// A cell is 2^CellLog2 x 2^CellLog2 pixels (7 = bin, 3 = tile).  The candidate range is the
// bounding box of the SAMPLE POSITIONS inside the triangle's AABB (pixel centres for 1 sample),
// which already drops most cells a tiny triangle merely touches.  Footprints larger than 2x2
// cells are refined with three exact S64 edge tests against the sample extent of each cell.
#pragma once
#include "Util.cuh"

namespace FW {

struct TriFootprint {
    S32 x0, y0, x1, y1, x2, y2;      // vertices, surface-corner subpixels
    S32 pxLoX, pxLoY, pxHiX, pxHiY;  // inclusive pixel range that can contain covered samples
    bool empty;
};

// header words: (v0y<<16 | v0x&0xffff), v1, v2 as stored by triangle setup.
template <int SamplesLog2>
__device__ __forceinline__ TriFootprint triFootprint(U32 hx, U32 hy, U32 hz, const crb_frame& f) {
    TriFootprint t;
    const S32 ox = f.originX, oy = f.originY;
    t.x0 = (S32)(S16)(hx & 0xFFFF) + ox; t.y0 = ((S32)hx >> 16) + oy;
    t.x1 = (S32)(S16)(hy & 0xFFFF) + ox; t.y1 = ((S32)hy >> 16) + oy;
    t.x2 = (S32)(S16)(hz & 0xFFFF) + ox; t.y2 = ((S32)hz >> 16) + oy;
    S32 lox = min(min(t.x0, t.x1), t.x2), hix = max(max(t.x0, t.x1), t.x2);
    S32 loy = min(min(t.y0, t.y1), t.y2), hiy = max(max(t.y0, t.y1), t.y2);
    const S32 upLo = SamplesLog2 == 0 ? 7 : 0, dnHi = SamplesLog2 == 0 ? 8 : 0;
    t.pxLoX = max((lox + upLo) >> CR_SUBPIXEL_LOG2, 0);
    t.pxLoY = max((loy + upLo) >> CR_SUBPIXEL_LOG2, 0);
    t.pxHiX = min((hix - dnHi) >> CR_SUBPIXEL_LOG2, f.widthPixels - 1);
    t.pxHiY = min((hiy - dnHi) >> CR_SUBPIXEL_LOG2, f.heightPixels - 1);
    t.empty = (t.pxLoX > t.pxHiX) | (t.pxLoY > t.pxHiY);
    return t;
}

// True when no sample inside pixel range [pxA..pxB] x [pyA..pyB] can be covered by the triangle.
template <int SamplesLog2>
__device__ __forceinline__ bool cellRejected(const TriFootprint& t, S32 pxA, S32 pyA, S32 pxB, S32 pyB) {
    const S32 in = SamplesLog2 == 0 ? 8 : 1, out = SamplesLog2 == 0 ? 8 : 15;
    const S64 X0 = (S64)pxA * 16 + in, X1 = (S64)pxB * 16 + out, Y0 = (S64)pyA * 16 + in, Y1 = (S64)pyB * 16 + out;
    const S32 ex[3] = {t.x0, t.x1, t.x0}, ey[3] = {t.y0, t.y1, t.y0};
    const S32 dx[3] = {t.x1 - t.x0, t.x2 - t.x1, t.x0 - t.x2}, dy[3] = {t.y1 - t.y0, t.y2 - t.y1, t.y0 - t.y2};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        // E(p) = (o.x - p.x)*d.y - (o.y - p.y)*d.x, maximised over the cell's sample extent
        S64 px = dy[i] < 0 ? X1 : X0, py = dx[i] > 0 ? Y1 : Y0;
        S64 e = ((S64)ex[i] - px) * dy[i] - ((S64)ey[i] - py) * dx[i];
        S64 tie = (dy[i] > 0 || (dy[i] == 0 && dx[i] <= 0)) ? 1 : 0;
        if (e - tie < 0) return true;
    }
    return false;
}

// Calls fn(cellX, cellY) for every cell of size 2^CellLog2 px overlapped by the triangle, in
// row-major order, restricted to the inclusive cell window [winLoX..winHiX] x [winLoY..winHiY].
template <int SamplesLog2, int CellLog2, class Fn>
__device__ __forceinline__ void forEachCell(const TriFootprint& t, S32 winLoX, S32 winLoY, S32 winHiX, S32 winHiY, Fn fn) {
    if (t.empty) return;
    S32 cLoX = t.pxLoX >> CellLog2, cHiX = t.pxHiX >> CellLog2, cLoY = t.pxLoY >> CellLog2, cHiY = t.pxHiY >> CellLog2;
    const bool refine = (cHiX - cLoX > 1) | (cHiY - cLoY > 1);   // decided on the UNclipped footprint
    cLoX = max(cLoX, winLoX); cHiX = min(cHiX, winHiX); cLoY = max(cLoY, winLoY); cHiY = min(cHiY, winHiY);
    for (S32 cy = cLoY; cy <= cHiY; cy++)
        for (S32 cx = cLoX; cx <= cHiX; cx++) {
            if (refine) {
                S32 pxA = max(cx << CellLog2, t.pxLoX), pxB = min((cx << CellLog2) + (1 << CellLog2) - 1, t.pxHiX);
                S32 pyA = max(cy << CellLog2, t.pxLoY), pyB = min((cy << CellLog2) + (1 << CellLog2) - 1, t.pxHiY);
                if (cellRejected<SamplesLog2>(t, pxA, pyA, pxB, pyB)) continue;
            }
            fn(cx, cy);
        }
}

// The same enumeration shared out over `stride` cooperating threads (thread `first` of them): used for the few
// triangles that span many cells, where one thread walking the whole rectangle would be a long tail.  Visits
// exactly the cells forEachCell visits (a rectangle larger than 2x2 cells is always refined).
template <int SamplesLog2, int CellLog2, class Fn>
__device__ __forceinline__ void forEachCellStrided(const TriFootprint& t, S32 winLoX, S32 winLoY, S32 winHiX, S32 winHiY, int first, int stride, Fn fn) {
    if (t.empty) return;
    S32 cLoX = t.pxLoX >> CellLog2, cHiX = t.pxHiX >> CellLog2, cLoY = t.pxLoY >> CellLog2, cHiY = t.pxHiY >> CellLog2;
    const bool refine = (cHiX - cLoX > 1) | (cHiY - cLoY > 1);
    cLoX = max(cLoX, winLoX); cHiX = min(cHiX, winHiX); cLoY = max(cLoY, winLoY); cHiY = min(cHiY, winHiY);
    const S32 nx = cHiX - cLoX + 1, ny = cHiY - cLoY + 1;
    if (nx <= 0 || ny <= 0) return;
    for (S32 k = first; k < nx * ny; k += stride) {
        const S32 cy = cLoY + k / nx, cx = cLoX + (k - (k / nx) * nx);
        if (refine) {
            S32 pxA = max(cx << CellLog2, t.pxLoX), pxB = min((cx << CellLog2) + (1 << CellLog2) - 1, t.pxHiX);
            S32 pyA = max(cy << CellLog2, t.pxLoY), pyB = min((cy << CellLog2) + (1 << CellLog2) - 1, t.pxHiY);
            if (cellRejected<SamplesLog2>(t, pxA, pyA, pxB, pyB)) continue;
        }
        fn(cx, cy);
    }
}

// Cell rectangle of a footprint inside an inclusive cell window: first cell, extent (0 = nothing)
// and whether cells must be refined with edge tests (decided on the UNclipped footprint, exactly
// like forEachCell, so that every pass that enumerates the cells of a triangle agrees).
struct CellRange { S32 x0, y0, nx, ny; bool refine; };
template <int CellLog2>
__device__ __forceinline__ CellRange cellRange(const TriFootprint& t, S32 winLoX, S32 winLoY, S32 winHiX, S32 winHiY) {
    CellRange r;
    S32 cLoX = t.pxLoX >> CellLog2, cHiX = t.pxHiX >> CellLog2, cLoY = t.pxLoY >> CellLog2, cHiY = t.pxHiY >> CellLog2;
    r.refine = (cHiX - cLoX > 1) | (cHiY - cLoY > 1);
    cLoX = max(cLoX, winLoX); cHiX = min(cHiX, winHiX); cLoY = max(cLoY, winLoY); cHiY = min(cHiY, winHiY);
    r.x0 = cLoX; r.y0 = cLoY;
    r.nx = t.empty ? 0 : max(cHiX - cLoX + 1, 0);
    r.ny = t.empty ? 0 : max(cHiY - cLoY + 1, 0);
    if (r.nx == 0 || r.ny == 0) r.nx = r.ny = 0;
    return r;
}

// Resolves a queue entry (triIdx*8 + sub, sub == 7 meaning "the only sub-triangle, stored at
// triIdx") to the slot of its header/data (reference: BinRaster.inl:190-197).
__device__ __forceinline__ S32 resolveDataIdx(S32 entry, const uint4* __restrict__ triHeader) {
    S32 tri = entry >> 3, sub = entry & 7;
    if (sub == 7) return tri;
    return (S32)__ldg(&triHeader[tri]).w + sub;
}

}  // namespace FW
