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

// The same enumeration shared out over `stride` cooperating threads (thread `first` of them), for the few triangles that span
// many cells, where one thread walking the whole rectangle would be a long tail.  Two forms that visit exactly the cells
// forEachCell visits (a rectangle larger than 2x2 cells is always refined):
//   forEachCellStrided    the cells of the rectangle dealt out one by one, three edge tests per cell -- footprints of a few
//                         hundred cells;
//   forEachCellRowSpans   threads take ROWS of cells; in a row the cells the edge tests accept form one span, because every
//                         test is monotone in the cell's x (the edge function is linear in the sample position and the sample
//                         extent of a cell, clamped to the footprint, moves monotonically with x): the ends of the span are
//                         found by bisection with the very predicate forEachCell applies, ~3 x log2(width) edge evaluations
//                         per row instead of 3 x width -- footprints of thousands of cells (a full-screen triangle has 32 400
//                         tiles at 1080p).
// forEachCellCoop picks by the size of the rectangle.
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

template <int SamplesLog2>
__device__ __forceinline__ bool cellEdgeAccepts(const TriFootprint& t, int i, S32 pxA, S32 pyA, S32 pxB, S32 pyB) {
    const S32 in = SamplesLog2 == 0 ? 8 : 1, out = SamplesLog2 == 0 ? 8 : 15;
    const S64 X0 = (S64)pxA * 16 + in, X1 = (S64)pxB * 16 + out, Y0 = (S64)pyA * 16 + in, Y1 = (S64)pyB * 16 + out;
    const S32 ex = i == 1 ? t.x1 : t.x0, ey = i == 1 ? t.y1 : t.y0;
    const S32 dx = i == 0 ? t.x1 - t.x0 : i == 1 ? t.x2 - t.x1 : t.x0 - t.x2, dy = i == 0 ? t.y1 - t.y0 : i == 1 ? t.y2 - t.y1 : t.y0 - t.y2;
    const S64 px = dy < 0 ? X1 : X0, py = dx > 0 ? Y1 : Y0;
    const S64 e = ((S64)ex - px) * dy - ((S64)ey - py) * dx;
    const S64 tie = (dy > 0 || (dy == 0 && dx <= 0)) ? 1 : 0;
    return e - tie >= 0;   // == !(cellRejected's test of edge i)
}

template <int SamplesLog2, int CellLog2, class SpanFn>
__device__ __forceinline__ void forEachCellRowSpans(const TriFootprint& t, S32 winLoX, S32 winLoY, S32 winHiX, S32 winHiY, int first, int stride, SpanFn spanFn) {
    if (t.empty) return;
    S32 cLoX = t.pxLoX >> CellLog2, cHiX = t.pxHiX >> CellLog2, cLoY = t.pxLoY >> CellLog2, cHiY = t.pxHiY >> CellLog2;
    const bool refine = (cHiX - cLoX > 1) | (cHiY - cLoY > 1);
    cLoX = max(cLoX, winLoX); cHiX = min(cHiX, winHiX); cLoY = max(cLoY, winLoY); cHiY = min(cHiY, winHiY);
    if (cHiX < cLoX || cHiY < cLoY) return;
    const S32 cell = 1 << CellLog2;
    for (S32 cy = cLoY + first; cy <= cHiY; cy += stride) {
        S32 xa = cLoX, xb = cHiX;   // the span of accepted cells of this row
        if (refine) {
            const S32 pyA = max(cy << CellLog2, t.pxLoY), pyB = min((cy << CellLog2) + cell - 1, t.pxHiY);
            auto accepts = [&](int i, S32 cx) { return cellEdgeAccepts<SamplesLog2>(t, i, max(cx << CellLog2, t.pxLoX), pyA, min((cx << CellLog2) + cell - 1, t.pxHiX), pyB); };
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const S32 dy = i == 0 ? t.y1 - t.y0 : i == 1 ? t.y2 - t.y1 : t.y0 - t.y2;
                if (xa > xb) break;
                if (dy < 0) {          // the edge function grows with x: accepted cells are the right part of the row
                    if (!accepts(i, xb)) { xa = xb + 1; break; }
                    S32 lo = xa, hi = xb;   // smallest accepted x in [lo, hi]; hi is accepted
                    while (lo < hi) { const S32 mid = (lo + hi) >> 1; if (accepts(i, mid)) hi = mid; else lo = mid + 1; }
                    xa = lo;
                } else if (dy > 0) {   // it falls with x: the left part
                    if (!accepts(i, xa)) { xa = xb + 1; break; }
                    S32 lo = xa, hi = xb;   // largest accepted x in [lo, hi]; lo is accepted
                    while (lo < hi) { const S32 mid = (lo + hi + 1) >> 1; if (accepts(i, mid)) lo = mid; else hi = mid - 1; }
                    xb = lo;
                } else if (!accepts(i, xa)) { xa = xb + 1; break; }   // constant along the row
            }
        }
        if (xa <= xb) spanFn(cy, xa, xb);   // cells (xa..xb, cy)
    }
}

// fn(cx, cy) takes one cell, spanFn(cy, xa, xb) the cells xa..xb of row cy (so that a caller whose per-cell work has a long
// latency -- an atomic whose result it needs -- can keep several cells of a span in flight).
template <int SamplesLog2, int CellLog2, class Fn, class SpanFn>
__device__ __forceinline__ void forEachCellCoop(const TriFootprint& t, S32 winLoX, S32 winLoY, S32 winHiX, S32 winHiY, int first, int stride, Fn fn, SpanFn spanFn) {
    const S32 nx = (t.pxHiX >> CellLog2) - (t.pxLoX >> CellLog2) + 1, ny = (t.pxHiY >> CellLog2) - (t.pxLoY >> CellLog2) + 1;
    if (nx >= 32 && ny * 4 >= stride) forEachCellRowSpans<SamplesLog2, CellLog2>(t, winLoX, winLoY, winHiX, winHiY, first, stride, spanFn);   // enough rows to keep a good part of the threads busy
    else forEachCellStrided<SamplesLog2, CellLog2>(t, winLoX, winLoY, winHiX, winHiY, first, stride, fn);
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
