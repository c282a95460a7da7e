// Stage 4 -- fine raster for sm_100a: one warp per 8x8 px tile, every lane OWNS two pixels.
//
// Computes what the reference's fineRasterImpl_SingleSample / _MultiSample compute
// (src/cudaraster/cuda/FineRaster.inl:499-752, :855-1126): exact top-left coverage at the
// sample positions, U32 plane-equation depth with a strict LESS test, one fragment-shader run per
// (triangle, pixel) at the pixel centre / MSAA centroid, blend, all in submission order.
//
// What is different (DESIGN.md "fine raster"):
//  * Pixel ownership.  The reference maps 32 *fragments* to lanes and resolves same-pixel
//    conflicts with racing shared-memory stores whose arbitration was deterministic on Fermi
//    only.  Here lane l owns pixels l and l+32 of the tile for the whole tile:
//    fragments of one pixel are applied by one thread in queue order, so the serial rule holds
//    by construction and the tile colour/depth live in registers, not shared memory.
//  * Refill: 32 queue entries at a time, one lane per triangle, computes the three edge
//    equations relative to the tile (exact integer arithmetic, no FP32 LUT) and the tile-relative
//    depth plane into a 64 B shared-memory record; __ballot_sync picks the live ones.
//  * Visibility-first shading: when the blend does not read dst and the shader cannot discard,
//    the loop only tracks (depth, queue position) of the winning fragment per sample and the
//    shader runs once per covered pixel at the end -- overdraw costs no shading.
//  * Explicit __syncwarp / __ballot_sync / __shfl_sync everywhere; no volatile-shared tricks.
#pragma once
#include "Overlap.cuh"
#include "PixelPipe.hpp"

namespace FW {

// The fine raster kernel is the last kernel of a frame: it leaves the per-frame scratch state the way
// the next frame expects it, so that a steady-state frame enqueues kernels only (no memsets):
//  * the counter block of the NEXT frame (double-buffered) is zeroed;
//  * the rows of the tile count matrix this frame used (one per coarse work item) are zeroed again --
//    plain coalesced 16-byte stores spread over the whole grid.  (Zeroing each row in the coarse scatter
//    kernel right after it is read was measured 8 us slower on C2: 31.2 vs 23.5 us for the coarse stage.)
// Must run after gridDepWait() and before any early return.
__device__ __forceinline__ void finishFrameState(const crb_frame& f) {
    // An overflowed frame leaves the self-cleaning scratch state dirty (its kernels return early), so the frames enqueued behind
    // it must not run on it: the flag is handed on (bit 4 = "an earlier frame of the batch overflowed") and every later frame
    // early-outs until the host has seen it (crb_finish) and reset the state.
    // the frame's counters are final by now (every other kernel of the frame has ended): the host's copy is stored from here --
    // a copy enqueued between two frames would hold the next frame's first kernel back by a copy-engine round trip
    if (blockIdx.x == 0 && threadIdx.x < (int)(sizeof(crb_atomics) / sizeof(int))) {
        reinterpret_cast<volatile int*>(f.hostCounters)[threadIdx.x] = reinterpret_cast<const int*>(f.atomics)[threadIdx.x];
        const int sticky = (threadIdx.x == (int)(offsetof(crb_atomics, overflow) / sizeof(int)) && f.atomics->overflow != 0) ? 16 : 0;
        reinterpret_cast<int*>(f.nextAtomics)[threadIdx.x] = sticky;
    }
    const int rows = min(f.atomics->numCoarseItems, f.maxItems);
    const int total = rows * (CR_BIN_SQR / 4);
    int4* mat = reinterpret_cast<int4*>(f.tileCountMat);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) mat[i] = make_int4(0, 0, 0, 0);
}

struct FineTriRec {  // 64 B, one per queued triangle of the current batch
    S32 a0, b0, c0, a1;   // edge i: E(sx, sy) = c_i + a_i*sx + b_i*sy >= 0, (sx, sy) = subpixel
    S32 b1, c1, a2, b2;   //         offset of the sample from the centre of the tile's pixel (0,0)
    S32 c2;
    U32 zx, zy, zb;       // depth = zb + zx*dsx + zy*dsy, (dsx, dsy) = sample offset in sample units
    S32 dataIdx, triIdx, seq, pad;
};

// Edge equations of a sub-triangle relative to the sample-space origin (bx, by) given in
// viewport-centred subpixels.  Exact: the constant is formed in S64 and clamped to +-2^30, which
// cannot change the sign of E anywhere inside a tile (|a*sx + b*sy| < 2^24 there).
__device__ __forceinline__ void setupTileEdges(const uint4& h, S32 bx, S32 by, S32 (&a)[3], S32 (&b)[3], S32 (&c)[3]) {
    const S32 x0 = (S32)(S16)(h.x & 0xFFFF), y0 = (S32)h.x >> 16;
    const S32 x1 = (S32)(S16)(h.y & 0xFFFF), y1 = (S32)h.y >> 16;
    const S32 x2 = (S32)(S16)(h.z & 0xFFFF), y2 = (S32)h.z >> 16;
    const S32 ox[3] = {x0, x1, x0}, oy[3] = {y0, y1, y0};
    const S32 dx[3] = {x1 - x0, x2 - x1, x0 - x2}, dy[3] = {y1 - y0, y2 - y1, y0 - y2};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        // E(s) = (o.x - s.x)*d.y - (o.y - s.y)*d.x - tie,  s = (bx + sx, by + sy)
        const S64 tie = (dy[i] > 0 || (dy[i] == 0 && dx[i] <= 0)) ? 1 : 0;
        S64 e = (S64)(ox[i] - bx) * dy[i] - (S64)(oy[i] - by) * dx[i] - tie;
        e = max(min(e, (S64)(1 << 30)), -(S64)(1 << 30));
        c[i] = (S32)e;
        a[i] = -dy[i];
        b[i] = dx[i];
    }
}

// Lanes that own the pixels of this lane's 2x2 quad: pixel (x, y & 3) of a 4-row block sits on lane x + 8*(y & 3).
__device__ __forceinline__ U32 quadLaneMask() { return 0x303u << (laneId() & 0x16u); }

// Perspective-correct barycentrics from the integer w/u/v planes (reference: FineRaster.inl:21-48).
template <int SamplesLog2>
__device__ __forceinline__ void computeBarys(Vec3f& bary, Vec3f& baryDX, Vec3f& baryDY, const int3& wp, const int3& up, const int3& vp, int sampleX, int sampleY) {
    const F32 w = __frcp_rn((F32)(wp.x * sampleX + wp.y * sampleY + wp.z));
    const F32 u = __fmul_rn(w, (F32)(up.x * sampleX + up.y * sampleY + up.z));
    const F32 v = __fmul_rn(w, (F32)(vp.x * sampleX + vp.y * sampleY + vp.z));
    bary = Vec3f(__fsub_rn(__fsub_rn(1.0f, u), v), u, v);
    const F32 wd = w * (F32)(1 << (SamplesLog2 + 1));
    const F32 udx = wd * ((F32)up.x - u * (F32)wp.x), udy = wd * ((F32)up.y - u * (F32)wp.y);
    const F32 vdx = wd * ((F32)vp.x - v * (F32)wp.x), vdy = wd * ((F32)vp.y - v * (F32)wp.y);
    baryDX = Vec3f(-udx - vdx, udx, vdx);
    baryDY = Vec3f(-udy - vdy, udy, vdy);
}

// Fills the shader inputs and runs it (reference: FineRaster.inl:52-119).  centroid = packed
// half-sample position of the shading point inside the pixel (low nibble x, high nibble y).
template <class VertexClass, class FragmentShaderClass, int SamplesLog2, U32 RenderModeFlags>
__device__ __forceinline__ void runFragmentShader(FragmentShaderClass& fs, const crb_frame& f, int triIdx, int dataIdx, int pixelX, int pixelY, U32 centroid) {
#if CRB_WIDE_LD
    uint4 t2, t3;   // rows 2 and 3 of the record in ONE 256-bit gather: uy, ub, vx, vy | vb, vi0, vi1, vi2
    ldg256(&f.triData[(size_t)dataIdx * 4 + 2], t2, t3);
#else
    const uint4 t3 = __ldg(&f.triData[(size_t)dataIdx * 4 + 3]);  // vb, vi0, vi1, vi2
#endif
    // quads mode: the caller runs this converged on the four lanes of the pixel's 2x2 quad
    fs.m_quadMask = (RenderModeFlags & RenderModeFlag_EnableQuads) != 0 ? quadLaneMask() : (1u << laneId());
    fs.m_triIdx = triIdx;
    fs.m_vertIdx = Vec3i((S32)t3.y, (S32)t3.z, (S32)t3.w);
    fs.m_pixelPos = Vec2i(pixelX, pixelY);
    fs.m_vertexBytes = (S32)sizeof(VertexClass);
    fs.m_vertexBuffer = f.vertexBuffer;
    fs.m_color = 0xFF0000FFu;
    fs.m_discard = false;
    if ((RenderModeFlags & RenderModeFlag_EnableLerp) == 0) {
        // interpolation off: varyings come from the last vertex
        fs.m_center = Vec3f(0.0f, 0.0f, 1.0f);
        fs.m_centerDX = Vec3f(0.0f); fs.m_centerDY = Vec3f(0.0f);
        fs.m_centroid = Vec3f(0.0f, 0.0f, 1.0f);
        fs.m_centroidDX = Vec3f(0.0f); fs.m_centroidDY = Vec3f(0.0f);
    } else {
        const uint4 t1 = ldg128Record(&f.triData[(size_t)dataIdx * 4 + 1]);  // wx, wy, wb, ux
#if !CRB_WIDE_LD
        const uint4 t2 = __ldg(&f.triData[(size_t)dataIdx * 4 + 2]);  // uy, ub, vx, vy
#endif
        const int3 wp = make_int3((S32)t1.x, (S32)t1.y, (S32)t1.z);
        const int3 up = make_int3((S32)t1.w, (S32)t2.x, (S32)t2.y);
        const int3 vp = make_int3((S32)t2.z, (S32)t2.w, (S32)t3.x);
        computeBarys<SamplesLog2>(fs.m_center, fs.m_centerDX, fs.m_centerDY, wp, up, vp, (pixelX * 2 + 1) << SamplesLog2, (pixelY * 2 + 1) << SamplesLog2);
        if (SamplesLog2 == 0) {
            fs.m_centroid = fs.m_center; fs.m_centroidDX = fs.m_centerDX; fs.m_centroidDY = fs.m_centerDY;
        } else {
            computeBarys<SamplesLog2>(fs.m_centroid, fs.m_centroidDX, fs.m_centroidDY, wp, up, vp, (pixelX << (SamplesLog2 + 1)) + (S32)(centroid & 0xF),
                                      (pixelY << (SamplesLog2 + 1)) + (S32)(centroid >> 4));
        }
    }
    fs.run();
}

template <class BlendShaderClass>
__device__ __forceinline__ void runBlendShader(BlendShaderClass& bs, int triIdx, int pixelX, int pixelY, int sampleIdx, U32 src, U32 dst) {
    bs.m_triIdx = triIdx;
    bs.m_pixelPos = Vec2i(pixelX, pixelY);
    bs.m_sampleIdx = sampleIdx;
    bs.m_src = src;
    bs.m_dst = dst;
    bs.m_color = 0xFF0000FFu;
    bs.m_writeColor = true;
    bs.run();
}

// Packed half-sample position of the shading point for a sample coverage mask
// (reference: FineRaster.inl:151-164): pixel centre when all or none are covered.
template <int SamplesLog2>
__device__ __forceinline__ U32 centroidCode(U32 sampleMask) {
    const int y = selectMSAACentroid(SamplesLog2, sampleMask);
    if (y < 0) return 0x11u << SamplesLog2;
    return (U32)(msaaSampleX(SamplesLog2, y) * 0x02 + y * 0x20 + 0x11);
}

// Lane-per-triangle refill of one batch of <= 32 queue entries.  Returns the ballot of triangles
// that can touch the tile.
template <int SamplesLog2, U32 RenderModeFlags>
__device__ __forceinline__ U32 fineRefill(const crb_frame& f, FineTriRec* recs, int queuePos, int remaining, int tileX, int tileY) {
    const int lane = laneId();
    bool live = false;
    if (lane < remaining) {
        const S32 entry = __ldg(&f.tileQueue[queuePos + lane]);
        const S32 dataIdx = resolveDataIdx(entry, f.triHeader);
        const uint4 h = __ldg(&f.triHeader[dataIdx]);
        // sample-space origin: centre of pixel (0,0) of the tile, viewport-centred subpixels
        const S32 bx = (tileX << (CR_TILE_LOG2 + CR_SUBPIXEL_LOG2)) + (CR_SUBPIXEL_SIZE >> 1) - f.originX;
        const S32 by = (tileY << (CR_TILE_LOG2 + CR_SUBPIXEL_LOG2)) + (CR_SUBPIXEL_SIZE >> 1) - f.originY;
        S32 a[3], b[3], c[3];
        setupTileEdges(h, bx, by, a, b, c);
        // sample offsets inside the tile span [-8, 120] subpixels on both axes
        const S32 lo = SamplesLog2 == 0 ? 0 : -8, hi = SamplesLog2 == 0 ? 112 : 120;
        live = true;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const S32 emax = c[i] + max(a[i] * lo, a[i] * hi) + max(b[i] * lo, b[i] * hi);
            live &= emax >= 0;
        }
        if (live) {
            FineTriRec r;
            r.a0 = a[0]; r.b0 = b[0]; r.c0 = c[0];
            r.a1 = a[1]; r.b1 = b[1]; r.c1 = c[1];
            r.a2 = a[2]; r.b2 = b[2]; r.c2 = c[2];
            r.zx = r.zy = r.zb = 0;
            if ((RenderModeFlags & RenderModeFlag_EnableDepth) != 0) {
                const uint4 z = __ldg(&f.triData[(size_t)dataIdx * 4]);
                r.zx = z.x; r.zy = z.y;
                r.zb = z.z + z.x * (U32)(tileX << (CR_TILE_LOG2 + SamplesLog2)) + z.y * (U32)(tileY << (CR_TILE_LOG2 + SamplesLog2));
            }
            r.dataIdx = dataIdx; r.triIdx = entry >> 3; r.seq = 0; r.pad = 0;
            uint4* dst = reinterpret_cast<uint4*>(&recs[lane]);
            dst[0] = make_uint4((U32)r.a0, (U32)r.b0, (U32)r.c0, (U32)r.a1);
            dst[1] = make_uint4((U32)r.b1, (U32)r.c1, (U32)r.a2, (U32)r.b2);
            dst[2] = make_uint4((U32)r.c2, r.zx, r.zy, r.zb);
            dst[3] = make_uint4((U32)r.dataIdx, (U32)r.triIdx, 0u, 0u);
        }
    }
    const U32 mask = __ballot_sync(0xFFFFFFFFu, live);
    __syncwarp();
    return mask;
}

//------------------------------------------------------------------------------------------------
// Coverage helpers shared by the single- and multi-sample kernels.
//------------------------------------------------------------------------------------------------

// 32x32 bit-matrix transpose across the warp: lane j passes row j (bit i = element (j, i)),
// lane i receives column i (bit j = element (j, i)).  Five butterfly stages of one shuffle each.
__device__ __forceinline__ U32 warpTranspose32(U32 x, int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const U32 lo = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
        const U32 o = __shfl_xor_sync(0xFFFFFFFFu, x, s);
        x = (lane & s) ? ((x & ~lo) | ((o >> s) & lo)) : ((x & lo) | ((o << s) & ~lo));
    }
    return x;
}

// 64-bit coverage of one 8x8 tile for edge equations given relative to the centre of pixel
// (0,0) of the tile, E_i(sx, sy) = c_i + a_i*sx + b_i*sy with (sx, sy) = 16 * (x, y).  Bit x + 8y.
// Exact (integer arithmetic, the tie rule is folded into c_i).  Only rows [rowLo, rowHi] are
// visited, so the cost follows the triangle's height inside the tile; SampleOfs shifts the
// sample point (used by the MSAA kernel to build its conservative pixel mask).
__device__ __forceinline__ void coverTileRows(const S32 (&a)[3], const S32 (&b)[3], const S32 (&c)[3], int rowLo, int rowHi, U32& maskLo, U32& maskHi) {
    maskLo = 0; maskHi = 0;
    const S32 a0 = a[0] << CR_SUBPIXEL_LOG2, a1 = a[1] << CR_SUBPIXEL_LOG2, a2 = a[2] << CR_SUBPIXEL_LOG2;
#pragma unroll 1
    for (int r = rowLo; r <= rowHi; r++) {
        // start at column 7 and walk left so that the funnel shift leaves column 0 in bit 0
        S32 e0 = c[0] + b[0] * (r << CR_SUBPIXEL_LOG2) + a0 * 7;
        S32 e1 = c[1] + b[1] * (r << CR_SUBPIXEL_LOG2) + a1 * 7;
        S32 e2 = c[2] + b[2] * (r << CR_SUBPIXEL_LOG2) + a2 * 7;
        U32 row = 0;
#pragma unroll
        for (int x = 0; x < 8; x++) {
            row = __funnelshift_l(~(U32)(e0 | e1 | e2), row, 1);   // row = row << 1 | inside
            e0 -= a0; e1 -= a1; e2 -= a2;
        }
        if (r < 4) maskLo |= row << (8 * r);
        else maskHi |= row << (8 * (r - 4));
    }
}

// Coverage of a SMALL triangle: its pixel-centre bounding box inside the tile is at most 4x4 pixels
// (corner pixel (colLo, rowLo), nc x nr pixels) and its extent is below 64 px, so every edge function
// fits S32 and the 16 candidate pixels are evaluated in straight-line code -- no per-lane loop whose
// trip count the warp would have to take the maximum of.  Same rule as coverTileRows, bit for bit.
// (x*, y*) = vertices in viewport-centred subpixels, (bx, by) = centre of the tile's pixel (0,0).
__device__ __forceinline__ void coverSmall4x4(S32 x0, S32 y0, S32 x1, S32 y1, S32 x2, S32 y2, S32 bx, S32 by, int colLo, int rowLo, int nc, int nr, U32& maskLo, U32& maskHi) {
    const S32 px = bx + (colLo << CR_SUBPIXEL_LOG2), py = by + (rowLo << CR_SUBPIXEL_LOG2);
    const S32 dx0 = x1 - x0, dy0 = y1 - y0, dx1 = x2 - x1, dy1 = y2 - y1, dx2 = x0 - x2, dy2 = y0 - y2;
    // E_i at the corner pixel; stepping one pixel right adds -dy*16, one pixel up adds dx*16
    S32 e0 = (x0 - px) * dy0 - (y0 - py) * dx0 - ((dy0 > 0 || (dy0 == 0 && dx0 <= 0)) ? 1 : 0);
    S32 e1 = (x1 - px) * dy1 - (y1 - py) * dx1 - ((dy1 > 0 || (dy1 == 0 && dx1 <= 0)) ? 1 : 0);
    S32 e2 = (x0 - px) * dy2 - (y0 - py) * dx2 - ((dy2 > 0 || (dy2 == 0 && dx2 <= 0)) ? 1 : 0);
    const S32 a0 = -(dy0 << CR_SUBPIXEL_LOG2), a1 = -(dy1 << CR_SUBPIXEL_LOG2), a2 = -(dy2 << CR_SUBPIXEL_LOG2);
    const S32 b0 = dx0 << CR_SUBPIXEL_LOG2, b1 = dx1 << CR_SUBPIXEL_LOG2, b2 = dx2 << CR_SUBPIXEL_LOG2;
    U32 acc = 0;   // byte r = row rowLo + r, bit x = column colLo + x
#pragma unroll
    for (int r = 0; r < 4; r++) {
        // walk from column 3 down to 0 so that the funnel shift leaves column 0 in bit 0
        S32 t0 = e0 + 3 * a0, t1 = e1 + 3 * a1, t2 = e2 + 3 * a2;
        U32 row = 0;
#pragma unroll
        for (int x = 0; x < 4; x++) {
            row = __funnelshift_l(~(U32)(t0 | t1 | t2), row, 1);
            t0 -= a0; t1 -= a1; t2 -= a2;
        }
        acc |= row << (8 * r);
        e0 += b0; e1 += b1; e2 += b2;
    }
    // pixels beyond the clipped box lie outside the tile (or outside the triangle): drop them
    const U32 colMask = (1u << nc) - 1u;
    acc &= colMask * 0x01010101u;
    acc &= nr >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nr)) - 1u);
    const unsigned long long m = (unsigned long long)acc << (8 * rowLo + colLo);
    maskLo = (U32)m;
    maskHi = (U32)(m >> 32);
}

// One queued sub-triangle as the refill stage fetched it.
struct FineFetch {
    S32 entry;     // triIdx*8 + sub, < 0 = none
    S32 dataIdx;
    uint4 h;       // CRTriangleHeader
    uint4 z;       // first row of CRTriangleData (zx, zy, zb, zslope)
};

template <U32 RenderModeFlags>
__device__ __forceinline__ void fineFetch(FineFetch& t, const crb_frame& f, S32 entry) {
    t.entry = entry;
    if (entry < 0) return;
    t.dataIdx = resolveDataIdx(entry, f.triHeader);
    t.h = __ldg(&f.triHeader[t.dataIdx]);
    if ((RenderModeFlags & RenderModeFlag_EnableDepth) != 0) t.z = __ldg(&f.triData[(size_t)t.dataIdx * 4]);
}

// Tile-level early Z (reference: FineRaster.inl:229-241: "zmin >= tileZMax -> skip the triangle").  The reference's cull is not a
// pure optimisation: the header's zmin bounds the plane depth from below only where the fixed-point plane is well conditioned --
// a sub-triangle left by clipping at w ~ 0 can span the guard band with a depth plane whose values at covered pixels lie BELOW
// its own zmin, and whether the reference culls it depends on how its fragments happened to be batched.  Here a triangle is
// dropped only when that is provably invisible in the result: the header test must say so AND the plane itself, evaluated
// without wrap-around over the box of samples the triangle can cover in this tile ([x0, x1] x [y0, y1], sample units relative
// to the tile's first sample), must stay at or behind the tile's farthest depth.  Every fragment of a culled triangle would then
// fail the LESS test (direct path: could not even tie), so the frame equals the one rendered without early Z -- the oracle's.
__device__ __forceinline__ bool earlyZCull(U32 zminHdr, U32 tileZMax, bool direct, U32 zbTile, U32 zx, U32 zy, int x0, int x1, int y0, int y1) {
    if (zminHdr < tileZMax || (direct && zminHdr == tileZMax)) return false;
    const S64 sx = (S64)(S32)zx, sy = (S64)(S32)zy;   // any representative of the slope mod 2^32 gives the same depths at integer samples
    const S64 lo = (S64)zbTile + min(sx * x0, sx * x1) + min(sy * y0, sy * y1);
    const S64 hi = (S64)zbTile + max(sx * x0, sx * x1) + max(sy * y0, sy * y1);
    if (lo < 0 || hi > (S64)0xFFFFFFFFll) return false;   // the U32 plane wraps inside the box: its minimum is not at a corner
    return direct ? lo > (S64)tileZMax : lo >= (S64)tileZMax;
}

// Per-warp staging of the current batch, lane-indexed SoA: the per-fragment gathers of the
// ownership loop hit distinct banks for distinct triangles and broadcast for equal ones.
struct FineBatch {
    U32 zx[32], zy[32], zb[32];
    S32 entry[32];
    S32 dataIdx[32];
};

//------------------------------------------------------------------------------------------------
// Single-sample kernel.
//
// Per batch of <= 32 queue entries:  (1) lane j builds the exact 64-bit coverage mask of triangle j
// (early-Z against the tile's max depth first);  (2) two warp transposes turn the 32 masks into,
// for every lane, the set of triangles covering each of its two pixels;  (3) every lane walks its
// own pixels' sets in queue order: depth test, winner update (or shade + blend in place).
// Work is proportional to fragments, not to triangles x 64 pixels, and fragments of one pixel
// are still applied by one thread in submission order.
//------------------------------------------------------------------------------------------------

template <class VertexClass, class FragmentShaderClass, class BlendShaderClass, U32 RenderModeFlags, int ProfMode = ProfilingMode_Default>
static __global__ void __launch_bounds__(CRB_FINE_WARPS * 32, CRB_FINE_WARPS_PER_SM / CRB_FINE_WARPS) fineRasterSingleKernel(const __grid_constant__ crb_frame f) {
    __shared__ FineBatch s_batch[CRB_FINE_WARPS];
    __shared__ __align__(128) U32 s_tileStage[CRB_FINE_WARPS][CR_TILE_SQR];   // tile-major colour surfaces: staging of the tile for the bulk store

    constexpr bool kDepth = (RenderModeFlags & RenderModeFlag_EnableDepth) != 0;
    constexpr bool kQuads = (RenderModeFlags & RenderModeFlag_EnableQuads) != 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int activeIdx = blockIdx.x * CRB_FINE_WARPS + warp;
    gridDepLaunchDependents();
    gridDepWait();
    finishFrameState(f);
    // {tile, queue start, queue count} in ONE load, issued together with the counters (the slot is
    // always inside the buffer; it only holds a real record when activeIdx < numActiveTiles).
    // Micro mode: every tile is active, so the warp's tile is known up front and the loads of its queue extent
    // and of its visibility-buffer entries (below) go out together -- one dependent round trip less per tile.
    int4 rec;
    unsigned long long* vis = nullptr;
    unsigned long long v0 = ~0ull, v1 = ~0ull;
    if (f.microMode != 0) {
#if CRB_FINE_REVERSE
        // tiles in REVERSE order: for a mesh submitted in screen order the records setup wrote LAST are still in L2, and the
        // fine raster, which starts right after setup, meets them first (a stack instead of a cyclic sweep through the L2)
        const int t = max(f.numTiles - 1 - activeIdx, 0);
#else
        const int t = min(activeIdx, f.numTiles - 1);
#endif
        rec = make_int4(t, __ldg(&f.tileStart[t]), __ldg(&f.tileCount[t]), 0);
    } else {
        rec = __ldg(&f.activeRecs[activeIdx]);
    }
    const int tileIdx = rec.x;
    const int tileY = tileIdx / f.widthTiles, tileX = tileIdx - tileY * f.widthTiles;   // (the kernel's one integer division)
    if (f.microMode != 0) {
        vis = f.visBuffer + (size_t)((tileY << CR_TILE_LOG2) + (lane >> 3)) * f.widthPixels + (tileX << CR_TILE_LOG2) + (lane & 7);
#if CRB_FINE_STREAM & 1
        v0 = __ldcs(vis);
        v1 = __ldcs(vis + (size_t)4 * f.widthPixels);
#else
        v0 = vis[0];
        v1 = vis[(size_t)4 * f.widthPixels];
#endif
    }
    if (f.atomics->overflow != 0) return;
    if (f.microMode != 0) {
        if (activeIdx >= f.numTiles) return;
        if (f.atomics->numQueuedCtas == 0) rec.z = 0;   // nothing was queued in this frame: directAllocKernel did not run, the extents are stale
    } else if (activeIdx >= f.atomics->numActiveTiles) return;

    BlendShaderClass blendProbe;
    const bool deferred = !blendProbe.needsDst() && (FragmentShaderClass::CanDiscard == 0);
    ProfTimer<ProfMode> tmTotal, tm;   // ProfilingMode_Timers only
    tmTotal.start();
    tm.start();

    FineBatch& sb = s_batch[warp];
    const int queueStart = rec.y;
    const int queueCount = rec.z;
    const S32* __restrict__ queue = f.tileQueue + queueStart;

    // software pipeline: entries two batches ahead, header + depth plane one batch ahead
    S32 entryB = (32 + lane < queueCount) ? __ldg(&queue[32 + lane]) : -1;
    FineFetch cur;
    fineFetch<RenderModeFlags>(cur, f, lane < queueCount ? __ldg(&queue[lane]) : -1);

    // this lane's two pixels: (lx, ly) and (lx, ly + 4)
    const int lx = lane & 7, ly = lane >> 3;
    const int pixelX = (tileX << CR_TILE_LOG2) + lx;
    const int pixelY0 = (tileY << CR_TILE_LOG2) + ly;
    // colour surface: the reference's row-major layout, or TILE-MAJOR (crb_set_color_layout: the 64 texels of a tile are
    // contiguous, pixel (x, y) of the tile at y*8 + x): a warp then writes its tile as two full 128-byte lines, which is
    // what a frame slot in a PEER GPU's memory wants -- the stores cross NVLink as two large packets instead of eight 32-byte ones
    U32* colorPtr = f.colorTiled ? f.colorBuffer + (size_t)tileIdx * CR_TILE_SQR + lane : f.colorBuffer + (size_t)pixelY0 * f.colorPitch + pixelX;
    const size_t colorStep = f.colorTiled ? (size_t)32 : (size_t)4 * f.colorPitch;
    U32* depthPtr = f.depthBuffer + (size_t)pixelY0 * f.surfacePitch + pixelX;
    const size_t rowStep = (size_t)4 * f.surfacePitch;

    U32 color[2], depth[2];
    S32 winner[2] = {-1, -1};   // queue entry of the visible fragment (deferred mode)
    if (f.deferredClear) {
        color[0] = color[1] = f.clearColor;
        depth[0] = depth[1] = f.clearDepth;
    } else {
        color[0] = colorPtr[0]; color[1] = colorPtr[colorStep];
        depth[0] = kDepth ? depthPtr[0] : 0u; depth[1] = kDepth ? depthPtr[rowStep] : 0u;
    }

    if (deferred && kDepth && f.microMode != 0) {
        // Micro-triangle visibility (TriangleSetup.cuh microRaster): the (depth, entry + 1) minimum over the triangles that
        // setup rasterized itself (loaded at the top of the kernel).  It joins the tile state under the same rule as a
        // queued fragment, and the buffer gets its neutral value back for the next frame.
#if CRB_FINE_STREAM & 1
        if (v0 != ~0ull) __stcs(vis, ~0ull);
        if (v1 != ~0ull) __stcs(vis + (size_t)4 * f.widthPixels, ~0ull);
#else
        if (v0 != ~0ull) vis[0] = ~0ull;
        if (v1 != ~0ull) vis[(size_t)4 * f.widthPixels] = ~0ull;
#endif
        if (v0 != ~0ull && (U32)(v0 >> 32) < depth[0]) { depth[0] = (U32)(v0 >> 32); winner[0] = (S32)(U32)v0 - 1; }
        if (v1 != ~0ull && (U32)(v1 >> 32) < depth[1]) { depth[1] = (U32)(v1 >> 32); winner[1] = (S32)(U32)v1 - 1; }
    }

    // sample-space origin: centre of pixel (0,0) of the tile, viewport-centred subpixels
    const S32 bx = (tileX << (CR_TILE_LOG2 + CR_SUBPIXEL_LOG2)) + (CR_SUBPIXEL_SIZE >> 1) - f.originX;
    const S32 by = (tileY << (CR_TILE_LOG2 + CR_SUBPIXEL_LOG2)) + (CR_SUBPIXEL_SIZE >> 1) - f.originY;
    U32 profFrags = 0, profZTests = 0, profZKills = 0;   // ProfilingMode_Counters only
    tm.stop(f, CRB_TIMER_FineReadTile);

    for (int base = 0; base < queueCount; base += 32) {
        // ---- issue the loads of the batches ahead
        FineFetch nxt;
        fineFetch<RenderModeFlags>(nxt, f, entryB);
        entryB = (base + 64 + lane < queueCount) ? __ldg(&queue[base + 64 + lane]) : -1;

        // ---- (1) lane j: coverage mask of triangle j
        tm.start();
        U32 tileZMax = 0xFFFFFFFFu;
        if (kDepth) tileZMax = __reduce_max_sync(0xFFFFFFFFu, max(depth[0], depth[1]));
        U32 maskLo = 0, maskHi = 0;
        {
            const S32 x0 = (S32)(S16)(cur.h.x & 0xFFFF), y0 = (S32)cur.h.x >> 16;
            const S32 x1 = (S32)(S16)(cur.h.y & 0xFFFF), y1 = (S32)cur.h.y >> 16;
            const S32 x2 = (S32)(S16)(cur.h.z & 0xFFFF), y2 = (S32)cur.h.z >> 16;
            const S32 loX = min(min(x0, x1), x2), hiX = max(max(x0, x1), x2), loY = min(min(y0, y1), y2), hiY = max(max(y0, y1), y2);
            // pixel centres of the tile inside the triangle's bounding box
            const int colLo = max((loX - bx + (CR_SUBPIXEL_SIZE - 1)) >> CR_SUBPIXEL_LOG2, 0), colHi = min((hiX - bx) >> CR_SUBPIXEL_LOG2, CR_TILE_SIZE - 1);
            const int rowLo = max((loY - by + (CR_SUBPIXEL_SIZE - 1)) >> CR_SUBPIXEL_LOG2, 0), rowHi = min((hiY - by) >> CR_SUBPIXEL_LOG2, CR_TILE_SIZE - 1);
            // early Z against the tile's farthest depth; on the direct path (unordered queue) a triangle AT that depth
            // may still beat a later-submitted winner of the same depth, so only strictly farther ones are dropped
            const U32 zminHdr = cur.h.w & 0xFFFFF000u;
            const U32 zbTile = cur.z.z + cur.z.x * (U32)(tileX << CR_TILE_LOG2) + cur.z.y * (U32)(tileY << CR_TILE_LOG2);
            const bool inTile = cur.entry >= 0 && colLo <= colHi && rowLo <= rowHi;
            const bool culledZ = kDepth && inTile && earlyZCull(zminHdr, tileZMax, f.directMode != 0, zbTile, cur.z.x, cur.z.y, colLo, colHi, rowLo, rowHi);
            const bool live = inTile && !culledZ;
            const bool small = (colHi - colLo < 4) & (rowHi - rowLo < 4) & (hiX - loX < (64 << CR_SUBPIXEL_LOG2)) & (hiY - loY < (64 << CR_SUBPIXEL_LOG2));
            if ((f.debugFlags & 1) == 0 && __all_sync(0xFFFFFFFFu, !live || small)) {
                if (live) coverSmall4x4(x0, y0, x1, y1, x2, y2, bx, by, colLo, rowLo, colHi - colLo + 1, rowHi - rowLo + 1, maskLo, maskHi);
            } else if (live) {
                S32 a[3], b[3], c[3];
                setupTileEdges(cur.h, bx, by, a, b, c);
                coverTileRows(a, b, c, rowLo, rowHi, maskLo, maskHi);
            }
            if (ProfMode == ProfilingMode_Counters) {   // reference: FineRaster.inl:235, :621-622
                const bool fetched = cur.entry >= 0;
                const bool earlyZ = culledZ;
                const bool considered = fetched && !earlyZ;
                profCountWarp<ProfMode>(f, CRB_PROF_FineEarlyZCull, earlyZ, fetched);
                profCountWarp<ProfMode>(f, CRB_PROF_FineEmptyCull, considered && (maskLo | maskHi) == 0, considered);
                const U32 frags = __reduce_add_sync(0xFFFFFFFFu, (U32)(__popc(maskLo) + __popc(maskHi)));
                const U32 tris = __popc(__ballot_sync(0xFFFFFFFFu, considered));
                if (lane == 0) profCount<ProfMode>(f, CRB_PROF_FineFragPerTri, frags, tris);
                profFrags += frags;
            }
        }
        if (kDepth) {
            sb.zx[lane] = cur.z.x; sb.zy[lane] = cur.z.y;
            sb.zb[lane] = cur.z.z + cur.z.x * (U32)(tileX << CR_TILE_LOG2) + cur.z.y * (U32)(tileY << CR_TILE_LOG2);
        }
        sb.entry[lane] = cur.entry;
        sb.dataIdx[lane] = cur.dataIdx;
        __syncwarp();

        tm.stop(f, CRB_TIMER_FinePixelCoverage);
        tm.start();
        // ---- (2) transpose: triangles covering this lane's two pixels
        const U32 cover[2] = {warpTranspose32(maskLo, lane), warpTranspose32(maskHi, lane)};

        // ---- (3) ownership loop, queue order
        // Quads mode, in-order shading (reference: FineRaster.inl:396-430, :705-724): the four lanes of a 2x2
        // quad walk the UNION of their triangle sets together, so that the shader runs converged on the quad
        // (helper pixels included, whatever their coverage or depth) and dFdx / dFdy can shuffle.
#pragma unroll
        for (int p = 0; p < 2; p++) {
            U32 w = cover[p];
            if (kQuads && !deferred) {
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 1);
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 8);
            }
            while (w) {
                const int j = __ffs(w) - 1;
                w &= w - 1;
                const bool covered = !(kQuads && !deferred) || ((cover[p] >> j) & 1) != 0;
                U32 z = 0;
                bool zkill = false;
                if (kDepth) {
                    z = sb.zb[j] + sb.zx[j] * (U32)lx + sb.zy[j] * (U32)(ly + 4 * p);
                    zkill = z >= depth[p];
                    if (ProfMode == ProfilingMode_Counters && covered) { profZTests++; profZKills += zkill ? 1 : 0; }
                    // Direct path: the queue is unordered, the survivor of a pixel is the (depth, submission index)
                    // minimum -- what the strict LESS test leaves when fragments arrive in submission order.  A tie
                    // with the depth the tile started with (winner < 0) still fails.
                    if (deferred && f.directMode != 0 && z == depth[p] && winner[p] >= 0 && sb.entry[j] < winner[p]) zkill = false;
                    if (!(kQuads && !deferred) && zkill) continue;
                }
                if (deferred) {
                    if (kDepth) depth[p] = z;
                    winner[p] = sb.entry[j];
                } else {
                    const S32 entry = sb.entry[j];
                    FragmentShaderClass fs;
                    runFragmentShader<VertexClass, FragmentShaderClass, 0, RenderModeFlags>(fs, f, entry >> 3, sb.dataIdx[j], pixelX, pixelY0 + 4 * p, 0x11u);
                    if (!covered || zkill || fs.m_discard) continue;
                    if (kDepth) depth[p] = z;
                    BlendShaderClass bs;
                    runBlendShader(bs, entry >> 3, pixelX, pixelY0 + 4 * p, 0, fs.m_color, color[p]);
                    if (bs.m_writeColor) color[p] = bs.m_color;
                }
            }
        }
        __syncwarp();
        tm.stop(f, CRB_TIMER_FineZKill);   // transposes + the per-pixel depth / ownership loop (in-order pipes: shading and blending too)
        cur = nxt;
    }

    tm.start();
    if (deferred && kQuads) {
        // Visibility is resolved; every quad shades each DISTINCT winner of its four pixels once, on all four
        // lanes (derivatives need the neighbours' values of the same triangle), and a lane keeps the colour
        // of its own winner.  The shuffles that collect the winners run before any divergence.
        S32 q[2][4];
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int k = 0; k < 4; k++) q[p][k] = __shfl_sync(0xFFFFFFFFu, winner[p], (lane & 0x16) | (k & 1) | ((k & 2) << 2));
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const S32 e = q[p][k];
                bool skip = e < 0;
#pragma unroll
                for (int kk = 0; kk < k; kk++) skip |= q[p][kk] == e;
                if (skip) continue;   // uniform over the quad
                FragmentShaderClass fs;
                runFragmentShader<VertexClass, FragmentShaderClass, 0, RenderModeFlags>(fs, f, e >> 3, resolveDataIdx(e, f.triHeader), pixelX, pixelY0 + 4 * p, 0x11u);
                if (winner[p] == e) {
                    BlendShaderClass bs;
                    runBlendShader(bs, e >> 3, pixelX, pixelY0 + 4 * p, 0, fs.m_color, 0u);
                    if (bs.m_writeColor) color[p] = bs.m_color;
                }
            }
    } else if (deferred && (winner[0] & winner[1]) >= 0) {
        // Shade only the visible fragment of each pixel.  Both pixels are shaded in one straight
        // line of code (a lane with a single covered pixel shades that fragment twice) so that the
        // two chains of dependent loads -- plane rows, then vertex varyings -- overlap.
        const S32 e0 = winner[0] >= 0 ? winner[0] : winner[1], e1 = winner[1] >= 0 ? winner[1] : winner[0];
        const S32 d0 = resolveDataIdx(e0, f.triHeader), d1 = resolveDataIdx(e1, f.triHeader);
        FragmentShaderClass fs0, fs1;
        runFragmentShader<VertexClass, FragmentShaderClass, 0, RenderModeFlags>(fs0, f, e0 >> 3, d0, pixelX, pixelY0 + (winner[0] >= 0 ? 0 : 4), 0x11u);
        runFragmentShader<VertexClass, FragmentShaderClass, 0, RenderModeFlags>(fs1, f, e1 >> 3, d1, pixelX, pixelY0 + (winner[1] >= 0 ? 4 : 0), 0x11u);
        BlendShaderClass bs0, bs1;
        runBlendShader(bs0, e0 >> 3, pixelX, pixelY0, 0, fs0.m_color, 0u);
        runBlendShader(bs1, e1 >> 3, pixelX, pixelY0 + 4, 0, fs1.m_color, 0u);
        if (winner[0] >= 0 && bs0.m_writeColor) color[0] = bs0.m_color;
        if (winner[1] >= 0 && bs1.m_writeColor) color[1] = bs1.m_color;
    }

    tm.stop(f, CRB_TIMER_FineShade);       // visibility-first pipes: the shading of the surviving fragments
    tm.start();
    if (ProfMode == ProfilingMode_Counters) {   // reference: FineRaster.inl:221, :684, :702, :733-734
        __syncwarp();
        const U32 zt = __reduce_add_sync(0xFFFFFFFFu, profZTests), zk = __reduce_add_sync(0xFFFFFFFFu, profZKills);
        if (lane == 0) {
            profCount<ProfMode>(f, CRB_PROF_FineZKill, 100ull * zk, zt);
            profCount<ProfMode>(f, CRB_PROF_FineTriPerTile, (U32)queueCount, 1);
            profCount<ProfMode>(f, CRB_PROF_FineFragPerTile, profFrags, 1);
            profCount<ProfMode>(f, CRB_PROF_SetupSamplesPerTri, profFrags, 0);
        }
    }

    if (f.colorTiled) {
        // Tile-major colour surface (a frame slot in a peer GPU's memory): the tile's 64 texels are contiguous, so the warp stages
        // them in shared memory and ONE lane hands the 256 bytes to the bulk-copy engine (cp.async.bulk shared -> global, TMA):
        // the tile crosses NVLink as one 256-byte write instead of 64 four-byte stores.
        U32* const st = s_tileStage[warp];
        st[lane] = color[0];
        st[lane + 32] = color[1];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the generic-proxy stores above, before the async proxy reads them
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 256;" ::"l"(f.colorBuffer + (size_t)tileIdx * CR_TILE_SQR), "r"((U32)__cvta_generic_to_shared(st)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the staging buffer must outlive the read
        }
    } else {
        colorPtr[0] = color[0];
        colorPtr[colorStep] = color[1];
    }
    if (kDepth || f.deferredClear) {
        depthPtr[0] = depth[0];
        depthPtr[rowStep] = depth[1];
    }
    tm.stop(f, CRB_TIMER_FineWriteTile);
    tmTotal.stop(f, CRB_TIMER_FineTotal);
}

}  // namespace FW
