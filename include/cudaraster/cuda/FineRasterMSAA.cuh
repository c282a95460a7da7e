// Stage 4, multi-sample variant (2/4/8 samples per pixel) + the launcher that picks the variant.
//
// Semantics follow the reference's fineRasterImpl_MultiSample (src/cudaraster/cuda/
// FineRaster.inl:855-1126): exact coverage at the N-rooks sample positions
// (cuda/Util.inl:340-383), per-sample depth zx*(px*N + X_i) + zy*(py*N + i) + zb with a strict
// LESS test, ONE shader run per (triangle, pixel) at the centroid sample chosen from the
// COVERAGE mask, per-sample blend, surface layout with the N samples of a tile stored as N
// horizontally adjacent 8x8 blocks.
//
// Design differences: like the single-sample kernel every lane owns two pixels (so all N samples
// of a pixel are updated by one thread, in queue order -- the reference does N read-modify-writes
// per fragment directly on the global surface and arbitrates intra-warp conflicts through shared
// memory).  The per-sample state (depth + colour or winner) of the tile lives in shared memory,
// [sample][pixel] so that warp accesses are conflict free, and goes to HBM once per tile.
#pragma once
#include "FineRaster.cuh"

namespace FW {

template <int SamplesLog2>
struct MsaaGeom {
    enum { N = 1 << SamplesLog2 };
    // subpixel offset of sample i from the pixel centre (reference: cuda/Util.inl:361-366)
    __device__ static __forceinline__ S32 offX(int i) { return (msaaSampleX(SamplesLog2, i) * 2 + 1 - N) << (CR_SUBPIXEL_LOG2 - SamplesLog2 - 1); }
    __device__ static __forceinline__ S32 offY(int i) { return (i * 2 + 1 - N) << (CR_SUBPIXEL_LOG2 - SamplesLog2 - 1); }
};

template <int SamplesLog2> struct FineWarps { enum { Value = SamplesLog2 >= 3 ? 4 : 8 }; };

// Sample coverage of one pixel whose centre is the origin of the edge equations (a,b,c).
template <int SamplesLog2>
__device__ __forceinline__ U32 pixelSampleMask(const S32 (&a)[3], const S32 (&b)[3], const S32 (&c)[3]) {
    U32 m = 0;
#pragma unroll
    for (int i = 0; i < (1 << SamplesLog2); i++) {
        const S32 ox = MsaaGeom<SamplesLog2>::offX(i), oy = MsaaGeom<SamplesLog2>::offY(i);
        const S32 e0 = c[0] + a[0] * ox + b[0] * oy, e1 = c[1] + a[1] * ox + b[1] * oy, e2 = c[2] + a[2] * ox + b[2] * oy;
        if ((e0 | e1 | e2) >= 0) m |= 1u << i;
    }
    return m;
}

// Per-warp staging of the current batch for the multi-sample kernel: edge equations relative to
// the centre of pixel (0,0) of the tile (needed again per fragment for the exact sample test),
// depth plane relative to the tile, queue entry and record slot.  Lane-indexed SoA.
struct FineBatchMSAA {
    S32 a0[32], b0[32], c0[32], a1[32], b1[32], c1[32], a2[32], b2[32], c2[32];
    U32 zx[32], zy[32], zb[32];
    S32 entry[32];
    S32 dataIdx[32];
    U32 zslope[32], zminHdr[32];   // quads mode only: inputs of the reference's per-pixel conservative depth kill
};

// Same scheme as the single-sample kernel (FineRaster.cuh): lane j builds a 64-bit PIXEL mask of
// triangle j -- conservative: the edge functions are relaxed by the largest sample offset, so the
// mask is a superset of the pixels with a covered sample -- two warp transposes hand every lane
// the triangles touching its two pixels, and the ownership loop tests the N samples of those
// fragments exactly, in queue order.
template <class VertexClass, class FragmentShaderClass, class BlendShaderClass, int SamplesLog2, U32 RenderModeFlags, int ProfMode = ProfilingMode_Default>
static __global__ void __launch_bounds__(FineWarps<SamplesLog2>::Value * 32) fineRasterMultiKernel(const __grid_constant__ crb_frame f) {
    constexpr int N = 1 << SamplesLog2;
    constexpr int kWarps = FineWarps<SamplesLog2>::Value;
    constexpr bool kDepth = (RenderModeFlags & RenderModeFlag_EnableDepth) != 0;
    constexpr bool kQuads = (RenderModeFlags & RenderModeFlag_EnableQuads) != 0;
    constexpr S32 kMaxOfs = (N - 1) << (CR_SUBPIXEL_LOG2 - SamplesLog2 - 1);   // largest |sample offset| from the pixel centre, subpixels
    __shared__ FineBatchMSAA s_batch[kWarps];
    __shared__ U32 s_depth[kWarps][N * CR_TILE_SQR];
    __shared__ U32 s_aux[kWarps][N * CR_TILE_SQR];   // colour (immediate mode) or winner entry + 1 (deferred mode)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int activeIdx = blockIdx.x * kWarps + warp;
    gridDepLaunchDependents();
    gridDepWait();
    finishFrameState(f);
    const int4 rec = __ldg(&f.activeRecs[activeIdx]);
    if (f.atomics->overflow != 0) return;
    if (activeIdx >= f.atomics->numActiveTiles) return;

    BlendShaderClass blendProbe;
    // quads mode shades in order: where a helper pixel is shaded depends on the depth state at that moment (below)
    const bool deferred = !kQuads && !blendProbe.needsDst() && (FragmentShaderClass::CanDiscard == 0);

    ProfTimer<ProfMode> tmTotal;   // ProfilingMode_Timers: the multi-sample kernel reports its total only
    tmTotal.start();
    FineBatchMSAA& sb = s_batch[warp];
    U32* tDepth = s_depth[warp];
    U32* tAux = s_aux[warp];
    const int tileIdx = rec.x;
    const int tileY = tileIdx / f.widthTiles, tileX = tileIdx - tileY * f.widthTiles;
    const int queueCount = rec.z;
    const S32* __restrict__ queue = f.tileQueue + rec.y;

    S32 entryB = (32 + lane < queueCount) ? __ldg(&queue[32 + lane]) : -1;
    FineFetch cur;
    fineFetch<RenderModeFlags>(cur, f, lane < queueCount ? __ldg(&queue[lane]) : -1);

    const int lx = lane & 7, ly = lane >> 3;
    const int pixelX = (tileX << CR_TILE_LOG2) + lx;
    const int pixelY0 = (tileY << CR_TILE_LOG2) + ly;
    // texel of sample i of this lane's pixel p: row pixelY0 + 4p, column tileX*8*N + i*8 + lx
    const size_t texel0 = (size_t)pixelY0 * f.surfacePitch + (size_t)tileX * (CR_TILE_SIZE * N) + lx;
    const size_t rowStep = (size_t)4 * f.surfacePitch;

    // ---- load / clear the tile state
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
        for (int i = 0; i < N; i++) {
            const int q = i * CR_TILE_SQR + lane + 32 * p;
            const size_t t = texel0 + p * rowStep + i * CR_TILE_SIZE;
            if (f.deferredClear) {
                tDepth[q] = f.clearDepth;
                tAux[q] = deferred ? 0u : f.clearColor;
            } else {
                tDepth[q] = kDepth ? f.depthBuffer[t] : 0u;
                tAux[q] = deferred ? 0u : f.colorBuffer[t];
            }
        }
    __syncwarp();

    const S32 bx = (tileX << (CR_TILE_LOG2 + CR_SUBPIXEL_LOG2)) + (CR_SUBPIXEL_SIZE >> 1) - f.originX;
    const S32 by = (tileY << (CR_TILE_LOG2 + CR_SUBPIXEL_LOG2)) + (CR_SUBPIXEL_SIZE >> 1) - f.originY;
    const S32 sx0 = lx << CR_SUBPIXEL_LOG2;
    const S32 sy0 = ly << CR_SUBPIXEL_LOG2;
    // quads mode: the reference kernel's tileDepth[pixel] -- CR_DEPTH_MAX until the pixel's first ROP of this
    // draw, then the maximum over its samples (FineRaster.inl:934-935, :1101-1108)
    U32 pixZMax[2] = {CR_DEPTH_MAX, CR_DEPTH_MAX};
    U32 profFrags = 0, profSamples = 0, profZTests = 0, profZKills = 0, profMsaaKills = 0;   // ProfilingMode_Counters only

    for (int base = 0; base < queueCount; base += 32) {
        FineFetch nxt;
        fineFetch<RenderModeFlags>(nxt, f, entryB);
        entryB = (base + 64 + lane < queueCount) ? __ldg(&queue[base + 64 + lane]) : -1;

        // ---- (1) lane j: conservative pixel mask of triangle j
        U32 tileZMax = 0xFFFFFFFFu;
        if (kDepth) {
            U32 m = 0;
#pragma unroll
            for (int i = 0; i < N; i++) m = max(m, max(tDepth[i * CR_TILE_SQR + lane], tDepth[i * CR_TILE_SQR + lane + 32]));
            tileZMax = __reduce_max_sync(0xFFFFFFFFu, m);
        }
        U32 maskLo = 0, maskHi = 0;
        bool culledZ = false;
        if (cur.entry >= 0) {
            S32 a[3], b[3], c[3];
            setupTileEdges(cur.h, bx, by, a, b, c);
            sb.a0[lane] = a[0]; sb.b0[lane] = b[0]; sb.c0[lane] = c[0];
            sb.a1[lane] = a[1]; sb.b1[lane] = b[1]; sb.c1[lane] = c[1];
            sb.a2[lane] = a[2]; sb.b2[lane] = b[2]; sb.c2[lane] = c[2];
            const U32 zminHdr = cur.h.w & 0xFFFFF000u;
            const S32 x0 = (S32)(S16)(cur.h.x & 0xFFFF), x1 = (S32)(S16)(cur.h.y & 0xFFFF), x2 = (S32)(S16)(cur.h.z & 0xFFFF);
            const S32 y0 = (S32)cur.h.x >> 16, y1 = (S32)cur.h.y >> 16, y2 = (S32)cur.h.z >> 16;
            // pixels of the tile with a sample inside the triangle's bounding box
            const int colLo = max((min(min(x0, x1), x2) - bx - kMaxOfs + (CR_SUBPIXEL_SIZE - 1)) >> CR_SUBPIXEL_LOG2, 0);
            const int colHi = min((max(max(x0, x1), x2) - bx + kMaxOfs) >> CR_SUBPIXEL_LOG2, CR_TILE_SIZE - 1);
            const int rowLo = max((min(min(y0, y1), y2) - by - kMaxOfs + (CR_SUBPIXEL_SIZE - 1)) >> CR_SUBPIXEL_LOG2, 0);
            const int rowHi = min((max(max(y0, y1), y2) - by + kMaxOfs) >> CR_SUBPIXEL_LOG2, CR_TILE_SIZE - 1);
            const U32 zbTile = cur.z.z + cur.z.x * (U32)(tileX << (CR_TILE_LOG2 + SamplesLog2)) + cur.z.y * (U32)(tileY << (CR_TILE_LOG2 + SamplesLog2));
            // early Z: only what provably cannot change the frame (earlyZCull, FineRaster.cuh); direct path: ties are kept
            culledZ = kDepth && colLo <= colHi && rowLo <= rowHi &&
                      earlyZCull(zminHdr, tileZMax, f.directMode != 0, zbTile, cur.z.x, cur.z.y, colLo * N, colHi * N + N - 1, rowLo * N, rowHi * N + N - 1);
            if (!culledZ && colLo <= colHi && rowLo <= rowHi) {
                // relax every edge by its largest possible gain over the sample offsets (|ox|, |oy| <= kMaxOfs);
                // c is clamped to +-2^30 and the relaxation is < 2^21, so nothing wraps
                S32 cr[3];
#pragma unroll
                for (int i = 0; i < 3; i++) cr[i] = c[i] + (abs(a[i]) + abs(b[i])) * kMaxOfs;
                coverTileRows(a, b, cr, rowLo, rowHi, maskLo, maskHi);
            }
        }
        if (ProfMode == ProfilingMode_Counters) {   // reference: FineRaster.inl:235, :969-970 (pixel mask = conservative coverage)
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
        if (kDepth) {
            sb.zx[lane] = cur.z.x; sb.zy[lane] = cur.z.y;
            sb.zb[lane] = cur.z.z + cur.z.x * (U32)(tileX << (CR_TILE_LOG2 + SamplesLog2)) + cur.z.y * (U32)(tileY << (CR_TILE_LOG2 + SamplesLog2));
        }
        sb.entry[lane] = cur.entry;
        sb.dataIdx[lane] = cur.dataIdx;
        if (kQuads && kDepth) {
            sb.zslope[lane] = cur.z.w;
            sb.zminHdr[lane] = cur.h.w & 0xFFFFF000u;
        }
        __syncwarp();

        // ---- (2) transpose
        const U32 cover[2] = {warpTranspose32(maskLo, lane), warpTranspose32(maskHi, lane)};

        // ---- (3) ownership loop, queue order
#pragma unroll
        for (int p = 0; p < 2; p++) {
            U32 w = cover[p];
            if (kQuads) {   // the four lanes of a 2x2 quad walk the union of their sets together (see FineRaster.cuh)
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 1);
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 8);
            }
            const S32 sy = sy0 + p * (4 << CR_SUBPIXEL_LOG2);
            const int qBase = lane + 32 * p;
            while (w) {
                const int j = __ffs(w) - 1;
                w &= w - 1;
                const S32 ea[3] = {sb.a0[j], sb.a1[j], sb.a2[j]}, eb[3] = {sb.b0[j], sb.b1[j], sb.b2[j]};
                const S32 ec[3] = {sb.c0[j] + ea[0] * sx0 + eb[0] * sy, sb.c1[j] + ea[1] * sx0 + eb[1] * sy, sb.c2[j] + ea[2] * sx0 + eb[2] * sy};
                const U32 cov = pixelSampleMask<SamplesLog2>(ea, eb, ec);
                if (ProfMode == ProfilingMode_Counters && ((cover[p] >> j) & 1)) { profMsaaKills += cov == 0 ? 1 : 0; profSamples += __popc(cov); }
                if (!kQuads && cov == 0) continue;
                const U32 zxv = sb.zx[j], zyv = sb.zy[j];
                const U32 zPix = sb.zb[j] + zxv * (U32)(lx * N) + zyv * (U32)((ly + 4 * p) * N);
                U32 pass = 0;
                U32 z[N];
#pragma unroll
                for (int i = 0; i < N; i++) {
                    z[i] = zPix + zxv * (U32)msaaSampleX(SamplesLog2, i) + zyv * (U32)i;
                    if (((cov >> i) & 1) && (!kDepth || z[i] < tDepth[i * CR_TILE_SQR + qBase])) pass |= 1u << i;
                    // direct path (unordered queue): among equal depths the earliest-submitted fragment survives
                    if (kDepth && deferred && f.directMode != 0 && ((cov >> i) & 1) && z[i] == tDepth[i * CR_TILE_SQR + qBase]) {
                        const U32 held = tAux[i * CR_TILE_SQR + qBase];
                        if (held != 0 && (U32)sb.entry[j] + 1u < held) pass |= 1u << i;
                    }
                }
                if (ProfMode == ProfilingMode_Counters && cov != 0) { profZTests++; profZKills += pass == 0 ? 1 : 0; }
                if (!kQuads && pass == 0) continue;
                const S32 entry = sb.entry[j];
                if (kQuads) {
                    // Reference semantics (FineRaster.inl:1034-1111): a pixel the conservative per-pixel depth test
                    // kills has an EMPTY sample mask -- it is shaded at its centre (as a helper) and gets no ROP;
                    // a live, covered pixel is shaded at the centroid of its mask and the ROP refreshes its bound.
                    U32 mask = cov;
                    if (kDepth && cov != 0) {
                        const U32 zslope = sb.zslope[j];
                        const U32 zmin = ((zxv + zyv) << (SamplesLog2 - 1 > 0 ? SamplesLog2 - 1 : 0)) + zPix - zslope;
                        if ((zmin >= pixZMax[p] && zmin < zmin + zslope * 2u) || sb.zminHdr[j] >= pixZMax[p]) mask = 0;
                    }
                    FragmentShaderClass fs;
                    runFragmentShader<VertexClass, FragmentShaderClass, SamplesLog2, RenderModeFlags>(fs, f, entry >> 3, sb.dataIdx[j], pixelX, pixelY0 + 4 * p, centroidCode<SamplesLog2>(mask));
                    if (mask == 0 || fs.m_discard) continue;
                    U32 zmaxNew = 0;
#pragma unroll
                    for (int i = 0; i < N; i++) {
                        if ((pass >> i) & 1) {
                            if (kDepth) tDepth[i * CR_TILE_SQR + qBase] = z[i];
                            BlendShaderClass bs;
                            runBlendShader(bs, entry >> 3, pixelX, pixelY0 + 4 * p, i, fs.m_color, tAux[i * CR_TILE_SQR + qBase]);
                            if (bs.m_writeColor) tAux[i * CR_TILE_SQR + qBase] = bs.m_color;
                        }
                        if (kDepth) zmaxNew = max(zmaxNew, tDepth[i * CR_TILE_SQR + qBase]);
                    }
                    if (kDepth && zmaxNew < pixZMax[p]) pixZMax[p] = zmaxNew;
                } else if (deferred) {
#pragma unroll
                    for (int i = 0; i < N; i++)
                        if ((pass >> i) & 1) {
                            if (kDepth) tDepth[i * CR_TILE_SQR + qBase] = z[i];
                            tAux[i * CR_TILE_SQR + qBase] = (U32)entry + 1u;
                        }
                } else {
                    FragmentShaderClass fs;
                    runFragmentShader<VertexClass, FragmentShaderClass, SamplesLog2, RenderModeFlags>(fs, f, entry >> 3, sb.dataIdx[j], pixelX, pixelY0 + 4 * p, centroidCode<SamplesLog2>(cov));
                    if (fs.m_discard) continue;
#pragma unroll
                    for (int i = 0; i < N; i++)
                        if ((pass >> i) & 1) {
                            if (kDepth) tDepth[i * CR_TILE_SQR + qBase] = z[i];
                            BlendShaderClass bs;
                            runBlendShader(bs, entry >> 3, pixelX, pixelY0 + 4 * p, i, fs.m_color, tAux[i * CR_TILE_SQR + qBase]);
                            if (bs.m_writeColor) tAux[i * CR_TILE_SQR + qBase] = bs.m_color;
                        }
                }
            }
        }
        __syncwarp();
        cur = nxt;
    }

    if (ProfMode == ProfilingMode_Counters) {   // reference: FineRaster.inl:1032, :1069-1070, :1120-1121
        __syncwarp();
        const U32 zt = __reduce_add_sync(0xFFFFFFFFu, profZTests), zk = __reduce_add_sync(0xFFFFFFFFu, profZKills);
        const U32 mk = __reduce_add_sync(0xFFFFFFFFu, profMsaaKills), ns = __reduce_add_sync(0xFFFFFFFFu, profSamples);
        if (lane == 0) {
            profCount<ProfMode>(f, CRB_PROF_FineZKill, 100ull * zk, zt);
            profCount<ProfMode>(f, CRB_PROF_FineMSAAKill, 100ull * mk, profFrags);
            profCount<ProfMode>(f, CRB_PROF_FineTriPerTile, (U32)queueCount, 1);
            profCount<ProfMode>(f, CRB_PROF_FineFragPerTile, profFrags, 1);
            profCount<ProfMode>(f, CRB_PROF_SetupSamplesPerTri, ns, 0);
        }
    }

    // ---- resolve + write back
#pragma unroll 1
    for (int p = 0; p < 2; p++) {
        const int qBase = lane + 32 * p;
        const int pixelY = pixelY0 + 4 * p;
        if (deferred) {
            U32 handled = 0;
#pragma unroll 1
            for (int i0 = 0; i0 < N; i0++) {
                const U32 win = tAux[i0 * CR_TILE_SQR + qBase];
                if (win == 0 || ((handled >> i0) & 1)) continue;
                U32 group = 0;
                for (int i = i0; i < N; i++)
                    if (tAux[i * CR_TILE_SQR + qBase] == win) group |= 1u << i;
                handled |= group;
                // shade this (triangle, pixel) once, at the centroid of its COVERAGE mask
                const S32 entry = (S32)(win - 1u);
                const S32 dataIdx = resolveDataIdx(entry, f.triHeader);
                const uint4 h = __ldg(&f.triHeader[dataIdx]);
                const S32 px = (pixelX << CR_SUBPIXEL_LOG2) + (CR_SUBPIXEL_SIZE >> 1) - f.originX;
                const S32 py = (pixelY << CR_SUBPIXEL_LOG2) + (CR_SUBPIXEL_SIZE >> 1) - f.originY;
                S32 a[3], b[3], c[3];
                setupTileEdges(h, px, py, a, b, c);
                const U32 cov = pixelSampleMask<SamplesLog2>(a, b, c);
                FragmentShaderClass fs;
                runFragmentShader<VertexClass, FragmentShaderClass, SamplesLog2, RenderModeFlags>(fs, f, entry >> 3, dataIdx, pixelX, pixelY, centroidCode<SamplesLog2>(cov));
                for (int i = i0; i < N; i++)
                    if ((group >> i) & 1) {
                        BlendShaderClass bs;
                        runBlendShader(bs, entry >> 3, pixelX, pixelY, i, fs.m_color, 0u);
                        if (bs.m_writeColor) f.colorBuffer[texel0 + p * rowStep + i * CR_TILE_SIZE] = bs.m_color;
                        else if (f.deferredClear) f.colorBuffer[texel0 + p * rowStep + i * CR_TILE_SIZE] = f.clearColor;
                    }
            }
            if (f.deferredClear)
                for (int i = 0; i < N; i++)
                    if (tAux[i * CR_TILE_SQR + qBase] == 0) f.colorBuffer[texel0 + p * rowStep + i * CR_TILE_SIZE] = f.clearColor;
        } else {
            for (int i = 0; i < N; i++) f.colorBuffer[texel0 + p * rowStep + i * CR_TILE_SIZE] = tAux[i * CR_TILE_SQR + qBase];
        }
        if (kDepth || f.deferredClear)
            for (int i = 0; i < N; i++) f.depthBuffer[texel0 + p * rowStep + i * CR_TILE_SIZE] = tDepth[i * CR_TILE_SQR + qBase];
    }
    tmTotal.stop(f, CRB_TIMER_FineTotal);
}

// Picks the kernel variant for a pipe; one warp per tile, surplus warps exit at once (the number
// of active tiles is only known on the device).
template <class VertexClass, class FragmentShaderClass, class BlendShaderClass, int SamplesLog2, U32 RenderModeFlags, int ProfMode = ProfilingMode_Default>
struct FineRasterLauncher {
    static int launch(const crb_frame* f, void* stream) {
        constexpr int kWarps = FineWarps<SamplesLog2>::Value;
        const int blocks = (f->numTiles + kWarps - 1) / kWarps;
        return launchChained(fineRasterMultiKernel<VertexClass, FragmentShaderClass, BlendShaderClass, SamplesLog2, RenderModeFlags, ProfMode>, blocks, kWarps * 32, (cudaStream_t)stream, *f) == cudaSuccess
                   ? CRB_OK : CRB_ERR_CUDA;
    }
};

template <class VertexClass, class FragmentShaderClass, class BlendShaderClass, U32 RenderModeFlags, int ProfMode>
struct FineRasterLauncher<VertexClass, FragmentShaderClass, BlendShaderClass, 0, RenderModeFlags, ProfMode> {
    static int launch(const crb_frame* f, void* stream) {
        const int blocks = (f->numTiles + CRB_FINE_WARPS - 1) / CRB_FINE_WARPS;
        return launchChained(fineRasterSingleKernel<VertexClass, FragmentShaderClass, BlendShaderClass, RenderModeFlags, ProfMode>, blocks, CRB_FINE_WARPS * 32, (cudaStream_t)stream, *f) == cudaSuccess
                   ? CRB_OK : CRB_ERR_CUDA;
    }
};

}  // namespace FW
