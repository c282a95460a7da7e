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

template <class VertexClass, class FragmentShaderClass, class BlendShaderClass, int SamplesLog2, U32 RenderModeFlags>
__global__ void __launch_bounds__(FineWarps<SamplesLog2>::Value * 32) fineRasterMultiKernel(const __grid_constant__ crb_frame f) {
    constexpr int N = 1 << SamplesLog2;
    constexpr int kWarps = FineWarps<SamplesLog2>::Value;
    constexpr bool kDepth = (RenderModeFlags & RenderModeFlag_EnableDepth) != 0;
    __shared__ __align__(16) FineTriRec s_recs[kWarps][32];
    __shared__ U32 s_depth[kWarps][N * CR_TILE_SQR];
    __shared__ U32 s_aux[kWarps][N * CR_TILE_SQR];   // colour (immediate mode) or winner position (deferred mode)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int activeIdx = blockIdx.x * kWarps + warp;
    if (f.atomics->overflow != 0) return;
    if (activeIdx >= f.atomics->numActiveTiles) return;

    BlendShaderClass blendProbe;
    const bool deferred = !blendProbe.needsDst() && (FragmentShaderClass::CanDiscard == 0);

    FineTriRec* recs = s_recs[warp];
    U32* tDepth = s_depth[warp];
    U32* tAux = s_aux[warp];
    const int tileIdx = __ldg(&f.activeTiles[activeIdx]);
    const int tileY = tileIdx / f.widthTiles, tileX = tileIdx - tileY * f.widthTiles;
    const int queueStart = __ldg(&f.tileStart[tileIdx]);
    const int queueCount = __ldg(&f.tileCount[tileIdx]);

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

    const S32 sx0 = lx << CR_SUBPIXEL_LOG2;
    const S32 sy0 = ly << CR_SUBPIXEL_LOG2;

    for (int base = 0; base < queueCount; base += 32) {
        U32 liveMask = fineRefill<SamplesLog2, RenderModeFlags>(f, recs, queueStart + base, queueCount - base, tileX, tileY);
        while (liveMask) {
            const int j = __ffs(liveMask) - 1;
            liveMask &= liveMask - 1;
            const uint4 r0 = reinterpret_cast<const uint4*>(&recs[j])[0];
            const uint4 r1 = reinterpret_cast<const uint4*>(&recs[j])[1];
            const uint4 r2 = reinterpret_cast<const uint4*>(&recs[j])[2];
            const S32 ea[3] = {(S32)r0.x, (S32)r0.w, (S32)r1.z}, eb[3] = {(S32)r0.y, (S32)r1.x, (S32)r1.w};
#pragma unroll
            for (int p = 0; p < 2; p++) {
                // edge values at the centre of pixel p, then per-sample offsets
                const S32 sy = sy0 + p * (4 << CR_SUBPIXEL_LOG2);
                const S32 ec[3] = {(S32)r0.z + ea[0] * sx0 + eb[0] * sy, (S32)r1.y + ea[1] * sx0 + eb[1] * sy, (S32)r2.x + ea[2] * sx0 + eb[2] * sy};
                const U32 cover = pixelSampleMask<SamplesLog2>(ea, eb, ec);
                if (cover == 0) continue;
                const int qBase = lane + 32 * p;
                const U32 zPix = r2.w + r2.y * (U32)(lx * N) + r2.z * (U32)((ly + 4 * p) * N);
                U32 pass = 0;
                U32 z[N];
#pragma unroll
                for (int i = 0; i < N; i++) {
                    z[i] = zPix + r2.y * (U32)msaaSampleX(SamplesLog2, i) + r2.z * (U32)i;
                    if (((cover >> i) & 1) && (!kDepth || z[i] < tDepth[i * CR_TILE_SQR + qBase])) pass |= 1u << i;
                }
                if (pass == 0) continue;
                if (deferred) {
#pragma unroll
                    for (int i = 0; i < N; i++)
                        if ((pass >> i) & 1) {
                            if (kDepth) tDepth[i * CR_TILE_SQR + qBase] = z[i];
                            tAux[i * CR_TILE_SQR + qBase] = (U32)(base + j + 1);
                        }
                } else {
                    const uint4 r3 = reinterpret_cast<const uint4*>(&recs[j])[3];
                    FragmentShaderClass fs;
                    runFragmentShader<VertexClass, FragmentShaderClass, SamplesLog2, RenderModeFlags>(fs, f, (int)r3.y, (int)r3.x, pixelX, pixelY0 + 4 * p, centroidCode<SamplesLog2>(cover));
                    if (fs.m_discard) continue;
#pragma unroll
                    for (int i = 0; i < N; i++)
                        if ((pass >> i) & 1) {
                            if (kDepth) tDepth[i * CR_TILE_SQR + qBase] = z[i];
                            BlendShaderClass bs;
                            runBlendShader(bs, (int)r3.y, pixelX, pixelY0 + 4 * p, i, fs.m_color, tAux[i * CR_TILE_SQR + qBase]);
                            if (bs.m_writeColor) tAux[i * CR_TILE_SQR + qBase] = bs.m_color;
                        }
                }
            }
        }
        __syncwarp();
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
                const S32 entry = __ldg(&f.tileQueue[queueStart + (int)win - 1]);
                const S32 dataIdx = resolveDataIdx(entry, f.triHeader);
                const uint4 h = __ldg(&f.triHeader[dataIdx]);
                const S32 bx = (pixelX << CR_SUBPIXEL_LOG2) + (CR_SUBPIXEL_SIZE >> 1) - f.originX;
                const S32 by = (pixelY << CR_SUBPIXEL_LOG2) + (CR_SUBPIXEL_SIZE >> 1) - f.originY;
                S32 a[3], b[3], c[3];
                setupTileEdges(h, bx, by, a, b, c);
                const U32 cover = pixelSampleMask<SamplesLog2>(a, b, c);
                FragmentShaderClass fs;
                runFragmentShader<VertexClass, FragmentShaderClass, SamplesLog2, RenderModeFlags>(fs, f, entry >> 3, dataIdx, pixelX, pixelY, centroidCode<SamplesLog2>(cover));
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
}

// Picks the kernel variant for a pipe; one warp per tile, surplus warps exit at once (the number
// of active tiles is only known on the device).
template <class VertexClass, class FragmentShaderClass, class BlendShaderClass, int SamplesLog2, U32 RenderModeFlags>
struct FineRasterLauncher {
    static int launch(const crb_frame* f, void* stream) {
        if ((RenderModeFlags & RenderModeFlag_EnableQuads) != 0) return CRB_ERR_INVALID;
        constexpr int kWarps = FineWarps<SamplesLog2>::Value;
        const int blocks = (f->numTiles + kWarps - 1) / kWarps;
        fineRasterMultiKernel<VertexClass, FragmentShaderClass, BlendShaderClass, SamplesLog2, RenderModeFlags><<<blocks, kWarps * 32, 0, (cudaStream_t)stream>>>(*f);
        return cudaGetLastError() == cudaSuccess ? CRB_OK : CRB_ERR_CUDA;
    }
};

template <class VertexClass, class FragmentShaderClass, class BlendShaderClass, U32 RenderModeFlags>
struct FineRasterLauncher<VertexClass, FragmentShaderClass, BlendShaderClass, 0, RenderModeFlags> {
    static int launch(const crb_frame* f, void* stream) {
        if ((RenderModeFlags & RenderModeFlag_EnableQuads) != 0) return CRB_ERR_INVALID;
        const int blocks = (f->numTiles + CRB_FINE_WARPS - 1) / CRB_FINE_WARPS;
        fineRasterSingleKernel<VertexClass, FragmentShaderClass, BlendShaderClass, RenderModeFlags><<<blocks, CRB_FINE_WARPS * 32, 0, (cudaStream_t)stream>>>(*f);
        return cudaGetLastError() == cudaSuccess ? CRB_OK : CRB_ERR_CUDA;
    }
};

}  // namespace FW
