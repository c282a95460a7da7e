// MSAA resolve and surface read-back helpers.
//
// The reference declares CudaSurface::resolveToScreen (src/cudaraster/CudaSurface.hpp:73: "Resolves MSAA
// and writes pixels into the current GL render target") but the Linux port dropped its body together
// with the GL blit; the demo calls its own resolveToScreen() after drawTriangles (test/SceneCR.cpp:297).
// Here the resolve is a CUDA kernel over the tile-replicated sample layout (sample i of pixel (x, y) at
// texel ((x>>3)*8N + 8i + (x&7), y), FineRaster.inl:909, :1088-1100) into a LINEAR width x height image:
// box filter, every 8-bit channel = (sum of the N samples + N/2) >> log2 N (round half up).
//
// HBM-bound streaming kernel: one thread resolves 4 horizontally adjacent pixels with 128-bit loads (one
// per sample) and one 128-bit store; algorithmic bytes = 4 (N + 1) per pixel.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>

#include "../../include/crb200.h"

namespace {

template <int N>
__global__ void __launch_bounds__(256) resolveKernel(const uint4* __restrict__ src, int srcPitch4, uint32_t* __restrict__ dst, int dstPitch, int width, int height, int flipY) {
    // thread -> 4 pixels: x4 = 4-pixel group in the row
    const int groupsPerRow = (width + 3) >> 2;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= groupsPerRow * height) return;
    const int y = gid / groupsPerRow, x4 = gid - y * groupsPerRow;
    const int x = x4 << 2;
    // texel column of sample 0: (x>>3)*8N + (x&7); in uint4 units: (x>>3)*2N + ((x&7)>>2)
    const uint4* p = src + (size_t)y * srcPitch4 + (size_t)(x >> 3) * (2 * N) + ((x & 7) >> 2);
    uint32_t rb[4] = {0, 0, 0, 0}, ga[4] = {0, 0, 0, 0};   // two 16-bit lanes each: no overflow for N <= 8 (8 * 255 < 65536)
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint4 t = __ldg(p + 2 * i);
        const uint32_t v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            rb[k] += v[k] & 0x00FF00FFu;
            ga[k] += (v[k] >> 8) & 0x00FF00FFu;
        }
    }
    constexpr int L = N == 1 ? 0 : N == 2 ? 1 : N == 4 ? 2 : 3;
    constexpr uint32_t half = (uint32_t)(N >> 1) * 0x00010001u;
    uint32_t out[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t a = ((rb[k] + half) >> L) & 0x00FF00FFu;
        const uint32_t b = ((ga[k] + half) >> L) & 0x00FF00FFu;
        out[k] = a | (b << 8);
    }
    const int oy = flipY ? height - 1 - y : y;
    uint32_t* q = dst + (size_t)oy * dstPitch + x;
    if (x + 3 < width && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) {
        *reinterpret_cast<uint4*>(q) = make_uint4(out[0], out[1], out[2], out[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (x + k < width) q[k] = out[k];
    }
}

}  // namespace

extern "C" int crb_resolve_surface(const void* d_src, int width, int height, int numSamples, void* d_dst, int dstPitch, int flipY, void* stream) {
    if (!d_src || !d_dst || width <= 0 || height <= 0 || dstPitch < width) return CRB_ERR_INVALID;
    if (width > CRB_MAX_VIEWPORT || height > CRB_MAX_VIEWPORT) return CRB_ERR_LIMIT;
    const int roundedW = (width + 7) & ~7;
    const int srcPitch4 = roundedW * numSamples / 4;
    const int groups = ((width + 3) >> 2) * height;
    const int block = 256, grid = (groups + block - 1) / block;
    cudaStream_t s = (cudaStream_t)stream;
    const uint4* src = (const uint4*)d_src;
    uint32_t* dst = (uint32_t*)d_dst;
    switch (numSamples) {
        case 1: resolveKernel<1><<<grid, block, 0, s>>>(src, srcPitch4, dst, dstPitch, width, height, flipY); break;
        case 2: resolveKernel<2><<<grid, block, 0, s>>>(src, srcPitch4, dst, dstPitch, width, height, flipY); break;
        case 4: resolveKernel<4><<<grid, block, 0, s>>>(src, srcPitch4, dst, dstPitch, width, height, flipY); break;
        case 8: resolveKernel<8><<<grid, block, 0, s>>>(src, srcPitch4, dst, dstPitch, width, height, flipY); break;
        default: return CRB_ERR_INVALID;
    }
    return cudaGetLastError() == cudaSuccess ? CRB_OK : CRB_ERR_CUDA;
}

// ---- multi-GPU composite over peer memory: CUDA IPC plumbing (crb200.h) ---------------------------------------------
namespace {
__global__ void ipcSignalKernel(volatile uint32_t* word, uint32_t value) {
    __threadfence_system();
    *word = value;
}
}  // namespace

// ---- sort-first geometry cull: clip-space bounds per chunk of CRB_CHUNK_BOUNDS_TRIS consecutive triangles ------------------------
// One CTA per chunk, one thread per triangle: min / max of x/w and y/w over the chunk's vertices (block reduction); a chunk with
// a vertex at w <= 0 (its projection is unbounded) gets the whole plane.
__global__ void __launch_bounds__(256) chunkBoundsKernel(const float4* __restrict__ verts, int stride4, const int32_t* __restrict__ idx, int numTris, float4* __restrict__ bounds) {
    __shared__ float s_red[4][8];
    const int tri = blockIdx.x * 256 + threadIdx.x;
    const float inf = __int_as_float(0x7F800000);
    float lox = inf, loy = inf, hix = -inf, hiy = -inf;
    if (tri < numTris) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float4 v = __ldg(&verts[(size_t)__ldg(&idx[tri * 3 + k]) * stride4]);
            if (v.w > 0.0f && v.x == v.x && v.y == v.y) {
                const float x = v.x / v.w, y = v.y / v.w;
                lox = fminf(lox, x); hix = fmaxf(hix, x); loy = fminf(loy, y); hiy = fmaxf(hiy, y);
            } else {
                lox = loy = -inf; hix = hiy = inf;
            }
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        lox = fminf(lox, __shfl_xor_sync(0xFFFFFFFFu, lox, d)); loy = fminf(loy, __shfl_xor_sync(0xFFFFFFFFu, loy, d));
        hix = fmaxf(hix, __shfl_xor_sync(0xFFFFFFFFu, hix, d)); hiy = fmaxf(hiy, __shfl_xor_sync(0xFFFFFFFFu, hiy, d));
    }
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = lox; s_red[1][threadIdx.x >> 5] = loy; s_red[2][threadIdx.x >> 5] = hix; s_red[3][threadIdx.x >> 5] = hiy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { lox = fminf(lox, s_red[0][w]); loy = fminf(loy, s_red[1][w]); hix = fmaxf(hix, s_red[2][w]); hiy = fmaxf(hiy, s_red[3][w]); }
        bounds[blockIdx.x] = make_float4(lox, loy, hix, hiy);
    }
}

extern "C" int crb_compute_chunk_bounds(const void* d_vertices, int vertexStride, const int32_t* d_indices, int numTris, float* d_bounds, void* stream) {
    if (numTris < 0 || vertexStride < 16 || (vertexStride & 15) || (numTris > 0 && (!d_vertices || !d_indices || !d_bounds))) return CRB_ERR_INVALID;
    if (numTris == 0) return CRB_OK;
    chunkBoundsKernel<<<(numTris + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)d_vertices, vertexStride / 16, d_indices, numTris, (float4*)d_bounds);
    return cudaGetLastError() == cudaSuccess ? CRB_OK : CRB_ERR_CUDA;
}

extern "C" int crb_ipc_alloc(size_t bytes, void** d_ptr, unsigned char handle[CRB_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == CRB_IPC_HANDLE_BYTES, "handle size");
    if (!d_ptr || !handle || bytes == 0) return CRB_ERR_INVALID;
    if (cudaMalloc(d_ptr, bytes) != cudaSuccess) return CRB_ERR_CUDA;
    if (cudaMemset(*d_ptr, 0, bytes) != cudaSuccess) return CRB_ERR_CUDA;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, *d_ptr) != cudaSuccess) { cudaFree(*d_ptr); *d_ptr = nullptr; return CRB_ERR_CUDA; }
    memcpy(handle, &h, sizeof(h));
    return CRB_OK;
}

extern "C" int crb_ipc_free(void* d_ptr) { return !d_ptr || cudaFree(d_ptr) == cudaSuccess ? CRB_OK : CRB_ERR_CUDA; }

extern "C" int crb_ipc_open(const unsigned char handle[CRB_IPC_HANDLE_BYTES], void** d_ptr) {
    if (!d_ptr || !handle) return CRB_ERR_INVALID;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    return cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? CRB_OK : CRB_ERR_CUDA;
}

extern "C" int crb_ipc_close(void* d_ptr) { return !d_ptr || cudaIpcCloseMemHandle(d_ptr) == cudaSuccess ? CRB_OK : CRB_ERR_CUDA; }

extern "C" int crb_ipc_signal(void* d_word, uint32_t value, void* stream) {
    if (!d_word) return CRB_ERR_INVALID;
    ipcSignalKernel<<<1, 1, 0, (cudaStream_t)stream>>>((volatile uint32_t*)d_word, value);
    return cudaGetLastError() == cudaSuccess ? CRB_OK : CRB_ERR_CUDA;
}

extern "C" int crb_ipc_copy_2d(void* d_dst, size_t dstPitchBytes, const void* d_src, size_t srcPitchBytes, size_t widthBytes, size_t height, void* stream) {
    if (widthBytes && height && (!d_dst || !d_src)) return CRB_ERR_INVALID;
    return cudaMemcpy2DAsync(d_dst, dstPitchBytes, d_src, srcPitchBytes, widthBytes, height, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) == cudaSuccess ? CRB_OK : CRB_ERR_CUDA;
}

// Stream-ordered pause (device spins on %globaltimer): ranks that composite into one display GPU start their frame loops out
// of phase with it, so that their frame pushes interleave instead of meeting at the display GPU's NVLink ingress all at once.
__global__ void ipcDelayKernel(unsigned long long ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < ns);
}
extern "C" int crb_ipc_delay(void* stream, unsigned int nanoseconds) {
    if (nanoseconds == 0) return CRB_OK;
    ipcDelayKernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long)(nanoseconds > 100000000u ? 100000000u : nanoseconds));
    return cudaGetLastError() == cudaSuccess ? CRB_OK : CRB_ERR_CUDA;
}

extern "C" int crb_ipc_copy(void* d_dst, const void* d_src, size_t bytes, void* stream) {
    if (bytes && (!d_dst || !d_src)) return CRB_ERR_INVALID;
    return cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) == cudaSuccess ? CRB_OK : CRB_ERR_CUDA;
}
