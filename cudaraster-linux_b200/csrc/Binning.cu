// Stages 2 and 3 -- bin raster and coarse raster for sm_100a (pipe independent, live in
// libcrb200.so; the per-pipe <name>_binRaster / <name>_coarseRaster launchers forward here).
//
// The reference sorts triangles into 128x128 px bins with 16 persistent CTAs that append to
// per-(bin, CTA) linked lists of 512-entry segments (src/cudaraster/cuda/BinRaster.inl:19-469), and
// then lets ONE CTA per bin merge the 16 streams back into submission order and emit per-tile
// linked lists of 32-entry segments (cuda/CoarseRaster.inl:28-809).  On a 148-SM part that leaves
// most of the chip idle and spends its time in block-wide barriers.
//
// B200 design: sort-middle is a stable two-digit MSD radix sort with key expansion
// (triangle -> bins, bin entry -> tiles).  Each digit is a count / scan / scatter:
//
//   bin stage     count   : fused into triangle setup (one column of binCountMat per chunk)
//                 scan    : binScanKernel      one CTA per bin, exclusive scan over chunks,
//                                              bin base from one atomicAdd, emits coarse work items
//                 scatter : binScatterKernel   one CTA per chunk; warp-ballot masks give every
//                                              (triangle, bin) pair its stable rank
//   coarse stage  count   : coarseCountKernel  one CTA per work item (<= CRB_ITEM_ENTRIES entries
//                                              of one bin), shared-memory tile histogram
//                 scan    : coarseScanKernel   one CTA per bin, scan over the bin's items and its
//                                              256 tiles, tile-queue base from one atomicAdd,
//                                              active-tile list
//                 scatter : coarseScatterKernel same ballot-rank scheme at tile granularity
//
// Queues are dense CSR arrays (binQueue/binStart/binTotal, tileQueue/tileStart/tileCount): order
// inside a bin/tile is submission order BY CONSTRUCTION (chunks and items are concatenated in
// order, ranks inside a CTA come from ordered ballots), every SM participates, and the fine
// stage reads contiguous runs instead of chasing segment pointers.
#include <cuda_runtime.h>

#include "../../include/cudaraster/cuda/Overlap.cuh"

using namespace FW;

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// Block-wide exclusive scan of one int per thread (kThreads threads).  Returns the exclusive
// prefix; *total receives the block sum.  s_warp must hold kWarps + 1 ints.
__device__ __forceinline__ int blockExclusiveScan(int v, int* s_warp, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int n = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kWarps ? s_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int d = 1; d < kWarps; d <<= 1) {
            int n = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= d) wi += n;
        }
        if (lane < kWarps) s_warp[lane] = wi - w;
        if (lane == kWarps - 1) s_warp[kWarps] = wi;
    }
    __syncthreads();
    const int res = s_warp[warp] + incl - v;
    *total = s_warp[kWarps];
    __syncthreads();
    return res;
}

//------------------------------------------------------------------------------------------------
// Bin stage.
//------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kThreads) binScanKernel(const __grid_constant__ crb_frame f) {
    __shared__ int s_warp[kWarps + 1];
    __shared__ int s_base[2];
    const int bin = blockIdx.x;
    int* row = f.binCountMat + (size_t)bin * f.numChunks;
    int running = 0;
    for (int base = 0; base < f.numChunks; base += kThreads) {
        const int i = base + threadIdx.x;
        const int v = i < f.numChunks ? row[i] : 0;
        int total;
        const int ex = blockExclusiveScan(v, s_warp, &total);
        if (i < f.numChunks) row[i] = running + ex;
        running += total;
    }
    const int total = running;
    const int numItems = (total + CRB_ITEM_ENTRIES - 1) / CRB_ITEM_ENTRIES;
    if (threadIdx.x == 0) {
        const int start = atomicAdd(&f.atomics->numBinEntries, total);
        const int itemBase = atomicAdd(&f.atomics->numCoarseItems, numItems);
        f.binStart[bin] = start;
        f.binTotal[bin] = total;
        f.binItemBase[bin] = itemBase;
        f.binItemCount[bin] = numItems;
        if (start + total > f.maxBinEntries) atomicOr(&f.atomics->overflow, 2);
        if (itemBase + numItems > f.maxItems) atomicOr(&f.atomics->overflow, 8);
        s_base[0] = start;
        s_base[1] = itemBase;
    }
    __syncthreads();
    const int start = s_base[0], itemBase = s_base[1];
    if (itemBase + numItems <= f.maxItems)
        for (int k = threadIdx.x; k < numItems; k += kThreads) {
            crb_item it;
            it.bin = bin;
            it.first = start + k * CRB_ITEM_ENTRIES;
            it.count = min(CRB_ITEM_ENTRIES, total - k * CRB_ITEM_ENTRIES);
            it.slot = k;
            f.items[itemBase + k] = it;
        }
}

// Shared scatter machinery: a batch of <= kThreads queue entries, one per thread, each touching
// any number of the 256 cells (bins or tiles of one bin).  Warp w records its lanes per cell in
// s_mask[w][cell]; cell-owner threads turn the masks into per-warp offsets from the running
// cursor; every thread then ranks itself with a popc of the lanes below it.  Stable because
// threads are entries in queue order and cursors advance batch by batch.
struct ScatterSmem {
    unsigned mask[kWarps][256];
    int warpOfs[kWarps][256];
    int cursor[256];
};

template <int SamplesLog2>
__global__ void __launch_bounds__(kThreads) binScatterKernel(const __grid_constant__ crb_frame f) {
    __shared__ ScatterSmem sm;
    __shared__ int s_list[kThreads * 7];
    __shared__ int s_warp[kWarps + 1];
    if (f.atomics->overflow != 0) return;

    const int chunk = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned ltMask = laneMaskLt();
    // one thread per bin: where this chunk's entries of that bin start
    if (threadIdx.x < f.numBins) sm.cursor[threadIdx.x] = f.binStart[threadIdx.x] + f.binCountMat[(size_t)threadIdx.x * f.numChunks + chunk];
    __syncthreads();

#pragma unroll 1
    for (int r = 0; r < CRB_CHUNK_TRIS / kThreads; r++) {
        const int triBase = chunk * CRB_CHUNK_TRIS + r * kThreads;
        if (triBase >= f.numTris) break;
        // expand triangles into (triangle, sub-triangle) entries, submission order
        const int tri = triBase + threadIdx.x;
        const int n = tri < f.numTris ? (int)f.triSubtris[tri] : 0;
        int numEntries;
        const int pos = blockExclusiveScan(n, s_warp, &numEntries);
        if (n == 1) s_list[pos] = tri * 8 + 7;
        else
            for (int k = 0; k < n; k++) s_list[pos + k] = tri * 8 + k;
        __syncthreads();

        for (int b = 0; b < numEntries; b += kThreads) {
            for (int i = threadIdx.x; i < kWarps * 256; i += kThreads) (&sm.mask[0][0])[i] = 0;
            __syncthreads();
            const int e = b + threadIdx.x;
            int entry = -1;
            TriFootprint fp;
            fp.empty = true;
            if (e < numEntries) {
                entry = s_list[e];
                const uint4 h = __ldg(&f.triHeader[resolveDataIdx(entry, f.triHeader)]);
                fp = triFootprint<SamplesLog2>(h.x, h.y, h.z, f);
                forEachCell<SamplesLog2, CR_BIN_LOG2 + CR_TILE_LOG2>(fp, 0, 0, f.widthBins - 1, f.heightBins - 1,
                                                                      [&](S32 bx, S32 by) { atomicOr(&sm.mask[warp][bx + by * f.widthBins], 1u << lane); });
            }
            __syncthreads();
            if (threadIdx.x < f.numBins) {
                int run = sm.cursor[threadIdx.x];
#pragma unroll
                for (int w = 0; w < kWarps; w++) {
                    sm.warpOfs[w][threadIdx.x] = run;
                    run += __popc(sm.mask[w][threadIdx.x]);
                }
                sm.cursor[threadIdx.x] = run;
            }
            __syncthreads();
            if (entry >= 0)
                forEachCell<SamplesLog2, CR_BIN_LOG2 + CR_TILE_LOG2>(fp, 0, 0, f.widthBins - 1, f.heightBins - 1, [&](S32 bx, S32 by) {
                    const int bin = bx + by * f.widthBins;
                    f.binQueue[sm.warpOfs[warp][bin] + __popc(sm.mask[warp][bin] & ltMask)] = entry;
                });
            __syncthreads();
        }
    }
}

//------------------------------------------------------------------------------------------------
// Coarse stage.
//------------------------------------------------------------------------------------------------

struct BinWindow { int tx0, ty0, tx1, ty1; };
__device__ __forceinline__ BinWindow binWindow(const crb_frame& f, int bin) {
    const int by = bin / f.widthBins, bx = bin - by * f.widthBins;
    BinWindow w;
    w.tx0 = bx << CR_BIN_LOG2; w.ty0 = by << CR_BIN_LOG2;
    w.tx1 = min(w.tx0 + CR_BIN_SIZE - 1, f.widthTiles - 1);
    w.ty1 = min(w.ty0 + CR_BIN_SIZE - 1, f.heightTiles - 1);
    return w;
}

template <int SamplesLog2>
__global__ void __launch_bounds__(kThreads) coarseCountKernel(const __grid_constant__ crb_frame f) {
    __shared__ int s_count[CR_BIN_SQR];
    if (f.atomics->overflow != 0) return;
    const int numItems = f.atomics->numCoarseItems;
    for (int item = blockIdx.x; item < numItems; item += gridDim.x) {
        s_count[threadIdx.x] = 0;
        __syncthreads();
        const crb_item it = f.items[item];
        const BinWindow w = binWindow(f, it.bin);
        for (int e = threadIdx.x; e < it.count; e += kThreads) {
            const S32 entry = __ldg(&f.binQueue[it.first + e]);
            const uint4 h = __ldg(&f.triHeader[resolveDataIdx(entry, f.triHeader)]);
            const TriFootprint fp = triFootprint<SamplesLog2>(h.x, h.y, h.z, f);
            forEachCell<SamplesLog2, CR_TILE_LOG2>(fp, w.tx0, w.ty0, w.tx1, w.ty1,
                                                   [&](S32 tx, S32 ty) { atomicAdd(&s_count[(tx - w.tx0) + ((ty - w.ty0) << CR_BIN_LOG2)], 1); });
        }
        __syncthreads();
        f.tileCountMat[(size_t)item * CR_BIN_SQR + threadIdx.x] = s_count[threadIdx.x];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads) coarseScanKernel(const __grid_constant__ crb_frame f) {
    __shared__ int s_warp[kWarps + 1];
    __shared__ int s_base[2];
    if (f.atomics->overflow != 0) return;
    const int bin = blockIdx.x;
    const int itemBase = f.binItemBase[bin], numItems = f.binItemCount[bin];
    // thread t owns tile t of the bin: exclusive prefix over the bin's items (coalesced rows)
    int run = 0;
    for (int k = 0; k < numItems; k++) {
        int* p = &f.tileCountMat[(size_t)(itemBase + k) * CR_BIN_SQR + threadIdx.x];
        const int c = *p;
        *p = run;
        run += c;
    }
    int binSum;
    const int ofs = blockExclusiveScan(run, s_warp, &binSum);

    const BinWindow w = binWindow(f, bin);
    const int tx = w.tx0 + (threadIdx.x & (CR_BIN_SIZE - 1)), ty = w.ty0 + (threadIdx.x >> CR_BIN_LOG2);
    const bool inside = tx <= w.tx1 && ty <= w.ty1;
    const bool active = inside && (run > 0 || f.deferredClear != 0);
    int numActive;
    const int activeOfs = blockExclusiveScan(active ? 1 : 0, s_warp, &numActive);
    if (threadIdx.x == 0) {
        const int base = atomicAdd(&f.atomics->numTileEntries, binSum);
        if (base + binSum > f.maxTileEntries) atomicOr(&f.atomics->overflow, 4);
        s_base[0] = base;
        s_base[1] = atomicAdd(&f.atomics->numActiveTiles, numActive);
    }
    __syncthreads();
    if (inside) {
        const int g = tx + ty * f.widthTiles;
        f.tileStart[g] = s_base[0] + ofs;
        f.tileCount[g] = run;
        if (active) f.activeTiles[s_base[1] + activeOfs] = g;
    }
}

template <int SamplesLog2>
__global__ void __launch_bounds__(kThreads) coarseScatterKernel(const __grid_constant__ crb_frame f) {
    __shared__ ScatterSmem sm;
    if (f.atomics->overflow != 0) return;
    const int numItems = f.atomics->numCoarseItems;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned ltMask = laneMaskLt();
    for (int item = blockIdx.x; item < numItems; item += gridDim.x) {
        const crb_item it = f.items[item];
        const BinWindow w = binWindow(f, it.bin);
        {
            const int tx = w.tx0 + (threadIdx.x & (CR_BIN_SIZE - 1)), ty = w.ty0 + (threadIdx.x >> CR_BIN_LOG2);
            int cur = 0;
            if (tx <= w.tx1 && ty <= w.ty1) cur = f.tileStart[tx + ty * f.widthTiles] + f.tileCountMat[(size_t)item * CR_BIN_SQR + threadIdx.x];
            sm.cursor[threadIdx.x] = cur;
        }
        __syncthreads();
        for (int b = 0; b < it.count; b += kThreads) {
            for (int i = threadIdx.x; i < kWarps * 256; i += kThreads) (&sm.mask[0][0])[i] = 0;
            __syncthreads();
            const int e = b + threadIdx.x;
            int entry = -1;
            TriFootprint fp;
            fp.empty = true;
            if (e < it.count) {
                entry = __ldg(&f.binQueue[it.first + e]);
                const uint4 h = __ldg(&f.triHeader[resolveDataIdx(entry, f.triHeader)]);
                fp = triFootprint<SamplesLog2>(h.x, h.y, h.z, f);
                forEachCell<SamplesLog2, CR_TILE_LOG2>(fp, w.tx0, w.ty0, w.tx1, w.ty1,
                                                       [&](S32 tx, S32 ty) { atomicOr(&sm.mask[warp][(tx - w.tx0) + ((ty - w.ty0) << CR_BIN_LOG2)], 1u << lane); });
            }
            __syncthreads();
            {
                int run = sm.cursor[threadIdx.x];
#pragma unroll
                for (int k = 0; k < kWarps; k++) {
                    sm.warpOfs[k][threadIdx.x] = run;
                    run += __popc(sm.mask[k][threadIdx.x]);
                }
                sm.cursor[threadIdx.x] = run;
            }
            __syncthreads();
            if (entry >= 0)
                forEachCell<SamplesLog2, CR_TILE_LOG2>(fp, w.tx0, w.ty0, w.tx1, w.ty1, [&](S32 tx, S32 ty) {
                    const int t = (tx - w.tx0) + ((ty - w.ty0) << CR_BIN_LOG2);
                    f.tileQueue[sm.warpOfs[warp][t] + __popc(sm.mask[warp][t] & ltMask)] = entry;
                });
            __syncthreads();
        }
    }
}

inline int checkLaunch() { return cudaGetLastError() == cudaSuccess ? CRB_OK : CRB_ERR_CUDA; }

}  // namespace

extern "C" int crb_launch_bin_raster(const crb_frame* f, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    binScanKernel<<<f->numBins, kThreads, 0, s>>>(*f);
    if (f->numTris > 0) {
        if (f->samplesLog2 == 0) binScatterKernel<0><<<f->numChunks, kThreads, 0, s>>>(*f);
        else binScatterKernel<1><<<f->numChunks, kThreads, 0, s>>>(*f);
    }
    return checkLaunch();
}

extern "C" int crb_launch_coarse_raster(const crb_frame* f, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = max(1, min(f->maxItems, f->numSMs * 4));
    if (f->samplesLog2 == 0) coarseCountKernel<0><<<grid, kThreads, 0, s>>>(*f);
    else coarseCountKernel<1><<<grid, kThreads, 0, s>>>(*f);
    coarseScanKernel<<<f->numBins, kThreads, 0, s>>>(*f);
    if (f->samplesLog2 == 0) coarseScatterKernel<0><<<grid, kThreads, 0, s>>>(*f);
    else coarseScatterKernel<1><<<grid, kThreads, 0, s>>>(*f);
    return checkLaunch();
}

extern "C" int crb_bin_launches(const crb_frame* f) { return f->numTris > 0 ? 2 : 1; }
extern "C" int crb_coarse_launches(const crb_frame*) { return 3; }
