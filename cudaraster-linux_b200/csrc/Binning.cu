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
//   bin stage     count   : fused into triangle setup (one row of binCountMat per chunk of
//                                              f.chunkTris consecutive triangles)
//                 scan    : binScanKernel      one CTA per bin: block scan over the chunks, bin base
//                                              from one atomicAdd, emits the coarse work items
//                 scatter : binScatterKernel   ONE WARP per chunk, no block barriers
//   coarse stage  count   : fused into binScatterKernel: when an entry receives its slot in the
//                                              bin queue, its coarse work item (CRB_ITEM_ENTRIES
//                                              consecutive entries of one bin) is known, and the
//                                              tiles it touches are counted with global reductions
//                 scan    : coarseScanKernel   one CTA per bin, scan over the bin's items and its
//                                              256 tiles, tile-queue base from one atomicAdd,
//                                              active-tile records for the fine stage
//                 scatter : coarseScatterKernel one warp per work item
//
// The scatter passes are warp-synchronous: a batch is 32 consecutive entries, one per lane; every
// lane ORs its lane bit into a warp-private shared-memory word per cell it touches, and the rank
// of an entry inside a cell is the popcount of the lower lanes in that word -- stable by
// construction, any number of cells per entry, three __syncwarp per batch and no __syncthreads.
//
// Queues are dense CSR arrays (binQueue/binStart/binTotal, tileQueue/tileStart/tileCount): order
// inside a bin/tile is submission order BY CONSTRUCTION (chunks and items are concatenated in
// order, batches and lanes inside a warp are in order), every SM participates, and the fine
// stage reads contiguous runs instead of chasing segment pointers.
#include <cuda_runtime.h>

#include <algorithm>

#include <cooperative_groups.h>

#include "../../include/cudaraster/cuda/Overlap.cuh"

namespace cg = cooperative_groups;

using namespace FW;

namespace {

constexpr int kWarps = 8;              // warps per CTA of the warp-centric kernels
constexpr int kThreads = kWarps * 32;
constexpr int kScanThreads = 256;      // binScanKernel
constexpr int kCells = 256;            // CR_MAXBINS_SQR == CR_BIN_SQR
constexpr int kMaxListEntries = 32 * 7;

// Warp-private staging of the scatter / count passes.
struct WarpCells {
    unsigned mask[kCells];   // lanes of the current batch that touch the cell
    int cursor[kCells];      // next free queue slot of the cell
};

__device__ __forceinline__ int warpExclusiveScan(int v, int lane, int* total) {
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += n;
    }
    *total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    return incl - v;
}

// Maps a cell (cx, cy) to its index in WarpCells: row-major over the whole bin grid (bin stage)
// or relative to the bin's 16x16 tile window (coarse stage).
struct CellIndexer {
    S32 loX, loY, rowShift, rowPitch;
    __device__ __forceinline__ int operator()(S32 cx, S32 cy) const { return rowShift >= 0 ? (cx - loX) + ((cy - loY) << rowShift) : cx + cy * rowPitch; }
};

// One batch of <= 32 entries (lane = entry, entry < 0 = none, fp = its footprint) scattered into
// the cells it touches.  All 32 lanes must call.  [winLo, winHi] = inclusive cell window;
// onPlaced(cx, cy, pos) is called for every (entry, cell) pair with the queue slot it received.
//
// Fast path (whole warp): every footprint is at most 2x2 cells -- four predicated slots, straight
// line code.  Otherwise the generic three-pass enumeration with edge-refined cells.
// Returns, in lane-local pieces that add up over the warp, the number of distinct cells the batch touched (every touched cell is
// counted by its lowest lane) -- used by the profiling counters only.
template <int SamplesLog2, int CellLog2, class OnPlaced>
__device__ __forceinline__ int scatterBatch(WarpCells& wc, int32_t* __restrict__ queue, S32 entry, const TriFootprint& fp, int lane, unsigned ltMask, S32 winLoX, S32 winLoY,
                                             S32 winHiX, S32 winHiY, const CellIndexer cellOf, OnPlaced onPlaced) {
    const CellRange r = cellRange<CellLog2>(fp, winLoX, winLoY, winHiX, winHiY);
    if (!__any_sync(0xFFFFFFFFu, r.refine | (r.nx > 2) | (r.ny > 2))) {
        const unsigned bit = 1u << lane;
        const bool v0 = r.nx > 0, v1 = r.nx > 1, v2 = r.ny > 1, v3 = v1 & v2;
        const int c0 = cellOf(r.x0, r.y0), c1 = cellOf(r.x0 + 1, r.y0), c2 = cellOf(r.x0, r.y0 + 1), c3 = cellOf(r.x0 + 1, r.y0 + 1);
        if (v0) atomicOr(&wc.mask[c0], bit);
        if (v1) atomicOr(&wc.mask[c1], bit);
        if (v2) atomicOr(&wc.mask[c2], bit);
        if (v3) atomicOr(&wc.mask[c3], bit);
        __syncwarp();
        unsigned m0 = 0, m1 = 0, m2 = 0, m3 = 0;
        if (v0) { m0 = wc.mask[c0]; const int pos = wc.cursor[c0] + __popc(m0 & ltMask); queue[pos] = entry; onPlaced(r.x0, r.y0, pos); }
        if (v1) { m1 = wc.mask[c1]; const int pos = wc.cursor[c1] + __popc(m1 & ltMask); queue[pos] = entry; onPlaced(r.x0 + 1, r.y0, pos); }
        if (v2) { m2 = wc.mask[c2]; const int pos = wc.cursor[c2] + __popc(m2 & ltMask); queue[pos] = entry; onPlaced(r.x0, r.y0 + 1, pos); }
        if (v3) { m3 = wc.mask[c3]; const int pos = wc.cursor[c3] + __popc(m3 & ltMask); queue[pos] = entry; onPlaced(r.x0 + 1, r.y0 + 1, pos); }
        __syncwarp();
        // the lowest lane of every touched cell advances the cursor and clears the word
        int touched = 0;
        if (v0 && (m0 & ltMask) == 0) { wc.cursor[c0] += __popc(m0); wc.mask[c0] = 0; touched++; }
        if (v1 && (m1 & ltMask) == 0) { wc.cursor[c1] += __popc(m1); wc.mask[c1] = 0; touched++; }
        if (v2 && (m2 & ltMask) == 0) { wc.cursor[c2] += __popc(m2); wc.mask[c2] = 0; touched++; }
        if (v3 && (m3 & ltMask) == 0) { wc.cursor[c3] += __popc(m3); wc.mask[c3] = 0; touched++; }
        __syncwarp();
        return touched;
    }
    int touched = 0;
    forEachCell<SamplesLog2, CellLog2>(fp, winLoX, winLoY, winHiX, winHiY, [&](S32 cx, S32 cy) { atomicOr(&wc.mask[cellOf(cx, cy)], 1u << lane); });
    __syncwarp();
    forEachCell<SamplesLog2, CellLog2>(fp, winLoX, winLoY, winHiX, winHiY, [&](S32 cx, S32 cy) {
        const int cell = cellOf(cx, cy);
        const int pos = wc.cursor[cell] + __popc(wc.mask[cell] & ltMask);
        queue[pos] = entry;
        onPlaced(cx, cy, pos);
    });
    __syncwarp();
    // the lowest lane of every touched cell advances its cursor and clears the word (a lane that
    // reads the word after the clear sees 0 and does nothing; lanes walk their cells at their own pace,
    // so the word is read and cleared with atomics -- any interleaving gives the same result)
    forEachCell<SamplesLog2, CellLog2>(fp, winLoX, winLoY, winHiX, winHiY, [&](S32 cx, S32 cy) {
        const int cell = cellOf(cx, cy);
        const unsigned m = atomicOr(&wc.mask[cell], 0u);
        if (m != 0 && (m & ltMask) == 0) {
            wc.cursor[cell] += __popc(m);
            atomicExch(&wc.mask[cell], 0u);
            touched++;
        }
    });
    __syncwarp();
    return touched;
}

// ProfilingMode_Counters for one round (= one batch of 32 entries) of the bin (Bin = true) or coarse scatter; all 32 lanes call.
// Reference counters: cuda/PrivateDefs.hpp:168-187, counted at BinRaster.inl:176, :222-264 and CoarseRaster.inl:291, :322-412, :463, :548-550.
template <int ProfMode, int CellLog2, bool Bin>
__device__ __forceinline__ void profScatterRound(const crb_frame& f, S32 entry, const TriFootprint& fp, S32 winLoX, S32 winLoY, S32 winHiX, S32 winHiY, int emits, int touched) {
    if (ProfMode != ProfilingMode_Counters) return;
    const CellRange r = cellRange<CellLog2>(fp, winLoX, winLoY, winHiX, winHiY);
    const bool valid = entry >= 0;
    const bool generic = __any_sync(0xFFFFFFFFu, r.refine | (r.nx > 2) | (r.ny > 2));   // the path scatterBatch took for this round
    const U32 numValid = __popc(__ballot_sync(0xFFFFFFFFu, valid));
    const U32 sumEmits = __reduce_add_sync(0xFFFFFFFFu, (U32)emits), sumTouched = __reduce_add_sync(0xFFFFFFFFu, (U32)touched);
    if (Bin) {
        const U32 area = __reduce_add_sync(0xFFFFFFFFu, valid ? (U32)(r.nx * r.ny) : 0u);
        const U32 single = __popc(__ballot_sync(0xFFFFFFFFu, valid && r.nx * r.ny <= 1)), slow = __popc(__ballot_sync(0xFFFFFFFFu, valid && (r.refine || r.nx > 2 || r.ny > 2)));
        if (laneId() == 0) {
            profCount<ProfMode>(f, CRB_PROF_BinTrisPerRound, numValid, 1);
            profCount<ProfMode>(f, CRB_PROF_BinTriBBArea, area, numValid);
            profCount<ProfMode>(f, CRB_PROF_BinTriSinglePath, 100ull * single, numValid);
            profCount<ProfMode>(f, CRB_PROF_BinTriFastPath, 100ull * (numValid - single - slow), numValid);
            profCount<ProfMode>(f, CRB_PROF_BinTriSlowPath, 100ull * slow, numValid);
        }
    } else if (laneId() == 0) {
        profCount<ProfMode>(f, CRB_PROF_CoarseRoundsPerBin, 1, 0);
        profCount<ProfMode>(f, CRB_PROF_CoarseTrisPerRound, numValid, 1);
        profCount<ProfMode>(f, CRB_PROF_CoarseTilesPerRound, sumTouched, 1);
        profCount<ProfMode>(f, CRB_PROF_CoarseEmitsPerRound, sumEmits, 1);
        profCount<ProfMode>(f, CRB_PROF_CoarseEmitsPerTri, sumEmits, numValid);
        profCount<ProfMode>(f, CRB_PROF_CoarseCaseA, generic ? 0 : 100, 1);
        profCount<ProfMode>(f, CRB_PROF_CoarseCaseC, generic ? 100 : 0, 1);
    }
}

template <int SamplesLog2>
__device__ __forceinline__ TriFootprint footprintOf(const crb_frame& f, S32 entry, const uint4& h) {
    TriFootprint fp;
    fp.empty = true;
    if (entry >= 0) fp = triFootprint<SamplesLog2>(h.x, h.y, h.z, f);
    return fp;
}

//------------------------------------------------------------------------------------------------
// Bin stage.
//------------------------------------------------------------------------------------------------

// A CLUSTER of kScanCluster CTAs per bin: exclusive scan of row `bin` of binCountMat[bin][chunk] over the chunks.  Every CTA of the
// cluster scans one quarter of the row (128-bit coalesced loads) and leaves its total in its own shared memory; after a cluster
// barrier each CTA reads the totals of the lower ranks straight out of THEIR shared memory (distributed shared memory,
// cluster.map_shared_rank) -- the cross-CTA merge of the per-chunk histograms costs no global round trip and no second kernel,
// and a bin is scanned by four SMs instead of one (135 bins at 1080p would otherwise leave the scan on 135 SMs' worth of
// single CTAs with a 4x longer serial chain).  Rank 0 owns the bin: queue base, work items.
constexpr int kScanCluster = 4;
template <int ProfMode>
__global__ void __launch_bounds__(kScanThreads) binScanKernel(const __grid_constant__ crb_frame f) {
    __shared__ int s_warp[33];
    __shared__ int s_base[2];
    __shared__ int s_total;   // this CTA's part of the bin total (read by the other CTAs of the cluster through DSMEM)
    cg::cluster_group cluster = cg::this_cluster();
    gridDepLaunchDependents();
    if (threadIdx.x < 33) s_warp[threadIdx.x] = 0;
    __syncthreads();
    gridDepWait();
    ProfTimer<ProfMode> tmScan;
    tmScan.start();
    const int bin = blockIdx.x / kScanCluster, rank = (int)cluster.block_rank(), lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // rank r owns the 4-aligned run of chunks [q0, q1); thread t of it the 4-aligned run [first, first + perThread): 128-bit loads, coalesced
    const int perRank = (((f.numChunks + kScanCluster - 1) / kScanCluster) + 3) & ~3;
    const int q0 = rank * perRank, q1 = min(q0 + perRank, f.numChunks);
    const int perThread = (((perRank + kScanThreads - 1) / kScanThreads) + 3) & ~3;
    const int first = q0 + threadIdx.x * perThread;
    int* __restrict__ row = f.binCountMat + (size_t)bin * f.matPitch;
    int sum = 0;
    for (int k = 0; k < perThread; k += 4) {
        const int c = first + k;
        if (c < q1) {   // matPitch is padded to a multiple of 4 and the padding is zero
            const int4 v = *reinterpret_cast<const int4*>(row + c);
            sum += v.x + v.y + v.z + v.w;
        }
    }
    int warpTotal;
    const int ex = warpExclusiveScan(sum, lane, &warpTotal);
    if (lane == 0) s_warp[warp] = warpTotal;
    __syncthreads();
    if (warp == 0) {
        int t;
        const int w = s_warp[lane];
        const int e = warpExclusiveScan(w, lane, &t);
        s_warp[lane] = e;
        if (lane == 0) { s_warp[32] = t; s_total = t; }
    }
    cluster.sync();   // every CTA's total is in its shared memory (also a block barrier)
    int lower = 0, total = 0;
#pragma unroll
    for (int r = 0; r < kScanCluster; r++) {
        const int t = *cluster.map_shared_rank(&s_total, r);   // DSMEM read
        if (r < rank) lower += t;
        total += t;
    }
    int run = lower + s_warp[warp] + ex;
    for (int k = 0; k < perThread; k += 4) {
        const int c = first + k;
        if (c < q1) {
            const int4 v = *reinterpret_cast<const int4*>(row + c);
            int4 o;
            o.x = run; run += v.x;
            o.y = run; run += v.y;
            o.z = run; run += v.z;
            o.w = run; run += v.w;
            // padding columns keep reading 0 in the next frame (nothing else ever writes them)
            if (c + 1 >= f.numChunks) o.y = 0;
            if (c + 2 >= f.numChunks) o.z = 0;
            if (c + 3 >= f.numChunks) o.w = 0;
            *reinterpret_cast<int4*>(row + c) = o;
        }
    }
    cluster.sync();   // nobody leaves while its shared memory may still be read
    if (rank != 0) {
        tmScan.stop(f, CRB_TIMER_BinScan);
        tmScan.stop(f, CRB_TIMER_BinTotal);
        return;
    }
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
        for (int k = threadIdx.x; k < numItems; k += kScanThreads) {
            crb_item it;
            it.bin = bin;
            it.first = start + k * CRB_ITEM_ENTRIES;
            it.count = min(CRB_ITEM_ENTRIES, total - k * CRB_ITEM_ENTRIES);
            it.slot = k;
            f.items[itemBase + k] = it;
        }
    tmScan.stop(f, CRB_TIMER_BinScan);
    tmScan.stop(f, CRB_TIMER_BinTotal);
}

// Tile-level count of one placed bin-queue entry: the entry sits at slot `pos` of bin (bx, by), i.e.
// in coarse work item binItemBase + (pos - binStart) / CRB_ITEM_ENTRIES; every tile of that bin the
// triangle touches gets +1 in the item's row of tileCountMat (fire-and-forget global reductions).
// This is the COUNT pass of the coarse stage, fused here because the footprint is already at hand.
template <int SamplesLog2>
__device__ __forceinline__ void countTilesOfPlacedEntry(const crb_frame& f, const TriFootprint& fp, S32 bx, S32 by, int pos) {
    const int bin = bx + by * f.widthBins;
    const int item = __ldg(&f.binItemBase[bin]) + ((pos - __ldg(&f.binStart[bin])) / CRB_ITEM_ENTRIES);
    if (item >= f.maxItems) return;   // overflow: flagged by the scan, the frame is redone
    int* row = f.tileCountMat + (size_t)item * CR_BIN_SQR;
    const S32 tx0 = bx << CR_BIN_LOG2, ty0 = by << CR_BIN_LOG2;
    const S32 tx1 = min(tx0 + CR_BIN_SIZE - 1, f.widthTiles - 1), ty1 = min(ty0 + CR_BIN_SIZE - 1, f.heightTiles - 1);
    forEachCell<SamplesLog2, CR_TILE_LOG2>(fp, tx0, ty0, tx1, ty1, [&](S32 tx, S32 ty) { atomicAdd(&row[(tx - tx0) + ((ty - ty0) << CR_BIN_LOG2)], 1); });
}

// One warp per chunk: expands the chunk's triangles into (triangle, sub-triangle) entries in
// submission order and scatters them into the bin queue.  Loads run one batch ahead.
template <int SamplesLog2, int ProfMode>
__global__ void __launch_bounds__(kThreads, 4) binScatterKernel(const __grid_constant__ crb_frame f) {
    __shared__ WarpCells s_cells[kWarps];
    __shared__ int s_list[kWarps][kMaxListEntries];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x * kWarps + warp;
    gridDepLaunchDependents();
    gridDepWait();
    if (chunk >= f.numChunks) return;
    ProfTimer<ProfMode> tmTotal, tm;
    tmTotal.start();
    tm.start();
    const int triBegin = chunk * f.chunkTris, triEnd = min(triBegin + f.chunkTris, f.numTris);
    // triSubtris runs two batches ahead, headers one batch ahead and only for surviving triangles
    // (on a cull-heavy scene most header slots are never written and must not be fetched)
    auto loadN = [&](int t) { return t < triEnd ? (int)f.triSubtris[t] : 0; };
    auto loadH = [&](int t, int n) { return n == 1 ? __ldg(&f.triHeader[t]) : make_uint4(0, 0, 0, 0); };
    int nCur = loadN(triBegin + lane), nNext = loadN(triBegin + 32 + lane);
    uint4 hCur = loadH(triBegin + lane, nCur);
    if (f.atomics->overflow != 0) return;
    WarpCells& wc = s_cells[warp];
    const unsigned ltMask = laneMaskLt();
    for (int b = lane; b < f.numBins; b += 32) {
        wc.cursor[b] = __ldg(&f.binStart[b]) + f.binCountMat[(size_t)b * f.matPitch + chunk];
        wc.mask[b] = 0;
    }
    __syncwarp();
    tm.stop(f, CRB_TIMER_BinReadTriHeader);   // the chunk's cursors and the first triangle records
    const CellIndexer cellOf = {0, 0, -1, f.widthBins};
    long long countClocks = 0;   // ProfilingMode_Timers: time inside the fused tile counting
    auto countTimed = [&](const TriFootprint& fp, S32 bx, S32 by, int pos) {
        if (ProfMode == ProfilingMode_Timers) {
            const long long t0 = clock64();
            countTilesOfPlacedEntry<SamplesLog2>(f, fp, bx, by, pos);
            countClocks += clock64() - t0;
        } else countTilesOfPlacedEntry<SamplesLog2>(f, fp, bx, by, pos);
    };

#pragma unroll 1
    for (int t0 = triBegin; t0 < triEnd; t0 += 32) {
        const int tri = t0 + lane;
        const int n = nCur;
        const uint4 h = hCur;
        hCur = loadH(t0 + 32 + lane, nNext);
        nCur = nNext;
        nNext = loadN(t0 + 64 + lane);
        if ((f.debugFlags & 2) == 0 && __all_sync(0xFFFFFFFFu, n == 0)) continue;   // nothing survived setup in this batch (the common case of a sub-pixel soup)
        if (__all_sync(0xFFFFFFFFu, n <= 1)) {
            const S32 entry = n == 1 ? tri * 8 + 7 : -1;
            const TriFootprint fp = footprintOf<SamplesLog2>(f, entry, h);
            tm.start();
            int emits = 0;
            const int touched = scatterBatch<SamplesLog2, CR_BIN_LOG2 + CR_TILE_LOG2>(wc, f.binQueue, entry, fp, lane, ltMask, 0, 0, f.widthBins - 1, f.heightBins - 1, cellOf,
                                                                                      [&](S32 bx, S32 by, int pos) { emits++; countTimed(fp, bx, by, pos); });
            tm.stop(f, CRB_TIMER_BinRasterize);
            profScatterRound<ProfMode, CR_BIN_LOG2 + CR_TILE_LOG2, true>(f, entry, fp, 0, 0, f.widthBins - 1, f.heightBins - 1, emits, touched);
        } else {
            // clipped triangles in the batch: lay the entries out in order first
            int numEntries;
            const int pos = warpExclusiveScan(n, lane, &numEntries);
            if (n == 1) s_list[warp][pos] = tri * 8 + 7;
            else
                for (int k = 0; k < n; k++) s_list[warp][pos + k] = tri * 8 + k;
            __syncwarp();
            for (int b = 0; b < numEntries; b += 32) {
                const S32 entry = b + lane < numEntries ? s_list[warp][b + lane] : -1;
                uint4 he = make_uint4(0, 0, 0, 0);
                if (entry >= 0) he = __ldg(&f.triHeader[resolveDataIdx(entry, f.triHeader)]);
                const TriFootprint fp = footprintOf<SamplesLog2>(f, entry, he);
                tm.start();
                int emits = 0;
                const int touched = scatterBatch<SamplesLog2, CR_BIN_LOG2 + CR_TILE_LOG2>(wc, f.binQueue, entry, fp, lane, ltMask, 0, 0, f.widthBins - 1, f.heightBins - 1, cellOf,
                                                                                          [&](S32 bx, S32 by, int pos) { emits++; countTimed(fp, bx, by, pos); });
                tm.stop(f, CRB_TIMER_BinRasterize);
                profScatterRound<ProfMode, CR_BIN_LOG2 + CR_TILE_LOG2, true>(f, entry, fp, 0, 0, f.widthBins - 1, f.heightBins - 1, emits, touched);
            }
            __syncwarp();
        }
    }
    if (ProfMode == ProfilingMode_Timers && lane == 0 && countClocks > 0) atomicAdd(&f.profCounters[2 * CRB_PROF_NUM + CRB_TIMER_BinCountTiles], (unsigned long long)countClocks);
    tmTotal.stop(f, CRB_TIMER_BinTotal);
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

// Exclusive scan of one int per thread over the 256 threads of group 0 of a CTA (other groups
// pass 0 and only take part in the barriers).  *total = sum.  s_warp holds kWarps + 1 ints.
__device__ __forceinline__ int blockExclusiveScan256g(int v, int group, int* s_warp, int* total) {
    const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & (kWarps - 1);
    int wt;
    const int ex = warpExclusiveScan(v, lane, &wt);
    if (group == 0 && lane == 0) s_warp[warp] = wt;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t;
        const int w = lane < kWarps ? s_warp[lane] : 0;
        const int e = warpExclusiveScan(w, lane, &t);
        if (lane < kWarps) s_warp[lane] = e;
        if (lane == 0) s_warp[kWarps] = t;
    }
    __syncthreads();
    const int res = s_warp[warp] + ex;
    *total = s_warp[kWarps];
    __syncthreads();
    return res;
}

// One CTA per bin.  Thread (g, t) owns tile t of the bin and the g-th quarter of the bin's work
// items: the four quarters are prefix-summed concurrently and stitched through shared memory, so the
// serial chain over the items of a crowded bin is four times shorter.
constexpr int kScanGroups = 4;
template <int ProfMode>
__global__ void __launch_bounds__(kThreads * kScanGroups) coarseScanKernel(const __grid_constant__ crb_frame f) {
    __shared__ int s_warp[kWarps + 1];
    __shared__ int s_base[2];
    __shared__ int s_group[kScanGroups][kCells];
    __shared__ int s_abort;
    gridDepLaunchDependents();
    gridDepWait();
    // one thread decides for the block: other blocks of THIS grid may raise the flag at any time, and a block
    // whose threads disagree would hang in the barriers below
    if (threadIdx.x == 0) s_abort = f.atomics->overflow;
    __syncthreads();
    if (s_abort != 0) return;
    ProfTimer<ProfMode> tmScan;
    tmScan.start();
    const int bin = blockIdx.x, t = threadIdx.x & (kCells - 1), g = threadIdx.x >> 8;
    const int itemBase = f.binItemBase[bin], numItems = f.binItemCount[bin];
    if (threadIdx.x == 0 && numItems > 0) {   // reference: CoarseRaster.inl:158-159 (bins a block picked; denominator of rounds per bin)
        profCount<ProfMode>(f, CRB_PROF_CoarseBins, 1, 0);
        profCount<ProfMode>(f, CRB_PROF_CoarseRoundsPerBin, 0, 1);
    }
    const int perGroup = (numItems + kScanGroups - 1) / kScanGroups;
    const int k0 = min(g * perGroup, numItems), k1 = min(k0 + perGroup, numItems);
    int* const col = &f.tileCountMat[(size_t)itemBase * CR_BIN_SQR + t];
    int sum = 0;
    for (int k = k0; k < k1; k++) sum += col[(size_t)k * CR_BIN_SQR];
    s_group[g][t] = sum;
    __syncthreads();
    int run = 0, tileTotal = 0;
#pragma unroll
    for (int i = 0; i < kScanGroups; i++) {
        const int v = s_group[i][t];
        if (i < g) run += v;
        tileTotal += v;
    }
    for (int k = k0; k < k1; k++) {
        int* p = &col[(size_t)k * CR_BIN_SQR];
        const int c = *p;
        *p = run;
        run += c;
    }
    // the first group finishes the bin: tile offsets inside the bin's block of the tile queue
    // (block barriers below are reached by all threads; only group 0 contributes / writes)
    int binSum;
    const int ofs = blockExclusiveScan256g(g == 0 ? tileTotal : 0, g, s_warp, &binSum);

    const BinWindow w = binWindow(f, bin);
    const int tx = w.tx0 + (t & (CR_BIN_SIZE - 1)), ty = w.ty0 + (t >> CR_BIN_LOG2);
    const bool inside = tx <= w.tx1 && ty <= w.ty1;
    const bool active = inside && (tileTotal > 0 || f.deferredClear != 0);
    int numActive;
    const int activeOfs = blockExclusiveScan256g((g == 0 && active) ? 1 : 0, g, s_warp, &numActive);
    if (threadIdx.x == 0) {
        const int base = atomicAdd(&f.atomics->numTileEntries, binSum);
        if (base + binSum > f.maxTileEntries) atomicOr(&f.atomics->overflow, 4);
        s_base[0] = base;
        s_base[1] = atomicAdd(&f.atomics->numActiveTiles, numActive);
    }
    __syncthreads();
    if (g == 0 && inside) {
        const int gi = tx + ty * f.widthTiles;
        f.tileStart[gi] = s_base[0] + ofs;
        f.tileCount[gi] = tileTotal;
        if (active) {
            f.activeTiles[s_base[1] + activeOfs] = gi;
            f.activeRecs[s_base[1] + activeOfs] = make_int4(gi, s_base[0] + ofs, tileTotal, 0);
        }
    }
    tmScan.stop(f, CRB_TIMER_CoarseScan);
    tmScan.stop(f, CRB_TIMER_CoarseTotal);
}

// One warp per work item.  Entries are loaded two batches ahead, headers one batch ahead.
template <int SamplesLog2, int ProfMode>
__global__ void __launch_bounds__(kThreads, 4) coarseScatterKernel(const __grid_constant__ crb_frame f) {
    __shared__ WarpCells s_cells[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = blockIdx.x * kWarps + warp;
    gridDepLaunchDependents();
    gridDepWait();
    if (f.atomics->overflow != 0 || item >= f.atomics->numCoarseItems) return;
    ProfTimer<ProfMode> tmTotal, tm;
    tmTotal.start();
    tm.start();
    WarpCells& wc = s_cells[warp];
    const unsigned ltMask = laneMaskLt();
    const crb_item it = f.items[item];
    const int32_t* __restrict__ src = f.binQueue + it.first;
    S32 entryCur = lane < it.count ? __ldg(&src[lane]) : -1;
    S32 entryNext = 32 + lane < it.count ? __ldg(&src[32 + lane]) : -1;
    uint4 hCur = make_uint4(0, 0, 0, 0);
    if (entryCur >= 0) hCur = __ldg(&f.triHeader[resolveDataIdx(entryCur, f.triHeader)]);

    const BinWindow w = binWindow(f, it.bin);
#pragma unroll
    for (int i = 0; i < kCells / 32; i++) {
        const int t = lane + 32 * i;
        const int tx = w.tx0 + (t & (CR_BIN_SIZE - 1)), ty = w.ty0 + (t >> CR_BIN_LOG2);
        int cur = 0;
        if (tx <= w.tx1 && ty <= w.ty1) cur = f.tileStart[tx + ty * f.widthTiles] + f.tileCountMat[(size_t)item * CR_BIN_SQR + t];
        wc.cursor[t] = cur;
        wc.mask[t] = 0;
    }
    __syncwarp();
    tm.stop(f, CRB_TIMER_CoarseStreamRead);   // the item, its cursors, the first entries and headers
    const CellIndexer cellOf = {w.tx0, w.ty0, CR_BIN_LOG2, 0};
#pragma unroll 1
    for (int b = 0; b < it.count; b += 32) {
        tm.start();
        uint4 hNext = make_uint4(0, 0, 0, 0);
        if (entryNext >= 0) hNext = __ldg(&f.triHeader[resolveDataIdx(entryNext, f.triHeader)]);
        const S32 entryNext2 = b + 64 + lane < it.count ? __ldg(&src[b + 64 + lane]) : -1;
        const TriFootprint fp = footprintOf<SamplesLog2>(f, entryCur, hCur);
        tm.stop(f, CRB_TIMER_CoarseStreamRead);
        tm.start();
        int emits = 0;
        const int touched = scatterBatch<SamplesLog2, CR_TILE_LOG2>(wc, f.tileQueue, entryCur, fp, lane, ltMask, w.tx0, w.ty0, w.tx1, w.ty1, cellOf, [&](S32, S32, int) { emits++; });
        tm.stop(f, CRB_TIMER_CoarseRasterize);
        profScatterRound<ProfMode, CR_TILE_LOG2, false>(f, entryCur, fp, w.tx0, w.ty0, w.tx1, w.ty1, emits, touched);
        entryCur = entryNext; hCur = hNext; entryNext = entryNext2;
    }
    tmTotal.stop(f, CRB_TIMER_CoarseTotal);
}

//------------------------------------------------------------------------------------------------
// Direct tile path (crb_frame::directMode, crb_set_binning_mode in crb200.h).
//
// For frames of small triangles and an order-independent pipe the two-level stable sort is more than the
// frame needs: the fine raster keeps, per sample, the (depth, submission index) minimum, which does not depend
// on the order the fragments arrive in.  So the tile queues may be filled in ANY order: triangle setup has
// already counted every tile's entries (tileCounter), directAllocKernel turns the counts into queue extents
// (block scan + one atomicAdd per 256 tiles -- the extents of different blocks need no particular order), zeroes
// the counters again for the next frame (no memset) and leaves, per tile, a cursor at the END of its extent;
// directScatterKernel gives every (triangle, tile) pair a slot with atomicSub on that cursor.
//------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kThreads) directAllocKernel(const __grid_constant__ crb_frame f) {
    __shared__ int s_warp[kWarps + 1];
    __shared__ int s_base[2];
    __shared__ int s_abort;
    gridDepLaunchDependents();
    gridDepWait();
    // one thread decides for the block: other blocks of THIS grid may raise the flag at any time, and a block
    // whose threads disagree would hang in the barriers below
    // micro mode with nothing queued (every triangle went the micro way): no queue extents to make -- the fine raster
    // addresses the tiles directly and takes every queue as empty (the counter is final: setup has ended)
    if (threadIdx.x == 0) s_abort = f.atomics->overflow | ((f.microMode != 0 && f.atomics->numQueuedCtas == 0) ? 1 : 0);
    __syncthreads();
    if (s_abort != 0) return;
    // The LARGE sub-triangles setup listed (final: setup has ended) are counted here, one whole CTA per triangle at a time and
    // every CTA of the grid taking its share, followed by a grid barrier: the scan below needs the complete counts.  The grid
    // is small (one CTA per 256 tiles, <= 256 CTAs of 256 threads) and therefore co-resident, which is what lets a CTA wait
    // for the others; frames without large triangles skip both.
    const int numLarge = min(f.atomics->numLargeTris, f.maxLarge);
    if (numLarge > 0) {
        for (int k = blockIdx.x; k < numLarge; k += gridDim.x) {
            const uint4 h = __ldg(&f.triHeader[f.largeList[k].y]);
            const TriFootprint fp = f.samplesLog2 == 0 ? triFootprint<0>(h.x, h.y, h.z, f) : triFootprint<1>(h.x, h.y, h.z, f);
            auto count = [&](S32 tx, S32 ty) { atomicAdd(&f.tileCounter[tx + ty * f.widthTiles], 1); };
            auto countSpan = [&](S32 ty, S32 xa, S32 xb) { for (S32 tx = xa; tx <= xb; tx++) atomicAdd(&f.tileCounter[tx + ty * f.widthTiles], 1); };
            if (f.samplesLog2 == 0) forEachCellCoop<0, CR_TILE_LOG2>(fp, 0, 0, f.widthTiles - 1, f.heightTiles - 1, threadIdx.x, kThreads, count, countSpan);
            else forEachCellCoop<1, CR_TILE_LOG2>(fp, 0, 0, f.widthTiles - 1, f.heightTiles - 1, threadIdx.x, kThreads, count, countSpan);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(&f.atomics->allocBarrier, 1);
            while (*(volatile int*)&f.atomics->allocBarrier < (int)gridDim.x) __nanosleep(64);
            __threadfence();
        }
        __syncthreads();
    }
    const int t = blockIdx.x * kThreads + threadIdx.x;
    const bool inside = t < f.numTiles;
    const int cnt = inside ? *(volatile int*)&f.tileCounter[t] : 0;   // (volatile: other CTAs' reductions landed behind the barrier above)
    if (cnt != 0) f.tileCounter[t] = 0;   // clean for the next frame's setup
    // micro mode: every tile is active (a tile without queue entries may hold fragments in the visibility buffer) and
    // the fine raster addresses tiles directly, numActiveTiles == numTiles
    const bool active = inside && (cnt > 0 || f.deferredClear != 0 || f.microMode != 0);
    int blockSum, numActive;
    const int ofs = blockExclusiveScan256g(cnt, 0, s_warp, &blockSum);
    const int activeOfs = blockExclusiveScan256g(active ? 1 : 0, 0, s_warp, &numActive);
    if (threadIdx.x == 0) {
        const int base = atomicAdd(&f.atomics->numTileEntries, blockSum);
        if (base + blockSum > f.maxTileEntries) atomicOr(&f.atomics->overflow, 4);
        s_base[0] = base;
        s_base[1] = atomicAdd(&f.atomics->numActiveTiles, numActive);
    }
    __syncthreads();
    if (inside) {
        f.tileStart[t] = s_base[0] + ofs;
        f.tileCount[t] = cnt;
        f.tileCursor[t] = s_base[0] + ofs + cnt;   // the scatter pass counts it down to tileStart
        if (active) {
            f.activeTiles[s_base[1] + activeOfs] = t;
            f.activeRecs[s_base[1] + activeOfs] = make_int4(t, s_base[0] + ofs, cnt, 0);
        }
    }
}

// Every thread places FOUR consecutive input triangles.  Setup left one word per triangle (crb_frame::triTileCode):
// the common case -- a single sub-triangle on at most 2x2 tiles -- needs nothing else, and its (at most four) slots
// are taken with up to sixteen independent atomics issued back to back before any result is used.  The cursor a slot comes from already holds the absolute queue position (directAllocKernel), so a
// placed entry costs two scattered memory operations: the atomic and the store -- the kernel is bound by the rate at
// which an SM issues scattered accesses, not by latency.  (Measured and rejected: warp-aggregated atomics with
// __match_any_sync -- 25 vs 19 us on C2, 109 vs 52 us on C4: the match costs more than the atomics it saves.)
// Clipped and refined triangles go through triSubtris / the headers; sub-triangles that span more than CRB_DIRECT_MAX_TILES tiles
// on an axis are on triangle setup's global list and are scattered by whole CTAs at the end of the kernel (same split as the
// count pass).
#ifndef CRB_SCATTER_MIN_BLOCKS
#define CRB_SCATTER_MIN_BLOCKS 6
#endif
constexpr int kScatterTris = 4;

// The uncommon triangle of the scatter pass (kept out of line: its S64 edge tests must not cost the common path registers).
template <int SamplesLog2>
static __device__ __noinline__ void scatterGeneralTriangle(const crb_frame& f, int tri) {
    auto place = [&](S32 tile, S32 entry) {
        f.tileQueue[atomicSub(&f.tileCursor[tile], 1) - 1] = entry;
    };
    const int n = (int)f.triSubtris[tri];
    const uint4 h = __ldg(&f.triHeader[tri]);
    for (int sub = 0; sub < n; sub++) {
        const S32 entry = n == 1 ? tri * 8 + 7 : tri * 8 + sub;
        const S32 slot = n == 1 ? tri : (S32)h.w + sub;
        const uint4 hs = n == 1 ? h : __ldg(&f.triHeader[slot]);
        const TriFootprint fp = triFootprint<SamplesLog2>(hs.x, hs.y, hs.z, f);
        const CellRange r = cellRange<CR_TILE_LOG2>(fp, 0, 0, f.widthTiles - 1, f.heightTiles - 1);
        if ((r.nx > CRB_DIRECT_MAX_TILES) | (r.ny > CRB_DIRECT_MAX_TILES)) continue;   // on the large list (same rule as triangle setup)
        forEachCell<SamplesLog2, CR_TILE_LOG2>(fp, 0, 0, f.widthTiles - 1, f.heightTiles - 1, [&](S32 tx, S32 ty) { place(tx + ty * f.widthTiles, entry); });
    }
}

template <int SamplesLog2>
__global__ void __launch_bounds__(kThreads, CRB_SCATTER_MIN_BLOCKS) directScatterKernel(const __grid_constant__ crb_frame f) {
    gridDepLaunchDependents();
    gridDepWait();
    if (f.atomics->numTileEntries == 0) return;   // nothing was queued (every triangle went the micro way): final since directAllocKernel ended
    const int base = (blockIdx.x * kThreads + threadIdx.x) * kScatterTris;
    uint4 c4 = make_uint4(0, 0, 0, 0);
    // (the buffer is padded to a multiple of 4 words; batches of 32 triangles none of which was queued left no words, only their flag)
    if (base < f.numTris && f.batchQueued[base >> 5] != 0) c4 = __ldg(reinterpret_cast<const uint4*>(f.triTileCode + base));
    const bool abort = f.atomics->overflow != 0;   // a queue overflowed: the frame is redone (the host resets the counters); nothing raises the flag during this grid
    __syncthreads();
    if (abort) return;
    U32 code[kScatterTris] = {c4.x, c4.y, c4.z, c4.w};
    S32 t0[kScatterTris];
#pragma unroll
    for (int k = 0; k < kScatterTris; k++) {
        if (base + k >= f.numTris) code[k] = 0;
        t0[k] = (S32)(code[k] & 0xFF) + (S32)((code[k] >> 8) & 0xFF) * f.widthTiles;
    }
    // all (up to 16) atomics of the thread are issued before the first result is used
    int pos[4][kScatterTris];
#pragma unroll
    for (int s = 0; s < 4; s++) {   // tile (s & 1, s >> 1) of the footprint
        const U32 need = 0x80000000u | ((s & 1) ? 0x10000u : 0u) | ((s & 2) ? 0x20000u : 0u);
        const S32 ofs = (s & 1) + ((s & 2) ? f.widthTiles : 0);
#pragma unroll
        for (int k = 0; k < kScatterTris; k++)
            if ((code[k] & need) == need) pos[s][k] = atomicSub(&f.tileCursor[t0[k] + ofs], 1) - 1;
    }
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const U32 need = 0x80000000u | ((s & 1) ? 0x10000u : 0u) | ((s & 2) ? 0x20000u : 0u);
#pragma unroll
        for (int k = 0; k < kScatterTris; k++)
            if ((code[k] & need) == need) f.tileQueue[pos[s][k]] = (base + k) * 8 + 7;
    }

#pragma unroll 1
    for (int k = 0; k < kScatterTris; k++)   // clipped, refined or large: through the headers
        if (code[k] == CRB_TILECODE_GENERAL) scatterGeneralTriangle<SamplesLog2>(f, base + k);
    auto place = [&](S32 tile, S32 entry) {
        f.tileQueue[atomicSub(&f.tileCursor[tile], 1) - 1] = entry;
    };
    // the large sub-triangles of the frame (triangle setup's global list), one whole CTA per triangle at a time, all CTAs sharing the list
    const int numLarge = min(f.atomics->numLargeTris, f.maxLarge);
    for (int k = blockIdx.x; k < numLarge; k += gridDim.x) {
        const int2 e = f.largeList[k];
        const uint4 hs = __ldg(&f.triHeader[e.y]);
        const TriFootprint fp = triFootprint<SamplesLog2>(hs.x, hs.y, hs.z, f);
        // a span of a row: eight slots are taken (eight independent atomics in flight) before the first entry is stored
        auto placeSpan = [&](S32 ty, S32 xa, S32 xb) {
            int* const cur = f.tileCursor + ty * f.widthTiles;
            for (S32 x0 = xa; x0 <= xb; x0 += 8) {
                int pos[8];
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (x0 + j <= xb) pos[j] = atomicSub(cur + x0 + j, 1) - 1;
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (x0 + j <= xb) f.tileQueue[pos[j]] = e.x;
            }
        };
        forEachCellCoop<SamplesLog2, CR_TILE_LOG2>(fp, 0, 0, f.widthTiles - 1, f.heightTiles - 1, threadIdx.x, kThreads, [&](S32 tx, S32 ty) { place(tx + ty * f.widthTiles, e.x); }, placeSpan);
    }
}

inline int checkLaunch() { return cudaGetLastError() == cudaSuccess ? CRB_OK : CRB_ERR_CUDA; }

}  // namespace

namespace {
template <int ProfMode>
int launchBinRaster(const crb_frame* f, cudaStream_t s) {
    cudaError_t e = launchChained(binScanKernel<ProfMode>, f->numBins * kScanCluster, kScanThreads, s, *f, kScanCluster);   // a cluster of CTAs per bin
    if (e == cudaSuccess && f->numTris > 0) {
        const int grid = (f->numChunks + kWarps - 1) / kWarps;
        e = f->samplesLog2 == 0 ? launchChained(binScatterKernel<0, ProfMode>, grid, kThreads, s, *f) : launchChained(binScatterKernel<1, ProfMode>, grid, kThreads, s, *f);
    }
    return e == cudaSuccess ? checkLaunch() : CRB_ERR_CUDA;
}
template <int ProfMode>
int launchCoarseRaster(const crb_frame* f, cudaStream_t s) {
    const int grid = max(1, (f->maxItems + kWarps - 1) / kWarps);
    cudaError_t e = launchChained(coarseScanKernel<ProfMode>, f->numBins, kThreads * kScanGroups, s, *f);
    if (e == cudaSuccess) e = f->samplesLog2 == 0 ? launchChained(coarseScatterKernel<0, ProfMode>, grid, kThreads, s, *f) : launchChained(coarseScatterKernel<1, ProfMode>, grid, kThreads, s, *f);
    return e == cudaSuccess ? checkLaunch() : CRB_ERR_CUDA;
}
}  // namespace

// The pipe's profiling mode (crb_frame::profilingMode, CR_PROFILING_MODE of the pipe's translation unit) picks the instrumented
// instances: the default instances carry no profiling code.
extern "C" int crb_launch_bin_raster(const crb_frame* f, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    return f->profilingMode == ProfilingMode_Counters ? launchBinRaster<ProfilingMode_Counters>(f, s)
         : f->profilingMode == ProfilingMode_Timers ? launchBinRaster<ProfilingMode_Timers>(f, s) : launchBinRaster<ProfilingMode_Default>(f, s);
}

extern "C" int crb_launch_coarse_raster(const crb_frame* f, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    return f->profilingMode == ProfilingMode_Counters ? launchCoarseRaster<ProfilingMode_Counters>(f, s)
         : f->profilingMode == ProfilingMode_Timers ? launchCoarseRaster<ProfilingMode_Timers>(f, s) : launchCoarseRaster<ProfilingMode_Default>(f, s);
}

// Direct tile path: the two kernels that stand where the bin and coarse stages stand.
extern "C" int crb_launch_direct_alloc(const crb_frame* f, void* stream) {
    const cudaError_t e = launchChained(directAllocKernel, (f->numTiles + kThreads - 1) / kThreads, kThreads, (cudaStream_t)stream, *f);
    return e == cudaSuccess ? checkLaunch() : CRB_ERR_CUDA;
}

extern "C" int crb_launch_direct_scatter(const crb_frame* f, void* stream) {
    if (f->numTris <= 0) return CRB_OK;
    // one thread per 4 triangles -- but never fewer CTAs than it takes to fill the GPU: the CTAs share out the list of large
    // sub-triangles, however few input triangles the frame has
    const int grid = std::max((f->numTris + kThreads * kScatterTris - 1) / (kThreads * kScatterTris), 2 * std::max(f->numSMs, 1));
    const cudaError_t e = f->samplesLog2 == 0 ? launchChained(directScatterKernel<0>, grid, kThreads, (cudaStream_t)stream, *f)
                                              : launchChained(directScatterKernel<1>, grid, kThreads, (cudaStream_t)stream, *f);
    return e == cudaSuccess ? checkLaunch() : CRB_ERR_CUDA;
}

extern "C" int crb_bin_launches(const crb_frame* f) { return f->directMode ? 1 : (f->numTris > 0 ? 2 : 1); }
extern "C" int crb_coarse_launches(const crb_frame* f) { return f->directMode ? (f->numTris > 0 ? 1 : 0) : 2; }
