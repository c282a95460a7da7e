// Format constants of the pipeline.  The VALUES are the reference's contract
// (src/cudaraster/cuda/Constants.hpp:21-84): 4 subpixel bits, 8x8 px tiles, 16x16-tile bins,
// viewport <= 2048, depth range shrunk by the interpolation error bound.  The scheduling
// constants of the Fermi design (16 bin streams, 512-entry segments, 20 fine warps) are gone:
// the B200 pipeline uses count/scan/scatter queues.
#pragma once

#define CR_MAXVIEWPORT_LOG2 11
#define CR_SUBPIXEL_LOG2 4
#define CR_MAXBINS_LOG2 4
#define CR_BIN_LOG2 4
#define CR_TILE_LOG2 3
#define CR_MAXSUBTRIS_LOG2 24

#define CR_FLIPBIT_FLIP_Y 2
#define CR_FLIPBIT_FLIP_X 3
#define CR_FLIPBIT_SWAP_XY 4
#define CR_FLIPBIT_COMPL 5

#define CR_MAXVIEWPORT_SIZE (1 << CR_MAXVIEWPORT_LOG2)
#define CR_SUBPIXEL_SIZE (1 << CR_SUBPIXEL_LOG2)
#define CR_MAXBINS_SIZE (1 << CR_MAXBINS_LOG2)
#define CR_MAXBINS_SQR (1 << (CR_MAXBINS_LOG2 * 2))
#define CR_BIN_SIZE (1 << CR_BIN_LOG2)
#define CR_BIN_SQR (1 << (CR_BIN_LOG2 * 2))
#define CR_MAXTILES_LOG2 (CR_MAXBINS_LOG2 + CR_BIN_LOG2)
#define CR_MAXTILES_SIZE (1 << CR_MAXTILES_LOG2)
#define CR_MAXTILES_SQR (1 << (CR_MAXTILES_LOG2 * 2))
#define CR_TILE_SIZE (1 << CR_TILE_LOG2)
#define CR_TILE_SQR (1 << (CR_TILE_LOG2 * 2))
#define CR_MAXSUBTRIS_SIZE (1 << CR_MAXSUBTRIS_LOG2)

// Interpolating Z/W/U/V at sample positions is accurate to +-CR_LERP_ERROR ULPs, so the depth
// range is shrunk from both ends to keep U32 arithmetic from wrapping.
#define CR_LERP_ERROR(SAMPLES_LOG2) (2200u << (SAMPLES_LOG2))
#define CR_DEPTH_MIN CR_LERP_ERROR(3)
#define CR_DEPTH_MAX (0xFFFFFFFFu - CR_LERP_ERROR(3))
#define CR_BARY_MAX ((1 << (30 - CR_SUBPIXEL_LOG2)) - 1)

// ---- B200 scheduling constants (new) --------------------------------------------------------
#ifndef CRB_SETUP_THREADS
#define CRB_SETUP_THREADS 256     // threads per setup CTA (one triangle each)
#endif
#define CRB_MIN_CHUNK_TRIS 256    // a chunk (one column of the bin count matrix, one warp of the bin scatter) holds at least this many triangles
#ifndef CRB_SETUP_MIN_BLOCKS
#define CRB_SETUP_MIN_BLOCKS 5    // resident setup CTAs per SM the register allocation must allow (5 -> 48 registers, 8 B spill; measured 4/5/6: 42.5/39.6/45.0 us on C2)
#endif
#define CRB_MAX_CHUNKS 16384       // chunkTris = CRB_SETUP_THREADS * 2^k, smallest k with numChunks <= this
#define CRB_BIN_THREADS 256       // == CR_MAXBINS_SQR: one thread per bin in the scan phases
#define CRB_COARSE_THREADS 256    // == CR_BIN_SQR: one thread per tile-in-bin in the scan phases
#define CRB_ITEM_ENTRIES 256      // bin-queue entries per coarse work item (one warp, 8 batches)
#ifndef CRB_FINE_REVERSE
#define CRB_FINE_REVERSE 1        // micro-mode fine raster walks the tiles last to first (C2: fine 39.0 -> 37.4 us, next setup 49.5 -> 46.0 us)
#endif
#ifndef CRB_FINE_WARPS
#ifndef CRB_FINE_WARPS_PER_SM
#define CRB_FINE_WARPS_PER_SM 32  // resident fine warps per SM the register allocation must allow (32 -> 64 registers)
#endif
#define CRB_FINE_WARPS 2          // warps (= tiles) per fine CTA: small CTAs, so a long tile does not hold idle warps' slots (measured 1/2/4/8: 51.2/51.1/51.5/53.3 us on C2)
#endif
