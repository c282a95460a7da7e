// Multi-GPU host layer for the kept C++ API (new design; the reference is single-GPU): one process per GPU,
// work partitioned only where it shards naturally --
//   sort-first     a large frame is cut into rectangles (FW::splitFrame == crb_split_frame); rank r renders rectangles
//                  r, r + world, ... STRAIGHT into the full frame that lives in the display GPU's memory (CUDA IPC peer memory
//                  over NVLink / NVSwitch + crb_set_color_pitch): geometry replicated, no gather, no paste;
//   view-parallel  independent views go round robin over the ranks, every rank delivers its frames into slots of the display GPU.
// The composite is therefore the render itself (or one DMA copy); the only thing processes exchange is the 64-byte IPC handle
// of the display rank's allocation, over whatever transport the application has (MPI_Bcast, a pipe, a file): PeerFrames does
// not care.  Header only, on top of the C ABI (include/crb200.h) like CudaRaster.hpp.
#pragma once
#include <cstring>
#include <vector>

#include "CudaRaster.hpp"

namespace FW {

struct FrameRect { int x0, y0, w, h; };

// At least `parts` rectangles, none straddling a 2048-px parent cell (crb_split_frame).
inline std::vector<FrameRect> splitFrame(int fullWidth, int fullHeight, int parts) {
    const int n = crb_split_frame(fullWidth, fullHeight, parts, NULL, 0);
    if (n < 0) fail("splitFrame: bad arguments!");
    std::vector<int> raw((size_t)4 * n);
    crb_split_frame(fullWidth, fullHeight, parts, raw.data(), n);
    std::vector<FrameRect> out((size_t)n);
    for (int i = 0; i < n; i++) out[i] = FrameRect{raw[4 * i], raw[4 * i + 1], raw[4 * i + 2], raw[4 * i + 3]};
    return out;
}

// Frame slots in the display rank's memory, mapped by every rank.  slot k (k < numSlots) is `slotBytes` bytes; behind the slots
// sits one 32-bit mark per (slot, rank) that a rank sets, stream ordered, after it has finished its part of the slot's frame.
class PeerFrames {
public:
    PeerFrames() : m_base(NULL), m_owner(false), m_slotBytes(0), m_numSlots(0), m_world(0) {}
    ~PeerFrames() { close(); }

    // display rank: allocates; `handle` (CRB_IPC_HANDLE_BYTES bytes) goes to the other ranks
    void create(size_t slotBytes, int numSlots, int world, unsigned char handle[CRB_IPC_HANDLE_BYTES]) {
        layout(slotBytes, numSlots, world);
        if (crb_ipc_alloc(totalBytes(), &m_base, handle) != CRB_OK) fail("PeerFrames: allocation failed!");
        m_owner = true;
    }
    // every other rank: maps the display rank's allocation
    void open(size_t slotBytes, int numSlots, int world, const unsigned char handle[CRB_IPC_HANDLE_BYTES]) {
        layout(slotBytes, numSlots, world);
        if (crb_ipc_open(handle, &m_base) != CRB_OK) fail("PeerFrames: CUDA IPC / peer access is not available between these GPUs!");
        m_owner = false;
    }
    void close(void) {
        if (!m_base) return;
        if (m_owner) crb_ipc_free(m_base); else crb_ipc_close(m_base);
        m_base = NULL;
    }
    U8* slot(int k) const { return (U8*)m_base + (size_t)k * m_slotBytes; }
    U32* mark(int k, int rank) const { return (U32*)((U8*)m_base + (size_t)m_numSlots * m_slotBytes) + (size_t)k * m_world + rank; }
    void publish(int k, int rank, U32 value, cudaStream_t stream = NULL) const {
        if (crb_ipc_signal(mark(k, rank), value, stream) != CRB_OK) fail("PeerFrames: frame mark failed!");
    }
    size_t totalBytes(void) const { return (size_t)m_numSlots * m_slotBytes + (size_t)m_numSlots * m_world * sizeof(U32); }

private:
    void layout(size_t slotBytes, int numSlots, int world) { close(); m_slotBytes = (slotBytes + 255) & ~(size_t)255; m_numSlots = numSlots; m_world = world; }
    void* m_base;
    bool m_owner;
    size_t m_slotBytes;
    int m_numSlots, m_world;
};

// Sort-first rendering of one frame: this rank's rectangles, rendered in place into `frame` (a fullWidth x fullHeight RGBA8 image,
// row pitch fullWidth texels: a PeerFrames slot on the display GPU, or local memory).  The caller has set the pixel pipe, the vertex
// and the index buffer; `depths[i]` is the depth surface of this rank's i-th rectangle.  Optionally a clip-space bounds table
// (crb_compute_chunk_bounds) lets the rank skip the chunks of the mesh outside each rectangle.
class SortFirstRenderer {
public:
    SortFirstRenderer(int fullWidth, int fullHeight, int rank, int world) : m_fw(fullWidth), m_fh(fullHeight) {
        const std::vector<FrameRect> all = splitFrame(fullWidth, fullHeight, world);
        m_numRects = (int)all.size();
        for (int i = rank; i < (int)all.size(); i += world) m_mine.push_back(all[i]);
    }
    const std::vector<FrameRect>& rects(void) const { return m_mine; }
    int numRectsOfFrame(void) const { return m_numRects; }

    void render(CudaRaster& cr, void* frame, const std::vector<CudaSurface*>& depths, const Vec4f& clearColor, F32 clearDepth, const float* d_chunkBounds = NULL,
                bool asynchronous = false) {
        for (size_t i = 0; i < m_mine.size(); i++) {
            const FrameRect& r = m_mine[i];
            cr.setSurfacePointers((U32*)frame + (size_t)r.y0 * m_fw + r.x0, depths[i]->getCudaPtr(), Vec2i(r.w, r.h), 1);
            cr.setColorPitch(m_fw);
            cr.setChunkBounds(d_chunkBounds);
            cr.setSubViewport(m_fw, m_fh, r.x0, r.y0);
            cr.deferredClear(clearColor, clearDepth);
            if (asynchronous) cr.drawTrianglesAsync(); else cr.drawTriangles();
        }
        cr.setSubViewport(0, 0, 0, 0);
        cr.setColorPitch(0);
    }

private:
    int m_fw, m_fh, m_numRects;
    std::vector<FrameRect> m_mine;
};

// View-parallel assignment: the views rank `rank` renders.
inline std::vector<int> viewsOfRank(int numViews, int rank, int world) {
    std::vector<int> v;
    for (int i = rank; i < numViews; i += world) v.push_back(i);
    return v;
}

}  // namespace FW
