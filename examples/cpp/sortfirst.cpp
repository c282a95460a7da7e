// Multi-GPU example of the kept C++ host API (include/cudaraster/MultiGpu.hpp): ONE frame of 2560 x 1440 (beyond the 2048 px
// viewport limit) rendered sort-first by `world` PROCESSES, one per GPU (rank r uses device r % deviceCount, so the example also
// runs on a single GPU), straight into the full frame in rank 0's memory -- CUDA IPC peer memory, no gather, no paste.  The only
// thing the processes exchange is the 64-byte IPC handle (here over pipes; MPI_Bcast in a real job).  Rank 0 then renders the
// whole frame alone and compares: "sortfirst: OK".
//
//     usage: sortfirst <world> [out.raw]
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cudaraster/MultiGpu.hpp>

using namespace FW;

struct Vertex { float x, y, z, w, r, g, b, a; };   // GouraudVertex

static void makeMesh(std::vector<Vertex>& v, std::vector<int>& idx, int nx, int ny) {
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.0f; };
    for (int j = 0; j <= ny; j++)
        for (int i = 0; i <= nx; i++) {
            const float x = -1.03f + 2.06f * i / nx + (rnd() - 0.5f) * 0.5f / nx, y = -1.03f + 2.06f * j / ny + (rnd() - 0.5f) * 0.5f / ny;
            const float w = 1.0f + 0.4f * y, z = 0.3f * x * y;
            v.push_back(Vertex{x * w, y * w, z * w, w, rnd(), rnd(), rnd(), 1.0f});
        }
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            const int a = j * (nx + 1) + i, b = a + 1, c = a + nx + 1, d = c + 1;
            idx.insert(idx.end(), {a, b, d, a, d, c});
        }
}

static int runRank(int rank, int world, int rfd, const std::vector<int>& wfds, const char* outPath) {
    const int fw = 2560, fh = 1440;
    int devices = 0;
    cudaGetDeviceCount(&devices);
    if (devices < 1) fail("sortfirst: no CUDA device!");
    const int device = rank % devices;
    cudaSetDevice(device);

    std::vector<Vertex> verts;
    std::vector<int> idx;
    makeMesh(verts, idx, 400, 300);
    const int numTris = (int)idx.size() / 3;
    Buffer vb(verts.data(), (S64)(verts.size() * sizeof(Vertex))), ib(idx.data(), (S64)(idx.size() * sizeof(int)));

    CudaRaster cr(device);
    cr.setPixelPipe(NULL, "PixelPipe_gouraud_s0_f3_BlendReplace");
    cr.setVertexBuffer(&vb, 0);
    cr.setIndexBuffer(&ib, 0, numTris);
    // once per mesh: clip-space bounds per chunk of 256 triangles (a rank skips the chunks outside its rectangles)
    Buffer bounds;
    bounds.resizeDiscard((S64)((numTris + 255) / 256) * 16);
    if (crb_compute_chunk_bounds(vb.getCudaPtr(), (int)sizeof(Vertex), (const int32_t*)ib.getCudaPtr(), numTris, (float*)bounds.getCudaPtr(), NULL) != CRB_OK) fail("sortfirst: bounds failed!");

    // the full frame lives in rank 0's memory; everybody maps it
    PeerFrames frames;
    unsigned char handle[CRB_IPC_HANDLE_BYTES];
    const size_t frameBytes = (size_t)fw * fh * 4;
    if (rank == 0) {
        frames.create(frameBytes, 1, world, handle);
        for (int fd : wfds)
            if (write(fd, handle, sizeof(handle)) != (ssize_t)sizeof(handle)) fail("sortfirst: pipe write failed!");
    } else {
        if (read(rfd, handle, sizeof(handle)) != (ssize_t)sizeof(handle)) fail("sortfirst: pipe read failed!");
        frames.open(frameBytes, 1, world, handle);
    }

    SortFirstRenderer sf(fw, fh, rank, world);
    std::vector<CudaSurface*> depths;
    for (const FrameRect& r : sf.rects()) depths.push_back(new CudaSurface(Vec2i(r.w, r.h), CudaSurface::FORMAT_DEPTH32));
    sf.render(cr, frames.slot(0), depths, Vec4f(0.2f, 0.4f, 0.8f, 1.0f), 1.0f, (const float*)bounds.getCudaPtr());
    frames.publish(0, rank, 1u);
    cudaDeviceSynchronize();
    printf("sortfirst: rank %d of %d on device %d rendered %d of %d rectangles\n", rank, world, device, (int)sf.rects().size(), sf.numRectsOfFrame());
    fflush(stdout);

    int rc = 0;
    if (rank == 0) {
        // wait for every rank's mark, then compare with the frame rendered by this rank alone (no bounds, local memory)
        std::vector<U32> marks((size_t)world);
        for (int spin = 0; spin < 20000; spin++) {
            cudaMemcpy(marks.data(), frames.mark(0, 0), sizeof(U32) * world, cudaMemcpyDeviceToHost);
            bool all = true;
            for (U32 m : marks) all = all && m == 1u;
            if (all) break;
            usleep(1000);
        }
        for (U32 m : marks)
            if (m != 1u) fail("sortfirst: a rank never delivered its rectangles!");
        std::vector<U32> got((size_t)fw * fh), want((size_t)fw * fh);
        cudaMemcpy(got.data(), frames.slot(0), frameBytes, cudaMemcpyDeviceToHost);
        Buffer local;
        local.resizeDiscard((S64)frameBytes);
        SortFirstRenderer whole(fw, fh, 0, 1);
        std::vector<CudaSurface*> d2;
        for (const FrameRect& r : whole.rects()) d2.push_back(new CudaSurface(Vec2i(r.w, r.h), CudaSurface::FORMAT_DEPTH32));
        whole.render(cr, local.getCudaPtr(), d2, Vec4f(0.2f, 0.4f, 0.8f, 1.0f), 1.0f);
        cudaMemcpy(want.data(), local.getCudaPtr(), frameBytes, cudaMemcpyDeviceToHost);
        size_t diff = 0, covered = 0;
        for (size_t i = 0; i < got.size(); i++) { diff += got[i] != want[i]; covered += want[i] != 0xFFCC6633u; }
        printf("sortfirst: %zu of %zu texels differ, %zu covered -> %s\n", diff, got.size(), covered, diff == 0 && covered > got.size() / 2 ? "OK" : "MISMATCH");
        rc = diff == 0 && covered > got.size() / 2 ? 0 : 1;
        if (outPath) {
            FILE* fp = fopen(outPath, "wb");
            if (fp) { fwrite(got.data(), 4, got.size(), fp); fclose(fp); }
        }
        for (CudaSurface* s : d2) delete s;
    }
    for (CudaSurface* s : depths) delete s;
    if (rank != 0) frames.close();
    return rc;
}

int main(int argc, char** argv) {
    const int world = argc > 1 ? atoi(argv[1]) : 2;
    if (world < 1 || world > 16) { printf("usage: sortfirst <world 1..16> [out.raw]\n"); return 2; }
    // fork BEFORE any CUDA call: one process per rank, a pipe from rank 0 to each of the others for the IPC handle
    std::vector<int> wfds;
    std::vector<pid_t> kids;
    for (int r = 1; r < world; r++) {
        int fd[2];
        if (pipe(fd) != 0) return 3;
        const pid_t pid = fork();
        if (pid == 0) {
            close(fd[1]);
            for (int w : wfds) close(w);
            return runRank(r, world, fd[0], {}, NULL);
        }
        close(fd[0]);
        wfds.push_back(fd[1]);
        kids.push_back(pid);
    }
    int rc = runRank(0, world, -1, wfds, argc > 2 ? argv[2] : NULL);
    for (pid_t k : kids) {
        int st = 0;
        waitpid(k, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = rc ? rc : 4;
    }
    return rc;
}
