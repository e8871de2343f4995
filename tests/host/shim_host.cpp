// shim_host.cpp -- a miniature of the reference server's main loop (server/src/server.cpp:38-146) that uses
// diff::cuda::CUDACore exactly as the reference does: alloc_arrays for the pinned ring (threads.cpp:95), the
// six-argument constructor (server.cpp:53) and exec_core per frame (server.cpp:139).  Frames come from a file
// instead of the webcam; the payload of every frame is written to another file for the parity test.
//
//   shim_host <in.bin> <out.bin> <text>
//   in.bin : int32 width, height, nframes | base frame | nframes frames | int32 gw, gh, nglyph | glyph atlas
//   out.bin: per frame: uint32 pos | int32 xs[pos] | uint8 diff[pos] | show frame (N bytes)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cvs_cuda_core.hpp"

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    FILE *fi = fopen(argv[1], "rb"), *fo = fopen(argv[2], "wb");
    if (!fi || !fo) return 3;
    int hdr[3];
    if (fread(hdr, sizeof hdr, 1, fi) != 1) return 4;
    const int width = hdr[0], height = hdr[1], nframes = hdr[2];
    const int total = 3 * width * height;
    std::vector<uint8_t> base(total);
    if (fread(base.data(), 1, total, fi) != (size_t)total) return 4;
    std::vector<std::vector<uint8_t>> frames(nframes, std::vector<uint8_t>(total));
    for (auto &f : frames)
        if (fread(f.data(), 1, total, fi) != (size_t)total) return 4;
    int gh[3];
    if (fread(gh, sizeof gh, 1, fi) != 1) return 4;
    std::vector<uint8_t> glyphs((size_t)3 * gh[0] * gh[1] * gh[2]);
    if (!glyphs.empty() && fread(glyphs.data(), 1, glyphs.size(), fi) != glyphs.size()) return 4;

    // threads.cpp:86-106: the ring of pinned buffers
    uint8_t *h_frame[2], *n_frame[2], *o_frame[2];
    int *h_xs[2];
    for (int i = 0; i < 2; i++)
        diff::cuda::CUDACore::alloc_arrays(&h_frame[i], &n_frame[i], &o_frame[i], &h_xs[i], height, width);

    float k[9];
    for (int i = 0; i < 9; i++) k[i] = 1.0f / 9.0f; // unused unless CVS_NOISE_FILTER=1
    diff::utils::matsz charsSz(gh[1], gh[0]), frameSz(height, width);
    diff::cuda::CUDACore cudaCore(glyphs.empty() ? nullptr : glyphs.data(), charsSz, k, total, base.data(), frameSz);
    if (cudaCore.chunkt_size() != 32) return 5;

    std::string text = argv[3];
    for (int t = 0; t < nframes; t++) {
        const int b = t & 1;
        memcpy(h_frame[b], frames[t].data(), total); // the capture thread's cap >> *pframe (threads.cpp:173-174)
        memset(n_frame[b], 0, total);
        unsigned int h_pos = 0;
        cudaCore.exec_core(h_frame[b], n_frame[b], text, &h_pos, h_xs[b]);
        // the send thread's wire format (threads.cpp:229-231): pos | xs[pos] | diff[pos]
        fwrite(&h_pos, sizeof h_pos, 1, fo);
        fwrite(h_xs[b], sizeof(int), h_pos, fo);
        fwrite(h_frame[b], 1, h_pos, fo);
        fwrite(n_frame[b], 1, total, fo);
    }
    fclose(fi);
    fclose(fo);
    return 0;
}
