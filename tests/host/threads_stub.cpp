// threads_stub.cpp -- a file-fed stand-in for the reference's capture/show threads, TEST INFRASTRUCTURE ONLY.
//
// Implements the nine members of diff::threads::ThreadsCore (server/include/threads.hpp:37-47) without OpenCV, a
// webcam or sockets, so that the reference's UNMODIFIED server/src/server.cpp can be compiled from where it lies
// under /root/reference and linked
//   * against libcvs_b200.so (its `#ifdef GPU` branch, server.cpp:53,139)   -> the drop-in boundary, end to end;
//   * with -DCPU and nothing else                                           -> the reference's own CPU filter chain
//                                                                              (server.cpp:96-135) as oracle/_ref;
//   * against the reference's own server/src/kernels.cu built for sm_100a    -> "reference GPU code on B200".
// The real implementation (server/src/threads.cpp:30-175) captures MJPG frames with OpenCV and hands them to main()
// through pipes; this one reads them from a file and records what main() hands back.
//
//   CVS_STUB_IN   input : int32 W, H, nframes, gw, gh; glyph atlas 22*3*gw*gh bytes; base frame; nframes frames
//   CVS_STUB_OUT  output: per frame  u32 pos, u32 kind, then kind 0: i32 xs[pos], u8 data[pos]
//                                                            kind 1: u8 data[N]   (STUB_CPU: whole frame comes back)
//                         followed by u8 show[N] when CVS_STUB_SHOW=1 (showReadyNData as writeNoise() sees it)
//   CVS_STUB_TIMES optional: text file, one line per frame: nanoseconds between readCap() returning and writeShow()
// readCap() ends the process with exit(0) once the file is exhausted (server.cpp's loop never ends on its own).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "../include/threads.hpp"
#ifndef STUB_CPU
#include "../include/kernels.cuh"
#endif

namespace {

struct StubState {
    FILE *in = nullptr, *out = nullptr, *times = nullptr;
    int w = 0, h = 0, nframes = 0, gw = 0, gh = 0, next = 0;
    size_t n = 0;
    std::vector<uint8_t> base, glyphs;
    uint8_t *frame = nullptr, *show = nullptr, *spare = nullptr;
    int *xs = nullptr;
    unsigned int pos = 0;
    bool want_show = false;
    std::chrono::high_resolution_clock::time_point t0;
};

StubState *S(void *p) { return static_cast<StubState *>(p); }

[[noreturn]] void die(const char *what)
{
    fprintf(stderr, "threads_stub: %s\n", what);
    exit(2);
}

void read_exact(FILE *f, void *dst, size_t bytes)
{
    if (bytes && fread(dst, 1, bytes, f) != bytes) die("short read on CVS_STUB_IN");
}

} // namespace

namespace diff {
namespace threads {

ThreadsCore::ThreadsCore()
{
    StubState *s = new StubState;
    this->pctx = s;
    const char *in = getenv("CVS_STUB_IN"), *out = getenv("CVS_STUB_OUT");
    if (!in || !out) die("CVS_STUB_IN and CVS_STUB_OUT must be set");
    s->in = fopen(in, "rb");
    s->out = fopen(out, "wb");
    if (!s->in || !s->out) die("cannot open CVS_STUB_IN / CVS_STUB_OUT");
    if (const char *t = getenv("CVS_STUB_TIMES")) s->times = fopen(t, "w");
    s->want_show = getenv("CVS_STUB_SHOW") && atoi(getenv("CVS_STUB_SHOW")) != 0;
    int32_t hdr[5];
    read_exact(s->in, hdr, sizeof hdr);
    s->w = hdr[0]; s->h = hdr[1]; s->nframes = hdr[2]; s->gw = hdr[3]; s->gh = hdr[4];
    s->n = (size_t)3 * s->w * s->h;
    s->glyphs.resize((size_t)22 * 3 * s->gw * s->gh + 64);
    read_exact(s->in, s->glyphs.data(), (size_t)22 * 3 * s->gw * s->gh);
    s->base.resize(s->n);
    read_exact(s->in, s->base.data(), s->n);
    this->frameSz = diff::utils::matsz(s->h, s->w);
    this->charSz = diff::utils::matsz(s->gh, s->gw);
    this->charsPx = s->glyphs.data();
#ifndef STUB_CPU
    // server/src/threads.cpp:95: every host buffer comes from CUDACore::alloc_arrays (pinned)
    diff::cuda::CUDACore::alloc_arrays(&s->frame, &s->show, &s->spare, &s->xs, s->h, s->w);
#else
    s->frame = new uint8_t[s->n + 32];
    s->show = new uint8_t[s->n + 32];
    s->xs = new int[s->n + 8];
#endif
    memset(s->show, 0, s->n);
}

diff::utils::matsz ThreadsCore::getFrameSize() { return this->frameSz; }
diff::utils::matsz ThreadsCore::getCharSize() { return this->charSz; }
uint8_t *ThreadsCore::getCharsPx() { return this->charsPx; }
uint8_t *ThreadsCore::getBaseFrameData() { return S(this->pctx)->base.data(); }
uint8_t *ThreadsCore::getShowReadyNData() { return S(this->pctx)->show; }

void ThreadsCore::readCap(struct preadymin &minready)
{
    StubState *s = S(this->pctx);
    if (s->next >= s->nframes) {
        fclose(s->out);
        if (s->times) fclose(s->times);
        exit(0);
    }
    read_exact(s->in, s->frame, s->n);
    s->next++;
    s->pos = 0;
    minready.data = s->frame;
    minready.h_pos = &s->pos;
    minready.h_xs = s->xs;
    minready.__ptr = nullptr;
    s->t0 = std::chrono::high_resolution_clock::now();
}

void ThreadsCore::writeNoise() {}

void ThreadsCore::writeShow(struct preadymin &minready)
{
    StubState *s = S(this->pctx);
    const auto t1 = std::chrono::high_resolution_clock::now();
    if (s->times)
        fprintf(s->times, "%lld\n", (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - s->t0).count());
#ifndef STUB_CPU
    const uint32_t pos = *minready.h_pos, kind = 0;
    fwrite(&pos, 4, 1, s->out);
    fwrite(&kind, 4, 1, s->out);
    fwrite(minready.h_xs, 4, pos, s->out);
    fwrite(minready.data, 1, pos, s->out);
#else
    const uint32_t pos = 0, kind = 1;
    fwrite(&pos, 4, 1, s->out);
    fwrite(&kind, 4, 1, s->out);
    fwrite(minready.data, 1, s->n, s->out);
#endif
    if (s->want_show) fwrite(s->show, 1, s->n, s->out);
}

} // namespace threads
} // namespace diff
