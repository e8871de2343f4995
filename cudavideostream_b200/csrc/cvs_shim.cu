// cvs_shim.cu -- diff::cuda::CUDACore re-implemented over the C ABI (include/cvs_b200.h).
//
// Replaces the class body of server/src/kernels.cu:377-536; signatures unchanged (see
// include/cvs_cuda_core.hpp).  Everything here is host glue: argument translation and the
// reference's print-and-exit error behaviour (kernels.cu:11-22).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/cvs_b200.h"
#include "../../include/cvs_cuda_core.hpp"

#define CVS_SHIM_API __attribute__((visibility("default")))

namespace {

[[noreturn]] void die(cvs_status st, const char *what)
{
    // kernels.cu:14-20: "<file>:<line> (<code>)\n<message>" then exit(status)
    fprintf(stderr, "%s (%d)\n%s\n", what, (int)st, cvs_last_error());
    exit((int)st);
}

int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

} // namespace

namespace diff {
namespace cuda {

CVS_SHIM_API CUDACore::CUDACore(uint8_t *charsPx, diff::utils::matsz &charsSz, float *k, int total,
                                uint8_t *sampleMatData, diff::utils::matsz &frameSz)
    : handle_(nullptr), frameSz_(frameSz), total_(total)
{
    cvs_config cfg;
    cvs_config_default(&cfg);
    cfg.width = frameSz.width;
    cfg.height = frameSz.height;
#ifdef LR_THRESHOLDS
    cfg.threshold = LR_THRESHOLDS;
#endif
#ifdef NOISE_VISUALIZER
    cfg.mode = NOISE_VISUALIZER;
#endif
#ifdef NOISE_FILTER
    cfg.noise_filter = 1;
#endif
#ifdef K
    cfg.ksize = K;
#endif
    cfg.threshold = env_int("CVS_LR_THRESHOLDS", cfg.threshold);
    cfg.mode = env_int("CVS_NOISE_VISUALIZER", cfg.mode);
    cfg.noise_filter = env_int("CVS_NOISE_FILTER", cfg.noise_filter);
    cfg.ksize = env_int("CVS_K", cfg.ksize);
    cfg.device = env_int("CVS_DEVICE", 0);
    cfg.kweights = k;
    cfg.base_frame = sampleMatData;
    cfg.glyphs = charsPx;
    cfg.glyph_w = charsSz.width;
    cfg.glyph_h = charsSz.height;
    const char *chars = getenv("CVS_CHARS_STR");
#ifdef CHARS_STR
    if (!chars) chars = CHARS_STR;
#endif
    cfg.glyph_chars = chars ? chars : "0123456789BFPSWbkps :/"; // common.h:13
    if (total != 3 * frameSz.width * frameSz.height) {
        fprintf(stderr, "CUDACore: total (%d) != 3*%d*%d\n", total, frameSz.width, frameSz.height);
        exit(1);
    }
    cvs_status st = cvs_create(&cfg, &handle_);
    if (st != CVS_OK) die(st, "cvs_create");
}

// kernels.cu:531-536: pinned host buffers with one chunk_t of slack
CVS_SHIM_API void CUDACore::alloc_arrays(uint8_t **h_frame, uint8_t **n_frame, uint8_t **o_frame, int **h_xs, int r, int c)
{
    const size_t n = (size_t)3 * r * c;
    void *p = nullptr;
    cvs_status st;
    if ((st = cvs_alloc_host(&p, n + 32)) != CVS_OK) die(st, "cvs_alloc_host");
    *h_frame = (uint8_t *)p;
    if ((st = cvs_alloc_host(&p, n + 32)) != CVS_OK) die(st, "cvs_alloc_host");
    *n_frame = (uint8_t *)p;
    if ((st = cvs_alloc_host(&p, n + 32)) != CVS_OK) die(st, "cvs_alloc_host");
    *o_frame = (uint8_t *)p;
    if ((st = cvs_alloc_host(&p, n * sizeof(int) + 32)) != CVS_OK) die(st, "cvs_alloc_host");
    *h_xs = (int *)p;
}

// kernels.cu:430-525
CVS_SHIM_API void CUDACore::exec_core(uint8_t *frameData, uint8_t *showReadyNData, std::string &text,
                                      unsigned int *h_pos, int *h_xs)
{
    cvs_status st = cvs_exec(handle_, frameData, showReadyNData, text.c_str(), h_pos, h_xs);
    if (st != CVS_OK) die(st, "cvs_exec");
}

// kernels.cu:527-529: sizeof(chunk_t) = sizeof(long4); kept for source compatibility
CVS_SHIM_API size_t CUDACore::chunkt_size() { return 32; }

} // namespace cuda
} // namespace diff
