/*
 * cvs_cuda_core.hpp -- drop-in declaration of diff::cuda::CUDACore for the reference server.
 *
 * Replaces server/include/kernels.cuh:13-43 of MatteoBattilana/CUDAVideoStream.  The four public
 * members keep the reference's exact signatures (kernels.cuh:38-41), so server/src/server.cpp:53,139
 * and server/src/threads.cpp:95 compile and link unchanged; the private state of the reference class
 * (a dozen raw device pointers) is replaced by one opaque handle of the C ABI in cvs_b200.h.
 * Implementation: cudavideostream_b200/csrc/cvs_shim.cu (inside libcvs_b200.so).
 *
 * The compile-time switches of server/include/common.h become run-time configuration read from the
 * environment at construction (defaults = the reference's defaults):
 *   CVS_NOISE_VISUALIZER = 0..7   (common.h:10 NOISE_VISUALIZER; 6/7 = average-gray variants)
 *   CVS_NOISE_FILTER     = 0|1    (common.h:5  NOISE_FILTER)
 *   CVS_K                = odd K  (common.h:6  K; must match the k[] array the caller passes)
 *   CVS_LR_THRESHOLDS    = T      (common.h:14 LR_THRESHOLDS)
 *   CVS_CHARS_STR        = atlas alphabet (common.h:13 CHARS_STR)
 *   CVS_DEVICE           = CUDA device ordinal (the reference hard-codes device 0, kernels.cu:385)
 * When built inside the reference tree the macros of common.h, if visible, provide the defaults.
 *
 * Errors follow the reference: message on stderr, then exit(status) (kernels.cu:11-22).
 */
#ifndef CVS_CUDA_CORE_HPP_
#define CVS_CUDA_CORE_HPP_

#include <stddef.h>
#include <stdint.h>

#include <string>

#if defined(__has_include)
#if __has_include("../include/utils.hpp")
#include "../include/utils.hpp" /* the reference's own diff::utils::matsz (server/include/utils.hpp:7-16) */
#endif
#endif

#ifndef UTILS_HPP_
#define UTILS_HPP_
namespace diff {
namespace utils {
/* layout-compatible stand-in for server/include/utils.hpp:7-16, used when this header is compiled
 * outside the reference tree (tests/host/) */
typedef struct matsz {
    int height;
    int width;
    matsz(int h, int w) : height(h), width(w) {}
    matsz() : matsz(0, 0) {}
    int area() { return height * width; }
} matsz;
} // namespace utils
} // namespace diff
#endif

struct cvs_stream_s;

namespace diff {
namespace cuda {

class CUDACore {
  private:
    cvs_stream_s *handle_;
    diff::utils::matsz frameSz_;
    int total_;

  public:
    CUDACore(uint8_t *charsPx, diff::utils::matsz &charsSz, float *k, int total, uint8_t *sampleMatData,
             diff::utils::matsz &frameSz);
    static void alloc_arrays(uint8_t **h_frame, uint8_t **n_frame, uint8_t **o_frame, int **h_xs, int r, int c);
    void exec_core(uint8_t *frameData, uint8_t *showReadyNData, std::string &text, unsigned int *h_pos, int *h_xs);
    size_t chunkt_size();
};

} // namespace cuda
} // namespace diff

#endif /* CVS_CUDA_CORE_HPP_ */
