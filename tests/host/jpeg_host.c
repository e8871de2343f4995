/* jpeg_host.c -- a capture thread in plain C (gcc, no CUDA headers) against the C ABI of include/cvs_b200.h: the camera's
 * MJPG buffers go to cvs_submit_jpeg instead of through OpenCV's decoder (what INTEGRATION.md proposes for
 * server/src/threads.cpp:32-41, :118 and server.cpp:139), up to four tickets in flight, payload written to a file.
 *
 *   jpeg_host <width> <height> <out.bin> <frame0.jpg> <frame1.jpg> ...      (frame0 seeds the reference frame all-zero)
 * out.bin: per frame  u32 pos, i32 xs[pos], u8 diff[pos]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cvs_b200.h"

#define CHECK(call)                                                                          \
    do {                                                                                     \
        cvs_status s_ = (call);                                                              \
        if (s_ != CVS_OK) {                                                                  \
            fprintf(stderr, "%s failed: %d (%s)\n", #call, (int)s_, cvs_last_error());       \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    const int w = atoi(argv[1]), h = atoi(argv[2]);
    const size_t n = (size_t)3 * w * h;
    const int nframes = argc - 4;
    if (cvs_device_count() < 1) {
        fprintf(stderr, "no sm_100 device\n");
        return 3;
    }
    uint8_t *base = (uint8_t *)calloc(n, 1);
    cvs_config cfg;
    cvs_config_default(&cfg);
    cfg.width = w;
    cfg.height = h;
    cfg.base_frame = base;
    cvs_handle hd;
    CHECK(cvs_create(&cfg, &hd));
    /* four slots: the JPEG as the camera delivered it, the payload buffers, the count */
    uint8_t *jbuf[4], *diff[4];
    int *xs[4];
    unsigned int pos[4];
    uint64_t ticket[4];
    for (int k = 0; k < 4; k++) {
        CHECK(cvs_alloc_host((void **)&jbuf[k], n));
        CHECK(cvs_alloc_host((void **)&diff[k], n + 32));
        CHECK(cvs_alloc_host((void **)&xs[k], 4 * n + 32));
    }
    FILE *out = fopen(argv[3], "wb");
    if (!out) return 4;
    for (int t = 0; t < nframes + 4; t++) {
        const int k = t & 3;
        if (t >= 4) { /* the slot's previous ticket: wait, hand the payload on (server.cpp:143 writeShow) */
            CHECK(cvs_wait(hd, ticket[k]));
            fwrite(&pos[k], 4, 1, out);
            fwrite(xs[k], 4, pos[k], out);
            fwrite(diff[k], 1, pos[k], out);
        }
        if (t < nframes) {
            FILE *f = fopen(argv[4 + t], "rb");
            if (!f) return 5;
            const size_t len = fread(jbuf[k], 1, n, f);
            fclose(f);
            CHECK(cvs_submit_jpeg(hd, jbuf[k], len, diff[k], NULL, "", &pos[k], xs[k], &ticket[k]));
        }
    }
    fclose(out);
    for (int k = 0; k < 4; k++) {
        cvs_free_host(jbuf[k]);
        cvs_free_host(diff[k]);
        cvs_free_host(xs[k]);
    }
    CHECK(cvs_destroy(hd));
    free(base);
    return 0;
}
