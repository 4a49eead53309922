/* Dev tool / record: call latency of ikc_resize_u8 through the C ABI when every call misses the weight-table cache
 * (a service fed arbitrary target sizes): 8 threads x N random target sizes on one context; prints p50 / p99 as JSON.
 *   gcc -O2 -std=c99 -Iinclude tools/table_miss_latency.c -o tools/table_miss_latency -Lrust-image-transform_b200 \
 *       -limagekit_cuda -lpthread -Wl,-rpath,$PWD/rust-image-transform_b200 && tools/table_miss_latency 500 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "imagekit_cuda.h"

enum { MAXT = 16, SW = 640, SH = 480, CH = 3 };
static ikc_ctx* ctx;
static unsigned char* src;
static int calls, warmup = 100, THREADS = 8;
static double* lat; /* [THREADS][calls] seconds */
static unsigned* dims; /* [THREADS][calls] dw << 16 | dh */

static double now(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
static unsigned next(unsigned long long* s) {
    *s = *s * 6364136223846793005ull + 1442695040888963407ull;
    return (unsigned)(*s >> 33);
}
static void* worker(void* arg) {
    const int t = (int)(size_t)arg;
    unsigned long long s = 88172645463325252ull + 977ull * (unsigned)t;
    unsigned char* dst = (unsigned char*)malloc(600 * 440 * CH);
    for (int i = -warmup; i < calls; ++i) {   /* the first `warmup` calls grow the lanes' buffers and the memory pool: not recorded */
        const unsigned dw = 40 + next(&s) % 560, dh = 40 + next(&s) % 400; /* almost always a size nobody asked for before */
        const double t0 = now();
        const int rc = ikc_resize_u8(ctx, src, SW, SH, (size_t)SW * CH, CH, dst, dw, dh, (size_t)dw * CH, IKC_FILTER_LANCZOS3);
        if (i >= 0) { lat[(size_t)t * calls + i] = now() - t0; dims[(size_t)t * calls + i] = dw << 16 | dh; }
        if (rc != IKC_OK) { fprintf(stderr, "rc=%d: %s\n", rc, ikc_last_error()); exit(1); }
    }
    free(dst);
    return NULL;
}
static int cmp(const void* a, const void* b) { return (*(const double*)a > *(const double*)b) - (*(const double*)a < *(const double*)b); }

int main(int argc, char** argv) {
    calls = argc > 1 ? atoi(argv[1]) : 500;
    const char* te = getenv("IKC_LAT_THREADS");
    THREADS = te ? atoi(te) : 8;
    if (THREADS < 1 || THREADS > MAXT) THREADS = 8;
    int dev = 0;
    if (ikc_create(&dev, 1, &ctx) != IKC_OK) { fprintf(stderr, "ikc_create: %s\n", ikc_last_error()); return 1; }
    src = (unsigned char*)malloc((size_t)SW * SH * CH);
    for (size_t i = 0; i < (size_t)SW * SH * CH; ++i) src[i] = (unsigned char)(i * 2654435761u >> 24);
    lat = (double*)malloc(sizeof(double) * MAXT * (size_t)calls);
    dims = (unsigned*)malloc(sizeof(unsigned) * MAXT * (size_t)calls);
    unsigned char* warm = (unsigned char*)malloc(100 * 75 * CH);
    for (int i = 0; i < 8; ++i) ikc_resize_u8(ctx, src, SW, SH, (size_t)SW * CH, CH, warm, 100, 75, 300, IKC_FILTER_LANCZOS3);
    pthread_t th[MAXT];
    const double t0 = now();
    for (int t = 0; t < THREADS; ++t) pthread_create(&th[t], NULL, worker, (void*)(size_t)t);
    for (int t = 0; t < THREADS; ++t) pthread_join(th[t], NULL);
    const double wall = now() - t0;
    const size_t n = (size_t)THREADS * calls;
    if (getenv("IKC_LAT_VERBOSE"))
        for (size_t i = 0; i < n; ++i)
            if (lat[i] > 5e-3) fprintf(stderr, "slow: thread %zu call %zu -> %ux%u %.2f ms\n", i / calls, i % calls, dims[i] >> 16, dims[i] & 0xffff, lat[i] * 1e3);
    qsort(lat, n, sizeof(double), cmp);
    printf("{\"threads\": %d, \"calls_per_thread\": %d, \"p50_ms\": %.4f, \"p90_ms\": %.4f, \"p99_ms\": %.4f, \"max_ms\": %.4f, \"calls_per_s\": %.1f, "
           "\"workload\": \"640x480 RGB8 -> random (40..600) x (40..440) Lanczos3 through ikc_resize_u8 (C, pthreads), pageable buffers, one context, "
           "every call a weight-table miss; 100 unrecorded warm-up calls per thread\"}\n",
           THREADS, calls, lat[n / 2] * 1e3, lat[n * 9 / 10] * 1e3, lat[n * 99 / 100] * 1e3, lat[n - 1] * 1e3, (double)n / wall);
    ikc_destroy(ctx);
    return 0;
}
