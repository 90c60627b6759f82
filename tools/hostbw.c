/* Host DRAM write-bandwidth probe (pthreads): how fast can T threads fill a large buffer with
 * (a) regular stores (memset per 4 KB piece), (b) non-temporal stores from an L1-resident chunk.
 * Used to size the packed device->host Jacobian transport.
 *   gcc -O2 -mavx2 -pthread tools/hostbw.c -o /tmp/hostbw && /tmp/hostbw [threads] [GiB] */
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

typedef struct { char* dst; size_t bytes; int mode; } job_t;

static void* worker(void* a) {
    job_t* j = (job_t*)a;
    static __thread double buf[512] __attribute__((aligned(64)));
    memset(buf, 0, sizeof buf);
    for (size_t o = 0; o + 4096 <= j->bytes; o += 4096) {
        if (j->mode == 0) {
            memset(j->dst + o, 0, 4096);
        } else {
            buf[(o >> 12) & 511] = 1.0;                       /* touch the chunk like a scatter would */
            for (int k = 0; k < 512; k += 4)
                _mm256_stream_pd((double*)(j->dst + o) + k, _mm256_load_pd(buf + k));
        }
    }
    _mm_sfence();
    return 0;
}

int main(int argc, char** argv) {
    int T = argc > 1 ? atoi(argv[1]) : 8;
    double gib = argc > 2 ? atof(argv[2]) : 2.0;
    size_t bytes = (size_t)(gib * (1u << 30)) / T / 4096 * 4096;
    char* base = aligned_alloc(4096, bytes * T);
    memset(base, 1, bytes * T);                               /* fault the pages in */
    for (int mode = 0; mode < 2; mode++)
        for (int rep = 0; rep < 3; rep++) {
            pthread_t th[256]; job_t jb[256];
            double t0 = now();
            for (int t = 0; t < T; t++) { jb[t] = (job_t){base + t * bytes, bytes, mode}; pthread_create(&th[t], 0, worker, &jb[t]); }
            for (int t = 0; t < T; t++) pthread_join(th[t], 0);
            double dt = now() - t0;
            printf("threads %d mode %s rep %d: %.1f GB/s\n", T, mode ? "nt-stream" : "memset", rep, bytes * T / dt / 1e9);
        }
    return 0;
}
