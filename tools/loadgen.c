/* loadgen.c -- closed-loop load generator for the in-library micro-batcher (measurement tool, not product code):
 * N native threads each call epp_submit + epp_wait in a loop for a fixed time, exactly the call shape one goroutine per
 * in-flight request has in the reference (requestcontrol/director.go:69-71, handlers/server.go:168).  The function
 * pointers are passed in by the Python driver (tools/batcher_load.py), so this file needs neither the header nor the
 * library at build time. */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int32_t (*submit_fn)(void *b, uint32_t model, const void *prompt, uint64_t len, uint32_t mm, uint64_t *ticket);
typedef int32_t (*wait_fn)(void *b, uint64_t ticket, void *dec, void *det);

typedef struct {
    submit_fn submit; wait_fn wait; void *batcher;
    const uint8_t *prompts; uint64_t prompt_len; int64_t n_prompts;
    double seconds; int tid, n_threads;
    float *lat_us; int64_t lat_cap; int64_t n_done; int32_t n_err;
    uint8_t *picks_out;              /* 32-byte decision of every prompt index this thread scheduled last (optional) */
} job;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *worker(void *arg) {
    job *j = (job *)arg;
    uint8_t dec[32], det[40];
    const double t_end = now_s() + j->seconds;
    int64_t i = j->tid;
    while (1) {
        const double t0 = now_s();
        if (t0 >= t_end) break;
        const int64_t p = i % j->n_prompts;
        uint64_t ticket = 0;
        int32_t rc = j->submit(j->batcher, 0, j->prompts + (uint64_t)p * j->prompt_len, j->prompt_len, 0, &ticket);
        if (rc == 0) rc = j->wait(j->batcher, ticket, dec, det);
        const double t1 = now_s();
        if (rc != 0) { j->n_err++; break; }
        if (j->n_done < j->lat_cap) j->lat_us[j->n_done] = (float)((t1 - t0) * 1e6);
        if (j->picks_out) memcpy(j->picks_out + 32 * p, dec, 32);
        j->n_done++;
        i += j->n_threads;
    }
    return NULL;
}

/* Runs n_threads workers for `seconds`; lat_us: [n_threads][lat_cap] per-request latencies; n_done: [n_threads].
 * Returns the number of failed calls. */
int64_t loadgen_run(void *submit, void *wait, void *batcher, const uint8_t *prompts, uint64_t prompt_len, int64_t n_prompts,
                    int n_threads, double seconds, float *lat_us, int64_t lat_cap, int64_t *n_done, uint8_t *picks_out) {
    job *jobs = (job *)calloc((size_t)n_threads, sizeof(job));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = (job){(submit_fn)submit, (wait_fn)wait, batcher, prompts, prompt_len, n_prompts, seconds, t, n_threads,
                        lat_us + (int64_t)t * lat_cap, lat_cap, 0, 0, picks_out};
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    int64_t err = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        n_done[t] = jobs[t].n_done;
        err += jobs[t].n_err;
    }
    free(jobs); free(th);
    return err;
}
