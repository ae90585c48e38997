/* oracle/par.c — TEST INFRASTRUCTURE ONLY. See par.h. */
#include "par.h"

#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

static int g_threads = 0;

int oracle_num_threads(void) {
  if (g_threads <= 0) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    g_threads = n > 0 ? (int)n : 1;
  }
  return g_threads;
}

void oracle_set_num_threads(int n) {
  if (n > 0) g_threads = n;
}

typedef struct {
  long long begin, end;
  oracle_body_fn body;
  void *ctx;
} job_t;

static void *run_job(void *p) {
  job_t *j = (job_t *)p;
  j->body(j->begin, j->end, j->ctx);
  return NULL;
}

void oracle_parallel_for(long long n, oracle_body_fn body, void *ctx) {
  int t = oracle_num_threads();
  if (t > n) t = n > 0 ? (int)n : 1;
  if (t <= 1) {
    body(0, n, ctx);
    return;
  }
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)t);
  job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)t);
  for (int i = 0; i < t; ++i) {
    jobs[i].begin = n * i / t;
    jobs[i].end = n * (i + 1) / t;
    jobs[i].body = body;
    jobs[i].ctx = ctx;
    pthread_create(&th[i], NULL, run_job, &jobs[i]);
  }
  for (int i = 0; i < t; ++i) pthread_join(th[i], NULL);
  free(jobs);
  free(th);
}
