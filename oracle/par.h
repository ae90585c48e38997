/* oracle/par.h — TEST INFRASTRUCTURE ONLY. A minimal pthread parallel-for (this image's gcc ships without libgomp). */
#ifndef ORACLE_PAR_H
#define ORACLE_PAR_H
typedef void (*oracle_body_fn)(long long begin, long long end, void *ctx);
/* Splits [0, n) into contiguous chunks over oracle_num_threads() threads and runs body on each. */
void oracle_parallel_for(long long n, oracle_body_fn body, void *ctx);
int oracle_num_threads(void);
void oracle_set_num_threads(int n);
#endif
