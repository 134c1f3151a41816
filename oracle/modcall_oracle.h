/* modcall_oracle.h -- TEST INFRASTRUCTURE ONLY: interface of the CPU oracle (see modcall_oracle.c).
 * Shares only the plain data types of include/minimod_cuda.h (batch layout and record structs). */
#ifndef MODCALL_ORACLE_H
#define MODCALL_ORACLE_H
#include <stdint.h>
#include "../include/minimod_cuda.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct oracle_ctx oracle_ctx;
/* thresh[i]: the double threshold of mods[i] (the oracle evaluates the reference's double comparison itself, it does
 * not use call_lut) */
oracle_ctx *oracle_create(int subtool, int n_mods, const mmc_mod_t *mods, const double *thresh, int insertions, int haplotypes,
                          int n_contigs, const uint32_t *lens);
void oracle_destroy(oracle_ctx *c);
int oracle_ref_add(oracle_ctx *c, int tid, const char *seq, uint32_t len);
int oracle_process_batch(oracle_ctx *c, const mmc_batch_t *b);
int oracle_freq_records(oracle_ctx *c, const mmc_freq_rec_t **recs, uint64_t *n);
int oracle_view_records(oracle_ctx *c, const mmc_view_rec_t **recs, uint64_t *n);
const char *oracle_code_name(const oracle_ctx *c, int code);
const char *oracle_strerror(const oracle_ctx *c);
#ifdef __cplusplus
}
#endif
#endif
