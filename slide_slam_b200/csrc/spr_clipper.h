// spr_clipper.h -- device-resident CLIPPER problem (affinity CSR + solver state), spr_clipper.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/slide_pr.h"

struct SprClipper;

SprClipper *spr_clipper_create();
void spr_clipper_destroy(SprClipper *c);

// scorePairwiseConsistency (clipper.cpp:21-65).  D1 / D2: dim x n column-major HOST arrays (Eigen's
// invariants::Data), or -- device_hint != NULL -- DEVICE arrays of the same layout (the generator's matched
// point lists); device_hint = {centre of cloud 1 [3], centre of cloud 2 [3], bound on |coordinate - centre|}
// then replaces the host pass that scales the fp32 prefilter's rounding margin.
// A: m x 2 host array, or NULL for the all-to-all hypothesis (utils.h:60-70).
int spr_clipper_score(SprClipper *c, const slide_clipper_params &p, const double *D1, int n1, const double *D2, int n2,
                      int dim, const int32_t *A, int m, const double *device_hint, int sm_count, cudaStream_t st,
                      long long *nnz_upper, float *kernel_ms, std::string &err);
int spr_clipper_size(const SprClipper *c, int *m, long long *nnz_sym);
const int32_t *spr_clipper_associations(const SprClipper *c);
// symmetric CSR without the diagonal -> host arrays (row_ptr: m + 1)
int spr_clipper_get_csr(SprClipper *c, int64_t *row_ptr, int32_t *col, double *val, long long cap, cudaStream_t st,
                        std::string &err);
// findDenseClique (clipper.cpp:172-323) with the given u0 (host, m doubles) + rounding
int spr_clipper_solve(SprClipper *c, const slide_clipper_params &p, const double *u0, int sm_count, cudaStream_t st,
                      int32_t *nodes_out, int32_t cap, slide_clipper_solution *sol, double *u_out, std::string &err);
