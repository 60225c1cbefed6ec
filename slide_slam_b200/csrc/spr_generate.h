// spr_generate.h -- launch wrappers of the device-side descriptor / hypothesis generator (spr_generate.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// stable LSD radix sort of the low `bits` bits of n 64-bit keys (tmp: n keys, hist: spr_radix_sort_hist_words(n) words)
cudaError_t spr_radix_sort_u64(unsigned long long *keys, unsigned long long *tmp, long long n, int bits, uint32_t *hist,
                               cudaStream_t st);
size_t spr_radix_sort_hist_words(long long n);

// counting sort of the data descriptors by bin of their first component: bin_start[n_bins + 1] offsets
// (bin_start needs n_bins + 2 words, bin_fill n_bins + 1), `binned` = spr_gen_binned_bytes(t_data) bytes
cudaError_t spr_launch_gen_bin(const double *desc_data, int t_data, double inv_w, uint32_t n_bins, uint32_t *bin_start,
                               uint32_t *bin_fill, void *binned, cudaStream_t st);
size_t spr_gen_binned_bytes(int t_data);
// windowed matching: appends keys model_idx * t_data + data_idx (unordered); *total counts all matches (may exceed cap)
cudaError_t spr_launch_gen_match(const double *desc_model, int t_model, const void *binned, const uint32_t *bin_start, uint32_t n_bins,
                                 double inv_w, double window, double Tstar, const double *sig_model, const double *sig_data,
                                 int t_data, unsigned long long *keys, unsigned long long cap, unsigned long long *total,
                                 int sm_count, cudaStream_t st);
cudaError_t spr_launch_gen_unpack(const unsigned long long *keys, long long n, int t_data, int32_t *model_idx, int32_t *data_idx,
                                  cudaStream_t st);
// per match: 2-D Kabsch hypothesis (c, s, x, y) data -> model (hyps4, optional) and the matched point
// lists (3 x 2 doubles per match and side, optional) in sorted-descriptor vertex order
cudaError_t spr_launch_gen_kabsch(const double *tris_model6, const double *tris_data6, const int32_t *perm_model,
                                  const int32_t *perm_data, const int32_t *model_idx, const int32_t *data_idx, long long n,
                                  double *hyps4, double *pts_model, double *pts_data, cudaStream_t st);
