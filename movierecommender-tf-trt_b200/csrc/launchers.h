// Host-side launch functions shared between the .cu files and the C-ABI layer (api.cu).
#pragma once

#include "common.cuh"

namespace mr {

// ---- neumf_kernels.cu ------------------------------------------------------------------------
struct TileLaunch {
  const MrModel* model;
  bool train;
  const float* wt;  // transposed hidden kernels, dense-block layout (train only)
  const int32_t* users;
  const int32_t* items;
  const float* labels;
  int64_t B;
  int32_t user_div;
  float inv_batch;
  float* logits;
  float* probs;
  float* loss_partial;   // (max_tile_ctas())
  float* dense_partial;  // (max_tile_ctas(), dense_stride)
  int64_t dense_stride;
  float* stage_u;
  float* stage_i;
  int32_t* flags;
};
int max_tile_ctas();
int choose_tile_rows(const MrModel& m, bool train);
int launch_neumf_tiles(const TileLaunch& a, cudaStream_t st, int* grid_out);
int launch_transpose_kernels(const MrModel& m, float* wt, cudaStream_t st);
int launch_dense_reduce(const MrModel& m, const float* partial, int64_t stride, int grid_ctas, float* out,
                        cudaStream_t st);
int launch_sum_partials(const float* partial, int n, float* out, cudaStream_t st);
int launch_l2_penalty(const float* x, int64_t n, float coef, float* out, cudaStream_t st);

// ---- gather.cu -------------------------------------------------------------------------------
int launch_gather_rows(const float* table, int64_t rows, int dim, const int32_t* idx, int64_t n, float* out,
                       cudaStream_t st);

// ---- radix_sort.cu ---------------------------------------------------------------------------
size_t sort_workspace_bytes(int64_t n);
int launch_sort_pairs(const int32_t* keys, int64_t n, int key_bits, int32_t* out_keys, int32_t* out_index,
                      void* ws, size_t ws_bytes, cudaStream_t st);

// ---- segreduce.cu ----------------------------------------------------------------------------
struct RowUpdate {       // what to do with each reduced row gradient
  int mode;              // MrTableMode
  int optimizer;         // MrOptimizer
  float lr_t, beta_1, beta_2, epsilon, lr;
  // two tables share one staged row: columns [0,d0) -> table 0, [d0,d0+d1) -> table 1
  int d0, d1;
  float *p0, *m0, *v0, *g0;  // table / Adam state / dense gradient table (mode DENSE writes g)
  float *p1, *m1, *v1, *g1;
};
size_t segreduce_workspace_bytes(int64_t n, int ld);
int launch_segreduce(const int32_t* sorted_keys, const int32_t* sorted_index, int64_t n, const float* staged,
                     const RowUpdate& u, void* ws, size_t ws_bytes, cudaStream_t st);

// ---- optimizer.cu ----------------------------------------------------------------------------
int launch_optimizer_flat(float* p, const float* g, float* m, float* v, int64_t n, int optimizer, float lr_t,
                          float beta_1, float beta_2, float epsilon, float l2, cudaStream_t st);

// ---- rank.cu ---------------------------------------------------------------------------------
size_t rank_partials_count(int64_t G);
int launch_rank_scores(const float* scores, int64_t G, int group, int k, const int32_t* label_col, int32_t* rank,
                       int32_t* pos, float* sums, float* partials, cudaStream_t st);

// ---- sampler.cu ------------------------------------------------------------------------------
int launch_sample_negatives(const int64_t* rowptr, const int32_t* csr_items, int32_t num_items,
                            const int32_t* pos_users, const int32_t* pos_items, int64_t P, int64_t first_index,
                            int negs, uint64_t seed, uint64_t epoch, int32_t* out_users, int32_t* out_items,
                            float* out_labels, cudaStream_t st);

// ---- tc_selftest.cu ---------------------------------------------------------------------------
int launch_tc_probe(const float* raw_a, int n_words, int start_off, int lbo, int sbo, int a_mn, float* D,
                    cudaStream_t st);
int launch_tc_selftest(const float* A, const float* B, float* D, int N, int K, int a_mn, int b_mn, int three_x,
                       cudaStream_t st);

}  // namespace mr
