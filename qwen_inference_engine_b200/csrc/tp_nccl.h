// tp_nccl.h -- tensor-parallel plumbing: NCCL (loaded at run time, only when tp_size > 1) for the
// all-reduce after o_proj / down_proj and the all-gather of arg-max candidates, plus the shard plan.
// The reference has no multi-GPU path at all; BASELINE configs[4] asks for
// Qwen2.5-7B-arch tensor parallel over NVLink with an NCCL all-reduce after o_proj and down_proj.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/qie_b200.h"

namespace qie {

// what rank `rank` of `size` owns (heads, intermediate columns, vocabulary rows)
struct TpPlan {
  int n_q, n_kv, inter, vocab;              // local sizes
  int q_row0, kv_row0, inter0, vocab0;      // first local row / column in the full tensors
};
// returns false if the shape does not divide
bool tp_plan(const qie_config& c, int rank, int size, TpPlan* out);

struct TpComm {
  void* comm = nullptr;
  int rank = 0, size = 1;
};
// 128-byte NCCL unique id (rank 0 creates it, everybody passes it to tp_comm_init)
int tp_unique_id(void* out128, char* err, size_t errlen);
int tp_comm_init(TpComm* t, const void* id128, int rank, int size, char* err, size_t errlen);
void tp_comm_destroy(TpComm* t);
// in-place sum over ranks of n bf16 / fp32 elements; all-gather of `bytes` bytes per rank
cudaError_t tp_allreduce_bf16(TpComm* t, void* buf, size_t n, cudaStream_t st);
cudaError_t tp_allreduce_f32(TpComm* t, float* buf, size_t n, cudaStream_t st);
cudaError_t tp_allgather(TpComm* t, const void* send, void* recv, size_t bytes, cudaStream_t st);

}  // namespace qie
