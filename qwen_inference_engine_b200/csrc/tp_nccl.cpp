// tp_nccl.cpp -- see tp_nccl.h.  NCCL is dlopen'ed so that single-GPU users of libqie_b200.so do
// not need it; the symbols are the stable C API of NCCL 2.x.
#include "tp_nccl.h"

#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace qie {

bool tp_plan(const qie_config& c, int rank, int size, TpPlan* out) {
  if (size < 1 || rank < 0 || rank >= size) return false;
  if (c.inter % size || c.n_q % c.n_kv) return false;
  TpPlan p;
  const int group = c.n_q / c.n_kv;  // query heads per kv head
  if (c.n_kv % size == 0) {
    // whole kv heads (and their query groups) per rank
    p.n_kv = c.n_kv / size;
    p.n_q = p.n_kv * group;
    p.kv_row0 = rank * p.n_kv * c.head_dim;
    p.q_row0 = rank * p.n_q * c.head_dim;
  } else if (size % c.n_kv == 0) {
    // more ranks than kv heads (Qwen2.5-7B: 28 q / 4 kv heads over 8 GPUs, SURVEY 7 "hard parts"): `rep` ranks share one
    // kv head -- each keeps its own copy of that head's K/V cache -- and split its query group between them as evenly
    // as it goes (7 heads over 2 ranks: 4 + 3).  No padded heads: o_proj takes the columns of the heads a rank owns.
    const int rep = size / c.n_kv, kvh = rank / rep, sub = rank % rep;
    if (group < rep) return false;
    const int base = group / rep, extra = group % rep;
    p.n_kv = 1;
    p.n_q = base + (sub < extra ? 1 : 0);
    p.kv_row0 = kvh * c.head_dim;
    p.q_row0 = (kvh * group + sub * base + std::min(sub, extra)) * c.head_dim;
  } else {
    return false;
  }
  p.inter = c.inter / size;
  if (p.inter & 7) return false;
  // vocabulary shards start at multiples of 256 so that the low byte of a token index (the
  // sampler's tie-break key, logit_decode.cu:15-33) is the same in the shard and in the full row
  const int chunks = (c.vocab + 255) / 256, per = (chunks + size - 1) / size;
  p.vocab0 = std::min(c.vocab, rank * per * 256);
  p.vocab = std::min(c.vocab, (rank + 1) * per * 256) - p.vocab0;
  if (p.vocab <= 0 || (p.vocab & 1)) return false;
  p.inter0 = rank * p.inter;
  *out = p;
  return true;
}

namespace {
struct NcclUniqueId {
  char internal[128];
};
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(void**, int, NcclUniqueId, int);
typedef int (*CommDestroyFn)(void*);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*GetErrorStringFn)(int);

struct Nccl {
  void* lib = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  AllReduceFn all_reduce = nullptr;
  AllGatherFn all_gather = nullptr;
  GetErrorStringFn error_string = nullptr;
};
Nccl g_nccl;
std::once_flag g_once;

const Nccl* nccl() {
  std::call_once(g_once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return;
    g_nccl.get_unique_id = (GetUniqueIdFn)dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.comm_init_rank = (CommInitRankFn)dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.comm_destroy = (CommDestroyFn)dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.all_reduce = (AllReduceFn)dlsym(g_nccl.lib, "ncclAllReduce");
    g_nccl.all_gather = (AllGatherFn)dlsym(g_nccl.lib, "ncclAllGather");
    g_nccl.error_string = (GetErrorStringFn)dlsym(g_nccl.lib, "ncclGetErrorString");
  });
  if (!g_nccl.lib || !g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.all_reduce || !g_nccl.all_gather) return nullptr;
  return &g_nccl;
}
// ncclDataType_t / ncclRedOp_t values of the NCCL 2.x ABI
constexpr int kNcclInt8 = 0, kNcclFloat32 = 7, kNcclBfloat16 = 9, kNcclSum = 0;
}  // namespace

int tp_unique_id(void* out128, char* err, size_t errlen) {
  const Nccl* n = nccl();
  if (!n) {
    snprintf(err, errlen, "libnccl.so.2 not found (tensor parallel needs NCCL)");
    return -1;
  }
  NcclUniqueId id;
  int rc = n->get_unique_id(&id);
  if (rc != 0) {
    snprintf(err, errlen, "ncclGetUniqueId: %s", n->error_string ? n->error_string(rc) : "error");
    return -1;
  }
  memcpy(out128, &id, 128);
  return 0;
}

int tp_comm_init(TpComm* t, const void* id128, int rank, int size, char* err, size_t errlen) {
  const Nccl* n = nccl();
  if (!n) {
    snprintf(err, errlen, "libnccl.so.2 not found (tensor parallel needs NCCL)");
    return -1;
  }
  NcclUniqueId id;
  memcpy(&id, id128, 128);
  int rc = n->comm_init_rank(&t->comm, size, id, rank);
  if (rc != 0) {
    snprintf(err, errlen, "ncclCommInitRank: %s", n->error_string ? n->error_string(rc) : "error");
    t->comm = nullptr;
    return -1;
  }
  t->rank = rank;
  t->size = size;
  return 0;
}

void tp_comm_destroy(TpComm* t) {
  const Nccl* n = nccl();
  if (n && t->comm && n->comm_destroy) n->comm_destroy(t->comm);
  t->comm = nullptr;
}

cudaError_t tp_allreduce_bf16(TpComm* t, void* buf, size_t nelem, cudaStream_t st) {
  const Nccl* n = nccl();
  if (!n || !t->comm) return cudaErrorNotReady;
  return n->all_reduce(buf, buf, nelem, kNcclBfloat16, kNcclSum, t->comm, st) == 0 ? cudaSuccess : cudaErrorUnknown;
}

cudaError_t tp_allreduce_f32(TpComm* t, float* buf, size_t nelem, cudaStream_t st) {
  const Nccl* n = nccl();
  if (!n || !t->comm) return cudaErrorNotReady;
  return n->all_reduce(buf, buf, nelem, kNcclFloat32, kNcclSum, t->comm, st) == 0 ? cudaSuccess : cudaErrorUnknown;
}

cudaError_t tp_allgather(TpComm* t, const void* send, void* recv, size_t bytes, cudaStream_t st) {
  const Nccl* n = nccl();
  if (!n || !t->comm) return cudaErrorNotReady;
  return n->all_gather(send, recv, bytes, kNcclInt8, t->comm, st) == 0 ? cudaSuccess : cudaErrorUnknown;
}

}  // namespace qie
