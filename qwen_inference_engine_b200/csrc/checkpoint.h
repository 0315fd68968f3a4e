// checkpoint.h -- the reference's on-disk model format: weights.bin (raw bf16 tensors
// back to back) + model_files/meta_data.txt (the text operator<< of
// /root/reference/layers/src/tensor_parser.cpp:19-28 emits, 6 lines per tensor).
// The reference only WRITES that text (it re-parses safetensors at start-up,
// tensor_parser.cpp:31-129); this loader READS it, addressing tensors by their
// [begin,end) byte offsets and never by file order.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/qie_b200.h"

namespace qie {

struct TensorInfo {
  std::string name;        // "model.layers.0.self_attn.k_proj.weight"
  std::string short_name;  // "self_attn.k_proj.weight" / "embed_tokens.weight" / "logits"
  int layer = -1;
  std::vector<size_t> shape;
  size_t begin = 0, end = 0;  // byte range in weights.bin
  int kind = 0;               // synthetic generator class: 0 matrix, 1 norm vector
};

struct Checkpoint {
  std::vector<TensorInfo> tensors;
  size_t total_bytes = 0;
  // short_name -> per-layer index into `tensors` (globals use slot 0), the same shape as
  // the reference's TensorTable (include/utils.hh:12, tensor_parser.cpp:132-165)
  std::unordered_map<std::string, std::vector<int>> index;

  const TensorInfo* find(const std::string& short_name, int layer) const;
  void build_index();
};

// parse meta_data.txt; returns false and sets err on malformed input
bool parse_meta(const std::string& path, Checkpoint* out, std::string* err);
// layout of a synthetic checkpoint for cfg (single shard, byte-lexicographic key order as
// nlohmann::json iterates, tensor_parser.cpp:71)
Checkpoint synth_layout(const qie_config& cfg);
// model shape from tensor shapes (replaces the literals of src/utills.cu:8-16)
bool derive_config(const Checkpoint& ck, int head_dim_hint, int context, qie_config* cfg, std::string* err);
void write_meta(const Checkpoint& ck, FILE* f);
// HF safetensors shards -> weights.bin + meta_data.txt in the reference's order (safetensors_convert.cpp);
// returns "" or the error text
std::string convert_safetensors(const std::vector<std::string>& shards, const std::string& meta_path,
                                const std::string& weights_path, bool tie_lm_head, size_t* total_bytes, int* n_tensors);
// host value of global element g (twin of synth_fill_kernel)
uint16_t synth_value(uint64_t seed, uint64_t g, int kind);

}  // namespace qie
