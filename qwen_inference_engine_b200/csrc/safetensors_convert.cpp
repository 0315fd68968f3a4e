// safetensors_convert.cpp -- HF safetensors shards -> the reference's weights.bin + meta_data.txt.
//
// What parsed_tensors() does in the reference (/root/reference/layers/src/tensor_parser.cpp:31-129): for every
// shard, in the order given, read the 8-byte little-endian header length and the JSON header, walk its keys in
// byte-lexicographic order (nlohmann::json objects are std::map, :71), keep keys that start with "model." or
// "lm_" (:74,102), give each kept tensor the running byte range [global_offset, +size) (:94-99), derive
// layer / short_name from the key (:85-92; every "lm_" tensor is called "logits", :107) and print the records
// with operator<< (:19-28, :124-126).  The reference then expects weights.bin to hold those bytes back to back
// (its own copy loop, :118-121, is commented out and would copy whole data sections, which only matches the
// offsets when a shard's data order equals its key order and nothing is filtered).  This converter writes the
// file the offsets describe: tensor by tensor, from data_offsets of the shard, in the reference's order.
//
// Differences, all explicit: only BF16 tensors are accepted (the engine is bf16; anything else is refused, not
// converted); with tie_lm_head != 0 a checkpoint without an "lm_" tensor (Qwen2.5-0.5B/1.5B tie the embeddings)
// gets "lm_head.weight" = a second copy of model.embed_tokens.weight appended, because the forward indexes
// `logits` unconditionally (qwen_main.cu:230).  No third-party JSON library: the safetensors header is a flat
// object of {dtype, shape, data_offsets} records plus an optional "__metadata__" string map.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "checkpoint.h"

namespace qie {
namespace {

struct StEntry {
  std::string key, dtype;
  std::vector<size_t> shape;
  size_t off[2] = {0, 0};
  bool has_off = false;
};

// minimal JSON reader for the safetensors header
struct Json {
  const char* p;
  const char* e;
  std::string err;
  bool fail(const char* m) {
    if (err.empty()) err = m;
    return false;
  }
  void ws() {
    while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
  }
  bool lit(char c) {
    ws();
    if (p < e && *p == c) {
      ++p;
      return true;
    }
    return false;
  }
  bool str(std::string* out) {
    ws();
    if (p >= e || *p != '"') return fail("expected string");
    ++p;
    out->clear();
    while (p < e && *p != '"') {
      if (*p == '\\') {
        if (++p >= e) return fail("bad escape");
        switch (*p) {
          case 'n': out->push_back('\n'); break;
          case 't': out->push_back('\t'); break;
          case 'r': out->push_back('\r'); break;
          case 'b': out->push_back('\b'); break;
          case 'f': out->push_back('\f'); break;
          case 'u':  // keys/dtypes are ASCII; keep the escape verbatim
            out->append("\\u");
            break;
          default: out->push_back(*p); break;
        }
        ++p;
      } else {
        out->push_back(*p++);
      }
    }
    if (p >= e) return fail("unterminated string");
    ++p;
    return true;
  }
  bool num(size_t* out) {
    ws();
    if (p >= e || *p < '0' || *p > '9') return fail("expected unsigned integer");
    size_t v = 0;
    while (p < e && *p >= '0' && *p <= '9') v = v * 10 + (size_t)(*p++ - '0');
    *out = v;
    return true;
  }
  bool skip_value() {  // any JSON value
    ws();
    if (p >= e) return fail("unexpected end");
    if (*p == '"') {
      std::string s;
      return str(&s);
    }
    if (*p == '{' || *p == '[') {
      const char open = *p, close = open == '{' ? '}' : ']';
      ++p;
      if (lit(close)) return true;
      for (;;) {
        if (open == '{') {
          std::string k;
          if (!str(&k) || !lit(':')) return fail("bad object");
        }
        if (!skip_value()) return false;
        if (lit(',')) continue;
        if (lit(close)) return true;
        return fail("bad container");
      }
    }
    while (p < e && *p != ',' && *p != '}' && *p != ']') ++p;  // number / true / false / null
    return true;
  }
  bool entry(StEntry* t) {
    if (!lit('{')) return fail("tensor record must be an object");
    if (lit('}')) return true;
    for (;;) {
      std::string k;
      if (!str(&k) || !lit(':')) return fail("bad tensor record");
      if (k == "dtype") {
        if (!str(&t->dtype)) return false;
      } else if (k == "shape" || k == "data_offsets") {
        if (!lit('[')) return fail("expected array");
        std::vector<size_t> v;
        if (!lit(']')) {
          for (;;) {
            size_t x;
            if (!num(&x)) return false;
            v.push_back(x);
            if (lit(',')) continue;
            if (lit(']')) break;
            return fail("bad array");
          }
        }
        if (k == "shape") {
          t->shape = v;
        } else {
          if (v.size() != 2) return fail("data_offsets must have two entries");
          t->off[0] = v[0];
          t->off[1] = v[1];
          t->has_off = true;
        }
      } else if (!skip_value()) {
        return false;
      }
      if (lit(',')) continue;
      if (lit('}')) return true;
      return fail("bad tensor record");
    }
  }
  bool header(std::vector<StEntry>* out) {
    if (!lit('{')) return fail("header must be a JSON object");
    if (lit('}')) return true;
    for (;;) {
      StEntry t;
      if (!str(&t.key) || !lit(':')) return fail("bad header");
      if (t.key == "__metadata__") {
        if (!skip_value()) return false;
      } else {
        if (!entry(&t)) return false;
        out->push_back(std::move(t));
      }
      if (lit(',')) continue;
      if (lit('}')) return true;
      return fail("bad header");
    }
  }
};

bool copy_range(FILE* in, size_t pos, size_t n, FILE* out, std::vector<char>* buf) {
  if (fseeko(in, (off_t)pos, SEEK_SET) != 0) return false;
  while (n) {
    const size_t m = std::min(n, buf->size());
    if (fread(buf->data(), 1, m, in) != m) return false;
    if (fwrite(buf->data(), 1, m, out) != m) return false;
    n -= m;
  }
  return true;
}

}  // namespace

// returns "" on success, else the error text
std::string convert_safetensors(const std::vector<std::string>& shards, const std::string& meta_path,
                                const std::string& weights_path, bool tie_lm_head, size_t* total_bytes, int* n_tensors) {
  Checkpoint ck;
  FILE* fw = fopen(weights_path.c_str(), "wb");
  if (!fw) return "cannot write " + weights_path;
  std::vector<char> buf((size_t)8 << 20);
  size_t global_offset = 0;
  bool have_lm = false;
  std::string embed_shard;
  size_t embed_pos = 0, embed_bytes = 0;
  std::vector<size_t> embed_shape;
  std::string err;
  for (const std::string& path : shards) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) {
      err = "cannot open " + path;
      break;
    }
    uint64_t hlen = 0;
    if (fread(&hlen, 8, 1, f) != 1 || hlen == 0 || hlen > ((uint64_t)1 << 30)) {
      fclose(f);
      err = path + ": bad safetensors header length";
      break;
    }
    std::string js((size_t)hlen, '\0');
    if (fread(&js[0], 1, (size_t)hlen, f) != (size_t)hlen) {
      fclose(f);
      err = path + ": truncated header";
      break;
    }
    std::vector<StEntry> ents;
    Json j{js.data(), js.data() + js.size(), ""};
    if (!j.header(&ents)) {
      fclose(f);
      err = path + ": " + j.err;
      break;
    }
    // nlohmann::json iterates object keys in std::map order = byte-lexicographic
    std::sort(ents.begin(), ents.end(), [](const StEntry& a, const StEntry& b) { return a.key < b.key; });
    const size_t data0 = 8 + (size_t)hlen;
    for (const StEntry& t : ents) {
      const bool is_model = t.key.rfind("model.", 0) == 0, is_lm = t.key.rfind("lm_", 0) == 0;
      if (!is_model && !is_lm) continue;
      if (!t.has_off || t.off[1] < t.off[0]) {
        err = path + ": " + t.key + ": missing data_offsets";
        break;
      }
      if (t.dtype != "BF16") {
        err = path + ": " + t.key + ": dtype " + t.dtype + " (only BF16 checkpoints are supported)";
        break;
      }
      size_t elems = 1;
      for (size_t d : t.shape) elems *= d;
      const size_t bytes = t.off[1] - t.off[0];
      if (bytes != elems * 2) {
        err = path + ": " + t.key + ": data_offsets do not match the shape";
        break;
      }
      TensorInfo ti;
      ti.name = t.key;
      ti.shape = t.shape;
      if (is_lm) {
        ti.short_name = "logits";  // tensor_parser.cpp:107
        have_lm = true;
      } else {
        const size_t lp = t.key.find("layers.");
        if (lp != std::string::npos) {
          const size_t dot = t.key.find('.', lp + 7);
          if (dot == std::string::npos) {
            err = path + ": " + t.key + ": malformed layer key";
            break;
          }
          ti.layer = atoi(t.key.substr(lp + 7, dot - (lp + 7)).c_str());
          ti.short_name = t.key.substr(dot + 1);
        } else {
          ti.short_name = t.key.substr(6);
        }
      }
      ti.begin = global_offset;
      ti.end = global_offset + bytes;
      if (!copy_range(f, data0 + t.off[0], bytes, fw, &buf)) {
        err = path + ": " + t.key + ": read/write error";
        break;
      }
      if (t.key == "model.embed_tokens.weight") {
        embed_shard = path;
        embed_pos = data0 + t.off[0];
        embed_bytes = bytes;
        embed_shape = t.shape;
      }
      global_offset += bytes;
      ck.tensors.push_back(std::move(ti));
    }
    fclose(f);
    if (!err.empty()) break;
  }
  if (err.empty() && !have_lm) {
    if (!tie_lm_head) {
      err = "no lm_head tensor in the shards (tied embeddings?): pass tie_lm_head to append a copy of model.embed_tokens.weight";
    } else if (embed_shard.empty()) {
      err = "tie_lm_head: model.embed_tokens.weight not found";
    } else {
      FILE* f = fopen(embed_shard.c_str(), "rb");
      TensorInfo ti;
      ti.name = "lm_head.weight";
      ti.short_name = "logits";
      ti.shape = embed_shape;
      ti.begin = global_offset;
      ti.end = global_offset + embed_bytes;
      if (!f || !copy_range(f, embed_pos, embed_bytes, fw, &buf)) err = "tie_lm_head: cannot copy the embedding";
      if (f) fclose(f);
      global_offset += embed_bytes;
      ck.tensors.push_back(std::move(ti));
    }
  }
  if (fclose(fw) != 0 && err.empty()) err = "short write to " + weights_path;
  if (!err.empty()) return err;
  ck.total_bytes = global_offset;
  FILE* fm = fopen(meta_path.c_str(), "w");
  if (!fm) return "cannot write " + meta_path;
  write_meta(ck, fm);
  fclose(fm);
  if (total_bytes) *total_bytes = global_offset;
  if (n_tensors) *n_tensors = (int)ck.tensors.size();
  return "";
}

}  // namespace qie
