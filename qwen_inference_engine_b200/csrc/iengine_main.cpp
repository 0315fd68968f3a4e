// iengine -- host driver, the B200 counterpart of the reference's main()
// (/root/reference/layers/src/iengine.cu:226-482): load weights.bin + meta_data.txt, create
// sequences, prefill, then decode token by token.  Differences that are the point:
// prompts come from the command line instead of a literal (iengine.cu:325), generation stops
// at EOS 151645 (the reference checks it at qwen_main.cu:257 but main() loops forever on
// getchar(), iengine.cu:422) or after --n tokens, several sequences decode as one batch, and
// errors are error codes rather than token 0.
//   iengine --meta model_files/meta_data.txt --weights weights.bin --prompt 151643,785,4767 --n 128
//   iengine --synthetic qwen2.5-0.5b --prompt 151643,785 --prompt 151643,3639 --n 32 --topk 50
//   iengine --convert model-00001-of-00002.safetensors --convert model-00002-of-00002.safetensors
//           --meta model_files/meta_data.txt --weights weights.bin [--tie-lm-head]     (host only, then exits)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/qie_b200.h"

static void die(const char* what) {
  fprintf(stderr, "iengine: %s: %s\n", what, qie_last_error());
  exit(1);
}

static bool arch_config(const std::string& a, qie_config* c) {
  struct { const char* n; qie_config c; } t[] = {
      {"qwen2.5-0.5b", {896, 4864, 24, 14, 2, 64, 151936, 32786}},
      {"qwen2.5-1.5b", {1536, 8960, 28, 12, 2, 128, 151936, 32786}},
      {"qwen2.5-7b", {3584, 18944, 28, 28, 4, 128, 152064, 32786}},
      {"tiny", {128, 256, 2, 4, 2, 64, 512, 32786}},
  };
  for (auto& e : t)
    if (a == e.n) {
      *c = e.c;
      return true;
    }
  return false;
}

int main(int argc, char** argv) {
  std::string meta, weights, synthetic;
  std::vector<const char*> shards;  // --convert: safetensors shards, in order (parsed_tensors(), tensor_parser.cpp:31-129)
  int tie_lm_head = 0;
  std::vector<std::vector<int32_t>> prompts;
  int n_new = 32, topk = 1, device = 0, fast = 0;
  const int EOS = 151645;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "--meta") meta = next();
    else if (a == "--weights") weights = next();
    else if (a == "--synthetic") synthetic = next();
    else if (a == "--n") n_new = atoi(next());
    else if (a == "--topk") topk = atoi(next());
    else if (a == "--device") device = atoi(next());
    else if (a == "--fast") fast = 1;
    else if (a == "--convert") shards.push_back(next());
    else if (a == "--tie-lm-head") tie_lm_head = 1;
    else if (a == "--prompt") {
      std::vector<int32_t> p;
      for (char* tok = strtok(const_cast<char*>(next()), ","); tok; tok = strtok(nullptr, ",")) p.push_back(atoi(tok));
      if (!p.empty()) prompts.push_back(p);
    } else {
      fprintf(stderr, "unknown argument %s\n", a.c_str());
      return 2;
    }
  }
  if (!shards.empty()) {  // checkpoint conversion only: no GPU needed
    if (meta.empty() || weights.empty()) {
      fprintf(stderr, "iengine: --convert needs --meta and --weights (output paths)\n");
      return 2;
    }
    size_t total = 0;
    int n = 0;
    if (qie_convert_safetensors(shards.data(), (int)shards.size(), meta.c_str(), weights.c_str(), tie_lm_head, &total, &n)) die("convert");
    printf("iengine: %d tensors, %zu bytes -> %s + %s\n", n, total, weights.c_str(), meta.c_str());
    return 0;
  }
  if (prompts.empty()) prompts.push_back({151643, 785, 4767, 315, 279, 3639, 4180, 374});  // iengine.cu:325
  qie_engine_opts o;
  qie_engine_opts_default(&o);
  o.device = device;
  o.max_seqs = (int)prompts.size() + 1;
  o.numerics = fast ? QIE_NUMERICS_FAST : QIE_NUMERICS_REFERENCE_ORDER;
  qie_engine* e = nullptr;
  auto t0 = std::chrono::steady_clock::now();
  if (!synthetic.empty()) {
    qie_config c;
    if (!arch_config(synthetic, &c)) {
      fprintf(stderr, "unknown --synthetic arch\n");
      return 2;
    }
    if (qie_engine_create_synthetic(&c, 1234, &o, &e)) die("engine_create_synthetic");
  } else {
    if (meta.empty() || weights.empty()) {
      fprintf(stderr, "need --meta and --weights (or --synthetic ARCH)\n");
      return 2;
    }
    if (qie_engine_create(meta.c_str(), weights.c_str(), &o, &e)) die("engine_create");
  }
  if (topk > 1) qie_engine_set_sampling(e, topk, 1.0f, 0.7f, 1234, 1);  // qwen_main.cu:241,381-388
  qie_config c;
  qie_engine_get_config(e, &c);
  fprintf(stderr, "model: H=%d I=%d L=%d heads=%d/%d hd=%d V=%d  (load %.2fs)\n", c.hidden, c.inter, c.layers, c.n_q, c.n_kv,
          c.head_dim, c.vocab, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  const int B = (int)prompts.size();
  std::vector<int> seqs(B);
  std::vector<int32_t> cur(B);
  std::vector<std::vector<int32_t>> out(B);
  std::vector<bool> done(B, false);
  for (int i = 0; i < B; ++i) {
    if (qie_seq_new(e, &seqs[i])) die("seq_new");
    if (qie_prefill(e, seqs[i], prompts[i].data(), (int)prompts[i].size(), &cur[i])) die("prefill");
    out[i].push_back(cur[i]);
    done[i] = cur[i] == EOS;
  }
  t0 = std::chrono::steady_clock::now();
  int steps = 0;
  for (int s = 1; s < n_new; ++s) {
    std::vector<int> live;
    std::vector<int32_t> in;
    for (int i = 0; i < B; ++i)
      if (!done[i]) {
        live.push_back(seqs[i]);
        in.push_back(cur[i]);
      }
    if (live.empty()) break;
    std::vector<int32_t> nxt(live.size());
    if (qie_decode_step(e, live.data(), in.data(), (int)live.size(), nxt.data())) die("decode_step");
    ++steps;
    for (int i = 0, j = 0; i < B; ++i)
      if (!done[i]) {
        cur[i] = nxt[j++];
        out[i].push_back(cur[i]);
        if (cur[i] == EOS) done[i] = true;
      }
  }
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (int i = 0; i < B; ++i) {
    printf("seq %d:", i);
    for (int32_t t : out[i]) printf(" %d", t);
    printf("\n");
  }
  if (steps) fprintf(stderr, "%d decode steps x %d sequences in %.3fs (%.1f tokens/s)\n", steps, B, dt, steps * B / dt);
  for (int i = 0; i < B; ++i) qie_seq_free(e, seqs[i]);
  qie_engine_destroy(e);
  return 0;
}
