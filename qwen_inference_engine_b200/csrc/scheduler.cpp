// scheduler.cpp -- continuous batching on top of the driver-level C ABI.
//
// The reference only sketches this: `enum State {prefill, decode}` and one `batch_metadata` per sequence
// (/root/reference/layers/include/iengine.cuh:23-37), a second sequence commented out in main()
// (src/iengine.cu:369-373,448-452), the EOS test `generated_token == 151645` at the top of the decode branch
// (src/qwen_main.cu:257) and a decode loop that never ends (iengine.cu:422).  Here requests queue up, are
// admitted FIFO while sequence slots and KV pages allow (prefill = llm() in state prefill), all running requests
// advance together with ONE batched decode step per iteration (llm() in state decode; <= 64 rows run as the
// persistent kernel), and a request that samples EOS or reaches its token budget leaves the batch at once: its
// pages go back to the pool (free_page_list, iengine.cu:98-109) and the next waiting request takes the slot.
// Reference-order numerics are batch invariant, so a request's tokens do not depend on who shares its steps.
// Host logic only: everything goes through qie_seq_new / qie_prefill / qie_decode_step / qie_seq_free.
#include <stdint.h>

#include <deque>
#include <vector>

#include "../../include/qie_b200.h"

struct qie_scheduler {
  qie_engine* eng = nullptr;
  int max_running = 0, eos = -1;
  int page_size = 16, total_pages = 0;
  int context = 0, vocab = 0, max_pages_per_seq = 0;  // engine limits a request is checked against when it is submitted
  struct Req {
    std::vector<int32_t> prompt, out;
    int max_new = 0, seq = -1, reserved = 0;
    bool running = false, done = false;
    int error = 0;  // QIE_E* code of the call that failed this request (0 = none)
  };
  std::vector<Req> reqs;
  std::deque<int> waiting;
  std::vector<int> running;
  int reserved_pages = 0;
  long steps = 0, decode_rows = 0, prefills = 0;
};

namespace {
int pages_for(const qie_scheduler* s, int positions) { return (positions + s->page_size - 1) / s->page_size; }

void finish(qie_scheduler* s, int id) {
  qie_scheduler::Req& r = s->reqs[id];
  if (r.seq >= 0) qie_seq_free(s->eng, r.seq);
  s->reserved_pages -= r.reserved;
  r.seq = -1;
  r.reserved = 0;
  r.running = false;
  r.done = true;
}
}  // namespace

extern "C" {

int qie_sched_create(qie_engine* e, int max_running, int eos_token, qie_scheduler** out) {
  if (!e || !out || max_running <= 0) return QIE_EINVAL;
  qie_kv_view v;
  int rc = qie_engine_kv_view(e, &v);
  if (rc) return rc;
  qie_scheduler* s = new qie_scheduler();
  s->eng = e;
  s->max_running = max_running;
  s->eos = eos_token;
  s->page_size = v.page_size;
  s->total_pages = v.n_pages;
  qie_config cfg;
  int max_rows = 0, max_pages_per_seq = 0, max_seqs = 0;
  if ((rc = qie_engine_get_config(e, &cfg)) || (rc = qie_engine_limits(e, &max_rows, &max_pages_per_seq, &max_seqs))) {
    delete s;
    return rc;
  }
  s->context = cfg.context;
  s->vocab = cfg.vocab;
  s->max_pages_per_seq = max_pages_per_seq;
  // never more rows in a decode step than the engine takes (max_batch_tokens / logits rows) or sequences it can hold
  if (s->max_running > max_rows) s->max_running = max_rows;
  if (s->max_running > max_seqs) s->max_running = max_seqs;
  *out = s;
  return QIE_OK;
}

void qie_sched_destroy(qie_scheduler* s) {
  if (!s) return;
  for (size_t i = 0; i < s->reqs.size(); ++i)
    if (s->reqs[i].running) finish(s, (int)i);
  delete s;
}

int qie_sched_submit(qie_scheduler* s, const int32_t* ids, int n, int max_new_tokens, int* request_id) {
  if (!s || !ids || n <= 0 || max_new_tokens <= 0 || !request_id) return QIE_EINVAL;
  // a request that could never run is refused HERE: once queued it would sit at the head of the FIFO and starve
  // everybody behind it (prefill / decode fail the same way every step)
  if (pages_for(s, n + max_new_tokens) > s->total_pages) return QIE_ENOMEM;  // could never be admitted
  if (n + max_new_tokens > s->context || pages_for(s, n + max_new_tokens) > s->max_pages_per_seq) return QIE_EINVAL;
  for (int i = 0; i < n; ++i)
    if (ids[i] < 0 || ids[i] >= s->vocab) return QIE_EINVAL;
  qie_scheduler::Req r;
  r.prompt.assign(ids, ids + n);
  r.max_new = max_new_tokens;
  s->reqs.push_back(std::move(r));
  *request_id = (int)s->reqs.size() - 1;
  s->waiting.push_back(*request_id);
  return QIE_OK;
}

// One iteration: admit (prefill) while slots and pages allow, then one decode step for everybody who is running.
// Returns the number of unfinished requests (waiting + running), or a negative QIE_E* code.
int qie_sched_step(qie_scheduler* s) {
  if (!s) return QIE_EINVAL;
  ++s->steps;
  // ---- admission, FIFO; pages for prompt + budget are reserved up front so a decode step can never run dry
  while (!s->waiting.empty() && (int)s->running.size() < s->max_running) {
    const int id = s->waiting.front();
    qie_scheduler::Req& r = s->reqs[id];
    const int need = pages_for(s, (int)r.prompt.size() + r.max_new);
    if (s->reserved_pages + need > s->total_pages) break;  // head of the queue waits for pages: no overtaking
    int seq = -1;
    int rc = qie_seq_new(s->eng, &seq);
    if (rc == QIE_ENOMEM) break;  // no sequence slot in the engine right now
    if (rc) return rc;
    int32_t tok = 0;
    rc = qie_prefill(s->eng, seq, r.prompt.data(), (int)r.prompt.size(), &tok);
    if (rc) {
      // this request fails (and leaves the queue); the others go on.  Only a device failure stops the scheduler.
      qie_seq_free(s->eng, seq);
      s->waiting.pop_front();
      r.error = rc;
      r.done = true;
      if (rc == QIE_ECUDA) return rc;
      continue;
    }
    ++s->prefills;
    s->waiting.pop_front();
    r.seq = seq;
    r.reserved = need;
    s->reserved_pages += need;
    r.running = true;
    r.out.push_back(tok);
    if (tok == s->eos || (int)r.out.size() >= r.max_new)
      finish(s, id);
    else
      s->running.push_back(id);
  }
  // ---- one batched decode step
  if (!s->running.empty()) {
    const int n = (int)s->running.size();
    std::vector<int> seqs(n);
    std::vector<int32_t> in(n), out(n);
    for (int i = 0; i < n; ++i) {
      const qie_scheduler::Req& r = s->reqs[s->running[i]];
      seqs[i] = r.seq;
      in[i] = r.out.back();
    }
    std::vector<int> row_rc(n, 0);
    int rc = qie_decode_step(s->eng, seqs.data(), in.data(), n, out.data());
    if (rc == QIE_ECUDA) return rc;
    if (rc) {
      // the batched call refused (no state was changed: arguments are checked before anything is launched): find
      // the offending request(s) by stepping every row on its own; they are failed, the others advance as usual
      for (int i = 0; i < n; ++i) {
        row_rc[i] = qie_decode_step(s->eng, &seqs[i], &in[i], 1, &out[i]);
        if (row_rc[i] == QIE_ECUDA) return row_rc[i];
      }
    }
    s->decode_rows += n;
    std::vector<int> still;
    for (int i = 0; i < n; ++i) {
      const int id = s->running[i];
      qie_scheduler::Req& r = s->reqs[id];
      if (row_rc[i]) {
        r.error = row_rc[i];
        finish(s, id);
        continue;
      }
      r.out.push_back(out[i]);
      if (out[i] == s->eos || (int)r.out.size() >= r.max_new)
        finish(s, id);
      else
        still.push_back(id);
    }
    s->running.swap(still);
  }
  return (int)(s->waiting.size() + s->running.size());
}

int qie_sched_result(const qie_scheduler* s, int request_id, int32_t* out, int max_tokens, int* finished) {
  if (!s || request_id < 0 || request_id >= (int)s->reqs.size()) return QIE_EINVAL;
  const qie_scheduler::Req& r = s->reqs[request_id];
  const int n = (int)r.out.size();
  if (out)
    for (int i = 0; i < n && i < max_tokens; ++i) out[i] = r.out[i];
  if (finished) *finished = r.error ? r.error : (r.done ? 1 : 0);  // negative: the QIE_E* code that failed the request
  return n;
}

int qie_sched_stats(const qie_scheduler* s, long* steps, long* decode_rows, long* prefills, int* running, int* waiting) {
  if (!s) return QIE_EINVAL;
  if (steps) *steps = s->steps;
  if (decode_rows) *decode_rows = s->decode_rows;
  if (prefills) *prefills = s->prefills;
  if (running) *running = (int)s->running.size();
  if (waiting) *waiting = (int)s->waiting.size();
  return QIE_OK;
}

}  // extern "C"
