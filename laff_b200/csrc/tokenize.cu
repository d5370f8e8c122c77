// Host-side half of the text front-end: the reference tokenises every caption with a Python regex and
// looks every word up in Python dicts, once per encoder (textlib.py:27-47, txt2vec.py:56-63, :97-104, :121-128).  At
// 10 000 queries per step that is ~130 ms of interpreter time — twice the GPU's whole ranking step — so the batch
// version lives here: one pass over the UTF-8 bytes per caption, an open-addressing hash table per vocabulary, CSR out.
//   TextTool.tokenize(clean=True, language='en'): replace every char outside [A-Za-z0-9] by a space, strip, lower, split
//   == maximal runs of ASCII alphanumerics, lower-cased (a non-ASCII character is one or more non-alnum bytes: a separator).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "host_util.cuh"

struct laff_vocab {
  struct Slot {
    uint64_t hash;
    uint32_t off, len;
    int32_t id;  // -1 = empty
  };
  std::vector<Slot> slots;
  std::string blob;
  uint64_t mask = 0;

  static uint64_t fnv(const char* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) h = (h ^ static_cast<unsigned char>(p[i])) * 1099511628211ull;
    return h ? h : 1;
  }
  void insert(const char* p, uint32_t n, int32_t id) {
    const uint64_t h = fnv(p, n);
    for (uint64_t i = h & mask;; i = (i + 1) & mask) {
      Slot& s = slots[i];
      if (s.id < 0) {
        s.hash = h;
        s.off = static_cast<uint32_t>(blob.size());
        s.len = n;
        s.id = id;
        blob.append(p, n);
        return;
      }
      if (s.hash == h && s.len == n && std::memcmp(blob.data() + s.off, p, n) == 0) {
        s.id = id;  // a repeated word keeps its last id, like dict(zip(names, range(n))) (bigfile.py:23)
        return;
      }
    }
  }
  int32_t find(const char* p, uint32_t n) const {
    const uint64_t h = fnv(p, n);
    for (uint64_t i = h & mask;; i = (i + 1) & mask) {
      const Slot& s = slots[i];
      if (s.id < 0) return -1;
      if (s.hash == h && s.len == n && std::memcmp(blob.data() + s.off, p, n) == 0) return s.id;
    }
  }
};

extern "C" laff_vocab* laff_vocab_create(const char* words_blob, const long long* offsets, const int32_t* ids, int n_words) {
  if (!words_blob || !offsets || n_words < 0) {
    laff::set_error("laff_vocab_create: bad arguments");
    return nullptr;
  }
  laff_vocab* v = new laff_vocab();
  uint64_t cap = 16;
  while (cap < static_cast<uint64_t>(n_words) * 2 + 2) cap <<= 1;
  v->slots.assign(cap, laff_vocab::Slot{0, 0, 0, -1});
  v->mask = cap - 1;
  for (int i = 0; i < n_words; ++i) {
    const long long a = offsets[i], b = offsets[i + 1];
    v->insert(words_blob + a, static_cast<uint32_t>(b - a), ids ? ids[i] : i);
  }
  return v;
}

extern "C" void laff_vocab_destroy(laff_vocab* v) { delete v; }

extern "C" long long laff_tokenize_lookup(const char* text_blob, const long long* cap_offsets, int n_caps, const laff_vocab* vocab,
                                          const laff_vocab* stopwords, int mode, int unk_id, int start_id, int end_id,
                                          long long* out_offsets, int32_t* out_ids, long long capacity) {
  LAFF_REQUIRE(text_blob && cap_offsets && vocab && out_offsets && (out_ids || capacity == 0) && n_caps >= 0 && mode >= 0 && mode <= 2,
               LAFF_EINVAL, "laff_tokenize_lookup: bad arguments");
  long long n = 0;
  std::string tok;
  std::vector<int32_t> uniq;
  out_offsets[0] = 0;
  for (int c = 0; c < n_caps; ++c) {
    const char* p = text_blob + cap_offsets[c];
    const char* e = text_blob + cap_offsets[c + 1];
    auto emit = [&](int32_t id) {
      if (n < capacity) out_ids[n] = id;
      ++n;
    };
    uniq.clear();
    if (mode == 0 && start_id >= 0) emit(start_id);
    while (p < e) {
      while (p < e && !((*p >= '0' && *p <= '9') || (*p >= 'a' && *p <= 'z') || (*p >= 'A' && *p <= 'Z'))) ++p;
      tok.clear();
      while (p < e && ((*p >= '0' && *p <= '9') || (*p >= 'a' && *p <= 'z') || (*p >= 'A' && *p <= 'Z'))) {
        tok.push_back((*p >= 'A' && *p <= 'Z') ? static_cast<char>(*p + 32) : *p);
        ++p;
      }
      if (tok.empty()) continue;
      if (stopwords && stopwords->find(tok.data(), static_cast<uint32_t>(tok.size())) >= 0) continue;
      const int32_t id = vocab->find(tok.data(), static_cast<uint32_t>(tok.size()));
      if (mode == 0) {
        emit(id >= 0 ? id : unk_id);  // IndexVec: '<unk>' for out-of-vocabulary words
      } else if (id >= 0) {
        if (mode == 1) emit(id);      // BowVec: every in-vocabulary token, in order
        else uniq.push_back(id);      // W2Vec: distinct in-vocabulary words, by position in the vector file
      }
    }
    if (mode == 0 && end_id >= 0) emit(end_id);
    if (mode == 2) {
      std::sort(uniq.begin(), uniq.end());
      uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
      for (int32_t id : uniq) emit(id);
    }
    out_offsets[c + 1] = n;
  }
  return n;
}
