// TEST INFRASTRUCTURE ONLY — C wrapper around the UNMODIFIED reference
// (gleb-kov/wordpiece) so that pytest / bench.py can call the reference's own
// word_piece::fast::encode and word_piece::linear::encode through ctypes.
// It is compiled together with the reference sources where they lie under
// /root/reference by oracle/Makefile; the outputs go to oracle/_ref/ only.
// Nothing under wordpiece_b200/ may link or load this.
//
// Wrapped entry points (reference file:line):
//   word_piece::fast::encode(text, vocab_vector)       src/fast.cpp:154-157
//   word_piece::fast::encode(text_file, vocab_file)    src/fast.cpp:159-163
//   word_piece::linear::encode(text, vocab_vector)     src/linear.cpp:332-335
//   word_piece::fast::decode(vocab_file, ids)          src/fast.cpp:165-187
//   word_piece::fast::encodeExternal(...)              src/fast.cpp:189-220
//   utils::globalThreadPool(n)                         src/utils.cpp:25-28
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "src/utils.hpp"
#include "src/word_piece.hpp"

namespace {
thread_local std::string g_err;

int copy_out(const std::vector<int> &ids, int **out_ids, size_t *out_n) {
  *out_n = ids.size();
  *out_ids = static_cast<int *>(std::malloc(ids.size() * sizeof(int) + 1));
  if (*out_ids == nullptr) {
    g_err = "malloc failed";
    return 2;
  }
  if (!ids.empty()) std::memcpy(*out_ids, ids.data(), ids.size() * sizeof(int));
  return 0;
}

std::vector<std::string> make_vocab(const char *const *toks, const size_t *lens, size_t n) {
  std::vector<std::string> v;
  v.reserve(n);
  for (size_t i = 0; i < n; i++) v.emplace_back(toks[i], lens[i]);
  return v;
}
}  // namespace

extern "C" {

// Must be the first call in the process to take effect: the pool size is
// frozen by the first caller (utils.cpp:25-28).  Returns the actual size.
size_t wpref_init_threads(size_t n_threads) {
  return utils::globalThreadPool(n_threads).maxThreads();
}

const char *wpref_last_error(void) { return g_err.c_str(); }

void wpref_free(void *p) { std::free(p); }

// algo: 0 = fast, 1 = linear.  seconds_out (optional) receives the
// steady_clock time around the reference call alone.
int wpref_encode(int algo, const char *text, size_t n_bytes, const char *const *toks,
                 const size_t *lens, size_t n_vocab, int **out_ids, size_t *out_n,
                 double *seconds_out) {
  try {
    const std::string s(text, n_bytes);
    const std::vector<std::string> vocab = make_vocab(toks, lens, n_vocab);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<int> ids = algo == 0 ? word_piece::fast::encode(s, vocab)
                                     : word_piece::linear::encode(s, vocab);
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    return copy_out(ids, out_ids, out_n);
  } catch (const std::exception &e) {
    g_err = e.what();
    return 1;
  }
}

int wpref_encode_files(int algo, const char *text_file, const char *vocab_file, int **out_ids,
                       size_t *out_n, double *seconds_out) {
  try {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<int> ids = algo == 0 ? word_piece::fast::encode(std::string(text_file), std::string(vocab_file))
                                     : word_piece::linear::encode(std::string(text_file), std::string(vocab_file));
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    return copy_out(ids, out_ids, out_n);
  } catch (const std::exception &e) {
    g_err = e.what();
    return 1;
  }
}

int wpref_encode_external(const char *text_file, const char *vocab_file, const char *out_file,
                          size_t memory_limit, double *seconds_out) {
  try {
    const auto t0 = std::chrono::steady_clock::now();
    word_piece::fast::encodeExternal(text_file, vocab_file, out_file, memory_limit);
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 1;
  }
}

// Decoded tokens are returned as one malloc'd buffer of '\n'-joined strings.
int wpref_decode(const char *vocab_file, const int *ids, size_t n_ids, char **out, size_t *out_len) {
  try {
    std::vector<int> v(ids, ids + n_ids);
    std::vector<std::string> toks = word_piece::fast::decode(vocab_file, v);
    std::string joined;
    for (size_t i = 0; i < toks.size(); i++) {
      if (i) joined.push_back('\n');
      joined += toks[i];
    }
    *out_len = joined.size();
    *out = static_cast<char *>(std::malloc(joined.size() + 1));
    std::memcpy(*out, joined.data(), joined.size());
    (*out)[joined.size()] = 0;
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 1;
  }
}

}  // extern "C"
