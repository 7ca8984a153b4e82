// Drop-in C++ entry points of the B200-native fast WordPiece encoder.
//
// Same names, namespaces, argument meaning, return types and error behaviour as
// the `fast` half of gleb-kov/wordpiece's public header (src/word_piece.hpp:23-34),
// so code written against the reference (tests/runner.cpp:38,53,
// tests/tests.cpp:86,95) recompiles against this header and links
// libwordpiece_b200.so instead of libword_piece.a.  Every call runs the sm_100a
// kernels through the C ABI in wordpiece_b200.h; there is no CPU path.
//
// The `linear` half of the reference header (suffix-array encoder,
// src/word_piece.hpp:10-21) is out of scope for this library and not declared.
#pragma once

#include <cstddef>
#include <string>
#include <vector>

namespace word_piece {
namespace fast {

// fast.cpp:154-157.  `vocab[i]` is the token with id i; text is UTF-8.
// Throws std::runtime_error("Vocab word is empty") like utils.cpp:99-101, and
// std::runtime_error on any CUDA failure.
std::vector<int> encode(const std::string &text, const std::vector<std::string> &vocab);

// fast.cpp:159-163.  Text and vocabulary (one token per line) read from files.
std::vector<int> encode(const std::string &text_file, const std::string &vocab_file);

// fast.cpp:165-187.  Token text for each id, "##" re-added for continuation tokens.
std::vector<std::string> decode(const std::string vocab_file, const std::vector<int> &ids);

// fast.cpp:189-220.  Streams text_file through the encoder in batches of at most
// memory_limit / 2 bytes cut after a space and appends "id id id " to out_file.
void encodeExternal(const std::string &text_file,
                    const std::string &vocab_file,
                    const std::string &out_file,
                    size_t memory_limit);

// ---- Extension, not in the reference: a vocabulary that stays on the GPU.
// The reference's entry points are stateless — every call re-parses the vocabulary and rebuilds both hash maps
// (fast.cpp:154-157, :21-35).  The functions above hide that behind a content-keyed cache; an Encoder makes
// the lifetime explicit and adds the call the reference lacks: many short texts in one launch.
class Encoder {
 public:
  explicit Encoder(const std::vector<std::string> &vocab, int device = 0);  // throws like encode(text, vocab)
  static Encoder fromFile(const std::string &vocab_file, int device = 0);   // utils.cpp:123-137 line rules
  ~Encoder();
  Encoder(Encoder &&other) noexcept;
  Encoder &operator=(Encoder &&other) noexcept;
  Encoder(const Encoder &) = delete;
  Encoder &operator=(const Encoder &) = delete;

  // == fast::encode(text, vocab) for the vocabulary this object holds
  std::vector<int> encode(const std::string &text) const;
  // Every text encoded as encode(texts[i]) would; ids of text i are ids[offsets[i] .. offsets[i + 1]).
  // One host->device copy, one pass of the kernels over the whole batch, one copy back.
  void encodeBatch(const std::vector<std::string> &texts, std::vector<int> &ids, std::vector<size_t> &offsets) const;
  std::vector<std::vector<int>> encodeBatch(const std::vector<std::string> &texts) const;
  // == fast::decode for this vocabulary
  std::vector<std::string> decode(const std::vector<int> &ids) const;
  size_t vocabSize() const;

 private:
  explicit Encoder(void *handle) : handle_(handle) {}
  void *handle_ = nullptr;  // wp_vocab* (wordpiece_b200.h)
};

}  // namespace fast
}  // namespace word_piece
