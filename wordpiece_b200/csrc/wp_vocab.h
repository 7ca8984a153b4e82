// Host-side vocabulary: token classification with the reference's rules and
// the hashed-trie table the kernels probe (layout in wp_table.h).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "wp_table.h"

namespace wp {

struct HostToken {
  std::string word;  // canonical UTF-8 of the decoded word, "##" stripped (utils.cpp:81-85)
  bool is_prefix = true;
  bool is_special = false;
  bool is_malformed = false;
  bool had_invalid = false;  // token line held bytes the decoder dropped
  uint32_t n_cp = 0;
};

struct HostVocab {
  std::vector<HostToken> tokens;
  int32_t unk_id = -1;  // utils.hpp:30
  size_t max_len = 0;   // code points, over non-special non-malformed tokens (fast.cpp:31)

  // device table image
  std::vector<Slot> slots;  // power-of-two size
  std::vector<uint32_t> long_ref;
  std::vector<uint32_t> long_entries;
  std::vector<uint8_t> long_bytes;
  size_t n_nodes = 0;
  size_t n_long = 0;
};

// Returns false (and sets *err) where the reference throws "Vocab word is empty".
bool build_host_vocab(const char *const *tokens, const size_t *lens, size_t n, HostVocab *out, std::string *err);

// Host mirror of the device longest-match query, used by unit tests of the table
// (wp_selftest) — NOT a fallback: nothing on the encode path calls it.
struct MatchResult {
  uint32_t len;  // bytes matched, 0 = miss
  int32_t id;
};
MatchResult host_longest_match(const HostVocab &v, const uint8_t *text, size_t window_bytes, uint32_t kind);

}  // namespace wp
