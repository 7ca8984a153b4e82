// Host-side vocabulary: token classification with the reference's rules and
// the two device tables the kernels use (layout in wp_table.h).
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

  // device table images
  std::vector<Edge> edges;      // power-of-two size
  std::vector<WordSlot> words;  // power-of-two size >= 4 x static words: the static word table
  size_t n_nodes = 0;           // trie nodes (both roots included)
  size_t n_static_words = 0;    // word-initial tokens of at most WORD_KEY_BYTES bytes
  size_t n_long = 0;            // kept tokens longer than WORD_KEY_BYTES bytes (statistics)
};

// Returns false (and sets *err) where the reference throws "Vocab word is empty".
bool build_host_vocab(const char *const *tokens, const size_t *lens, size_t n, HostVocab *out, std::string *err);

// Host mirrors of the device queries, used by unit tests of the table images — NOT a fallback: nothing on
// the encode path calls them.
struct MatchResult {
  uint32_t len;  // bytes matched, 0 = miss
  int32_t id;
};
// longest token of `kind` that is a prefix of text[0, window_bytes)
MatchResult host_longest_match(const HostVocab &v, const uint8_t *text, size_t window_bytes, uint32_t kind);
// whole-segment lookup in the static word table: id count (0 = absent), ids[0..count)
uint32_t host_word_lookup(const HostVocab &v, const uint8_t *text, size_t len, int32_t *ids, uint32_t *slot_out);

}  // namespace wp
