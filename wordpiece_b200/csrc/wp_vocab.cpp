// Host vocabulary builder.  Reference behaviour reproduced here (paths relative
// to gleb-kov/wordpiece):
//   utils.cpp:108-121  parseVocab        id = index, "[UNK]" (last) -> unk id
//   utils.cpp:81-106   WordPieceToken    decode (invalid bytes dropped), "##"
//                                        strip, special / malformed, empty => throw
//   fast.cpp:21-36     map build         skip special/malformed, max_len over the
//                                        rest, duplicate key: last index wins
#include "wp_vocab.h"

#include <algorithm>
#include <cstring>

namespace wp {
namespace {

struct Node {
  int32_t term_id = WP_NO_ID;
  uint32_t n_children = 0;
};

uint32_t log2_ceil(size_t x) {
  uint32_t l = 0;
  while ((size_t(1) << l) < x) l++;
  return l;
}

// the four key words of a word-table key (bytes zero padded)
void word_key(const uint8_t *b, size_t len, uint32_t k[4]) {
  uint8_t buf[WORD_KEY_BYTES] = {0};
  std::memcpy(buf, b, len);
  for (int i = 0; i < 4; i++)
    k[i] = uint32_t(buf[4 * i]) | (uint32_t(buf[4 * i + 1]) << 8) | (uint32_t(buf[4 * i + 2]) << 16) |
           (uint32_t(buf[4 * i + 3]) << 24);
}

}  // namespace

bool build_host_vocab(const char *const *tokens, const size_t *lens, size_t n, HostVocab *out, std::string *err) {
  HostVocab &hv = *out;
  hv = HostVocab();
  hv.tokens.resize(n);

  // (parent, char) -> child while the trie grows: a flat open-addressed map, sized by the byte count of the
  // vocabulary (an upper bound on the number of edges)
  size_t bound = 2;
  for (size_t i = 0; i < n; i++) bound += lens[i];
  const uint32_t tmp_log2 = std::max<uint32_t>(6, log2_ceil(2 * bound + 2));
  const uint32_t tmp_shift = 32 - tmp_log2;
  const uint32_t tmp_mask = (1u << tmp_log2) - 1u;
  constexpr uint64_t TMP_EMPTY = ~0ull;
  std::vector<uint64_t> tmp_key(size_t(1) << tmp_log2, TMP_EMPTY);  // parent << 32 | char
  std::vector<uint32_t> tmp_child(size_t(1) << tmp_log2, 0u);
  std::vector<Node> nodes(2);  // node 0: root of the word-initial map, node 1: root of the "##" map
  std::vector<uint32_t> edge_order;  // temp-map slots in creation order (parents before children)
  bool too_many = false;
  auto child_of = [&](uint32_t parent, uint32_t ch) -> uint32_t {
    const uint64_t key = (static_cast<uint64_t>(parent) << 32) | ch;
    uint32_t h = edge_hash(parent, ch, tmp_shift);
    while (tmp_key[h] != TMP_EMPTY) {
      if (tmp_key[h] == key) return tmp_child[h];
      h = (h + 1) & tmp_mask;
    }
    if (nodes.size() >= EDGE_MAX_NODES) {
      too_many = true;
      return parent;
    }
    tmp_key[h] = key;
    tmp_child[h] = static_cast<uint32_t>(nodes.size());
    nodes.emplace_back();
    nodes[parent].n_children++;
    edge_order.push_back(h);
    return tmp_child[h];
  };

  for (size_t i = 0; i < n; i++) {
    const uint8_t *b = reinterpret_cast<const uint8_t *>(tokens[i]);
    const size_t len = lens[i];
    if (len == 5 && std::memcmp(b, "[UNK]", 5) == 0) hv.unk_id = static_cast<int32_t>(i);  // utils.cpp:112-114

    // decode, dropping invalid bytes (utf8.cpp:130-147)
    std::vector<uint32_t> cps;
    cps.reserve(len);
    HostToken &t = hv.tokens[i];
    for (size_t p = 0; p < len;) {
      uint32_t cp = 0;
      const size_t rem = len - p;
      const uint32_t l = utf8_decode(b[p], rem > 1 ? b[p + 1] : 0, rem > 2 ? b[p + 2] : 0, rem > 3 ? b[p + 3] : 0,
                                     rem > 4 ? 4u : static_cast<uint32_t>(rem), &cp);
      if (l == 0) {
        t.had_invalid = true;
        p += 1;
      } else {
        cps.push_back(cp);
        p += l;
      }
    }
    size_t first = 0;
    if (cps.size() >= 2 && cps[0] == '#' && cps[1] == '#') {  // utils.cpp:83-85,139-141
      t.is_prefix = false;
      first = 2;
    } else if (cps.size() > 2 && cps[0] == '[' && cps.back() == ']') {  // utils.cpp:86-88,143-146
      t.is_special = true;
    }
    if (cps.size() == first) {  // utils.cpp:99-101
      if (err) *err = "Vocab word is empty";
      return false;
    }
    bool all_punct = true;
    for (size_t k = first; k < cps.size(); k++) {
      if (!cp_is_punct(cps[k]) && !cp_is_space(cps[k])) all_punct = false;
      uint8_t enc[4];
      const uint32_t l = utf8_encode(cps[k], enc);
      t.word.append(reinterpret_cast<const char *>(enc), l);
    }
    t.n_cp = static_cast<uint32_t>(cps.size() - first);
    t.is_malformed = all_punct && t.n_cp > 1;  // utils.cpp:102-105

    if (t.is_special || t.is_malformed) continue;  // fast.cpp:28-30
    hv.max_len = std::max<size_t>(hv.max_len, t.n_cp);  // fast.cpp:31

    uint32_t node = t.is_prefix ? WP_KIND_PREFIX : WP_KIND_SUFFIX;
    for (size_t p = 0; p < t.word.size();) {  // canonical UTF-8: every lead byte tells its char's length
      const uint8_t *wb = reinterpret_cast<const uint8_t *>(t.word.data()) + p;
      const uint32_t cl = utf8_lead_len(wb[0]);
      uint32_t ch = 0;
      for (uint32_t q = 0; q < cl; q++) ch |= static_cast<uint32_t>(wb[q]) << (8 * q);
      node = child_of(node, ch);
      p += cl;
    }
    if (too_many) {
      if (err) *err = "vocabulary too large (more than 2^24 trie nodes)";
      return false;
    }
    if (nodes[node].term_id == WP_NO_ID && t.word.size() > WORD_KEY_BYTES) hv.n_long++;
    nodes[node].term_id = static_cast<int32_t>(i);  // fast.cpp:34: assignment => last duplicate wins
  }
  hv.n_nodes = nodes.size();

  // ---- E: the edge table, load factor <= 0.25, linear probing; parents are inserted before their children, so
  // the edges near the roots — the hot ones — sit in (or next to) their home slots.  Every piece ends with a
  // step that FAILS (unless it ends at a leaf), and a failing lookup walks to the next empty slot: 1.4
  // dependent loads on average at a quarter full against 2.5 at half full (4 MB for a 29k vocabulary, 32 MB
  // for a 120k one — well inside the 126 MB L2)
  const size_t n_edges = edge_order.size();
  const uint32_t e_log2 = std::max<uint32_t>(6, log2_ceil(4 * n_edges + 2));
  const uint32_t e_shift = 32 - e_log2, e_mask = (1u << e_log2) - 1u;
  hv.edges.assign(size_t(1) << e_log2, Edge{EDGE_EMPTY, 0u, 0u, WP_NO_ID});
  for (uint32_t h : edge_order) {
    const uint32_t parent = static_cast<uint32_t>(tmp_key[h] >> 32), ch = static_cast<uint32_t>(tmp_key[h]);
    const uint32_t child = tmp_child[h];
    uint32_t idx = edge_hash(parent, ch, e_shift);
    while (hv.edges[idx].parent != EDGE_EMPTY) idx = (idx + 1) & e_mask;
    hv.edges[idx] = Edge{parent, ch, child | (nodes[child].n_children ? EDGE_HAS_CHILDREN : 0u), nodes[child].term_id};
  }

  // ---- W: static part of the word table — every word-initial token of at most WORD_KEY_BYTES bytes, with the
  // id its trie node ended up with (last duplicate wins)
  // K1 settles a single-char segment that is absent from W as UNK, so the static part must be complete: the
  // table is sized so that the static words fill at most a quarter of it
  size_t n_short = 0;
  for (size_t i = 0; i < n; i++) {
    const HostToken &t = hv.tokens[i];
    if (!t.is_special && !t.is_malformed && t.is_prefix && t.word.size() <= WORD_KEY_BYTES) n_short++;
  }
  uint32_t word_slots_log2 = 6;
  while (word_slots_log2 < 25 && (size_t(1) << word_slots_log2) < 4 * n_short) word_slots_log2++;
  const uint32_t w_shift = 32 - word_slots_log2, w_mask = (1u << word_slots_log2) - 1u;
  hv.words.assign(size_t(1) << word_slots_log2, WordSlot{});
  for (size_t i = 0; i < n; i++) {
    const HostToken &t = hv.tokens[i];
    if (t.is_special || t.is_malformed || !t.is_prefix || t.word.size() > WORD_KEY_BYTES) continue;
    const uint8_t *wb = reinterpret_cast<const uint8_t *>(t.word.data());
    const uint32_t len = static_cast<uint32_t>(t.word.size());
    const MatchResult whole = host_longest_match(hv, wb, len, WP_KIND_PREFIX);  // == {len, final id of this word}
    uint32_t k[4];
    word_key(wb, len, k);
    uint32_t idx = word_hash(k[0], k[1], k[2], k[3], len, w_shift);
    bool present = false;
    while (hv.words[idx].meta != 0) {
      const WordSlot &s = hv.words[idx];
      if (word_meta_len(s.meta) == len && s.key[0] == k[0] && s.key[1] == k[1] && s.key[2] == k[2] && s.key[3] == k[3]) {
        present = true;
        break;
      }
      idx = (idx + 1) & w_mask;
    }
    if (present) continue;
    WordSlot &s = hv.words[idx];
    for (int q = 0; q < 4; q++) s.key[q] = k[q];
    s.meta = word_meta(len, 1, false);
    s.ids[0] = whole.id;
    hv.n_static_words++;
  }
  return true;
}

MatchResult host_longest_match(const HostVocab &v, const uint8_t *text, size_t window, uint32_t kind) {
  const uint32_t mask = static_cast<uint32_t>(v.edges.size() - 1);
  const uint32_t shift = 32 - log2_ceil(v.edges.size());
  uint32_t node = kind ? WP_KIND_SUFFIX : WP_KIND_PREFIX;
  MatchResult best{0, WP_NO_ID};
  for (size_t d = 0; d < window;) {
    const uint32_t cl = utf8_lead_len(text[d]);
    if (cl == 0 || d + cl > window) break;  // (the kernels only ever walk clean text: whole, valid chars)
    uint32_t ch = 0;
    for (uint32_t q = 0; q < cl; q++) ch |= static_cast<uint32_t>(text[d + q]) << (8 * q);
    uint32_t idx = edge_hash(node, ch, shift);
    while (!(v.edges[idx].parent == node && v.edges[idx].ch == ch) && v.edges[idx].parent != EDGE_EMPTY) idx = (idx + 1) & mask;
    const Edge &e = v.edges[idx];
    if (e.parent != node) break;
    node = e.child & EDGE_CHILD_MASK;
    d += cl;
    if (e.term_id != WP_NO_ID) best = MatchResult{static_cast<uint32_t>(d), e.term_id};
    if (!(e.child & EDGE_HAS_CHILDREN)) break;
  }
  return best;
}

uint32_t host_word_lookup(const HostVocab &v, const uint8_t *text, size_t len, int32_t *ids, uint32_t *slot_out) {
  if (len == 0 || len > WORD_KEY_BYTES) return 0;
  const uint32_t mask = static_cast<uint32_t>(v.words.size() - 1);
  const uint32_t shift = 32 - log2_ceil(v.words.size());
  uint32_t k[4];
  word_key(text, len, k);
  uint32_t idx = word_hash(k[0], k[1], k[2], k[3], static_cast<uint32_t>(len), shift);
  while (v.words[idx].meta != 0) {
    const WordSlot &s = v.words[idx];
    if (word_meta_len(s.meta) == len && s.key[0] == k[0] && s.key[1] == k[1] && s.key[2] == k[2] && s.key[3] == k[3]) {
      const uint32_t cnt = word_meta_count(s.meta);
      for (uint32_t t = 0; t < cnt; t++) ids[t] = s.ids[t];
      if (slot_out) *slot_out = idx;
      return cnt;
    }
    idx = (idx + 1) & mask;
  }
  return 0;
}

}  // namespace wp
