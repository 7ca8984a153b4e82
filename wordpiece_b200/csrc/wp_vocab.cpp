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
  uint32_t kw[6];  // the node's key: its bytes (<= WP_KEY_BYTES, zero padded), length and kind, as in the table
  int32_t term_id = WP_NO_ID;
  uint32_t best_len = 0;
  int32_t best_id = WP_NO_ID;
  int32_t long_list = -1;  // index into the side table of long-token lists (tokens longer than WP_KEY_BYTES)
};

struct LongList {
  std::vector<LongEntry> longs;               // byte_off filled at emission
  std::vector<const std::string *> long_str;  // parallel to longs
};

inline void key_words(const uint8_t *b, uint32_t len, uint32_t kind, uint32_t kw[6]) {
  uint8_t buf[24] = {0};
  std::memcpy(buf, b, len);
  for (int i = 0; i < 5; i++)
    kw[i] = uint32_t(buf[4 * i]) | (uint32_t(buf[4 * i + 1]) << 8) | (uint32_t(buf[4 * i + 2]) << 16) |
            (uint32_t(buf[4 * i + 3]) << 24);
  kw[5] = make_w5(uint32_t(buf[20]) | (uint32_t(buf[21]) << 8), len, kind);
}

}  // namespace

bool build_host_vocab(const char *const *tokens, const size_t *lens, size_t n, HostVocab *out, std::string *err) {
  HostVocab &hv = *out;
  hv = HostVocab();
  hv.tokens.resize(n);

  // node key -> nodes[]: a flat open-addressed map on the key words themselves (no strings, no allocations;
  // the per-prefix std::string keys of the first version made a 120k vocabulary take 0.8 s to build)
  size_t bound = 0;
  for (size_t i = 0; i < n; i++) bound += std::min<size_t>(lens[i], WP_KEY_BYTES);
  size_t map_size = 64;
  while (map_size < 2 * bound + 2) map_size <<= 1;
  std::vector<uint32_t> index(map_size, 0u);  // node index + 1, 0 = empty
  const uint32_t map_mask = static_cast<uint32_t>(map_size - 1);
  std::vector<Node> nodes;
  nodes.reserve(bound);
  std::vector<LongList> long_lists;
  auto same_key = [](const uint32_t a[6], const uint32_t b[6]) {
    return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3] && a[4] == b[4] && a[5] == b[5];
  };
  // find (create = false: SIZE_MAX if absent) / find-or-create the node with key kw
  auto get_node_kw = [&](const uint32_t kw[6], bool create) -> size_t {
    uint32_t h = key_hash(kw[0], kw[1], kw[2], kw[3], kw[4], kw[5]) & map_mask;
    for (;;) {
      const uint32_t e = index[h];
      if (e == 0) break;
      if (same_key(nodes[e - 1].kw, kw)) return e - 1;
      h = (h + 1) & map_mask;
    }
    if (!create) return static_cast<size_t>(-1);
    nodes.emplace_back();
    for (int q = 0; q < 6; q++) nodes.back().kw[q] = kw[q];
    index[h] = static_cast<uint32_t>(nodes.size());
    return nodes.size() - 1;
  };
  auto get_node = [&](uint32_t kind, const char *b, size_t k) -> size_t {
    uint32_t kw[6];
    key_words(reinterpret_cast<const uint8_t *>(b), static_cast<uint32_t>(k), kind, kw);
    return get_node_kw(kw, true);
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

    const uint32_t kind = t.is_prefix ? WP_KIND_PREFIX : WP_KIND_SUFFIX;
    const size_t L = t.word.size();
    const size_t depth = std::min<size_t>(L, WP_KEY_BYTES);
    size_t ni = 0;
    for (size_t k = 1; k <= depth; k++) ni = get_node(kind, t.word.data(), k);
    if (L <= WP_KEY_BYTES) {
      nodes[ni].term_id = static_cast<int32_t>(i);  // fast.cpp:34: assignment => last duplicate wins
    } else {
      if (nodes[ni].long_list < 0) {
        nodes[ni].long_list = static_cast<int32_t>(long_lists.size());
        long_lists.emplace_back();
      }
      LongList &ll = long_lists[nodes[ni].long_list];
      bool replaced = false;
      for (size_t e = 0; e < ll.longs.size(); e++) {
        if (*ll.long_str[e] == t.word) {
          ll.longs[e].id = static_cast<int32_t>(i);
          replaced = true;
          break;
        }
      }
      if (!replaced) {
        ll.longs.push_back(LongEntry{static_cast<uint32_t>(L), static_cast<int32_t>(i), 0});
        ll.long_str.push_back(&t.word);
      }
    }
  }

  // best_len / best_id: longest token that is a PROPER prefix of the node.  A node is created after all its
  // shorter prefixes (get_node is called with k = 1, 2, ...), so index order has parents first.
  for (size_t oi = 0; oi < nodes.size(); oi++) {
    Node &nd = nodes[oi];
    const uint32_t k = slot_len(nd.kw[5]);
    if (k <= 1) continue;
    // the parent's key: the same bytes without the last one
    uint32_t pk[6];
    for (int q = 0; q < 6; q++) pk[q] = nd.kw[q];
    const uint32_t last = k - 1;  // index of the byte to clear
    if (last < 20) pk[last >> 2] &= ~(0xFFu << (8 * (last & 3)));
    else pk[5] &= ~(0xFFu << (8 * (last - 20)));
    pk[5] = (pk[5] & ~(0xFFu << 16)) | (last << 16);
    const Node &par = nodes[get_node_kw(pk, false)];
    if (par.term_id != WP_NO_ID) {
      nd.best_len = k - 1;
      nd.best_id = par.term_id;
    } else {
      nd.best_len = par.best_len;
      nd.best_id = par.best_id;
    }
  }

  // open-addressed table, linear probing.  Load factor <= 0.25 while the table stays small (a miss — most
  // binary-search probes are misses — then costs ~1.4 slot loads instead of ~2.5), <= 0.5 for huge vocabularies.
  size_t n_slots = 64;
  while (n_slots < 4 * nodes.size()) n_slots <<= 1;
  if (n_slots * sizeof(Slot) > (size_t(64) << 20)) n_slots >>= 1;
  hv.slots.assign(n_slots, Slot{{0, 0, 0, 0, 0, 0, 0, 0}});
  hv.long_ref.assign(n_slots, 0);
  hv.long_entries.clear();
  hv.long_entries.push_back(0);  // index 0 = "no list"
  hv.long_bytes.clear();
  const uint32_t mask = static_cast<uint32_t>(n_slots - 1);
  for (Node &nd : nodes) {
    const uint32_t *kw = nd.kw;
    const bool has_long = nd.long_list >= 0;
    uint32_t idx = key_hash(kw[0], kw[1], kw[2], kw[3], kw[4], kw[5]) & mask;
    while (slot_len(hv.slots[idx].w[5]) != 0) idx = (idx + 1) & mask;
    Slot &s = hv.slots[idx];
    for (int i = 0; i < 5; i++) s.w[i] = kw[i];
    s.w[5] = kw[5] | (has_long ? (1u << 25) : 0u) | (nd.best_len << 26);
    s.w[6] = static_cast<uint32_t>(nd.term_id);
    s.w[7] = static_cast<uint32_t>(nd.best_id);
    if (has_long) {
      // longest first, so the first full match is the longest (fast.cpp:66-77 probes longest first)
      const LongList &ll = long_lists[nd.long_list];
      std::vector<size_t> ord(ll.longs.size());
      for (size_t i = 0; i < ord.size(); i++) ord[i] = i;
      std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return ll.longs[a].len > ll.longs[b].len; });
      hv.long_ref[idx] = static_cast<uint32_t>(hv.long_entries.size());
      hv.long_entries.push_back(static_cast<uint32_t>(ord.size()));
      for (size_t oi : ord) {
        LongEntry e = ll.longs[oi];
        e.byte_off = static_cast<uint32_t>(hv.long_bytes.size());
        hv.long_bytes.insert(hv.long_bytes.end(), ll.long_str[oi]->begin(), ll.long_str[oi]->end());
        hv.long_entries.push_back(e.len);
        hv.long_entries.push_back(static_cast<uint32_t>(e.id));
        hv.long_entries.push_back(e.byte_off);
        hv.n_long++;
      }
    }
  }
  hv.n_nodes = nodes.size();
  // pad the pools so that the device never sees a null / zero-sized buffer
  while (hv.long_bytes.size() % 16 != 0 || hv.long_bytes.empty()) hv.long_bytes.push_back(0);
  return true;
}

MatchResult host_longest_match(const HostVocab &v, const uint8_t *text, size_t window, uint32_t kind) {
  const uint32_t mask = static_cast<uint32_t>(v.slots.size() - 1);
  auto probe = [&](uint32_t k, const Slot **out) -> bool {
    uint32_t kw[6];
    key_words(text, k, kind, kw);
    uint32_t idx = key_hash(kw[0], kw[1], kw[2], kw[3], kw[4], kw[5]) & mask;
    for (;;) {
      const Slot &s = v.slots[idx];
      if (slot_len(s.w[5]) == 0) return false;
      if (s.w[0] == kw[0] && s.w[1] == kw[1] && s.w[2] == kw[2] && s.w[3] == kw[3] && s.w[4] == kw[4] &&
          ((s.w[5] ^ kw[5]) & WP_W5_KEYMASK) == 0) {
        *out = &s;
        return true;
      }
      idx = (idx + 1) & mask;
    }
  };
  const uint32_t k0 = static_cast<uint32_t>(std::min<size_t>(window, WP_KEY_BYTES));
  if (k0 == 0) return MatchResult{0, WP_NO_ID};
  const Slot *node = nullptr;
  uint32_t lo = 0;
  if (probe(k0, &node)) {
    lo = k0;
  } else {
    uint32_t hi = k0;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) / 2;
      const Slot *s = nullptr;
      if (probe(mid, &s)) {
        lo = mid;
        node = s;
      } else {
        hi = mid;
      }
    }
  }
  if (lo == 0) return MatchResult{0, WP_NO_ID};
  if (lo == WP_KEY_BYTES && slot_has_long(node->w[5]) && window > WP_KEY_BYTES) {
    const uint32_t ref = v.long_ref[static_cast<size_t>(node - v.slots.data())];
    const uint32_t cnt = v.long_entries[ref];
    for (uint32_t e = 0; e < cnt; e++) {
      const uint32_t len = v.long_entries[ref + 1 + 3 * e];
      if (len > window) continue;
      const uint32_t off = v.long_entries[ref + 3 + 3 * e];
      if (std::memcmp(text + WP_KEY_BYTES, v.long_bytes.data() + off + WP_KEY_BYTES, len - WP_KEY_BYTES) == 0)
        return MatchResult{len, static_cast<int32_t>(v.long_entries[ref + 2 + 3 * e])};
    }
  }
  if (static_cast<int32_t>(node->w[6]) != WP_NO_ID) return MatchResult{lo, static_cast<int32_t>(node->w[6])};
  if (slot_best_len(node->w[5]) != 0) return MatchResult{slot_best_len(node->w[5]), static_cast<int32_t>(node->w[7])};
  return MatchResult{0, WP_NO_ID};
}

}  // namespace wp
