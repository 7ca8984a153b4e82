// Device vocabulary table layout, shared by the host builder (wp_vocab.cpp) and
// the kernels (wp_encode.cu).
//
// The reference keeps two unordered_map<VectorSegment,int> (word-initial and
// "##" continuation, fast.cpp:21-35) and finds the longest match by probing the
// window longest -> shortest (fast.cpp:66-77), O(window) probes per piece.  The
// dictionary semantics are exact (utf8.hpp:60-70), so any exact structure that
// returns the same longest match yields the same ids.  Ours is a HASHED TRIE
// over canonical UTF-8 bytes:
//
//   * one open-addressed table of 32-byte slots (one L2 sector each); a slot is
//     a NODE = one byte-prefix (length 1..WP_KEY_BYTES) of some kept token, of
//     one kind (0 word-initial, 1 continuation), with its bytes stored INLINE so
//     that a probe is verified exactly by the same load that found it;
//   * node existence is monotone in the prefix length, so the deepest node along
//     a window is found by binary search (O(log) probes, first probe = the whole
//     window, which settles the common whole-word hit in one probe);
//   * each node carries the id of the token ending exactly there (term_id) and
//     the longest token that is a proper prefix of it (best_len/best_id), so the
//     longest match is read off the deepest node;
//   * tokens longer than WP_KEY_BYTES hang off their depth-WP_KEY_BYTES node as a
//     list sorted by length (longest first) and are compared byte by byte.
#pragma once
#include <stdint.h>

#include "wp_utf8.h"

namespace wp {

constexpr uint32_t WP_KEY_BYTES = 22;        // inline key bytes per node
constexpr int32_t WP_NO_ID = -2;             // "no token ends here" (ids are >= 0, UNK may be -1)
constexpr uint32_t WP_KIND_PREFIX = 0;       // word-initial map (prefix_to_id)
constexpr uint32_t WP_KIND_SUFFIX = 1;       // "##" map (suffix_to_id)

// Slot = 8 x u32:
//   w[0..4]            key bytes 0..19 (little endian, zero padded)
//   w[5] bits  0..15   key bytes 20..21
//        bits 16..23   len (1..22; 0 = empty slot)
//        bit  24       kind
//        bit  25       has_long (len == 22 and longer tokens share this prefix)
//        bits 26..30   best_len (0 = none; < len)
//   w[6]               term_id  (WP_NO_ID if no token ends at this node)
//   w[7]               best_id  (valid if best_len != 0)
struct Slot {
  uint32_t w[8];
};
static_assert(sizeof(Slot) == 32, "slot must be one 32-byte sector");

constexpr uint32_t WP_W5_KEYMASK = 0x01FFFFFFu;  // key bytes 20..21, len, kind

WP_HD uint32_t slot_len(uint32_t w5) { return (w5 >> 16) & 0xFFu; }
WP_HD uint32_t slot_has_long(uint32_t w5) { return (w5 >> 25) & 1u; }
WP_HD uint32_t slot_best_len(uint32_t w5) { return (w5 >> 26) & 0x1Fu; }
WP_HD uint32_t make_w5(uint32_t bytes2021, uint32_t len, uint32_t kind) {
  return (bytes2021 & 0xFFFFu) | (len << 16) | (kind << 24);
}

// Hash of the six key words (w[5] already reduced to WP_W5_KEYMASK bits).  Multiply-xor: the six products
// are independent (they issue back to back on the GPU), then one xor-shift-multiply finaliser; the table
// index is taken from the low bits, so the finaliser folds the high halves down.
WP_HD uint32_t key_hash(uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3, uint32_t k4, uint32_t k5) {
  uint32_t h = (k0 * 0x9E3779B1u) ^ (k1 * 0x85EBCA77u) ^ (k2 * 0xC2B2AE3Du) ^ (k3 * 0x27D4EB2Fu) ^ (k4 * 0x165667B1u) ^
               ((k5 + 0x7F4A7C15u) * 0xD6E8FEB9u);
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 13;
  h *= 0x297A2D39u;
  h ^= h >> 16;
  return h;
}

// One entry of a long-token list (tokens with more than WP_KEY_BYTES bytes).
struct LongEntry {
  uint32_t len;       // byte length of the token
  int32_t id;         // token id
  uint32_t byte_off;  // offset of the token's canonical bytes in the byte pool
};

// Everything a kernel needs to know about a vocabulary (device pointers).
struct DeviceVocab {
  const Slot *slots;            // n_slots entries, n_slots a power of two
  uint32_t slot_mask;           // n_slots - 1
  const uint32_t *long_ref;     // per slot: index into long_entries of {count, entries...}; only for has_long slots
  const uint32_t *long_entries; // [count, (len,id,byte_off) x count] groups
  const uint8_t *long_bytes;    // byte pool of long tokens
  int32_t unk_id;               // utils.hpp:30 / utils.cpp:112-114
  uint32_t han_swallow;         // 1 iff max_len >= 2 (an OOV Han char swallows the following run)
  uint32_t probe_pairs;         // K2 looks at two slots per probe (pays while the table is L2-resident: <= 16 MiB)
};

}  // namespace wp
