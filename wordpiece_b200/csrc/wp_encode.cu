// WordPiece encode kernels for sm_100a: word split -> longest match -> scan +
// scatter.  Text is read from HBM once (plus the bytes of the ~20 % of words
// that need more than one probe); ids are written once; the intermediates are a
// 4-byte result per segment and a 16-byte entry per unsettled segment.
//
// What they reproduce (gleb-kov/wordpiece):
//   utils.cpp:37-79 / utf8.cpp:130-147   strict UTF-8 decode, invalid bytes dropped
//   utf8.cpp:10-29                       space / punctuation / Han classes
//   fast.cpp:38-41                       word-initial positions
//   fast.cpp:43-99                       the greedy longest-match worker with
//                                        whole-word UNK roll-back
//   fast.cpp:101-138                     chunk + concat (here: tiles + decoupled
//                                        look-back scans)
//
// K1 wp_split_kernel (one 4 KB tile per CTA)
//   S1 split : the tile + halo is staged in shared memory with 16-byte loads;
//              each thread classifies a 32-byte chunk into bit masks (valid lead
//              / space / punct / Han) with SWAR arithmetic; tiles holding invalid
//              UTF-8 are compacted in shared memory; segment starts and ends
//              ("safe starts") fall out of mask operations and are
//              compacted into lists in text order.
//   S2a probe: one whole-window probe per segment into the hashed-trie table
//              (wp_table.h), uniform work for every lane.  It settles every
//              segment that is a single token (~80 % of English words) — its id
//              goes straight to seg_result[] — and appends the rest to the
//              global slow list.
// K2 wp_match_kernel (whole GPU, no tiles, no barriers)
//   S2b match: every lane owns many slow segments and runs the greedy matcher as
//              a FLATTENED state machine — one table probe per loop iteration,
//              whatever the lane is doing (binary-search step, next piece,
//              collision) — so a warp stays converged on the probe and chains of
//              very different length average out over a lane's share.  Segments
//              that left their tile's window or hold invalid UTF-8 are walked
//              from global memory by the exact byte-wise lane (walk_segment).
// K3 wp_scatter_kernel
//   S3 scatter: per-segment id counts -> block scan -> decoupled look-back ->
//              ids staged in shared memory and written out coalesced.
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "wp_encode.h"
#include "wp_table.h"

namespace wp {

// ------------------------------------------------------------------ geometry
constexpr int TILE = 4096;                       // text bytes owned by one CTA of K1
constexpr int CHUNK = 32;                        // bytes classified by one thread
constexpr int THREADS = 160;                     // K1: 5 warps
constexpr int HALO = 256;                        // classified bytes past the tile (segment completion)
constexpr int LOOKAHEAD = 32;                    // loaded, not classified (UTF-8 validation look-ahead)
constexpr int LEFT = 16;                         // bytes before the tile (ownership of leading continuation bytes)
constexpr int WINDOW = TILE + HALO;              // classified window
constexpr int NCHUNK = WINDOW / CHUNK;           // 136
constexpr int OWNED_CHUNKS = TILE / CHUNK;       // 128
constexpr int RAW_BYTES = LEFT + WINDOW + LOOKAHEAD;  // 4400
constexpr int WARPS = THREADS / 32;
constexpr int MAX_TILE_SLOW = TILE / 2 + 8;      // slow segments have >= 2 bytes
constexpr int PREFETCH_TILES = 2 * 6 * 148;      // K1 pulls the text of the tile this far ahead into L2
constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr uint32_t POS_MASK = 0x3FFFu;           // window positions fit 14 bits
constexpr uint32_t SLOW_FIRST_MISSED = 0x8000u;  // tile slow-list flag: the whole-window probe already missed
constexpr uint32_t SLOW_WALK = 0x4000u;          // tile slow-list flag: segment leaves the window

// Rows of the key-mask table are 48 bytes apart (32 used): lanes read rows by their own k, and with this
// stride the rows of k = 1..8 — nearly all probes — start in eight different groups of four banks, so the
// two 16-byte loads of a probe do not serialise on bank conflicts.
constexpr int KEY_MASK_ROW = 3;                  // in uint4 units

constexpr int MATCH_THREADS = 256;               // K2
constexpr int SCATTER_THREADS = 256;             // K3
constexpr int SCATTER_ITEMS = 8;                 // segments per thread and block iteration
constexpr int SCATTER_SEGS = SCATTER_THREADS * SCATTER_ITEMS;  // 2048
constexpr int SCATTER_STAGE = 6144;              // ids staged in shared memory per block iteration

static_assert(RAW_BYTES % 16 == 0, "raw buffer is loaded in 16-byte units");
static_assert(WP_KEY_BYTES + 4 <= LOOKAHEAD, "key window reads stay inside the loaded bytes");
static_assert(NCHUNK + 1 <= THREADS, "one thread per chunk in the classification and compaction passes");
static_assert(WINDOW <= static_cast<int>(POS_MASK), "positions must fit the packed list entries");
static_assert(TILE <= 4096, "segment ordinals must fit 12 bits of the tile slow list");

struct __align__(16) TileSmem {
  uint8_t raw[RAW_BYTES];              // [0,LEFT) left halo, then the window, then look-ahead
  uint16_t seg_s[TILE];                // owned segment k (text order): start position | class << 14
  uint16_t seg_e[WINDOW + 64];         // j-th segment end in the window
  uint16_t slow[MAX_TILE_SLOW];        // segments the whole-window probe did not settle: ordinal | flags
  uint32_t settled[TILE / 32];         // bit k: segment k was settled by S2a (its result is parked in seg_s/seg_e)
  uint32_t n_slow2;                    // entries of slow[] left after the word memo settled its share
  uint32_t memo_hits;
  uint32_t memo_use;                   // this tile runs the memo phase (decided by one thread)
  uint32_t any_single;                 // some single-char segment is still to be settled (see S2a)
  uint32_t m_lead[NCHUNK + 1];         // valid lead bytes
  uint32_t m_space[NCHUNK + 1];
  uint32_t m_punct[NCHUNK + 1];
  uint32_t m_han[NCHUNK + 1];
  uint32_t m_cover[NCHUNK + 1];        // bytes covered by a valid sequence that starts in this chunk
  uint32_t m_kept[NCHUNK + 1];         // bytes of the RAW window that survive the strict decoder (dirty tiles)
  uint32_t kept_scan[NCHUNK + 2];      // exclusive scan of kept bytes per chunk (dirty tiles)
  uint8_t spill[NCHUNK + 1];           // bytes by which the chunk's last sequence runs into the next chunk
  uint4 key_mask[KEY_MASK_ROW * (WP_KEY_BYTES + 1)];  // row k: masks of the six key words for a k-byte key, then k << 16
  uint32_t warp_sums[WARPS];
  uint32_t prev_class;                 // class of the last valid char before the tile
  uint32_t left_spill;                 // bytes of the tile start covered by a sequence that began before it
  uint32_t n_segs;                     // owned segments in this tile
  uint32_t n_ends;                     // segment ends found in the window
  uint32_t n_slow;                     // entries of slow[]
  uint32_t slow_base;                  // first global slow index of this tile
  uint32_t tok_base;                   // first id-scratch slot reserved for this tile
  unsigned long long seg_base;         // segments of all earlier tiles of the range
};

// ------------------------------------------------------------------- helpers

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ uint32_t ld_u32(const uint8_t *p) { return *reinterpret_cast<const uint32_t *>(p); }

// One 32-byte table slot with ONE 256-bit load (sm_100: LDG.E.256) through the read-only path: half the
// L1 wavefronts of two 16-byte gathers — the probes are the dominant L1 traffic of K2.
// Shared-memory add by ONE lane that already speaks for its warp (ballot + popc done by the caller).  Plain
// atomicAdd() here makes the compiler wrap its own warp aggregation (vote, flo, popc, shfl: 16 instructions)
// around the single active lane.
__device__ __forceinline__ uint32_t smem_add(uint32_t *p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;"
               : "=r"(old)
               : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))), "r"(v)
               : "memory");
  return old;
}
__device__ __forceinline__ void smem_add_noret(uint32_t *p, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))), "r"(v) : "memory");
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void ld_slot(const uint4 *tab, uint32_t idx, uint4 *a, uint4 *b) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a->x), "=r"(a->y), "=r"(a->z), "=r"(a->w), "=r"(b->x), "=r"(b->y), "=r"(b->z), "=r"(b->w)
               : "l"(tab + 2 * static_cast<size_t>(idx)));
}

// 0x80 in every byte lane whose (7-bit) value lies in [lo, hi]; lanes must be < 0x80.
__device__ __forceinline__ uint32_t swar_range(uint32_t w, uint32_t lo, uint32_t hi) {
  const uint32_t ge_lo = w + (0x80808080u - lo * 0x01010101u);
  const uint32_t gt_hi = w + (0x80808080u - (hi + 1u) * 0x01010101u);
  return ge_lo & ~gt_hi & 0x80808080u;
}

// gather the four 0x80 flags of a word into a nibble (lane 0 -> bit 0)
__device__ __forceinline__ uint32_t swar_nibble(uint32_t flags) { return (((flags >> 7) * 0x01020408u) >> 24) & 0xFu; }

// Class of a decoded char from its UTF-8 bytes (valid sequence of length len).
__device__ __forceinline__ uint32_t class_of(uint32_t cp) { return cp_class(cp); }

// ---------------------------------------------------------------- table probe

// Outcome of looking at two consecutive slots (idx, idx+1) for one key: with linear probing at load <= 0.25
// the key, or the empty slot that proves its absence, is in this pair for all but ~1 % of the probes, so a
// probe is one turn (two 256-bit loads issued together) instead of a chain of dependent single-slot turns.
enum PairOutcome : uint32_t { PAIR_MISS = 0, PAIR_HIT0 = 1, PAIR_HIT1 = 2, PAIR_BOTH_OTHER = 3 };

__device__ __forceinline__ bool slot_matches(const uint4 &a, const uint4 &b, const uint32_t kw[6]) {
  return a.x == kw[0] && a.y == kw[1] && a.z == kw[2] && a.w == kw[3] && b.x == kw[4] &&
         ((b.y ^ kw[5]) & WP_W5_KEYMASK) == 0;
}

__device__ __forceinline__ uint32_t pair_outcome(const uint4 &a0, const uint4 &b0, const uint4 &a1, const uint4 &b1,
                                                 const uint32_t kw[6]) {
  if (slot_len(b0.y) == 0) return PAIR_MISS;
  if (slot_matches(a0, b0, kw)) return PAIR_HIT0;
  if (slot_len(b1.y) == 0) return PAIR_MISS;
  if (slot_matches(a1, b1, kw)) return PAIR_HIT1;
  return PAIR_BOTH_OTHER;
}

struct NodeHit {
  uint32_t w5;
  int32_t term_id;
  int32_t best_id;
  uint32_t slot;
};

__device__ __forceinline__ bool probe_node(const DeviceVocab &V, const uint32_t kw[6], NodeHit *hit) {
  uint32_t idx = key_hash(kw[0], kw[1], kw[2], kw[3], kw[4], kw[5]) & V.slot_mask;
  const uint4 *tab = reinterpret_cast<const uint4 *>(V.slots);
  for (;;) {
    const uint4 a = __ldg(tab + 2 * idx);
    const uint4 b = __ldg(tab + 2 * idx + 1);
    if (slot_len(b.y) == 0) return false;
    if (a.x == kw[0] && a.y == kw[1] && a.z == kw[2] && a.w == kw[3] && b.x == kw[4] &&
        ((b.y ^ kw[5]) & WP_W5_KEYMASK) == 0) {
      hit->w5 = b.y;
      hit->term_id = static_cast<int32_t>(b.z);
      hit->best_id = static_cast<int32_t>(b.w);
      hit->slot = idx;
      return true;
    }
    idx = (idx + 1) & V.slot_mask;
  }
}

// Key words for the first k (1..22) bytes of the 24 raw window bytes r[0..5].
__device__ __forceinline__ void make_key(const uint32_t r[6], uint32_t k, uint32_t kind, uint32_t kw[6]) {
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const int nb = static_cast<int>(k) - 4 * i;
    kw[i] = nb >= 4 ? r[i] : (nb <= 0 ? 0u : (r[i] & ((1u << (8 * nb)) - 1u)));
  }
  const int nb5 = static_cast<int>(k) - 20;
  const uint32_t tail = nb5 <= 0 ? 0u : (r[5] & ((1u << (8 * nb5)) - 1u));
  kw[5] = make_w5(tail, k, kind);
}

// Same, with the byte masks read from a row of a shared-memory table
// (two 16-byte loads and six ANDs instead of a compare/select chain per word).
__device__ __forceinline__ void make_key_tab(const uint4 *key_mask, const uint32_t r[6], uint32_t k, uint32_t kind,
                                             uint32_t kw[6]) {
  const uint4 ma = key_mask[KEY_MASK_ROW * k];
  const uint4 mb = key_mask[KEY_MASK_ROW * k + 1];
  kw[0] = r[0] & ma.x;
  kw[1] = r[1] & ma.y;
  kw[2] = r[2] & ma.z;
  kw[3] = r[3] & ma.w;
  kw[4] = r[4] & mb.x;
  kw[5] = (r[5] & mb.y) | mb.z | (kind << 24);
}

// Deepest trie node along the window whose first min(window,22) bytes are r[];
// returns its depth (0 = none).  Node existence is monotone in the depth, so
// after the whole-window probe a binary search suffices.
__device__ __forceinline__ uint32_t deepest_node(const DeviceVocab &V, const uint32_t r[6], uint32_t k0, uint32_t kind,
                                                 NodeHit *node) {
  uint32_t kw[6];
  make_key(r, k0, kind, kw);
  if (probe_node(V, kw, node)) return k0;
  uint32_t lo = 0, hi = k0;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    NodeHit h;
    make_key(r, mid, kind, kw);
    if (probe_node(V, kw, &h)) {
      lo = mid;
      *node = h;
    } else {
      hi = mid;
    }
  }
  return lo;
}

// --------------------------------------------- global-memory walker (slow lane)
// Walks ONE segment straight from the text in global memory, dropping invalid
// bytes on the fly.  Used for the (at most one per tile) segment that does not
// end inside the tile's window.  Single thread; two passes: count, then emit.

struct TextView {
  const uint8_t *t;
  size_t n;
};

// Decode the char at pos; returns its length (0 = invalid byte) and class.
__device__ uint32_t gdecode(const TextView &tv, size_t pos, uint32_t *cls) {
  const size_t rem = tv.n - pos;
  const uint32_t b0 = tv.t[pos];
  if (b0 < 0x80u) {
    *cls = cp_class(b0);
    return 1;
  }
  const uint32_t b1 = rem > 1 ? tv.t[pos + 1] : 0u;
  const uint32_t b2 = rem > 2 ? tv.t[pos + 2] : 0u;
  const uint32_t b3 = rem > 3 ? tv.t[pos + 3] : 0u;
  uint32_t cp = 0;
  const uint32_t len = utf8_decode(b0, b1, b2, b3, rem > 4 ? 4u : static_cast<uint32_t>(rem), &cp);
  if (len) *cls = cp_class(cp);
  return len;
}

// First valid lead at or after pos (n if none); returns its length and class.
__device__ size_t gnext(const TextView &tv, size_t pos, uint32_t *len, uint32_t *cls) {
  while (pos < tv.n) {
    const uint32_t l = gdecode(tv, pos, cls);
    if (l) {
      *len = l;
      return pos;
    }
    pos++;
  }
  *len = 0;
  *cls = CLS_SPACE;
  return tv.n;
}

// Class of the last valid char that starts before pos (SPACE at the text start).
__device__ uint32_t gprev_class(const TextView &tv, size_t pos) {
  size_t p = pos;
  while (p > 0) {
    size_t q = p - 1;
    while (q > 0 && is_cont_byte(tv.t[q])) q--;
    if (!is_cont_byte(tv.t[q])) {
      uint32_t cls;
      if (gdecode(tv, q, &cls)) return cls;
    }
    p = q;
  }
  return CLS_SPACE;
}

// Longest match for the window that starts at the valid lead `p` (first char of
// any non-space class, then ordinary chars only).
__device__ uint32_t longest_match_global(const DeviceVocab &V, const TextView &tv, size_t p, uint32_t kind,
                                         int32_t *id) {
  uint8_t kb[24];
#pragma unroll
  for (int i = 0; i < 24; i++) kb[i] = 0;
  uint32_t klen = 0;
  bool more = false;
  {
    size_t q = p;
    bool first = true;
    while (q < tv.n) {
      uint32_t len, cls;
      q = gnext(tv, q, &len, &cls);
      if (q >= tv.n) break;
      if (!first && cls != CLS_OTHER) break;
      for (uint32_t i = 0; i < len; i++) {
        if (klen < WP_KEY_BYTES) {
          kb[klen++] = tv.t[q + i];
        } else {
          more = true;
        }
      }
      if (more) break;
      if (first && cls == CLS_PUNCT) break;
      first = false;
      q += len;
    }
    if (!more && klen == WP_KEY_BYTES && q < tv.n) {
      // exactly 22 bytes gathered: does the window go on?
      uint32_t len, cls;
      const size_t q2 = gnext(tv, q, &len, &cls);
      more = q2 < tv.n && cls == CLS_OTHER;
    }
  }
  uint32_t r[6];
#pragma unroll
  for (int i = 0; i < 6; i++)
    r[i] = uint32_t(kb[4 * i]) | (uint32_t(kb[4 * i + 1]) << 8) | (uint32_t(kb[4 * i + 2]) << 16) |
           (uint32_t(kb[4 * i + 3]) << 24);
  NodeHit node;
  const uint32_t d = deepest_node(V, r, klen, kind, &node);
  if (d == 0) return 0;
  if (d == WP_KEY_BYTES && slot_has_long(node.w5) && more) {
    const uint32_t ref = V.long_ref[node.slot];
    const uint32_t cnt = V.long_entries[ref];
    for (uint32_t e = 0; e < cnt; e++) {
      const uint32_t len = V.long_entries[ref + 1 + 3 * e];
      const uint8_t *tok = V.long_bytes + V.long_entries[ref + 3 + 3 * e];
      // compare clean bytes [0,len) of the window with the token
      uint32_t o = 0;
      size_t q = p;
      bool ok = true, first = true;
      while (o < len) {
        uint32_t clen, cls;
        q = gnext(tv, q, &clen, &cls);
        if (q >= tv.n || (!first && cls != CLS_OTHER)) {
          ok = false;
          break;
        }
        for (uint32_t i = 0; i < clen && ok; i++) {
          if (o >= len || tv.t[q + i] != tok[o]) ok = false;
          o++;
        }
        if (!ok) break;
        first = false;
        q += clen;
      }
      if (ok && o == len) {
        *id = static_cast<int32_t>(V.long_entries[ref + 2 + 3 * e]);
        return len;
      }
    }
  }
  if (node.term_id != WP_NO_ID) {
    *id = node.term_id;
    return d;
  }
  const uint32_t bl = slot_best_len(node.w5);
  if (bl != 0) {
    *id = node.best_id;
    return bl;
  }
  return 0;
}

// Walk the segment that starts at the valid, non-space lead gs.  out == nullptr:
// count only.  Otherwise ids go to out[0..) (bounded by cap_left); ids at
// indices >= unk_at are not written (they are rolled back, fast.cpp:80-84) and
// the UNK id is written at unk_at.  Returns the id count; *unk_at_out = index
// of the final UNK or -1.
__device__ uint32_t walk_segment(const DeviceVocab &V, const TextView &tv, size_t gs, int32_t *out,
                                 unsigned long long cap_left, int32_t unk_at, int32_t *unk_at_out) {
  uint32_t len0, cls0;
  gnext(tv, gs, &len0, &cls0);
  int32_t id = 0;
  auto emit = [&](uint32_t index, int32_t v, bool is_unk) {
    if (out == nullptr) return;
    if (!is_unk && unk_at >= 0 && index >= static_cast<uint32_t>(unk_at)) return;
    if (index < cap_left) out[index] = v;
  };
  *unk_at_out = -1;
  if (cls0 == CLS_PUNCT) {
    const uint32_t k = longest_match_global(V, tv, gs, WP_KIND_PREFIX, &id);
    if (k == len0) {
      emit(0, id, false);
    } else {
      *unk_at_out = 0;
      emit(0, V.unk_id, true);
    }
    return 1;
  }
  size_t p = gs;
  uint32_t n = 0, word_first = 0, kind = WP_KIND_PREFIX;
  bool at_han = (cls0 == CLS_HAN);
  for (;;) {
    const uint32_t k = longest_match_global(V, tv, p, kind, &id);
    if (k == 0) {
      if (at_han && !V.han_swallow) {
        // max_len < 2: the Han char alone is UNK, the run after it is a new word
        emit(n, V.unk_id, true);  // not rolled back later: word_first moves past it
        n += 1;
        word_first = n;
        at_han = false;
        uint32_t l, c;
        p = gnext(tv, p + len0, &l, &c);
        if (p >= tv.n || c != CLS_OTHER) return n;
        continue;
      }
      *unk_at_out = static_cast<int32_t>(word_first);
      emit(word_first, V.unk_id, true);
      return word_first + 1;
    }
    emit(n, id, false);
    n += 1;
    // advance k clean bytes
    uint32_t o = 0, l = 0, c = CLS_SPACE;
    while (o < k) {
      p = gnext(tv, p, &l, &c);
      o += l;
      p += l;
    }
    p = gnext(tv, p, &l, &c);
    if (p >= tv.n || c != CLS_OTHER) return n;
    if (at_han && k == len0) {
      word_first = n;
      kind = WP_KIND_PREFIX;
    } else {
      kind = WP_KIND_SUFFIX;
    }
    at_han = false;
  }
}

// ------------------------------------------------------------- classification

// exact: 0x80 in every byte lane whose byte is zero
__device__ __forceinline__ uint32_t swar_zero(uint32_t x) {
  return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
}

// flags of the byte lane d (1..3) positions earlier: lane j gets cur/prev lane j-d
__device__ __forceinline__ uint32_t lanes_back(uint32_t prev, uint32_t cur, int d) {
  return __funnelshift_l(prev, cur, 8 * d);
}

struct WordFlags {
  uint32_t cont;     // 10xxxxxx
  uint32_t m1;       // 11xxxxxx  (any multi-byte lead)
  uint32_t m2;       // 111xxxxx
  uint32_t m3;       // 1111xxxx
  uint32_t suspect;  // leads whose validity depends on their value: C0 C1 E0 ED F0..FF
};

__device__ __forceinline__ WordFlags word_flags(uint32_t w) {
  WordFlags f;
  const uint32_t hi = w & 0x80808080u;
  const uint32_t t1 = w << 1, t2 = w << 2, t3 = w << 3;
  f.cont = hi & ~t1;
  f.m1 = hi & t1;
  f.m2 = f.m1 & t2;
  f.m3 = f.m2 & t3;
  const uint32_t lead2 = f.m1 & ~t2;
  const uint32_t lead3 = f.m2 & ~t3;
  const uint32_t low = w & 0x0F0F0F0Fu;
  f.suspect = f.m3 | (lead2 & swar_zero(w & 0x1E1E1E1Eu)) | (lead3 & (swar_zero(low) | swar_zero(low ^ 0x0D0D0D0Du)));
  return f;
}

// Classify chunk c of `buf` (window coordinates) into the mask arrays.  `limit`
// is the number of meaningful bytes in buf (positions >= limit are ignored).
//
// Fast lane (no per-byte loop): ASCII classes by SWAR range tests; multi-byte
// text is validated STRUCTURALLY (every lead followed by exactly its
// continuation bytes) with byte-lane shifts; only leads that can be space /
// punctuation / Han (C2, E2..E9, EF) are decoded.  Chunks holding a lead whose
// validity depends on its value (overlong / surrogate / 4-byte forms) or any
// structural error take the exact per-byte lane below.
__device__ __forceinline__ void classify_chunk(TileSmem &sm, const uint8_t *buf, int c, int limit) {
  const uint8_t *cb = buf + c * CHUNK;
  uint32_t lead = 0xFFFFFFFFu, sp = 0, pu = 0, ha = 0, cover = 0xFFFFFFFFu, spill = 0;
  uint32_t any_high = 0;
  // (the loops over the chunk's eight words are deliberately NOT fully unrolled: K1's executed code must fit
  // the SM's 32 KB instruction cache — six tiles in different phases share it — and these two loops alone
  // were a sixth of it; the words are re-read from shared memory where they are needed)
#pragma unroll 2
  for (int i = 0; i < 8; i++) {
    const uint32_t wi = ld_u32(cb + 4 * i);
    const uint32_t w7 = wi & 0x7F7F7F7Fu;
    const uint32_t asc = ~wi & 0x80808080u;
    const uint32_t s = (swar_range(w7, 0x09, 0x0D) | swar_range(w7, 0x20, 0x20)) & asc;
    const uint32_t q = (swar_range(w7, 0x21, 0x2F) | swar_range(w7, 0x3A, 0x40) | swar_range(w7, 0x5B, 0x60) |
                        swar_range(w7, 0x7B, 0x7E)) & asc;
    sp |= swar_nibble(s) << (4 * i);
    pu |= swar_nibble(q) << (4 * i);
    any_high |= wi;
  }
  if (any_high & 0x80808080u) {
    // ---- structural validation in the byte-lane domain
    WordFlags prev = word_flags(ld_u32(cb - 4));
    uint32_t bad = prev.suspect, contm = 0, cand = 0;
#pragma unroll 1
    for (int i = 0; i < 8; i++) {
      const uint32_t wi = ld_u32(cb + 4 * i);
      const WordFlags f = word_flags(wi);
      const uint32_t expect = lanes_back(prev.m1, f.m1, 1) | lanes_back(prev.m2, f.m2, 2) | lanes_back(prev.m3, f.m3, 3);
      bad |= (expect ^ f.cont) | f.suspect;
      contm |= swar_nibble(f.cont) << (4 * i);
      // leads that may be a spacing char: C2 (Latin-1 punctuation), E2 (U+2010.., U+2581), E3..E9, EF (Han)
      const uint32_t low = wi & 0x0F0F0F0Fu;
      const uint32_t lead3 = f.m2 & ~(wi << 3);
      const uint32_t cf = swar_zero(wi ^ 0xC2C2C2C2u) | (lead3 & (swar_range(low, 2, 9) | swar_zero(low ^ 0x0F0F0F0Fu)));
      cand |= swar_nibble(cf) << (4 * i);
      prev = f;
    }
    {
      // sequences that run past the chunk must find their continuation bytes in the next word
      const WordFlags nx = word_flags(ld_u32(cb + CHUNK));
      const uint32_t expect = lanes_back(prev.m1, 0u, 1) | lanes_back(prev.m2, 0u, 2) | lanes_back(prev.m3, 0u, 3);
      bad |= expect & ~nx.cont;
      spill = __popc(expect);
    }
    if (bad == 0) {
      lead = ~contm;
      while (cand) {
        const int j = __ffs(cand) - 1;
        cand &= cand - 1;
        const uint32_t b0 = cb[j], b1 = cb[j + 1], b2 = cb[j + 2];
        const uint32_t cp = b0 < 0xE0u ? (((b0 & 0x1Fu) << 6) | (b1 & 0x3Fu))
                                       : (((b0 & 0x0Fu) << 12) | ((b1 & 0x3Fu) << 6) | (b2 & 0x3Fu));
        const uint32_t cls = cp_class(cp);
        sp |= (cls == CLS_SPACE ? 1u : 0u) << j;
        pu |= (cls == CLS_PUNCT ? 1u : 0u) << j;
        ha |= (cls == CLS_HAN ? 1u : 0u) << j;
      }
    } else {
      // ---- exact per-byte lane (utf8.cpp:54-90): overlongs, surrogates, 4-byte forms, stray bytes
      lead = 0;
      cover = 0;
      spill = 0;
      sp = 0;
      pu = 0;
      for (int j = 0; j < CHUNK; j++) {
        const uint32_t b0 = cb[j];
        if (is_cont_byte(b0)) continue;
        uint32_t cp = 0;
        const uint32_t len = b0 < 0x80u ? (cp = b0, 1u) : utf8_decode(b0, cb[j + 1], cb[j + 2], cb[j + 3], 4u, &cp);
        if (len == 0) continue;
        lead |= 1u << j;
        cover |= ((1u << len) - 1u) << j;
        if (j + static_cast<int>(len) > CHUNK) spill = j + len - CHUNK;
        const uint32_t cls = cp_class(cp);
        sp |= (cls == CLS_SPACE ? 1u : 0u) << j;
        pu |= (cls == CLS_PUNCT ? 1u : 0u) << j;
        ha |= (cls == CLS_HAN ? 1u : 0u) << j;
      }
    }
  }
  // ignore everything at or past `limit`
  const int left = limit - c * CHUNK;
  const uint32_t in = left >= CHUNK ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
  sm.m_lead[c] = lead & in;
  sm.m_space[c] = sp & in;
  sm.m_punct[c] = pu & in;
  sm.m_han[c] = ha & in;
  sm.m_cover[c] = cover;
  sm.spill[c] = static_cast<uint8_t>(spill);
}

// block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total = sum
template <int NWARPS>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t *warp_sums, uint32_t v, uint32_t *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += y;
  }
  __syncthreads();  // warp_sums may still be in use by an earlier scan
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0, tot = 0;
#pragma unroll
  for (int wi = 0; wi < NWARPS; wi++) {
    const uint32_t ws = warp_sums[wi];
    if (wi < warp) wbase += ws;
    tot += ws;
  }
  *total = tot;
  return wbase + incl - v;
}

// K1 calls the scan from three places; one out-of-line copy keeps its executed code inside the instruction
// cache (see classify_chunk).  Returns the exclusive prefix in the low and the total in the high 32 bits.
__device__ __noinline__ unsigned long long tile_exclusive_scan(uint32_t *warp_sums, uint32_t v) {
  uint32_t total;
  const uint32_t at = block_exclusive_scan<WARPS>(warp_sums, v, &total);
  return (static_cast<unsigned long long>(total) << 32) | at;
}

// Decoupled look-back over a chain of units (tiles of K1, blocks of K3).  state[i] = flag << 62 | value;
// flag 1 = the unit's own total, 2 = inclusive prefix.  Units are handed out in launch order by a
// ticket, so every predecessor is already running: the spin always ends.
//
// lookback_publish (one thread) makes the unit's total visible as early as possible; lookback_walk (one
// full warp) later returns the sum over all earlier units and publishes the inclusive prefix.  Work that
// does not need the prefix goes in between: by then the predecessors have published theirs and the
// walk is short (when every unit walked right away, it crossed hundreds of concurrent units).
__device__ __forceinline__ void lookback_publish(volatile unsigned long long *state, uint32_t index,
                                                 unsigned long long total) {
  state[index] = ((index == 0 ? 2ull : 1ull) << 62) | total;
}

__device__ __forceinline__ unsigned long long lookback_walk(volatile unsigned long long *state, uint32_t index,
                                                            unsigned long long total, int lane) {
  constexpr unsigned long long VALUE_MASK = (1ull << 62) - 1;
  unsigned long long base = 0;
  if (index == 0) return 0;
  long long pred = static_cast<long long>(index) - 1 - lane;  // lane i looks at unit index-1-i
  for (;;) {
    unsigned long long sv = 2ull << 62;  // units before 0 count as a zero prefix
    if (pred >= 0) {
      do {
        sv = state[pred];
      } while ((sv >> 62) == 0);
    }
    const uint32_t is_prefix = __ballot_sync(FULL, (sv >> 62) == 2);
    // add the totals of the lanes before the first inclusive prefix, and that prefix
    const int stop = is_prefix ? __ffs(is_prefix) - 1 : 31;
    unsigned long long v = lane <= stop ? (sv & VALUE_MASK) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    base += v;
    if (is_prefix) break;
    pred -= 32;
  }
  if (lane == 0) state[index] = (2ull << 62) | (base + total);
  return base;
}

// 24 window bytes starting at p (any alignment) as six little-endian words
__device__ __forceinline__ void load_window(const uint8_t *buf, int p, uint32_t r[6]) {
  const int a = p & ~3;
  const uint32_t sh = (p & 3) * 8;
  uint32_t x[7];
#pragma unroll
  for (int i = 0; i < 7; i++) x[i] = ld_u32(buf + a + 4 * i);
#pragma unroll
  for (int i = 0; i < 6; i++) r[i] = __funnelshift_r(x[i], x[i + 1], sh);
}

// the same from the text in global memory (bytes past the end read as spaces)
__device__ __forceinline__ void load_window_global(const uint8_t *text, size_t n_bytes, size_t pos, uint32_t r[6]) {
  const uint8_t *addr = text + pos;
  if (pos + 32 <= n_bytes) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(addr) & ~static_cast<uintptr_t>(3);
    const uint32_t sh = (reinterpret_cast<uintptr_t>(addr) & 3u) * 8u;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a);
    uint32_t x[7];
#pragma unroll
    for (int i = 0; i < 7; i++) x[i] = __ldg(w + i);
#pragma unroll
    for (int i = 0; i < 6; i++) r[i] = __funnelshift_r(x[i], x[i + 1], sh);
  } else {
#pragma unroll
    for (int i = 0; i < 6; i++) {
      uint32_t v = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const size_t q = pos + 4 * i + b;
        v |= (q < n_bytes ? static_cast<uint32_t>(text[q]) : 0x20u) << (8 * b);
      }
      r[i] = v;
    }
  }
}

__device__ __forceinline__ void init_key_mask(uint4 *key_mask, int tid) {
  if (tid <= static_cast<int>(WP_KEY_BYTES)) {
    const uint32_t k = tid;  // row k of the key mask table
    uint32_t m[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const int nb = static_cast<int>(k) - 4 * i;
      m[i] = nb >= 4 ? 0xFFFFFFFFu : (nb <= 0 ? 0u : ((1u << (8 * nb)) - 1u));
    }
    key_mask[KEY_MASK_ROW * k] = make_uint4(m[0], m[1], m[2], m[3]);
    key_mask[KEY_MASK_ROW * k + 1] = make_uint4(m[4], m[5] & 0xFFFFu, k << 16, 0u);
  }
}

// ================================================================ K1: split

// A single-char segment whose home slot holds another key: walk the probe sequence to the key or to an
// empty slot.  Rare (a few lanes per tile at most), so kept out of line; returns the node's term id
// (possibly WP_NO_ID) or SINGLE_WALK_MISS.
constexpr int32_t SINGLE_WALK_MISS = WP_NO_ID - 1;
constexpr uint16_t PARK_PENDING = 0xFFFFu;  // parked high half of a result is < 0x4000 (results are < 2^30)
__device__ __noinline__ int32_t single_char_walk(const uint4 *tab, uint32_t slot_mask, uint32_t k0, uint32_t k1,
                                                 uint32_t k2, uint32_t k3, uint32_t k4, uint32_t k5) {
  const uint32_t kw[6] = {k0, k1, k2, k3, k4, k5};
  uint32_t idx = (key_hash(k0, k1, k2, k3, k4, k5) + 1) & slot_mask;
  for (;;) {
    uint4 a, b;
    ld_slot(tab, idx, &a, &b);
    if (slot_len(b.y) == 0) return SINGLE_WALK_MISS;
    if (slot_matches(a, b, kw)) return static_cast<int32_t>(b.z);
    idx = (idx + 1) & slot_mask;
  }
}

__global__ void __launch_bounds__(THREADS) wp_split_kernel(EncodeParams P) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  TileSmem &sm = *reinterpret_cast<TileSmem *>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const DeviceVocab &V = P.vocab;

  // tile = block index: blocks of a 1-D grid are dispatched in index order, so a tile's predecessors are
  // always already running (decoupled look-back needs forward progress); no ticket round trip
  if (tid == 0) {
    sm.left_spill = 0;
    sm.n_slow = 0;
    sm.n_slow2 = 0;
    sm.memo_hits = 0;
    sm.any_single = 0;
  }
  if (tid == 32) {
    // whether the memo is still worth its lookups was decided by K2 of the previous range (uniform over this
    // range: the phase has barriers); read here, next to the ticket, so that the round trip is hidden
    sm.memo_use = P.memo != nullptr && (P.range_index < 2 || P.call->memo_off == 0u);
  }
  init_key_mask(sm.key_mask, tid);
  __syncthreads();
  const uint32_t rel_tile = blockIdx.x;                    // within the range
  const size_t t0 = (static_cast<size_t>(P.first_tile) + rel_tile) * TILE;
  const size_t n = P.n_bytes;
  const size_t avail = n - t0;  // > 0
  const bool more_text = avail > static_cast<size_t>(WINDOW);
  const TextView tv{P.text, n};
  {
    // tiles run in index order, about 6 x 148 at a time: pull the tile two such waves ahead into L2 so that
    // its own loads do not wait on DRAM
    const size_t pf = t0 + static_cast<size_t>(PREFETCH_TILES) * TILE + static_cast<size_t>(tid) * 128u;
    if (tid < TILE / 128 && pf < n) prefetch_l2(P.text + pf);
  }

  // ---- S1a: stage raw bytes [t0-LEFT, t0+WINDOW+LOOKAHEAD) in shared memory
  {
    const bool aligned = (reinterpret_cast<uintptr_t>(P.text) & 15u) == 0;
#pragma unroll 1
    for (int u = tid; u < RAW_BYTES / 16; u += THREADS) {
      const long long g = static_cast<long long>(t0) - LEFT + 16ll * u;  // text offset of this unit
      uint4 v;
      if (aligned && g >= 0 && static_cast<size_t>(g) + 16 <= n) {
        v = ldg_stream(reinterpret_cast<const uint4 *>(P.text + g));
      } else {
        uint32_t q[4] = {0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u};
        for (int i = 0; i < 16; i++) {
          const long long gi = g + i;
          if (gi >= 0 && static_cast<size_t>(gi) < n) {
            q[i >> 2] = (q[i >> 2] & ~(0xFFu << (8 * (i & 3)))) | (uint32_t(P.text[gi]) << (8 * (i & 3)));
          }
        }
        v = make_uint4(q[0], q[1], q[2], q[3]);
      }
      *reinterpret_cast<uint4 *>(sm.raw + 16 * u) = v;
    }
  }
  __syncthreads();
  if (tid == THREADS - 1) {
    // class of the last valid char before the tile: from the left halo in shared memory when it holds one
    // (always, for valid UTF-8), else by walking back through the text in global memory
    uint32_t pc = CLS_SPACE;
    bool found = t0 == 0;
    if (!found) {
      const uint8_t *h = sm.raw + LEFT;  // h[-1] is the byte before the tile
      for (int j = -1; j >= -4 && !found; j--) {
        const uint32_t b0 = h[j];
        if (is_cont_byte(b0)) continue;
        uint32_t cp = 0;
        const uint32_t len = b0 < 0x80u ? (cp = b0, 1u) : utf8_decode(b0, h[j + 1], h[j + 2], h[j + 3], 4u, &cp);
        if (len != 0) {
          pc = cp_class(cp);
          found = true;
        }
        break;  // an invalid lead: the exact answer needs the walk below
      }
      if (!found) pc = gprev_class(tv, t0);
    }
    sm.prev_class = pc;
  }

  // ---- S1b: classify the window; find bytes that the strict decoder drops
  uint8_t *const buf = sm.raw + LEFT;
  int limit = WINDOW;
  if (tid < NCHUNK) classify_chunk(sm, buf, tid, WINDOW);
  if (tid == THREADS - 2) {
    // a sequence that starts in the last 3 bytes before the tile may own its first bytes
    uint32_t ls = 0;
    for (int j = -3; j < 0; j++) {
      const uint32_t b0 = buf[j];
      if (is_cont_byte(b0) || b0 < 0x80u) continue;
      uint32_t cp;
      const uint32_t len = utf8_decode(b0, buf[j + 1], buf[j + 2], buf[j + 3], 4u, &cp);
      if (len && j + static_cast<int>(len) > 0) ls = j + len;
    }
    sm.left_spill = ls;
  }
  __syncthreads();
  uint32_t kept = 0xFFFFFFFFu;  // surviving bytes of chunk `tid` (chunk NCHUNK = look-ahead: only a spilled tail)
  if (tid <= NCHUNK) {
    const uint32_t sp_in = tid == 0 ? sm.left_spill : sm.spill[tid - 1];
    kept = (tid < NCHUNK ? sm.m_cover[tid] : 0u) | ((1u << sp_in) - 1u);
  }
  const bool dirty = __syncthreads_or(tid < NCHUNK && kept != 0xFFFFFFFFu);

  if (dirty) {
    // ---- S1c (rare): drop the invalid bytes by compacting the window IN PLACE
    // (every thread first pulls its chunk into registers), then classify again.
    // utf8.cpp:130-147.
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t my_cnt = 0;
    if (tid <= NCHUNK) {
      const uint4 a = *reinterpret_cast<const uint4 *>(buf + tid * CHUNK);
      const uint4 b = *reinterpret_cast<const uint4 *>(buf + tid * CHUNK + 16);
      w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
      w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
      my_cnt = __popc(kept);
      sm.m_kept[tid] = kept;
    }
    uint32_t packed_len;
    const unsigned long long sc0 = tile_exclusive_scan(sm.warp_sums, my_cnt);  // syncs: chunks are in registers
    const uint32_t dst0 = static_cast<uint32_t>(sc0);
    packed_len = static_cast<uint32_t>(sc0 >> 32);
    if (tid <= NCHUNK) {
      sm.kept_scan[tid] = dst0;
      if (tid == NCHUNK) sm.kept_scan[NCHUNK + 1] = packed_len;
      uint32_t dst = dst0, m = kept;
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        buf[dst++] = static_cast<uint8_t>(w[j >> 2] >> (8 * (j & 3)));
      }
    }
    __syncthreads();
    for (int i = static_cast<int>(packed_len) + tid; i < WINDOW + LOOKAHEAD; i += THREADS) buf[i] = 0x20;
    __syncthreads();
    limit = static_cast<int>(packed_len) < WINDOW ? static_cast<int>(packed_len) : WINDOW;
    if (tid < NCHUNK) classify_chunk(sm, buf, tid, limit);
    __syncthreads();
  }
  // owned range in buffer coordinates: segments that start in [0, own_end)
  const int own_end = dirty ? static_cast<int>(sm.kept_scan[OWNED_CHUNKS]) : TILE;

  // ---- S1d: segment starts (SURVEY A.2 safe starts) and segment ends, by mask
  // arithmetic, compacted in text order into seg_s[] / seg_e[].  A segment ends
  // at a lead p whose previous char is not a space and (is punctuation, or p
  // itself is a spacing char).  Starts and ends alternate, so owned segment k
  // pairs with end k + skip, skip = 1 iff a segment of the previous tile is
  // still open at the tile border.
  {
    uint32_t starts = 0, ends = 0, pu = 0, ha = 0;
    const int c = tid;
    if (c < NCHUNK) {
      const uint32_t lead = sm.m_lead[c];
      pu = sm.m_punct[c];
      ha = sm.m_han[c];
      const uint32_t sp = sm.m_space[c];
      uint32_t carry_s, carry_p;
      if (c == 0) {
        carry_s = sm.prev_class == CLS_SPACE;
        carry_p = sm.prev_class == CLS_PUNCT;
      } else {
        const uint32_t pl = sm.m_lead[c - 1];
        uint32_t xs = sm.m_space[c - 1], xp = sm.m_punct[c - 1];
#pragma unroll
        for (int i = 0; i < 3; i++) {
          xs |= (xs << 1) & ~pl;
          xp |= (xp << 1) & ~pl;
        }
        carry_s = xs >> 31;
        carry_p = xp >> 31;
      }
      uint32_t xs = sp | (carry_s & ~lead & 1u), xp = pu | (carry_p & ~lead & 1u);
#pragma unroll
      for (int i = 0; i < 3; i++) {
        xs |= (xs << 1) & ~lead;
        xp |= (xp << 1) & ~lead;
      }
      const uint32_t prev_s = (xs << 1) | carry_s;
      const uint32_t prev_p = (xp << 1) | carry_p;
      ends = lead & ~prev_s & (prev_p | sp | pu | ha);
      starts = lead & ~sp & (pu | ha | prev_s | prev_p);
      const int left = own_end - c * CHUNK;
      starts &= left >= CHUNK ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
    }
    uint32_t totals;
    const unsigned long long sc1 = tile_exclusive_scan(sm.warp_sums, __popc(starts) | (__popc(ends) << 16));
    const uint32_t at = static_cast<uint32_t>(sc1);
    totals = static_cast<uint32_t>(sc1 >> 32);
    uint32_t at_s = at & 0xFFFFu, at_e = at >> 16;
    while (starts) {
      const int j = __ffs(starts) - 1;
      starts &= starts - 1;
      const uint32_t bit = 1u << j;
      const uint32_t cls = (pu & bit) ? CLS_PUNCT : ((ha & bit) ? CLS_HAN : CLS_OTHER);
      sm.seg_s[at_s++] = static_cast<uint16_t>((c * CHUNK + j) | (cls << 14));
    }
    while (ends) {
      const int j = __ffs(ends) - 1;
      ends &= ends - 1;
      sm.seg_e[at_e++] = static_cast<uint16_t>(c * CHUNK + j);
    }
    if (tid == 0) {
      sm.n_segs = totals & 0xFFFFu;
      sm.n_ends = totals >> 16;
    }
  }
  __syncthreads();
  const uint32_t n_segs = sm.n_segs;
  const uint32_t n_ends = sm.n_ends;
  const uint32_t skip = sm.prev_class != CLS_SPACE ? 1u : 0u;

  // ---- segment numbering across tiles: publish this tile's segment count now, walk back after S2a
  if (tid == 0) {
    lookback_publish(P.tile_state, rel_tile, n_segs);
    if (dirty) atomicAdd(&P.call->dirty_tiles, 1ull);
  }
  const uint4 *tab = reinterpret_cast<const uint4 *>(V.slots);

  // ---- S2a: one whole-window probe per segment, statically assigned (uniform
  // work: every lane does the same thing; two segments per lane and turn so that
  // two table loads are in flight).  It settles every segment that is a single
  // token (fast.cpp:66-72 hit on the first, longest candidate) and every
  // single-char segment; the rest — and the rare probe that lands on another
  // key's slot — go to the slow list.
  constexpr int PER_TURN = 2;
  for (uint32_t base = 0; base < n_segs; base += PER_TURN * THREADS) {
    uint32_t kk[PER_TURN], slow[PER_TURN], wlen[PER_TURN], first_len[PER_TURN], kw[PER_TURN][6];
    uint4 sa[PER_TURN], sb[PER_TURN];
#pragma unroll
    for (int u = 0; u < PER_TURN; u++) {
      const uint32_t k = base + u * THREADS + tid;
      kk[u] = k;
      slow[u] = 0;  // 0 = settled (or no segment), else the flags of the tile slow-list entry | 1
      wlen[u] = 0;
      first_len[u] = 0;
      sa[u] = make_uint4(0, 0, 0, 0);
      sb[u] = make_uint4(0, 0, 0, 0);
      if (k < n_segs) {
        const uint32_t sv = sm.seg_s[k];
        const int s = static_cast<int>(sv & POS_MASK);
        const uint32_t j = k + skip;
        int e = limit;
        if (j < n_ends) {
          e = sm.seg_e[j];
        } else if (more_text) {  // leaves the window: walked from global memory in K2
          slow[u] = SLOW_WALK | 1u;
        }
        if (!slow[u]) {
          wlen[u] = static_cast<uint32_t>(e - s);
          first_len[u] = utf8_lead_len(buf[s]);
          const uint32_t k0 = wlen[u] < WP_KEY_BYTES ? wlen[u] : WP_KEY_BYTES;
          uint32_t r[6];
          load_window(buf, s, r);
          make_key_tab(sm.key_mask, r, k0, WP_KIND_PREFIX, kw[u]);
          const uint32_t idx = key_hash(kw[u][0], kw[u][1], kw[u][2], kw[u][3], kw[u][4], kw[u][5]) & V.slot_mask;
          ld_slot(tab, idx, &sa[u], &sb[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < PER_TURN; u++) {
      bool settled = false;
      if (wlen[u] != 0) {
        // one slot only here (K2 looks at pairs): a probe that lands on another key's slot is left to K2 —
        // except for single-char segments: every segment handed to K2 must have at least two bytes (the
        // slow-list capacities rely on it), so those are marked and settled by the pass below (rare)
        const bool empty = slot_len(sb[u].y) == 0;
        const bool match = !empty && slot_matches(sa[u], sb[u], kw[u]);
        const int32_t term = static_cast<int32_t>(sb[u].z);
        const bool hit = match && term != WP_NO_ID && wlen[u] <= WP_KEY_BYTES;
        const bool single = wlen[u] == first_len[u];  // no token (empty), only a prefix of tokens, or a collision
        if (hit || single) {
          // settled: park the result (id + 1) in the two list entries of the segment, which are no longer
          // needed, until the tile knows its first global segment number
          const uint32_t res = static_cast<uint32_t>((hit ? term : V.unk_id) + 1);
          if (single && !empty && !match) {
            sm.seg_e[kk[u] + skip] = PARK_PENDING;  // seg_s keeps the position for the pass below
            sm.any_single = 1u;
          } else {
            sm.seg_s[kk[u]] = static_cast<uint16_t>(res);
            sm.seg_e[kk[u] + skip] = static_cast<uint16_t>(res >> 16);
          }
          settled = true;
        } else {
          // miss on an empty slot: K2 may skip the whole-window probe; match or collision: K2 redoes it
          slow[u] = (empty ? SLOW_FIRST_MISSED : 0u) | 1u;
        }
      }
      const uint32_t settledm = __ballot_sync(FULL, settled);
      if (lane == 0 && base + u * THREADS + (tid & ~31) < n_segs) sm.settled[(base + u * THREADS + tid) >> 5] = settledm;
      const uint32_t slowm = __ballot_sync(FULL, slow[u] != 0);
      if (slowm) {
        uint32_t at = 0;
        const int leader = __ffs(slowm) - 1;
        if (lane == leader) at = smem_add(&sm.n_slow, static_cast<uint32_t>(__popc(slowm)));
        at = __shfl_sync(FULL, at, leader);
        if (slow[u]) sm.slow[at + __popc(slowm & ((1u << lane) - 1u))] = static_cast<uint16_t>(kk[u] | (slow[u] & ~1u));
      }
    }
  }
  __syncthreads();

  // ---- (rare) single-char segments whose home slot holds another key: follow the probe sequence
  if (sm.any_single) {  // uniform
#pragma unroll 1
    for (uint32_t k = tid; k < n_segs; k += THREADS) {
      if (!((sm.settled[k >> 5] >> (k & 31)) & 1u) || sm.seg_e[k + skip] != PARK_PENDING) continue;
      uint32_t r[6], key[6];
      load_window(buf, static_cast<int>(sm.seg_s[k] & POS_MASK), r);
      make_key_tab(sm.key_mask, r, utf8_lead_len(r[0] & 0xFFu), WP_KIND_PREFIX, key);
      const int32_t term = single_char_walk(tab, V.slot_mask, key[0], key[1], key[2], key[3], key[4], key[5]);
      const uint32_t res = static_cast<uint32_t>((term >= 0 ? term : V.unk_id) + 1);
      sm.seg_s[k] = static_cast<uint16_t>(res);
      sm.seg_e[k + skip] = static_cast<uint16_t>(res >> 16);
    }
  }

  // ---- the walk back over earlier tiles (short by now), then the settled results go out coalesced
  if (warp == 0) {
    const unsigned long long base = lookback_walk(P.tile_state, rel_tile, n_segs, lane);
    if (lane == 0) {
      sm.seg_base = base;
      if (rel_tile == P.n_tiles - 1) P.counters->n_segs = base + n_segs;
    }
  }
  __syncthreads();
  const unsigned long long seg_base = sm.seg_base;
  if (seg_base + n_segs > P.seg_capacity) {
    if (tid == 0) P.call->overflow = 1u;
    return;  // uniform
  }
#pragma unroll 1
  for (uint32_t k = tid; k < n_segs; k += THREADS) {
    if ((sm.settled[k >> 5] >> (k & 31)) & 1u)
      P.seg_result[seg_base + k] = static_cast<uint32_t>(sm.seg_s[k]) | (static_cast<uint32_t>(sm.seg_e[k + skip]) << 16);
  }

  // ---- word memo: an unsettled segment of at most 16 bytes whose exact bytes were matched before (by K2,
  // in an earlier range of this call) is settled here with one lookup; the rest form the final slow list
  uint32_t n_slow = sm.n_slow;
  if (n_slow == 0) return;  // uniform
  if (sm.memo_use && !dirty) {  // uniform
    // two entries per lane and round, all lookups of a round in flight together; the list is compacted in
    // place (a round's survivors land below the entries the next round reads)
    constexpr int MP = 2;
    uint32_t my_lookups = 0;
    for (uint32_t base = 0; base < n_slow; base += MP * THREADS) {
      uint32_t ent[MP], mlen[MP], midx[MP], mk[MP][4];
      uint4 ma[MP], mb[MP];
      bool keep[MP], have[MP];
#pragma unroll
      for (int u = 0; u < MP; u++) {
        const uint32_t i = base + u * THREADS + tid;
        have[u] = i < n_slow;
        keep[u] = have[u];
        ent[u] = 0;
        mlen[u] = 0;
        midx[u] = 0;
        ma[u] = make_uint4(0, 0, 0, 0);
        mb[u] = make_uint4(0, 0, 0, 0);
        if (have[u]) {
          ent[u] = sm.slow[i];
          my_lookups++;  // counted per unsettled segment, eligible or not: the memo must pay for the whole slow lane
          if (!(ent[u] & SLOW_WALK)) {
            const uint32_t k = ent[u] & 0xFFFu;
            const int s = static_cast<int>(sm.seg_s[k] & POS_MASK);
            const uint32_t j = k + skip;
            const int e = j < n_ends ? static_cast<int>(sm.seg_e[j]) : limit;
            const uint32_t len = static_cast<uint32_t>(e - s);
            if (len <= MEMO_KEY_BYTES) {
              uint32_t r[6];
              load_window(buf, s, r);
              const uint4 km = sm.key_mask[KEY_MASK_ROW * len];
              mk[u][0] = r[0] & km.x; mk[u][1] = r[1] & km.y; mk[u][2] = r[2] & km.z; mk[u][3] = r[3] & km.w;
              midx[u] = key_hash(mk[u][0], mk[u][1], mk[u][2], mk[u][3], len, MEMO_SALT) & P.memo_mask;
              ma[u] = __ldg(P.memo + 2 * static_cast<size_t>(midx[u]));
              mb[u] = __ldg(P.memo + 2 * static_cast<size_t>(midx[u]) + 1);
              mlen[u] = len;
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < MP; u++) {
        if (mlen[u] == 0) continue;
        for (int t = 0;; t++) {
          if (mb[u].x == 0) break;
          if ((mb[u].x & MEMO_READY) && (mb[u].x & 0xFFu) == mlen[u] && ma[u].x == mk[u][0] && ma[u].y == mk[u][1] &&
              ma[u].z == mk[u][2] && ma[u].w == mk[u][3]) {
            P.seg_result[seg_base + (ent[u] & 0xFFFu)] =
                SEG_RESULT_MEMO | (((mb[u].x >> 8) & 3u) << SEG_MEMO_SLOT_BITS) | midx[u];
            keep[u] = false;
            break;
          }
          if (t == 1) break;  // K2 inserts within a few slots of home; two are looked at here
          midx[u] = (midx[u] + 1) & P.memo_mask;
          ma[u] = __ldg(P.memo + 2 * static_cast<size_t>(midx[u]));
          mb[u] = __ldg(P.memo + 2 * static_cast<size_t>(midx[u]) + 1);
        }
      }
      __syncthreads();  // every entry of this round has been read
#pragma unroll
      for (int u = 0; u < MP; u++) {
        const uint32_t keepm = __ballot_sync(FULL, keep[u]);
        const uint32_t hitm = __ballot_sync(FULL, have[u] && !keep[u]);
        if (keepm) {
          uint32_t at = 0;
          const int leader = __ffs(keepm) - 1;
          if (lane == leader) {
            at = smem_add(&sm.n_slow2, static_cast<uint32_t>(__popc(keepm)));
            if (hitm) smem_add_noret(&sm.memo_hits, static_cast<uint32_t>(__popc(hitm)));
          }
          at = __shfl_sync(FULL, at, leader);
          if (keep[u]) sm.slow[at + __popc(keepm & ((1u << lane) - 1u))] = static_cast<uint16_t>(ent[u]);
        } else if (hitm && lane == 0) {
          smem_add_noret(&sm.memo_hits, static_cast<uint32_t>(__popc(hitm)));
        }
      }
    }
    __syncthreads();
    n_slow = sm.n_slow2;
    if (tid == 0 && sm.memo_hits) atomicAdd(&P.call->memo_hits, static_cast<unsigned long long>(sm.memo_hits));
    if (P.range_index >= 1) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) my_lookups += __shfl_xor_sync(FULL, my_lookups, o);
      if (lane == 0 && my_lookups) atomicAdd(&P.call->memo_lookups, static_cast<unsigned long long>(my_lookups));
    }
    if (n_slow == 0) return;  // uniform
  }

  // ---- hand the unsettled segments to K2: 16-byte entries in the global slow
  // list (one reservation per tile), id-scratch space reserved by byte length
  // (a segment never has more ids than bytes).
  {
    const uint32_t per = (n_slow + THREADS - 1) / THREADS;
    const uint32_t lo = min(n_slow, static_cast<uint32_t>(tid) * per);
    const uint32_t hi = min(n_slow, lo + per);
    uint32_t my_len = 0, n_walk = 0;
#pragma unroll 1
    for (uint32_t i = lo; i < hi; i++) {
      const uint32_t ent = sm.slow[i];
      const uint32_t k = ent & 0xFFFu;
      if (dirty || (ent & SLOW_WALK)) {
        n_walk += (ent & SLOW_WALK) ? 1u : 0u;
        continue;
      }
      const uint32_t j = k + skip;
      const int e = j < n_ends ? static_cast<int>(sm.seg_e[j]) : limit;
      my_len += static_cast<uint32_t>(e - static_cast<int>(sm.seg_s[k] & POS_MASK));
    }
    uint32_t total_len;
    const unsigned long long sc2 = tile_exclusive_scan(sm.warp_sums, my_len);
    uint32_t run = static_cast<uint32_t>(sc2);
    total_len = static_cast<uint32_t>(sc2 >> 32);
    if (tid == 0) {
      // n_slow and tok_reserved sit side by side: one 64-bit atomic reserves both (one round trip to L2)
      static_assert(offsetof(RangeCounters, tok_reserved) == offsetof(RangeCounters, n_slow) + 4 &&
                        offsetof(RangeCounters, n_slow) % 8 == 0,
                    "n_slow / tok_reserved must form one aligned 64-bit word");
      const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(&P.counters->n_slow),
                                               (static_cast<unsigned long long>(total_len) << 32) | n_slow);
      sm.slow_base = static_cast<uint32_t>(old);
      sm.tok_base = static_cast<uint32_t>(old >> 32);
    }
    if (n_walk) atomicAdd(&P.call->long_segments, static_cast<unsigned long long>(n_walk));
    __syncthreads();
    const uint32_t slow_base = sm.slow_base;
    const uint32_t tok_base = sm.tok_base;
    if (static_cast<unsigned long long>(slow_base) + n_slow > P.slow_capacity ||
        static_cast<unsigned long long>(tok_base) + total_len > P.tok_capacity) {
      if (tid == 0) P.call->overflow = 1u;
      return;  // uniform
    }
#pragma unroll 1
    for (uint32_t i = lo; i < hi; i++) {
      const uint32_t ent = sm.slow[i];
      const uint32_t k = ent & 0xFFFu;
      const uint32_t sv = sm.seg_s[k];
      int wpos = static_cast<int>(sv & POS_MASK);
      const bool walk = dirty || (ent & SLOW_WALK);
      uint32_t len = 0;
      if (!walk) {
        const uint32_t j = k + skip;
        const int e = j < n_ends ? static_cast<int>(sm.seg_e[j]) : limit;
        len = static_cast<uint32_t>(e - wpos);
      }
      if (dirty) {
        // window position -> raw window position: undo the compaction
        int a = 0, b = NCHUNK;  // last chunk whose first surviving byte is at or before wpos
        while (a < b) {
          const int m = (a + b + 1) >> 1;
          if (sm.kept_scan[m] <= static_cast<uint32_t>(wpos)) a = m; else b = m - 1;
        }
        const uint32_t kk = static_cast<uint32_t>(wpos) - sm.kept_scan[a];
        wpos = a * CHUNK + static_cast<int>(__fns(sm.m_kept[a], 0, static_cast<int>(kk) + 1));
      }
      const size_t pos = t0 + static_cast<size_t>(wpos);
      SlowEntry out;
      out.pos_lo = static_cast<uint32_t>(pos);
      out.meta = static_cast<uint32_t>((pos >> 32) & 0xFFu) | (len << 8) | ((sv >> 14) << 24) |
                 ((ent & SLOW_FIRST_MISSED) ? SLOW_META_MISSED : 0u) | (walk ? SLOW_META_WALK : 0u);
      out.tok_off = tok_base + run;
      out.seg = static_cast<uint32_t>(seg_base + k);
      run += len;
      if (!walk && len <= SLOW_TEXT_BYTES) {
        // K2 would otherwise fetch these bytes from DRAM again, one random sector per piece
        out.meta |= SLOW_META_TEXT;
        const int a = wpos & ~3;
        const uint32_t sh = (wpos & 3) * 8;
        uint32_t x[9];
#pragma unroll
        for (int q = 0; q < 9; q++) x[q] = ld_u32(buf + a + 4 * q);
        uint32_t y[8];
#pragma unroll
        for (int q = 0; q < 8; q++) y[q] = __funnelshift_r(x[q], x[q + 1], sh);
        uint4 *dst = P.slow_text + 2 * static_cast<size_t>(slow_base + i);
        dst[0] = make_uint4(y[0], y[1], y[2], y[3]);
        dst[1] = make_uint4(y[4], y[5], y[6], y[7]);
      }
      *reinterpret_cast<uint4 *>(&P.slow[slow_base + i]) = *reinterpret_cast<const uint4 *>(&out);
      P.seg_result[seg_base + k] = SEG_RESULT_SLOW | (slow_base + i);
    }
  }
}

// ================================================================ K2: match

constexpr uint32_t SEG_HAN_FIRST = 1u;   // about to match the first piece of a Han-led segment
constexpr uint32_t SEG_KNOWN_MISS = 2u;  // the whole-window probe of the first piece is known to miss (K1 did it)

// The exact byte-wise lane for one entry: the segment left its tile's window or holds invalid UTF-8.
__device__ __noinline__ void match_walk(const EncodeParams &P, uint32_t i, size_t seg_pos, uint32_t spill_base) {
  const TextView tv{P.text, P.n_bytes};
  int32_t unk_at, tmp;
  const uint32_t cnt = walk_segment(P.vocab, tv, seg_pos, nullptr, 0, -1, &unk_at);
  const uint32_t off = spill_base + atomicAdd(&P.counters->tok_spill, cnt);
  uint32_t written = 0;
  if (static_cast<unsigned long long>(off) + cnt <= P.tok_capacity) {
    walk_segment(P.vocab, tv, seg_pos, P.tok + off, cnt, unk_at, &tmp);
    written = cnt;
  } else {
    P.call->overflow = 1u;
  }
  const uint32_t seg = P.slow[i].seg;
  *reinterpret_cast<uint4 *>(&P.slow[i]) = make_uint4(written, off, 0u, 0u);  // result form, not inline
  P.seg_result[seg] = SEG_RESULT_SLOW | (min(written, SEG_SLOW_COUNT_MAX) << SEG_SLOW_INDEX_BITS) | i;
}

// Every lane owns a long run of slow-list entries (its warp's share / 32), so
// chains of very different length average out.  The loop is aligned on PIECES:
//   round:  lanes without a segment take the next entry
//           every lane loads the 24-byte window of its next piece
//           inner loop: one table probe per iteration until every lane's piece
//                       is settled (whole-window probe, then binary search for
//                       the deepest trie node; a collision is one more turn)
//           every lane reads its longest match off the deepest node and applies
//           the piece (fast.cpp:66-91)
// so the refill / window / apply code runs once per piece and warp, fully
// converged, and only the short probe body repeats.
constexpr int LANE_TEXT_WORDS = 17;  // 68 bytes per lane: 32 + 28 readable past any piece start, odd stride (banks)

__global__ void __launch_bounds__(MATCH_THREADS, 3) wp_match_kernel(EncodeParams P) {
  __shared__ uint4 key_mask[KEY_MASK_ROW * (WP_KEY_BYTES + 1)];
  __shared__ uint32_t lane_text[MATCH_THREADS * LANE_TEXT_WORDS];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const DeviceVocab &V = P.vocab;
  init_key_mask(key_mask, tid);
  __syncthreads();

  if (P.call->overflow) return;  // a K1 tile gave up (scratch too small): its entries are unwritten, the host retries
  const uint32_t n_slow = min(P.counters->n_slow, P.slow_capacity);
  const uint32_t spill_base = min(P.counters->tok_reserved, P.tok_capacity);
  const uint32_t n_warps = gridDim.x * (MATCH_THREADS / 32);
  const uint32_t gw = blockIdx.x * (MATCH_THREADS / 32) + (tid >> 5);
  const uint32_t per = ((n_slow + n_warps - 1) / n_warps + 31u) & ~31u;
  uint32_t cursor = min(n_slow, gw * per);             // warp-uniform: next unassigned entry of this warp
  const uint32_t cursor_end = min(n_slow, cursor + per);
  const uint4 *tab = reinterpret_cast<const uint4 *>(V.slots);

  bool have = false;       // this lane holds an unfinished segment
  // K1 of this range is done, so the counters are final: every lane reads the same verdict, and one thread
  // records it for the K1 tiles of the next range (in a cache line of its own: the counters' line is hot)
  const bool worth = memo_worthwhile(P.call->memo_lookups, P.call->memo_hits, P.range_index <= 1);
  if (blockIdx.x == 0 && tid == 0 && !worth) P.call->memo_off = 1u;
  bool memo_on = P.memo != nullptr && (P.range_index < 2 || worth);
  bool in_smem = false;    // ... whose bytes sit in this lane's shared-memory buffer
  uint32_t *const my_text = lane_text + tid * LANE_TEXT_WORDS;
  size_t seg_pos = 0;
  uint32_t ent_index = 0, seg_len = 0, p = 0, tok_off = 0;
  uint32_t nid = 0, word_first = 0, kind = WP_KIND_PREFIX, flags = 0, first_len = 0;
  int32_t t0 = 0, t1 = 0, t2 = 0;  // the first three ids of the segment stay in registers (see the result form)

  for (;;) {
    // -- refill: lanes without a segment take the next entries of this warp's share
    const uint32_t needm = __ballot_sync(FULL, !have);
    if (needm && cursor < cursor_end) {
      const uint32_t i = cursor + __popc(needm & ((1u << lane) - 1u));
      cursor += __popc(needm);  // may pass cursor_end; entries beyond it are simply not taken
      if (cursor + lane < cursor_end) {
        // the next 32 entries of this warp's share: by the time a lane takes one, it sits in L1
        prefetch_l1(&P.slow[cursor + lane]);
        prefetch_l1(P.slow_text + 2 * static_cast<size_t>(cursor + lane));
      }
      if (!have && i < cursor_end) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(&P.slow[i]));
        const uint32_t meta = raw.y;
        seg_pos = static_cast<size_t>(raw.x) | (static_cast<size_t>(meta & 0xFFu) << 32);
        if (meta & SLOW_META_WALK) {
          match_walk(P, i, seg_pos, spill_base);
        } else {
          ent_index = i;
          seg_len = (meta >> 8) & 0xFFFFu;
          tok_off = raw.z;
          my_text[LANE_TEXT_WORDS - 1] = raw.w;  // the segment's number (a spare word of the lane buffer)
          in_smem = (meta & SLOW_META_TEXT) != 0;
          if (in_smem) {
            const uint4 ta = __ldg(P.slow_text + 2 * static_cast<size_t>(i));
            const uint4 tb = __ldg(P.slow_text + 2 * static_cast<size_t>(i) + 1);
            my_text[0] = ta.x; my_text[1] = ta.y; my_text[2] = ta.z; my_text[3] = ta.w;
            my_text[4] = tb.x; my_text[5] = tb.y; my_text[6] = tb.z; my_text[7] = tb.w;
            first_len = utf8_lead_len(ta.x & 0xFFu);
          } else {
            first_len = utf8_lead_len(P.text[seg_pos]);
          }
          have = true;
          p = 0;
          nid = 0;
          word_first = 0;
          kind = WP_KIND_PREFIX;
          flags = (((meta >> 24) & 3u) == CLS_HAN ? SEG_HAN_FIRST : 0u) | ((meta & SLOW_META_MISSED) ? SEG_KNOWN_MISS : 0u);
        }
      }
    }
    if (!__any_sync(FULL, have)) {
      if (cursor >= cursor_end) break;
      continue;
    }

    // -- piece start: window bytes and search bounds
    uint32_t r[6] = {0, 0, 0, 0, 0, 0};
    uint32_t k = 0, lo = 0, hi = 0, poff = 0;
    uint32_t node_w5 = 0, node_slot = 0;
    int32_t node_term = WP_NO_ID, node_best = WP_NO_ID;
    bool searching = have;
    if (have) {
      if (in_smem) {
        // p < 32, so the 28 bytes read stay inside the 68-byte lane buffer; bytes past the segment are
        // stale but never enter a key (k <= remaining length)
        const uint32_t a = p >> 2, sh = (p & 3u) * 8u;
        uint32_t x[7];
#pragma unroll
        for (int q = 0; q < 7; q++) x[q] = my_text[a + q];
#pragma unroll
        for (int q = 0; q < 6; q++) r[q] = __funnelshift_r(x[q], x[q + 1], sh);
      } else {
        load_window_global(P.text, P.n_bytes, seg_pos + p, r);
      }
      const uint32_t wlen = seg_len - p;
      const uint32_t k0 = wlen < WP_KEY_BYTES ? wlen : WP_KEY_BYTES;
      if (flags & SEG_KNOWN_MISS) {  // k0 >= 2 here
        hi = k0;
        k = k0 >> 1;
        flags &= ~SEG_KNOWN_MISS;
      } else {
        hi = k0 + 1;
        k = k0;
      }
    }

    // -- probes: deepest trie node along the window (existence is monotone in the depth)
    while (__any_sync(FULL, searching)) {
      if (searching) {
        uint32_t kw[6];
        make_key_tab(key_mask, r, k, kind, kw);
        const uint32_t idx = (key_hash(kw[0], kw[1], kw[2], kw[3], kw[4], kw[5]) + poff) & V.slot_mask;
        uint4 sa, sb, sc, sd;
        ld_slot(tab, idx, &sa, &sb);
        uint32_t oc;
        if (V.probe_pairs) {  // uniform
          ld_slot(tab, (idx + 1) & V.slot_mask, &sc, &sd);
          oc = pair_outcome(sa, sb, sc, sd, kw);
        } else {
          sc = sa;
          sd = sb;
          oc = slot_len(sb.y) == 0 ? PAIR_MISS : (slot_matches(sa, sb, kw) ? PAIR_HIT0 : PAIR_BOTH_OTHER);
        }
        if (oc == PAIR_BOTH_OTHER) {
          poff += V.probe_pairs ? 2u : 1u;  // the slot(s) hold other keys: walk on, same key
        } else {
          poff = 0;
          if (oc != PAIR_MISS) {
            const bool second = oc == PAIR_HIT1;
            lo = k;
            node_w5 = second ? sd.y : sb.y;
            node_term = static_cast<int32_t>(second ? sd.z : sb.z);
            node_best = static_cast<int32_t>(second ? sd.w : sb.w);
            node_slot = (idx + (second ? 1u : 0u)) & V.slot_mask;
          } else {
            hi = k;
          }
          k = (lo + hi) >> 1;
          searching = hi - lo > 1;
        }
      }
    }

    // -- the deepest node is at depth lo: read the longest match off it and apply the piece
    if (have) {
      uint32_t mlen = 0;
      int32_t mid = WP_NO_ID;
      if (lo != 0) {
        if (node_term != WP_NO_ID) {
          mlen = lo;
          mid = node_term;
        } else if (slot_best_len(node_w5) != 0) {
          mlen = slot_best_len(node_w5);
          mid = node_best;
        }
        if (lo == WP_KEY_BYTES && slot_has_long(node_w5) && seg_len - p > WP_KEY_BYTES) {
          // tokens longer than the inline key hang off this node, longest first
          const uint32_t ref = V.long_ref[node_slot];
          const uint32_t cnt = V.long_entries[ref];
          const uint8_t *txt = P.text + seg_pos + p;
          for (uint32_t li = 0; li < cnt; li++) {
            const uint32_t len = V.long_entries[ref + 1 + 3 * li];
            if (len > seg_len - p) continue;
            const uint8_t *tok = V.long_bytes + V.long_entries[ref + 3 + 3 * li];
            uint32_t o = WP_KEY_BYTES;
            while (o < len && txt[o] == tok[o]) o++;
            if (o == len) {
              mlen = len;
              mid = static_cast<int32_t>(V.long_entries[ref + 2 + 3 * li]);
              break;
            }
          }
        }
      }
      bool done = false;
      int32_t *out = P.tok + tok_off;
      // ids 0..2 go to registers, later ones straight to the id scratch
      auto put = [&](uint32_t index, int32_t v) {
        if (index == 0) t0 = v;
        else if (index == 1) t1 = v;
        else if (index == 2) t2 = v;
        else out[index] = v;
      };
      if (flags & SEG_HAN_FIRST) {
        flags = 0;
        nid = 1;
        if (mlen == 0) {
          t0 = V.unk_id;
          if (V.han_swallow) {
            done = true;  // fast.cpp:85-88: begin += word_len swallows the run
          } else {
            word_first = 1;
            p += first_len;
          }
        } else {
          t0 = mid;
          p += mlen;
          if (mlen == first_len) {
            word_first = 1;  // fast.cpp:89-91: the next position follows a spacing char => new word
          } else {
            kind = WP_KIND_SUFFIX;
          }
        }
      } else if (mlen == 0) {  // fast.cpp:79-88: whole-word UNK, earlier pieces rolled back
        put(word_first, V.unk_id);
        nid = word_first + 1;
        done = true;
      } else {
        put(nid, mid);
        nid++;
        p += mlen;
        kind = WP_KIND_SUFFIX;
      }
      if (done || p >= seg_len) {
        // result form of the entry (read by K3): up to three ids inline — one 16-byte store and nothing
        // else for most segments — else the count and where the ids sit in the id scratch
        uint4 res;
        if (nid <= 3) {
          res = make_uint4(nid | SLOW_RESULT_INLINE, static_cast<uint32_t>(t0), static_cast<uint32_t>(t1),
                           static_cast<uint32_t>(t2));
        } else {
          out[0] = t0;
          out[1] = t1;
          out[2] = t2;
          res = make_uint4(nid, tok_off, 0u, 0u);
        }
        *reinterpret_cast<uint4 *>(&P.slow[ent_index]) = res;
        P.seg_result[my_text[LANE_TEXT_WORDS - 1]] =
            SEG_RESULT_SLOW | (min(nid, SEG_SLOW_COUNT_MAX) << SEG_SLOW_INDEX_BITS) | ent_index;
        have = false;
        if (memo_on && in_smem && seg_len <= MEMO_KEY_BYTES && nid <= 3) {
          // record bytes -> ids in the word memo so that K1 settles every later occurrence itself
          const uint4 ma = key_mask[KEY_MASK_ROW * seg_len];
          const uint32_t k0 = my_text[0] & ma.x, k1 = my_text[1] & ma.y, k2 = my_text[2] & ma.z, k3 = my_text[3] & ma.w;
          uint32_t idx = key_hash(k0, k1, k2, k3, seg_len, MEMO_SALT) & P.memo_mask;
          bool placed = false;
          for (int t = 0; t < 4 && !placed; t++) {
            uint4 *slot = P.memo + 2 * static_cast<size_t>(idx);
            unsigned int *state = reinterpret_cast<unsigned int *>(slot + 1);
            const unsigned int old = atomicCAS(state, 0u, 1u);
            if (old == 0u) {
              // claimed.  No fence between the payload and READY: the readers that need a complete slot (K1
              // and K3) run in later kernels; a concurrent K2 lane that sees READY early can at worst fail to
              // recognise its own word here and store a harmless duplicate one slot further.
              slot[0] = make_uint4(k0, k1, k2, k3);
              state[1] = static_cast<unsigned int>(t0);
              state[2] = static_cast<unsigned int>(t1);
              state[3] = static_cast<unsigned int>(t2);
              *reinterpret_cast<volatile unsigned int *>(state) = MEMO_READY | (nid << 8) | seg_len;
              placed = true;
            } else if (old == 1u) {
              placed = true;  // another lane is writing this slot right now (most likely the same word)
            } else if ((old & 0xFFu) == seg_len) {
              const uint4 a = __ldcg(slot);
              if (a.x == k0 && a.y == k1 && a.z == k2 && a.w == k3) placed = true;  // already there
            }
            idx = (idx + 1) & P.memo_mask;
          }
          if (!placed) memo_on = false;  // crowded neighbourhood: this lane stops feeding the memo
        }
      }
    }
  }
}

// ============================================================== K3: scatter

struct ScatterSmem {
  int32_t stage[SCATTER_STAGE];
  uint32_t desc_pos[SCATTER_SEGS];     // segments not settled by K1: stage position << 16 | id count
  uint32_t desc_src[SCATTER_SEGS];     // ... and their seg_result word (where the ids are)
  uint32_t n_desc;
  uint32_t warp_sums[SCATTER_THREADS / 32];
  uint32_t block_index[2];
  unsigned long long base;
};

// ids of one segment that K1 did not settle -> dst[0..cnt): from the memo slot, inline in the slow entry, or
// from the id scratch
__device__ __forceinline__ void scatter_fetch(const EncodeParams &P, uint32_t res, uint32_t cnt, int32_t *dst) {
  if (res & SEG_RESULT_SLOW) {
    const uint32_t si = res & SEG_SLOW_INDEX_MASK;
    if (si >= P.slow_capacity) return;
    const uint4 e = *reinterpret_cast<const uint4 *>(&P.slow[si]);
    if (e.x & SLOW_RESULT_INLINE) {
      dst[0] = static_cast<int32_t>(e.y);
      if (cnt > 1) dst[1] = static_cast<int32_t>(e.z);
      if (cnt > 2) dst[2] = static_cast<int32_t>(e.w);
    } else if (static_cast<unsigned long long>(e.y) + cnt <= P.tok_capacity) {
      const int32_t *src = P.tok + e.y;
      for (uint32_t t = 0; t < cnt; t++) dst[t] = src[t];
    }
  } else {
    const uint4 e = *reinterpret_cast<const uint4 *>(P.memo + 2 * static_cast<size_t>(res & SEG_MEMO_SLOT_MASK) + 1);
    dst[0] = static_cast<int32_t>(e.y);
    if (cnt > 1) dst[1] = static_cast<int32_t>(e.z);
    if (cnt > 2) dst[2] = static_cast<int32_t>(e.w);
  }
}

__global__ void __launch_bounds__(SCATTER_THREADS) wp_scatter_kernel(EncodeParams P) {
  __shared__ ScatterSmem sm;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const unsigned long long n_segs = min(P.counters->n_segs, static_cast<unsigned long long>(P.seg_capacity));
  const uint32_t n_blocks = static_cast<uint32_t>((n_segs + SCATTER_SEGS - 1) / SCATTER_SEGS);
  const unsigned long long ids_in = P.call->ids_total[P.range_parity];
  if (P.call->overflow) return;  // uniform over the grid (K1 and K2 are done): the host retries the call
  if (n_blocks == 0) {
    if (blockIdx.x == 0 && tid == 0) P.call->ids_total[P.range_parity ^ 1u] = ids_in;
    return;
  }
  // The ticket of the NEXT block is taken at the start of an iteration, so that its round trip is hidden
  // behind the work on this one (two slots, used alternately).  A CTA that holds a ticket while it still
  // works on an earlier block cannot stall the chain: its current block only waits for smaller indices.
  if (tid == 0) sm.block_index[0] = atomicAdd(&P.counters->scatter_ticket, 1u);
  for (uint32_t it = 0;; it++) {
    __syncthreads();
    const uint32_t b = sm.block_index[it & 1u];
    if (b >= n_blocks) break;
    if (tid == 0) {
      sm.block_index[(it + 1u) & 1u] = atomicAdd(&P.counters->scatter_ticket, 1u);
      sm.n_desc = 0;
    }
    const unsigned long long first = static_cast<unsigned long long>(b) * SCATTER_SEGS + tid * SCATTER_ITEMS;
    {
      // blocks are handed out in order to gridDim.x CTAs: pull the words of the block two rounds ahead into L2
      const unsigned long long pf = (static_cast<unsigned long long>(b) + 2ull * gridDim.x) * SCATTER_SEGS + tid * 32ull;
      if (tid < SCATTER_SEGS / 32 && pf < n_segs) prefetch_l2(P.seg_result + pf);
    }

    // per-segment id counts, straight from the seg_result words (K1: 1; memo and slow: count bits)
    uint32_t res[SCATTER_ITEMS], cnt[SCATTER_ITEMS];
    if (first + SCATTER_ITEMS <= n_segs) {
      const uint4 a = *reinterpret_cast<const uint4 *>(P.seg_result + first);
      const uint4 c = *reinterpret_cast<const uint4 *>(P.seg_result + first + 4);
      res[0] = a.x; res[1] = a.y; res[2] = a.z; res[3] = a.w;
      res[4] = c.x; res[5] = c.y; res[6] = c.z; res[7] = c.w;
    } else {
#pragma unroll
      for (int j = 0; j < SCATTER_ITEMS; j++) res[j] = first + j < n_segs ? P.seg_result[first + j] : 0u;
    }
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < SCATTER_ITEMS; j++) {
      const uint32_t slow_cnt = (res[j] >> SEG_SLOW_INDEX_BITS) & SEG_SLOW_COUNT_MAX;
      const uint32_t memo_cnt = (res[j] >> SEG_MEMO_SLOT_BITS) & 3u;
      uint32_t c = (res[j] & SEG_RESULT_SLOW) ? slow_cnt : ((res[j] & SEG_RESULT_MEMO) ? memo_cnt : 1u);
      if (first + j >= n_segs) c = 0;
      if ((res[j] & SEG_RESULT_SLOW) && slow_cnt == SEG_SLOW_COUNT_MAX && c != 0) {  // rare: 31 ids or more
        const uint32_t si = res[j] & SEG_SLOW_INDEX_MASK;
        c = si < P.slow_capacity ? (P.slow[si].pos_lo & ~SLOW_RESULT_INLINE) : 0u;  // word 0 of the result form
      }
      cnt[j] = c;
      mine += c;
    }
    uint32_t total;
    uint32_t at = block_exclusive_scan<SCATTER_THREADS / 32>(sm.warp_sums, mine, &total);

    if (tid == 0) lookback_publish(P.block_state, b, total);

    if (total <= SCATTER_STAGE) {
      // stage in shared memory, then write out coalesced.  Ids settled by K1 are placed at once; the other
      // segments are listed and fetched one per thread, all lanes busy (inline they would leave most
      // lanes of a warp idle behind the few that have one).
#pragma unroll
      for (int j = 0; j < SCATTER_ITEMS; j++) {
        const bool other = (res[j] & (SEG_RESULT_SLOW | SEG_RESULT_MEMO)) != 0 && cnt[j] != 0;
        if (cnt[j] != 0 && !other) sm.stage[at] = static_cast<int32_t>(res[j]) - 1;
        const uint32_t om = __ballot_sync(FULL, other);
        if (om) {
          uint32_t d = 0;
          const int leader = __ffs(om) - 1;
          if (lane == leader) d = smem_add(&sm.n_desc, static_cast<uint32_t>(__popc(om)));
          d = __shfl_sync(FULL, d, leader) + __popc(om & ((1u << lane) - 1u));
          if (other) {
            sm.desc_pos[d] = (at << 16) | cnt[j];
            sm.desc_src[d] = res[j];
          }
        }
        at += cnt[j];
      }
      __syncthreads();
      const uint32_t n_desc = sm.n_desc;
      for (uint32_t d = tid; d < n_desc; d += SCATTER_THREADS) {
        const uint32_t dp = sm.desc_pos[d];
        scatter_fetch(P, sm.desc_src[d], dp & 0xFFFFu, sm.stage + (dp >> 16));
      }
      // the ids are staged; only now the block needs its place in the output
      if (warp == 0) {
        const unsigned long long base = lookback_walk(P.block_state, b, total, lane);
        if (lane == 0) {
          sm.base = base;
          if (b == n_blocks - 1) P.call->ids_total[P.range_parity ^ 1u] = ids_in + base + total;
        }
      }
      __syncthreads();
      const unsigned long long out0 = ids_in + sm.base;
      for (uint32_t i = tid; i < total; i += SCATTER_THREADS) {
        if (out0 + i < P.capacity) P.ids[out0 + i] = sm.stage[i];
      }
    } else {
      // a block with unusually many ids (long words cut into many pieces): write directly
      if (warp == 0) {
        const unsigned long long base = lookback_walk(P.block_state, b, total, lane);
        if (lane == 0) {
          sm.base = base;
          if (b == n_blocks - 1) P.call->ids_total[P.range_parity ^ 1u] = ids_in + base + total;
        }
      }
      __syncthreads();
      const unsigned long long out0 = ids_in + sm.base;
#pragma unroll
      for (int j = 0; j < SCATTER_ITEMS; j++) {
        if (cnt[j] == 0) continue;
        const unsigned long long o = out0 + at;
        if (!(res[j] & (SEG_RESULT_SLOW | SEG_RESULT_MEMO))) {
          if (o < P.capacity) P.ids[o] = static_cast<int32_t>(res[j]) - 1;
        } else if (o + cnt[j] <= P.capacity) {
          scatter_fetch(P, res[j], cnt[j], P.ids + o);
        } else {
          // the caller's buffer ends inside this segment (the call reports WP_ERR_CAPACITY): id by id
          int32_t tmp[3];
          if (cnt[j] <= 3) {
            scatter_fetch(P, res[j], cnt[j], tmp);
            for (uint32_t t = 0; t < cnt[j]; t++) {
              if (o + t < P.capacity) P.ids[o + t] = tmp[t];
            }
          } else {
            const uint32_t si = res[j] & SEG_SLOW_INDEX_MASK;
            if (si < P.slow_capacity) {
              const uint4 e = *reinterpret_cast<const uint4 *>(&P.slow[si]);
              if (static_cast<unsigned long long>(e.y) + cnt[j] <= P.tok_capacity) {
                for (uint32_t t = 0; t < cnt[j]; t++) {
                  if (o + t < P.capacity) P.ids[o + t] = P.tok[e.y + t];
                }
              }
            }
          }
        }
        at += cnt[j];
      }
    }
  }
}

// ============================================================== K4: format
//
// ids -> the reference's output wire format, every id in decimal followed by one space (fast.cpp:214-216,
// utils.cpp:30-35: `"id id id "`), for the streaming entry point (encodeExternal).  Two kernels: the total
// length, then the text (per-block scan of the lengths, decoupled look-back, chars staged in shared memory).

constexpr int FORMAT_THREADS = 256;
constexpr int FORMAT_ITEMS = 8;
constexpr int FORMAT_IDS = FORMAT_THREADS * FORMAT_ITEMS;  // 2048 ids per block iteration
constexpr int FORMAT_MAX_CHARS = 12;                       // "-2147483648 "

__device__ __forceinline__ uint32_t decimal_chars(int32_t v) {  // digits + sign + the trailing space
  const uint32_t a = v < 0 ? 0u - static_cast<uint32_t>(v) : static_cast<uint32_t>(v);
  uint32_t n = 2u + (v < 0 ? 1u : 0u);
  n += a >= 10u;
  n += a >= 100u;
  n += a >= 1000u;
  n += a >= 10000u;
  n += a >= 100000u;
  n += a >= 1000000u;
  n += a >= 10000000u;
  n += a >= 100000000u;
  n += a >= 1000000000u;
  return n;
}

__global__ void __launch_bounds__(FORMAT_THREADS) wp_format_total_kernel(const int32_t *__restrict__ ids,
                                                                         unsigned long long n,
                                                                         unsigned long long *total) {
  __shared__ unsigned long long warp_sums[FORMAT_THREADS / 32];
  unsigned long long mine = 0;
  for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * FORMAT_THREADS + threadIdx.x; i < n;
       i += static_cast<unsigned long long>(gridDim.x) * FORMAT_THREADS)
    mine += decimal_chars(ids[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(FULL, mine, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int w = 0; w < FORMAT_THREADS / 32; w++) s += warp_sums[w];
    if (s) atomicAdd(total, s);
  }
}

struct FormatSmem {
  uint8_t stage[FORMAT_IDS * FORMAT_MAX_CHARS];
  uint32_t warp_sums[FORMAT_THREADS / 32];
  uint32_t block_index;
  unsigned long long base;
};

__global__ void __launch_bounds__(FORMAT_THREADS) wp_format_kernel(const int32_t *__restrict__ ids, unsigned long long n,
                                                                   char *__restrict__ out, unsigned long long *block_state,
                                                                   unsigned int *ticket) {
  __shared__ FormatSmem sm;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const uint32_t n_blocks = static_cast<uint32_t>((n + FORMAT_IDS - 1) / FORMAT_IDS);
  for (;;) {
    __syncthreads();
    if (tid == 0) sm.block_index = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t b = sm.block_index;
    if (b >= n_blocks) break;
    const unsigned long long first = static_cast<unsigned long long>(b) * FORMAT_IDS + tid * FORMAT_ITEMS;
    int32_t v[FORMAT_ITEMS];
    uint32_t len[FORMAT_ITEMS], mine = 0;
#pragma unroll
    for (int j = 0; j < FORMAT_ITEMS; j++) {
      v[j] = first + j < n ? ids[first + j] : 0;
      len[j] = first + j < n ? decimal_chars(v[j]) : 0u;
      mine += len[j];
    }
    uint32_t total;
    uint32_t at = block_exclusive_scan<FORMAT_THREADS / 32>(sm.warp_sums, mine, &total);
    if (tid == 0) lookback_publish(block_state, b, total);
#pragma unroll
    for (int j = 0; j < FORMAT_ITEMS; j++) {
      if (len[j] == 0) continue;
      uint32_t a = v[j] < 0 ? 0u - static_cast<uint32_t>(v[j]) : static_cast<uint32_t>(v[j]);
      uint32_t p = at + len[j] - 1;
      sm.stage[p--] = ' ';
      do {
        sm.stage[p--] = static_cast<uint8_t>('0' + a % 10u);
        a /= 10u;
      } while (a != 0u);
      if (v[j] < 0) sm.stage[p] = '-';
      at += len[j];
    }
    __syncthreads();
    if (tid < 32) {
      const unsigned long long base = lookback_walk(block_state, b, total, lane);
      if (lane == 0) sm.base = base;
    }
    __syncthreads();
    char *dst = out + sm.base;
    for (uint32_t i = tid; i < total; i += FORMAT_THREADS) dst[i] = static_cast<char>(sm.stage[i]);
  }
}

cudaError_t launch_format_total(const int32_t *ids, size_t n, unsigned long long *total, int sm_count,
                                cudaStream_t stream, uint64_t *launches) {
  if (sm_count <= 0) sm_count = 148;
  wp_format_total_kernel<<<sm_count * 8, FORMAT_THREADS, 0, stream>>>(ids, n, total);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_format(const int32_t *ids, size_t n, char *out, unsigned long long *block_state, unsigned int *ticket,
                          int sm_count, cudaStream_t stream, uint64_t *launches) {
  if (sm_count <= 0) sm_count = 148;
  wp_format_kernel<<<sm_count * 4, FORMAT_THREADS, 0, stream>>>(ids, n, out, block_state, ticket);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

uint32_t format_block_ids() { return FORMAT_IDS; }

// -------------------------------------------------------------------- launch

uint32_t encode_tile_bytes() { return TILE; }
uint32_t scatter_block_segments() { return SCATTER_SEGS; }

cudaError_t launch_encode_range(const EncodeParams &P, int sm_count, cudaStream_t stream, uint64_t *launches,
                                cudaEvent_t *timing) {
  // the opt-in to > 48 KB of dynamic shared memory is per device (K1 stays below it, but keep it explicit)
  static bool configured[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    e = cudaFuncSetAttribute(wp_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(sizeof(TileSmem)));
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  if (sm_count <= 0) sm_count = 148;
  // K1 and K2 probe the vocabulary table at random: ask L2 to keep it resident while text, ids and the
  // intermediates stream through (access policy window = the slot array, everything else streaming).
  cudaLaunchAttribute attr[1];
  unsigned n_attr = 0;
  if (P.persist_bytes > 0) {
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow.base_ptr = const_cast<Slot *>(P.vocab.slots);
    attr[0].val.accessPolicyWindow.num_bytes = P.persist_bytes;
    attr[0].val.accessPolicyWindow.hitRatio = P.persist_ratio;
    attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    n_attr = 1;
  }
  cudaLaunchConfig_t cfg{};
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;

  if (timing) cudaEventRecord(timing[0], stream);
  cfg.gridDim = dim3(P.n_tiles);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = sizeof(TileSmem);
  e = cudaLaunchKernelEx(&cfg, wp_split_kernel, P);
  if (e != cudaSuccess) return e;

  if (timing) cudaEventRecord(timing[1], stream);
  cfg.gridDim = dim3(sm_count * 3);  // = resident capacity (__launch_bounds__(256, 3)): one wave, large shares
  cfg.blockDim = dim3(MATCH_THREADS);
  cfg.dynamicSmemBytes = 0;
  e = cudaLaunchKernelEx(&cfg, wp_match_kernel, P);
  if (e != cudaSuccess) return e;

  if (timing) cudaEventRecord(timing[2], stream);
  cfg.gridDim = dim3(sm_count * 4);
  cfg.blockDim = dim3(SCATTER_THREADS);
  cfg.numAttrs = 0;
  e = cudaLaunchKernelEx(&cfg, wp_scatter_kernel, P);
  if (e != cudaSuccess) return e;
  if (timing) cudaEventRecord(timing[3], stream);
  if (launches) *launches += 3;
  return cudaSuccess;
}

}  // namespace wp
