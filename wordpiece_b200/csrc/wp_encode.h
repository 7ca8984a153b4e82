// Launch interface of the three encode kernels (wp_encode.cu).
//
//   K1 wp_split_kernel   word split + whole-window probe, one text tile per CTA
//   K2 wp_match_kernel   greedy longest-match chains of the segments K1 could
//                        not settle with one probe, load-balanced over the GPU
//   K3 wp_scatter_kernel scan of per-segment id counts + scatter of the ids
//
// The three run back to back on one stream over one RANGE of tiles of a text
// (a whole text, or a block of it when the host bounds the scratch memory);
// nothing is synchronised in between — K2 and K3 read their work sizes from
// device memory.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "wp_table.h"

namespace wp {

// One segment that needs more than the whole-window probe (16 bytes).
// K1 writes the WORK form below; K2 overwrites the entry with the RESULT form: word 0 = id count
// (| SLOW_RESULT_INLINE if at most three ids, which then sit in words 1..3), else word 1 = id-scratch offset.
struct SlowEntry {
  uint32_t pos_lo;   // text position of the segment start, low 32 bits
  uint32_t meta;     // bits 0..7 pos high bits, 8..23 byte length (0 for WALK), 24..25 char class,
                     // bit 26 whole-window probe known to miss, bit 27 WALK (walk from global memory, see K2)
  uint32_t tok_off;  // where this segment's ids go in the id scratch (K1 for plain entries, K2 for WALK)
  uint32_t seg;      // the segment's number within the range: K2 puts the id count into seg_result[seg]
};
static_assert(sizeof(SlowEntry) == 16, "slow entries are read as one 16-byte load");

constexpr uint32_t SLOW_META_MISSED = 1u << 26;
constexpr uint32_t SLOW_META_WALK = 1u << 27;
constexpr uint32_t SLOW_META_TEXT = 1u << 28;    // the segment's bytes (<= 32) were copied to slow_text[] by K1
constexpr uint32_t SLOW_TEXT_BYTES = 32;
constexpr uint32_t SLOW_RESULT_INLINE = 0x80000000u;  // result form of an entry (after K2): word 0 = id count | this
// seg_result, one word per segment: settled by K1 = id + 1 (< 2^30); slow = SEG_RESULT_SLOW | id count << 26
// (31 = 31 or more: the count is in the slow entry; filled in by K2) | slow index; memo = SEG_RESULT_MEMO |
// id count << 20 | memo slot.  K3 so knows every segment's id count without a dependent load.
constexpr uint32_t SEG_RESULT_SLOW = 0x80000000u;
constexpr uint32_t SEG_RESULT_MEMO = 0x40000000u;
constexpr uint32_t SEG_SLOW_INDEX_BITS = 26;
constexpr uint32_t SEG_SLOW_INDEX_MASK = (1u << SEG_SLOW_INDEX_BITS) - 1u;
constexpr uint32_t SEG_SLOW_COUNT_MAX = 31;
constexpr uint32_t SEG_MEMO_SLOT_BITS = 20;
constexpr uint32_t SEG_MEMO_SLOT_MASK = (1u << SEG_MEMO_SLOT_BITS) - 1u;

// Word memo (per encode call): exact bytes of a short segment -> its ids.  Text repeats its rare words; the
// first occurrence of a word that needs more than one probe is matched by K2, which records the result, and
// later tiles settle every further occurrence in K1 with one lookup.  Slot = 2 x uint4: the 16 key bytes
// (zero padded), then {state, id0, id1, id2}; state 0 = empty, 1 = being written, else MEMO_READY | count << 8
// | byte length.  Cleared at the start of every call, so results never depend on earlier calls.
constexpr uint32_t MEMO_READY = 0x80000000u;
constexpr uint32_t MEMO_KEY_BYTES = 16;
constexpr uint32_t MEMO_SALT = 0x5BD1E995u;

// The memo pays only if words repeat (natural-language text); on text whose unsettled words never repeat
// (random strings, long CJK runs) it is switched off for the rest of the call once enough lookups of the
// ranges >= 1 have shown that it settles less than 1/3 of ALL unsettled segments (a lookup per short
// unsettled word, an atomic insert per miss and a dependent read in K3 per hit cost about that much).
// Judged once per range, by K2, for the ranges after it.  The first verdict (after range 1) sees a memo
// that was warmed by 2 MiB only — English and Russian text hit 37 % there and 65-75 % later, Japanese 5 %,
// random strings 1 % — so it only asks for 1/6.  A heuristic on speed only: ids never depend on it.
__host__ __device__ inline bool memo_worthwhile(unsigned long long lookups, unsigned long long hits, bool early) {
  return !(lookups > 20000ull && hits * (early ? 6ull : 3ull) < lookups);
}

// Counters in device memory, zeroed before every range.
struct RangeCounters {
  unsigned int split_ticket;         // tile dispenser of K1
  unsigned int scatter_ticket;       // block dispenser of K3
  unsigned int n_slow;               // slow entries appended by K1
  unsigned int tok_reserved;         // id scratch reserved by K1 (plain entries)
  unsigned int tok_spill;            // id scratch reserved by K2 past the K1 part (WALK entries)
  unsigned int pad;
  unsigned long long n_segs;         // segments of the range (K1, last tile)
};

// Counters in device memory, zeroed once per encode call.
struct CallCounters {
  unsigned long long ids_total[2];   // running id count; range r reads [r & 1] and writes [(r + 1) & 1]
  unsigned long long dirty_tiles;
  unsigned long long long_segments;
  unsigned int overflow;             // set if a scratch capacity was exceeded (the host retries with more)
  unsigned int pad;
  unsigned long long memo_hits;      // segments settled by the word memo in K1
  unsigned long long memo_lookups;   // unsettled segments seen by K1's memo phase in ranges >= 1 (range 0 cannot hit)
  unsigned int pad2[18];             // (the counters above are hit by atomics from every tile)
  unsigned int memo_off;             // set by K2 once the memo has shown not to pay (memo_worthwhile); read by K1
  unsigned int pad3[31];
};
static_assert(offsetof(CallCounters, memo_off) % 128 == 0, "memo_off sits in a cache line of its own");

struct EncodeParams {
  DeviceVocab vocab;
  const uint8_t *text;              // device, the whole text (tiles read their halo from it)
  size_t n_bytes;                   // size of the whole text
  uint32_t first_tile;              // range = tiles [first_tile, first_tile + n_tiles)
  uint32_t n_tiles;
  // outputs
  int32_t *ids;                     // device, capacity entries
  unsigned long long capacity;
  CallCounters *call;               // K3 appends at call->ids_total[range_parity]
  uint32_t range_parity;
  // scratch
  RangeCounters *counters;
  unsigned long long *tile_state;   // n_tiles look-back words (K1), zeroed
  unsigned long long *block_state;  // look-back words of K3, zeroed
  uint32_t *seg_result;             // seg_capacity entries
  uint32_t seg_capacity;
  SlowEntry *slow;                  // slow_capacity entries
  uint4 *slow_text;                 // 2 x uint4 per entry: the first 32 bytes of the segment (SLOW_META_TEXT)
  uint32_t slow_capacity;
  int32_t *tok;                     // tok_capacity ids
  uint32_t tok_capacity;
  uint32_t n_scatter_blocks;        // size of block_state
  uint4 *memo;                      // word memo, memo_mask + 1 slots of 2 x uint4 (nullptr = off)
  uint32_t memo_mask;
  uint32_t range_index;             // 0, 1, 2, ... within the call
  // L2 residency hint for the vocabulary table (0 bytes = none)
  size_t persist_bytes;
  float persist_ratio;
};

uint32_t encode_tile_bytes();
uint32_t scatter_block_segments();
// Enqueue K1, K2, K3 for one range.  *launches is incremented per kernel launched.
// `timing` (optional): 4 events recorded around the three launches (before K1, K1|K2, K2|K3, after K3).
cudaError_t launch_encode_range(const EncodeParams &P, int sm_count, cudaStream_t stream, uint64_t *launches,
                                cudaEvent_t *timing = nullptr);

// ids -> "id id id " (decimal, one space after every id).  launch_format_total adds the text length to *total
// (zeroed by the caller); launch_format writes it to `out` (block_state: one zeroed word per
// format_block_ids() ids, ticket: one zeroed word).
uint32_t format_block_ids();
cudaError_t launch_format_total(const int32_t *ids, size_t n, unsigned long long *total, int sm_count, cudaStream_t stream,
                                uint64_t *launches);
cudaError_t launch_format(const int32_t *ids, size_t n, char *out, unsigned long long *block_state, unsigned int *ticket,
                          int sm_count, cudaStream_t stream, uint64_t *launches);

}  // namespace wp
