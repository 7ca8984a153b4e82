/* wordpiece_b200 — C ABI of the B200-native fast WordPiece encoder.
 *
 * This is the drop-in boundary for ONE path of gleb-kov/wordpiece: the Fast
 * greedy longest-match encoder (reference src/fast.cpp).  The reference has no
 * FFI of its own — its public surface is seven C++ free functions in
 * src/word_piece.hpp:10-36 — so the boundary is two-layered:
 *
 *   include/word_piece.hpp   the reference's own C++ signatures (namespace
 *                            word_piece::fast), implemented by a host shim that
 *   include/wordpiece_b200.h calls THIS C ABI, which launches the sm_100a kernels.
 *
 * Each entry point below names the reference interface it replaces.  Plain
 * pointers and sizes only; no C++ or torch types; nothing throws across the
 * boundary (status codes + wp_last_error()).  There is NO CPU fallback: every
 * encode call runs the CUDA kernels or fails with WP_ERR_CUDA / WP_ERR_NO_DEVICE.
 *
 * Threading: a wp_vocab handle owns one CUDA stream and scratch buffers on one
 * device; use a handle from one host thread at a time (the reference is not
 * re-entrant either: one process-global pool, utils.cpp:25-28).  Several handles
 * (e.g. one per GPU) may be used concurrently.
 */
#ifndef WORDPIECE_B200_H_
#define WORDPIECE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wp_vocab wp_vocab;

typedef enum wp_status {
  WP_OK = 0,
  WP_ERR_INVALID_ARG = 1,
  WP_ERR_EMPTY_VOCAB_WORD = 2, /* reference: throw runtime_error("Vocab word is empty"), utils.cpp:99-101 */
  WP_ERR_CUDA = 3,
  WP_ERR_NO_DEVICE = 4,
  WP_ERR_CAPACITY = 5, /* caller's id buffer too small; *n_ids still holds the exact count */
  WP_ERR_IO = 6,
  WP_ERR_NOMEM = 7,
  WP_ERR_ID_RANGE = 8 /* wp_decode: id == vocab size (reference: .at() throws out_of_range, fast.cpp:175) */
} wp_status;

/* Counters of the last encode call on a handle (diagnostics; tests use them to
 * prove that the slow lanes were exercised). */
typedef struct wp_stats {
  uint64_t n_bytes;        /* text bytes encoded */
  uint64_t n_ids;          /* ids produced */
  uint64_t n_tiles;        /* text tiles processed */
  uint64_t dirty_tiles;    /* tiles that held invalid UTF-8 (bytes dropped, utf8.cpp:130-147) */
  uint64_t long_segments;  /* segments longer than a tile's window (walked from global memory) */
  uint64_t kernel_launches;/* kernels launched by this call */
  uint64_t memo_hits;      /* segments settled by the per-call word memo (repeats of a word matched earlier) */
} wp_stats;

/* Message for the last non-OK status returned on the calling thread. */
const char *wp_last_error(void);

/* Number of CUDA kernels this library has launched in this process. */
uint64_t wp_kernel_launch_count(void);

/* Text bytes one thread block owns (the tile size of the encode kernel). */
uint32_t wp_tile_bytes(void);

/* ---- vocabulary ---------------------------------------------------------
 * Replaces utils::parseVocab (utils.cpp:108-121) + the WordPieceToken
 * constructor (utils.cpp:81-106) + the two-map build (fast.cpp:21-36), done
 * ONCE per vocabulary instead of once per encode call: id = index; "[UNK]"
 * (last such line) gives the UNK id, else -1; "##" prefix => continuation
 * token; "[...]" tokens and all-punctuation tokens of length > 1 are never
 * matched; duplicate tokens: the last index wins.  The table is uploaded to
 * `device` (CUDA ordinal).  device == -1 builds a HOST-ONLY handle (vocabulary
 * queries and wp_decode work; every encode call returns WP_ERR_NO_DEVICE). */
wp_status wp_vocab_create(const char *const *tokens, const size_t *token_lens, size_t n_tokens, int device,
                          wp_vocab **out);

/* Replaces utils::readVocabFromFile (utils.cpp:123-137): one token per line
 * ('\n' separated, a trailing '\r' stays in the token, an empty line is
 * WP_ERR_EMPTY_VOCAB_WORD). */
wp_status wp_vocab_create_from_file(const char *vocab_file, int device, wp_vocab **out);

void wp_vocab_destroy(wp_vocab *v);

size_t wp_vocab_size(const wp_vocab *v);        /* number of lines/tokens */
int32_t wp_vocab_unk_id(const wp_vocab *v);     /* WordPieceVocabulary::unk_token_id, utils.hpp:33 */
size_t wp_vocab_max_len(const wp_vocab *v);     /* max code points over matchable tokens, fast.cpp:31 */
int wp_vocab_device(const wp_vocab *v);
/* bit0 is_prefix, bit1 is_special, bit2 is_malformed (utils.hpp:23-25); bit3: token held invalid UTF-8 */
int wp_vocab_token_flags(const wp_vocab *v, size_t index);
/* Size in bytes of the device-resident table (slots + long-token lists). */
size_t wp_vocab_device_bytes(const wp_vocab *v);

/* ---- encode -------------------------------------------------------------
 * All three replace encodeFastWordPiece (fast.cpp:143-150), i.e. parseText
 * (utils.cpp:37-79) + encodeFastWordPieceImpl (fast.cpp:19-141): UTF-8 text in,
 * int32 token ids out, bit-identical to the reference.  n_bytes == 0 => 0 ids. */

/* Host text -> host ids.  *ids_out is malloc'd by the library (wp_free). */
wp_status wp_encode(wp_vocab *v, const char *text, size_t n_bytes, int32_t **ids_out, size_t *n_ids);

/* Host text -> the ids as decimal text, every id followed by one space ("id id id "): the wire format that
 * encodeExternal appends to its output file (fast.cpp:214-216) and utils::writeToFile writes
 * (utils.cpp:30-35).  Formatted on the device.  *out is malloc'd by the library (wp_free), NUL-terminated,
 * *out_len excludes the NUL. */
wp_status wp_encode_text(wp_vocab *v, const char *text, size_t n_bytes, char **out, size_t *out_len, size_t *n_ids);

/* Host text -> caller's host buffer of `capacity` ids. */
wp_status wp_encode_into(wp_vocab *v, const char *text, size_t n_bytes, int32_t *ids, size_t capacity,
                         size_t *n_ids);

/* Device text -> device ids, both resident on the handle's device.  Enqueues
 * the kernels on `stream` (a cudaStream_t; NULL = the CUDA default stream),
 * then synchronises that stream to return the count.  At most `capacity` ids
 * are written; if more were produced the status is WP_ERR_CAPACITY and *n_ids
 * is the count needed (n_bytes ids always suffice). */
wp_status wp_encode_device(wp_vocab *v, const void *d_text, size_t n_bytes, int32_t *d_ids, size_t capacity,
                           size_t *n_ids, void *stream);

/* As above but fully asynchronous: the id count is written to the device
 * word *d_n_ids (uint64) and nothing is synchronised. */
wp_status wp_encode_device_async(wp_vocab *v, const void *d_text, size_t n_bytes, int32_t *d_ids, size_t capacity,
                                 uint64_t *d_n_ids, void *stream);

/* Counters of the last completed wp_encode / wp_encode_into / wp_encode_device call. */
wp_status wp_last_stats(wp_vocab *v, wp_stats *out);

/* Per-kernel device time (diagnostics for bench.py's roofline): when enabled, every following encode call
 * on the handle records CUDA events around its three kernels — K1 split + whole-window probe, K2 match,
 * K3 scan + scatter — on the stream the kernels are launched on.  wp_last_kernel_ms waits for the last
 * call and returns the summed milliseconds of each kernel over the call's ranges. */
wp_status wp_set_kernel_timing(wp_vocab *v, int enabled);
wp_status wp_last_kernel_ms(wp_vocab *v, float ms[3], uint32_t *n_ranges);

/* ---- decode -------------------------------------------------------------
 * Replaces word_piece::fast::decode (fast.cpp:165-187): id -> token text with
 * "##" re-added for continuation tokens.  The texts are concatenated into one
 * malloc'd buffer *out; token i is (*out)[(*offsets_out)[i] .. (*offsets_out)[i+1])
 * with *n_tokens + 1 offsets (both buffers: wp_free).  ids < 0 or > size and
 * malformed tokens are skipped (the reference prints a warning and skips them);
 * id == size is WP_ERR_ID_RANGE.  *n_skipped (optional) counts skipped ids. */
wp_status wp_decode(const wp_vocab *v, const int32_t *ids, size_t n_ids, char **out, size_t **offsets_out,
                    size_t *n_tokens, size_t *n_skipped);

void wp_free(void *p);

/* ---- test hooks (not part of the drop-in surface) -------------------------
 * Host mirror of the device longest-match query over the table image, and table
 * statistics; used by the CPU unit tests of the vocabulary builder. */
wp_status wp_debug_longest_match(const wp_vocab *v, const char *text, size_t window_bytes, int kind,
                                 uint32_t *len_out, int32_t *id_out);
size_t wp_debug_table_slots(const wp_vocab *v);
size_t wp_debug_table_nodes(const wp_vocab *v);
size_t wp_debug_long_tokens(const wp_vocab *v);
/* chunk plan of the host-buffer pipeline (cut offsets, first 0, last n); 0 if the text cannot be cut */
size_t wp_debug_plan_chunks(const char *text, size_t n, size_t chunk, size_t *cuts, size_t cap);
/* code points of single-char word-initial nodes displaced from their home slot; returns their number */
size_t wp_debug_displaced_singles(const wp_vocab *v, uint32_t *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* WORDPIECE_B200_H_ */
