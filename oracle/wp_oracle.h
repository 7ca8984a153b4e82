/* TEST INFRASTRUCTURE ONLY — CPU oracle for the fast WordPiece encode path.
 *
 * A plain-C restatement, in the code-point domain, of the algorithm that
 * gleb-kov/wordpiece implements in src/fast.cpp, src/utils.cpp and
 * src/third_party/utf8.{hpp,cpp}.  It exists to CHECK the CUDA path; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 * The product (wordpiece_b200/) never links, loads or falls back to it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (1) the 28 golden id vectors of the reference's tests/tests.cpp:137-217,
 *   (2) tests/golden/*.json — outputs of the UNMODIFIED reference compiled here
 *       (oracle/_ref/libwpref.so; generator: tests/golden/make_golden.py),
 *   (3) when oracle/_ref is present, live differential fuzzing against it.
 */
#ifndef WP_ORACLE_H_
#define WP_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wpo_vocab wpo_vocab;

enum { WPO_OK = 0, WPO_ERR_EMPTY_WORD = 1, WPO_ERR_NOMEM = 2 };

/* utf8.cpp:10-29 — character classes (C locale). */
int wpo_is_space(uint32_t cp);
int wpo_is_punct(uint32_t cp);
int wpo_is_han(uint32_t cp);
int wpo_is_spacing(uint32_t cp);

/* utf8.cpp:54-90,130-147 — strict decode, invalid bytes dropped.  out must hold
 * n_bytes entries; returns the number of code points; *had_invalid set to 1 if
 * at least one byte was dropped. */
size_t wpo_decode_utf8(const char *bytes, size_t n_bytes, uint32_t *out, int *had_invalid);

/* utils.cpp:81-137 + fast.cpp:21-36 — token classification and the two maps.
 * Returns WPO_ERR_EMPTY_WORD where the reference throws "Vocab word is empty". */
int wpo_vocab_create(const char *const *toks, const size_t *lens, size_t n, wpo_vocab **out);
void wpo_vocab_free(wpo_vocab *v);
int32_t wpo_vocab_unk_id(const wpo_vocab *v);
size_t wpo_vocab_max_len(const wpo_vocab *v);
/* per-token classification: bit0 prefix, bit1 special, bit2 malformed */
int wpo_vocab_token_flags(const wpo_vocab *v, size_t i);

/* fast.cpp:143-150 + :19-99 — encode; *ids is malloc'd (free with wpo_free). */
int wpo_encode(const wpo_vocab *v, const char *text, size_t n_bytes, int32_t **ids, size_t *n_ids);
void wpo_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* WP_ORACLE_H_ */
