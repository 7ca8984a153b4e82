/* TEST INFRASTRUCTURE ONLY — see wp_oracle.h.  Code-point-domain restatement
 * of the reference fast path; every function cites the reference lines it
 * follows (paths relative to /root/reference). */
#include "wp_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ classes */

/* utf8.cpp:10-12 — C-locale isspace for cp < 256 (09-0D, 20) plus U+2581. */
int wpo_is_space(uint32_t cp) { return (cp >= 0x09 && cp <= 0x0D) || cp == 0x20 || cp == 0x2581; }

/* utf8.cpp:14-17 — C-locale ispunct for cp < 256 (the four ASCII ranges; no
 * byte >= 0x80 is punctuation in the C locale) plus the explicit list. */
int wpo_is_punct(uint32_t cp) {
  if (cp < 0x80) {
    return (cp >= 0x21 && cp <= 0x2F) || (cp >= 0x3A && cp <= 0x40) || (cp >= 0x5B && cp <= 0x60) ||
           (cp >= 0x7B && cp <= 0x7E);
  }
  return cp == 183 || cp == 171 || cp == 187 || cp == 8249 || cp == 8250 || (cp >= 8208 && cp <= 8248);
}

/* utf8.cpp:19-27 */
int wpo_is_han(uint32_t cp) {
  return (cp >= 0x4E00 && cp <= 0x9FFF) || (cp >= 0x3400 && cp <= 0x4DBF) || (cp >= 0x20000 && cp <= 0x2A6DF) ||
         (cp >= 0x2A700 && cp <= 0x2B73F) || (cp >= 0x2B740 && cp <= 0x2B81F) || (cp >= 0x2B820 && cp <= 0x2CEAF) ||
         (cp >= 0xF900 && cp <= 0xFAFF) || (cp >= 0x2F800 && cp <= 0x2FA1F);
}

/* utf8.cpp:29 */
int wpo_is_spacing(uint32_t cp) { return wpo_is_space(cp) || wpo_is_punct(cp) || wpo_is_han(cp); }

/* ------------------------------------------------------------------- decode */

/* utf8.cpp:37-52 (lead-byte length), :54-90 (one sequence), :130-147 (loop).
 * A sequence is accepted iff its lead announces 1..4 bytes, that many bytes
 * are available, every trail byte is 10xxxxxx, the value is not overlong, not
 * a surrogate and below 0x110000.  Anything else: consume ONE byte, emit none. */
size_t wpo_decode_utf8(const char *bytes, size_t n_bytes, uint32_t *out, int *had_invalid) {
  const unsigned char *b = (const unsigned char *)bytes;
  size_t p = 0, n = 0;
  int bad = 0;
  while (p < n_bytes) {
    unsigned c = b[p];
    size_t len = c < 0x80 ? 1 : (c & 0xE0) == 0xC0 ? 2 : (c & 0xF0) == 0xE0 ? 3 : (c & 0xF8) == 0xF0 ? 4 : 0;
    if (len == 1) {
      out[n++] = c;
      p += 1;
      continue;
    }
    int ok = len != 0 && n_bytes - p >= len;
    uint32_t cp = 0;
    if (ok) {
      static const unsigned lead_mask[5] = {0, 0, 0x1F, 0x0F, 0x07};
      static const uint32_t min_value[5] = {0, 0, 0x80, 0x800, 0x10000};
      cp = c & lead_mask[len];
      for (size_t k = 1; k < len; k++) {
        if ((b[p + k] & 0xC0) != 0x80) ok = 0;
        cp = (cp << 6) | (b[p + k] & 0x3F);
      }
      if (cp < min_value[len]) ok = 0;                 /* overlong            */
      if (cp >= 0xD800 && cp <= 0xDFFF) ok = 0;        /* surrogate  :35      */
      if (cp >= 0x110000) ok = 0;                      /* out of range :35    */
    }
    if (ok) {
      out[n++] = cp;
      p += len;
    } else {
      bad = 1;
      p += 1;
    }
  }
  if (had_invalid) *had_invalid = bad;
  return n;
}

/* -------------------------------------------------------------------- vocab */

typedef struct {
  uint32_t *cps; /* word without the leading ## */
  size_t len;
  int is_prefix, is_special, is_malformed;
} wpo_token;

typedef struct {
  /* open-addressed exact dictionary over code-point strings; value = token id.
   * The reference uses unordered_map<VectorSegment,int> with a polynomial hash
   * (utf8.hpp:75-111); the hash value never reaches the output, so any exact
   * dictionary is equivalent. */
  int64_t *slot_tok; /* index of the token whose string is the key, -1 = empty */
  int32_t *slot_id;  /* mapped id (last duplicate wins, fast.cpp:34) */
  size_t cap;
} wpo_map;

struct wpo_vocab {
  wpo_token *tok;
  size_t n;
  int32_t unk;
  size_t max_len;
  wpo_map map[2]; /* [0] word-initial (prefix_to_id), [1] ## (suffix_to_id) */
};

static uint64_t wpo_hash(const uint32_t *s, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; i++) {
    h ^= s[i];
    h *= 1099511628211ull;
  }
  return h ^ (h >> 29);
}

static void map_put(wpo_map *m, const wpo_token *toks, size_t ti, int32_t id) {
  const wpo_token *t = &toks[ti];
  size_t i = (size_t)wpo_hash(t->cps, t->len) & (m->cap - 1);
  for (;;) {
    int64_t o = m->slot_tok[i];
    if (o < 0) {
      m->slot_tok[i] = (int64_t)ti;
      m->slot_id[i] = id;
      return;
    }
    if (toks[o].len == t->len && memcmp(toks[o].cps, t->cps, t->len * sizeof(uint32_t)) == 0) {
      m->slot_id[i] = id; /* duplicate key: operator[] assignment overwrites */
      return;
    }
    i = (i + 1) & (m->cap - 1);
  }
}

static int map_get(const wpo_map *m, const wpo_token *toks, const uint32_t *s, size_t n, int32_t *id) {
  size_t i = (size_t)wpo_hash(s, n) & (m->cap - 1);
  for (;;) {
    int64_t o = m->slot_tok[i];
    if (o < 0) return 0;
    if (toks[o].len == n && memcmp(toks[o].cps, s, n * sizeof(uint32_t)) == 0) {
      *id = m->slot_id[i];
      return 1;
    }
    i = (i + 1) & (m->cap - 1);
  }
}

void wpo_vocab_free(wpo_vocab *v) {
  if (!v) return;
  for (size_t i = 0; i < v->n; i++) free(v->tok[i].cps);
  free(v->tok);
  for (int k = 0; k < 2; k++) {
    free(v->map[k].slot_tok);
    free(v->map[k].slot_id);
  }
  free(v);
}

int wpo_vocab_create(const char *const *toks, const size_t *lens, size_t n, wpo_vocab **out) {
  wpo_vocab *v = (wpo_vocab *)calloc(1, sizeof(*v));
  if (!v) return WPO_ERR_NOMEM;
  v->tok = (wpo_token *)calloc(n ? n : 1, sizeof(wpo_token));
  v->n = n;
  v->unk = -1; /* utils.hpp:30 kDefaultUnkTokenId */
  size_t cap = 16;
  while (cap < 4 * n) cap <<= 1;
  for (int k = 0; k < 2; k++) {
    v->map[k].cap = cap;
    v->map[k].slot_tok = (int64_t *)malloc(cap * sizeof(int64_t));
    v->map[k].slot_id = (int32_t *)malloc(cap * sizeof(int32_t));
    for (size_t i = 0; i < cap; i++) v->map[k].slot_tok[i] = -1;
  }
  for (size_t i = 0; i < n; i++) {
    /* utils.cpp:112-114 / :129-131 — the LAST line equal to "[UNK]" gives the id */
    if (lens[i] == 5 && memcmp(toks[i], "[UNK]", 5) == 0) v->unk = (int32_t)i;

    /* utils.cpp:81-106 — WordPieceToken constructor */
    wpo_token *t = &v->tok[i];
    uint32_t *cps = (uint32_t *)malloc((lens[i] + 1) * sizeof(uint32_t));
    size_t len = wpo_decode_utf8(toks[i], lens[i], cps, NULL);
    t->is_prefix = 1;
    if (len >= 2 && cps[0] == '#' && cps[1] == '#') { /* utils.cpp:139-141 */
      t->is_prefix = 0;
      memmove(cps, cps + 2, (len - 2) * sizeof(uint32_t));
      len -= 2;
    } else if (len > 2 && cps[0] == '[' && cps[len - 1] == ']') { /* utils.cpp:143-146 */
      t->is_special = 1;
    }
    t->cps = cps;
    t->len = len;
    if (len == 0) { /* utils.cpp:99-101 */
      v->n = i + 1;
      wpo_vocab_free(v);
      return WPO_ERR_EMPTY_WORD;
    }
    int all_punct = 1;
    for (size_t k = 0; k < len; k++)
      if (!wpo_is_punct(cps[k]) && !wpo_is_space(cps[k])) all_punct = 0;
    t->is_malformed = all_punct && len > 1; /* utils.cpp:102-105 */

    /* fast.cpp:26-35 */
    if (t->is_special || t->is_malformed) continue;
    if (len > v->max_len) v->max_len = len;
    map_put(&v->map[t->is_prefix ? 0 : 1], v->tok, i, (int32_t)i);
  }
  *out = v;
  return WPO_OK;
}

int32_t wpo_vocab_unk_id(const wpo_vocab *v) { return v->unk; }
size_t wpo_vocab_max_len(const wpo_vocab *v) { return v->max_len; }
int wpo_vocab_token_flags(const wpo_vocab *v, size_t i) {
  return (v->tok[i].is_prefix ? 1 : 0) | (v->tok[i].is_special ? 2 : 0) | (v->tok[i].is_malformed ? 4 : 0);
}

/* ------------------------------------------------------------------- encode */

void wpo_free(void *p) { free(p); }

typedef struct {
  int32_t *v;
  size_t n, cap;
} ivec;

static int ivec_push(ivec *a, int32_t x) {
  if (a->n == a->cap) {
    size_t nc = a->cap ? a->cap * 2 : 64;
    int32_t *nv = (int32_t *)realloc(a->v, nc * sizeof(int32_t));
    if (!nv) return 0;
    a->v = nv;
    a->cap = nc;
  }
  a->v[a->n++] = x;
  return 1;
}

/* fast.cpp:143-150 (entry) and :19-99 (the serial worker; the chunked form
 * :101-138 yields the same ids because cuts are at is_space code points).
 *
 * Outside the reference's domain (it divides by max_len == 0 at fast.cpp:45
 * when the text decodes to zero code points or no token is usable) this
 * restatement continues with the natural reading of the loop: no code points
 * -> no ids; no usable token -> every window has length 1 and misses. */
int wpo_encode(const wpo_vocab *v, const char *text, size_t n_bytes, int32_t **ids, size_t *n_ids) {
  *ids = NULL;
  *n_ids = 0;
  if (n_bytes == 0) return WPO_OK; /* fast.cpp:145 */
  uint32_t *cp = (uint32_t *)malloc(n_bytes * sizeof(uint32_t));
  if (!cp) return WPO_ERR_NOMEM;
  const size_t n = wpo_decode_utf8(text, n_bytes, cp, NULL); /* utils.cpp:37-79 */
  size_t max_len = v->max_len < n ? v->max_len : n;        /* fast.cpp:36 */
  ivec out = {0, 0, 0};
  size_t i = 0, since_prefix = 0;

#define WP(ix) ((ix) == 0 || wpo_is_spacing(cp[ix]) || wpo_is_spacing(cp[(ix)-1])) /* fast.cpp:38-41 */
  while (i < n && wpo_is_space(cp[i])) i++; /* :47-49 */
  while (i < n) {
    size_t win = 1; /* :54-60 */
    if (!wpo_is_punct(cp[i])) {
      size_t lim = max_len < n - i ? max_len : n - i;
      while (win < lim && !wpo_is_spacing(cp[i + win])) win++;
    }
    const wpo_map *m = &v->map[WP(i) ? 0 : 1]; /* :64 */
    size_t k = win;
    int32_t id = 0;
    while (k > 0 && !map_get(m, v->tok, cp + i, k, &id)) k--; /* :66-77 longest first */
    if (k > 0) { /* :69-73 */
      since_prefix++;
      ivec_push(&out, id);
      i += k;
      if (i < n && WP(i)) since_prefix = 0; /* :89-91 */
    } else { /* :79-88 whole-word UNK: roll back this word's pieces */
      out.n -= since_prefix;
      since_prefix = 0;
      ivec_push(&out, v->unk);
      i += win;
      while (i < n && !WP(i)) i++;
    }
    while (i < n && wpo_is_space(cp[i])) i++; /* :93-95 */
  }
#undef WP
  free(cp);
  if (out.n == 0) {
    free(out.v);
    out.v = NULL;
  }
  *ids = out.v;
  *n_ids = out.n;
  return WPO_OK;
}
