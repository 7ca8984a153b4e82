/* Synthetic corpus generator (bench / test data, not part of the encode path).
 *
 * There is no network in the build or GPU containers, so the workloads of
 * BASELINE.json are synthetic: a weighted lexicon of UTF-8 items and a weighted
 * set of separators are sampled into a byte buffer.  The corpus is defined in
 * independent 1 MiB blocks — block b is generated from hash(seed, b) alone — so
 * any block-aligned shard of a 10 GB corpus can be produced on its own rank, in
 * parallel, and is bit-identical to the same range of the whole corpus.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WP_SYNTH_BLOCK (1u << 20)

typedef struct {
  const uint8_t *bytes;
  const uint32_t *off;   /* n + 1 offsets */
  const uint32_t *prob;  /* alias method: acceptance threshold (scaled to 2^32) */
  const uint32_t *alias;
  uint32_t n;
} wp_table;

typedef struct {
  uint8_t *out;
  size_t n_bytes;
  size_t first_block;
  uint64_t seed;
  wp_table items, seps;
  uint32_t cap_threshold; /* P(capitalise an ASCII lower-case first letter) * 2^32 */
  size_t block_begin, block_end; /* blocks of this worker, relative */
} wp_job;

static inline uint64_t splitmix64(uint64_t *s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static inline uint32_t draw(const wp_table *t, uint64_t *s) {
  const uint64_t r = splitmix64(s);
  const uint32_t i = (uint32_t)(((r >> 32) * (uint64_t)t->n) >> 32);
  return ((uint32_t)r < t->prob[i]) ? i : t->alias[i];
}

static void fill_block(const wp_job *j, size_t rel_block) {
  const size_t begin = rel_block * (size_t)WP_SYNTH_BLOCK;
  if (begin >= j->n_bytes) return;
  size_t len = j->n_bytes - begin;
  if (len > WP_SYNTH_BLOCK) len = WP_SYNTH_BLOCK;
  uint8_t *o = j->out + begin;
  uint64_t s = j->seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(j->first_block + rel_block + 1));
  splitmix64(&s);
  /* the block is always generated as a full 1 MiB block; a shorter request takes its prefix */
  size_t pos = 0;
  uint8_t tmp[512];
  while (pos < WP_SYNTH_BLOCK) {
    const uint32_t wi = draw(&j->items, &s);
    const uint32_t si = draw(&j->seps, &s);
    const uint32_t capr = (uint32_t)splitmix64(&s);
    const uint32_t wl = j->items.off[wi + 1] - j->items.off[wi];
    const uint32_t sl = j->seps.off[si + 1] - j->seps.off[si];
    if (pos + wl + sl > WP_SYNTH_BLOCK || wl + sl > sizeof(tmp)) {
      /* does not fit: pad the block with spaces */
      for (size_t k = pos; k < WP_SYNTH_BLOCK && k < len; k++) o[k] = ' ';
      break;
    }
    memcpy(tmp, j->items.bytes + j->items.off[wi], wl);
    memcpy(tmp + wl, j->seps.bytes + j->seps.off[si], sl);
    if (capr < j->cap_threshold && wl > 0 && tmp[0] >= 'a' && tmp[0] <= 'z') tmp[0] = (uint8_t)(tmp[0] - 32);
    const size_t tot = (size_t)wl + sl;
    if (pos + tot <= len) {
      memcpy(o + pos, tmp, tot);
    } else if (pos < len) {
      memcpy(o + pos, tmp, len - pos);
    }
    pos += tot;
    if (pos >= len) break;
  }
}

static void *worker(void *arg) {
  const wp_job *j = (const wp_job *)arg;
  for (size_t b = j->block_begin; b < j->block_end; b++) fill_block(j, b);
  return NULL;
}

/* Fill out[0, n_bytes) with the bytes [first_block * 1 MiB, first_block * 1 MiB + n_bytes) of corpus `seed`. */
int wp_synth_fill(uint8_t *out, size_t n_bytes, size_t first_block, uint64_t seed, const uint8_t *item_bytes,
                  const uint32_t *item_off, const uint32_t *item_prob, const uint32_t *item_alias, uint32_t n_items,
                  const uint8_t *sep_bytes, const uint32_t *sep_off, const uint32_t *sep_prob,
                  const uint32_t *sep_alias, uint32_t n_seps, uint32_t cap_threshold, int n_threads) {
  if (n_items == 0 || n_seps == 0) return 1;
  wp_job base;
  base.out = out;
  base.n_bytes = n_bytes;
  base.first_block = first_block;
  base.seed = seed;
  base.items.bytes = item_bytes;
  base.items.off = item_off;
  base.items.prob = item_prob;
  base.items.alias = item_alias;
  base.items.n = n_items;
  base.seps.bytes = sep_bytes;
  base.seps.off = sep_off;
  base.seps.prob = sep_prob;
  base.seps.alias = sep_alias;
  base.seps.n = n_seps;
  base.cap_threshold = cap_threshold;
  const size_t n_blocks = (n_bytes + WP_SYNTH_BLOCK - 1) / WP_SYNTH_BLOCK;
  if (n_threads < 1) n_threads = 1;
  if ((size_t)n_threads > n_blocks) n_threads = (int)(n_blocks ? n_blocks : 1);
  if (n_threads == 1) {
    base.block_begin = 0;
    base.block_end = n_blocks;
    worker(&base);
    return 0;
  }
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  wp_job *jobs = (wp_job *)malloc(sizeof(wp_job) * (size_t)n_threads);
  if (!th || !jobs) return 2;
  for (int t = 0; t < n_threads; t++) {
    jobs[t] = base;
    jobs[t].block_begin = n_blocks * (size_t)t / (size_t)n_threads;
    jobs[t].block_end = n_blocks * (size_t)(t + 1) / (size_t)n_threads;
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
  return 0;
}
