// UTF-8 sequence validation and the reference's character classes, shared by
// the host vocab builder and the device kernels (byte-domain forms).
//
// Semantics follow gleb-kov/wordpiece src/third_party/utf8.cpp:
//   utf_length :37-52, chars_to_utf8 :54-90 (strict: no overlongs, no
//   surrogates, < 0x110000; on any violation ONE byte is consumed and nothing
//   is emitted, decode_utf8 :130-147), is_space :10-12, is_punctuation :14-17,
//   is_chinese :19-27, is_spacing_char :29 — all in the C locale.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define WP_HD __host__ __device__ __forceinline__
#else
#define WP_HD inline
#endif

namespace wp {

// Character classes.  SPACE/PUNCT/HAN are the reference's "spacing" chars.
enum CharClass : uint32_t { CLS_OTHER = 0, CLS_SPACE = 1, CLS_PUNCT = 2, CLS_HAN = 3 };

WP_HD bool is_cont_byte(uint32_t b) { return (b & 0xC0u) == 0x80u; }

// Announced sequence length of a lead byte; 0 for continuation bytes and >= 0xF8.
WP_HD uint32_t utf8_lead_len(uint32_t b) {
  if (b < 0x80u) return 1;
  if ((b & 0xE0u) == 0xC0u) return 2;
  if ((b & 0xF0u) == 0xE0u) return 3;
  if ((b & 0xF8u) == 0xF0u) return 4;
  return 0;
}

// Decode the sequence whose bytes are b0..b3 (`avail` of them exist).
// Returns the sequence length (1..4) and the code point, or 0 if invalid.
WP_HD uint32_t utf8_decode(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, uint32_t avail, uint32_t *cp) {
  const uint32_t len = utf8_lead_len(b0);
  if (len == 1) {
    *cp = b0;
    return 1;
  }
  if (len == 0 || avail < len) return 0;
  uint32_t v;
  if (len == 2) {
    if (!is_cont_byte(b1)) return 0;
    v = ((b0 & 0x1Fu) << 6) | (b1 & 0x3Fu);
    if (v < 0x80u) return 0;
  } else if (len == 3) {
    if (!is_cont_byte(b1) || !is_cont_byte(b2)) return 0;
    v = ((b0 & 0x0Fu) << 12) | ((b1 & 0x3Fu) << 6) | (b2 & 0x3Fu);
    if (v < 0x800u || (v >= 0xD800u && v <= 0xDFFFu)) return 0;
  } else {
    if (!is_cont_byte(b1) || !is_cont_byte(b2) || !is_cont_byte(b3)) return 0;
    v = ((b0 & 0x07u) << 18) | ((b1 & 0x3Fu) << 12) | ((b2 & 0x3Fu) << 6) | (b3 & 0x3Fu);
    if (v < 0x10000u || v >= 0x110000u) return 0;
  }
  *cp = v;
  return len;
}

WP_HD bool cp_is_space(uint32_t cp) { return (cp >= 0x09u && cp <= 0x0Du) || cp == 0x20u || cp == 0x2581u; }

WP_HD bool cp_is_punct(uint32_t cp) {
  if (cp < 0x80u)
    return (cp >= 0x21u && cp <= 0x2Fu) || (cp >= 0x3Au && cp <= 0x40u) || (cp >= 0x5Bu && cp <= 0x60u) ||
           (cp >= 0x7Bu && cp <= 0x7Eu);
  return cp == 0xABu || cp == 0xB7u || cp == 0xBBu || (cp >= 0x2010u && cp <= 0x203Au);
}

WP_HD bool cp_is_han(uint32_t cp) {
  return (cp >= 0x4E00u && cp <= 0x9FFFu) || (cp >= 0x3400u && cp <= 0x4DBFu) || (cp >= 0xF900u && cp <= 0xFAFFu) ||
         (cp >= 0x20000u && cp <= 0x2A6DFu) || (cp >= 0x2A700u && cp <= 0x2B73Fu) ||
         (cp >= 0x2B740u && cp <= 0x2B81Fu) || (cp >= 0x2B820u && cp <= 0x2CEAFu) ||
         (cp >= 0x2F800u && cp <= 0x2FA1Fu);
}

WP_HD uint32_t cp_class(uint32_t cp) {
  if (cp_is_space(cp)) return CLS_SPACE;
  if (cp_is_punct(cp)) return CLS_PUNCT;
  if (cp_is_han(cp)) return CLS_HAN;
  return CLS_OTHER;
}

// Canonical UTF-8 encoding of a valid code point (utf8.cpp:98-120); returns the length.
WP_HD uint32_t utf8_encode(uint32_t cp, uint8_t *out) {
  if (cp <= 0x7Fu) {
    out[0] = (uint8_t)cp;
    return 1;
  }
  if (cp <= 0x7FFu) {
    out[0] = (uint8_t)(0xC0u | (cp >> 6));
    out[1] = (uint8_t)(0x80u | (cp & 0x3Fu));
    return 2;
  }
  if (cp <= 0xFFFFu) {
    out[0] = (uint8_t)(0xE0u | (cp >> 12));
    out[1] = (uint8_t)(0x80u | ((cp >> 6) & 0x3Fu));
    out[2] = (uint8_t)(0x80u | (cp & 0x3Fu));
    return 3;
  }
  out[0] = (uint8_t)(0xF0u | (cp >> 18));
  out[1] = (uint8_t)(0x80u | ((cp >> 12) & 0x3Fu));
  out[2] = (uint8_t)(0x80u | ((cp >> 6) & 0x3Fu));
  out[3] = (uint8_t)(0x80u | (cp & 0x3Fu));
  return 4;
}

}  // namespace wp
