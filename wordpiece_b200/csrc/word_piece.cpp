// Host shim: the reference's C++ entry points (include/word_piece.hpp) on top of
// the C ABI (include/wordpiece_b200.h).  Mirrors src/fast.cpp:143-220 of
// gleb-kov/wordpiece; the encoding itself happens on the GPU.
//
// The reference is stateless — it re-parses the vocabulary and rebuilds both hash
// maps on every call (fast.cpp:154-157, :21-35; ~14 ms for a 29k vocabulary).
// Here the device table is built once per distinct vocabulary and cached by a
// content hash, so repeated calls with the same vocabulary only pay for the text.
#include "word_piece.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <list>
#include <mutex>
#include <stdexcept>
#include <thread>

#include "word_piece_utils.hpp"
#include "wordpiece_b200.h"

namespace {

int shim_device() {
  const char *e = std::getenv("WORDPIECE_B200_DEVICE");
  return e ? std::atoi(e) : 0;
}

[[noreturn]] void raise(wp_status st) {
  // utils.cpp:100 throws exactly this text; everything else is a runtime_error too (fast path has no other throws
  // except Boost's mmap failure).
  if (st == WP_ERR_EMPTY_VOCAB_WORD) throw std::runtime_error("Vocab word is empty");
  if (st == WP_ERR_ID_RANGE) throw std::out_of_range(wp_last_error());
  throw std::runtime_error(std::string("wordpiece_b200: ") + wp_last_error());
}

uint64_t fnv1a(const void *p, size_t n, uint64_t h) {
  const unsigned char *b = static_cast<const unsigned char *>(p);
  for (size_t i = 0; i < n; i++) {
    h ^= b[i];
    h *= 1099511628211ull;
  }
  return h;
}

struct CacheEntry {
  uint64_t key;
  size_t n_tokens;
  size_t n_bytes;
  std::vector<std::string> tokens;  // the vocabulary itself: a hit is confirmed by comparing contents, not by the hash
  wp_vocab *handle;
};

// Small LRU of device vocabularies.  Guarded by one mutex which is also held
// while a handle is in use: the reference's API is not re-entrant either
// (one process-global pool whose waitCompletion waits for all tasks,
// thread_pool.hpp:72-77).
class VocabCache {
 public:
  ~VocabCache() {
    for (auto &e : entries_) wp_vocab_destroy(e.handle);
  }
  std::mutex mu;

  wp_vocab *get(const std::vector<std::string> &vocab) {
    uint64_t h = 14695981039346656037ull;
    size_t bytes = 0;
    for (const std::string &t : vocab) {
      const uint64_t len = t.size();
      h = fnv1a(&len, sizeof(len), h);
      h = fnv1a(t.data(), t.size(), h);
      bytes += t.size();
    }
    for (auto it = entries_.begin(); it != entries_.end(); ++it) {
      if (it->key == h && it->n_tokens == vocab.size() && it->n_bytes == bytes && it->tokens == vocab) {
        entries_.splice(entries_.begin(), entries_, it);
        return entries_.front().handle;
      }
    }
    std::vector<const char *> ptrs(vocab.size());
    std::vector<size_t> lens(vocab.size());
    for (size_t i = 0; i < vocab.size(); i++) {
      ptrs[i] = vocab[i].data();
      lens[i] = vocab[i].size();
    }
    wp_vocab *handle = nullptr;
    const wp_status st = wp_vocab_create(ptrs.data(), lens.data(), vocab.size(), shim_device(), &handle);
    if (st != WP_OK) raise(st);
    for (size_t i = 0; i < vocab.size(); i++) {
      const int fl = wp_vocab_token_flags(handle, i);
      if (fl & 4) std::cerr << "Vocab word is malformed: " << vocab[i] << std::endl;  // utils.cpp:104
    }
    entries_.push_front(CacheEntry{h, vocab.size(), bytes, vocab, handle});
    if (entries_.size() > kMaxEntries) {
      wp_vocab_destroy(entries_.back().handle);
      entries_.pop_back();
    }
    return handle;
  }

 private:
  static constexpr size_t kMaxEntries = 4;
  std::list<CacheEntry> entries_;
};

VocabCache &cache() {
  static VocabCache c;
  return c;
}

// utils.cpp:123-137: std::getline over the file; no check that it opened.
std::vector<std::string> read_vocab_lines(const std::string &file) {
  std::vector<std::string> lines;
  std::ifstream fin(file);
  std::string word;
  while (std::getline(fin, word)) lines.push_back(word);
  return lines;
}

// Read-only mapping of a text file (the reference uses boost::iostreams::mapped_file, fast.cpp:161,196).
class MappedFile {
 public:
  explicit MappedFile(const std::string &path) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("cannot open file: " + path);
    struct stat st;
    if (::fstat(fd, &st) != 0) {
      ::close(fd);
      throw std::runtime_error("cannot stat file: " + path);
    }
    size_ = static_cast<size_t>(st.st_size);
    if (size_ > 0) {
      void *p = ::mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
      if (p == MAP_FAILED) {
        ::close(fd);
        throw std::runtime_error("cannot map file: " + path);
      }
      data_ = static_cast<const char *>(p);
    }
    ::close(fd);
  }
  MappedFile(const MappedFile &) = delete;
  MappedFile &operator=(const MappedFile &) = delete;
  ~MappedFile() {
    if (data_) ::munmap(const_cast<char *>(data_), size_);
  }
  const char *data() const { return data_; }
  size_t size() const { return size_; }

 private:
  const char *data_ = nullptr;
  size_t size_ = 0;
};

// A std::vector<int> of n elements whose storage has NOT been written: std::vector<int>(n) zero-fills, which
// for the 1.1 GB of ids of a 1 GiB text means touching every page once with one thread (0.4 s, more than the
// whole encode) before the real data overwrites it.  libstdc++ keeps three pointers; after reserve() the
// library fills the storage (several threads, first touch included) and then moves the end pointer.  Other
// standard libraries take the portable road: ids into a scratch buffer, then assign().
#if defined(__GLIBCXX__)
struct VectorTail : std::vector<int> {
  void set_size(size_t n) { this->_M_impl._M_finish = this->_M_impl._M_start + n; }
};
constexpr bool kUninitializedVectors = true;
#else
constexpr bool kUninitializedVectors = false;
#endif

// The storage of a large fresh vector comes straight from mmap and is touched for the first time by whoever copies
// the ids into it: with 4 KiB pages that is 280 000 page faults for the ids of a 1 GiB text.  Asking for transparent
// huge pages (MADV_HUGEPAGE) makes them a few hundred.  Measured on B200 boxes (16 cores, THP and defrag in "madvise"
// mode; profiles/r2z_hugepage_ab.jsonl):
//  * Encoder::encodeBatch, where ONE thread (the driver's pageable device-to-host copy) touches the storage:
//    10 000 x 4 KiB texts into a fresh vector 22.5 ms without the advice, 10.8 ms with it -> on by default there;
//  * fast::encode on 1 GiB, where the sixteen threads of the copy pool touch it: SLOWER with the advice, 15.2 ->
//    14.0 GB/s in three interleaved pairs (the threads wait while single faults zero 2 MiB each) -> off by default.
// WORDPIECE_B200_HUGEPAGES=0 / =1 forces it off / on everywhere.  Only blocks the allocator certainly took from mmap
// (>= 64 MiB) are advised; where THP is off the advice changes nothing.
void advise_huge_pages(void *p, size_t bytes, bool by_default) {
#if defined(MADV_HUGEPAGE)
  constexpr uintptr_t kHuge = uintptr_t(2) << 20;
  if (p == nullptr || bytes < (size_t(64) << 20)) return;
  static const int forced = [] {
    const char *e = std::getenv("WORDPIECE_B200_HUGEPAGES");
    return e == nullptr ? -1 : (std::atoi(e) != 0 ? 1 : 0);
  }();
  if (forced == 0 || (forced < 0 && !by_default)) return;
  const uintptr_t lo = (reinterpret_cast<uintptr_t>(p) + kHuge - 1) & ~(kHuge - 1);
  const uintptr_t hi = (reinterpret_cast<uintptr_t>(p) + bytes) & ~(kHuge - 1);
  if (hi > lo) (void)::madvise(reinterpret_cast<void *>(lo), hi - lo, MADV_HUGEPAGE);
#else
  (void)p;
  (void)bytes;
  (void)by_default;
#endif
}

// fast.cpp:143-150
std::vector<int> encode_buffer(wp_vocab *v, const char *text, size_t size) {
  if (size == 0) return {};
  static_assert(sizeof(int) == sizeof(int32_t), "ids are 32-bit");
  std::vector<int> out;
  // first guess: half an id per byte (English-like text needs a quarter; small texts get the worst case, one
  // id per byte, right away); the call reports the exact count if the guess was short
  size_t cap = size <= (size_t(1) << 20) ? size : size / 2 + 1024;
  for (int attempt = 0;; attempt++) {
    size_t n = 0;
    wp_status st;
    if (kUninitializedVectors) {
      out.reserve(cap);
      advise_huge_pages(out.data(), out.capacity() * sizeof(int), /*by_default=*/false);
      st = wp_encode_into(v, text, size, reinterpret_cast<int32_t *>(out.data()), cap, &n);
#if defined(__GLIBCXX__)
      if (st == WP_OK) static_cast<VectorTail &>(out).set_size(n);
#endif
    } else {
      std::vector<int32_t> tmp(cap);
      st = wp_encode_into(v, text, size, tmp.data(), cap, &n);
      if (st == WP_OK) out.assign(tmp.begin(), tmp.begin() + static_cast<std::ptrdiff_t>(n));
    }
    if (st == WP_ERR_CAPACITY && attempt == 0) {
      cap = n;
      continue;
    }
    if (st != WP_OK) raise(st);
    break;
  }
  wp_stats stats{};
  wp_last_stats(v, &stats);
  if (stats.dirty_tiles > 0)  // utf8.cpp:143-145
    std::cerr << "WARNING Input contains invalid unicode characters." << std::endl;
  return out;
}

// utf8.cpp:92-96 with :54-90 — does the sequence at `p` decode to an is_space char?
bool starts_with_space(const char *p, size_t size) {
  if (size == 0) return false;
  const unsigned char c = static_cast<unsigned char>(p[0]);
  if (c < 0x80) return (c >= 0x09 && c <= 0x0D) || c == 0x20;
  return size >= 3 && c == 0xE2 && static_cast<unsigned char>(p[1]) == 0x96 &&
         static_cast<unsigned char>(p[2]) == 0x81;  // U+2581
}

wp_vocab *create_handle(const std::vector<std::string> &vocab, int device) {
  std::vector<const char *> ptrs(vocab.size());
  std::vector<size_t> lens(vocab.size());
  for (size_t i = 0; i < vocab.size(); i++) {
    ptrs[i] = vocab[i].data();
    lens[i] = vocab[i].size();
  }
  wp_vocab *handle = nullptr;
  const wp_status st = wp_vocab_create(ptrs.data(), lens.data(), vocab.size(), device, &handle);
  if (st != WP_OK) raise(st);
  for (size_t i = 0; i < vocab.size(); i++) {
    if (wp_vocab_token_flags(handle, i) & 4) std::cerr << "Vocab word is malformed: " << vocab[i] << std::endl;  // utils.cpp:104
  }
  return handle;
}

std::vector<std::string> decode_with(wp_vocab *v, const std::vector<int> &ids) {
  const size_t size = wp_vocab_size(v);
  for (int id : ids) {  // fast.cpp:171-178 diagnostics
    if (id < 0 || static_cast<size_t>(id) > size) {
      std::cerr << "no token " << id << std::endl;
    } else if (static_cast<size_t>(id) < size && (wp_vocab_token_flags(v, static_cast<size_t>(id)) & 4)) {
      std::cerr << "trying to access malformed token" << std::endl;
    }
  }
  char *buf = nullptr;
  size_t *offs = nullptr;
  size_t n = 0;
  static_assert(sizeof(int) == sizeof(int32_t), "ids are 32-bit");
  const wp_status st = wp_decode(v, reinterpret_cast<const int32_t *>(ids.data()), ids.size(), &buf, &offs, &n, nullptr);
  if (st != WP_OK) raise(st);
  std::vector<std::string> result;
  result.reserve(n);
  for (size_t i = 0; i < n; i++) result.emplace_back(buf + offs[i], offs[i + 1] - offs[i]);
  wp_free(buf);
  wp_free(offs);
  return result;
}

}  // namespace

namespace word_piece {
namespace fast {

std::vector<int> encode(const std::string &text, const std::vector<std::string> &vocab) {
  std::lock_guard<std::mutex> lock(cache().mu);
  wp_vocab *v = cache().get(vocab);
  return encode_buffer(v, text.data(), text.size());
}

std::vector<int> encode(const std::string &text_file, const std::string &vocab_file) {
  std::lock_guard<std::mutex> lock(cache().mu);
  wp_vocab *v = cache().get(read_vocab_lines(vocab_file));
  MappedFile map(text_file);
  return encode_buffer(v, map.data(), map.size());
}

std::vector<std::string> decode(const std::string vocab_file, const std::vector<int> &ids) {
  std::lock_guard<std::mutex> lock(cache().mu);
  return decode_with(cache().get(read_vocab_lines(vocab_file)), ids);
}

void encodeExternal(const std::string &text_file,
                    const std::string &vocab_file,
                    const std::string &out_file,
                    size_t memory_limit) {
  std::lock_guard<std::mutex> lock(cache().mu);
  wp_vocab *v = cache().get(read_vocab_lines(vocab_file));
  const size_t max_text_batch = memory_limit / 2;  // fast.cpp:194
  MappedFile map(text_file);
  const char *begin = map.data();
  size_t size = map.size();
  std::ofstream fout(out_file);
  while (size > 0) {
    size_t batch;
    if (size > max_text_batch) {  // fast.cpp:202-211: extend until the last byte of the batch starts a space
      batch = max_text_batch;
      if (batch == 0) batch = 1;
      while (batch < size && !starts_with_space(begin + batch - 1, size - batch)) batch++;
    } else {
      batch = size;
    }
    // fast.cpp:212-216: encode the batch and append "id id id " — the ids are formatted on the device
    char *txt = nullptr;
    size_t len = 0, n = 0;
    const wp_status st = wp_encode_text(v, begin, batch, &txt, &len, &n);
    if (st != WP_OK) raise(st);
    wp_stats stats{};
    wp_last_stats(v, &stats);
    if (stats.dirty_tiles > 0)  // utf8.cpp:143-145
      std::cerr << "WARNING Input contains invalid unicode characters." << std::endl;
    fout.write(txt, static_cast<std::streamsize>(len));
    wp_free(txt);
    begin += batch;
    size -= batch;
  }
}

// ---- Encoder (extension): one vocabulary, resident on the GPU for the object's lifetime

Encoder::Encoder(const std::vector<std::string> &vocab, int device) : handle_(create_handle(vocab, device)) {}

Encoder Encoder::fromFile(const std::string &vocab_file, int device) {
  return Encoder(static_cast<void *>(create_handle(read_vocab_lines(vocab_file), device)));
}

Encoder::~Encoder() { wp_vocab_destroy(static_cast<wp_vocab *>(handle_)); }

Encoder::Encoder(Encoder &&other) noexcept : handle_(other.handle_) { other.handle_ = nullptr; }

Encoder &Encoder::operator=(Encoder &&other) noexcept {
  if (this != &other) {
    wp_vocab_destroy(static_cast<wp_vocab *>(handle_));
    handle_ = other.handle_;
    other.handle_ = nullptr;
  }
  return *this;
}

std::vector<int> Encoder::encode(const std::string &text) const {
  return encode_buffer(static_cast<wp_vocab *>(handle_), text.data(), text.size());
}

void Encoder::encodeBatch(const std::vector<std::string> &texts, std::vector<int> &ids, std::vector<size_t> &offsets) const {
  wp_vocab *v = static_cast<wp_vocab *>(handle_);
  std::vector<const char *> ptrs(texts.size());
  std::vector<size_t> lens(texts.size());
  size_t bytes = 0;
  for (size_t i = 0; i < texts.size(); i++) {
    ptrs[i] = texts[i].data();
    lens[i] = texts[i].size();
    bytes += lens[i];
  }
  offsets.assign(texts.size() + 1, 0);
  // first guess: half an id per byte (English-like text needs a quarter); the call reports the exact count if short
  size_t cap = bytes / 2 + 64;
  for (int attempt = 0;; attempt++) {
    size_t n = 0;
    wp_status st;
    if (kUninitializedVectors) {
      // no zero-fill of the guess (resize() would touch 2 bytes of ids per text byte with one thread, several
      // times what the call itself takes); a vector the caller reuses keeps its storage and its touched pages
      ids.clear();
      ids.reserve(cap);
      advise_huge_pages(ids.data(), ids.capacity() * sizeof(int), /*by_default=*/true);
      st = wp_encode_batch(v, ptrs.data(), lens.data(), texts.size(), reinterpret_cast<int32_t *>(ids.data()),
                           ids.capacity(), offsets.data(), &n);
#if defined(__GLIBCXX__)
      if (st == WP_OK) static_cast<VectorTail &>(ids).set_size(n);
#endif
    } else {
      ids.resize(cap);
      st = wp_encode_batch(v, ptrs.data(), lens.data(), texts.size(), reinterpret_cast<int32_t *>(ids.data()),
                           ids.size(), offsets.data(), &n);
      if (st == WP_OK) ids.resize(n);
    }
    if (st == WP_ERR_CAPACITY && attempt == 0) {
      cap = n;
      continue;
    }
    if (st != WP_OK) raise(st);
    break;
  }
  wp_stats stats{};
  wp_last_stats(v, &stats);
  if (stats.dirty_tiles > 0)  // utf8.cpp:143-145
    std::cerr << "WARNING Input contains invalid unicode characters." << std::endl;
}

std::vector<std::vector<int>> Encoder::encodeBatch(const std::vector<std::string> &texts) const {
  std::vector<int> ids;
  std::vector<size_t> offsets;
  encodeBatch(texts, ids, offsets);
  std::vector<std::vector<int>> out(texts.size());
  for (size_t i = 0; i < texts.size(); i++) out[i].assign(ids.begin() + offsets[i], ids.begin() + offsets[i + 1]);
  return out;
}

std::vector<std::string> Encoder::decode(const std::vector<int> &ids) const {
  return decode_with(static_cast<wp_vocab *>(handle_), ids);
}

size_t Encoder::vocabSize() const { return wp_vocab_size(static_cast<wp_vocab *>(handle_)); }

}  // namespace fast
}  // namespace word_piece

namespace utils {

ThreadPool::ThreadPool(size_t n_threads) : n_threads_(n_threads) {
  if (n_threads_ == 0) {
    n_threads_ = std::thread::hardware_concurrency();
    if (n_threads_ == 0) n_threads_ = 8;  // thread_pool.hpp:24-29
  }
}

ThreadPool &globalThreadPool(size_t n_threads) {
  static ThreadPool pool(n_threads);  // frozen by the first caller, utils.cpp:25-28
  return pool;
}

void writeToFile(const std::string &file, const std::vector<int> &ids) {
  std::ofstream fout(file);
  std::string line;
  for (int id : ids) {
    line += std::to_string(id);
    line.push_back(' ');
    if (line.size() > (1u << 16)) {
      fout << line;
      line.clear();
    }
  }
  fout << line;
}

int64_t currentTs() {
  return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch())
      .count();
}

}  // namespace utils
