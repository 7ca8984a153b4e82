// The few names of the reference's src/utils.hpp that its drivers use besides
// the encoder itself (tests/runner.cpp:35,41; tests/tests.cpp:17), so that those
// drivers build against this library unchanged.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace utils {

// utils.hpp:29-34 — only the constant is part of the drivers' contract.
struct WordPieceVocabulary {
  static constexpr int kDefaultUnkTokenId = -1;
};

// thread_pool.hpp:16-89 / utils.cpp:25-28.  The CPU pool has no role on the GPU
// path; the object is kept so that `utils::globalThreadPool(n)` still compiles.
// maxThreads() reports the size the first caller asked for (0 => hardware
// concurrency, as in thread_pool.hpp:21-30).
class ThreadPool {
 public:
  explicit ThreadPool(size_t n_threads);
  size_t maxThreads() const { return n_threads_; }

 private:
  size_t n_threads_;
};

ThreadPool &globalThreadPool(size_t n_threads = 0);

// utils.cpp:30-35 — ids as decimal text, each followed by one space.
void writeToFile(const std::string &file, const std::vector<int> &ids);

// utils.cpp:19-23 — wall clock in milliseconds.
int64_t currentTs();

}  // namespace utils
