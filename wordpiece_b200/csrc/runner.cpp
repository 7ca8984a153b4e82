// Command-line driver of the GPU encoder.  Its argv contract is the one of the reference's
// tests/runner.cpp:13-65 — the binary tests/speed_test.py:92-101 spawns and times — for the modes
// this library covers:
//   runner fast          <text_file> <vocab_file> [n_threads] [out_file]
//   runner fast-external <text_file> <vocab_file> n_threads out_file memory_limit_mb
// Kept from that contract, quirks included: errors are uncaught std::runtime_error (the process
// aborts with the message); `n_threads` is read only when there are exactly five arguments
// (runner.cpp:23) and only sizes utils::globalThreadPool, which the GPU path does not use;
// memory_limit_mb below 50 is refused and the value is decimal megabytes (runner.cpp:27-32);
// `fast` prints "Total ids N" and writes the ids only when an out_file is given (runner.cpp:38-43).
// The suffix-array modes (linear, linear-external) are not part of this library and are refused.
#include <cstddef>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "src/utils.hpp"
#include "src/word_piece.hpp"

namespace {

struct Args {
  std::string mode, text_file, vocab_file, out_file;
  size_t n_threads = 0;
  size_t memory_limit = 0;  // bytes
  bool has_out_file = false, has_memory_limit = false;
};

[[noreturn]] void refuse(const std::string &why) { throw std::runtime_error(why); }

Args parse(int argc, char *argv[]) {
  const int n = argc - 1;  // arguments after the program name
  if (n < 3 || n > 6)
    refuse("Usage: ./runner <mode> <text_file> <vocab_file> [n_threads] [out_file] [memory_limit_mb]. "
           "Modes: fast, fast-external.");
  Args a;
  a.mode = argv[1];
  a.text_file = argv[2];
  a.vocab_file = argv[3];
  if (n == 4) a.n_threads = std::stoull(argv[4]);
  if (n >= 5) {
    a.out_file = argv[5];
    a.has_out_file = true;
  }
  if (n == 6) {
    const size_t mb = std::stoull(argv[6]);
    if (mb < 50) refuse("memory_limit cannot be less than 50Mb");
    a.memory_limit = mb * 1'000'000;
    a.has_memory_limit = true;
  }
  return a;
}

void run_fast(const Args &a) {
  const std::vector<int> ids = word_piece::fast::encode(a.text_file, a.vocab_file);
  std::cout << "Total ids " << ids.size() << std::endl;
  if (a.has_out_file) utils::writeToFile(a.out_file, ids);
}

void run_fast_external(const Args &a) {
  if (!a.has_memory_limit) refuse("For external mode provide out_file and memory_limit");
  word_piece::fast::encodeExternal(a.text_file, a.vocab_file, a.out_file, a.memory_limit);
}

void run_linear(const Args &a) {
  refuse("mode '" + a.mode + "' (suffix-array encoder) is not part of wordpiece_b200");
}

}  // namespace

int main(int argc, char *argv[]) {
  const Args a = parse(argc, argv);
  utils::globalThreadPool(a.n_threads);
  static const struct {
    const char *name;
    void (*run)(const Args &);
  } kModes[] = {{"fast", run_fast}, {"fast-external", run_fast_external}, {"linear", run_linear},
                {"linear-external", run_linear}};
  for (const auto &m : kModes) {
    if (a.mode == m.name) {
      m.run(a);
      return 0;
    }
  }
  refuse("Unknown mode");
}
