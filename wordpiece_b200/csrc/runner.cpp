// Command-line driver with the argv contract of the reference's tests/runner.cpp:13-65
// (what tests/speed_test.py:92-101 spawns), for the modes this library covers:
//   runner fast          <text_file> <vocab_file> [n_threads] [out_file]
//   runner fast-external <text_file> <vocab_file> n_threads out_file memory_limit_mb
// `n_threads` is accepted and ignored by the GPU path (it sizes the reference's
// CPU pool).  The suffix-array modes (linear, linear-external) are not part of
// this library and are rejected.
#include <iostream>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "src/utils.hpp"
#include "src/word_piece.hpp"

int main(int argc, char *argv[]) {
  if (argc < 4 || argc > 7) {
    throw std::runtime_error("Usage: ./runner <mode> <text_file> <vocab_file> [n_threads] "
                             "[out_file] [memory_limit_mb]. "
                             "Modes: fast, fast-external.");
  }
  const std::string mode = argv[1];
  const std::string text_file = argv[2];
  const std::string vocab_file = argv[3];
  const size_t n_threads = argc == 5 ? std::stoull(argv[4]) : 0;  // runner.cpp:23 (only honoured with exactly 5 args)
  const std::optional<std::string> out_file = argc >= 6 ? std::optional<std::string>(argv[5]) : std::nullopt;
  std::optional<size_t> memory_limit = argc >= 7 ? std::optional<size_t>(std::stoull(argv[6])) : std::nullopt;
  if (memory_limit.has_value()) {
    if (*memory_limit < 50) throw std::runtime_error("memory_limit cannot be less than 50Mb");
    *memory_limit *= 1'000'000;
  }
  [[maybe_unused]] auto &pool = utils::globalThreadPool(n_threads);

  if (mode == "fast") {
    const std::vector<int> ids = word_piece::fast::encode(text_file, vocab_file);
    std::cout << "Total ids " << ids.size() << std::endl;
    if (out_file) utils::writeToFile(*out_file, ids);
  } else if (mode == "fast-external") {
    if (!memory_limit.has_value()) throw std::runtime_error("For external mode provide out_file and memory_limit");
    word_piece::fast::encodeExternal(text_file, vocab_file, out_file.value(), memory_limit.value());
  } else if (mode == "linear" || mode == "linear-external") {
    throw std::runtime_error("mode '" + mode + "' (suffix-array encoder) is not part of wordpiece_b200");
  } else {
    throw std::runtime_error("Unknown mode");
  }
  return 0;
}
