// Times the reference's own C++ signature of this library end to end:
//   std::vector<int> word_piece::fast::encode(const std::string &text, const std::vector<std::string> &vocab)
// (include/word_piece.hpp, reference src/word_piece.hpp:27 / fast.cpp:154-157) with what a reference caller has:
// a pageable std::string in, a std::vector<int> out.  Prints one JSON line.
//
//   dropin_bench <text_file> <vocab_file> <reps> [<batch_texts> <batch_text_bytes>]
// With the two optional arguments it also times the C++ batch call of the extension class,
//   word_piece::fast::Encoder::encodeBatch(const std::vector<std::string>&, std::vector<int>&, std::vector<size_t>&),
// on <batch_texts> slices of about <batch_text_bytes> bytes cut at spaces ("batch": {...} in the JSON line).
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "word_piece.hpp"

int main(int argc, char **argv) {
  if (argc != 4 && argc != 6) {
    std::cerr << "usage: dropin_bench <text_file> <vocab_file> <reps> [<batch_texts> <batch_text_bytes>]" << std::endl;
    return 2;
  }
  std::string text;
  {
    std::ifstream in(argv[1], std::ios::binary | std::ios::ate);
    if (!in) return 2;
    const std::streamsize n = in.tellg();
    in.seekg(0);
    text.resize(static_cast<size_t>(n));
    in.read(&text[0], n);
  }
  std::vector<std::string> vocab;
  {
    std::ifstream in(argv[2]);
    std::string line;
    while (std::getline(in, line)) vocab.push_back(line);
  }
  const int reps = std::atoi(argv[3]);
  using clock = std::chrono::steady_clock;
  // warm-up: CUDA context, vocabulary tables, scratch (the reference pays its table build on every call; here it
  // is cached by vocabulary content after the first)
  size_t n_ids = word_piece::fast::encode(text.substr(0, text.size() < (1u << 20) ? text.size() : (1u << 20)), vocab).size();
  std::vector<double> secs;
  for (int r = 0; r < reps; r++) {
    const auto t0 = clock::now();
    const std::vector<int> ids = word_piece::fast::encode(text, vocab);
    secs.push_back(std::chrono::duration<double>(clock::now() - t0).count());
    n_ids = ids.size();
  }
  // the floor the signature itself sets: creating (and thereby touching) a result vector of that size
  const auto t0 = clock::now();
  {
    std::vector<int> probe(n_ids);
    volatile int sink = probe[n_ids / 2];
    (void)sink;
  }
  const double vec = std::chrono::duration<double>(clock::now() - t0).count();
  double best = secs.empty() ? 0 : secs[0];
  for (double s : secs) best = s < best ? s : best;
  std::cout << "{\"bytes\": " << text.size() << ", \"n_ids\": " << n_ids << ", \"reps\": " << reps
            << ", \"best_seconds\": " << best << ", \"vector_seconds\": " << vec << ", \"seconds\": [";
  for (size_t i = 0; i < secs.size(); i++) std::cout << (i ? ", " : "") << secs[i];
  std::cout << "]";
  if (argc == 6) {
    const size_t n_texts = static_cast<size_t>(std::atoll(argv[4])), size = static_cast<size_t>(std::atoll(argv[5]));
    std::vector<std::string> texts;
    size_t pos = 0, batch_bytes = 0;
    while (texts.size() < n_texts && pos < text.size()) {
      size_t end = pos + size < text.size() ? pos + size : text.size();
      while (end > pos + 1 && end < text.size() && text[end - 1] != ' ') end--;  // (cut after a space)
      texts.emplace_back(text, pos, end - pos);
      batch_bytes += end - pos;
      pos = end;
    }
    word_piece::fast::Encoder enc(vocab);
    std::vector<int> reused;
    std::vector<size_t> offsets;
    enc.encodeBatch(texts, reused, offsets);  // warm-up: staging buffers; `reused` keeps its storage from here on
    double fresh_best = 0, reused_best = 0;
    size_t batch_ids = 0;
    for (int r = 0; r < 5; r++) {
      std::vector<int> ids;
      auto b0 = clock::now();
      enc.encodeBatch(texts, ids, offsets);
      double s = std::chrono::duration<double>(clock::now() - b0).count();
      fresh_best = (r == 0 || s < fresh_best) ? s : fresh_best;
      batch_ids = ids.size();
      b0 = clock::now();
      enc.encodeBatch(texts, reused, offsets);
      s = std::chrono::duration<double>(clock::now() - b0).count();
      reused_best = (r == 0 || s < reused_best) ? s : reused_best;
      if (reused != ids) batch_ids = 0;  // (reported as a mismatch by the caller)
    }
    std::cout << ", \"batch\": {\"texts\": " << texts.size() << ", \"bytes\": " << batch_bytes << ", \"n_ids\": " << batch_ids
              << ", \"fresh_vector_best_seconds\": " << fresh_best << ", \"reused_vector_best_seconds\": " << reused_best
              << "}";
  }
  std::cout << "}" << std::endl;
  return 0;
}
