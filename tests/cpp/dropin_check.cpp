// Drop-in check: drives the C++ entry points of include/word_piece.hpp the way the reference's own test
// harness does (reference tests/tests.cpp:80-97 `check`: fast::encode(text, vocab_vector) compared with
// the expected ids; :259-272 stress shape: one long space-free word), linked against libwordpiece_b200.so.
// Cases come from a file written by tests/test_gpu_parity.py (golden vectors of tests/tests.cpp:137-217
// restated in tests/cases.py, plus seeded texts with ids from the CPU oracle), so this binary holds no
// reference code and reads nothing outside the repo.
//
//   dropin_check <case_file> <scratch_dir>
//
// case file:  N, then per case:  text (len \n bytes \n), vocab (count \n, per token len \n bytes \n),
//             expected ids (count \n ids...\n), expected decode of those ids (count \n, per token len \n bytes \n;
//             count -1 = skip the decode check, e.g. a token holds a newline)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "src/utils.hpp"
#include "src/word_piece.hpp"

namespace {

std::string read_blob(std::istream &in) {
  size_t len = 0;
  in >> len;
  in.get();  // the newline after the length
  std::string s(len, '\0');
  if (len) in.read(&s[0], static_cast<std::streamsize>(len));
  in.get();
  return s;
}

int failures = 0, checks = 0;

void expect(bool ok, const std::string &what) {
  checks++;
  if (!ok) {
    failures++;
    std::cerr << "FAIL: " << what << std::endl;
  }
}

}  // namespace

int main(int argc, char **argv) {
  if (argc != 3) {
    std::cerr << "usage: dropin_check <case_file> <scratch_dir>" << std::endl;
    return 2;
  }
  std::ifstream in(argv[1], std::ios::binary);
  if (!in) {
    std::cerr << "cannot open " << argv[1] << std::endl;
    return 2;
  }
  const std::string dir = argv[2];
  size_t n_cases = 0;
  in >> n_cases;
  double encode_seconds = 0;
  size_t encode_bytes = 0;
  // cases grouped by vocabulary for the Encoder class (a vocabulary that stays on the GPU; batch call)
  std::map<std::vector<std::string>, std::vector<std::pair<std::string, std::vector<int>>>> by_vocab;
  for (size_t c = 0; c < n_cases; c++) {
    const std::string text = read_blob(in);
    size_t n_vocab = 0;
    in >> n_vocab;
    std::vector<std::string> vocab(n_vocab);
    bool newline_in_token = false;
    for (auto &t : vocab) {
      t = read_blob(in);
      if (t.find('\n') != std::string::npos) newline_in_token = true;
    }
    size_t n_ids = 0;
    in >> n_ids;
    std::vector<int> expected(n_ids);
    for (auto &id : expected) in >> id;
    long long n_dec = 0;
    in >> n_dec;
    std::vector<std::string> expected_dec;
    for (long long i = 0; i < n_dec; i++) expected_dec.push_back(read_blob(in));

    // (1) the in-memory overload, fast.cpp:154-157 — what tests/tests.cpp:86,95 call
    const auto t0 = std::chrono::steady_clock::now();
    const std::vector<int> got = word_piece::fast::encode(text, vocab);
    encode_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    encode_bytes += text.size();
    expect(got == expected, "case " + std::to_string(c) + ": fast::encode(text, vocab) ids differ (" +
                                std::to_string(got.size()) + " vs " + std::to_string(expected.size()) + " expected)");
    by_vocab[vocab].emplace_back(text, expected);
    if (newline_in_token) continue;

    // (2) the file overload, fast.cpp:159-163, and decode, fast.cpp:165-187
    const std::string vocab_file = dir + "/vocab_" + std::to_string(c) + ".txt";
    const std::string text_file = dir + "/text_" + std::to_string(c) + ".txt";
    {
      std::ofstream vf(vocab_file, std::ios::binary);
      for (const auto &t : vocab) vf << t << '\n';
      std::ofstream tf(text_file, std::ios::binary);
      tf << text;
    }
    if (!text.empty()) {  // (mapping an empty file is an error in the reference's Boost mapping as well)
      const std::vector<int> got_files = word_piece::fast::encode(text_file, vocab_file);
      expect(got_files == expected, "case " + std::to_string(c) + ": fast::encode(text_file, vocab_file) ids differ");
    }
    if (n_dec >= 0) {
      const std::vector<std::string> dec = word_piece::fast::decode(vocab_file, got);
      expect(dec == expected_dec, "case " + std::to_string(c) + ": fast::decode differs (" + std::to_string(dec.size()) +
                                      " vs " + std::to_string(expected_dec.size()) + " tokens)");
    }
    std::remove(vocab_file.c_str());
    std::remove(text_file.c_str());
  }
  // (3) Encoder (extension): encode == fast::encode, encodeBatch == one encode per text, decode round trip
  for (const auto &group : by_vocab) {
    word_piece::fast::Encoder enc(group.first);
    std::vector<std::string> texts;
    for (const auto &item : group.second) {
      texts.push_back(item.first);
      expect(enc.encode(item.first) == item.second, "Encoder::encode differs from the expected ids");
    }
    const std::vector<std::vector<int>> batch = enc.encodeBatch(texts);
    expect(batch.size() == texts.size(), "Encoder::encodeBatch: wrong number of results");
    for (size_t i = 0; i < batch.size() && i < texts.size(); i++)
      expect(batch[i] == group.second[i].second, "Encoder::encodeBatch: text " + std::to_string(i) + " differs");
    std::vector<int> ids;
    std::vector<size_t> offsets;
    enc.encodeBatch(texts, ids, offsets);
    expect(offsets.size() == texts.size() + 1 && offsets.front() == 0 && offsets.back() == ids.size(),
           "Encoder::encodeBatch: offsets are not a partition of the ids");
    word_piece::fast::Encoder moved = std::move(enc);
    expect(moved.vocabSize() == group.first.size(), "Encoder: vocabulary size after a move");
    expect(moved.encode(texts.front()) == group.second.front().second, "Encoder: encode after a move");
  }
  std::cout << "Passed " << (checks - failures) << " of " << checks << " checks; in-memory encode: " << encode_bytes
            << " bytes in " << encode_seconds << " s" << std::endl;
  return failures ? 1 : 0;
}
