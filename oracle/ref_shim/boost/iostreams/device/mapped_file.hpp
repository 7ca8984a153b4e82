// Test infrastructure only (oracle build): a minimal stand-in for
// boost::iostreams::mapped_file so that the UNMODIFIED reference sources under
// /root/reference compile without Boost.  The reference uses exactly three
// members (fast.cpp:161,196; linear.cpp:339,350): the (path, readonly)
// constructor, const_data() and size().  Implemented with open/fstat/mmap.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstddef>
#include <stdexcept>
#include <string>

namespace boost {
namespace iostreams {

class mapped_file {
 public:
  enum mapmode { readonly = 1, readwrite = 2, priv = 4 };

  mapped_file(const std::string &path, mapmode) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("mapped_file: cannot open " + path);
    struct stat st;
    if (::fstat(fd, &st) != 0) {
      ::close(fd);
      throw std::runtime_error("mapped_file: cannot stat " + path);
    }
    size_ = static_cast<size_t>(st.st_size);
    if (size_ > 0) {
      void *p = ::mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
      if (p == MAP_FAILED) {
        ::close(fd);
        throw std::runtime_error("mapped_file: mmap failed for " + path);
      }
      data_ = static_cast<const char *>(p);
    }
    ::close(fd);
  }
  mapped_file(const mapped_file &) = delete;
  mapped_file &operator=(const mapped_file &) = delete;
  ~mapped_file() {
    if (data_ != nullptr) ::munmap(const_cast<char *>(data_), size_);
  }

  const char *const_data() const { return data_; }
  size_t size() const { return size_; }

 private:
  const char *data_ = nullptr;
  size_t size_ = 0;
};

}  // namespace iostreams
}  // namespace boost
