#include "CsvIngest.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <climits>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <thread>

namespace {

struct Mapped {
  const char* p = nullptr;
  size_t n = 0;
  int fd = -1;
  explicit Mapped(const std::string& name) {
    fd = ::open(name.c_str(), O_RDONLY);
    if (fd < 0) return;  // the reference's ifstream silently yields no lines for an unreadable file
    struct stat st;
    if (::fstat(fd, &st) == 0 && st.st_size > 0) {
      void* m = ::mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m != MAP_FAILED) { p = (const char*)m; n = (size_t)st.st_size; }
    }
  }
  ~Mapped() {
    if (p) ::munmap((void*)p, n);
    if (fd >= 0) ::close(fd);
  }
};

// std::stoi on the token [b, e): optional whitespace, sign, digits; stops at the first other character.
inline int stoi_token(const char* b, const char* e) {
  while (b < e && (*b == ' ' || *b == '\t' || *b == '\r' || *b == '\n' || *b == '\v' || *b == '\f')) b++;
  bool neg = false;
  if (b < e && (*b == '+' || *b == '-')) { neg = *b == '-'; b++; }
  if (b >= e || *b < '0' || *b > '9') throw std::invalid_argument("stoi");
  long long v = 0;
  for (; b < e && *b >= '0' && *b <= '9'; b++) {
    v = v * 10 + (*b - '0');
    if (v > (long long)INT_MAX + 1) throw std::out_of_range("stoi");
  }
  if (neg) v = -v;
  if (v > INT_MAX || v < INT_MIN) throw std::out_of_range("stoi");
  return (int)v;
}

void parse_file(const std::string& name, int offset, int recordSize, int startFrame, int endFrame, Cameras& cam) {
  Mapped file(name);
  const char* p = file.p;
  const char* end = file.p + file.n;
  int n_line = 0, frame = -1;
  std::vector<int> fields;
  while (p && p < end) {
    const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* line_end = eol ? eol : end;
    const char* next = eol ? eol + 1 : end;
    if (offset > n_line++) { p = next; continue; }
    fields.clear();
    for (const char* t = p; t <= line_end;) {
      const char* comma = (const char*)memchr(t, ',', (size_t)(line_end - t));
      const char* te = comma ? comma : line_end;
      if (te > t) fields.push_back(stoi_token(t, te));
      if (!comma) break;
      t = comma + 1;
    }
    p = next;
    if (fields.empty()) throw std::runtime_error("Invalid CSV file!");
    if ((fields[0] <= startFrame || fields[0] > endFrame) && startFrame != endFrame) { frame = fields[0]; continue; }
    for (int i = 0; i < fields[0] - frame - 1; i++) cam.emplace_back();
    frame = fields[0];
    if ((fields.size() - 1) % (size_t)recordSize != 0) throw std::runtime_error("Invalid CSV file!");
    cam.emplace_back();
    Detections& dets = cam.back();
    const size_t n = fields.size() / (size_t)recordSize;
    dets.reserve(n);
    for (size_t j = 0; j < n; j++) dets.emplace_back((double)fields[j * recordSize + 5], (double)fields[j * recordSize + 6]);
  }
}

}  // namespace

void ingestDetectionFiles(const std::vector<std::string>& files, int offset, int recordSize, int startFrame, int endFrame,
                          std::vector<Cameras>& out, int n_threads) {
  out.assign(files.size(), Cameras());
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)files.size()));
  std::atomic<size_t> next{0};
  std::vector<std::exception_ptr> errors(files.size());
  auto work = [&]() {
    for (size_t i = next++; i < files.size(); i = next++) {
      try {
        parse_file(files[i], offset, recordSize, startFrame, endFrame, out[i]);
      } catch (...) {
        errors[i] = std::current_exception();
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; t++) pool.emplace_back(work);
  work();
  for (std::thread& t : pool) t.join();
  for (const std::exception_ptr& e : errors)
    if (e) std::rethrow_exception(e);  // the first failing file in camera order, like the sequential reader
}
