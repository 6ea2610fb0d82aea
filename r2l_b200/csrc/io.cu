// Host-side ray-shard reader: payloads of `.npy` ray shards straight into one (pinned) batch buffer,
// several files at a time on native threads - no Python per file, no intermediate arrays, no GIL.
// Reference path: BlenderDataset_v2.__getitem__ = np.load + torch.Tensor per shard in DataLoader worker processes, then
// collate + pin_memory copies (dataset/load_blender.py:304-318, main.py:795-808).  The file format is numpy's .npy
// (versions 1-3) as written by np.save in utils/create_data.py:866-869: C-ordered little-endian float32 [rows, 9].
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace r2l {
namespace {

// Reads one shard; returns an empty string on success, else the reason.
std::string read_one(const char* path, float* dst, int64_t n_floats) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return std::string("cannot open ") + path;
  struct Closer { int fd; ~Closer() { close(fd); } } closer{fd};
  unsigned char head[12];
  if (pread(fd, head, sizeof(head), 0) != (ssize_t)sizeof(head) || memcmp(head, "\x93NUMPY", 6) != 0)
    return std::string(path) + ": not a .npy file";
  const int major = head[6];
  size_t hlen, hoff;
  if (major == 1) { hlen = head[8] | (head[9] << 8); hoff = 10; }
  else if (major == 2 || major == 3) { hlen = head[8] | (head[9] << 8) | (head[10] << 16) | ((size_t)head[11] << 24); hoff = 12; }
  else return std::string(path) + ": unsupported .npy version";
  if (hlen > 65536) return std::string(path) + ": implausible .npy header";
  std::string header(hlen, '\0');
  if (pread(fd, &header[0], hlen, hoff) != (ssize_t)hlen) return std::string(path) + ": truncated .npy header";
  if (header.find("'descr': '<f4'") == std::string::npos) return std::string(path) + ": dtype is not little-endian float32";
  if (header.find("'fortran_order': False") == std::string::npos) return std::string(path) + ": not C-ordered";
  struct stat st;
  if (fstat(fd, &st) != 0) return std::string(path) + ": fstat failed";
  const int64_t payload = (int64_t)st.st_size - (int64_t)(hoff + hlen);
  if (payload != n_floats * 4) {
    char msg[160];
    snprintf(msg, sizeof(msg), ": holds %lld bytes of data, expected %lld", (long long)payload, (long long)(n_floats * 4));
    return std::string(path) + msg;
  }
  char* out = reinterpret_cast<char*>(dst);
  int64_t done = 0;
  while (done < payload) {
    const ssize_t got = pread(fd, out + done, (size_t)(payload - done), (off_t)(hoff + hlen + done));
    if (got <= 0) return std::string(path) + ": short read";
    done += got;
  }
  return std::string();
}

}  // namespace

// dst[i * floats_per_shard ...] = payload of paths[i].  Returns the first error (empty = success).
std::string read_ray_shards(const char* const* paths, int n_paths, float* dst, int64_t floats_per_shard, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_paths) n_threads = n_paths;
  std::atomic<int> next{0};
  std::vector<std::string> errors((size_t)n_threads);
  auto work = [&](int tid) {
    for (int i = next.fetch_add(1); i < n_paths; i = next.fetch_add(1)) {
      std::string e = read_one(paths[i], dst + (int64_t)i * floats_per_shard, floats_per_shard);
      if (!e.empty() && errors[(size_t)tid].empty()) errors[(size_t)tid] = e;
    }
  };
  std::vector<std::thread> threads;
  for (int t = 1; t < n_threads; ++t) threads.emplace_back(work, t);
  work(0);
  for (auto& t : threads) t.join();
  for (const auto& e : errors)
    if (!e.empty()) return e;
  return std::string();
}

}  // namespace r2l
