// io.cu — descriptor / match file IO feeding the matcher.
//
// Once matching takes milliseconds, Regions_Provider::load (sfm_regions_provider.hpp:88-138: one cereal/ifstream
// per image into std::vector<Descriptor>) and matching::Save are the wall clock.  Two host-side entry points:
//   * omvg_match_load_desc_files — reads openMVG ".desc" files (features/descriptor.hpp:182-203: a size_t count,
//     then count x 128 bytes) with a pool of host threads straight into ONE page-locked staging buffer laid out
//     like the device arena, and uploads each image as soon as its read completes;
//   * omvg_matches_save — writes a CSR result as "matches.putative.{txt,bin}" in the two formats of
//     matching::Save (matching/indMatch_utils.cpp:80-131): text "I J\nN\n" + "i j\n" lines, or cereal's
//     PortableBinaryOutputArchive of std::map<Pair, std::vector<IndMatch>> (1 endianness byte, u64 map size, per
//     entry u32 I, u32 J, u64 n, n x {u32 i, u32 j}); map order = ascending (I, J), empty pairs are not written.
// Both formats are pinned byte for byte by fixtures the reference's own writers produced (tests/golden/).
#include "common.cuh"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

using namespace omvg;

namespace {

// header count, checked against the file size (a corrupt header must not size a page-locked allocation)
bool read_count(const char *path, uint64_t &n) {
  FILE *f = std::fopen(path, "rb");
  if (!f) return false;
  bool ok = std::fread(&n, sizeof(uint64_t), 1, f) == 1;                // std::size_t on the platforms openMVG builds on
  if (ok && std::fseek(f, 0, SEEK_END) == 0) { const long long sz = ftello(f); ok = n < (1ull << 40) && sz == (long long)(8 + n * OMVG_DESC_LEN); } else ok = false;
  std::fclose(f);
  return ok;
}

}  // namespace

extern "C" {

int omvg_match_load_desc_files(omvg_match_ctx *ctx, uint32_t n_images, const char *const *desc_paths, uint32_t *counts_out) {
  if (!ctx || (!desc_paths && n_images)) return fail(OMVG_E_ARG, "bad arguments");
  std::vector<uint32_t> counts(n_images, 0);
  for (uint32_t k = 0; k < n_images; ++k) {
    uint64_t n = 0;
    if (!desc_paths[k] || !read_count(desc_paths[k], n)) return fail(OMVG_E_ARG, "cannot read descriptor file %u (missing, or its size is not 8 + count x 128 bytes): %s", k, desc_paths[k] ? desc_paths[k] : "(null)");
    if (n >= (1ull << 31)) return fail(OMVG_E_ARG, "descriptor file %u claims %llu descriptors", k, (unsigned long long)n);
    counts[k] = (uint32_t)n;
  }
  int rc = omvg_match_set_images(ctx, n_images, counts.data());
  if (rc) return rc;
  std::vector<size_t> off(n_images + 1, 0);
  for (uint32_t k = 0; k < n_images; ++k) off[k + 1] = off[k] + (size_t)counts[k] * OMVG_DESC_LEN;
  uint8_t *stage = nullptr;
  if (off[n_images]) OMVG_CUDA(cudaMallocHost(&stage, off[n_images]));
  // reader pool: files are claimed in order; every image is uploaded by the thread that read it (the async copy is
  // queued on the context's stream from page-locked memory, so reads and H2D overlap)
  std::atomic<uint32_t> next{0}; std::atomic<int> err{OMVG_OK}; std::atomic<uint32_t> bad{0};
  std::string worker_msg;                                     // omvg_last_error() is thread-local: carried over by hand
  const unsigned nthreads = std::max(1u, std::min<unsigned>(std::min<unsigned>(16u, std::thread::hardware_concurrency()), n_images));
  std::vector<std::thread> pool;
  std::mutex up;
  for (unsigned t = 0; t < nthreads; ++t)
    pool.emplace_back([&]() {
      for (;;) {
        const uint32_t k = next.fetch_add(1);
        if (k >= n_images || err.load() != OMVG_OK) return;
        if (!counts[k]) continue;
        FILE *f = std::fopen(desc_paths[k], "rb");
        const size_t want = (size_t)counts[k] * OMVG_DESC_LEN;
        bool ok = f && std::fseek(f, sizeof(uint64_t), SEEK_SET) == 0 && std::fread(stage + off[k], 1, want, f) == want;
        if (f) std::fclose(f);
        if (!ok) { bad = k; err = OMVG_E_ARG; return; }
        std::lock_guard<std::mutex> g(up);                      // the C ABI of one context is not re-entrant
        const int r = omvg_match_upload_host(ctx, k, stage + off[k]);
        if (r) { worker_msg = omvg_last_error(); bad = k; err = r; return; }      // (still under the lock)
      }
    });
  for (auto &th : pool) th.join();
  if (err.load() == OMVG_OK) rc = omvg_match_prepare(ctx);
  if (rc == OMVG_OK && err.load() == OMVG_OK) rc = omvg_match_sync(ctx);   // staging may only be released after the copies
  else omvg_match_sync(ctx);
  if (stage) cudaFreeHost(stage);
  if (err.load() == OMVG_E_ARG) return fail(OMVG_E_ARG, "short read in descriptor file %u: %s", bad.load(), desc_paths[bad.load()]);
  if (err.load() != OMVG_OK) return fail(err.load(), "upload of descriptor file %u failed: %s", bad.load(), worker_msg.c_str());
  if (rc) return rc;
  if (counts_out) std::memcpy(counts_out, counts.data(), n_images * sizeof(uint32_t));
  return OMVG_OK;
}

int omvg_matches_save(const char *path, uint64_t n_pairs, const uint32_t *pair_I, const uint32_t *pair_J, const uint64_t *offsets, const uint32_t *ij) {
  if (!path || ((!pair_I || !pair_J || !offsets) && n_pairs)) return fail(OMVG_E_ARG, "bad arguments");
  const std::string p(path);
  const size_t dot = p.find_last_of('.');
  const std::string ext = dot == std::string::npos ? "" : p.substr(dot + 1);
  if (ext != "txt" && ext != "bin") return fail(OMVG_E_ARG, "unknown PairWiseMatches file extension: %s", path);
  // std::map<Pair, IndMatches> order; pairs without matches are never in the reference's map
  std::vector<uint64_t> order;
  for (uint64_t q = 0; q < n_pairs; ++q) if (offsets[q + 1] > offsets[q]) order.push_back(q);
  std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return pair_I[a] != pair_I[b] ? pair_I[a] < pair_I[b] : pair_J[a] < pair_J[b]; });
  for (size_t t = 1; t < order.size(); ++t)
    if (pair_I[order[t]] == pair_I[order[t - 1]] && pair_J[order[t]] == pair_J[order[t - 1]]) return fail(OMVG_E_ARG, "pair (%u,%u) appears twice", pair_I[order[t]], pair_J[order[t]]);
  FILE *f = std::fopen(path, ext == "bin" ? "wb" : "w");
  if (!f) return fail(OMVG_E_ARG, "cannot open %s for writing", path);
  bool ok = true;
  if (ext == "txt") {
    std::string buf; buf.reserve(1 << 20);
    char line[64];
    for (const uint64_t q : order) {
      const uint64_t b = offsets[q], e = offsets[q + 1];
      std::snprintf(line, sizeof line, "%u %u\n%llu\n", pair_I[q], pair_J[q], (unsigned long long)(e - b)); buf += line;
      for (uint64_t k = b; k < e; ++k) { std::snprintf(line, sizeof line, "%u %u\n", ij[2 * k], ij[2 * k + 1]); buf += line; }
      if (buf.size() > (1u << 20)) { ok = ok && std::fwrite(buf.data(), 1, buf.size(), f) == buf.size(); buf.clear(); }
    }
    ok = ok && std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
  } else {
    std::vector<uint8_t> buf;
    auto put = [&](const void *src, size_t n) { const uint8_t *s = static_cast<const uint8_t *>(src); buf.insert(buf.end(), s, s + n); };
    const uint8_t little = 1; put(&little, 1);                    // PortableBinaryOutputArchive: endianness of the writer
    const uint64_t n_map = order.size(); put(&n_map, 8);
    for (const uint64_t q : order) {
      const uint64_t b = offsets[q], e = offsets[q + 1], n = e - b;
      put(&pair_I[q], 4); put(&pair_J[q], 4); put(&n, 8);
      put(ij + 2 * b, (size_t)n * 8);                             // {u32 i_, u32 j_} per match, as IndMatch::serialize writes them
      if (buf.size() > (1u << 22)) { ok = ok && std::fwrite(buf.data(), 1, buf.size(), f) == buf.size(); buf.clear(); }
    }
    ok = ok && std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
  }
  ok = std::fclose(f) == 0 && ok;
  return ok ? OMVG_OK : fail(OMVG_E_ARG, "short write to %s", path);
}

}  // extern "C"
