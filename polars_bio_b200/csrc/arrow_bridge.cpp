// arrow_bridge.cpp -- Arrow C Stream level of the C ABI: pbgpu_range_op (include/pbgpu.h).
//
// This is what the reference's PyO3 entry points range_operation_frame / _lazy
// (/root/reference/src/lib.rs:79-88,154-166) and plan builder do_range_operation
// (/root/reference/src/operation.rs:27-98) do around the providers, restated for the GPU engine:
//   drain both ArrowArrayStreams (the callee moves them: released before returning, lib.rs:63-72)
//   -> dictionary-encode the contig columns over a shared dictionary, narrow positions to int32
//      (reference domain limit, docs/features/operations.md:37) into cached pinned staging
//   -> H2D, pbgpu_index_build over the indexed side, provider kernel, D2H of row-id results
//   -> an output ArrowArrayStream that materialises the reference's column contract batch by
//      batch (operation.rs:170-195 nearest, :272-299 overlap, :347 count/coverage) or hands out
//      the raw index pairs (emit = 1).
// Host-side only; every device computation goes through the device-level C ABI in pbgpu.cu.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <future>
#include <mutex>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/pbgpu.h"
#include "arrow_c_abi.h"

struct pbgpu_index;
namespace pbgpu {
int set_error(int code, const char *fmt, ...);
extern thread_local char g_err[512];
int widen_codes_u8(const uint8_t *d_in, int64_t n, int32_t n_contigs, int32_t *d_out, void *stream);
int gather_i32_u8(const int32_t *d_src, const uint32_t *d_rows, int64_t n, uint8_t *d_out, void *stream);
int count_overlaps_u32(const pbgpu_index *ix, const int32_t *c, const int32_t *s, const int32_t *e, int64_t n, int filter_op,
                       uint32_t *d_counts, void *stream);  // pbgpu.cu (internal)
int gather_fixed(const void *d_src, int width, const uint32_t *d_rows, int64_t n, void *d_out, void *stream);
int gather_valid_bits(const uint8_t *d_valid, const uint32_t *d_rows, int64_t n, uint32_t *d_bits, unsigned long long *d_null_count, void *stream);
int gather_str_plan(const long long *d_off, const uint8_t *d_valid, const uint32_t *d_rows, int64_t n, unsigned long long *d_pos,
                    unsigned long long *d_total, void *stream);
int gather_str_bytes(const long long *d_off, const char *d_chars, const uint32_t *d_rows, int64_t n, const unsigned long long *d_pos,
                     const unsigned long long *d_total, void *d_out_off, int large, char *d_out_chars, void *stream);
int rebase_offsets(const void *d_in, int large, int64_t len, long long base, long long *d_out, void *stream);
void release_thread_state();  // pbgpu.cu: host mailbox + stage events of the calling (helper) thread
int dev_alloc(void **p, size_t bytes, cudaStream_t s);  // stream-ordered device blocks through pbgpu.cu's block cache
void dev_free(void *p, cudaStream_t s);
}  // namespace pbgpu
using pbgpu::set_error;

namespace {

// PBGPU_TRACE=1: host-stage wall times of every pbgpu_range_op call on stderr (tuning aid)
struct Trace {
  bool on;
  std::chrono::steady_clock::time_point t0;
  Trace() {
    const char *e = getenv("PBGPU_TRACE");
    on = e && (e[0] == '1' || e[0] == '2');
    t0 = std::chrono::steady_clock::now();
  }
  // PBGPU_TRACE=2: additionally drain `s` first, so the lap holds the device work queued so far (serialises the pipeline)
  void lap_drained(const char *what, cudaStream_t s) {
    const char *e = getenv("PBGPU_TRACE");
    if (!(e && e[0] == '2')) return;
    on = true;
    cudaStreamSynchronize(s);
    lap(what);
  }
  void lap(const char *what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[pbgpu] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// ---------------------------------------------------------------------------------------------
// tiny persistent thread pool: parallel_for over [0,n) in chunks
class Pool {
 public:
  static Pool &get() {
    static Pool p;
    return p;
  }
  int size() const { return (int)workers_.size() + 1; }
  // Returns false when a task threw (std::bad_alloc in practice): nothing may unwind out of a worker thread or
  // across the C ABI, so the failure is latched and reported to the caller instead.
  bool parallel_for(int64_t n_tasks, const std::function<void(int64_t)> &fn) {
    if (n_tasks <= 0) return true;
    if (n_tasks == 1 || workers_.empty()) {
      try { for (int64_t i = 0; i < n_tasks; ++i) fn(i); } catch (...) { return false; }
      return true;
    }
    std::unique_lock<std::mutex> run_lock(run_mu_);  // one parallel_for at a time
    failed_.store(false);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      next_.store(0);
      total_ = n_tasks;
      pending_ = (int)workers_.size();
      ++epoch_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
    return !failed_.load();
  }

 private:
  Pool() {
    unsigned hc = std::thread::hardware_concurrency();
    int n = (int)std::min<unsigned>(hc ? hc : 4, 64);
    const char *e = getenv("PBGPU_HOST_THREADS");
    if (e && atoi(e) > 0) n = atoi(e);
    for (int i = 1; i < n; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
  }
  void work() {
    for (;;) {
      int64_t i = next_.fetch_add(1);
      if (i >= total_) break;
      try { (*fn_)(i); } catch (...) { failed_.store(true); }
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) return;
      }
      work();
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_cv_.notify_all();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int64_t)> *fn_ = nullptr;
  std::atomic<int64_t> next_{0};
  std::atomic<bool> failed_{false};
  int64_t total_ = 0;
  int pending_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

constexpr int64_t kChunk = 1 << 14;        // rows per key-encoding task (small: a 1M-row H2D slice still feeds 64 threads)
constexpr int64_t kGatherChunk = 1 << 14;  // rows per gather task (multiple of 8: validity bytes stay task-private)

// ---------------------------------------------------------------------------------------------
// cached pinned staging (cudaHostAlloc is milliseconds per 100 MB: never on the per-call path twice)
struct PinnedBuf {
  void *p = nullptr;
  size_t cap = 0;
  bool busy = false;
  bool wc = false;  // write-combined: H2D staging only (CPU writes, never reads)
};
std::mutex g_pin_mu;
std::vector<PinnedBuf> g_pin;
size_t g_pin_busy = 0, g_pin_peak = 0;  // page-locked staging bytes in use right now / high-water mark (pbgpu_pinned_stats)

void *pinned_get(size_t bytes, bool wc = false) {
  if (bytes == 0) bytes = 1;
  std::lock_guard<std::mutex> lk(g_pin_mu);
  int best = -1;
  for (int i = 0; i < (int)g_pin.size(); ++i)
    if (!g_pin[i].busy && g_pin[i].wc == wc && g_pin[i].cap >= bytes && (best < 0 || g_pin[i].cap < g_pin[best].cap)) best = i;
  if (best >= 0) {
    g_pin[best].busy = true;
    g_pin_busy += g_pin[best].cap;
    if (g_pin_busy > g_pin_peak) g_pin_peak = g_pin_busy;
    return g_pin[best].p;
  }
  PinnedBuf b;
  size_t cap = (bytes + (1 << 20) - 1) & ~(size_t)((1 << 20) - 1);
  if (cudaHostAlloc(&b.p, cap, cudaHostAllocPortable | (wc ? cudaHostAllocWriteCombined : 0)) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  b.cap = cap;
  b.busy = true;
  b.wc = wc;
  g_pin.push_back(b);
  g_pin_busy += cap;
  if (g_pin_busy > g_pin_peak) g_pin_peak = g_pin_busy;
  return b.p;
}
size_t pinned_cap_bytes() {
  static size_t cap = [] { const char *e = getenv("PBGPU_PINNED_CACHE_MB"); return (size_t)(e ? atoll(e) : 8192) << 20; }();
  return cap;
}
void pinned_put(void *p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_pin_mu);
  size_t total = 0;
  for (auto &b : g_pin) { if (b.p == p && b.busy) { b.busy = false; g_pin_busy -= b.cap; } total += b.cap; }
  // keep the cache bounded (PBGPU_PINNED_CACHE_MB, default 8 GiB): release idle buffers, largest first
  while (total > pinned_cap_bytes()) {
    int victim = -1;
    for (int i = 0; i < (int)g_pin.size(); ++i)
      if (!g_pin[i].busy && (victim < 0 || g_pin[i].cap > g_pin[victim].cap)) victim = i;
    if (victim < 0) break;
    cudaFreeHost(g_pin[victim].p);
    total -= g_pin[victim].cap;
    g_pin.erase(g_pin.begin() + victim);
  }
}
struct PinnedHold {  // returns its buffers to the cache on destruction
  std::vector<void *> v;
  bool wc = false;
  template <typename T>
  T *get(size_t count) {
    void *p = pinned_get(count * sizeof(T), wc);
    if (p) v.push_back(p);
    return (T *)p;
  }
  ~PinnedHold() {
    for (void *p : v) pinned_put(p);
  }
};

// ---------------------------------------------------------------------------------------------
// Host buffer cache for everything handed to the consumer (result columns, gathered payload).  Large result
// buffers are otherwise mmap'ed, page-faulted on first touch and munmap'ed on release by glibc on every call
// (tens of ms per 100 MB): the same allocator sensitivity the reference notes for mimalloc/jemalloc
// (Cargo.toml:9-11).  Freed blocks >= 64 KiB are kept (up to PBGPU_HOST_CACHE_MB, default 4096) and reused.
struct HostCache {
  std::mutex mu;
  std::unordered_map<size_t, std::vector<void *>> free_;
  size_t cached = 0, limit = (size_t)4096 << 20;
  HostCache() {
    const char *e = getenv("PBGPU_HOST_CACHE_MB");
    if (e) limit = (size_t)atoll(e) << 20;
  }
  // process exit: the CUDA context may already be gone, so page-locked blocks are left to the OS
  ~HostCache() {
    for (auto &kv : free_) for (void *p : kv.second) if (!((size_t *)p)[1]) free(p);
  }
};
HostCache &host_cache() { static HostCache c; return c; }
constexpr size_t kHdr = 64;  // block header: [0] capacity, [1] 1 = page-locked (cudaHostRegister'ed) for direct D2H landing
constexpr size_t kDmaMin = (size_t)1 << 20;
inline size_t size_class(size_t bytes) {
  size_t need = bytes + kHdr;
  if (need <= ((size_t)64 << 10)) return need;                                  // small: not cached
  if (need <= ((size_t)1 << 20)) { size_t c = (size_t)64 << 10; while (c < need) c <<= 1; return c; }
  return (need + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);           // multiples of 1 MiB
}
void free_block(void *base) {
  if (((size_t *)base)[1]) { if (cudaHostUnregister(base) != cudaSuccess) cudaGetLastError(); }
  free(base);
}
void *hmalloc(size_t bytes) {
  const size_t cap = size_class(bytes ? bytes : 1);
  void *base = nullptr;
  if (cap > ((size_t)64 << 10)) {
    HostCache &c = host_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.free_.find(cap);
    if (it != c.free_.end() && !it->second.empty()) { base = it->second.back(); it->second.pop_back(); c.cached -= cap; }
  }
  if (!base) {
    if (posix_memalign(&base, cap >= kDmaMin ? 4096 : 64, cap) != 0) return nullptr;
    ((size_t *)base)[1] = 0;
  }
  *(size_t *)base = cap;
  return (char *)base + kHdr;
}
// A result buffer the DMA engine can write directly: the block is page-locked once (cudaHostRegister is
// milliseconds per 100 MB) and stays so while it cycles through the cache, so steady-state calls pay nothing and the
// pinned-staging -> host copy of every result column disappears.  *dma = false when registration is not possible
// (small block, PBGPU_DIRECT_D2H=0, or the driver refuses): the caller then stages through a pinned buffer.
void *hmalloc_dma(size_t bytes, bool *dma) {
  static const bool enabled = [] { const char *e = getenv("PBGPU_DIRECT_D2H"); return !(e && e[0] == '0'); }();
  void *p = hmalloc(bytes);
  *dma = false;
  if (!p || !enabled) return p;
  size_t *base = (size_t *)((char *)p - kHdr);
  if (base[0] < kDmaMin) return p;
  if (!base[1]) {
    if (cudaHostRegister(base, base[0], cudaHostRegisterPortable) == cudaSuccess) base[1] = 1;
    else cudaGetLastError();
  }
  *dma = base[1] != 0;
  return p;
}
void *hcalloc(size_t bytes) {
  void *p = hmalloc(bytes);
  if (p) memset(p, 0, bytes);
  return p;
}
void hfree(void *p) {
  if (!p) return;
  void *base = (char *)p - kHdr;
  const size_t cap = *(size_t *)base;
  if (cap > ((size_t)64 << 10)) {
    HostCache &c = host_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.cached + cap <= c.limit) { c.free_[cap].push_back(base); c.cached += cap; return; }
  }
  free_block(base);
}

// Result arrays handed to the consumer live in cached host blocks (they may outlive the call by a long time).
// Large ones are page-locked in place (hmalloc_dma) and the D2H lands in them directly; otherwise the D2H lands in a
// cached pinned staging buffer and the pool copies it out in parallel (copy_out).
struct HostBufs {
  std::vector<void *> v;
  ~HostBufs() { for (void *p : v) hfree(p); }
};
void copy_into(void *dst, const void *pinned, size_t bytes) {
  const size_t chunk = (size_t)4 << 20;
  const int64_t nt = (int64_t)((bytes + chunk - 1) / chunk);
  Pool::get().parallel_for(nt, [&](int64_t i) {
    const size_t off = (size_t)i * chunk;
    memcpy((char *)dst + off, (const char *)pinned + off, std::min(chunk, bytes - off));
  });
}
// ---------------------------------------------------------------------------------------------
inline bool bit_get(const uint8_t *bits, int64_t i) { return (bits[i >> 3] >> (i & 7)) & 1; }
inline void bit_set(uint8_t *bits, int64_t i) { bits[i >> 3] |= (uint8_t)(1u << (i & 7)); }

size_t metadata_size(const char *md) {
  if (!md) return 0;
  const char *p = md;
  int32_t n;
  memcpy(&n, p, 4);
  p += 4;
  for (int32_t i = 0; i < n; ++i) {
    int32_t l;
    memcpy(&l, p, 4); p += 4 + l;
    memcpy(&l, p, 4); p += 4 + l;
  }
  return (size_t)(p - md);
}

// An input table: the drained stream (schema + batches), owned.
struct Table {
  ArrowSchema schema{};
  std::vector<ArrowArray> batches;
  std::vector<int64_t> start;  // prefix sums of batch lengths, size nb+1
  int64_t rows = 0;
  int key[3] = {-1, -1, -1};
  Table() = default;
  Table(const Table &) = delete;
  Table &operator=(const Table &) = delete;
  ~Table() {
    for (auto &b : batches)
      if (b.release) b.release(&b);
    if (schema.release) schema.release(&schema);
  }
  int64_t n_cols() const { return schema.n_children; }
  // (batch, index inside batch) of a global row
  inline void locate(uint32_t row, int &b, int64_t &i) const {
    if (batches.size() == 1) { b = 0; i = row; return; }
    auto it = std::upper_bound(start.begin(), start.end(), (int64_t)row);
    b = (int)(it - start.begin()) - 1;
    i = (int64_t)row - start[b];
  }
};

int drain(ArrowArrayStream *s, Table &t, const char *side) {
  if (!s || !s->get_schema || !s->get_next) return set_error(PBGPU_EINVAL, "%s stream is NULL/invalid", side);
  if (s->get_schema(s, &t.schema) != 0) {
    const char *m = s->get_last_error ? s->get_last_error(s) : nullptr;
    return set_error(PBGPU_ESTREAM, "%s stream get_schema failed: %s", side, m ? m : "?");
  }
  if (!t.schema.format || strcmp(t.schema.format, "+s") != 0)
    return set_error(PBGPU_ESCHEMA, "%s stream must yield struct (record batch) arrays, got format '%s'", side,
                     t.schema.format ? t.schema.format : "?");
  t.start.push_back(0);
  for (;;) {
    ArrowArray a{};
    if (s->get_next(s, &a) != 0) {
      const char *m = s->get_last_error ? s->get_last_error(s) : nullptr;
      return set_error(PBGPU_ESTREAM, "%s stream get_next failed: %s", side, m ? m : "?");
    }
    if (!a.release) break;  // end of stream
    if (a.n_children != t.schema.n_children) {
      a.release(&a);
      return set_error(PBGPU_ESCHEMA, "%s stream: batch has %lld children, schema %lld", side, (long long)a.n_children,
                       (long long)t.schema.n_children);
    }
    if (a.length == 0) { a.release(&a); continue; }
    t.rows += a.length;
    t.start.push_back(t.rows);
    t.batches.push_back(a);
  }
  if (t.rows >= 0xFFFFFFFFll) return set_error(PBGPU_ERANGE, "%s table has %lld rows; limit is 2^32-2", side, (long long)t.rows);
  return PBGPU_OK;
}

int find_col(const Table &t, const char *name, const char *side, int *out) {
  if (!name) return set_error(PBGPU_EINVAL, "NULL column name for %s", side);
  for (int64_t i = 0; i < t.schema.n_children; ++i)
    if (t.schema.children[i]->name && strcmp(t.schema.children[i]->name, name) == 0) { *out = (int)i; return PBGPU_OK; }
  return set_error(PBGPU_ESCHEMA, "column '%s' not found in %s table", name, side);
}

// ---- contig dictionary shared by both sides ----------------------------------------------------
struct ContigDict {
  std::mutex mu;
  std::unordered_map<std::string, int32_t> map;
  int32_t intern(std::string_view sv) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = map.find(std::string(sv));
    if (it != map.end()) return it->second;
    int32_t c = (int32_t)map.size();
    map.emplace(std::string(sv), c);
    return c;
  }
};

enum class StrKind { Utf8, LargeUtf8, View, None };
StrKind str_kind(const char *f) {
  if (!strcmp(f, "u") || !strcmp(f, "z")) return StrKind::Utf8;
  if (!strcmp(f, "U") || !strcmp(f, "Z")) return StrKind::LargeUtf8;
  if (!strcmp(f, "vu") || !strcmp(f, "vz")) return StrKind::View;
  return StrKind::None;
}

// string at physical index j of a (non-dictionary) string array
inline std::string_view str_at(const ArrowArray *a, StrKind k, int64_t j) {
  switch (k) {
    case StrKind::Utf8: {
      const int32_t *o = (const int32_t *)a->buffers[1];
      return std::string_view((const char *)a->buffers[2] + o[j], (size_t)(o[j + 1] - o[j]));
    }
    case StrKind::LargeUtf8: {
      const int64_t *o = (const int64_t *)a->buffers[1];
      return std::string_view((const char *)a->buffers[2] + o[j], (size_t)(o[j + 1] - o[j]));
    }
    case StrKind::View: {
      const uint8_t *v = (const uint8_t *)a->buffers[1] + 16 * j;
      int32_t len;
      memcpy(&len, v, 4);
      if (len <= 12) return std::string_view((const char *)v + 4, (size_t)len);
      int32_t bi, off;
      memcpy(&bi, v + 8, 4);
      memcpy(&off, v + 12, 4);
      return std::string_view((const char *)a->buffers[2 + bi] + off, (size_t)len);
    }
    default: return {};
  }
}

inline int64_t int_at(const void *buf, char f, int64_t j) {
  switch (f) {
    case 'c': return ((const int8_t *)buf)[j];
    case 'C': return ((const uint8_t *)buf)[j];
    case 's': return ((const int16_t *)buf)[j];
    case 'S': return ((const uint16_t *)buf)[j];
    case 'i': return ((const int32_t *)buf)[j];
    case 'I': return ((const uint32_t *)buf)[j];
    case 'l': return ((const int64_t *)buf)[j];
    case 'L': { uint64_t v = ((const uint64_t *)buf)[j]; return v > (uint64_t)INT64_MAX ? INT64_MAX : (int64_t)v; }
    default: return 0;
  }
}
inline bool is_int_format(const char *f) { return f && f[0] && !f[1] && strchr("cCsSiIlL", f[0]); }

// Encode the three key columns of a table into int32 staging (code = -1 for null keys).
// code8 != NULL: contig codes are written as one byte each instead (255 = null key or code >= 255): a quarter of the
// H2D bytes of that column when the index holds at most 255 contigs (codes it does not know cannot match anyway).
// ---- streaming stores into the H2D staging buffers -------------------------------------------------------------
// The staging buffers are written by pool threads and then read once by the copy engine.  Measured on the B200 boxes
// (scripts/h2d_probe.cu): a 12 MB pinned buffer just written with ordinary stores by other threads is DMA-read at
// 5 GB/s (the lines sit dirty in those cores' caches), the same buffer written with non-temporal stores at 51 GB/s --
// 1.4 ms per call for the indexed side alone; 96 MB: 26 vs 50 GB/s.  (cudaHostAllocWriteCombined made no difference
// there.)  So everything that lands in staging goes through these, and every task ends with an sfence.
#if defined(__SSE2__)
#include <emmintrin.h>
inline void nt_copy(void *dst, const void *src, size_t bytes) {
  char *d = (char *)dst;
  const char *s = (const char *)src;
  size_t head = (16 - ((uintptr_t)d & 15)) & 15;
  if (head > bytes) head = bytes;
  memcpy(d, s, head);
  d += head; s += head; bytes -= head;
  const size_t blocks = bytes / 16;
  for (size_t i = 0; i < blocks; ++i) _mm_stream_si128((__m128i *)d + i, _mm_loadu_si128((const __m128i *)s + i));
  memcpy(d + blocks * 16, s + blocks * 16, bytes - blocks * 16);
}
inline void nt_fill(void *dst, size_t bytes, __m128i pattern /*period divides 16; dst aligned to the period*/, const void *pat_bytes) {
  char *d = (char *)dst;
  size_t head = (16 - ((uintptr_t)d & 15)) & 15;
  if (head > bytes) head = bytes;
  memcpy(d, (const char *)pat_bytes + (16 - head) % 16, head);  // pattern is periodic with a divisor of 16: any 16-aligned phase
  d += head; bytes -= head;
  const size_t blocks = bytes / 16;
  for (size_t i = 0; i < blocks; ++i) _mm_stream_si128((__m128i *)d + i, pattern);
  memcpy(d + blocks * 16, pat_bytes, bytes - blocks * 16);
}
inline void nt_fill8(uint8_t *dst, uint8_t v, size_t n) {
  alignas(16) uint8_t pat[16];
  memset(pat, v, 16);
  nt_fill(dst, n, _mm_set1_epi8((char)v), pat);
}
inline void nt_fill32(int32_t *dst, int32_t v, size_t n) {  // dst is 4-byte aligned: head/tail are whole elements
  alignas(16) int32_t pat[4] = {v, v, v, v};
  nt_fill(dst, 4 * n, _mm_set1_epi32(v), pat);
}
inline void nt_fence() { _mm_sfence(); }
#else
inline void nt_copy(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
inline void nt_fill8(uint8_t *dst, uint8_t v, size_t n) { memset(dst, v, n); }
inline void nt_fill32(int32_t *dst, int32_t v, size_t n) { for (size_t i = 0; i < n; ++i) dst[i] = v; }
inline void nt_fence() {}
#endif

// Collision-free table for names of at most 7 bytes (key = the bytes as a little-endian word | length << 56; code >= 0).
// slot = (key * mult) >> 54 over 1024 entries; when a new key lands on an occupied slot the multiplier is replaced until
// all keys sit alone (a handful of tries for the few dozen contig names of a genome; beyond kMaxKeys names new ones simply
// stay on the caller's slow path).  A lookup is then one multiply, one load and a compare that virtually always succeeds.
struct ShortNames {
  static constexpr int32_t kMiss = INT32_MIN;
  static constexpr int kBits = 10, kMaxKeys = 56;
  struct Ent { uint64_t key; int32_t code; };
  Ent e[1 << kBits];
  uint64_t mult = 0x9E3779B97F4A7C15ull;
  uint64_t keys[kMaxKeys];
  int32_t codes[kMaxKeys];
  int n = 0;
  ShortNames() { for (auto &x : e) { x.key = ~0ull; x.code = kMiss; } }
  static inline uint64_t mask(size_t len) { return len >= 8 ? ~0ull : ((1ull << (8 * len)) - 1ull); }
  inline unsigned slot(uint64_t key) const { return (unsigned)((key * mult) >> (64 - kBits)); }
  inline int32_t find(uint64_t key) const { const Ent &x = e[slot(key)]; return x.key == key ? x.code : kMiss; }
  void insert(uint64_t key, int32_t code) {
    if (code < 0 || n >= kMaxKeys) return;
    keys[n] = key; codes[n] = code; ++n;
    Ent &x = e[slot(key)];
    if (x.key == ~0ull) { x.key = key; x.code = code; return; }
    for (int tries = 0; tries < 4096; ++tries) {  // occupied: clear what was placed, try another multiplier
      for (int k = 0; k < n; ++k) { Ent &y = e[slot(keys[k])]; y.key = ~0ull; y.code = kMiss; }
      mult = (mult * 6364136223846793005ull + 1442695040888963407ull) | 1ull;
      bool ok = true;
      int placed = 0;
      for (; placed < n && ok; ++placed) {
        Ent &y = e[slot(keys[placed])];
        if (y.key != ~0ull) ok = false; else { y.key = keys[placed]; y.code = codes[placed]; }
      }
      if (ok) return;
      for (int k = 0; k < placed - 1; ++k) { Ent &y = e[slot(keys[k])]; y.key = ~0ull; y.code = kMiss; }  // undo this try
    }
    --n;  // no luck (cannot happen in practice): leave the table without the new key
    for (int k = 0; k < n; ++k) { Ent &y = e[slot(keys[k])]; if (y.key == ~0ull) { y.key = keys[k]; y.code = codes[k]; } }
  }
};

int encode_keys(const Table &t, const char *side, ContigDict &dict, int32_t *code, int32_t *st, int32_t *en,
                int64_t row_lo = 0, int64_t row_hi = INT64_MAX, uint8_t *code8 = nullptr) {
  const ArrowSchema *fc = t.schema.children[t.key[0]];
  const ArrowSchema *fs = t.schema.children[t.key[1]];
  const ArrowSchema *fe = t.schema.children[t.key[2]];
  const bool is_dict = fc->dictionary != nullptr;
  const StrKind sk = str_kind(is_dict ? fc->dictionary->format : fc->format);
  if (sk == StrKind::None || (is_dict && !is_int_format(fc->format)))
    return set_error(PBGPU_ESCHEMA, "%s contig column '%s' has unsupported type '%s' (want utf8 / large_utf8 / utf8_view / dictionary of those)",
                     side, fc->name, fc->format);
  if (!is_int_format(fs->format) || !is_int_format(fe->format))
    return set_error(PBGPU_ESCHEMA, "%s position columns must be integers, got '%s' / '%s'", side, fs->format, fe->format);
  const char sf = fs->format[0], ef = fe->format[0];

  struct Task { int b; int64_t lo, hi; };
  std::vector<Task> tasks;
  for (int b = 0; b < (int)t.batches.size(); ++b) {  // the part of every batch inside [row_lo,row_hi), in kChunk pieces
    const int64_t b_lo = std::max<int64_t>(0, row_lo - t.start[b]), b_hi = std::min<int64_t>(t.batches[b].length, row_hi - t.start[b]);
    for (int64_t lo = b_lo; lo < b_hi; lo += kChunk) tasks.push_back({b, lo, std::min(lo + kChunk, b_hi)});
  }
  std::atomic<int> range_err{0};
  const bool pool_ok = Pool::get().parallel_for((int64_t)tasks.size(), [&](int64_t ti) {
    const Task &tk = tasks[ti];
    const ArrowArray &ba = t.batches[tk.b];
    const ArrowArray *ac = ba.children[t.key[0]], *as = ba.children[t.key[1]], *ae = ba.children[t.key[2]];
    const int64_t oc = ac->offset + ba.offset, os = as->offset + ba.offset, oe = ae->offset + ba.offset;
    const uint8_t *vc = ac->null_count != 0 ? (const uint8_t *)ac->buffers[0] : nullptr;
    const uint8_t *vs = as->null_count != 0 ? (const uint8_t *)as->buffers[0] : nullptr;
    const uint8_t *ve = ae->null_count != 0 ? (const uint8_t *)ae->buffers[0] : nullptr;
    const int64_t g0 = t.start[tk.b];
    std::unordered_map<std::string_view, int32_t> local;  // views into Arrow buffers (alive for the call)
    std::string_view prev;
    int32_t prev_code = -1;
    bool have_prev = false;
    std::vector<int32_t> dict_codes;
    if (is_dict) {  // map this batch's dictionary values once
      const ArrowArray *d = ac->dictionary;
      dict_codes.resize((size_t)d->length);
      const uint8_t *vd = d->null_count != 0 ? (const uint8_t *)d->buffers[0] : nullptr;
      for (int64_t j = 0; j < d->length; ++j)
        dict_codes[j] = (vd && !bit_get(vd, j + d->offset)) ? -1 : dict.intern(str_at(d, sk, j + d->offset));
    }
    // ---- fast paths (the common shape: int32 positions and string contigs in long runs, no nulls) ----
    // positions: a straight copy into the staging buffer
    const bool fast_s = !vs && sf == 'i', fast_e = !ve && ef == 'i';
    if (fast_s) nt_copy(st + g0 + tk.lo, (const int32_t *)as->buffers[1] + os + tk.lo, 4 * (size_t)(tk.hi - tk.lo));
    if (fast_e) nt_copy(en + g0 + tk.lo, (const int32_t *)ae->buffers[1] + oe + tk.lo, 4 * (size_t)(tk.hi - tk.lo));
    // contigs: genomic tables are (mostly) sorted by contig, so a block of rows usually repeats one string.  For
    // utf8 / large_utf8 that is: equal lengths and a data region that is periodic with that length -- two
    // vectorisable sweeps instead of a hash lookup or memcmp per row.
    auto block_is_run = [&](int64_t a, int64_t b) -> bool {  // rows [a,b) of this batch (b - a >= 2), no nulls
      if (sk == StrKind::Utf8) {
        const int32_t *o = (const int32_t *)ac->buffers[1] + oc;
        const int32_t L = o[a + 1] - o[a];
        if (o[a + 2] - o[a + 1] != L || memcmp((const char *)ac->buffers[2] + o[a], (const char *)ac->buffers[2] + o[a + 1], (size_t)L) != 0)
          return false;  // the first two rows differ (mixed contigs): no need to look at the other 1022
        int bad = 0;
        for (int64_t i = a + 1; i < b; ++i) bad |= (o[i + 1] - o[i]) ^ L;
        if (bad) return false;
        const char *d = (const char *)ac->buffers[2] + o[a];
        return L == 0 || memcmp(d, d + L, (size_t)L * (size_t)(b - a - 1)) == 0;
      }
      if (sk == StrKind::LargeUtf8) {
        const int64_t *o = (const int64_t *)ac->buffers[1] + oc;
        const int64_t L = o[a + 1] - o[a];
        if (o[a + 2] - o[a + 1] != L || memcmp((const char *)ac->buffers[2] + o[a], (const char *)ac->buffers[2] + o[a + 1], (size_t)L) != 0)
          return false;
        int64_t bad = 0;
        for (int64_t i = a + 1; i < b; ++i) bad |= (o[i + 1] - o[i]) ^ L;
        if (bad) return false;
        const char *d = (const char *)ac->buffers[2] + o[a];
        return L == 0 || memcmp(d, d + L, (size_t)L * (size_t)(b - a - 1)) == 0;
      }
      return false;
    };
    auto lookup_slow = [&](std::string_view sv) -> int32_t {
      if (have_prev && sv.size() == prev.size() && memcmp(sv.data(), prev.data(), sv.size()) == 0) return prev_code;
      int32_t c;
      auto it = local.find(sv);
      if (it != local.end()) c = it->second;
      else { c = dict.intern(sv); local.emplace(sv, c); }
      prev = sv; prev_code = c; have_prev = true;
      return c;
    };
    // Short names (<= 7 bytes: "chr1" .. "chrUn", "1" .. "MT") in arbitrary order -- rows of all contigs mixed, the shape
    // of BASELINE config 3 -- would pay a std::hash + compare per row (~40 ns); instead the string is read as ONE 64-bit
    // word (length in the top byte) and looked up in a COLLISION-FREE multiplicative hash table (ShortNames below): one
    // multiply, one load, one always-taken compare per row.  (Round 2's first version was a 128-entry direct-mapped table:
    // two of the 24 human contig names shared a slot and evicted each other on every alternation -- 20 % of the rows took
    // the slow path, and the unpredictable hit / miss branch cost more than the lookup: 25 ns per row; now ~4.5.)  The
    // 8-byte read must stay inside the data buffer, so the last few bytes of it take the ordinary path.
    ShortNames small;
    const char *data_end = nullptr;
    if (!is_dict && (sk == StrKind::Utf8 || sk == StrKind::LargeUtf8) && ac->buffers[2]) {
      const int64_t last = ac->offset + ac->length;  // offsets[last] = end of the data this array can reference
      data_end = (const char *)ac->buffers[2] +
                 (sk == StrKind::Utf8 ? (int64_t)((const int32_t *)ac->buffers[1])[last] : ((const int64_t *)ac->buffers[1])[last]);
    }
    auto lookup = [&](std::string_view sv) -> int32_t {
      const size_t L = sv.size();
      if (L <= 7 && data_end && sv.data() + 8 <= data_end) {
        uint64_t w;
        memcpy(&w, sv.data(), 8);
        const uint64_t key = (w & ShortNames::mask(L)) | ((uint64_t)L << 56);
        const int32_t hit = small.find(key);
        if (__builtin_expect(hit != ShortNames::kMiss, 1)) return hit;
        const int32_t c = lookup_slow(sv);
        small.insert(key, c);
        return c;
      }
      return lookup_slow(sv);
    };
    const bool run_ok = !is_dict && !vc && fast_s && fast_e && (sk == StrKind::Utf8 || sk == StrKind::LargeUtf8);
    constexpr int64_t kRun = 1024;
    for (int64_t blk = tk.lo; blk < tk.hi; blk += kRun) {
      const int64_t bhi = std::min(tk.hi, blk + kRun);
      if (run_ok && bhi - blk >= 2 && block_is_run(blk, bhi)) {
        const int32_t c = lookup(str_at(ac, sk, blk + oc));
        if (code8) nt_fill8(code8 + g0 + blk, (uint8_t)((c < 0 || c >= 255) ? 255 : c), (size_t)(bhi - blk));
        else nt_fill32(code + g0 + blk, c, (size_t)(bhi - blk));
        continue;
      }
      // row by row into block-local arrays (they stay in this core's cache), streamed out to staging afterwards
      int32_t tc[kRun], ts[kRun], te[kRun];
      uint8_t tc8[kRun];
      if (run_ok) {  // the common shape (utf8 / large_utf8 contig without nulls, int32 positions already copied): a loop
        auto tight = [&](auto *o) {  // with every pointer hoisted and nothing but the short-name lookup in it
          const char *d = (const char *)ac->buffers[2];
          for (int64_t i = blk; i < bhi; ++i) {
            const int64_t off = (int64_t)o[i];
            const size_t L = (size_t)((int64_t)o[i + 1] - off);
            const char *p = d + off;
            int32_t c;
            if (L <= 7 && p + 8 <= data_end) {
              uint64_t w;
              memcpy(&w, p, 8);
              const uint64_t key = (w & ShortNames::mask(L)) | ((uint64_t)L << 56);
              c = small.find(key);
              if (__builtin_expect(c == ShortNames::kMiss, 0)) { c = lookup_slow(std::string_view(p, L)); small.insert(key, c); }
            } else c = lookup_slow(std::string_view(p, L));
            tc[i - blk] = c;
          }
        };
        if (sk == StrKind::Utf8) tight((const int32_t *)ac->buffers[1] + oc); else tight((const int64_t *)ac->buffers[1] + oc);
        const size_t nb = (size_t)(bhi - blk);
        if (code8) {
          for (size_t k = 0; k < nb; ++k) tc8[k] = (uint8_t)(((uint32_t)tc[k] >= 255u) ? 255 : tc[k]);  // negative or >= 255 -> 255
          nt_copy(code8 + g0 + blk, tc8, nb);
        } else nt_copy(code + g0 + blk, tc, 4 * nb);
        continue;
      }
      for (int64_t i = blk; i < bhi; ++i) {
        const int64_t li = i - blk;
        int32_t c;
        if (vc && !bit_get(vc, i + oc)) c = -1;
        else if (is_dict) {
          int64_t k = int_at(ac->buffers[1], fc->format[0], i + oc);
          c = (k >= 0 && k < (int64_t)dict_codes.size()) ? dict_codes[k] : -1;
        } else c = lookup(str_at(ac, sk, i + oc));
        if (!fast_s || !fast_e) {
          int64_t s = fast_s ? 0 : int_at(as->buffers[1], sf, i + os), e = fast_e ? 0 : int_at(ae->buffers[1], ef, i + oe);
          bool null_pos = (vs && !bit_get(vs, i + os)) || (ve && !bit_get(ve, i + oe));
          if (s < INT32_MIN || s > INT32_MAX || e < INT32_MIN || e > INT32_MAX) { range_err.store(1); null_pos = true; }
          if (null_pos) { c = -1; s = 0; e = 0; }  // null-keyed row: never matches; a fast-copied position may stay as it is
          ts[li] = (int32_t)s;
          te[li] = (int32_t)e;
        }
        if (code8) tc8[li] = (uint8_t)((c < 0 || c >= 255) ? 255 : c);
        else tc[li] = c;
      }
      const size_t nb = (size_t)(bhi - blk);
      if (!fast_s) nt_copy(st + g0 + blk, ts, 4 * nb);
      if (!fast_e) nt_copy(en + g0 + blk, te, 4 * nb);
      if (code8) nt_copy(code8 + g0 + blk, tc8, nb);
      else nt_copy(code + g0 + blk, tc, 4 * nb);
    }
    nt_fence();
  });
  if (!pool_ok) return set_error(PBGPU_ENOMEM, "host allocation failed while encoding the %s table", side);
  if (range_err.load())
    return set_error(PBGPU_ERANGE, "%s table has a coordinate outside the int32 domain (reference limit: contigs < 2 Gb)", side);
  return PBGPU_OK;
}

// ---- output construction ------------------------------------------------------------------------
struct OwnedArray {  // private_data of every ArrowArray we produce
  std::vector<void *> bufs;        // malloc'd, freed on release
  std::vector<const void *> bptr;  // the `buffers` array
  std::vector<ArrowArray> kids;
  std::vector<ArrowArray *> kptr;
};
void release_array(ArrowArray *a) {
  if (!a || !a->release) return;
  OwnedArray *o = (OwnedArray *)a->private_data;
  for (auto &k : o->kids)
    if (k.release) k.release(&k);
  for (void *p : o->bufs) hfree(p);
  delete o;
  a->release = nullptr;
}
ArrowArray finish_array(OwnedArray *o, int64_t length, int64_t null_count) {
  ArrowArray a{};
  a.length = length;
  a.null_count = null_count;
  a.offset = 0;
  a.n_buffers = (int64_t)o->bptr.size();
  a.buffers = o->bptr.data();
  o->kptr.clear();
  for (auto &k : o->kids) o->kptr.push_back(&k);
  a.n_children = (int64_t)o->kptr.size();
  a.children = o->kptr.empty() ? nullptr : o->kptr.data();
  a.dictionary = nullptr;
  a.release = release_array;
  a.private_data = o;
  return a;
}

struct OwnedSchema {
  std::string format, name, metadata;
  bool has_md = false;
  std::vector<ArrowSchema> kids;
  std::vector<ArrowSchema *> kptr;
  std::unique_ptr<ArrowSchema> dict;
};
void release_schema(ArrowSchema *s) {
  if (!s || !s->release) return;
  OwnedSchema *o = (OwnedSchema *)s->private_data;
  for (auto &k : o->kids)
    if (k.release) k.release(&k);
  if (o->dict && o->dict->release) o->dict->release(o->dict.get());
  delete o;
  s->release = nullptr;
}
ArrowSchema make_schema(const std::string &format, const std::string &name, const char *metadata, int64_t flags,
                        std::vector<ArrowSchema> &&kids = {}) {
  OwnedSchema *o = new OwnedSchema();
  o->format = format;
  o->name = name;
  if (metadata) { o->metadata.assign(metadata, metadata_size(metadata)); o->has_md = true; }
  o->kids = std::move(kids);
  for (auto &k : o->kids) o->kptr.push_back(&k);
  ArrowSchema s{};
  s.format = o->format.c_str();
  s.name = o->name.c_str();
  s.metadata = o->has_md ? o->metadata.data() : nullptr;
  s.flags = flags;
  s.n_children = (int64_t)o->kptr.size();
  s.children = o->kptr.empty() ? nullptr : o->kptr.data();
  s.dictionary = nullptr;
  s.release = release_schema;
  s.private_data = o;
  return s;
}

// byte width of fixed-width primitive formats (0 = not fixed width / unsupported)
int fixed_width(const char *f) {
  if (!f || !f[0]) return 0;
  if (!f[1]) switch (f[0]) {
      case 'c': case 'C': return 1;
      case 's': case 'S': case 'e': return 2;
      case 'i': case 'I': case 'f': return 4;
      case 'l': case 'L': case 'g': return 8;
      default: return 0;
    }
  if (f[0] == 't') {
    if (!strncmp(f, "tdD", 3)) return 4;
    if (!strncmp(f, "tdm", 3)) return 8;
    if (!strncmp(f, "tts", 3) || !strncmp(f, "ttm", 3)) return 4;
    if (!strncmp(f, "ttu", 3) || !strncmp(f, "ttn", 3)) return 8;
    if (!strncmp(f, "ts", 2) || !strncmp(f, "tD", 2)) return 8;
    if (!strncmp(f, "tiM", 3)) return 4;
    if (!strncmp(f, "tiD", 3)) return 8;
    if (!strncmp(f, "tin", 3)) return 16;
    return 0;
  }
  if (f[0] == 'd' && f[1] == ':') {  // decimal: d:p,s[,bits]
    int p = 0, s = 0, bits = 128;
    if (sscanf(f, "d:%d,%d,%d", &p, &s, &bits) >= 2) return bits / 8;
    return 0;
  }
  if (f[0] == 'w' && f[1] == ':') return atoi(f + 2);
  return 0;
}

// output format of an input field (string views and dictionary-encoded strings come out as large_utf8)
std::string out_format(const ArrowSchema *f, bool *ok) {
  *ok = true;
  if (f->dictionary) {
    if (str_kind(f->dictionary->format) != StrKind::None && is_int_format(f->format)) return "U";
    *ok = false;
    return "";
  }
  StrKind k = str_kind(f->format);
  if (k == StrKind::View) return f->format[1] == 'u' ? "U" : "Z";
  if (k != StrKind::None) return f->format;
  if (!strcmp(f->format, "b") || !strcmp(f->format, "n") || fixed_width(f->format) > 0) return f->format;
  *ok = false;
  return "";
}

// Gather of one output column, split in two phases so that every column of a batch can share the same two
// parallel_for sweeps (tasks = columns x 16K-row chunks): pass 1 sizes variable-width data, pass 2 copies.
// rows[i] == PBGPU_NO_PARTNER -> null.
struct ColGather {
  const Table *t = nullptr;
  int col = 0;
  const uint32_t *rows = nullptr;
  int64_t n = 0, nchunks = 0;
  const ArrowSchema *f = nullptr;
  bool is_dict = false, is_bool = false, is_null_type = false, large = false;
  StrKind sk = StrKind::None;
  int w = 0;
  std::unique_ptr<OwnedArray> o;
  uint8_t *valid = nullptr, *vals = nullptr;
  void *offs = nullptr;
  char *data = nullptr;
  std::vector<int64_t> chunk_nulls, chunk_bytes, chunk_off;
  std::vector<uint32_t> lens;
  int64_t total = 0;

  int init(const Table &tab, int c, const uint32_t *r, int64_t count) {
    t = &tab; col = c; rows = r; n = count;
    nchunks = (n + kGatherChunk - 1) / kGatherChunk;
    f = tab.schema.children[c];
    o.reset(new OwnedArray());
    is_dict = f->dictionary != nullptr;
    sk = str_kind(is_dict ? f->dictionary->format : f->format);
    w = (!is_dict && sk == StrKind::None) ? fixed_width(f->format) : 0;
    is_bool = !is_dict && !strcmp(f->format, "b");
    is_null_type = !is_dict && !strcmp(f->format, "n");
    if (is_null_type) return PBGPU_OK;
    if (!(w > 0 || is_bool || sk != StrKind::None))
      return set_error(PBGPU_ESCHEMA, "payload column '%s' has unsupported Arrow type '%s' for materialised output (use emit=1 index pairs)",
                       f->name, f->format);
    const size_t vbytes = (size_t)((n + 7) / 8);
    valid = (uint8_t *)hcalloc(vbytes ? vbytes : 1);
    if (!valid) return set_error(PBGPU_ENOMEM, "host allocation failed");
    o->bufs.push_back(valid);
    chunk_nulls.assign((size_t)std::max<int64_t>(nchunks, 1), 0);
    if (sk != StrKind::None) {
      large = is_dict || sk != StrKind::Utf8;  // views / dictionaries come out as large_utf8
      chunk_bytes.assign(chunk_nulls.size(), 0);
      chunk_off.assign(chunk_nulls.size(), 0);
      lens.resize((size_t)n);
    } else {
      vals = (uint8_t *)(is_bool ? hcalloc(vbytes ? vbytes : 1) : hmalloc((size_t)n * w + 1));
      if (!vals) return set_error(PBGPU_ENOMEM, "host allocation failed");
      o->bufs.push_back(vals);
    }
    return PBGPU_OK;
  }
  static inline bool src_valid(const ArrowArray *a, int64_t j) {
    return a->null_count == 0 || !a->buffers[0] || bit_get((const uint8_t *)a->buffers[0], j);
  }
  inline bool fetch(uint32_t r, std::string_view *sv) const {
    if (r == PBGPU_NO_PARTNER) return false;
    int b; int64_t li;
    t->locate(r, b, li);
    const ArrowArray &ba = t->batches[b];
    const ArrowArray *a = ba.children[col];
    const int64_t j = li + a->offset + ba.offset;
    if (!src_valid(a, j)) return false;
    if (is_dict) {
      const ArrowArray *d = a->dictionary;
      const int64_t k = int_at(a->buffers[1], f->format[0], j);
      if (k < 0 || k >= d->length || !src_valid(d, k + d->offset)) return false;
      *sv = str_at(d, sk, k + d->offset);
    } else *sv = str_at(a, sk, j);
    return true;
  }
  void pass1(int64_t ci) {  // strings: lengths + validity
    if (sk == StrKind::None || is_null_type) return;
    const int64_t lo = ci * kGatherChunk, hi = std::min(n, lo + kGatherChunk);
    int64_t nulls = 0, bytes = 0;
    for (int64_t i = lo; i < hi; ++i) {
      std::string_view sv;
      if (fetch(rows[i], &sv)) { bit_set(valid, i); lens[i] = (uint32_t)sv.size(); bytes += (int64_t)sv.size(); }
      else { ++nulls; lens[i] = 0; }
    }
    chunk_nulls[ci] = nulls;
    chunk_bytes[ci] = bytes;
  }
  int alloc() {  // between the passes: offsets of the chunks, data buffers
    if (sk == StrKind::None || is_null_type) return PBGPU_OK;
    total = 0;
    for (int64_t ci = 0; ci < nchunks; ++ci) { chunk_off[ci] = total; total += chunk_bytes[ci]; }
    if (!large && total > INT32_MAX)
      return set_error(PBGPU_ERANGE, "utf8 column '%s' would exceed 2 GiB in one output batch; lower max_batch_rows or use large_utf8", f->name);
    offs = hmalloc((size_t)(n + 1) * (large ? 8 : 4));
    data = (char *)hmalloc((size_t)total + 1);
    if (!offs || !data) { hfree(offs); hfree(data); return set_error(PBGPU_ENOMEM, "host allocation failed"); }
    o->bufs.push_back(offs);
    o->bufs.push_back(data);
    if (large) ((int64_t *)offs)[n] = total; else ((int32_t *)offs)[n] = (int32_t)total;
    return PBGPU_OK;
  }
  void pass2(int64_t ci) {
    if (is_null_type) return;
    const int64_t lo = ci * kGatherChunk, hi = std::min(n, lo + kGatherChunk);
    if (sk != StrKind::None) {
      int64_t pos = chunk_off[ci];
      for (int64_t i = lo; i < hi; ++i) {
        if (large) ((int64_t *)offs)[i] = pos; else ((int32_t *)offs)[i] = (int32_t)pos;
        if (lens[i]) {
          std::string_view sv;
          fetch(rows[i], &sv);
          memcpy(data + pos, sv.data(), sv.size());
          pos += (int64_t)sv.size();
        }
      }
      return;
    }
    int64_t nulls = 0;
    const bool one = t->batches.size() == 1;
    for (int64_t i = lo; i < hi; ++i) {
      const uint32_t r = rows[i];
      if (r == PBGPU_NO_PARTNER) { ++nulls; if (!is_bool) memset(vals + (size_t)i * w, 0, w); continue; }
      int b = 0; int64_t li = r;
      if (!one) t->locate(r, b, li);
      const ArrowArray &ba = t->batches[b];
      const ArrowArray *a = ba.children[col];
      const int64_t j = li + a->offset + ba.offset;
      if (!src_valid(a, j)) { ++nulls; if (!is_bool) memset(vals + (size_t)i * w, 0, w); continue; }
      bit_set(valid, i);
      if (is_bool) { if (bit_get((const uint8_t *)a->buffers[1], j)) bit_set(vals, i); }
      else if (w == 4) ((uint32_t *)vals)[i] = ((const uint32_t *)a->buffers[1])[j];
      else if (w == 8) ((uint64_t *)vals)[i] = ((const uint64_t *)a->buffers[1])[j];
      else memcpy(vals + (size_t)i * w, (const uint8_t *)a->buffers[1] + (size_t)j * w, w);
    }
    chunk_nulls[ci] = nulls;
  }
  ArrowArray finish() {
    if (is_null_type) { o->bptr = {}; return finish_array(o.release(), n, n); }
    int64_t nulls = 0;
    for (auto v : chunk_nulls) nulls += v;
    if (sk != StrKind::None) o->bptr = {nulls ? (const void *)valid : nullptr, offs, data};
    else o->bptr = {nulls ? (const void *)valid : nullptr, vals};
    return finish_array(o.release(), n, nulls);
  }
};

// gather several (table, column, row list) jobs of one output batch together
struct GatherJob { const Table *t; int col; const uint32_t *rows; };
int gather_columns(const std::vector<GatherJob> &jobs, int64_t n, std::vector<ArrowArray> *out) {
  std::vector<ColGather> g(jobs.size());
  for (size_t k = 0; k < jobs.size(); ++k) { int rc = g[k].init(*jobs[k].t, jobs[k].col, jobs[k].rows, n); if (rc != PBGPU_OK) return rc; }
  const int64_t nchunks = (n + kGatherChunk - 1) / kGatherChunk, nj = (int64_t)jobs.size();
  bool any_str = false;
  for (auto &c : g) any_str |= c.sk != StrKind::None;
  bool ok = true;
  if (any_str) ok = Pool::get().parallel_for(nj * nchunks, [&](int64_t ti) { g[ti / nchunks].pass1(ti % nchunks); });
  for (auto &c : g) { int rc = c.alloc(); if (rc != PBGPU_OK) return rc; }
  ok = Pool::get().parallel_for(nj * nchunks, [&](int64_t ti) { g[ti / nchunks].pass2(ti % nchunks); }) && ok;
  if (!ok) return set_error(PBGPU_ENOMEM, "host allocation failed while gathering payload columns");
  for (auto &c : g) out->push_back(c.finish());
  return PBGPU_OK;
}

template <typename T>
int plain_column(const T *src, int64_t n, const uint8_t *null_if_zero_flag, ArrowArray *out) {
  (void)null_if_zero_flag;
  std::unique_ptr<OwnedArray> o(new OwnedArray());
  T *v = (T *)hmalloc(sizeof(T) * (size_t)(n ? n : 1));
  if (!v) return set_error(PBGPU_ENOMEM, "host allocation failed");
  memcpy(v, src, sizeof(T) * (size_t)n);
  o->bufs.push_back(v);
  o->bptr = {nullptr, v};
  *out = finish_array(o.release(), n, 0);
  return PBGPU_OK;
}

// int64 column where value < 0 means null (nearest distance)
int nullable_i64_column(const int64_t *src, int64_t n, ArrowArray *out) {
  std::unique_ptr<OwnedArray> o(new OwnedArray());
  int64_t *v = (int64_t *)hmalloc(8 * (size_t)(n ? n : 1));
  uint8_t *valid = (uint8_t *)hcalloc((size_t)((n + 7) / 8) + 1);
  if (!v || !valid) { hfree(v); hfree(valid); return set_error(PBGPU_ENOMEM, "host allocation failed"); }
  o->bufs.push_back(valid);
  o->bufs.push_back(v);
  int64_t nulls = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (src[i] < 0) { v[i] = 0; ++nulls; }
    else { v[i] = src[i]; bit_set(valid, i); }
  }
  o->bptr = {nulls ? (const void *)valid : nullptr, v};
  *out = finish_array(o.release(), n, nulls);
  return PBGPU_OK;
}

int nullable_u32_column(const uint32_t *src, int64_t n, ArrowArray *out) {
  std::unique_ptr<OwnedArray> o(new OwnedArray());
  uint32_t *v = (uint32_t *)hmalloc(4 * (size_t)(n ? n : 1));
  uint8_t *valid = (uint8_t *)hcalloc((size_t)((n + 7) / 8) + 1);
  if (!v || !valid) { hfree(v); hfree(valid); return set_error(PBGPU_ENOMEM, "host allocation failed"); }
  o->bufs.push_back(valid);
  o->bufs.push_back(v);
  int64_t nulls = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (src[i] == PBGPU_NO_PARTNER) { v[i] = 0; ++nulls; }
    else { v[i] = src[i]; bit_set(valid, i); }
  }
  o->bptr = {nulls ? (const void *)valid : nullptr, v};
  *out = finish_array(o.release(), n, nulls);
  return PBGPU_OK;
}

// ---- zero-copy re-export of input columns (count_overlaps / coverage return the iterated table's rows
// unchanged plus one column: operation.rs:316-347) ----------------------------------------------------
struct ViewPriv { std::shared_ptr<Table> keep; };
void release_view(ArrowArray *a) {
  if (!a || !a->release) return;
  delete (ViewPriv *)a->private_data;  // buffers / children / dictionary stay owned by the Table
  a->release = nullptr;
}
ArrowArray view_column(const std::shared_ptr<Table> &t, int batch, int col, int64_t off, int64_t len) {
  const ArrowArray &ba = t->batches[batch];
  const ArrowArray *c = ba.children[col];
  ArrowArray v = *c;
  v.offset = c->offset + ba.offset + off;
  v.length = len;
  v.null_count = c->null_count == 0 ? 0 : -1;  // unknown for a slice
  v.release = release_view;
  v.private_data = new ViewPriv{t};
  return v;
}
// zero-copy view of a slice of an engine result buffer (pinned, owned by the stream via `keep`)
struct BufViewPriv { std::shared_ptr<void> keep; const void *bufs[2]; };
void release_buf_view(ArrowArray *a) {
  if (!a || !a->release) return;
  delete (BufViewPriv *)a->private_data;
  a->release = nullptr;
}
ArrowArray view_buffer(const std::shared_ptr<void> &keep, const void *data, int64_t len) {
  BufViewPriv *p = new BufViewPriv{keep, {nullptr, data}};
  ArrowArray v{};
  v.length = len;
  v.null_count = 0;
  v.n_buffers = 2;
  v.buffers = p->bufs;
  v.release = release_buf_view;
  v.private_data = p;
  return v;
}
ArrowSchema copy_schema(const ArrowSchema *f, const std::string &name) {
  std::vector<ArrowSchema> kids;
  for (int64_t i = 0; i < f->n_children; ++i) kids.push_back(copy_schema(f->children[i], f->children[i]->name ? f->children[i]->name : ""));
  ArrowSchema s = make_schema(f->format, name, f->metadata, f->flags, std::move(kids));
  if (f->dictionary) {
    OwnedSchema *o = (OwnedSchema *)s.private_data;
    o->dict.reset(new ArrowSchema(copy_schema(f->dictionary, f->dictionary->name ? f->dictionary->name : "")));
    s.dictionary = o->dict.get();
  }
  return s;
}

// ---- key columns rebuilt from device-gathered int32 values ----------------------------------------------
// position column: the input dtype is preserved (Int32 in -> Int32 out, docs/supplement.md:328-333); result rows
// never carry null keys, so no validity bitmap
int pos_column(const std::shared_ptr<void> &keep, const int32_t *src, int64_t n, const char *fmt, ArrowArray *out) {
  const char f = fmt[0];
  if (f == 'i' || f == 'I') { *out = view_buffer(keep, src, n); return PBGPU_OK; }  // same 4-byte pattern (values >= 0 for 'I')
  const int w = fixed_width(fmt);
  std::unique_ptr<OwnedArray> o(new OwnedArray());
  uint8_t *v = (uint8_t *)hmalloc((size_t)n * w + 1);
  if (!v) return set_error(PBGPU_ENOMEM, "host allocation failed");
  o->bufs.push_back(v);
  const int64_t nch = (n + kGatherChunk - 1) / kGatherChunk;
  Pool::get().parallel_for(nch, [&](int64_t ci) {
    const int64_t lo = ci * kGatherChunk, hi = std::min(n, lo + kGatherChunk);
    switch (f) {
      case 'l': case 'L': for (int64_t i = lo; i < hi; ++i) ((int64_t *)v)[i] = src[i]; break;
      case 's': case 'S': for (int64_t i = lo; i < hi; ++i) ((int16_t *)v)[i] = (int16_t)src[i]; break;
      default: for (int64_t i = lo; i < hi; ++i) ((int8_t *)v)[i] = (int8_t)src[i]; break;
    }
  });
  o->bptr = {nullptr, v};
  *out = finish_array(o.release(), n, 0);
  return PBGPU_OK;
}

// contig column from dictionary codes: offsets + data filled sequentially (no random reads of the source table);
// the buffers are shared by the left and right contig columns of a join batch (equal by the join condition)
struct StrBufs { std::shared_ptr<HostBufs> keep; void *offs = nullptr; char *data = nullptr; bool large = false; };
template <typename CodeT>  // int32 codes, or uint8 when they travelled as bytes (<= 255 indexed contigs)
int contig_buffers(const CodeT *codes, int64_t n, const std::vector<std::string> &names, bool large, StrBufs *sb) {
  sb->keep = std::make_shared<HostBufs>();
  sb->large = large;
  const int64_t nch = (n + kGatherChunk - 1) / kGatherChunk;
  std::vector<int64_t> bytes((size_t)std::max<int64_t>(nch, 1), 0), off((size_t)std::max<int64_t>(nch, 1), 0);
  std::vector<uint8_t> uniform((size_t)std::max<int64_t>(nch, 1), 0);  // chunk repeats one contig (sorted output: nearly always)
  std::vector<uint32_t> len(names.size());
  for (size_t k = 0; k < names.size(); ++k) len[k] = (uint32_t)names[k].size();
  Pool::get().parallel_for(nch, [&](int64_t ci) {
    const int64_t lo = ci * kGatherChunk, hi = std::min(n, lo + kGatherChunk);
    const CodeT c0 = codes[lo];
    int diff = 0;
    for (int64_t i = lo; i < hi; ++i) diff |= (int)(codes[i] ^ c0);
    if (!diff) { uniform[ci] = 1; bytes[ci] = (int64_t)len[c0] * (hi - lo); return; }
    int64_t b = 0;
    for (int64_t i = lo; i < hi; ++i) b += len[codes[i]];
    bytes[ci] = b;
  });
  int64_t total = 0;
  for (int64_t ci = 0; ci < nch; ++ci) { off[ci] = total; total += bytes[ci]; }
  if (!large && total > INT32_MAX) return set_error(PBGPU_ERANGE, "utf8 contig column would exceed 2 GiB in one output batch; lower max_batch_rows");
  sb->offs = hmalloc((size_t)(n + 1) * (large ? 8 : 4));
  sb->data = (char *)hmalloc((size_t)total + 16);
  if (!sb->offs || !sb->data) { hfree(sb->offs); hfree(sb->data); return set_error(PBGPU_ENOMEM, "host allocation failed"); }
  sb->keep->v.push_back(sb->offs);
  sb->keep->v.push_back(sb->data);
  Pool::get().parallel_for(nch, [&](int64_t ci) {
    const int64_t lo = ci * kGatherChunk, hi = std::min(n, lo + kGatherChunk);
    int64_t pos = off[ci];
    if (uniform[ci]) {  // arithmetic offsets + one pattern replicated by doubling copies
      const std::string &nm = names[codes[lo]];
      const int64_t L = (int64_t)nm.size();
      if (large) { int64_t *o = (int64_t *)sb->offs; for (int64_t i = lo; i < hi; ++i) o[i] = pos + (i - lo) * L; }
      else { int32_t *o = (int32_t *)sb->offs; const int32_t p0 = (int32_t)pos, l32 = (int32_t)L; for (int64_t i = lo; i < hi; ++i) o[i] = p0 + (int32_t)(i - lo) * l32; }
      const int64_t tot = L * (hi - lo);
      if (tot > 0) {
        char *d = sb->data + pos;
        memcpy(d, nm.data(), (size_t)L);
        for (int64_t done = L; done < tot; ) { const int64_t c = std::min(done, tot - done); memcpy(d + done, d, (size_t)c); done += c; }
      }
      return;
    }
    for (int64_t i = lo; i < hi; ++i) {
      if (large) ((int64_t *)sb->offs)[i] = pos; else ((int32_t *)sb->offs)[i] = (int32_t)pos;
      const std::string &nm = names[codes[i]];
      memcpy(sb->data + pos, nm.data(), nm.size());
      pos += (int64_t)nm.size();
    }
  });
  if (large) ((int64_t *)sb->offs)[n] = total; else ((int32_t *)sb->offs)[n] = (int32_t)total;
  return PBGPU_OK;
}
struct StrViewPriv { std::shared_ptr<HostBufs> keep; const void *bufs[3]; };
void release_str_view(ArrowArray *a) {
  if (!a || !a->release) return;
  delete (StrViewPriv *)a->private_data;
  a->release = nullptr;
}
ArrowArray str_view(const StrBufs &sb, int64_t n) {
  StrViewPriv *p = new StrViewPriv{sb.keep, {nullptr, sb.offs, sb.data}};
  ArrowArray v{};
  v.length = n;
  v.null_count = 0;
  v.n_buffers = 3;
  v.buffers = p->bufs;
  v.release = release_str_view;
  v.private_data = p;
  return v;
}

#define BR_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) return set_error(PBGPU_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)
#define BR_TRY(expr) do { int _rc = (expr); if (_rc != PBGPU_OK) return _rc; } while (0)

struct DevBufs {  // device scratch of one call, freed stream-ordered
  cudaStream_t s = nullptr;
  std::vector<void *> v;
  template <typename T>
  T *get(size_t count) {
    void *p = nullptr;
    if (pbgpu::dev_alloc(&p, sizeof(T) * (count ? count : 1), s) != PBGPU_OK) return nullptr;  // block cache of pbgpu.cu
    v.push_back(p);
    return (T *)p;
  }
  void release() { for (void *p : v) pbgpu::dev_free(p, s); v.clear(); }
  ~DevBufs() { release(); }
};

// Everything one pbgpu_range_op call holds on the device.  It normally dies when the call returns; a streaming
// overlap (result larger than one ring slot) hands it to the output stream, which keeps emitting from it.
// The indexed ("build") side of a join, resident on the device: encoded key columns + the search structure.  One
// pbgpu_range_op call owns one; a pbgpu_range_open session shares it between all its pbgpu_range_probe calls (the
// build-once / probe-many shape of the reference's IntervalJoinExec: src/scan.rs:103-139 streams the probe side batch
// by batch against the collected build side).
struct IndexSide {
  int device = -1;
  cudaStream_t s = nullptr;        // the stream the columns were uploaded and the index was built on
  cudaEvent_t ready = nullptr;     // recorded behind the build: probing streams wait for it
  DevBufs dev;
  PinnedHold stage_wc;             // H2D staging of the key columns
  std::shared_ptr<Table> tab;
  ContigDict dict;                 // contig name -> code, as the index knows them
  int32_t n_contigs = 0;
  int64_t m = 0;
  int32_t *dc_x = nullptr, *ds_x = nullptr, *de_x = nullptr;
  int32_t *hc_x = nullptr, *hs_x = nullptr, *he_x = nullptr;  // encoded keys in pinned staging (until the upload has run)
  uint8_t *hc8_x = nullptr, *dc8_x = nullptr;  // <= 255 contigs: the codes are staged and uploaded as bytes, widened on the device
  pbgpu_index *ix = nullptr;
  // pbgpu_range_op runs the upload + index build on a helper thread while the calling thread encodes the iterated table
  // (the build has host round trips of its own: on the calling thread it would keep the encoder waiting); index_ready()
  // joins it.  The task holds a raw pointer to this object: the destructor waits for it first.
  std::future<int> pending;
  std::string pending_err;
  IndexSide() { stage_wc.wc = true; }
  IndexSide(const IndexSide &) = delete;
  IndexSide &operator=(const IndexSide &) = delete;
  ~IndexSide() {
    if (pending.valid()) pending.wait();
    int prev = -1;
    if (device >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != device) cudaSetDevice(device); else prev = -1;
    if (s) cudaStreamSynchronize(s);
    if (ix) pbgpu_index_free(ix);  // probing calls have synchronised their own streams before dropping their reference
    dev.release();
    if (ready) cudaEventDestroy(ready);
    if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Everything one probing call (pbgpu_range_op / pbgpu_range_probe) holds on the device.  It normally dies when the call
// returns; a streaming overlap (result larger than one ring slot) hands it to the output stream, which keeps emitting from it.
struct CallState {
  int device = -1;      // device the state lives on
  cudaStream_t s = nullptr;
  cudaStream_t s2 = nullptr;  // count_overlaps / coverage: kernel + D2H of slice k while slice k+1 uploads on `s`
  DevBufs dev;
  PinnedHold stage;     // D2H landing buffers the host reads (returned to the cache when the state dies)
  PinnedHold stage_wc;  // H2D staging: write-combined pinned memory, written once by the encoders, read by DMA
  std::shared_ptr<IndexSide> xs;  // the indexed side (shared with the session, if any)
  pbgpu_index *ix = nullptr;      // = xs->ix
  pbgpu_overlap_plan *plan = nullptr;
  CallState() { stage_wc.wc = true; }
  CallState(const CallState &) = delete;
  CallState &operator=(const CallState &) = delete;
  ~CallState() {
    int prev = -1;
    if (device >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != device) cudaSetDevice(device); else prev = -1;
    if (s2) { cudaStreamSynchronize(s2); cudaStreamDestroy(s2); }
    if (s) cudaStreamSynchronize(s);  // the plan is freed on the legacy stream: nothing of ours may still read it
    if (plan) pbgpu_overlap_plan_free(plan);
    dev.release();
    if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    xs.reset();  // after our streams have drained: the last reference frees the index
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Streaming sink of a materialised / index-pair overlap (BASELINE config 5): pass 2 runs over
// runs of 256-probe blocks whose pairs fit one ring slot; the slot's pairs (and the key columns gathered from them)
// are copied to the host, and the next chunk is enqueued while the consumer materialises the current one.
struct Sink {
  std::vector<uint64_t> offs;  // exclusive pair offset of every block (+ total)
  int64_t nblk = 0, next_blk = 0, cap = 0;
  bool mat = false, join = false, need_l = false, need_r = false;
  bool code8 = false;  // the contig code column of the result rows comes down as bytes (<= 255 indexed contigs)
  const int32_t *dc_i = nullptr, *ds_i = nullptr, *de_i = nullptr, *ds_x = nullptr, *de_x = nullptr;
  uint32_t *d_p = nullptr, *d_b = nullptr;  // the device slot
  int32_t *d_k[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  void *h_stage[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // pinned landing (only when direct D2H is off)
  // Payload columns gathered ON THE DEVICE (the reference materialises the joined rows itself,
  // operation.rs:272-303): fixed-width values of 1..16 bytes and utf8 / large_utf8 / binary, nulls included.  The column
  // is uploaded once per call; every chunk of pairs gathers from it and brings down ready-made Arrow buffers.  Other
  // types (bool, string views, dictionaries, nested) stay on the host gather path.
  struct DevCol {
    int table = 0, col = 0;    // 0 = left, 1 = right
    int width = 0;             // > 0: fixed width
    bool is_str = false, large = false;
    const void *d_data = nullptr;       // values / characters
    const uint8_t *d_valid = nullptr;   // one byte per source row, NULL = no nulls
    const long long *d_off = nullptr;   // strings: rows + 1 offsets into d_data
    void *d_out = nullptr;              // per chunk: cap values, or cap + 1 offsets
    uint32_t *d_bits = nullptr;
    unsigned long long *d_pos = nullptr, *d_meta = nullptr;  // strings: scanned lengths; [0] total bytes, [1] null count
    void *h_stage[3] = {nullptr, nullptr, nullptr};          // pinned landing buffers (bits / values or offsets / characters),
    size_t h_cap[3] = {0, 0, 0};                             // used when the final buffer is not page-locked; reused by every chunk
  };
  std::vector<DevCol> dcols;
  struct ColOut { const void *data = nullptr, *bits = nullptr, *offs = nullptr; };
  struct Chunk {
    bool valid = false;
    int64_t base = 0, rows = 0;
    std::shared_ptr<HostBufs> pins;
    std::vector<ColOut> cols;  // one per dcols entry
    const uint32_t *lrow = nullptr, *rrow = nullptr;
    const int32_t *k[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    struct Staged { void *fin; const void *landing; size_t bytes; };
    std::vector<Staged> staged;  // pinned landing -> final copies still to do after the sync
  } infl;  // the chunk in flight
};

// ---- the output stream ---------------------------------------------------------------------------
struct OutStream {
  std::shared_ptr<Table> left, right;  // shared with zero-copy output views of their columns
  PbRangeOptions opt{};
  std::string suffix1 = "_1", suffix2 = "_2";
  std::shared_ptr<HostBufs> pins = std::make_shared<HostBufs>();  // result arrays (host memory), shared with the zero-copy output views
  int64_t n_out = 0;         // result rows
  const uint32_t *lrow = nullptr;  // per result row: row of `left`  (may be NULL when not needed)
  const uint32_t *rrow = nullptr;  // per result row: row of `right` (PBGPU_NO_PARTNER = null)
  const int64_t *extra = nullptr;  // count / coverage / distance (distance < 0 = null)
  std::vector<uint32_t> own_l, own_r;  // host-built row lists (nearest expansion, distinct)
  std::vector<int64_t> own_x;
  // overlap, materialised: key columns of the result rows gathered on the device (NULL = not available)
  const uint8_t *k_code8 = nullptr;  // the same column as k_code when it travelled as bytes
  const int32_t *k_code = nullptr, *k_ls = nullptr, *k_le = nullptr, *k_rs = nullptr, *k_re = nullptr;
  std::vector<Sink::ColOut> dev_cols;     // payload columns of the resident chunk that were gathered on the device
  std::vector<std::string> contig_names;  // dictionary: code -> contig string
  int64_t cursor = 0;
  int64_t chunk_base = 0, chunk_rows = 0;  // result rows [chunk_base, chunk_base + chunk_rows) are behind lrow / rrow / k_*
  std::unique_ptr<CallState> call;         // device state (kept past the call only by a streaming overlap)
  std::unique_ptr<Sink> sink;
  int view_batch = 0;        // pass-through modes: input batch being re-exported
  int64_t view_off = 0;      //   and the offset inside it
  uint32_t batch_rows = 1 << 20;
  std::string last_error;
};

// a batch-sized window [lo, lo + n) of a device-gathered payload column of the resident chunk: zero-copy, the chunk's
// buffers (validity bits, values or offsets + characters) with the Arrow `offset` field doing the slicing
struct DevColPriv { std::shared_ptr<void> keep; const void *bufs[3]; };
void release_dev_col(ArrowArray *a) {
  if (!a || !a->release) return;
  delete (DevColPriv *)a->private_data;
  a->release = nullptr;
}
ArrowArray dev_col_view(const std::shared_ptr<void> &keep, const Sink::DevCol &dc, const Sink::ColOut &co, int64_t lo, int64_t n) {
  DevColPriv *p = new DevColPriv{keep, {co.bits, dc.is_str ? co.offs : co.data, dc.is_str ? co.data : nullptr}};
  ArrowArray v{};
  v.length = n;
  v.offset = lo;
  v.null_count = co.bits ? -1 : 0;
  v.n_buffers = dc.is_str ? 3 : 2;
  v.buffers = p->bufs;
  v.release = release_dev_col;
  v.private_data = p;
  return v;
}

// enqueue pass 2 + key gathers + D2H of the next chunk on the call's stream (no sync)
int sink_enqueue(OutStream *st) {
  Sink &sk = *st->sink;
  CallState &cs = *st->call;
  Sink::Chunk &ch = sk.infl;
  ch = Sink::Chunk();
  int64_t lo = sk.next_blk, hi = lo, rows = 0;
  while (lo < sk.nblk) {  // a run of whole blocks that fits the slot (at least one block), skipping empty runs
    hi = (int64_t)(std::upper_bound(sk.offs.begin() + lo, sk.offs.end(), sk.offs[lo] + (uint64_t)sk.cap) - sk.offs.begin()) - 1;
    if (hi <= lo) hi = lo + 1;
    rows = (int64_t)(sk.offs[hi] - sk.offs[lo]);
    if (rows > 0) break;
    lo = hi;
  }
  sk.next_blk = hi;
  if (lo >= sk.nblk || rows <= 0) return PBGPU_OK;  // nothing left
  cudaStream_t s = cs.s;
  BR_TRY(pbgpu_overlap_emit_blocks(cs.plan, lo, hi, sk.d_p, sk.d_b, s));
  ch.pins = std::make_shared<HostBufs>();
  int slot = 0;
  auto fetch = [&](const void *d_src, const void **dst, size_t width) -> int {
    bool dma = false;
    void *fin = hmalloc_dma(width * (size_t)rows, &dma);
    if (!fin) return set_error(PBGPU_ENOMEM, "host allocation failed");
    ch.pins->v.push_back(fin);
    *dst = fin;
    void *h = fin;
    if (!dma) {
      if (!sk.h_stage[slot]) sk.h_stage[slot] = cs.stage.get<uint32_t>((size_t)sk.cap);
      h = sk.h_stage[slot];
      if (!h) return set_error(PBGPU_ENOMEM, "pinned allocation failed");
      ch.staged.push_back({fin, h, width * (size_t)rows});
    }
    ++slot;
    BR_CUDA(cudaMemcpyAsync(h, d_src, width * (size_t)rows, cudaMemcpyDeviceToHost, s));
    return PBGPU_OK;
  };
  if (sk.mat) {  // key columns of the result rows: gathered where they already live
    struct G { const int32_t *src; const uint32_t *rows; bool on; };
    const G gs[5] = {{sk.dc_i, sk.d_p, true}, {sk.ds_i, sk.d_p, true}, {sk.de_i, sk.d_p, true}, {sk.ds_x, sk.d_b, sk.join}, {sk.de_x, sk.d_b, sk.join}};
    for (int j = 0; j < 5; ++j) {
      if (!gs[j].on) continue;
      if (j == 0 && sk.code8) {
        BR_TRY(pbgpu::gather_i32_u8(gs[j].src, gs[j].rows, rows, (uint8_t *)sk.d_k[j], s));
        BR_TRY(fetch(sk.d_k[j], (const void **)&ch.k[j], 1));
        continue;
      }
      BR_TRY(pbgpu_gather_i32(gs[j].src, gs[j].rows, rows, sk.d_k[j], s));
      BR_TRY(fetch(sk.d_k[j], (const void **)&ch.k[j], 4));
    }
  }
  if (sk.need_l) BR_TRY(fetch(sk.d_p, (const void **)&ch.lrow, 4));
  if (sk.need_r) BR_TRY(fetch(sk.d_b, (const void **)&ch.rrow, 4));
  // payload columns that live on the device: gather by the chunk's row ids, bring down ready-made Arrow buffers
  auto fetch_bytes = [&](Sink::DevCol &dc, int which, const void *d_src, size_t bytes, const void **dst) -> int {
    bool dma = false;
    void *fin = hmalloc_dma(bytes ? bytes : 1, &dma);
    if (!fin) return set_error(PBGPU_ENOMEM, "host allocation failed");
    ch.pins->v.push_back(fin);
    *dst = fin;
    if (!bytes) return PBGPU_OK;
    void *h = fin;
    if (!dma) {  // small buffers are not page-locked: land in the column's pinned slot (free again once sink_advance has copied it out)
      if (dc.h_cap[which] < bytes) {
        const size_t cap = std::max(bytes, 2 * dc.h_cap[which]);
        dc.h_stage[which] = cs.stage.get<char>(cap);
        dc.h_cap[which] = dc.h_stage[which] ? cap : 0;
      }
      h = dc.h_stage[which];
      if (!h) return set_error(PBGPU_ENOMEM, "pinned allocation failed");
      ch.staged.push_back({fin, h, bytes});
    }
    BR_CUDA(cudaMemcpyAsync(h, d_src, bytes, cudaMemcpyDeviceToHost, s));
    return PBGPU_OK;
  };
  ch.cols.assign(sk.dcols.size(), Sink::ColOut());
  for (size_t k = 0; k < sk.dcols.size(); ++k) {
    Sink::DevCol &dc = sk.dcols[k];
    const uint32_t *d_rows = dc.table == 0 ? sk.d_p : sk.d_b;
    if (dc.d_valid) {
      BR_TRY(pbgpu::gather_valid_bits(dc.d_valid, d_rows, rows, dc.d_bits, nullptr, s));
      BR_TRY(fetch_bytes(dc, 0, dc.d_bits, 4 * (size_t)((rows + 31) / 32), &ch.cols[k].bits));
    }
    if (!dc.is_str) {
      BR_TRY(pbgpu::gather_fixed(dc.d_data, dc.width, d_rows, rows, dc.d_out, s));
      BR_TRY(fetch_bytes(dc, 1, dc.d_out, (size_t)dc.width * (size_t)rows, &ch.cols[k].data));
      continue;
    }
    BR_TRY(pbgpu::gather_str_plan(dc.d_off, dc.d_valid, d_rows, rows, dc.d_pos, dc.d_meta, s));
    unsigned long long total = 0;  // bytes of this chunk's strings: sizes the host and device buffers (one sync per string column)
    BR_CUDA(cudaMemcpyAsync(&total, dc.d_meta, 8, cudaMemcpyDeviceToHost, s));
    BR_CUDA(cudaStreamSynchronize(s));
    if (!dc.large && total > (unsigned long long)INT32_MAX)
      return set_error(PBGPU_ERANGE, "utf8 column would exceed 2 GiB in one chunk of the result; lower sink_pairs or use large_utf8");
    void *d_chars_v = nullptr;
    BR_TRY(pbgpu::dev_alloc(&d_chars_v, (size_t)total + 16, s));
    char *d_chars = (char *)d_chars_v;
    int rc_s = pbgpu::gather_str_bytes(dc.d_off, (const char *)dc.d_data, d_rows, rows, dc.d_pos, dc.d_meta, dc.d_out, dc.large ? 1 : 0, d_chars, s);
    if (rc_s == PBGPU_OK) rc_s = fetch_bytes(dc, 1, dc.d_out, (size_t)(dc.large ? 8 : 4) * (size_t)(rows + 1), &ch.cols[k].offs);
    if (rc_s == PBGPU_OK) rc_s = fetch_bytes(dc, 2, d_chars, (size_t)total, &ch.cols[k].data);
    pbgpu::dev_free(d_chars_v, s);  // stream-ordered: behind the copy that reads it
    if (rc_s != PBGPU_OK) return rc_s;
  }
  ch.base = (int64_t)sk.offs[lo];
  ch.rows = rows;
  ch.valid = true;
  return PBGPU_OK;
}

// make the chunk in flight the current one (sync), then start the next; the device state is dropped after the last
int sink_advance(OutStream *st) {
  Sink &sk = *st->sink;
  if (!st->call) return set_error(PBGPU_EINVAL, "streaming overlap: device state already released");
  CallState &cs = *st->call;
  int prev = -1;
  if (cudaGetDevice(&prev) == cudaSuccess && prev != cs.device) cudaSetDevice(cs.device); else prev = -1;
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};
  if (!sk.infl.valid) BR_TRY(sink_enqueue(st));
  if (!sk.infl.valid) return set_error(PBGPU_EINVAL, "streaming overlap: ran out of chunks before the last row");
  BR_CUDA(cudaStreamSynchronize(cs.s));
  Sink::Chunk &ch = sk.infl;
  for (auto &c : ch.staged) copy_into(c.fin, c.landing, c.bytes);
  st->pins = ch.pins;
  st->lrow = ch.lrow;
  st->rrow = ch.rrow ? ch.rrow : ch.lrow;  // never dereferenced when not needed; keeps emit=1 paths well defined
  st->k_code = sk.code8 ? nullptr : ch.k[0];
  st->k_code8 = sk.code8 ? (const uint8_t *)ch.k[0] : nullptr;
  st->k_ls = ch.k[1]; st->k_le = ch.k[2]; st->k_rs = ch.k[3]; st->k_re = ch.k[4];
  st->dev_cols = ch.cols;
  st->chunk_base = ch.base;
  st->chunk_rows = ch.rows;
  ch.valid = false;
  if (st->chunk_base + st->chunk_rows < st->n_out) BR_TRY(sink_enqueue(st));  // prefetch while the consumer materialises
  if (!sk.infl.valid) st->call.reset();  // last chunk is on the host: free the device side now
  return PBGPU_OK;
}

bool want_distance(const PbRangeOptions &o) { return o.range_op == PBGPU_OP_NEAREST && o.compute_distance; }

int out_get_schema(ArrowArrayStream *s, ArrowSchema *out) {
  OutStream *st = (OutStream *)s->private_data;
  const PbRangeOptions &o = st->opt;
  std::vector<ArrowSchema> kids;
  auto add_table = [&](const Table &t, const std::string &suffix, bool force_nullable) -> int {
    for (int64_t i = 0; i < t.schema.n_children; ++i) {
      const ArrowSchema *f = t.schema.children[i];
      bool ok;
      std::string fmt = out_format(f, &ok);
      if (!ok) { st->last_error = std::string("unsupported payload type '") + f->format + "' in column '" + f->name + "'"; return 1; }
      int64_t flags = f->flags | (force_nullable ? ARROW_FLAG_NULLABLE : 0);
      kids.push_back(make_schema(fmt, std::string(f->name) + suffix, f->metadata, flags));
    }
    return 0;
  };
  int rc = 0;
  if (o.emit == 1) {
    kids.push_back(make_schema("I", "left_row", nullptr, 0));
    kids.push_back(make_schema("I", "right_row", nullptr, o.range_op == PBGPU_OP_NEAREST ? ARROW_FLAG_NULLABLE : 0));
    if (want_distance(o)) kids.push_back(make_schema("l", "distance", nullptr, ARROW_FLAG_NULLABLE));
  } else if (o.range_op == PBGPU_OP_OVERLAP) {
    if (o.output_mode == PBGPU_OUT_JOIN) { rc = add_table(*st->left, st->suffix1, false); if (!rc) rc = add_table(*st->right, st->suffix2, false); }
    else rc = add_table(*st->left, "", false);
  } else if (o.range_op == PBGPU_OP_NEAREST) {
    rc = add_table(*st->left, st->suffix1, false);
    if (!rc) rc = add_table(*st->right, st->suffix2, true);
    if (!rc && want_distance(o)) kids.push_back(make_schema("l", "distance", nullptr, ARROW_FLAG_NULLABLE));
  } else {  // count_overlaps / coverage: the iterated table's rows + one int64 column (operation.rs:316-347)
    for (int64_t i = 0; i < st->right->schema.n_children; ++i)  // columns are re-exported untouched, any Arrow type
      kids.push_back(copy_schema(st->right->schema.children[i], st->right->schema.children[i]->name ? st->right->schema.children[i]->name : ""));
    kids.push_back(make_schema("l", o.range_op == PBGPU_OP_COVERAGE ? "coverage" : "count", nullptr, 0));
  }
  if (rc) {
    for (auto &k : kids) k.release(&k);
    return 22;  // EINVAL
  }
  *out = make_schema("+s", "", nullptr, 0, std::move(kids));
  return 0;
}

int out_get_next(ArrowArrayStream *s, ArrowArray *out) {
  OutStream *st = (OutStream *)s->private_data;
  const PbRangeOptions &o = st->opt;
  memset(out, 0, sizeof(*out));
  if (st->cursor >= st->n_out) return 0;  // end of stream: released (release == NULL) array
  Trace tr;
  struct LapEnd { Trace &t; ~LapEnd() { t.lap("get_next (materialise batch)"); } } lap_end{tr};
  if (st->sink && st->cursor >= st->chunk_base + st->chunk_rows) {  // streaming overlap: next chunk of pairs
    if (sink_advance(st) != PBGPU_OK) { st->last_error = pbgpu::g_err; return 5; }
  }
  const int64_t row0 = st->cursor;  // first result row of this batch; `lo` = its position inside the resident chunk
  const int64_t lo = row0 - st->chunk_base;
  const int64_t n = std::min<int64_t>(std::min<int64_t>(st->batch_rows, st->n_out - row0), st->chunk_rows - lo);
  std::unique_ptr<OwnedArray> top(new OwnedArray());
  int rc = PBGPU_OK;
  auto push = [&](ArrowArray &&a) { top->kids.push_back(a); };
  std::vector<GatherJob> jobs;
  auto add_table = [&](const Table &t, const uint32_t *rows) {
    for (int c = 0; c < (int)t.n_cols(); ++c) jobs.push_back(GatherJob{&t, c, rows + lo});
  };
  auto run_jobs = [&]() {
    std::vector<ArrowArray> cols;
    rc = gather_columns(jobs, n, &cols);
    if (rc == PBGPU_OK) for (auto &a : cols) push(std::move(a));
    else for (auto &a : cols) if (a.release) a.release(&a);
  };
  const std::shared_ptr<void> keep = st->pins;
  const bool pinned_rows = st->own_l.empty();  // row lists straight from the engine (pinned) vs host-built vectors
  if (o.emit == 1) {
    if (pinned_rows) push(view_buffer(keep, st->lrow + lo, n));
    else { ArrowArray a{}; rc = plain_column<uint32_t>(st->lrow + lo, n, nullptr, &a); if (rc == PBGPU_OK) push(std::move(a)); }
    if (rc == PBGPU_OK) {
      if (o.range_op == PBGPU_OP_NEAREST) { ArrowArray b{}; rc = nullable_u32_column(st->rrow + lo, n, &b); if (rc == PBGPU_OK) push(std::move(b)); }
      else if (pinned_rows) push(view_buffer(keep, st->rrow + lo, n));
      else { ArrowArray b{}; rc = plain_column<uint32_t>(st->rrow + lo, n, nullptr, &b); if (rc == PBGPU_OK) push(std::move(b)); }
    }
    if (rc == PBGPU_OK && want_distance(o)) { ArrowArray d{}; rc = nullable_i64_column(st->extra + lo, n, &d); if (rc == PBGPU_OK) push(std::move(d)); }
  } else if (o.range_op == PBGPU_OP_OVERLAP && (st->k_code || st->k_code8)) {
    // key columns come from the device-gathered int32 arrays; only true payload columns are gathered on the host
    struct Slot { int kind; const Table *t; int col; const int32_t *pos; const uint32_t *rows; int dev; };  // 0 contig, 1 position, 2 payload (host gather), 3 payload (device-gathered buffers)
    std::vector<Slot> slots;
    auto plan_table = [&](const Table &t, int which, const uint32_t *rows, const int32_t *ks, const int32_t *ke) {
      for (int c = 0; c < (int)t.n_cols(); ++c) {
        if (c == t.key[0]) slots.push_back({0, &t, c, nullptr, nullptr, -1});
        else if (c == t.key[1]) slots.push_back({1, &t, c, ks + lo, nullptr, -1});
        else if (c == t.key[2]) slots.push_back({1, &t, c, ke + lo, nullptr, -1});
        else {
          int dev = -1;
          if (st->sink)
            for (size_t k = 0; k < st->sink->dcols.size() && k < st->dev_cols.size(); ++k)
              if (st->sink->dcols[k].table == which && st->sink->dcols[k].col == c) dev = (int)k;
          if (dev >= 0) slots.push_back({3, &t, c, nullptr, nullptr, dev});
          else { slots.push_back({2, &t, c, nullptr, rows + lo, -1}); jobs.push_back(GatherJob{&t, c, rows + lo}); }
        }
      }
    };
    plan_table(*st->left, 0, st->lrow, st->k_ls, st->k_le);
    if (o.output_mode == PBGPU_OUT_JOIN) plan_table(*st->right, 1, st->rrow, st->k_rs, st->k_re);
    std::vector<ArrowArray> payload;
    if (!jobs.empty()) rc = gather_columns(jobs, n, &payload);
    StrBufs sb_small, sb_large;
    size_t next_payload = 0;
    for (size_t k = 0; k < slots.size() && rc == PBGPU_OK; ++k) {
      const Slot &sl = slots[k];
      const ArrowSchema *f = sl.t->schema.children[sl.col];
      if (sl.kind == 2) { push(std::move(payload[next_payload++])); continue; }
      if (sl.kind == 3) { push(dev_col_view(keep, st->sink->dcols[(size_t)sl.dev], st->dev_cols[(size_t)sl.dev], lo, n)); continue; }
      if (sl.kind == 1) { ArrowArray a{}; rc = pos_column(keep, sl.pos, n, f->format, &a); if (rc == PBGPU_OK) push(std::move(a)); continue; }
      bool ok;
      const bool large = out_format(f, &ok)[0] == 'U';
      StrBufs &sb = large ? sb_large : sb_small;
      if (!sb.keep) rc = st->k_code8 ? contig_buffers(st->k_code8 + lo, n, st->contig_names, large, &sb)
                                     : contig_buffers(st->k_code + lo, n, st->contig_names, large, &sb);
      if (rc == PBGPU_OK) push(str_view(sb, n));
    }
    for (; next_payload < payload.size(); ++next_payload) if (payload[next_payload].release) payload[next_payload].release(&payload[next_payload]);
  } else if (o.range_op == PBGPU_OP_OVERLAP) {
    add_table(*st->left, st->lrow);
    if (o.output_mode == PBGPU_OUT_JOIN) add_table(*st->right, st->rrow);
    run_jobs();
  } else if (o.range_op == PBGPU_OP_NEAREST) {
    add_table(*st->left, st->lrow);
    add_table(*st->right, st->rrow);
    run_jobs();
    if (rc == PBGPU_OK && want_distance(o)) { ArrowArray d{}; rc = nullable_i64_column(st->extra + lo, n, &d); if (rc == PBGPU_OK) push(std::move(d)); }
  } else {
    // zero-copy: slice the current input batch (output batches follow the input batch boundaries)
    const Table &t = *st->right;
    while (st->view_batch < (int)t.batches.size() && st->view_off >= t.batches[st->view_batch].length) { ++st->view_batch; st->view_off = 0; }
    const int64_t avail = t.batches[st->view_batch].length - st->view_off;
    const int64_t len = std::min<int64_t>(std::min<int64_t>(n, avail), st->n_out - lo);
    for (int c = 0; c < (int)t.n_cols(); ++c) push(view_column(st->right, st->view_batch, c, st->view_off, len));
    push(view_buffer(keep, st->extra + lo, len));  // the count / coverage column: a view of the D2H buffer
    if (rc == PBGPU_OK) {
      top->bptr = {nullptr};
      *out = finish_array(top.release(), len, 0);
      st->cursor += len;
      st->view_off += len;
      return 0;
    }
  }
  if (rc != PBGPU_OK) {
    st->last_error = pbgpu::g_err;
    for (auto &k : top->kids) if (k.release) k.release(&k);
    return 5;  // EIO
  }
  top->bptr = {nullptr};  // struct validity
  *out = finish_array(top.release(), n, 0);
  st->cursor += n;
  return 0;
}

const char *out_last_error(ArrowArrayStream *s) {
  OutStream *st = (OutStream *)s->private_data;
  return st->last_error.empty() ? nullptr : st->last_error.c_str();
}
void out_release(ArrowArrayStream *s) {
  if (!s || !s->release) return;
  delete (OutStream *)s->private_data;
  s->release = nullptr;
}

// Payload columns of one input table to the device (Sink::DevCol): fixed-width values and utf8 / binary columns, each as
// ONE contiguous device column whatever the batch structure of the input, plus a byte-per-row validity column when any
// batch has nulls.  PBGPU_DEV_GATHER=0 keeps every payload column on the host gather path.
int upload_payload(const Table &t, int which, Sink &sk, CallState &cs) {
  static const bool enabled = [] { const char *e = getenv("PBGPU_DEV_GATHER"); return !(e && e[0] == '0'); }();
  if (!enabled || t.rows == 0) return PBGPU_OK;
  cudaStream_t s = cs.s;
  const int64_t rows = t.rows;
  for (int c = 0; c < (int)t.n_cols(); ++c) {
    if (c == t.key[0] || c == t.key[1] || c == t.key[2]) continue;
    const ArrowSchema *f = t.schema.children[c];
    if (f->dictionary) continue;
    const int w = fixed_width(f->format);
    const StrKind k = str_kind(f->format);
    const bool fixed = (w == 1 || w == 2 || w == 4 || w == 8 || w == 16);
    const bool str = k == StrKind::Utf8 || k == StrKind::LargeUtf8;
    if (!fixed && !str) continue;
    Sink::DevCol dc;
    dc.table = which; dc.col = c; dc.width = fixed ? w : 0; dc.is_str = str; dc.large = k == StrKind::LargeUtf8;
    bool any_null = false;
    for (const ArrowArray &ba : t.batches) { const ArrowArray *a = ba.children[c]; if (a->null_count != 0 && a->buffers[0]) any_null = true; }
    if (any_null) {
      uint8_t *hv = cs.stage_wc.get<uint8_t>((size_t)rows);
      uint8_t *dv = cs.dev.get<uint8_t>((size_t)rows);
      if (!hv || !dv) return set_error(PBGPU_ENOMEM, "allocation failed");
      for (size_t b = 0; b < t.batches.size(); ++b) {
        const ArrowArray &ba = t.batches[b];
        const ArrowArray *a = ba.children[c];
        const uint8_t *bits = (a->null_count != 0) ? (const uint8_t *)a->buffers[0] : nullptr;
        const int64_t off = a->offset + ba.offset, len = ba.length, g0 = t.start[b];
        const int64_t nch = (len + kGatherChunk - 1) / kGatherChunk;
        Pool::get().parallel_for(nch, [&](int64_t ci) {
          const int64_t lo = ci * kGatherChunk, hi = std::min(len, lo + kGatherChunk);
          if (!bits) memset(hv + g0 + lo, 1, (size_t)(hi - lo));
          else for (int64_t i = lo; i < hi; ++i) hv[g0 + i] = bit_get(bits, off + i) ? 1 : 0;
        });
      }
      BR_CUDA(cudaMemcpyAsync(dv, hv, (size_t)rows, cudaMemcpyHostToDevice, s));
      dc.d_valid = dv;
      dc.d_bits = cs.dev.get<uint32_t>((size_t)(sk.cap + 31) / 32 + 1);
      if (!dc.d_bits) return set_error(PBGPU_ENOMEM, "device allocation failed");
    }
    if (fixed) {
      char *dd = cs.dev.get<char>((size_t)rows * w);
      dc.d_out = cs.dev.get<char>((size_t)sk.cap * w);
      if (!dd || !dc.d_out) return set_error(PBGPU_ENOMEM, "device allocation failed");
      for (size_t b = 0; b < t.batches.size(); ++b) {
        const ArrowArray &ba = t.batches[b];
        const ArrowArray *a = ba.children[c];
        const size_t bytes = (size_t)ba.length * w;
        if (a->buffers[1]) BR_CUDA(cudaMemcpyAsync(dd + (size_t)t.start[b] * w, (const char *)a->buffers[1] + (size_t)(a->offset + ba.offset) * w, bytes, cudaMemcpyHostToDevice, s));
        else BR_CUDA(cudaMemsetAsync(dd + (size_t)t.start[b] * w, 0, bytes, s));
      }
      dc.d_data = dd;
    } else {
      const bool large = dc.large;
      auto off_at = [&](const ArrowArray *a, int64_t i) -> int64_t {
        return large ? ((const int64_t *)a->buffers[1])[i] : (int64_t)((const int32_t *)a->buffers[1])[i];
      };
      int64_t total = 0;
      std::vector<int64_t> base(t.batches.size());
      for (size_t b = 0; b < t.batches.size(); ++b) {
        const ArrowArray &ba = t.batches[b];
        const ArrowArray *a = ba.children[c];
        const int64_t o0 = a->offset + ba.offset;
        base[b] = total;
        total += off_at(a, o0 + ba.length) - off_at(a, o0);
      }
      char *dd = cs.dev.get<char>((size_t)total + 16);
      long long *doff = cs.dev.get<long long>((size_t)rows + 1);
      dc.d_out = cs.dev.get<char>((size_t)(sk.cap + 1) * (large ? 8 : 4));
      dc.d_pos = cs.dev.get<unsigned long long>((size_t)sk.cap + 1);
      dc.d_meta = cs.dev.get<unsigned long long>(2);
      if (!dd || !doff || !dc.d_out || !dc.d_pos || !dc.d_meta) return set_error(PBGPU_ENOMEM, "device allocation failed");
      for (size_t b = 0; b < t.batches.size(); ++b) {
        const ArrowArray &ba = t.batches[b];
        const ArrowArray *a = ba.children[c];
        const int64_t o0 = a->offset + ba.offset, len = ba.length;
        const size_t ow = large ? 8 : 4;
        char *d_raw = cs.dev.get<char>((size_t)(len + 1) * ow);
        if (!d_raw) return set_error(PBGPU_ENOMEM, "device allocation failed");
        BR_CUDA(cudaMemcpyAsync(d_raw, (const char *)a->buffers[1] + (size_t)o0 * ow, (size_t)(len + 1) * ow, cudaMemcpyHostToDevice, s));
        BR_TRY(pbgpu::rebase_offsets(d_raw, large ? 1 : 0, len, (long long)base[b], doff + t.start[b], s));
        const int64_t c0 = off_at(a, o0), c1 = off_at(a, o0 + len);
        if (c1 > c0) BR_CUDA(cudaMemcpyAsync(dd + base[b], (const char *)a->buffers[2] + c0, (size_t)(c1 - c0), cudaMemcpyHostToDevice, s));
      }
      dc.d_data = dd;
      dc.d_off = doff;
    }
    sk.dcols.push_back(dc);
  }
  return PBGPU_OK;
}

// Upload + index the indexed side.  `side`: what to call it in error messages.
// Indexed side, phase A (calling thread; uses the host pool): stream, staging, key encoding + the contig dictionary.
int prepare_index_encode(const PbRangeOptions &o, std::shared_ptr<Table> IXt, const char *side, std::shared_ptr<IndexSide> *out) {
  auto xs = std::make_shared<IndexSide>();
  xs->tab = IXt;
  Table *IX = IXt.get();
  Trace tr;
  const int64_t m = IX->rows;
  xs->m = m;
  int prev_dev = -1;
  BR_CUDA(cudaGetDevice(&prev_dev));
  if (o.device >= 0 && o.device != prev_dev) BR_CUDA(cudaSetDevice(o.device)); else prev_dev = -1;
  struct DevRestore { int d; ~DevRestore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev};
  BR_CUDA(cudaGetDevice(&xs->device));
  BR_CUDA(cudaStreamCreateWithFlags(&xs->s, cudaStreamNonBlocking));
  BR_CUDA(cudaEventCreateWithFlags(&xs->ready, cudaEventDisableTiming));
  xs->dev.s = xs->s;
  xs->hc8_x = xs->stage_wc.get<uint8_t>(m); xs->hs_x = xs->stage_wc.get<int32_t>(m); xs->he_x = xs->stage_wc.get<int32_t>(m);
  if (!xs->hc8_x || !xs->hs_x || !xs->he_x) return set_error(PBGPU_ENOMEM, "pinned staging allocation failed");
  xs->dc_x = xs->dev.get<int32_t>(m); xs->ds_x = xs->dev.get<int32_t>(m); xs->de_x = xs->dev.get<int32_t>(m);
  xs->dc8_x = xs->dev.get<uint8_t>(m);
  if (!xs->dc_x || !xs->ds_x || !xs->de_x || !xs->dc8_x) return set_error(PBGPU_ENOMEM, "device allocation failed");
  // Contigs that only occur on the iterated side get codes >= n_contigs of the index and are treated as null keys by
  // the kernels (they cannot match anything anyway).
  // The encoder is bound by its streaming stores into the staging buffers (r2A: 0.55 ns per row with 12 bytes written, 0.41
  // with 9): contig codes are staged as BYTES (255 = null key), like the iterated side's, and widened on the device.  A
  // table with more than 255 contigs is encoded again with 32-bit codes (the dictionary is complete by then: same codes).
  BR_TRY(encode_keys(*IX, side, xs->dict, nullptr, xs->hs_x, xs->he_x, 0, INT64_MAX, xs->hc8_x));
  xs->n_contigs = (int32_t)xs->dict.map.size();
  if (xs->n_contigs > 255) {
    xs->hc_x = xs->stage_wc.get<int32_t>(m);
    if (!xs->hc_x) return set_error(PBGPU_ENOMEM, "pinned staging allocation failed");
    BR_TRY(encode_keys(*IX, side, xs->dict, xs->hc_x, xs->hs_x, xs->he_x));
    xs->hc8_x = nullptr;
  }
  tr.lap("encode indexed side");
  *out = xs;
  return PBGPU_OK;
}
// Phase B (any thread; no pool use): upload, index build, `ready` recorded behind it on the side's stream.
int prepare_index_build(IndexSide *xs) {
  Trace tr;
  int prev_dev = -1;
  BR_CUDA(cudaGetDevice(&prev_dev));
  if (xs->device != prev_dev) BR_CUDA(cudaSetDevice(xs->device)); else prev_dev = -1;
  struct DevRestore { int d; ~DevRestore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev};
  cudaStream_t s = xs->s;
  const int64_t m = xs->m;
  if (xs->hc8_x) {
    BR_CUDA(cudaMemcpyAsync(xs->dc8_x, xs->hc8_x, (size_t)m, cudaMemcpyHostToDevice, s));
    BR_TRY(pbgpu::widen_codes_u8(xs->dc8_x, m, xs->n_contigs, xs->dc_x, s));
  } else BR_CUDA(cudaMemcpyAsync(xs->dc_x, xs->hc_x, 4 * (size_t)m, cudaMemcpyHostToDevice, s));
  BR_CUDA(cudaMemcpyAsync(xs->ds_x, xs->hs_x, 4 * (size_t)m, cudaMemcpyHostToDevice, s));
  BR_CUDA(cudaMemcpyAsync(xs->de_x, xs->he_x, 4 * (size_t)m, cudaMemcpyHostToDevice, s));
  BR_TRY(pbgpu_index_build(xs->dc_x, xs->ds_x, xs->de_x, m, xs->n_contigs, s, &xs->ix));
  BR_CUDA(cudaEventRecord(xs->ready, s));
  tr.lap("H2D indexed side + index build");
  return PBGPU_OK;
}
int prepare_index(const PbRangeOptions &o, std::shared_ptr<Table> IXt, const char *side, std::shared_ptr<IndexSide> *out) {
  std::shared_ptr<IndexSide> xs;
  BR_TRY(prepare_index_encode(o, IXt, side, &xs));
  BR_TRY(prepare_index_build(xs.get()));
  *out = xs;
  return PBGPU_OK;
}
// joins the helper thread of pbgpu_range_op (no-op for sessions): afterwards xs->ix is valid and `ready` is recorded
int index_ready(IndexSide &xs) {
  if (!xs.pending.valid()) return PBGPU_OK;
  const int rc = xs.pending.get();
  if (rc != PBGPU_OK) return set_error(rc, "%s", xs.pending_err.c_str());
  return PBGPU_OK;
}

int run_iter(Table *L, Table *R, OutStream *os, std::shared_ptr<IndexSide> xs);

int run(Table *L, Table *R, OutStream *os) {
  const PbRangeOptions &o = os->opt;
  // roles: which table is indexed, which is iterated (see pbgpu.h)
  //   overlap: probe/iterate = left (df1), index = right (df2)             operation.rs:253-263
  //   nearest: iterate = left (df1), index = right (df2)                   operation.rs:143-158
  //   count/coverage: index = left (s1), iterate = right (s2), rows of right returned   operation.rs:316-340
  const bool iter_is_left = (o.range_op == PBGPU_OP_OVERLAP || o.range_op == PBGPU_OP_NEAREST);
  std::shared_ptr<IndexSide> xs;
  BR_TRY(prepare_index_encode(o, iter_is_left ? os->right : os->left, iter_is_left ? "right" : "left", &xs));
  static const bool async_build = [] { const char *e = getenv("PBGPU_ASYNC_BUILD"); return !(e && e[0] == '0'); }();
  // a fresh thread allocates its own host mailbox and stage events (~0.3 ms): only worth it when the build it hides is longer
  static const int64_t async_min_rows = [] { const char *e = getenv("PBGPU_ASYNC_MIN_ROWS"); return e ? (int64_t)atoll(e) : (int64_t)4 << 20; }();
  if (async_build && os->left->rows + os->right->rows >= async_min_rows) {
    IndexSide *px = xs.get();
    xs->pending = std::async(std::launch::async, [px]() -> int {
      const int rc = prepare_index_build(px);
      if (rc != PBGPU_OK) px->pending_err = pbgpu::g_err;  // the message lives in this thread's buffer
      pbgpu::release_thread_state();
      return rc;
    });
  } else BR_TRY(prepare_index_build(xs.get()));
  return run_iter(L, R, os, xs);
}

// One iterated ("probe") table against a resident indexed side.
int run_iter(Table *L, Table *R, OutStream *os, std::shared_ptr<IndexSide> xs) {
  const PbRangeOptions &o = os->opt;
  const bool iter_is_left = (o.range_op == PBGPU_OP_OVERLAP || o.range_op == PBGPU_OP_NEAREST);
  Table *IT = iter_is_left ? L : R, *IX = iter_is_left ? R : L;
  Trace tr;
  ContigDict dict;  // a private copy: encoding the iterated side adds the contigs only it has
  { std::lock_guard<std::mutex> lk(xs->dict.mu); dict.map = xs->dict.map; }
  os->call.reset(new CallState());
  CallState &cs = *os->call;  // dies with this call (end of run) unless a streaming overlap keeps it
  cs.xs = xs;
  PinnedHold &stage = cs.stage, &stage_wc = cs.stage_wc;
  const int64_t n = IT->rows, m = IX->rows;
  int32_t *hs_i = stage_wc.get<int32_t>(n), *he_i = stage_wc.get<int32_t>(n);
  if (!hs_i || !he_i) return set_error(PBGPU_ENOMEM, "pinned staging allocation failed");

  int prev_dev = -1;
  BR_CUDA(cudaGetDevice(&prev_dev));
  if (xs->device != prev_dev) BR_CUDA(cudaSetDevice(xs->device)); else prev_dev = -1;
  struct DevRestore { int d; ~DevRestore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev};
  cs.device = xs->device;
  BR_CUDA(cudaStreamCreateWithFlags(&cs.s, cudaStreamNonBlocking));
  cudaStream_t s = cs.s;
  cs.dev.s = s;
  DevBufs &dev = cs.dev;
  int32_t *dc_x = xs->dc_x, *ds_x = xs->ds_x, *de_x = xs->de_x;
  (void)dc_x; (void)m;
  int32_t *dc_i = dev.get<int32_t>(n), *ds_i = dev.get<int32_t>(n), *de_i = dev.get<int32_t>(n);
  if (!dc_i || !ds_i || !de_i) return set_error(PBGPU_ENOMEM, "device allocation failed");
  tr.lap_drained("  stream + device allocations", s);
  const int32_t n_contigs = xs->n_contigs;
  // The index may still be under construction on the helper thread of pbgpu_range_op: encoding and uploading the iterated
  // table needs only the dictionary.  need_index() joins the build and orders this call's stream behind it (the event must
  // have been RECORDED before a wait on it is enqueued, hence join first).
  pbgpu_index *ix = nullptr;
  auto need_index = [&]() -> int {
    if (ix) return PBGPU_OK;
    BR_TRY(index_ready(*xs));
    BR_CUDA(cudaStreamWaitEvent(s, xs->ready, 0));  // uploads and the index build ran on the indexed side's stream
    ix = xs->ix;
    cs.ix = ix;
    return PBGPU_OK;
  };
  // iterated side in slices: the DMA of slice k runs while the host encodes slice k+1.  With at most 255 indexed
  // contigs the contig codes travel as bytes and are widened on the device.
  // count_overlaps / coverage are row-local, so they join the pipeline: the kernel of slice k and the D2H of its
  // result run on a second stream while slice k+1 is still on its way up (PCIe is full duplex), and the host widens
  // slice k while slice k+1 comes down.
  const bool row_local = o.range_op == PBGPU_OP_COUNT_OVERLAPS_NAIVE || o.range_op == PBGPU_OP_COVERAGE;
  const bool is_cov = o.range_op == PBGPU_OP_COVERAGE;
  struct SliceOut { int64_t lo, hi; cudaEvent_t up, down; };
  std::vector<SliceOut> slices;
  struct EvGuard { std::vector<SliceOut> &v; ~EvGuard() { for (auto &x : v) { if (x.up) cudaEventDestroy(x.up); if (x.down) cudaEventDestroy(x.down); } } } evg{slices};
  uint32_t *d_cnt = nullptr, *h_cnt = nullptr;  // count: u32 on the wire (half the D2H bytes), widened on the host
  int64_t *d_cov = nullptr, *h_cov = nullptr, *fin64 = nullptr;
  bool cov_dma = false;
  if (row_local) {
    BR_CUDA(cudaStreamCreateWithFlags(&cs.s2, cudaStreamNonBlocking));
    fin64 = (int64_t *)(is_cov ? hmalloc_dma(8 * (size_t)(n ? n : 1), &cov_dma) : hmalloc(8 * (size_t)(n ? n : 1)));
    if (!fin64) return set_error(PBGPU_ENOMEM, "host allocation failed");
    os->pins->v.push_back(fin64);
    if (is_cov) { d_cov = dev.get<int64_t>(n); h_cov = cov_dma ? fin64 : stage.get<int64_t>(n); if (!d_cov || !h_cov) return set_error(PBGPU_ENOMEM, "allocation failed"); }
    else { d_cnt = dev.get<uint32_t>(n); h_cnt = stage.get<uint32_t>(n); if (!d_cnt || !h_cnt) return set_error(PBGPU_ENOMEM, "allocation failed"); }
  }
  size_t launched = 0;  // row-local operations: slices whose kernel + D2H have been enqueued
  auto launch_slice = [&](SliceOut &so) -> int {
    const int64_t lo = so.lo, hi = so.hi;
    BR_CUDA(cudaStreamWaitEvent(cs.s2, xs->ready, 0));  // the index build (its event is recorded: need_index() ran)
    BR_CUDA(cudaStreamWaitEvent(cs.s2, so.up, 0));
    if (is_cov) {
      BR_TRY(pbgpu_coverage(ix, dc_i + lo, ds_i + lo, de_i + lo, hi - lo, o.filter_op, d_cov + lo, cs.s2));
      BR_CUDA(cudaMemcpyAsync(h_cov + lo, d_cov + lo, 8 * (size_t)(hi - lo), cudaMemcpyDeviceToHost, cs.s2));
    } else {
      BR_TRY(pbgpu::count_overlaps_u32(ix, dc_i + lo, ds_i + lo, de_i + lo, hi - lo, o.filter_op, d_cnt + lo, cs.s2));
      BR_CUDA(cudaMemcpyAsync(h_cnt + lo, d_cnt + lo, 4 * (size_t)(hi - lo), cudaMemcpyDeviceToHost, cs.s2));
    }
    BR_CUDA(cudaEventRecord(so.down, cs.s2));
    return PBGPU_OK;
  };
  {
    const bool narrow = n_contigs <= 255;
    int32_t *hc_i = narrow ? nullptr : stage_wc.get<int32_t>(n);
    uint8_t *hc8 = narrow ? stage_wc.get<uint8_t>(n) : nullptr;
    uint8_t *dc8 = narrow ? dev.get<uint8_t>(n) : nullptr;
    if ((!narrow && !hc_i) || (narrow && (!hc8 || !dc8))) return set_error(PBGPU_ENOMEM, "staging allocation failed");
    const int64_t slice = std::max<int64_t>(1 << 20, (n + 7) / 8);
    slices.reserve(8);
    for (int64_t lo = 0; lo < n; lo += slice) {
      const int64_t hi = std::min(n, lo + slice);
      int rc = encode_keys(*IT, iter_is_left ? "left" : "right", dict, hc_i, hs_i, he_i, lo, hi, hc8);
      if (rc != PBGPU_OK) return rc;
      if (narrow) {
        cudaMemcpyAsync(dc8 + lo, hc8 + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, s);
        rc = pbgpu::widen_codes_u8(dc8 + lo, hi - lo, n_contigs, dc_i + lo, s);
        if (rc != PBGPU_OK) return rc;
      } else
        cudaMemcpyAsync(dc_i + lo, hc_i + lo, 4 * (size_t)(hi - lo), cudaMemcpyHostToDevice, s);
      cudaMemcpyAsync(ds_i + lo, hs_i + lo, 4 * (size_t)(hi - lo), cudaMemcpyHostToDevice, s);
      cudaMemcpyAsync(de_i + lo, he_i + lo, 4 * (size_t)(hi - lo), cudaMemcpyHostToDevice, s);
      if (row_local) {
        slices.push_back({lo, hi, nullptr, nullptr});
        SliceOut &so = slices.back();
        BR_CUDA(cudaEventCreateWithFlags(&so.up, cudaEventDisableTiming));
        BR_CUDA(cudaEventCreateWithFlags(&so.down, cudaEventDisableTiming));
        BR_CUDA(cudaEventRecord(so.up, s));
        // kernels of the slices uploaded so far start as soon as the index is there; while it is still being built on the
        // helper thread the loop keeps encoding and uploading (the last slice waits for it)
        const bool last = hi >= n;
        const bool ready = !xs->pending.valid() || xs->pending.wait_for(std::chrono::seconds(0)) == std::future_status::ready;
        if (ready || last) {
          BR_TRY(need_index());
          for (; launched < slices.size(); ++launched) BR_TRY(launch_slice(slices[launched]));
        }
      }
    }
    if (cudaGetLastError() != cudaSuccess) return set_error(PBGPU_ECUDA, "H2D copy of the iterated table failed");
  }
  tr.lap("encode + H2D iterated side");
  BR_TRY(need_index());
  const uint64_t limit = o.limit;

  if (row_local) {
    for (const SliceOut &so : slices) {
      BR_CUDA(cudaEventSynchronize(so.down));
      if (is_cov && cov_dma) continue;  // landed in the result buffer directly
      const int64_t lo = so.lo, hi = so.hi, nch = (hi - lo + kGatherChunk - 1) / kGatherChunk;
      Pool::get().parallel_for(nch, [&](int64_t ci) {
        const int64_t a = lo + ci * kGatherChunk, b = std::min(hi, a + kGatherChunk);
        if (is_cov) memcpy(fin64 + a, h_cov + a, 8 * (size_t)(b - a));
        else for (int64_t i = a; i < b; ++i) fin64[i] = (int64_t)h_cnt[i];
      });
    }
    BR_CUDA(cudaStreamSynchronize(cs.s2));
    BR_CUDA(cudaStreamSynchronize(s));
    os->extra = fin64;
    os->n_out = n;
  } else if (o.range_op == PBGPU_OP_OVERLAP && o.output_mode == PBGPU_OUT_LEFT_DISTINCT) {
    int64_t *d_out = dev.get<int64_t>(n);
    int64_t *h_out = stage.get<int64_t>(n);
    if (!d_out || !h_out) return set_error(PBGPU_ENOMEM, "allocation failed");
    BR_TRY(pbgpu_count_overlaps(ix, dc_i, ds_i, de_i, n, o.filter_op, d_out, s));
    BR_CUDA(cudaMemcpyAsync(h_out, d_out, 8 * (size_t)n, cudaMemcpyDeviceToHost, s));
    BR_CUDA(cudaStreamSynchronize(s));
    for (int64_t i = 0; i < n; ++i) if (h_out[i] > 0) os->own_l.push_back((uint32_t)i);
    os->lrow = os->own_l.data();
    os->rrow = os->own_l.data();  // unused in Left modes; keeps emit=1 well defined
    os->n_out = (int64_t)os->own_l.size();
  } else if (o.range_op == PBGPU_OP_OVERLAP) {
    pbgpu_overlap_plan *plan = nullptr;
    int64_t total = 0;
    BR_TRY(pbgpu_overlap_count(ix, dc_i, ds_i, de_i, n, o.filter_op, s, &plan, &total));
    cs.plan = plan;
    os->n_out = total;
    if (total > 0) {
      // pass 2 goes through the streaming sink: one chunk when the result fits a ring slot (the common case; the
      // device state is then released before this call returns), else chunk by chunk from out->get_next
      os->sink.reset(new Sink());
      Sink &sk = *os->sink;
      const bool join = o.output_mode == PBGPU_OUT_JOIN;
      auto has_payload = [](const Table &t) { for (int c = 0; c < (int)t.n_cols(); ++c) if (c != t.key[0] && c != t.key[1] && c != t.key[2]) return true; return false; };
      sk.mat = o.emit == 0;
      sk.join = join;
      sk.need_l = !sk.mat || has_payload(*L);
      sk.need_r = !sk.mat || (join && has_payload(*R));
      sk.code8 = n_contigs <= 255;
      sk.dc_i = dc_i; sk.ds_i = ds_i; sk.de_i = de_i; sk.ds_x = ds_x; sk.de_x = de_x;
      sk.nblk = pbgpu_overlap_plan_blocks(plan);
      sk.offs.resize((size_t)sk.nblk + 1);
      BR_TRY(pbgpu_overlap_plan_block_offsets(plan, sk.offs.data(), s));
      uint64_t widest = 0;
      for (int64_t b = 0; b < sk.nblk; ++b) widest = std::max(widest, sk.offs[b + 1] - sk.offs[b]);
      uint64_t cap = o.sink_pairs;
      if (!cap) { const char *e = getenv("PBGPU_SINK_PAIRS"); cap = e ? strtoull(e, nullptr, 10) : 0; }
      if (!cap) cap = (uint64_t)1 << 24;
      const uint64_t want = limit ? std::min<uint64_t>((uint64_t)total, limit + widest) : (uint64_t)total;  // rows ever read
      sk.cap = (int64_t)std::max<uint64_t>(std::min<uint64_t>(cap, want), widest);
      sk.d_p = dev.get<uint32_t>((size_t)sk.cap);
      sk.d_b = dev.get<uint32_t>((size_t)sk.cap);
      if (!sk.d_p || !sk.d_b) return set_error(PBGPU_ENOMEM, "device allocation failed for a ring slot of %lld pairs", (long long)sk.cap);
      if (sk.mat) {
        for (int j = 0; j < 5; ++j) {
          if (j >= 3 && !join) continue;
          sk.d_k[j] = dev.get<int32_t>((size_t)sk.cap);
          if (!sk.d_k[j]) return set_error(PBGPU_ENOMEM, "device allocation failed");
        }
        os->contig_names.resize(dict.map.size());
        for (auto &kv : dict.map) os->contig_names[(size_t)kv.second] = kv.first;
        BR_TRY(upload_payload(*L, 0, sk, cs));
        if (join) BR_TRY(upload_payload(*R, 1, sk, cs));
      }
      if (limit && (uint64_t)os->n_out > limit) os->n_out = (int64_t)limit;  // before the first prefetch decision
      BR_TRY(sink_enqueue(os));
      BR_TRY(sink_advance(os));  // first chunk resident; drops the device state when it was the only one
    }
  } else if (o.range_op == PBGPU_OP_NEAREST) {
    const int64_t k = o.nearest_k ? (int64_t)o.nearest_k : 1;
    uint32_t *d_p = dev.get<uint32_t>((size_t)(n * k));
    int64_t *d_d = dev.get<int64_t>((size_t)(n * k));
    uint32_t *h_p = stage.get<uint32_t>((size_t)(n * k));
    int64_t *h_d = stage.get<int64_t>((size_t)(n * k));
    if (!d_p || !d_d || !h_p || !h_d) return set_error(PBGPU_ENOMEM, "allocation failed");
    BR_TRY(pbgpu_nearest(ix, dc_i, ds_i, de_i, n, o.filter_op, k, o.include_overlaps, d_p, d_d, s));
    BR_CUDA(cudaMemcpyAsync(h_p, d_p, 4 * (size_t)(n * k), cudaMemcpyDeviceToHost, s));
    BR_CUDA(cudaMemcpyAsync(h_d, d_d, 8 * (size_t)(n * k), cudaMemcpyDeviceToHost, s));
    BR_CUDA(cudaStreamSynchronize(s));
    // expand: one row per found partner; a probe row with no partner at all yields one null-partner row
    os->own_l.reserve((size_t)n); os->own_r.reserve((size_t)n); os->own_x.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
      bool any = false;
      for (int64_t j = 0; j < k; ++j) {
        if (h_p[i * k + j] == PBGPU_NO_PARTNER) break;
        os->own_l.push_back((uint32_t)i); os->own_r.push_back(h_p[i * k + j]); os->own_x.push_back(h_d[i * k + j]);
        any = true;
      }
      if (!any) { os->own_l.push_back((uint32_t)i); os->own_r.push_back(PBGPU_NO_PARTNER); os->own_x.push_back(-1); }
    }
    os->lrow = os->own_l.data(); os->rrow = os->own_r.data(); os->extra = os->own_x.data();
    os->n_out = (int64_t)os->own_l.size();
  } else {
    return set_error(PBGPU_EINVAL, "unsupported range_op %d", o.range_op);
  }
  if (!os->sink) {  // everything is on the host already
    os->chunk_base = 0;
    os->chunk_rows = os->n_out;
    os->call.reset();
  }
  if (limit && (uint64_t)os->n_out > limit) os->n_out = (int64_t)limit;
  tr.lap("provider kernels + D2H");
  return PBGPU_OK;
}


// ================================================================================================================
// Unary sweeps at the Arrow level: merge / cluster / complement / subtract -- the host glue of
// /root/reference/src/operation.rs:352-510 (do_merge / do_cluster / do_complement / do_subtract) over the device calls
// pbgpu_merge / pbgpu_cluster / pbgpu_subtract (unary.cuh).  Contig codes are ranked in lexicographic name order before
// they go to the device, so merged intervals come out ordered by contig name and cluster ids count the way bioframe's do
// (tests/test_bioframe.py:392-411).  Output schemas (positions always Int64, like the reference's providers):
//   merge       contig, start, end (named like the input's interval columns), n_intervals    tests/_expected.py:174-181
//   cluster     every input column (zero-copy) + cluster, cluster_start, cluster_end          ..._regressions.py:49-59
//   complement  contig, start, end                                                            ..._regressions.py:33-39
//   subtract    every df1 column, the interval columns holding the remaining pieces           ..._regressions.py:41-47
// Rows with a null contig / start / end take no part (cluster drops them from its output as well): parity unpinned.
// ================================================================================================================
struct VecStream {  // a finished result: schema + batches handed out one by one
  ArrowSchema schema{};
  std::vector<ArrowArray> batches;
  size_t next = 0;
  ~VecStream() {
    for (auto &b : batches) if (b.release) b.release(&b);
    if (schema.release) schema.release(&schema);
  }
};
int vs_get_schema(ArrowArrayStream *s, ArrowSchema *out) {
  VecStream *v = (VecStream *)s->private_data;
  *out = copy_schema(&v->schema, "");
  return 0;
}
int vs_get_next(ArrowArrayStream *s, ArrowArray *out) {
  VecStream *v = (VecStream *)s->private_data;
  memset(out, 0, sizeof(*out));
  if (v->next >= v->batches.size()) return 0;
  *out = v->batches[v->next];
  v->batches[v->next].release = nullptr;  // moved
  ++v->next;
  return 0;
}
const char *vs_last_error(ArrowArrayStream *) { return nullptr; }
void vs_release(ArrowArrayStream *s) {
  if (!s || !s->release) return;
  delete (VecStream *)s->private_data;
  s->release = nullptr;
}

// int64 column widened from int32 values; open_end: INT32_MAX stands for "no upper bound" (complement's default view)
int widened_column(const int32_t *src, int64_t n, bool open_end, ArrowArray *out) {
  std::unique_ptr<OwnedArray> o(new OwnedArray());
  int64_t *v = (int64_t *)hmalloc(8 * (size_t)(n ? n : 1));
  if (!v) return set_error(PBGPU_ENOMEM, "host allocation failed");
  o->bufs.push_back(v);
  for (int64_t i = 0; i < n; ++i) v[i] = (open_end && src[i] == INT32_MAX) ? INT64_MAX : (int64_t)src[i];
  o->bptr = {nullptr, v};
  *out = finish_array(o.release(), n, 0);
  return PBGPU_OK;
}

struct UnaryKeys {  // (contig code, start, end) of one table on the host (page-locked, ordinary caching: they are read back)
  PinnedHold hold;
  int32_t *c = nullptr, *s = nullptr, *e = nullptr;
  int64_t n = 0;
  int alloc(int64_t rows) {
    n = rows;
    c = hold.get<int32_t>((size_t)(rows ? rows : 1)); s = hold.get<int32_t>((size_t)(rows ? rows : 1)); e = hold.get<int32_t>((size_t)(rows ? rows : 1));
    return (c && s && e) ? PBGPU_OK : set_error(PBGPU_ENOMEM, "pinned staging allocation failed");
  }
};

struct UnaryDev {  // stream + device columns of one unary call
  int prev_dev = -1;
  cudaStream_t s = nullptr;
  DevBufs dev;
  ~UnaryDev() {
    if (s) cudaStreamSynchronize(s);
    dev.release();
    if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
  }
  int open(int device) {
    int cur = -1;
    BR_CUDA(cudaGetDevice(&cur));
    if (device >= 0 && device != cur) { BR_CUDA(cudaSetDevice(device)); prev_dev = cur; }
    BR_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    dev.s = s;
    return PBGPU_OK;
  }
  int upload(const UnaryKeys &k, int32_t **dc, int32_t **ds, int32_t **de) {
    const size_t n = (size_t)k.n;
    *dc = dev.get<int32_t>(n); *ds = dev.get<int32_t>(n); *de = dev.get<int32_t>(n);
    if (!*dc || !*ds || !*de) return set_error(PBGPU_ENOMEM, "device allocation failed");
    if (n) {
      BR_CUDA(cudaMemcpyAsync(*dc, k.c, 4 * n, cudaMemcpyHostToDevice, s));
      BR_CUDA(cudaMemcpyAsync(*ds, k.s, 4 * n, cudaMemcpyHostToDevice, s));
      BR_CUDA(cudaMemcpyAsync(*de, k.e, 4 * n, cudaMemcpyHostToDevice, s));
    }
    return PBGPU_OK;
  }
  // device column -> freshly allocated host buffer owned by `keep`
  template <typename T>
  int download(const T *d_src, int64_t n, const std::shared_ptr<HostBufs> &keep, T **out) {
    T *h = (T *)hmalloc(sizeof(T) * (size_t)(n ? n : 1));
    if (!h) return set_error(PBGPU_ENOMEM, "host allocation failed");
    keep->v.push_back(h);
    if (n) BR_CUDA(cudaMemcpyAsync(h, d_src, sizeof(T) * (size_t)n, cudaMemcpyDeviceToHost, s));
    *out = h;
    return PBGPU_OK;
  }
};

int run_unary(const PbRangeOptions &o, std::shared_ptr<Table> L, std::shared_ptr<Table> R, ArrowArrayStream *out) {
  const int op = o.range_op;
  const bool two = R != nullptr;  // subtract: the right table; complement: the view table
  if (op == PBGPU_OP_SUBTRACT && !two) return set_error(PBGPU_EINVAL, "subtract needs two tables");
  if (o.min_dist < 0) return set_error(PBGPU_EINVAL, "min_dist must be >= 0");
  // 1. keys of both tables against one dictionary, then codes re-ranked in lexicographic name order
  ContigDict dict;
  UnaryKeys kl, kr;
  BR_TRY(kl.alloc(L->rows));
  BR_TRY(encode_keys(*L, "left", dict, kl.c, kl.s, kl.e));
  if (two) {
    BR_TRY(kr.alloc(R->rows));
    BR_TRY(encode_keys(*R, op == PBGPU_OP_COMPLEMENT ? "view" : "right", dict, kr.c, kr.s, kr.e));
  }
  const int32_t n_contigs = (int32_t)dict.map.size();
  std::vector<std::string> by_code((size_t)n_contigs), names((size_t)n_contigs);
  for (auto &kv : dict.map) by_code[(size_t)kv.second] = kv.first;
  std::vector<int32_t> order((size_t)n_contigs), rank((size_t)n_contigs);
  for (int32_t i = 0; i < n_contigs; ++i) order[(size_t)i] = i;
  std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return by_code[(size_t)a] < by_code[(size_t)b]; });
  for (int32_t r = 0; r < n_contigs; ++r) { rank[(size_t)order[(size_t)r]] = r; names[(size_t)r] = by_code[(size_t)order[(size_t)r]]; }
  auto rerank = [&](UnaryKeys &k) {
    const int64_t nch = (k.n + kChunk - 1) / kChunk;
    Pool::get().parallel_for(nch, [&](int64_t ci) {
      const int64_t lo = ci * kChunk, hi = std::min(k.n, lo + kChunk);
      for (int64_t i = lo; i < hi; ++i) { const int32_t c = k.c[i]; k.c[i] = (c >= 0 && c < n_contigs) ? rank[(size_t)c] : -1; }
    });
  };
  rerank(kl);
  if (two) rerank(kr);
  // complement without a view table: every contig present spans [0, INT64_MAX) (polars_bio/range_op.py:726-729); swept as
  // [0, INT32_MAX] -- no int32 coordinate reaches that end -- and the open end is put back afterwards
  bool open_end = false;
  if (op == PBGPU_OP_COMPLEMENT && !two) {
    std::vector<uint8_t> present((size_t)n_contigs, 0);
    for (int64_t i = 0; i < kl.n; ++i) if (kl.c[i] >= 0) present[(size_t)kl.c[i]] = 1;
    int64_t nv = 0;
    for (auto v : present) nv += v;
    BR_TRY(kr.alloc(nv));
    int64_t j = 0;
    for (int32_t c = 0; c < n_contigs; ++c) if (present[(size_t)c]) { kr.c[j] = c; kr.s[j] = 0; kr.e[j] = INT32_MAX; ++j; }
    open_end = true;
  }
  // 2. device
  UnaryDev ud;
  BR_TRY(ud.open(o.device));
  cudaStream_t s = ud.s;
  auto keep = std::make_shared<HostBufs>();
  std::unique_ptr<VecStream> vs(new VecStream());
  std::vector<ArrowSchema> fields;
  struct FieldGuard { std::vector<ArrowSchema> &f; bool armed = true; ~FieldGuard() { if (armed) for (auto &k : f) if (k.release) k.release(&k); } } fguard{fields};
  const ArrowSchema *fc = L->schema.children[L->key[0]];
  bool ok_fmt = true;
  const std::string contig_fmt = out_format(fc, &ok_fmt);  // utf8 stays utf8; views / dictionaries come out as large_utf8
  const bool contig_large = contig_fmt == "U";
  auto name_of = [&](const Table &t, int c) { return std::string(t.schema.children[c]->name ? t.schema.children[c]->name : ""); };
  auto push_batch = [&](std::unique_ptr<OwnedArray> &top, int64_t rows) {
    top->bptr = {nullptr};
    vs->batches.push_back(finish_array(top.release(), rows, 0));
  };
  int32_t *dlc = nullptr, *dls = nullptr, *dle = nullptr;
  if (op == PBGPU_OP_MERGE) {
    BR_TRY(ud.upload(kl, &dlc, &dls, &dle));
    pbgpu_intervals *h = nullptr;
    BR_TRY(pbgpu_merge(dlc, dls, dle, kl.n, n_contigs, o.filter_op, o.min_dist, s, &h));
    struct Free { pbgpu_intervals *h; cudaStream_t s; ~Free() { pbgpu_intervals_free(h, s); } } fr{h, s};
    const int64_t n = pbgpu_intervals_rows(h);
    const int32_t *dc = nullptr, *ds = nullptr, *de = nullptr;
    const int64_t *dn = nullptr;
    BR_TRY(pbgpu_intervals_columns(h, &dc, nullptr, &ds, &de, &dn));
    int32_t *hc = nullptr, *hs = nullptr, *he = nullptr;
    int64_t *hn = nullptr;
    BR_TRY(ud.download(dc, n, keep, &hc)); BR_TRY(ud.download(ds, n, keep, &hs)); BR_TRY(ud.download(de, n, keep, &he)); BR_TRY(ud.download(dn, n, keep, &hn));
    BR_CUDA(cudaStreamSynchronize(s));
    fields.push_back(make_schema(contig_fmt, name_of(*L, L->key[0]), nullptr, 0));
    fields.push_back(make_schema("l", name_of(*L, L->key[1]), nullptr, 0));
    fields.push_back(make_schema("l", name_of(*L, L->key[2]), nullptr, 0));
    fields.push_back(make_schema("l", "n_intervals", nullptr, 0));
    if (n > 0) {
      std::unique_ptr<OwnedArray> top(new OwnedArray());
      StrBufs sb;
      BR_TRY(contig_buffers(hc, n, names, contig_large, &sb));
      top->kids.push_back(str_view(sb, n));
      ArrowArray a{};
      int rc = widened_column(hs, n, false, &a);
      if (rc == PBGPU_OK) { top->kids.push_back(a); rc = widened_column(he, n, false, &a); }
      if (rc == PBGPU_OK) { top->kids.push_back(a); top->kids.push_back(view_buffer(keep, hn, n)); }
      if (rc != PBGPU_OK) { for (auto &k : top->kids) if (k.release) k.release(&k); return rc; }
      push_batch(top, n);
    }
  } else if (op == PBGPU_OP_CLUSTER) {
    BR_TRY(ud.upload(kl, &dlc, &dls, &dle));
    const int64_t n = kl.n;
    int64_t *d_id = ud.dev.get<int64_t>((size_t)n);
    int32_t *d_cs = ud.dev.get<int32_t>((size_t)n), *d_ce = ud.dev.get<int32_t>((size_t)n);
    if (!d_id || !d_cs || !d_ce) return set_error(PBGPU_ENOMEM, "device allocation failed");
    int64_t n_clusters = 0;
    BR_TRY(pbgpu_cluster(dlc, dls, dle, n, n_contigs, o.filter_op, o.min_dist, d_id, d_cs, d_ce, &n_clusters, s));
    int64_t *h_id = nullptr;
    int32_t *h_cs = nullptr, *h_ce = nullptr;
    BR_TRY(ud.download(d_id, n, keep, &h_id)); BR_TRY(ud.download(d_cs, n, keep, &h_cs)); BR_TRY(ud.download(d_ce, n, keep, &h_ce));
    BR_CUDA(cudaStreamSynchronize(s));
    int64_t *w_cs = (int64_t *)hmalloc(8 * (size_t)(n ? n : 1)), *w_ce = (int64_t *)hmalloc(8 * (size_t)(n ? n : 1));
    if (!w_cs || !w_ce) { hfree(w_cs); hfree(w_ce); return set_error(PBGPU_ENOMEM, "host allocation failed"); }
    keep->v.push_back(w_cs); keep->v.push_back(w_ce);
    for (int64_t i = 0; i < n; ++i) { w_cs[i] = h_cs[i]; w_ce[i] = h_ce[i]; }
    for (int64_t c = 0; c < L->schema.n_children; ++c) fields.push_back(copy_schema(L->schema.children[c], name_of(*L, (int)c)));
    fields.push_back(make_schema("l", "cluster", nullptr, 0));
    fields.push_back(make_schema("l", "cluster_start", nullptr, 0));
    fields.push_back(make_schema("l", "cluster_end", nullptr, 0));
    // the input columns are re-exported zero-copy: one output batch per maximal run of rows that take part inside an
    // input batch (null-keyed rows, cluster id -1, are dropped; without them: one output batch per input batch)
    for (int b = 0; b < (int)L->batches.size(); ++b) {
      const int64_t g0 = L->start[(size_t)b], len = L->batches[(size_t)b].length;
      int64_t a = 0;
      while (a < len) {
        while (a < len && h_id[g0 + a] < 0) ++a;
        int64_t z = a;
        while (z < len && h_id[g0 + z] >= 0) ++z;
        if (z > a) {
          std::unique_ptr<OwnedArray> top(new OwnedArray());
          for (int c = 0; c < (int)L->n_cols(); ++c) top->kids.push_back(view_column(L, b, c, a, z - a));
          top->kids.push_back(view_buffer(keep, h_id + g0 + a, z - a));
          top->kids.push_back(view_buffer(keep, w_cs + g0 + a, z - a));
          top->kids.push_back(view_buffer(keep, w_ce + g0 + a, z - a));
          push_batch(top, z - a);
        }
        a = z;
      }
    }
  } else {  // subtract / complement: pieces of every "left" row that no "right" row covers
    const bool comp = op == PBGPU_OP_COMPLEMENT;
    const UnaryKeys &ka = comp ? kr : kl, &kb = comp ? kl : kr;  // complement = subtract with the view table on the left
    int32_t *dac = nullptr, *das = nullptr, *dae = nullptr, *dbc = nullptr, *dbs = nullptr, *dbe = nullptr;
    BR_TRY(ud.upload(ka, &dac, &das, &dae));
    BR_TRY(ud.upload(kb, &dbc, &dbs, &dbe));
    pbgpu_intervals *h = nullptr;
    BR_TRY(pbgpu_subtract(dac, das, dae, ka.n, dbc, dbs, dbe, kb.n, n_contigs, o.filter_op, s, &h));
    struct Free { pbgpu_intervals *h; cudaStream_t s; ~Free() { pbgpu_intervals_free(h, s); } } fr{h, s};
    const int64_t n = pbgpu_intervals_rows(h);
    const uint32_t *dr = nullptr;
    const int32_t *ds = nullptr, *de = nullptr;
    BR_TRY(pbgpu_intervals_columns(h, nullptr, &dr, &ds, &de, nullptr));
    uint32_t *hr = nullptr;
    int32_t *hs = nullptr, *he = nullptr;
    BR_TRY(ud.download(dr, n, keep, &hr)); BR_TRY(ud.download(ds, n, keep, &hs)); BR_TRY(ud.download(de, n, keep, &he));
    BR_CUDA(cudaStreamSynchronize(s));
    if (comp) {
      fields.push_back(make_schema(contig_fmt, name_of(*L, L->key[0]), nullptr, 0));
      fields.push_back(make_schema("l", name_of(*L, L->key[1]), nullptr, 0));
      fields.push_back(make_schema("l", name_of(*L, L->key[2]), nullptr, 0));
    } else {
      for (int c = 0; c < (int)L->n_cols(); ++c) {
        if (c == L->key[1] || c == L->key[2]) { fields.push_back(make_schema("l", name_of(*L, c), nullptr, 0)); continue; }
        bool ok = true;
        const std::string fmt = out_format(L->schema.children[c], &ok);
        if (!ok) return set_error(PBGPU_ESCHEMA, "payload column '%s' has unsupported Arrow type '%s' for subtract", name_of(*L, c).c_str(), L->schema.children[c]->format);
        fields.push_back(make_schema(fmt, name_of(*L, c), L->schema.children[c]->metadata, L->schema.children[c]->flags | ARROW_FLAG_NULLABLE));
      }
    }
    if (n > 0) {
      std::unique_ptr<OwnedArray> top(new OwnedArray());
      int rc = PBGPU_OK;
      ArrowArray a{};
      if (comp) {
        std::vector<int32_t> codes((size_t)n);
        for (int64_t i = 0; i < n; ++i) codes[(size_t)i] = ka.c[hr[i]];  // the view row's contig
        StrBufs sb;
        rc = contig_buffers(codes.data(), n, names, contig_large, &sb);
        if (rc == PBGPU_OK) { top->kids.push_back(str_view(sb, n)); rc = widened_column(hs, n, false, &a); }
        if (rc == PBGPU_OK) { top->kids.push_back(a); rc = widened_column(he, n, open_end, &a); }
        if (rc == PBGPU_OK) top->kids.push_back(a);
      } else {
        std::vector<GatherJob> jobs;
        for (int c = 0; c < (int)L->n_cols(); ++c)
          if (c != L->key[1] && c != L->key[2]) jobs.push_back(GatherJob{L.get(), c, hr});
        std::vector<ArrowArray> cols;
        if (!jobs.empty()) rc = gather_columns(jobs, n, &cols);
        size_t next = 0;
        for (int c = 0; c < (int)L->n_cols() && rc == PBGPU_OK; ++c) {
          if (c == L->key[1]) { rc = widened_column(hs, n, false, &a); if (rc == PBGPU_OK) top->kids.push_back(a); }
          else if (c == L->key[2]) { rc = widened_column(he, n, false, &a); if (rc == PBGPU_OK) top->kids.push_back(a); }
          else top->kids.push_back(cols[next++]);
        }
        for (; next < cols.size(); ++next) if (cols[next].release) cols[next].release(&cols[next]);
      }
      if (rc != PBGPU_OK) { for (auto &k : top->kids) if (k.release) k.release(&k); return rc; }
      push_batch(top, n);
    }
  }
  fguard.armed = false;
  vs->schema = make_schema("+s", "", nullptr, 0, std::move(fields));
  out->get_schema = vs_get_schema;
  out->get_next = vs_get_next;
  out->get_last_error = vs_last_error;
  out->release = vs_release;
  out->private_data = vs.release();
  return PBGPU_OK;
}
}  // namespace

namespace {
int check_opts(const PbRangeOptions *opts) {
  if (!opts) return set_error(PBGPU_EINVAL, "opts is NULL");
  if (opts->filter_op != PBGPU_FILTER_WEAK && opts->filter_op != PBGPU_FILTER_STRICT) return set_error(PBGPU_EINVAL, "bad filter_op %d", opts->filter_op);
  if (opts->range_op < PBGPU_OP_OVERLAP || opts->range_op > PBGPU_OP_MERGE)
    return set_error(PBGPU_EINVAL, "unknown range_op %d", opts->range_op);
  if (opts->output_mode < PBGPU_OUT_JOIN || opts->output_mode > PBGPU_OUT_LEFT_DISTINCT) return set_error(PBGPU_EINVAL, "bad output_mode %d", opts->output_mode);
  return PBGPU_OK;
}
bool is_unary_op(int op) { return op == PBGPU_OP_MERGE || op == PBGPU_OP_CLUSTER || op == PBGPU_OP_COMPLEMENT || op == PBGPU_OP_SUBTRACT; }
std::unique_ptr<OutStream> new_outstream(const PbRangeOptions &opts, const std::string *sfx1, const std::string *sfx2) {
  std::unique_ptr<OutStream> os(new OutStream());
  os->opt = opts;
  if (sfx1) os->suffix1 = *sfx1;
  if (sfx2) os->suffix2 = *sfx2;
  os->opt.suffixes[0] = os->opt.suffixes[1] = nullptr;
  os->opt.cols1[0] = os->opt.cols1[1] = os->opt.cols1[2] = nullptr;  // caller strings are not retained
  os->opt.cols2[0] = os->opt.cols2[1] = os->opt.cols2[2] = nullptr;
  if (opts.max_batch_rows) os->batch_rows = opts.max_batch_rows;
  os->batch_rows = (os->batch_rows + 7u) & ~7u;  // keep bitmap bytes chunk-private
  return os;
}
void publish(ArrowArrayStream *out, std::unique_ptr<OutStream> os) {
  out->get_schema = out_get_schema;
  out->get_next = out_get_next;
  out->get_last_error = out_last_error;
  out->release = out_release;
  out->private_data = os.release();
}
}  // namespace

// ---- build once, probe many (include/pbgpu.h: pbgpu_range_open / _probe / _close) ----------------------------------------
struct pbgpu_range_session {
  PbRangeOptions opt{};
  std::string cols1[3], cols2[3], suffix1 = "_1", suffix2 = "_2";
  bool iter_is_left = true;
  std::shared_ptr<Table> indexed;
  std::shared_ptr<IndexSide> xs;
};

extern "C" int pbgpu_range_open(struct ArrowArrayStream *indexed, const PbRangeOptions *opts, pbgpu_range_session **out) {
  struct Releaser { ArrowArrayStream *s; ~Releaser() { if (s && s->release) s->release(s); } } rl{indexed};
  if (!out) return set_error(PBGPU_EINVAL, "out is NULL");
  *out = nullptr;
  int rc = check_opts(opts);
  if (rc != PBGPU_OK) return rc;
  if (is_unary_op(opts->range_op)) return set_error(PBGPU_EINVAL, "range_op %d is a unary sweep: use pbgpu_range_op", opts->range_op);
  try {
    std::unique_ptr<pbgpu_range_session> ss(new pbgpu_range_session());
    ss->opt = *opts;
    for (int i = 0; i < 3; ++i) {
      if (!opts->cols1[i] || !opts->cols2[i]) return set_error(PBGPU_EINVAL, "NULL column name");
      ss->cols1[i] = opts->cols1[i];
      ss->cols2[i] = opts->cols2[i];
    }
    if (opts->suffixes[0]) ss->suffix1 = opts->suffixes[0];
    if (opts->suffixes[1]) ss->suffix2 = opts->suffixes[1];
    ss->iter_is_left = (opts->range_op == PBGPU_OP_OVERLAP || opts->range_op == PBGPU_OP_NEAREST);
    ss->indexed = std::make_shared<Table>();
    const char *side = ss->iter_is_left ? "right" : "left";
    rc = drain(indexed, *ss->indexed, side);
    for (int i = 0; i < 3 && rc == PBGPU_OK; ++i)
      rc = find_col(*ss->indexed, (ss->iter_is_left ? ss->cols2[i] : ss->cols1[i]).c_str(), side, &ss->indexed->key[i]);
    if (rc == PBGPU_OK) rc = prepare_index(ss->opt, ss->indexed, side, &ss->xs);
    if (rc != PBGPU_OK) return rc;
    *out = ss.release();
    return PBGPU_OK;
  } catch (const std::bad_alloc &) {
    return set_error(PBGPU_ENOMEM, "host allocation failed");
  } catch (const std::exception &e) {
    return set_error(PBGPU_EINVAL, "internal error: %s", e.what());
  }
}

extern "C" int pbgpu_range_probe(pbgpu_range_session *ss, struct ArrowArrayStream *iterated, struct ArrowArrayStream *out) {
  struct Releaser { ArrowArrayStream *s; ~Releaser() { if (s && s->release) s->release(s); } } rl{iterated};
  if (!ss || !out) return set_error(PBGPU_EINVAL, "session/out is NULL");
  try {
    std::unique_ptr<OutStream> os = new_outstream(ss->opt, &ss->suffix1, &ss->suffix2);
    auto it = std::make_shared<Table>();
    const char *side = ss->iter_is_left ? "left" : "right";
    int rc = drain(iterated, *it, side);
    for (int i = 0; i < 3 && rc == PBGPU_OK; ++i)
      rc = find_col(*it, (ss->iter_is_left ? ss->cols1[i] : ss->cols2[i]).c_str(), side, &it->key[i]);
    if (rc != PBGPU_OK) return rc;
    os->left = ss->iter_is_left ? it : ss->indexed;
    os->right = ss->iter_is_left ? ss->indexed : it;
    rc = run_iter(os->left.get(), os->right.get(), os.get(), ss->xs);
    if (rc != PBGPU_OK) return rc;
    publish(out, std::move(os));
    return PBGPU_OK;
  } catch (const std::bad_alloc &) {
    return set_error(PBGPU_ENOMEM, "host allocation failed");
  } catch (const std::exception &e) {
    return set_error(PBGPU_EINVAL, "internal error: %s", e.what());
  }
}

extern "C" void pbgpu_range_close(pbgpu_range_session *ss) { delete ss; }

extern "C" void pbgpu_pinned_stats(uint64_t *busy_bytes, uint64_t *peak_bytes, int reset_peak) {
  std::lock_guard<std::mutex> lk(g_pin_mu);
  if (busy_bytes) *busy_bytes = g_pin_busy;
  if (peak_bytes) *peak_bytes = g_pin_peak;
  if (reset_peak) g_pin_peak = g_pin_busy;
}

extern "C" int pbgpu_range_op(struct ArrowArrayStream *left, struct ArrowArrayStream *right, const PbRangeOptions *opts,
                              struct ArrowArrayStream *out) {
  // The inputs are moved: whatever happens below, both are released exactly once before returning.
  struct Releaser { ArrowArrayStream *s; ~Releaser() { if (s && s->release) s->release(s); } } rl{left}, rr{right};
  if (!out) return set_error(PBGPU_EINVAL, "opts/out is NULL");
  int rc0 = check_opts(opts);
  if (rc0 != PBGPU_OK) return rc0;
  try {
    if (is_unary_op(opts->range_op)) {  // merge / cluster / complement / subtract: `right` is the second table (subtract), the
      auto L = std::make_shared<Table>();  // view table (complement; may be NULL) or NULL
      std::shared_ptr<Table> R;
      int rc = drain(left, *L, "left");
      for (int i = 0; i < 3 && rc == PBGPU_OK; ++i) rc = find_col(*L, opts->cols1[i], "left", &L->key[i]);
      const bool want_right = right && right->get_schema && (opts->range_op == PBGPU_OP_SUBTRACT || opts->range_op == PBGPU_OP_COMPLEMENT);
      if (rc == PBGPU_OK && want_right) {
        R = std::make_shared<Table>();
        rc = drain(right, *R, "right");
        for (int i = 0; i < 3 && rc == PBGPU_OK; ++i) rc = find_col(*R, opts->cols2[i], "right", &R->key[i]);
      }
      if (rc != PBGPU_OK) return rc;
      return run_unary(*opts, L, R, out);
    }
    if (!right) return set_error(PBGPU_EINVAL, "right stream is NULL");
    std::string sfx1 = opts->suffixes[0] ? opts->suffixes[0] : "_1", sfx2 = opts->suffixes[1] ? opts->suffixes[1] : "_2";
    std::unique_ptr<OutStream> os = new_outstream(*opts, &sfx1, &sfx2);
    os->left = std::make_shared<Table>();
    os->right = std::make_shared<Table>();
    Trace tr;
    int rc = drain(left, *os->left, "left");
    if (rc == PBGPU_OK) rc = drain(right, *os->right, "right");
    tr.lap("drain input streams");
    for (int i = 0; i < 3 && rc == PBGPU_OK; ++i) rc = find_col(*os->left, opts->cols1[i], "left", &os->left->key[i]);
    for (int i = 0; i < 3 && rc == PBGPU_OK; ++i) rc = find_col(*os->right, opts->cols2[i], "right", &os->right->key[i]);
    if (rc != PBGPU_OK) return rc;
    rc = run(os->left.get(), os->right.get(), os.get());
    if (rc != PBGPU_OK) return rc;
    publish(out, std::move(os));
    return PBGPU_OK;
  } catch (const std::bad_alloc &) {
    return set_error(PBGPU_ENOMEM, "host allocation failed");
  } catch (const std::exception &e) {
    return set_error(PBGPU_EINVAL, "internal error: %s", e.what());
  }
}
