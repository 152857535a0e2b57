// pbgpu.cu -- device-level C ABI of libpbgpu.so (see include/pbgpu.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 (see __graft_entry__.build()).
#include <stdarg.h>

#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"
#include "index.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "sweep.cuh"
#include "bins.cuh"

namespace pbgpu {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static std::once_flag g_pool_once[64];
static void tune_pool(int dev) {
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;  // keep freed scratch cached in the pool between calls
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
}

// ---- device block cache ---------------------------------------------------------------------------------------
// cudaMallocAsync splits and re-merges the blocks of its pool.  A build over 90 M rows asks for a dozen blocks of
// 0.3-3.6 GB in varying order; the pool then keeps mapping fresh physical memory (30-360 ms per block, measured:
// profiles/r01x) although the same sizes were freed a moment ago.  Blocks of >= 64 KB are therefore never handed
// back on free: they are parked, whole, with an event recorded on the freeing stream, and the next request of the
// same size class gets one of them -- stream-ordered like cudaMallocAsync itself: reuse on the freeing stream
// needs nothing, reuse on another stream first waits for the event.  Size classes are 1/16-octave steps, so the
// sizes of one workload map onto themselves call after call and a steady state makes no allocator calls at all.
// Parked bytes are capped (PBGPU_DEV_CACHE_MB, default 1/3 of the device); on overflow or on an allocation
// failure the oldest parked blocks go back to the pool.  Smaller requests use the pool directly.
struct ParkedBlock { void *p; size_t bytes; cudaStream_t stream; cudaEvent_t ev; int dev; uint64_t age; };
struct BlockCache {
  std::mutex mu;
  std::vector<ParkedBlock> parked;
  std::vector<std::pair<void *, size_t>> live;  // blocks of cacheable size handed out: pointer -> class size
  std::vector<std::pair<int, cudaEvent_t>> spare_events;  // (device, event)
  size_t parked_bytes = 0, cap = 0;
  uint64_t clock = 0;
  bool enabled = true;
  BlockCache() {
    const char *e = getenv("PBGPU_DEV_CACHE_MB");
    if (e) { const long long mb = atoll(e); if (mb <= 0) enabled = false; else cap = (size_t)mb << 20; }
  }
};
static BlockCache &block_cache() { static BlockCache *c = new BlockCache(); return *c; }  // never destroyed: no CUDA calls at exit
constexpr size_t kCacheMinBytes = 64 << 10;
static size_t size_class(size_t b) {
  int hb = 63 - __builtin_clzll((unsigned long long)b);
  const size_t step = (size_t)1 << (hb > 4 ? hb - 4 : 0);
  return (b + step - 1) & ~(step - 1);
}
// give parked blocks back to the pool, oldest first, until at most `keep` bytes stay parked (mu held)
static void cache_release_locked(BlockCache &c, int dev, size_t keep) {
  while (c.parked_bytes > keep && !c.parked.empty()) {
    size_t k = 0;
    for (size_t i = 1; i < c.parked.size(); ++i) if (c.parked[i].age < c.parked[k].age) k = i;
    ParkedBlock b = c.parked[k];
    c.parked[k] = c.parked.back();
    c.parked.pop_back();
    c.parked_bytes -= b.bytes;
    int cur = dev;
    if (b.dev != cur) cudaSetDevice(b.dev);
    if (b.ev) { cudaEventSynchronize(b.ev); c.spare_events.emplace_back(b.dev, b.ev); }
    cudaFreeAsync(b.p, cudaStreamPerThread);  // its last use has completed: any stream will do
    if (b.dev != cur) cudaSetDevice(cur);
  }
}

int dev_alloc(void **p, size_t bytes, cudaStream_t s) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64) std::call_once(g_pool_once[dev], tune_pool, dev);
  if (!bytes) bytes = 1;
  BlockCache &c = block_cache();
  const bool cacheable = c.enabled && bytes >= kCacheMinBytes;
  if (cacheable) {
    bytes = size_class(bytes);
    std::lock_guard<std::mutex> lk(c.mu);
    int hit = -1;
    for (size_t i = 0; i < c.parked.size(); ++i) {
      const ParkedBlock &b = c.parked[i];
      if (b.dev != dev || b.bytes != bytes) continue;
      if (b.stream == s) { hit = (int)i; break; }  // same stream: plain stream order, no wait
      if (hit < 0) hit = (int)i;
    }
    if (hit >= 0) {
      ParkedBlock b = c.parked[hit];
      c.parked[hit] = c.parked.back();
      c.parked.pop_back();
      c.parked_bytes -= b.bytes;
      if (b.ev) {
        if (b.stream != s && cudaStreamWaitEvent(s, b.ev, 0) != cudaSuccess) { cudaGetLastError(); cudaEventSynchronize(b.ev); }
        c.spare_events.emplace_back(b.dev, b.ev);
      }
      c.live.emplace_back(b.p, b.bytes);
      *p = b.p;
      return PBGPU_OK;
    }
  }
  cudaError_t e = cudaMallocAsync(p, bytes, s);
  if (e == cudaErrorMemoryAllocation && c.enabled) {  // hand every parked block back and try once more
    cudaGetLastError();
    {
      std::lock_guard<std::mutex> lk(c.mu);
      cache_release_locked(c, dev, 0);
    }
    cudaStreamSynchronize(cudaStreamPerThread);
    e = cudaMallocAsync(p, bytes, s);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(e == cudaErrorMemoryAllocation ? PBGPU_ENOMEM : PBGPU_ECUDA, "cudaMallocAsync(%zu) failed: %s", bytes,
                     cudaGetErrorString(e));
  }
  if (cacheable) {
    std::lock_guard<std::mutex> lk(c.mu);
    c.live.emplace_back(*p, bytes);
  }
  return PBGPU_OK;
}
void dev_free(void *p, cudaStream_t s) {
  if (!p) return;
  BlockCache &c = block_cache();
  if (c.enabled) {
    std::lock_guard<std::mutex> lk(c.mu);
    for (size_t i = c.live.size(); i-- > 0;) {
      if (c.live[i].first != p) continue;
      const size_t bytes = c.live[i].second;
      c.live[i] = c.live.back();
      c.live.pop_back();
      int dev = 0;
      cudaGetDevice(&dev);
      if (!c.cap) {
        size_t fr = 0, tot = 0;
        c.cap = cudaMemGetInfo(&fr, &tot) == cudaSuccess ? tot / 3 : ((size_t)16 << 30);
      }
      cudaEvent_t ev = nullptr;
      for (size_t k = c.spare_events.size(); k-- > 0;) {
        if (c.spare_events[k].first != dev) continue;
        ev = c.spare_events[k].second;
        c.spare_events[k] = c.spare_events.back();
        c.spare_events.pop_back();
        break;
      }
      if (!ev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); ev = nullptr; }
      if (!ev || cudaEventRecord(ev, s) != cudaSuccess) {  // cannot order a later reuse: ordinary free
        cudaGetLastError();
        if (ev) c.spare_events.emplace_back(dev, ev);
        cudaFreeAsync(p, s);
        return;
      }
      c.parked.push_back(ParkedBlock{p, bytes, s, ev, dev, ++c.clock});
      c.parked_bytes += bytes;
      if (c.parked_bytes > c.cap) cache_release_locked(c, dev, c.cap - c.cap / 4);
      return;
    }
  }
  cudaFreeAsync(p, s);
}

// CUDA-event stage marks, always on (an event record is ~1 us of host time and no device sync): pairs of
// events bracket the index build and each provider kernel on the launching stream; pbgpu_last_stage_times()
// turns them into durations after the fact.  One set per host thread and device.
enum { EV_BUILD0, EV_BUILD1, EV_COUNT0, EV_COUNT1, EV_P1_0, EV_P1_1, EV_SCAN1, EV_EMIT0, EV_EMIT1, EV_BIN0, EV_BIN1, EV_UNBIN0, EV_UNBIN1, EV_N };
struct StageEvents {
  cudaEvent_t ev[EV_N] = {};
  bool set[EV_N] = {};
  int device = -1;
  void mark(int which, cudaStream_t s) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != device) {
      for (int i = 0; i < EV_N; ++i) { if (ev[i]) cudaEventDestroy(ev[i]); ev[i] = nullptr; set[i] = false; }
      device = dev;
    }
    if (!ev[which] && cudaEventCreate(&ev[which]) != cudaSuccess) { cudaGetLastError(); ev[which] = nullptr; return; }
    set[which] = cudaEventRecord(ev[which], s) == cudaSuccess;
  }
  uint64_t span_ns(int a, int b) {
    if (!set[a] || !set[b]) return 0;
    if (cudaEventSynchronize(ev[b]) != cudaSuccess) { cudaGetLastError(); return 0; }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ev[a], ev[b]) != cudaSuccess) { cudaGetLastError(); return 0; }
    return (uint64_t)(ms * 1e6f);
  }
};
static thread_local StageEvents g_ev;

// ---- host scalars without a copy engine round trip ------------------------------------------------------------
// The build needs two small device results on the host (coordinate statistics; nested? + axis span) and pass 1 one
// (the pair total).  cudaMemcpyAsync into pageable memory + cudaStreamSynchronize costs 10-20 us of idle GPU per
// scalar; instead a one-thread kernel posts the words into a page-locked, device-mapped mailbox followed by a
// sequence number and the host spins on that number (the write is visible ~1-2 us after it is issued).  If the
// number does not show up within a bounded spin the stream is synchronised the ordinary way, which also surfaces
// any kernel fault.  One mailbox per host thread.  PBGPU_SYNC=memcpy keeps the copy-engine path (A/B runs).
constexpr int kMailboxWords = 16;  // word 15 = sequence number
__global__ void mailbox_post_kernel(const unsigned long long *__restrict__ src, int n_words, volatile unsigned long long *mb,
                                    unsigned long long seq) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_words; ++i) mb[i] = src[i];
    __threadfence_system();
    mb[kMailboxWords - 1] = seq;
  }
}
struct Mailbox {
  volatile unsigned long long *h = nullptr;
  unsigned long long *d = nullptr;
  unsigned long long seq = 0;
  bool tried = false;
  ~Mailbox() { if (h) cudaFreeHost((void *)h); }
};
static thread_local Mailbox g_mailbox;
static bool sync_by_memcpy() {
  static bool v = [] { const char *e = getenv("PBGPU_SYNC"); return e && !strcmp(e, "memcpy"); }();
  return v;
}
struct MailboxSlot { unsigned long long *d; unsigned long long seq; };
// next sequence number of the calling thread's mailbox; d == nullptr when there is none (allocation failed / disabled)
MailboxSlot mailbox_open() {
  Mailbox &mb = g_mailbox;
  if (!mb.tried && !sync_by_memcpy()) {
    mb.tried = true;
    void *hp = nullptr, *dp = nullptr;
    if (cudaHostAlloc(&hp, sizeof(unsigned long long) * kMailboxWords, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
        cudaHostGetDevicePointer(&dp, hp, 0) == cudaSuccess) {
      memset(hp, 0, sizeof(unsigned long long) * kMailboxWords);
      mb.h = (volatile unsigned long long *)hp;
      mb.d = (unsigned long long *)dp;
    } else {
      cudaGetLastError();
      if (hp) cudaFreeHost(hp);
    }
  }
  if (!mb.h) return MailboxSlot{nullptr, 0};
  return MailboxSlot{mb.d, ++mb.seq};
}
// waits until the device has posted sequence number slot.seq (by a kernel enqueued on s), then copies n_words out
int mailbox_wait(const MailboxSlot &slot, int n_words, unsigned long long *out, cudaStream_t s) {
  Mailbox &mb = g_mailbox;
  const unsigned long long seq = slot.seq;
  bool arrived = false;
  for (long spin = 0; spin < 20000000L; ++spin) {  // ~ tens of ms at most
    if (mb.h[kMailboxWords - 1] == seq) { arrived = true; break; }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  if (!arrived) {
    PB_CUDA(cudaStreamSynchronize(s));  // long-running predecessor or a fault: wait the ordinary way
    if (mb.h[kMailboxWords - 1] != seq) return set_error(PBGPU_ECUDA, "device result did not arrive in the host mailbox");
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  for (int i = 0; i < n_words; ++i) out[i] = mb.h[i];
  return PBGPU_OK;
}
// Per-thread state (host mailbox, stage events) of a short-lived helper thread, released before it exits: the Arrow level
// runs index builds on one (arrow_bridge.cpp); long-lived caller threads keep theirs for the life of the process.
void release_thread_state() {
  for (int i = 0; i < EV_N; ++i) {
    if (g_ev.ev[i]) cudaEventDestroy(g_ev.ev[i]);
    g_ev.ev[i] = nullptr;
    g_ev.set[i] = false;
  }
  g_ev.device = -1;
  Mailbox &mb = g_mailbox;
  if (mb.h) cudaFreeHost((void *)mb.h);
  mb.h = nullptr;
  mb.d = nullptr;
  mb.tried = false;
}
// copies n_words (<= 15) 64-bit words from device memory to `out`; returns after they have arrived
int fetch_words(const void *d_src, int n_words, unsigned long long *out, cudaStream_t s) {
  const MailboxSlot slot = n_words <= kMailboxWords - 1 ? mailbox_open() : MailboxSlot{nullptr, 0};
  if (!slot.d) {
    PB_CUDA(cudaMemcpyAsync(out, d_src, sizeof(unsigned long long) * (size_t)n_words, cudaMemcpyDeviceToHost, s));
    PB_CUDA(cudaStreamSynchronize(s));
    return PBGPU_OK;
  }
  PB_LAUNCH(mailbox_post_kernel, 1, 32, 0, s, (const unsigned long long *)d_src, n_words, slot.d, slot.seq);
  PB_CHECK_LAUNCH();
  return mailbox_wait(slot, n_words, out, s);
}

}  // namespace pbgpu

using namespace pbgpu;

extern "C" {

const char *pbgpu_last_error(void) { return g_err; }
const char *pbgpu_version(void) { return "pbgpu 0.1.0 (sm_100a)"; }
uint64_t pbgpu_launch_count(void) { return g_launches.load(); }

int pbgpu_device_count(int *count) {
  if (!count) return set_error(PBGPU_EINVAL, "count is NULL");
  PB_CUDA(cudaGetDeviceCount(count));
  return PBGPU_OK;
}

int pbgpu_last_stage_times(pbgpu_stage_times *out) {
  if (!out) return set_error(PBGPU_EINVAL, "out is NULL");
  out->partition_sort_ns = g_ev.span_ns(EV_BUILD0, EV_BUILD1);
  out->count_ns = g_ev.span_ns(EV_P1_0, EV_P1_1);
  out->scan_ns = g_ev.span_ns(EV_P1_1, EV_SCAN1);
  out->emit_ns = g_ev.span_ns(EV_EMIT0, EV_EMIT1);
  out->count_overlaps_ns = g_ev.span_ns(EV_COUNT0, EV_COUNT1);
  out->bin_ns = g_ev.span_ns(EV_BIN0, EV_BIN1);
  out->unbin_ns = g_ev.span_ns(EV_UNBIN0, EV_UNBIN1);
  return PBGPU_OK;
}

// ---------------------------------------------------------------------------------------------
void pbgpu_index_free(pbgpu_index *ix) {
  if (!ix) return;
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != ix->device) cudaSetDevice(ix->device);
  // stream-ordered free on the legacy default stream: ordered after everything blocking streams have queued;
  // users of non-blocking streams synchronise before freeing (pbgpu.h)
  dev_free(ix->slab, 0);
  dev_free(ix->slab2, 0);
  dev_free(ix->slab_n, 0);
  dev_free(ix->slab_e, 0);
  if (ix->end_ready) cudaEventDestroy(ix->end_ready);
  if (cur != ix->device) cudaSetDevice(cur);
  delete ix;
}

// release an index in stream order on `s` (the stream its last reader was enqueued on)
static void index_free_on(pbgpu_index *ix, cudaStream_t s) {
  if (!ix) return;
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != ix->device) cudaSetDevice(ix->device);
  dev_free(ix->slab, s);
  dev_free(ix->slab2, s);
  dev_free(ix->slab_n, s);
  dev_free(ix->slab_e, s);
  if (ix->end_ready) cudaEventDestroy(ix->end_ready);
  if (cur != ix->device) cudaSetDevice(cur);
  delete ix;
}
void pbgpu_index_free_async(pbgpu_index *ix, void *stream) { index_free_on(ix, (cudaStream_t)stream); }

int64_t pbgpu_index_rows(const pbgpu_index *ix) { return ix ? ix->m : 0; }
size_t pbgpu_index_bytes(const pbgpu_index *ix) { return ix ? ix->bytes : 0; }

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// PBGPU_TRACE_BUILD=1: host wall time of every build stage on stderr, with the stream drained at each lap (tuning aid;
// changes the timing it reports on by serialising host and device)
struct BuildTrace {
  bool on;
  cudaStream_t s;
  double t0;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
  explicit BuildTrace(cudaStream_t st) : s(st) {
    const char *e = getenv("PBGPU_TRACE_BUILD");
    on = e && e[0] == '1';
    t0 = now();
  }
  void lap(const char *what, bool drain = true) {
    if (!on) return;
    const double t1 = now();
    double t2 = t1;
    if (drain) { cudaStreamSynchronize(s); t2 = now(); }
    fprintf(stderr, "[pbgpu build] %-34s host %8.3f ms  + drain %8.3f ms\n", what, t1 - t0, t2 - t1);
    t0 = now();
  }
};

// PBGPU_JDIR_SHIFT=k: k more bits per directory bucket than the 0.6-1.2-rows-per-bucket rule picks.  Default: 1 when the
// directory would otherwise exceed the L2 by far (1.2-2.4 rows per bucket, ~6 of a record's 12 key slots used: half the
// directory bytes to build and to stream through the L2 for the same count-kernel time; r2f: build -0.75 ms on 90 M rows),
// 0 for directories that stay L2-resident (r2h: config 2's count kernels lose 30 % with the fuller records: more of them
// are crowded and take the search path).
static int jdir_extra_shift(unsigned long long n_buckets) {
  static int v = [] { const char *e = getenv("PBGPU_JDIR_SHIFT"); return e ? atoi(e) : -1; }();
  if (v >= 0) return v;
  return n_buckets * sizeof(JRec) > ((unsigned long long)192 << 20) ? 1 : 0;
}
constexpr int kCrowdedLinearMax = 64;  // crowded records of an unsorted-ends index are counted linearly up to this many ends

namespace pbgpu {
__global__ void widen_word_kernel(const unsigned int *__restrict__ in, unsigned long long *__restrict__ out) {
  if (threadIdx.x == 0) *out = *in;
}
__global__ void __launch_bounds__(256) widen_u32_keys_kernel(const uint32_t *__restrict__ in, int64_t n, uint64_t *__restrict__ k, uint64_t *__restrict__ v) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) { k[i] = in[i]; v[i] = 0; }
}
__global__ void __launch_bounds__(256) narrow_u32_keys_kernel(const uint64_t *__restrict__ k, int64_t n, uint32_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) out[i] = (uint32_t)k[i];
}
}  // namespace pbgpu

// End order of an index with nested intervals: ends sorted by (contig, end, start, row) and their positions in start
// order (nearest's upstream walk, the generic kernels' rank identity).  Stable sort over the start order, so ties keep
// (start, row) order.  Enqueued on `s`; the arrays live in their own slab.
static int build_end_order(pbgpu_index *ix, cudaStream_t s) {
  const int64_t m = ix->m;
  if (!ix->nested || ix->en_sorted || m == 0) return PBGPU_OK;
  const size_t arr_b = (sizeof(int32_t) * (size_t)m + 255) & ~(size_t)255;
  PB_TRY(dev_alloc(&ix->slab_e, 2 * arr_b, s));
  ix->bytes += 2 * arr_b;
  int32_t *en_sorted = (int32_t *)ix->slab_e;
  uint32_t *en_pos = (uint32_t *)((char *)ix->slab_e + arr_b);
  Scratch sc(s);
  uint64_t *ek = nullptr, *ev = nullptr, *ek2 = nullptr, *ev2 = nullptr;
  PB_TRY(sc.get(&ek, (size_t)m));
  PB_TRY(sc.get(&ev, (size_t)m));
  PB_TRY(sc.get(&ek2, (size_t)m));
  PB_TRY(sc.get(&ev2, (size_t)m));
  constexpr int pos_bits = 32;
  constexpr uint32_t bias = 0x80000000u;
  PB_LAUNCH(make_end_keys_seg_kernel, (unsigned)cdiv(m, 256), 256, 0, s, ix->seg, ix->n_contigs, ix->en, m, pos_bits, bias, ek, ev);
  PB_CHECK_LAUNCH();
  int epos[kRsMaxPasses], ne = 0;
  const int vb = bit_length_u32((uint32_t)ix->min_end ^ (uint32_t)ix->max_end);
  for (int p = 0; p * 8 < vb; ++p) epos[ne++] = p;
  const int contig_digits = (bit_length_u32((uint32_t)ix->n_contigs) + 7) / 8;
  if (ix->n_contigs > 1) for (int q = 0; q < contig_digits; ++q) epos[ne++] = 4 + q;
  SortedPairs es;
  PB_TRY(radix_sort_digits(ek, ev, ek2, ev2, m, epos, ne, nullptr, s, &es));
  PB_LAUNCH(unpack_ends_kernel, (unsigned)cdiv(m, 256), 256, 0, s, es.keys, es.vals, m, pos_bits, bias, en_sorted, en_pos);
  PB_CHECK_LAUNCH();
  ix->en_sorted = en_sorted;
  ix->en_pos = en_pos;
  return PBGPU_OK;
}
// nearest on a fast-path index: the end order is built by the first call that needs it (on that call's stream); later
// calls on other streams wait for the event
static int ensure_end_order(const pbgpu_index *cix, cudaStream_t s) {
  pbgpu_index *ix = const_cast<pbgpu_index *>(cix);
  if (!ix->nested) return PBGPU_OK;
  std::lock_guard<std::mutex> lk(ix->mu);
  if (!ix->en_sorted) {
    PB_TRY(build_end_order(ix, s));
    if (!ix->end_ready && cudaEventCreateWithFlags(&ix->end_ready, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); ix->end_ready = nullptr; }
    if (ix->end_ready) PB_CUDA(cudaEventRecord(ix->end_ready, s));
    else PB_CUDA(cudaStreamSynchronize(s));
    ix->end_stream = s;
  } else if (ix->end_ready && s != ix->end_stream) {
    PB_CUDA(cudaStreamWaitEvent(s, ix->end_ready, 0));
  }
  return PBGPU_OK;
}

// PBGPU_BUILD=generic: always the generic front half (16-byte pairs, one pass per varying key byte); default: the
// global-key front half whenever it applies (A/B runs; the parity tests run both)
static bool build_generic_only() {
  static bool v = [] { const char *e = getenv("PBGPU_BUILD"); return e && !strcmp(e, "generic"); }();
  return v;
}

// what the two front halves of the build hand to the common back half
struct BuildFront {
  bool nested = false;
  unsigned long long total_span = 0;   // 0: no fast path (not computed / does not apply)
  long long max_len = 0;
  ContigMap *d_cmap = nullptr;         // device, n_contigs entries (scratch), offsets final
  const uint64_t *keys = nullptr;      // generic: the sorted (contig | start) keys (contig of position i); gkey: NULL (gs is ready)
  void *idle = nullptr;                // scratch of at least 8 bytes per input row, free for reuse
  uint64_t *idle2 = nullptr;           // a second such buffer (generic) or NULL
  bool gs_ready = false;
};

static void set_slab1(pbgpu_index *ix, int32_t n_contigs, int64_t m, size_t *arr_b_out) {
  const size_t seg_b = align_up(sizeof(int32_t) * ((size_t)n_contigs + 2)), arr_b = align_up(sizeof(int32_t) * (size_t)(m ? m : 1));
  char *base = (char *)ix->slab;
  ix->seg = (int32_t *)base;
  ix->st = (int32_t *)(base + seg_b);
  ix->en = (int32_t *)(base + seg_b + arr_b);
  ix->row = (uint32_t *)(base + seg_b + 2 * arr_b);
  ix->er = (uint2 *)(base + seg_b + 3 * arr_b);
  ix->pmax = ix->en;       // until nested intervals are detected: running max == end,
  ix->en_sorted = ix->en;  // end order == start order,
  ix->en_pos = nullptr;    // identity
  *arr_b_out = arr_b;
}
static size_t slab1_bytes(int32_t n_contigs, int64_t m) {
  return align_up(sizeof(int32_t) * ((size_t)n_contigs + 2)) + 3 * align_up(sizeof(int32_t) * (size_t)(m ? m : 1)) +
         align_up(sizeof(uint2) * (size_t)(m ? m : 1));
}
// bucket width of the rank directory: the smallest power of two that leaves at most m/0.6 buckets (0.6-1.2 indexed rows
// per bucket: a record then holds ~3 keys on average of its 12 and crowded records stay below ~0.2 %), capped at 2^13
static int jdir_shift_for(unsigned long long total_span, int64_t m) {
  int shift = 0;
  while (shift < kJMaxShift && (total_span >> shift) > ((unsigned long long)m * 5ull) / 3ull) ++shift;
  shift += jdir_extra_shift(total_span >> shift);
  return shift > kJMaxShift ? kJMaxShift : shift;
}
static int alloc_slab2(pbgpu_index *ix, unsigned long long total_span, int64_t m, int32_t n_contigs, cudaStream_t s) {
  const int shift = jdir_shift_for(total_span, m);
  const uint32_t nb = (uint32_t)(total_span >> shift) + 1;
  const size_t cm_b = align_up(sizeof(ContigMap) * ((size_t)n_contigs + 1)), g_b = align_up(4 * (size_t)m);
  const size_t d_b = align_up(sizeof(JRec) * ((size_t)nb + 1));
  const size_t cm32_b = align_up(sizeof(ContigMap32) * ((size_t)n_contigs + 1));
  PB_TRY(dev_alloc(&ix->slab2, cm_b + 2 * g_b + d_b + cm32_b, s));
  ix->bytes += cm_b + 2 * g_b + d_b + cm32_b;
  char *b2 = (char *)ix->slab2;
  ix->cmap = (ContigMap *)b2;
  ix->gs = (uint32_t *)(b2 + cm_b);
  ix->ge = (uint32_t *)(b2 + cm_b + g_b);
  ix->jdir = (JRec *)(b2 + cm_b + 2 * g_b);
  ix->cmap32 = (ContigMap32 *)(b2 + cm_b + 2 * g_b + d_b);
  ix->shift = shift;
  ix->n_buckets = nb;
  ix->axis_span = (uint32_t)total_span;
  return PBGPU_OK;
}

// ---- front half, global-key variant (index.cuh): returns *applies = false when the table needs the generic one ------
static int build_front_gkey(pbgpu_index *ix, const int32_t *d_c, const int32_t *d_s, const int32_t *d_e, const uint32_t *d_ids, int64_t m_in,
                            int32_t n_contigs, cudaStream_t s, Scratch &sc, BuildTrace &bt, BuildFront *fr, bool *applies) {
  *applies = false;
  int32_t *cmin = nullptr, *cmax = nullptr;
  GStats *d_st = nullptr;
  unsigned long long *d_words = nullptr;
  ContigMap *d_cmap = nullptr;
  PB_TRY(sc.get(&cmin, (size_t)n_contigs + 1));
  PB_TRY(sc.get(&cmax, (size_t)n_contigs + 1));
  PB_TRY(sc.get(&d_st, 1));
  PB_TRY(sc.get(&d_words, 8));
  PB_TRY(sc.get(&d_cmap, (size_t)n_contigs + 1));
  PB_LAUNCH(gstats_init_kernel, (unsigned)cdiv(n_contigs > 0 ? n_contigs : 1, 256), 256, 0, s, cmin, cmax, n_contigs, d_st);
  int64_t grid = cdiv(m_in, 512 * 2);
  if (grid > kSMs * 4) grid = kSMs * 4;
  PB_LAUNCH(contig_stats_kernel, (unsigned)grid, 512, 0, s, d_c, d_s, d_e, m_in, n_contigs, cmin, cmax, d_st);
  const MailboxSlot slot = mailbox_open();
  PB_LAUNCH(contig_layout_mm_kernel, 1, 1024, 0, s, cmin, cmax, d_st, n_contigs, d_cmap, d_words, slot.d, slot.seq);
  PB_CHECK_LAUNCH();
  unsigned long long w[6] = {0, 0, 0, 0, 0, 0};
  if (slot.d) PB_TRY(mailbox_wait(slot, 6, w, s));
  else PB_TRY(fetch_words(d_words, 6, w, s));
  bt.lap("contig ranges + layout + stats fetch");
  const int64_t m = (int64_t)w[0];
  const unsigned long long total_span = w[5];
  if (w[1] != 0 || m == 0 || total_span == 0 || total_span >= 0xFFFFFFF0ull) return PBGPU_OK;  // inverted rows / empty / axis too long
  *applies = true;
  ix->m = m;
  ix->has_inverted = 0;
  ix->min_end = (int32_t)(long long)w[3];
  ix->max_end = (int32_t)(long long)w[4];
  fr->max_len = (long long)w[2];
  fr->total_span = total_span;
  fr->d_cmap = d_cmap;
  size_t arr_b = 0;
  PB_TRY(dev_alloc(&ix->slab, slab1_bytes(n_contigs, m), s));
  ix->bytes = slab1_bytes(n_contigs, m);
  set_slab1(ix, n_contigs, m, &arr_b);
  PB_TRY(alloc_slab2(ix, total_span, m, n_contigs, s));
  PB_CUDA(cudaMemcpyAsync(ix->cmap, d_cmap, sizeof(ContigMap) * (size_t)n_contigs, cudaMemcpyDeviceToDevice, s));
  PB_LAUNCH(cmap32_kernel, (unsigned)cdiv(n_contigs, 256), 256, 0, s, d_cmap, n_contigs, ix->cmap32);
  bt.lap("slab allocs");
  uint64_t *k1 = nullptr, *k2 = nullptr;
  uint32_t *v1 = nullptr, *v2 = nullptr, *d_totals = nullptr;
  PB_TRY(sc.get(&k1, (size_t)m_in));
  PB_TRY(sc.get(&k2, (size_t)m_in));
  PB_TRY(sc.get(&v1, (size_t)m_in));
  PB_TRY(sc.get(&v2, (size_t)m_in));
  PB_TRY(sc.get(&d_totals, (size_t)kRsMaxPasses * kRsRadix));
  PB_CUDA(cudaMemsetAsync(d_totals, 0, sizeof(uint32_t) * kRsMaxPasses * kRsRadix, s));
  grid = cdiv(m_in, kPrepThreads * 2);
  if (grid > kSMs * 2) grid = kSMs * 2;
  PB_LAUNCH(gkeys_kernel, (unsigned)grid, kPrepThreads, 0, s, d_c, d_s, d_e, m_in, n_contigs, d_cmap, k1, v1, d_totals, d_ids);
  PB_CHECK_LAUNCH();
  // key bytes 4..7 hold the global start; without null-keyed rows (their sentinel sets every bit) only the bytes below
  // the top bit of the axis length vary
  int dpos[4], ndig = 0;
  const int vbits = m < m_in ? 32 : bit_length_u32((uint32_t)(total_span - 1));
  for (int p = 0; p * 8 < vbits; ++p) dpos[ndig++] = 4 + p;
  SortedKV<uint32_t> sorted;
  PB_TRY(radix_sort_digits<uint32_t>(k1, v1, k2, v2, m_in, dpos, ndig, d_totals, s, &sorted));
  bt.lap("global keys + start sort");
  fr->idle = sorted.keys == k1 ? (void *)k2 : (void *)k1;
  unsigned long long *d_inv = nullptr;
  PB_TRY(sc.get(&d_inv, 1));
  PB_CUDA(cudaMemsetAsync(d_inv, 0, sizeof(unsigned long long), s));
  PB_LAUNCH(unpack_gsorted_kernel, (unsigned)cdiv(m, 256), 256, 0, s, sorted.keys, sorted.vals, m, d_cmap, n_contigs, ix->st, ix->en, ix->row,
            ix->er, ix->gs, ix->seg, d_inv);
  PB_CHECK_LAUNCH();
  unsigned long long inv = 0;
  PB_TRY(fetch_words(d_inv, 1, &inv, s));
  bt.lap("unpack + nesting fetch");
  fr->nested = inv != 0;
  fr->gs_ready = true;
  return PBGPU_OK;
}

// ---- front half, generic variant: any contig count, inverted rows, axes beyond 2^32 --------------------------------------
static int build_front_generic(pbgpu_index *ix, const int32_t *d_c, const int32_t *d_s, const int32_t *d_e, const uint32_t *d_ids, int64_t m_in,
                               int32_t n_contigs, cudaStream_t s, Scratch &sc, BuildTrace &bt, BuildFront *fr) {
  // 1. ONE pass over the input: coordinate statistics (key width, fast-path eligibility), the sort keys
  //    contig << 32 | biased start  with values  end << 32 | row, and the digit totals of every radix pass.
  //    The key format does not depend on the statistics, so nothing waits for the host before it.
  const int contig_bits = bit_length_u32((uint32_t)n_contigs);  // codes 0..n_contigs (sentinel included)
  const int contig_digits = (contig_bits + 7) / 8;
  constexpr int pos_bits = 32;
  constexpr uint32_t bias = 0x80000000u;
  BuildStats hs = {INT32_MAX, INT32_MIN, INT32_MAX, INT32_MIN, 0ull, 0ull, 0ull, 0ull};
  uint64_t *keys = nullptr, *vals = nullptr, *keys2 = nullptr, *vals2 = nullptr;
  uint32_t *d_totals = nullptr;
  if (m_in > 0) {
    PB_TRY(sc.get(&keys, (size_t)m_in));
    PB_TRY(sc.get(&vals, (size_t)m_in));
    PB_TRY(sc.get(&keys2, (size_t)m_in));
    PB_TRY(sc.get(&vals2, (size_t)m_in));
    char *d_prep = nullptr;  // BuildStats | digit totals [kRsMaxPasses][256]
    const size_t stats_b = align_up(sizeof(BuildStats)), tot_b = sizeof(uint32_t) * kRsMaxPasses * kRsRadix;
    PB_TRY(sc.get(&d_prep, stats_b + tot_b));
    bt.lap("scratch alloc (keys, vals x2)");
    BuildStats *d_stats = (BuildStats *)d_prep;
    d_totals = (uint32_t *)(d_prep + stats_b);
    PB_CUDA(cudaMemsetAsync(d_prep, 0, stats_b + tot_b, s));
    PB_CUDA(cudaMemcpyAsync(d_stats, &hs, sizeof(hs), cudaMemcpyHostToDevice, s));
    int64_t grid = cdiv(m_in, kPrepThreads * 2);
    if (grid > kSMs * 2) grid = kSMs * 2;
    static_assert(sizeof(BuildStats) == 48, "BuildStats: 5 payload words + the block counter word, posted by prep_kernel");
    const MailboxSlot slot = mailbox_open();
    PB_LAUNCH(prep_kernel, (unsigned)grid, kPrepThreads, 0, s, d_c, d_s, d_e, m_in, n_contigs, 4 + contig_digits, d_stats, keys, vals, d_totals,
              slot.d, slot.seq, d_ids);
    PB_CHECK_LAUNCH();
    if (slot.d) PB_TRY(mailbox_wait(slot, 5, reinterpret_cast<unsigned long long *>(&hs), s));
    else PB_TRY(fetch_words(d_stats, 5, reinterpret_cast<unsigned long long *>(&hs), s));
    bt.lap("prep kernel + stats fetch");
  }
  const int64_t m = (int64_t)hs.valid;
  ix->m = m;
  ix->has_inverted = hs.inverted != 0;
  ix->min_end = hs.min_end;
  ix->max_end = hs.max_end;
  fr->max_len = (long long)hs.max_len;
  size_t arr_b = 0;
  PB_TRY(dev_alloc(&ix->slab, slab1_bytes(n_contigs, m), s));
  bt.lap("slab 1 alloc");
  ix->bytes = slab1_bytes(n_contigs, m);
  set_slab1(ix, n_contigs, m, &arr_b);
  if (m == 0) {
    PB_CUDA(cudaMemsetAsync(ix->seg, 0, sizeof(int32_t) * ((size_t)n_contigs + 2), s));
    return PBGPU_OK;
  }
  // 2. radix partition by contig + sort by start: one stable LSD sort of the (contig | start) keys.  Only digits
  //    that can differ are sorted: every biased start lies between the biased min and max, so they agree above the
  //    highest bit of min ^ max; the contig digits are needed when there is more than one contig or a null key
  //    (sentinel code n_contigs, which must end up behind every real contig).
  int dpos[kRsMaxPasses], ndig = 0;
  {
    const int vb = bit_length_u32((uint32_t)hs.min_start ^ (uint32_t)hs.max_start);
    for (int p = 0; p * 8 < vb; ++p) dpos[ndig++] = p;
  }
  const bool contig_passes = n_contigs > 1 || m < m_in;
  if (contig_passes) for (int q = 0; q < contig_digits; ++q) dpos[ndig++] = 4 + q;
  SortedPairs sorted;
  PB_TRY(radix_sort_digits<uint64_t>(keys, vals, keys2, vals2, m_in, dpos, ndig, d_totals, s, &sorted));
  bt.lap("start sort");
  fr->idle = sorted.keys == keys ? (void *)keys2 : (void *)keys;
  fr->idle2 = sorted.vals == vals ? vals2 : vals;
  keys = sorted.keys;
  vals = sorted.vals;
  fr->keys = keys;
  // 3. unpack + segments + nested-interval detection; contig slices of the global axis.  One host round trip then
  //    decides two things: nested intervals?  global axis fits 32 bits (fast path)?
  const bool try_fast = !ix->has_inverted;
  unsigned long long *d_meta = nullptr;  // [0] end inversions, [1] total span
  unsigned long long *d_span = nullptr;
  PB_TRY(sc.get(&d_meta, 2));
  PB_TRY(sc.get(&d_span, (size_t)n_contigs + 1));
  ContigMap *d_cmap_tmp = nullptr;
  PB_TRY(sc.get(&d_cmap_tmp, (size_t)n_contigs + 1));
  PB_CUDA(cudaMemsetAsync(d_meta, 0, 2 * sizeof(unsigned long long), s));
  PB_LAUNCH(unpack_sorted_kernel, (unsigned)cdiv(m, 256), 256, 0, s, keys, vals, m, pos_bits, bias, n_contigs, ix->st, ix->en, ix->row,
            ix->er, ix->seg, d_meta);
  const bool small_table = n_contigs <= 1024;
  MailboxSlot meta_slot{nullptr, 0};  // set when the layout kernel posts d_meta to the host itself
  if (try_fast) {
    if (small_table) {
      meta_slot = mailbox_open();
      PB_LAUNCH(contig_layout_kernel, 1, 1024, 0, s, ix->seg, ix->st, (long long)hs.max_len, n_contigs, d_cmap_tmp, d_meta, meta_slot.d, meta_slot.seq);
    } else {
      PB_LAUNCH(contig_span_kernel, (unsigned)cdiv(n_contigs, 128), 128, 0, s, ix->seg, ix->st, (long long)hs.max_len, n_contigs, d_cmap_tmp, d_span);
      PB_TRY((device_scan<SumU64, false>(d_span, d_span, n_contigs, d_meta + 1, s)));
      PB_LAUNCH(contig_off_kernel, (unsigned)cdiv(n_contigs, 128), 128, 0, s, d_span, n_contigs, d_cmap_tmp);
    }
  }
  PB_CHECK_LAUNCH();
  unsigned long long h_meta[2] = {0, 0};
  if (meta_slot.d) PB_TRY(mailbox_wait(meta_slot, 2, h_meta, s));
  else PB_TRY(fetch_words(d_meta, 2, h_meta, s));
  bt.lap("unpack + layout + meta fetch");
  fr->nested = h_meta[0] != 0;
  fr->total_span = try_fast ? h_meta[1] : 0;
  fr->d_cmap = d_cmap_tmp;
  return PBGPU_OK;
}

// sweep_only: the caller wants the (contig, start, row) order, the segments and the running max of the ends only (the
// unary sweeps of unary.cuh): the end order and the rank directory are skipped
static int index_build_impl(pbgpu_index *ix, const int32_t *d_c, const int32_t *d_s, const int32_t *d_e, int64_t m_in,
                            int32_t n_contigs, cudaStream_t s, bool sweep_only = false, const uint32_t *d_ids = nullptr) {
  ix->m_in = m_in;
  ix->n_contigs = n_contigs;
  PB_CUDA(cudaGetDevice(&ix->device));
  g_ev.mark(EV_BUILD0, s);
  struct MarkEnd { cudaStream_t s; ~MarkEnd() { g_ev.mark(EV_BUILD1, s); } } mark_end{s};
  Scratch sc(s);
  BuildTrace bt(s);
  BuildFront fr;
  bool gkey = false;
  if (!build_generic_only() && m_in > 0 && n_contigs >= 1 && n_contigs <= 1024)
    PB_TRY(build_front_gkey(ix, d_c, d_s, d_e, d_ids, m_in, n_contigs, s, sc, bt, &fr, &gkey));
  if (!gkey) PB_TRY(build_front_generic(ix, d_c, d_s, d_e, d_ids, m_in, n_contigs, s, sc, bt, &fr));
  const int64_t m = ix->m;
  if (m == 0) return PBGPU_OK;
  const bool nested = fr.nested;
  ix->nested = nested;
  const size_t arr_b = align_up(sizeof(int32_t) * (size_t)m);

  // 4. nested intervals: running maximum of the ends inside every contig (two levels: tile maxima, their scan, tiles).
  //    The ends are NOT sorted here any more (round 1: a second five-pass radix sort, 7 of the 15.8 ms of a 90 M-row
  //    build): the fast path's directory only needs them grouped by bucket (step 5), and the end order proper
  //    (en_sorted / en_pos: nearest and the generic kernels) is built on first use (ensure_end_order) or right away
  //    when the fast path is not available.
  if (nested) {
    PB_TRY(dev_alloc(&ix->slab_n, arr_b, s));
    ix->bytes += arr_b;
    ix->pmax = (int32_t *)ix->slab_n;
    ix->en_sorted = nullptr;  // lazy
    ix->en_pos = nullptr;
    const int64_t tiles = cdiv(m, kPmTile);
    unsigned long long *tile_max = nullptr;
    PB_TRY(sc.get(&tile_max, (size_t)tiles + 1));
    PB_LAUNCH(pmax_tile_max_kernel, (unsigned)tiles, kPmThreads, 0, s, ix->seg, n_contigs, ix->en, m, tile_max);
    PB_CHECK_LAUNCH();
    PB_TRY((device_scan<MaxU64, false>(tile_max, tile_max, tiles, nullptr, s)));
    PB_LAUNCH(pmax_tile_final_kernel, (unsigned)tiles, kPmThreads, 0, s, ix->seg, n_contigs, ix->en, m, tile_max, ix->pmax);
    PB_CHECK_LAUNCH();
    bt.lap("nested: running max");
  }
  if (sweep_only) return PBGPU_OK;

  // 5. fast path: global axis + rank directory
  const unsigned long long total_span = fr.total_span;
  if (!ix->has_inverted && total_span > 0 && total_span < 0xFFFFFFF0ull) {
    if (!ix->slab2) {  // generic front half: the directory slab and the gs column are still to come
      PB_TRY(alloc_slab2(ix, total_span, m, n_contigs, s));
      bt.lap("slab 2 alloc");
      PB_CUDA(cudaMemcpyAsync(ix->cmap, fr.d_cmap, sizeof(ContigMap) * (size_t)n_contigs, cudaMemcpyDeviceToDevice, s));
      PB_LAUNCH(cmap32_kernel, (unsigned)cdiv(n_contigs, 256), 256, 0, s, fr.d_cmap, n_contigs, ix->cmap32);
      PB_LAUNCH(gs_from_keys_kernel, (unsigned)cdiv(m, 256), 256, 0, s, fr.keys, 32, ix->st, m, ix->cmap, ix->gs);
      PB_CHECK_LAUNCH();
    }
    const int shift = ix->shift;
    const uint32_t nb = ix->n_buckets;
    uint32_t *rank_s = nullptr, *rank_e = nullptr;
    PB_TRY(sc.get(&rank_s, (size_t)nb + 1));
    PB_TRY(sc.get(&rank_e, (size_t)nb + 2));  // + the crowded-ends word of the nested variant
    if (!nested) {  // end order == start order: ge and both rank arrays by run filling, one pass over the sorted rows
      PB_LAUNCH(jdir_mark_g_kernel<false>, (unsigned)cdiv(m, 256), 256, 0, s, ix->gs, ix->st, ix->en, m, shift, nb, ix->ge, rank_s, rank_e);
      PB_LAUNCH(jdir_pack2_kernel<false>, (unsigned)cdiv((int64_t)nb + 1, 256), 256, 0, s, ix->gs, ix->ge, m, shift, nb, rank_s, rank_e, ix->jdir,
                (unsigned int *)nullptr);
      PB_CHECK_LAUNCH();
      ix->ge_sorted = 1;
    } else {  // ends grouped by bucket (counting sort with the buckets as bins), unordered inside a bucket
      uint32_t *ge_tmp = (uint32_t *)fr.idle;  // the sort's idle buffer
      unsigned int *d_crowded = rank_e + nb + 1;
      PB_CUDA(cudaMemsetAsync(rank_e, 0, sizeof(uint32_t) * ((size_t)nb + 2), s));
      PB_LAUNCH(jdir_mark_g_kernel<true>, (unsigned)cdiv(m, 256), 256, 0, s, ix->gs, ix->st, ix->en, m, shift, nb, ge_tmp, rank_s, rank_e);
      PB_CHECK_LAUNCH();
      PB_TRY((device_scan<SumU32, false>(rank_e, rank_e, (int64_t)nb + 1, nullptr, s)));
      PB_LAUNCH(jdir_place_ends_kernel, (unsigned)cdiv(m, 256), 256, 0, s, ge_tmp, m, shift, rank_e, ix->ge);
      PB_LAUNCH(jdir_pack2_kernel<true>, (unsigned)cdiv((int64_t)nb + 1, 256), 256, 0, s, ix->gs, ix->ge, m, shift, nb, rank_s, rank_e, ix->jdir,
                d_crowded);
      PB_CHECK_LAUNCH();
      // probes that land in a crowded record count its ends one by one (they are not sorted): fine for a few dozen,
      // not for thousands of equal ends -- then the ends are sorted after all (one pass per varying digit)
      unsigned long long crowded = 0;
      {
        unsigned long long *d_word = nullptr;
        PB_TRY(sc.get(&d_word, 1));
        PB_LAUNCH(widen_word_kernel, 1, 32, 0, s, d_crowded, d_word);
        PB_TRY(fetch_words(d_word, 1, &crowded, s));
      }
      ix->ge_sorted = 0;
      if (crowded > (unsigned long long)kCrowdedLinearMax) {
        uint64_t *gk = nullptr, *gv = nullptr, *gk2 = nullptr, *gv2 = nullptr;
        PB_TRY(sc.get(&gk, (size_t)m));
        PB_TRY(sc.get(&gv, (size_t)m));
        PB_TRY(sc.get(&gk2, (size_t)m));
        PB_TRY(sc.get(&gv2, (size_t)m));
        PB_LAUNCH(widen_u32_keys_kernel, (unsigned)cdiv(m, 256), 256, 0, s, ix->ge, m, gk, gv);
        int gpos[4] = {0, 1, 2, 3};
        SortedPairs gsrt;
        PB_TRY(radix_sort_digits<uint64_t>(gk, gv, gk2, gv2, m, gpos, 4, nullptr, s, &gsrt));
        PB_LAUNCH(narrow_u32_keys_kernel, (unsigned)cdiv(m, 256), 256, 0, s, gsrt.keys, m, ix->ge);
        PB_CHECK_LAUNCH();
        ix->ge_sorted = 1;
      }
    }
    bt.lap("directory");
    ix->fast = 1;
  }
  if (nested && !ix->fast) PB_TRY(build_end_order(ix, s));  // the generic kernels search the end order
  return PBGPU_OK;
}

int pbgpu_index_build(const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t m, int32_t n_contigs,
                      void *stream, pbgpu_index **out) {
  return pbgpu_index_build_ids(d_contig, d_start, d_end, nullptr, m, n_contigs, stream, out);
}

int pbgpu_index_build_ids(const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, const uint32_t *d_row_ids, int64_t m,
                          int32_t n_contigs, void *stream, pbgpu_index **out) {
  if (!out) return set_error(PBGPU_EINVAL, "out is NULL");
  *out = nullptr;
  if (m < 0 || n_contigs < 0) return set_error(PBGPU_EINVAL, "negative size");
  if (m > 0 && (!d_contig || !d_start || !d_end)) return set_error(PBGPU_EINVAL, "NULL column");
  if (m >= (int64_t)INT32_MAX) return set_error(PBGPU_ERANGE, "indexed table has %lld rows; limit is 2^31-2", (long long)m);
  if (n_contigs >= (1 << 30)) return set_error(PBGPU_ERANGE, "too many contigs");
  pbgpu_index *ix = new (std::nothrow) pbgpu_index();
  if (!ix) return set_error(PBGPU_ENOMEM, "host allocation failed");
  int rc = index_build_impl(ix, d_contig, d_start, d_end, m, n_contigs, (cudaStream_t)stream, false, d_row_ids);
  if (rc != PBGPU_OK) {
    pbgpu_index_free(ix);
    return rc;
  }
  *out = ix;
  return PBGPU_OK;
}

// ---------------------------------------------------------------------------------------------
static int check_probe_args(const pbgpu_index *ix, const int32_t *c, const int32_t *s, const int32_t *e, int64_t n, int filter_op) {
  if (!ix) return set_error(PBGPU_EINVAL, "index is NULL");
  if (n < 0) return set_error(PBGPU_EINVAL, "negative row count");
  if (n > 0 && (!c || !s || !e)) return set_error(PBGPU_EINVAL, "NULL column");
  if (n >= 0xFFFFFFFFll) return set_error(PBGPU_ERANGE, "iterated table has %lld rows; limit is 2^32-2", (long long)n);
  if (filter_op != PBGPU_FILTER_WEAK && filter_op != PBGPU_FILTER_STRICT) return set_error(PBGPU_EINVAL, "bad filter_op %d", filter_op);
  return PBGPU_OK;
}

}  // extern "C"

namespace pbgpu {
// ---- probe partition (bins.cuh) ---------------------------------------------------------------------------------------
// When: fast-path index whose rank directory is well beyond the L2 and enough probes to pay for the extra pass.
// PBGPU_BIN=0 never, PBGPU_BIN=1 whenever the fast path is available (tests force it on small inputs).
static int bin_mode() {
  static int v = [] { const char *e = getenv("PBGPU_BIN"); return e ? (e[0] == '0' ? 0 : 2) : 1; }();
  return v;
}
static bool want_bins(const pbgpu_index *ix, int64_t n) {
  if (!ix->fast || n <= 0 || n >= (int64_t)kLbMask || bin_mode() == 0) return false;
  if (bin_mode() == 2) return true;
  return n >= (1 << 22) && (size_t)ix->n_buckets * sizeof(JRec) > ((size_t)96 << 20);
}
struct BinnedProbes {
  int4 *recs = nullptr;
  uint32_t *pos = nullptr;
  void *slab = nullptr;
  int bin_shift = 0;
};
// hist + one stable partition pass, enqueued on s.  The slab (records | pos | partition scratch) is the caller's to free.
static int bin_probes(const pbgpu_index *ix, const int32_t *pc, const int32_t *ps, const int32_t *pe, const uint32_t *ids, int64_t n,
                      int filter_op, bool write_pos, cudaStream_t s, BinnedProbes *out) {
  const int64_t tiles = cdiv(n, kBinTile);
  const size_t rec_b = align_up(sizeof(int4) * (size_t)n), pos_b = write_pos ? align_up(sizeof(uint32_t) * (size_t)n) : 0;
  const size_t work_w = (size_t)kBinRadix + 64 + (size_t)tiles * kBinRadix;  // totals | ticket | status
  PB_TRY(dev_alloc(&out->slab, rec_b + pos_b + sizeof(uint32_t) * work_w, s));
  out->recs = (int4 *)out->slab;
  out->pos = write_pos ? (uint32_t *)((char *)out->slab + rec_b) : nullptr;
  uint32_t *work = (uint32_t *)((char *)out->slab + rec_b + pos_b);
  uint32_t *totals = work, *ticket = work + kBinRadix, *status = work + kBinRadix + 64;
  int span_bits = 0;
  for (uint32_t v = ix->axis_span; v; v >>= 1) ++span_bits;
  out->bin_shift = span_bits > 8 ? span_bits - 8 : 0;
  g_ev.mark(EV_BIN0, s);
  PB_CUDA(cudaMemsetAsync(work, 0, sizeof(uint32_t) * work_w, s));
  int64_t hgrid = cdiv(n, 512 * 4);
  if (hgrid > kSMs * 4) hgrid = kSMs * 4;
  PB_LAUNCH(bin_hist_kernel, (unsigned)hgrid, 512, 0, s, view_of(ix), pc, ps, pe, n, out->bin_shift, filter_op == PBGPU_FILTER_STRICT, totals);
  constexpr size_t stage_b = sizeof(int4) * kBinTile;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(bin_partition_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_b);
    cudaFuncSetAttribute(bin_partition_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_b);
    cudaFuncSetAttribute(bin_partition_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_b);
    cudaFuncSetAttribute(bin_partition_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_b);
  });
  const int strict = filter_op == PBGPU_FILTER_STRICT;
  // PBGPU_BIN_STORE=ldst: copy the staged runs out through registers instead of TMA bulk stores (A/B)
  static const int bulk = [] { const char *e = getenv("PBGPU_BIN_STORE"); return (e && !strcmp(e, "ldst")) ? 0 : 1; }();
#define PB_BINPART(WP, OCC)                                                                                                         \
  PB_LAUNCH((bin_partition_kernel<WP, OCC>), (unsigned)tiles, kBinThreads, stage_b, s, view_of(ix), pc, ps, pe, ids, n, out->bin_shift, strict, \
            totals, status, ticket, out->recs, out->pos, bulk)
  // r2f: the 40-register cap of the 3-blocks-per-SM build spills the 16-byte records (1.74 vs 1.55 ms per 100 M probes)
  static const int bin_occ = [] { const char *e = getenv("PBGPU_BIN_OCC"); return (e && e[0] == '3') ? 3 : 2; }();
  if (bin_occ == 3) { if (write_pos) PB_BINPART(true, 3); else PB_BINPART(false, 3); }
  else { if (write_pos) PB_BINPART(true, 2); else PB_BINPART(false, 2); }
#undef PB_BINPART
  PB_CHECK_LAUNCH();
  g_ev.mark(EV_BIN1, s);
  return PBGPU_OK;
}

// probes per thread in the fast count kernels (PBGPU_ITEMS=1|2; default 2)
static int sweep_items() {
  static int v = [] { const char *e = getenv("PBGPU_ITEMS"); return (e && e[0] == '1') ? 1 : ((e && e[0] == '4') ? 4 : 2); }();
  return v;
}
template <typename OutT>
int count_overlaps_impl(const pbgpu_index *ix, const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t n,
                        int filter_op, OutT *d_counts, cudaStream_t s) {
  PB_TRY(check_probe_args(ix, d_contig, d_start, d_end, n, filter_op));
  if (n == 0) return PBGPU_OK;
  if (!d_counts) return set_error(PBGPU_EINVAL, "d_counts is NULL");
  g_ev.mark(EV_COUNT0, s);
  const unsigned grid = (unsigned)cdiv(n, kSweepThreads);
  const bool strict = filter_op == PBGPU_FILTER_STRICT;
  if (want_bins(ix, n)) {  // index beyond the L2: partition the probes by coordinate, count in bin order, counts back to row order
    BinnedProbes bp;
    int rc = bin_probes(ix, d_contig, d_start, d_end, nullptr, n, filter_op, true, s, &bp);
    uint32_t *cnt_b = nullptr;
    if (rc == PBGPU_OK) rc = dev_alloc_t(&cnt_b, (size_t)n, s);
    if (rc == PBGPU_OK) {
      const unsigned g2 = (unsigned)cdiv(n, kSweepThreads * 2);
      if (strict) PB_LAUNCH((binned_count_kernel<true, 2>), g2, kSweepThreads, 0, s, view_of(ix), bp.recs, n, cnt_b);
      else PB_LAUNCH((binned_count_kernel<false, 2>), g2, kSweepThreads, 0, s, view_of(ix), bp.recs, n, cnt_b);
      g_ev.mark(EV_UNBIN0, s);
      PB_LAUNCH(unbin_counts_kernel<OutT>, (unsigned)cdiv(n, 256 * 4), 256, 0, s, cnt_b, bp.pos, n, d_counts);
      g_ev.mark(EV_UNBIN1, s);
      if (cudaGetLastError() != cudaSuccess) rc = set_error(PBGPU_ECUDA, "binned count launch failed");
    }
    dev_free(cnt_b, s);
    dev_free(bp.slab, s);
    if (rc != PBGPU_OK) return rc;
  } else if (ix->fast) {
    const int items = sweep_items();
    const unsigned g2 = (unsigned)cdiv(n, kSweepThreads * 2), g4 = (unsigned)cdiv(n, kSweepThreads * 4);
    if (items == 4) {
      if (strict) PB_LAUNCH((count_overlaps_fast_kernel<true, OutT, 4>), g4, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, d_counts);
      else PB_LAUNCH((count_overlaps_fast_kernel<false, OutT, 4>), g4, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, d_counts);
    } else if (items == 2) {
      if (strict) PB_LAUNCH((count_overlaps_fast_kernel<true, OutT, 2>), g2, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, d_counts);
      else PB_LAUNCH((count_overlaps_fast_kernel<false, OutT, 2>), g2, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, d_counts);
    } else {
      if (strict) PB_LAUNCH((count_overlaps_fast_kernel<true, OutT, 1>), grid, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, d_counts);
      else PB_LAUNCH((count_overlaps_fast_kernel<false, OutT, 1>), grid, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, d_counts);
    }
  } else {
    if (strict) PB_LAUNCH((count_overlaps_kernel<true, OutT>), grid, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, d_counts);
    else PB_LAUNCH((count_overlaps_kernel<false, OutT>), grid, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, d_counts);
  }
  PB_CHECK_LAUNCH();
  g_ev.mark(EV_COUNT1, s);
  return PBGPU_OK;
}
// contig codes that travelled as bytes (Arrow bridge, <= 255 indexed contigs): widen, unknown codes -> null key
__global__ void __launch_bounds__(256) widen_codes_u8_kernel(const uint8_t *__restrict__ in, int64_t n, int32_t n_contigs,
                                                             int32_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) { const int32_t v = in[i]; out[i] = v < n_contigs ? v : -1; }
}
int widen_codes_u8(const uint8_t *d_in, int64_t n, int32_t n_contigs, int32_t *d_out, void *stream) {
  if (n <= 0) return PBGPU_OK;
  PB_LAUNCH(widen_codes_u8_kernel, (unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream, d_in, n, n_contigs, d_out);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}
// result-row contig codes for the Arrow bridge, as bytes (matched rows only: 0 <= code < n_contigs <= 255)
__global__ void __launch_bounds__(256) gather_i32_u8_kernel(const int32_t *__restrict__ src, const uint32_t *__restrict__ rows,
                                                            int64_t n, uint8_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) out[i] = (uint8_t)__ldg(src + rows[i]);
}
int gather_i32_u8(const int32_t *d_src, const uint32_t *d_rows, int64_t n, uint8_t *d_out, void *stream) {
  if (n <= 0) return PBGPU_OK;
  PB_LAUNCH(gather_i32_u8_kernel, (unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream, d_src, d_rows, n, d_out);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}
// internal (same .so, not part of the C ABI): 32-bit counts for the Arrow bridge, widened on the host
int count_overlaps_u32(const pbgpu_index *ix, const int32_t *c, const int32_t *s_, const int32_t *e, int64_t n, int filter_op,
                       uint32_t *d_counts, void *stream) {
  return count_overlaps_impl<uint32_t>(ix, c, s_, e, n, filter_op, d_counts, (cudaStream_t)stream);
}
}  // namespace pbgpu

extern "C" {

int pbgpu_count_overlaps(const pbgpu_index *ix, const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t n,
                         int filter_op, int64_t *d_counts, void *stream) {
  return count_overlaps_impl<int64_t>(ix, d_contig, d_start, d_end, n, filter_op, d_counts, (cudaStream_t)stream);
}

int pbgpu_coverage(const pbgpu_index *ix, const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t n,
                   int filter_op, int64_t *d_coverage, void *stream) {
  PB_TRY(check_probe_args(ix, d_contig, d_start, d_end, n, filter_op));
  if (n == 0) return PBGPU_OK;
  if (!d_coverage) return set_error(PBGPU_EINVAL, "d_coverage is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)cdiv(n, kSweepThreads);
  if (filter_op == PBGPU_FILTER_STRICT)
    PB_LAUNCH(coverage_kernel<true>, grid, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, d_coverage);
  else
    PB_LAUNCH(coverage_kernel<false>, grid, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, d_coverage);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}

}  // extern "C"

// PBGPU_EMIT=walk: pass 2 always walks the candidate window (the first implementation; kept for A/B runs).  The flat
// expansion needs the fast path, no nested intervals and 32 x indexed rows < 2^32 (32-bit warp scan of the counts).
static bool emit_by_walk() {
  static bool walk = [] { const char *e = getenv("PBGPU_EMIT"); return e && !strcmp(e, "walk"); }();
  return walk;
}
static bool emit_flat_ok(const pbgpu_index *ix) { return !emit_by_walk() && ix->fast && !ix->nested && ix->m < (1ll << 27); }

// Pass 1 leaves raw block totals and a device scan turns them into offsets.  PBGPU_P1SCAN=lookback does the scan
// inside pass 1 instead (decoupled look-back over its blocks): measured 2.5x SLOWER (r01s: 181 vs 72 us at 10M
// probes -- a 512-probe tile lives ~5 us, the ticket + status round trips add ~3 us to every one), kept for A/B runs.
static bool p1_scan_by_kernels() {
  static bool v = [] { const char *e = getenv("PBGPU_P1SCAN"); return !(e && !strcmp(e, "lookback")); }();
  return v;
}

struct pbgpu_overlap_plan {
  const pbgpu_index *ix;
  const int32_t *pc, *ps, *pe;
  int64_t n;
  int filter_op;
  int device;
  int64_t nblk;
  uint32_t *counts;                 // [n]
  uint32_t *his;                    // [n] start-rank of every probe (fast path only)
  unsigned long long *block_base;   // [nblk] exclusive-scanned block totals
  unsigned long long *warp_off;     // [n/32] offset of every 32-probe group inside its block (flat pass 2 only)
  void *slab;                       // one stream-ordered allocation behind the arrays
  int64_t total;
  const int4 *recs;                 // probes partitioned by coordinate (bins.cuh) when the index is beyond the L2, else NULL
  void *bin_slab;
  const uint32_t *probe_ids;        // id reported for probe row i (NULL: i itself)
};

extern "C" {

void pbgpu_overlap_plan_free(pbgpu_overlap_plan *p) {
  if (!p) return;
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != p->device) cudaSetDevice(p->device);
  dev_free(p->slab, 0);  // legacy stream: ordered after the emit kernel of blocking streams
  dev_free(p->bin_slab, 0);
  if (cur != p->device) cudaSetDevice(cur);
  delete p;
}

void pbgpu_overlap_plan_free_async(pbgpu_overlap_plan *p, void *stream) {
  if (!p) return;
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != p->device) cudaSetDevice(p->device);
  dev_free(p->slab, (cudaStream_t)stream);  // ordered after the pass-2 launches enqueued on `stream`
  dev_free(p->bin_slab, (cudaStream_t)stream);
  if (cur != p->device) cudaSetDevice(cur);
  delete p;
}

const uint32_t *pbgpu_overlap_plan_counts(const pbgpu_overlap_plan *plan) { return plan ? plan->counts : nullptr; }

int pbgpu_overlap_count(const pbgpu_index *ix, const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t n,
                        int filter_op, void *stream, pbgpu_overlap_plan **plan, int64_t *total_pairs) {
  return pbgpu_overlap_count_ids(ix, d_contig, d_start, d_end, nullptr, n, filter_op, stream, plan, total_pairs);
}

int pbgpu_overlap_count_ids(const pbgpu_index *ix, const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end,
                            const uint32_t *d_probe_ids, int64_t n, int filter_op, void *stream, pbgpu_overlap_plan **plan,
                            int64_t *total_pairs) {
  PB_TRY(check_probe_args(ix, d_contig, d_start, d_end, n, filter_op));
  if (!plan || !total_pairs) return set_error(PBGPU_EINVAL, "plan/total_pairs is NULL");
  *plan = nullptr;
  *total_pairs = 0;
  cudaStream_t s = (cudaStream_t)stream;
  pbgpu_overlap_plan *p = new (std::nothrow) pbgpu_overlap_plan();
  if (!p) return set_error(PBGPU_ENOMEM, "host allocation failed");
  *p = pbgpu_overlap_plan{ix, d_contig, d_start, d_end, n, filter_op, 0, cdiv(n, kSweepThreads), nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, d_probe_ids};
  cudaGetDevice(&p->device);
  auto fail = [&](int rc) { pbgpu_overlap_plan_free(p); return rc; };
  if (n == 0) { *plan = p; return PBGPU_OK; }
  const int items = sweep_items();
  const unsigned grid = (unsigned)p->nblk, grid_fast = (unsigned)cdiv(n, (int64_t)kSweepThreads * items);
  const bool binned = want_bins(ix, n);
  const bool lookback = ix->fast && !binned && !p1_scan_by_kernels();
  unsigned long long *d_status = nullptr;
  {
    const size_t cb = align_up(sizeof(uint32_t) * (size_t)n), bb = align_up(sizeof(unsigned long long) * (size_t)(p->nblk + 1));
    const bool flat = emit_flat_ok(ix) && !binned;
    const size_t wb = flat ? align_up(sizeof(unsigned long long) * (size_t)(p->nblk * (kSweepThreads / 32))) : 0;
    const size_t sb = lookback ? align_up(sizeof(unsigned long long) * ((size_t)grid_fast + 1)) : 0;  // status words + ticket
    int rc0 = dev_alloc(&p->slab, 2 * cb + bb + wb + sb, s);
    if (rc0 != PBGPU_OK) return fail(rc0);
    p->counts = (uint32_t *)p->slab;
    p->his = (uint32_t *)((char *)p->slab + cb);
    p->block_base = (unsigned long long *)((char *)p->slab + 2 * cb);
    if (flat) p->warp_off = (unsigned long long *)((char *)p->slab + 2 * cb + bb);
    if (lookback) {
      d_status = (unsigned long long *)((char *)p->slab + 2 * cb + bb + wb);
      if (cudaMemsetAsync(d_status, 0, sb, s) != cudaSuccess) return fail(set_error(PBGPU_ECUDA, "memset failed"));
    }
  }
  g_ev.mark(EV_P1_0, s);
  unsigned long long *d_total = p->block_base + p->nblk;
  unsigned long long h_total = 0;
  int rc = PBGPU_OK;
  if (binned) {  // index beyond the L2: pass 1 and pass 2 run over the probes partitioned by coordinate
    BinnedProbes bp;
    rc = bin_probes(ix, d_contig, d_start, d_end, d_probe_ids, n, filter_op, false, s, &bp);
    p->recs = bp.recs;
    p->bin_slab = bp.slab;
    if (rc != PBGPU_OK) return fail(rc);
    const unsigned g2 = (unsigned)cdiv(n, kSweepThreads * 2);
    if (filter_op == PBGPU_FILTER_STRICT)
      PB_LAUNCH((binned_p1_kernel<true, 2>), g2, kSweepThreads, 0, s, view_of(ix), p->recs, n, p->counts, p->his, p->block_base, p->warp_off);
    else
      PB_LAUNCH((binned_p1_kernel<false, 2>), g2, kSweepThreads, 0, s, view_of(ix), p->recs, n, p->counts, p->his, p->block_base, p->warp_off);
    if (cudaGetLastError() != cudaSuccess) return fail(set_error(PBGPU_ECUDA, "binned_p1_kernel launch failed"));
    g_ev.mark(EV_P1_1, s);
  } else if (ix->fast) {
    const bool strict = filter_op == PBGPU_FILTER_STRICT;
    const MailboxSlot slot = lookback ? mailbox_open() : MailboxSlot{nullptr, 0};
    unsigned int *d_ticket = lookback ? (unsigned int *)(d_status + grid_fast) : nullptr;
#define PB_P1(ST, IT, LB)                                                                                                   \
  PB_LAUNCH((overlap_count_fast_kernel<ST, IT, LB>), grid_fast, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, \
            p->counts, p->his, p->block_base, p->warp_off, d_status, d_ticket, slot.d, slot.seq)
#define PB_P1_ITEMS(ST, LB) do { if (items == 4) PB_P1(ST, 4, LB); else if (items == 2) PB_P1(ST, 2, LB); else PB_P1(ST, 1, LB); } while (0)
    if (lookback) { if (strict) PB_P1_ITEMS(true, true); else PB_P1_ITEMS(false, true); }
    else { if (strict) PB_P1_ITEMS(true, false); else PB_P1_ITEMS(false, false); }
#undef PB_P1_ITEMS
#undef PB_P1
    if (cudaGetLastError() != cudaSuccess) return fail(set_error(PBGPU_ECUDA, "overlap_count_fast_kernel launch failed"));
    g_ev.mark(EV_P1_1, s);
    if (lookback) {  // offsets and total came out of pass 1 itself; the total is already on its way to the mailbox
      g_ev.mark(EV_SCAN1, s);
      if (slot.d) rc = mailbox_wait(slot, 1, &h_total, s);
      else if (cudaMemcpyAsync(&h_total, d_total, sizeof(h_total), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
        rc = set_error(PBGPU_ECUDA, "overlap pass 1 failed: %s", cudaGetErrorString(cudaGetLastError()));
      if (rc != PBGPU_OK) return fail(rc);
    }
  } else {
    if (filter_op == PBGPU_FILTER_STRICT)
      PB_LAUNCH(overlap_count_kernel<true>, grid, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, p->counts, p->block_base);
    else
      PB_LAUNCH(overlap_count_kernel<false>, grid, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, p->counts, p->block_base);
    if (cudaGetLastError() != cudaSuccess) return fail(set_error(PBGPU_ECUDA, "overlap_count_kernel launch failed"));
    g_ev.mark(EV_P1_1, s);
  }
  if (!lookback) {
    const MailboxSlot slot = mailbox_open();
    bool posted = false;
    rc = device_scan<SumU64, false>(p->block_base, p->block_base, p->nblk, d_total, s, slot.d, slot.seq, &posted);
    if (rc != PBGPU_OK) return fail(rc);
    g_ev.mark(EV_SCAN1, s);
    rc = posted ? mailbox_wait(slot, 1, &h_total, s) : fetch_words(d_total, 1, &h_total, s);
    if (rc != PBGPU_OK) return fail(rc);
  }
  p->total = (int64_t)h_total;
  *total_pairs = p->total;
  *plan = p;
  return PBGPU_OK;
}

namespace pbgpu {
__global__ void __launch_bounds__(256) probe_ids_in_place_kernel(uint32_t *__restrict__ rows, const unsigned long long *__restrict__ block_base,
                                                                 int64_t blk_lo, int64_t blk_hi, const uint32_t *__restrict__ ids) {
  const unsigned long long cnt = block_base[blk_hi] - block_base[blk_lo];
  for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < cnt; i += (unsigned long long)gridDim.x * 256) rows[i] = ids[rows[i]];
}
}  // namespace pbgpu

static int emit_blocks_impl(const pbgpu_overlap_plan *p, int64_t blk_lo, int64_t blk_hi,
                            uint32_t *d_probe_rows, uint32_t *d_build_rows, cudaStream_t s) {
  g_ev.mark(EV_EMIT0, s);
  const unsigned grid = (unsigned)(blk_hi - blk_lo);
  if (p->warp_off) {  // fast path over an index without nested intervals: pure expansion of (count, start rank)
    const unsigned g2 = (unsigned)cdiv(blk_hi - blk_lo, 2);
    if (p->filter_op == PBGPU_FILTER_STRICT)
      PB_LAUNCH((overlap_emit_flat_kernel<true, 2>), g2, kSweepThreads, 0, s, view_of(p->ix), p->pc, p->ps, p->pe, p->n, p->counts,
                p->his, p->block_base, p->warp_off, blk_lo, blk_hi, d_probe_rows, d_build_rows, p->probe_ids);
    else
      PB_LAUNCH((overlap_emit_flat_kernel<false, 2>), g2, kSweepThreads, 0, s, view_of(p->ix), p->pc, p->ps, p->pe, p->n, p->counts,
                p->his, p->block_base, p->warp_off, blk_lo, blk_hi, d_probe_rows, d_build_rows, p->probe_ids);
  } else if (p->recs) {  // partitioned probes: pairs leave in bin order
    if (p->filter_op == PBGPU_FILTER_STRICT)
      PB_LAUNCH((overlap_emit_staged_kernel<true, true>), grid, kSweepThreads, 0, s, view_of(p->ix), p->recs, p->pc, p->ps, p->pe, (const uint32_t *)nullptr, p->n, p->counts,
                p->his, p->block_base, blk_lo, d_probe_rows, d_build_rows);
    else
      PB_LAUNCH((overlap_emit_staged_kernel<false, true>), grid, kSweepThreads, 0, s, view_of(p->ix), p->recs, p->pc, p->ps, p->pe, (const uint32_t *)nullptr, p->n, p->counts,
                p->his, p->block_base, blk_lo, d_probe_rows, d_build_rows);
  } else if (p->ix->fast && !emit_by_walk()) {
    if (p->filter_op == PBGPU_FILTER_STRICT)
      PB_LAUNCH((overlap_emit_staged_kernel<true, false>), grid, kSweepThreads, 0, s, view_of(p->ix), (const int4 *)nullptr, p->pc, p->ps, p->pe, p->probe_ids, p->n,
                p->counts, p->his, p->block_base, blk_lo, d_probe_rows, d_build_rows);
    else
      PB_LAUNCH((overlap_emit_staged_kernel<false, false>), grid, kSweepThreads, 0, s, view_of(p->ix), (const int4 *)nullptr, p->pc, p->ps, p->pe, p->probe_ids, p->n,
                p->counts, p->his, p->block_base, blk_lo, d_probe_rows, d_build_rows);
  } else if (p->ix->fast) {
    if (p->filter_op == PBGPU_FILTER_STRICT)
      PB_LAUNCH(overlap_emit_fast_kernel<true>, grid, kSweepThreads, 0, s, view_of(p->ix), p->pc, p->ps, p->pe, p->n, p->counts,
                p->his, p->block_base, blk_lo, d_probe_rows, d_build_rows);
    else
      PB_LAUNCH(overlap_emit_fast_kernel<false>, grid, kSweepThreads, 0, s, view_of(p->ix), p->pc, p->ps, p->pe, p->n, p->counts,
                p->his, p->block_base, blk_lo, d_probe_rows, d_build_rows);
  } else if (p->filter_op == PBGPU_FILTER_STRICT)
    PB_LAUNCH(overlap_emit_kernel<true>, grid, kSweepThreads, 0, s, view_of(p->ix), p->pc, p->ps, p->pe, p->n, p->counts,
              p->block_base, blk_lo, d_probe_rows, d_build_rows);
  else
    PB_LAUNCH(overlap_emit_kernel<false>, grid, kSweepThreads, 0, s, view_of(p->ix), p->pc, p->ps, p->pe, p->n, p->counts,
              p->block_base, blk_lo, d_probe_rows, d_build_rows);
  // the window-walking kernels of the generic path report probe rows: map them to the caller's ids in place
  if (p->probe_ids && !p->recs && !p->warp_off && !(p->ix->fast && !emit_by_walk()))
    PB_LAUNCH(probe_ids_in_place_kernel, kSMs * 8, 256, 0, s, d_probe_rows, p->block_base, blk_lo, blk_hi, p->probe_ids);
  PB_CHECK_LAUNCH();
  g_ev.mark(EV_EMIT1, s);
  return PBGPU_OK;
}

int pbgpu_overlap_emit(const pbgpu_overlap_plan *p, uint32_t *d_probe_rows, uint32_t *d_build_rows, void *stream) {
  if (!p) return set_error(PBGPU_EINVAL, "plan is NULL");
  if (p->n == 0 || p->total == 0) return PBGPU_OK;
  if (!d_probe_rows || !d_build_rows) return set_error(PBGPU_EINVAL, "output buffer is NULL");
  return emit_blocks_impl(p, 0, p->nblk, d_probe_rows, d_build_rows, (cudaStream_t)stream);
}

int64_t pbgpu_overlap_plan_blocks(const pbgpu_overlap_plan *p) { return p ? p->nblk : 0; }

int pbgpu_overlap_plan_block_offsets(const pbgpu_overlap_plan *p, uint64_t *h_offsets, void *stream) {
  if (!p || !h_offsets) return set_error(PBGPU_EINVAL, "plan/h_offsets is NULL");
  if (p->n == 0) { h_offsets[0] = 0; return PBGPU_OK; }
  cudaStream_t s = (cudaStream_t)stream;
  PB_CUDA(cudaMemcpyAsync(h_offsets, p->block_base, sizeof(uint64_t) * (size_t)(p->nblk + 1), cudaMemcpyDeviceToHost, s));
  PB_CUDA(cudaStreamSynchronize(s));
  return PBGPU_OK;
}

int pbgpu_overlap_emit_blocks(const pbgpu_overlap_plan *p, int64_t blk_lo, int64_t blk_hi, uint32_t *d_probe_rows,
                              uint32_t *d_build_rows, void *stream) {
  if (!p) return set_error(PBGPU_EINVAL, "plan is NULL");
  if (blk_lo < 0 || blk_hi > p->nblk || blk_lo > blk_hi) return set_error(PBGPU_EINVAL, "block range [%lld,%lld) outside [0,%lld)", (long long)blk_lo, (long long)blk_hi, (long long)p->nblk);
  if (blk_lo == blk_hi || p->total == 0) return PBGPU_OK;
  if (!d_probe_rows || !d_build_rows) return set_error(PBGPU_EINVAL, "output buffer is NULL");
  return emit_blocks_impl(p, blk_lo, blk_hi, d_probe_rows, d_build_rows, (cudaStream_t)stream);
}

int pbgpu_nearest(const pbgpu_index *ix, const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t n,
                  int filter_op, int64_t k, int include_overlaps, uint32_t *d_partner, int64_t *d_distance, void *stream) {
  PB_TRY(check_probe_args(ix, d_contig, d_start, d_end, n, filter_op));
  if (k < 1) return set_error(PBGPU_EINVAL, "k must be >= 1");
  if (n == 0) return PBGPU_OK;
  if (!d_partner) return set_error(PBGPU_EINVAL, "d_partner is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  PB_TRY(ensure_end_order(ix, s));
  const unsigned grid = (unsigned)cdiv(n, kSweepThreads);
  if (filter_op == PBGPU_FILTER_STRICT)
    PB_LAUNCH(nearest_kernel<true>, grid, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, k, include_overlaps,
              d_partner, d_distance);
  else
    PB_LAUNCH(nearest_kernel<false>, grid, kSweepThreads, 0, s, view_of(ix), d_contig, d_start, d_end, n, k, include_overlaps,
              d_partner, d_distance);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}

}  // extern "C"

// ---- multi-GPU plumbing: bucket rows by owning rank ---------------------------------------------
namespace pbgpu {

__global__ void __launch_bounds__(256) owner_keys_kernel(const int32_t *__restrict__ c, int64_t n, const int32_t *__restrict__ owner,
                                                         int32_t n_contigs, int32_t n_ranks, uint64_t *__restrict__ keys,
                                                         uint64_t *__restrict__ vals, unsigned long long *__restrict__ rank_counts) {
  __shared__ unsigned int bins[256];  // per-block rank histogram: one global atomic per rank per block
  for (int i = threadIdx.x; i < 256; i += blockDim.x) bins[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int32_t cc = c[i];
    int32_t r = (cc >= 0 && cc < n_contigs) ? owner[cc] : n_ranks;
    if (r < 0 || r > n_ranks) r = n_ranks;
    keys[i] = (uint64_t)r;
    vals[i] = (uint64_t)i;
    if (r < n_ranks) atomicAdd(&bins[r], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_ranks; i += blockDim.x) if (bins[i]) atomicAdd(rank_counts + i, (unsigned long long)bins[i]);
}

__global__ void __launch_bounds__(256) pack_records_kernel(const uint64_t *__restrict__ perm, int64_t kept, const int32_t *__restrict__ c,
                                                           const int32_t *__restrict__ s, const int32_t *__restrict__ e,
                                                           uint32_t row_id_base, int4 *__restrict__ packed) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kept) return;
  const uint64_t r = perm[i];
  packed[i] = make_int4(c[r], s[r], e[r], (int)(row_id_base + (uint32_t)r));
}

}  // namespace pbgpu

namespace pbgpu {
__global__ void __launch_bounds__(256) unpack_records_kernel(const int4 *__restrict__ rec, int64_t n, int32_t *__restrict__ c,
                                                             int32_t *__restrict__ s, int32_t *__restrict__ e, uint32_t *__restrict__ row) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 r = rec[i];
  c[i] = r.x; s[i] = r.y; e[i] = r.z; row[i] = (uint32_t)r.w;
}
__global__ void __launch_bounds__(256) translate_rows_kernel(const uint32_t *__restrict__ idx, int64_t n, const uint32_t *__restrict__ table,
                                                             uint32_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const uint32_t k = idx[i]; out[i] = k == PBGPU_NO_PARTNER ? PBGPU_NO_PARTNER : table[k]; }
}
}  // namespace pbgpu

namespace pbgpu {
// rows per contig (null keys ignored), added onto hist: block-private shared-memory bins, one flush per block
__global__ void __launch_bounds__(256) contig_hist_kernel(const int32_t *__restrict__ c, int64_t n, int32_t n_contigs,
                                                          unsigned long long *__restrict__ hist) {
  extern __shared__ unsigned int bins[];
  const bool use_smem = n_contigs <= 4096;
  if (use_smem) { for (int i = threadIdx.x; i < n_contigs; i += blockDim.x) bins[i] = 0; __syncthreads(); }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t cc = c[i];
    if (cc < 0 || cc >= n_contigs) continue;
    if (use_smem) atomicAdd(&bins[cc], 1u); else atomicAdd(hist + cc, 1ull);
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < n_contigs; i += blockDim.x) if (bins[i]) atomicAdd(hist + i, (unsigned long long)bins[i]);
  }
}
}  // namespace pbgpu

namespace pbgpu {
__global__ void __launch_bounds__(256) gather_i32_kernel(const int32_t *__restrict__ src, const uint32_t *__restrict__ rows, int64_t n,
                                                         int32_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const uint32_t r = rows[i]; out[i] = r == PBGPU_NO_PARTNER ? 0 : __ldg(src + r); }
}
}  // namespace pbgpu

// ---- payload gather on the device (the reference's output IS the joined rows, operation.rs:272-303) -----
// Result row j takes column value src[rows[j]] (rows[j] == PBGPU_NO_PARTNER: null).  Fixed-width values of 1..16 bytes,
// validity as one byte per source row in, one BIT per result row out (Arrow layout, packed by warp ballot), and
// utf8 / binary columns as (lengths -> exclusive scan -> byte copy).
namespace pbgpu {
template <typename T>
__global__ void __launch_bounds__(256) gather_fixed_kernel(const T *__restrict__ src, const uint32_t *__restrict__ rows, int64_t n, T *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = rows[i];
  T z{};
  out[i] = r == PBGPU_NO_PARTNER ? z : src[r];
}
// bit j of out = row present and (valid == NULL or valid[rows[j]]); out has ceil(n / 32) words, n may end inside a word
__global__ void __launch_bounds__(256) gather_valid_bits_kernel(const uint8_t *__restrict__ valid, const uint32_t *__restrict__ rows, int64_t n,
                                                                uint32_t *__restrict__ out, unsigned long long *__restrict__ null_count) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  bool ok = false;
  if (i < n) {
    const uint32_t r = rows[i];
    ok = r != PBGPU_NO_PARTNER && (!valid || valid[r] != 0);
  }
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && i < n) {
    out[i >> 5] = m;
    const int64_t in_word = n - i < 32 ? n - i : 32;
    const unsigned nulls = (unsigned)in_word - (unsigned)__popc(m);
    if (nulls && null_count) atomicAdd(null_count, (unsigned long long)nulls);
  }
}
// lengths of the gathered strings (0 for nulls / missing partners), as 64-bit for the scan
__global__ void __launch_bounds__(256) gather_str_len_kernel(const long long *__restrict__ off, const uint8_t *__restrict__ valid,
                                                             const uint32_t *__restrict__ rows, int64_t n, unsigned long long *__restrict__ len) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = rows[i];
  len[i] = (r == PBGPU_NO_PARTNER || (valid && !valid[r])) ? 0ull : (unsigned long long)(off[r + 1] - off[r]);
}
// one warp per result row: copies its bytes; OffT = the output offset type (int32 utf8 / int64 large_utf8); also writes the
// Arrow offsets (n + 1 entries) from the scanned lengths
template <typename OffT>
__global__ void __launch_bounds__(256) gather_str_bytes_kernel(const long long *__restrict__ off, const char *__restrict__ chars,
                                                               const uint32_t *__restrict__ rows, int64_t n,
                                                               const unsigned long long *__restrict__ out_pos /*exclusive scan of len, n entries*/,
                                                               const unsigned long long *__restrict__ total, OffT *__restrict__ out_off,
                                                               char *__restrict__ out_chars) {
  const int64_t w = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w > n) return;
  if (w == n) { if (lane == 0) out_off[n] = (OffT)*total; return; }
  const unsigned long long p = out_pos[w];
  if (lane == 0) out_off[w] = (OffT)p;
  const uint32_t r = rows[w];
  if (r == PBGPU_NO_PARTNER) return;
  const long long a = off[r], b = off[r + 1];
  const unsigned long long next = w + 1 < n ? out_pos[w + 1] : *total;
  const long long L = (long long)(next - p);  // 0 for nulls even when the source slot holds bytes
  for (long long k = lane; k < L && a + k < b; k += 32) out_chars[p + k] = chars[a + k];
}
// per-batch Arrow offsets (int32 or int64, first entry anywhere) -> one global int64 offset column
template <typename OffT>
__global__ void __launch_bounds__(256) rebase_offsets_kernel(const OffT *__restrict__ in /*len + 1 entries*/, int64_t len, long long base,
                                                             long long *__restrict__ out /*written at [0, len] */) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i <= len) out[i] = (long long)in[i] - (long long)in[0] + base;
}

int gather_fixed(const void *d_src, int width, const uint32_t *d_rows, int64_t n, void *d_out, void *stream) {
  if (n <= 0) return PBGPU_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)cdiv(n, 256);
  switch (width) {
    case 1: PB_LAUNCH(gather_fixed_kernel<uint8_t>, grid, 256, 0, s, (const uint8_t *)d_src, d_rows, n, (uint8_t *)d_out); break;
    case 2: PB_LAUNCH(gather_fixed_kernel<uint16_t>, grid, 256, 0, s, (const uint16_t *)d_src, d_rows, n, (uint16_t *)d_out); break;
    case 4: PB_LAUNCH(gather_fixed_kernel<uint32_t>, grid, 256, 0, s, (const uint32_t *)d_src, d_rows, n, (uint32_t *)d_out); break;
    case 8: PB_LAUNCH(gather_fixed_kernel<uint2>, grid, 256, 0, s, (const uint2 *)d_src, d_rows, n, (uint2 *)d_out); break;
    case 16: PB_LAUNCH(gather_fixed_kernel<uint4>, grid, 256, 0, s, (const uint4 *)d_src, d_rows, n, (uint4 *)d_out); break;
    default: return set_error(PBGPU_EINVAL, "gather_fixed: unsupported width %d", width);
  }
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}
int gather_valid_bits(const uint8_t *d_valid, const uint32_t *d_rows, int64_t n, uint32_t *d_bits, unsigned long long *d_null_count, void *stream) {
  if (n <= 0) return PBGPU_OK;
  PB_LAUNCH(gather_valid_bits_kernel, (unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream, d_valid, d_rows, n, d_bits, d_null_count);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}
// lengths + exclusive scan: d_pos[n] (scratch, 8 bytes per row) and *d_total = bytes of the gathered column
int gather_str_plan(const long long *d_off, const uint8_t *d_valid, const uint32_t *d_rows, int64_t n, unsigned long long *d_pos,
                    unsigned long long *d_total, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n <= 0) { PB_CUDA(cudaMemsetAsync(d_total, 0, 8, s)); return PBGPU_OK; }
  PB_LAUNCH(gather_str_len_kernel, (unsigned)cdiv(n, 256), 256, 0, s, d_off, d_valid, d_rows, n, d_pos);
  PB_CHECK_LAUNCH();
  return device_scan<SumU64, false>(d_pos, d_pos, n, d_total, s);
}
int gather_str_bytes(const long long *d_off, const char *d_chars, const uint32_t *d_rows, int64_t n, const unsigned long long *d_pos,
                     const unsigned long long *d_total, void *d_out_off, int large, char *d_out_chars, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)cdiv((n + 1) * 32, 256);
  if (large) PB_LAUNCH(gather_str_bytes_kernel<long long>, grid, 256, 0, s, d_off, d_chars, d_rows, n, d_pos, d_total, (long long *)d_out_off, d_out_chars);
  else PB_LAUNCH(gather_str_bytes_kernel<int>, grid, 256, 0, s, d_off, d_chars, d_rows, n, d_pos, d_total, (int *)d_out_off, d_out_chars);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}
int rebase_offsets(const void *d_in, int large, int64_t len, long long base, long long *d_out, void *stream) {
  const unsigned grid = (unsigned)cdiv(len + 1, 256);
  if (large) PB_LAUNCH(rebase_offsets_kernel<long long>, grid, 256, 0, (cudaStream_t)stream, (const long long *)d_in, len, base, d_out);
  else PB_LAUNCH(rebase_offsets_kernel<int>, grid, 256, 0, (cudaStream_t)stream, (const int *)d_in, len, base, d_out);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}
}  // namespace pbgpu

extern "C" int pbgpu_gather_i32(const int32_t *d_src, const uint32_t *d_rows, int64_t n, int32_t *d_out, void *stream) {
  if (n < 0) return set_error(PBGPU_EINVAL, "negative n");
  if (n == 0) return PBGPU_OK;
  if (!d_src || !d_rows || !d_out) return set_error(PBGPU_EINVAL, "NULL argument");
  PB_LAUNCH(gather_i32_kernel, (unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream, d_src, d_rows, n, d_out);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}

extern "C" int pbgpu_contig_histogram(const int32_t *d_contig, int64_t n, int32_t n_contigs, int64_t *d_hist, void *stream) {
  if (n < 0 || n_contigs < 0) return set_error(PBGPU_EINVAL, "negative size");
  if (n == 0 || n_contigs == 0) return PBGPU_OK;
  if (!d_contig || !d_hist) return set_error(PBGPU_EINVAL, "NULL argument");
  int64_t grid = cdiv(n, 256 * 16);
  if (grid > kSMs * 8) grid = kSMs * 8;
  const size_t smem = n_contigs <= 4096 ? sizeof(unsigned int) * (size_t)n_contigs : 0;
  PB_LAUNCH(contig_hist_kernel, (unsigned)grid, 256, smem, (cudaStream_t)stream, d_contig, n, n_contigs, (unsigned long long *)d_hist);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}

extern "C" int pbgpu_unpack_records(const int32_t *d_packed, int64_t n, int32_t *d_contig, int32_t *d_start, int32_t *d_end,
                                    uint32_t *d_row, void *stream) {
  if (n < 0) return set_error(PBGPU_EINVAL, "negative n");
  if (n == 0) return PBGPU_OK;
  if (!d_packed || !d_contig || !d_start || !d_end || !d_row) return set_error(PBGPU_EINVAL, "NULL argument");
  PB_LAUNCH(unpack_records_kernel, (unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream, (const int4 *)d_packed, n, d_contig, d_start, d_end, d_row);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}

extern "C" int pbgpu_translate_rows(const uint32_t *d_local, int64_t n, const uint32_t *d_global_of_local, uint32_t *d_out, void *stream) {
  if (n < 0) return set_error(PBGPU_EINVAL, "negative n");
  if (n == 0) return PBGPU_OK;
  if (!d_local || !d_global_of_local || !d_out) return set_error(PBGPU_EINVAL, "NULL argument");
  PB_LAUNCH(translate_rows_kernel, (unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream, d_local, n, d_global_of_local, d_out);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}

extern "C" int pbgpu_pack_by_owner(const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t n,
                                   const int32_t *d_owner, int32_t n_contigs, int32_t n_ranks, uint32_t row_id_base,
                                   int32_t *d_packed, int64_t *d_rank_counts, void *stream) {
  if (n < 0 || n_ranks < 1 || n_ranks > 255) return set_error(PBGPU_EINVAL, "bad n / n_ranks");
  if (!d_rank_counts || !d_owner) return set_error(PBGPU_EINVAL, "NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  PB_CUDA(cudaMemsetAsync(d_rank_counts, 0, sizeof(int64_t) * (size_t)n_ranks, s));
  if (n == 0) return PBGPU_OK;
  Scratch sc(s);
  uint64_t *keys = nullptr, *vals = nullptr;
  PB_TRY(sc.get(&keys, (size_t)n));
  PB_TRY(sc.get(&vals, (size_t)n));
  {
    int64_t grid = cdiv(n, 256 * 8);
    if (grid > kSMs * 16) grid = kSMs * 16;
    PB_LAUNCH(owner_keys_kernel, (unsigned)grid, 256, 0, s, d_contig, n, d_owner, n_contigs, n_ranks, keys, vals,
              (unsigned long long *)d_rank_counts);
  }
  PB_CHECK_LAUNCH();
  PB_TRY(radix_sort_pairs(keys, vals, n, 8, s));
  if (d_packed) {
    // kept rows = everything whose key < n_ranks; they sit at the front after the sort, so packing all n
    // rows is safe only up to `kept`; compute it from the counts on the host side of the caller, or pack
    // the full prefix here: the caller allocates 4*n int32 and reads only sum(rank_counts) records.
    PB_LAUNCH(pack_records_kernel, (unsigned)cdiv(n, 256), 256, 0, s, vals, n, d_contig, d_start, d_end, row_id_base, (int4 *)d_packed);
    PB_CHECK_LAUNCH();
  }
  return PBGPU_OK;
}

#include "unary.cuh"
#include "peer.cuh"
