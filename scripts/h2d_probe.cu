// h2d_probe.cu -- tuning aid: where do 1.6 ms go when 3 x 4 MB are copied H2D at the start of pbgpu_range_op?
// Build: nvcc -O2 -o scripts/h2d_probe scripts/h2d_probe.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  const size_t n = 4u << 20;
  void *h_plain, *h_wc, *d[3];
  cudaHostAlloc(&h_plain, 3 * n, cudaHostAllocDefault);
  cudaHostAlloc(&h_wc, 3 * n, cudaHostAllocWriteCombined);
  for (auto &p : d) cudaMalloc(&p, n);
  cudaStream_t s0;
  cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking);
  int mode = 0;  // how the CPU writes: 0 memset, 1 non-temporal stores, 2 memset + clflushopt, 3 memset on the calling thread only
  auto run = [&](const char *what, void *h, bool fresh_stream, int idle_ms, bool touch) {
    for (int rep = 0; rep < 4; ++rep) {
      if (touch) {  // CPU writes the staging buffer from 8 threads, as the encoder does
        std::vector<std::thread> th;
        auto body = [&](int t) {
          char *p = (char *)h + t * (3 * n / 8);
          const size_t len = 3 * n / 8;
          if (mode == 1) {
            const __m128i v = _mm_set1_epi8((char)(rep + t));
            for (size_t i = 0; i < len; i += 16) _mm_stream_si128((__m128i *)(p + i), v);
            _mm_sfence();
          } else {
            memset(p, rep + t, len);
            if (mode == 2) { for (size_t i = 0; i < len; i += 64) _mm_clflushopt(p + i); _mm_sfence(); }
          }
        };
        if (mode == 3) { for (int t = 0; t < 8; ++t) body(t); }
        else {
          for (int t = 0; t < 8; ++t) th.emplace_back(body, t);
          for (auto &x : th) x.join();
        }
      }
      if (idle_ms) std::this_thread::sleep_for(std::chrono::milliseconds(idle_ms));
      cudaStream_t s = s0;
      if (fresh_stream) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
      const double t0 = now_ms();
      for (int k = 0; k < 3; ++k) cudaMemcpyAsync(d[k], (char *)h + k * n, n, cudaMemcpyHostToDevice, s);
      const double t1 = now_ms();
      cudaStreamSynchronize(s);
      const double t2 = now_ms();
      if (rep) printf("%-44s enqueue %.3f ms  drain %.3f ms  (%.1f GB/s)\n", what, t1 - t0, t2 - t1, 3 * n / (t2 - t0) / 1e6);
      if (fresh_stream) cudaStreamDestroy(s);
    }
  };
  run("plain pinned, same stream, no idle", h_plain, false, 0, false);
  run("WC pinned, same stream, no idle", h_wc, false, 0, false);
  run("WC pinned, fresh stream, no idle", h_wc, true, 0, false);
  run("WC pinned, same stream, 5 ms idle", h_wc, false, 5, false);
  run("WC pinned, same stream, 50 ms idle", h_wc, false, 50, false);
  run("plain pinned, same stream, 50 ms idle", h_plain, false, 50, false);
  run("WC pinned, CPU-written, no idle", h_wc, false, 0, true);
  run("plain pinned, CPU-written, no idle", h_plain, false, 0, true);
  run("WC pinned, CPU-written, fresh stream, 5 ms", h_wc, true, 5, true);
  mode = 1;
  run("plain pinned, NT stores", h_plain, false, 0, true);
  run("WC pinned, NT stores", h_wc, false, 0, true);
  mode = 2;
  run("plain pinned, memset + clflushopt", h_plain, false, 0, true);
  mode = 3;
  run("plain pinned, memset on calling thread", h_plain, false, 0, true);
  run("WC pinned, memset on calling thread", h_wc, false, 0, true);
  mode = 0;
  {  // larger buffer: 96 MB written by 8 threads, then copied
    const size_t big = 96u << 20;
    void *hb, *db;
    cudaHostAlloc(&hb, big, cudaHostAllocDefault);
    cudaMalloc(&db, big);
    for (int rep = 0; rep < 3; ++rep) {
      std::vector<std::thread> th;
      for (int t = 0; t < 8; ++t) th.emplace_back([&, t] { memset((char *)hb + t * (big / 8), rep + t, big / 8); });
      for (auto &x : th) x.join();
      const double t0 = now_ms();
      cudaMemcpyAsync(db, hb, big, cudaMemcpyHostToDevice, s0);
      cudaStreamSynchronize(s0);
      const double t2 = now_ms();
      printf("plain pinned 96 MB, CPU-written             %.3f ms (%.1f GB/s)\n", t2 - t0, big / (t2 - t0) / 1e6);
    }
  }
  return 0;
}
