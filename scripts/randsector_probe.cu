// randsector_probe.cu -- what does B200 HBM deliver for the access patterns the probe side can choose between?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/randsector_probe scripts/randsector_probe.cu
//   run  : scripts/randsector_probe            (prints one line per pattern; tuning aid, not part of the product)
// Patterns (n = 100 M accesses unless noted):
//   gather32   one random 32-byte record per access out of a table of S bytes (the joint rank directory today)
//   gather32+s the same next to a 12 B/access streaming read and an 8 B/access streaming write (the count kernel)
//   binned     accesses grouped into B bins of the table in bin order, random inside the bin (what a one-pass probe
//              partition gives), with and without a bulk L2 prefetch of the bin's slice ahead of its accesses
//   scatter4   one random 4-byte store per access (counts written straight back to row order)
//   fronts16   one 16-byte store per access into K advancing fronts (direct scatter of probe records into K bins)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
struct alignas(32) Rec { uint32_t w[8]; };
__device__ __forceinline__ uint32_t ld_rec(const Rec *p) {
  uint32_t w[8];
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
  return w[0] ^ w[3] ^ w[7];
}

template <int ITEMS, bool STREAM>
__global__ void __launch_bounds__(256) gather32(const Rec *__restrict__ tab, uint64_t nrec, int64_t n, const int32_t *__restrict__ in3,
                                               int64_t *__restrict__ out) {
  const int64_t base = (int64_t)blockIdx.x * (256 * ITEMS) + threadIdx.x;
  uint32_t acc[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + (int64_t)j * 256;
    uint64_t h = mix((uint64_t)i);
    if (STREAM && i < n) h ^= (uint64_t)(in3[i] + in3[n + i] + in3[2 * n + i]);
    acc[j] = i < n ? ld_rec(tab + (h % nrec)) : 0u;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + (int64_t)j * 256;
    if (i < n) { if (STREAM) out[i] = acc[j]; else if (acc[j] == 0x12345u) out[0] = 1; }
  }
}

// bin b owns accesses [b*n/B, (b+1)*n/B) and records [b*nrec/B, (b+1)*nrec/B)
template <int ITEMS>
__global__ void __launch_bounds__(256) binned32(const Rec *__restrict__ tab, uint64_t nrec, int64_t n, int B, int prefetch_ahead,
                                               int64_t *__restrict__ out) {
  const int64_t base = (int64_t)blockIdx.x * (256 * ITEMS) + threadIdx.x;
  const uint64_t per_bin_rec = nrec / B;
  const int64_t per_bin_acc = n / B;
  if (prefetch_ahead > 0) {
    // the block prefetches its share of the slice `prefetch_ahead` bins ahead of the one it reads: every record of a bin is
    // requested once, sequentially, by the blocks that work `prefetch_ahead` bins earlier
    const int64_t blk_first = (int64_t)blockIdx.x * (256 * ITEMS);
    const int64_t b = blk_first / per_bin_acc + prefetch_ahead;
    if (b < B) {
      const int64_t blocks_per_bin = (per_bin_acc + 256 * ITEMS - 1) / (256 * ITEMS);
      const int64_t k = (blk_first % per_bin_acc) / (256 * ITEMS);  // my index among the blocks of my bin
      const uint64_t bytes = per_bin_rec * sizeof(Rec);
      const uint64_t chunk = ((bytes / blocks_per_bin) + 127) & ~127ull;
      const uint64_t lo = k * chunk;
      if (lo < bytes && threadIdx.x == 0) {
        const uint64_t len = (lo + chunk <= bytes ? chunk : bytes - lo) & ~15ull;
        const char *src = (const char *)(tab + (uint64_t)b * per_bin_rec) + lo;
        if (len) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)len) : "memory");
      }
    }
  }
  uint32_t acc[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + (int64_t)j * 256;
    const uint64_t h = mix((uint64_t)i);
    int64_t b = i / per_bin_acc;
    if (b >= B) b = B - 1;
    acc[j] = i < n ? ld_rec(tab + ((uint64_t)b * per_bin_rec + h % per_bin_rec)) : 0u;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + (int64_t)j * 256;
    if (i < n && acc[j] == 0x12345u) out[0] = 1;
  }
}

__global__ void __launch_bounds__(256) scatter4(uint32_t *__restrict__ dst, uint64_t nslots, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) dst[mix((uint64_t)i) % nslots] = (uint32_t)i;
}
// direct scatter of 16-byte records into K fronts: element i -> front hash(i) % K, slot ~ i / K
__global__ void __launch_bounds__(256) fronts16(uint4 *__restrict__ dst, int64_t n, int K) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint64_t f = mix((uint64_t)i) % (uint64_t)K;
  const int64_t cap = n / K + 1;
  dst[f * cap + i / K] = make_uint4((uint32_t)i, 1u, 2u, 3u);
}
__global__ void fill(uint32_t *p, size_t nwords) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i * 2654435761u;
}

template <typename F>
static float time_ms(F f, int reps = 3) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  const int64_t n = 100000000;
  const size_t max_tab = (size_t)3 << 30;
  Rec *tab; int32_t *in3; int64_t *out; uint32_t *sc; uint4 *fr;
  CK(cudaMalloc(&tab, max_tab));
  CK(cudaMalloc(&in3, sizeof(int32_t) * 3 * n));
  CK(cudaMalloc(&out, sizeof(int64_t) * n));
  CK(cudaMalloc(&sc, sizeof(uint32_t) * n));
  CK(cudaMalloc(&fr, sizeof(uint4) * (n + 70000)));
  fill<<<1184, 256>>>((uint32_t *)tab, max_tab / 4);
  fill<<<1184, 256>>>((uint32_t *)in3, (size_t)3 * n);
  CK(cudaDeviceSynchronize());
  const size_t sizes[] = {(size_t)32 << 20, (size_t)128 << 20, (size_t)512 << 20, (size_t)1 << 30, (size_t)3 << 30};
  for (size_t S : sizes) {
    const uint64_t nrec = S / sizeof(Rec);
    float t1 = time_ms([&] { gather32<1, false><<<(unsigned)((n + 255) / 256), 256>>>(tab, nrec, n, in3, out); });
    float t2 = time_ms([&] { gather32<2, false><<<(unsigned)((n + 511) / 512), 256>>>(tab, nrec, n, in3, out); });
    float t4 = time_ms([&] { gather32<4, false><<<(unsigned)((n + 1023) / 1024), 256>>>(tab, nrec, n, in3, out); });
    float t8 = time_ms([&] { gather32<8, false><<<(unsigned)((n + 2047) / 2048), 256>>>(tab, nrec, n, in3, out); });
    float ts = time_ms([&] { gather32<2, true><<<(unsigned)((n + 511) / 512), 256>>>(tab, nrec, n, in3, out); });
    float ts4 = time_ms([&] { gather32<4, true><<<(unsigned)((n + 1023) / 1024), 256>>>(tab, nrec, n, in3, out); });
    printf("gather32 table %5zu MB: items1 %.3f ms (%.1f G/s)  items2 %.3f (%.1f)  items4 %.3f (%.1f)  items8 %.3f (%.1f) | +stream 20B: items2 %.3f (%.1f)  items4 %.3f (%.1f)\n",
           S >> 20, t1, n / t1 * 1e-6, t2, n / t2 * 1e-6, t4, n / t4 * 1e-6, t8, n / t8 * 1e-6, ts, n / ts * 1e-6, ts4, n / ts4 * 1e-6);
  }
  {
    const size_t S = (size_t)3 << 30;
    const uint64_t nrec = S / sizeof(Rec);
    const int Bs[] = {64, 256, 1024, 4096};
    for (int B : Bs) {
      for (int pf : {0, 1, 2, 4}) {
        float t = time_ms([&] { binned32<2><<<(unsigned)((n + 511) / 512), 256>>>(tab, nrec, n, B, pf, out); });
        printf("binned32 table 3072 MB, %4d bins (%.1f MB each), prefetch %d bins ahead: %.3f ms (%.1f G/s)\n", B, (double)S / B / 1048576.0, pf, t,
               n / t * 1e-6);
      }
    }
  }
  {
    float t = time_ms([&] { scatter4<<<(unsigned)((n + 255) / 256), 256>>>(sc, (uint64_t)n, n); });
    printf("scatter4 into 400 MB: %.3f ms (%.1f G/s)\n", t, n / t * 1e-6);
    for (int K : {256, 4096, 16384, 65536}) {
      float tf = time_ms([&] { fronts16<<<(unsigned)((n + 255) / 256), 256>>>(fr, n, K); });
      printf("fronts16 %6d fronts: %.3f ms (%.1f G rec/s, %.0f GB/s written)\n", K, tf, n / tf * 1e-6, 16.0 * n / tf * 1e-6);
    }
  }
  return 0;
}
