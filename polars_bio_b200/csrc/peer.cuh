// peer.cuh -- the contig exchange as kernels over NVLink peer memory (one process per GPU, one node).
//
// rows arrive on arbitrary ranks and must reach the rank that owns their contig before the local join.
// The NCCL form of that step is pack (bucket by owner) -> all_to_all of 16-byte records -> unpack, with three host
// round trips (dist.py: shard_tables, kept as the fallback).  Here every rank maps every other rank's receive arena
// (CUDA IPC; NVSwitch gives full bandwidth to every peer) and ONE kernel per table reads the rank's slice once and
// stores each row straight into the column arrays of its owner: pack + transfer + unpack fused, no staging copy of
// the records, no collective on the data path.
//
//   peer_hist_kernel     per-contig row counts of this rank's slice (+ the slice size)         -> all_gather (tiny)
//   peer_plan_kernel     ON THE DEVICE, identically on every rank: contig -> owner by LPT bin packing (same table as
//                        dist.owner_table), rows(source, table, destination), this source's region in every
//                        destination's arena (regions are laid out in source order, so the received rows are ordered
//                        by global row id exactly as after the stable NCCL exchange), received rows, id bases, overflow
//   peer_block_count / peer_block_scan / peer_scatter_kernel
//                        1024 rows per block: destination counts per block, exclusive scan per destination, then every
//                        row takes its slot (warp multisplit + per-warp prefix: stable) and the block writes position
//                        by position out of shared memory, so the remote stores of a warp are runs of consecutive
//                        4-byte elements of one column (NVLink packets carry whole sectors), not 32 scattered rows.
//
// The host never waits between these launches: the only host read (received rows per table, needed to size the join)
// overlaps the scatter.  One tiny all_reduce afterwards orders "all peers have written" before "I read".
#pragma once
#include "common.cuh"

namespace pbgpu {

constexpr int kPeerThreads = 256;
constexpr int kPeerItems = 4;
constexpr int kPeerTile = kPeerThreads * kPeerItems;
constexpr int kPeerWarps = kPeerThreads / 32;
constexpr int kPeerMaxRanks = 16;
constexpr int kPeerMaxTables = 4;

struct PeerDst {  // per destination rank: this source's region there, one pointer per column
  int32_t *contig;
  int32_t *start;
  int32_t *end;
  uint32_t *row;
};

struct PeerPlanArgs {
  unsigned long long arena[kPeerMaxRanks];  // base address of every destination's arena in THIS process
  long long cap[kPeerMaxTables];            // rows per column of table t
  long long tab_off[kPeerMaxTables];        // byte offset of table t inside an arena: columns contig|start|end|row
};

// rows per contig (null keys ignored) of one slice, added onto hist[0..n_contigs); hist[n_contigs] = slice size
__global__ void __launch_bounds__(256) peer_hist_kernel(const int32_t *__restrict__ c, int64_t n, int32_t n_contigs,
                                                        unsigned long long *__restrict__ hist) {
  extern __shared__ unsigned int bins[];
  const bool use_smem = n_contigs <= 4096;
  if (blockIdx.x == 0 && threadIdx.x == 0) hist[n_contigs] = (unsigned long long)n;
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

// g: [world][T][n_contigs+1] all-gathered histograms.  result: int64 [3T+1] = received rows[T] | global row id base[T] |
// largest region any destination needs[T] | overflow flag.
__global__ void __launch_bounds__(1024) peer_plan_kernel(const long long *__restrict__ g, int world, int rank, int T, int nc,
                                                         PeerPlanArgs a, unsigned long long *__restrict__ w, int32_t *__restrict__ order,
                                                         int32_t *__restrict__ owner, PeerDst *__restrict__ dst, long long *__restrict__ result) {
  __shared__ unsigned long long cnt[kPeerMaxRanks * kPeerMaxTables * kPeerMaxRanks];
  __shared__ unsigned long long load[kPeerMaxRanks];
  __shared__ unsigned long long need[kPeerMaxTables];
  __shared__ int overflow;
  const int tid = threadIdx.x, stride = nc + 1;
  for (int i = tid; i < world * T * world; i += blockDim.x) cnt[i] = 0;
  if (tid < kPeerMaxTables) need[tid] = 0;
  if (tid == 0) overflow = 0;
  for (int c = tid; c < nc; c += blockDim.x) {
    unsigned long long sum = 0;
    for (int st = 0; st < world * T; ++st) sum += (unsigned long long)g[(size_t)st * stride + c];
    w[c] = sum;
  }
  __syncthreads();
  // position of every contig in (weight descending, id ascending) order
  for (int c = tid; c < nc; c += blockDim.x) {
    const unsigned long long wc = w[c];
    int r = 0;
    for (int j = 0; j < nc; ++j) {
      const unsigned long long wj = w[j];
      r += (wj > wc) || (wj == wc && j < c);
    }
    order[r] = c;
  }
  __syncthreads();
  // longest-processing-time-first bin packing, ties to the lowest rank (dist.owner_table)
  if (tid == 0) {
    for (int k = 0; k < world; ++k) load[k] = 0;
    for (int i = 0; i < nc; ++i) {
      const int c = order[i];
      int best = 0;
      for (int k = 1; k < world; ++k) if (load[k] < load[best]) best = k;
      owner[c] = best;
      load[best] += w[c];
    }
  }
  __syncthreads();
  for (int c = tid; c < nc; c += blockDim.x) {
    const int d = owner[c];
    for (int st = 0; st < world * T; ++st) {
      const unsigned long long v = (unsigned long long)g[(size_t)st * stride + c];
      if (v) atomicAdd(&cnt[st * world + d], v);
    }
  }
  __syncthreads();
  if (tid < T * world) {
    const int t = tid / world, d = tid % world;
    unsigned long long run = 0, mine = 0;
    for (int s = 0; s < world; ++s) {
      if (s == rank) mine = run;
      run += cnt[(s * T + t) * world + d];
    }
    if ((long long)run > a.cap[t]) overflow = 1;
    atomicMax(&need[t], run);
    const long long cap = a.cap[t];
    int32_t *base = (int32_t *)((char *)a.arena[d] + a.tab_off[t]);
    dst[t * world + d] = PeerDst{base + mine, base + cap + mine, base + 2 * cap + mine, (uint32_t *)(base + 3 * cap + mine)};
    if (d == rank) result[t] = (long long)run;
  }
  if (tid < T) {
    long long b = 0;
    for (int s = 0; s < rank; ++s) b += g[(size_t)(s * T + tid) * stride + nc];
    result[T + tid] = b;
  }
  __syncthreads();
  if (tid < T) result[2 * T + tid] = (long long)need[tid];
  if (tid == 0) result[3 * T] = overflow;
}

__device__ __forceinline__ int peer_dest(int32_t cc, const int32_t *__restrict__ owner, int32_t n_contigs, int32_t n_ranks) {
  if (cc < 0 || cc >= n_contigs) return -1;
  const int32_t o = __ldg(owner + cc);
  return (o >= 0 && o < n_ranks) ? o : -1;
}

// bc[d][blk] = rows of block blk that go to destination d
__global__ void __launch_bounds__(kPeerThreads) peer_block_count_kernel(const int32_t *__restrict__ c, int64_t n, const int32_t *__restrict__ owner,
                                                                        int32_t n_contigs, int32_t n_ranks, const long long *__restrict__ flag,
                                                                        unsigned int *__restrict__ bc, int64_t nblk) {
  __shared__ unsigned int bcount[kPeerMaxRanks];
  if (*flag) return;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < kPeerMaxRanks) bcount[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kPeerTile;
#pragma unroll
  for (int j = 0; j < kPeerItems; ++j) {
    const int64_t i = base + (int64_t)j * kPeerThreads + threadIdx.x;
    const int d = i < n ? peer_dest(c[i], owner, n_contigs, n_ranks) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (d >= 0 && lane == __ffs(peers) - 1) atomicAdd(&bcount[d], (unsigned)__popc(peers));
  }
  __syncthreads();
  if (threadIdx.x < n_ranks) bc[(int64_t)threadIdx.x * nblk + blockIdx.x] = bcount[threadIdx.x];
}

// in-place exclusive scan of every destination's row bc[d][0..nblk): one block per destination
__global__ void __launch_bounds__(1024) peer_block_scan_kernel(unsigned int *__restrict__ bc, int64_t nblk, const long long *__restrict__ flag) {
  __shared__ unsigned int wsum[32];
  __shared__ unsigned int carry;
  if (*flag) return;
  unsigned int *row = bc + (int64_t)blockIdx.x * nblk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t t0 = 0; t0 < nblk; t0 += 1024) {
    const int64_t i = t0 + threadIdx.x;
    const unsigned int v = i < nblk ? row[i] : 0u;
    unsigned int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      unsigned int s = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
      wsum[lane] = s;
    }
    __syncthreads();
    const unsigned int before = carry + (warp ? wsum[warp - 1] : 0u) + x - v;
    if (i < nblk) row[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kPeerThreads) peer_scatter_kernel(const int32_t *__restrict__ c, const int32_t *__restrict__ s,
                                                                    const int32_t *__restrict__ e, int64_t n, const int32_t *__restrict__ owner,
                                                                    int32_t n_contigs, int32_t n_ranks, const long long *__restrict__ row_id_base,
                                                                    const PeerDst *__restrict__ dst_table, const unsigned int *__restrict__ bc,
                                                                    int64_t nblk, const long long *__restrict__ flag) {
  __shared__ int32_t sc[kPeerTile], ss[kPeerTile], se[kPeerTile];
  __shared__ uint32_t sr[kPeerTile];
  __shared__ unsigned int wcnt[kPeerItems * kPeerWarps][kPeerMaxRanks];  // rows of (item, warp) per destination -> exclusive prefix
  __shared__ unsigned int boff[kPeerMaxRanks + 1];
  __shared__ unsigned int gbase[kPeerMaxRanks];
  __shared__ PeerDst dst[kPeerMaxRanks];
  if (*flag) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = lanemask_lt();
  for (int i = threadIdx.x; i < kPeerItems * kPeerWarps * kPeerMaxRanks; i += kPeerThreads) (&wcnt[0][0])[i] = 0;
  if (threadIdx.x < n_ranks) {
    dst[threadIdx.x] = dst_table[threadIdx.x];
    gbase[threadIdx.x] = bc[(int64_t)threadIdx.x * nblk + blockIdx.x];
  }
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kPeerTile;
  const uint32_t id0 = (uint32_t)*row_id_base;
  int32_t rc[kPeerItems], rs[kPeerItems], re[kPeerItems];
  int rd[kPeerItems];
  unsigned int slot[kPeerItems];
#pragma unroll
  for (int j = 0; j < kPeerItems; ++j) {
    const int64_t i = base + (int64_t)j * kPeerThreads + threadIdx.x;
    rd[j] = -1;
    if (i < n) {
      rc[j] = c[i]; rs[j] = s[i]; re[j] = e[i];
      rd[j] = peer_dest(rc[j], owner, n_contigs, n_ranks);
    }
  }
  // rank of every row among the rows of its (item, warp) with the same destination; (item, warp, lane) order = row order
#pragma unroll
  for (int j = 0; j < kPeerItems; ++j) {
    const unsigned peers = __match_any_sync(0xffffffffu, rd[j]);
    if (rd[j] >= 0 && lane == __ffs(peers) - 1) wcnt[j * kPeerWarps + warp][rd[j]] = (unsigned)__popc(peers);
    slot[j] = (unsigned)__popc(peers & lt);
  }
  __syncthreads();
  if (threadIdx.x < n_ranks) {  // exclusive prefix over the 32 (item, warp) groups of this destination
    unsigned int run = 0;
    for (int k = 0; k < kPeerItems * kPeerWarps; ++k) { const unsigned int v = wcnt[k][threadIdx.x]; wcnt[k][threadIdx.x] = run; run += v; }
    boff[threadIdx.x + 1] = run;  // block total, turned into offsets below
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int run = 0;
    for (int r = 0; r < n_ranks; ++r) { const unsigned int v = boff[r + 1]; boff[r] = run; run += v; }
    boff[n_ranks] = run;
  }
  __syncthreads();
  // stage in (destination, slot) order
#pragma unroll
  for (int j = 0; j < kPeerItems; ++j) {
    if (rd[j] < 0) continue;
    const unsigned int p = boff[rd[j]] + wcnt[j * kPeerWarps + warp][rd[j]] + slot[j];
    const int64_t i = base + (int64_t)j * kPeerThreads + threadIdx.x;
    sc[p] = rc[j]; ss[p] = rs[j]; se[p] = re[j]; sr[p] = id0 + (uint32_t)i;
  }
  __syncthreads();
  // write out: consecutive positions of one destination are consecutive elements of its region
  const unsigned int total = boff[n_ranks];
  for (unsigned int p = threadIdx.x; p < total; p += kPeerThreads) {
    int r = 0;
    while (p >= boff[r + 1]) ++r;
    const unsigned long long k = (unsigned long long)gbase[r] + (p - boff[r]);
    const PeerDst d = dst[r];
    d.contig[k] = sc[p]; d.start[k] = ss[p]; d.end[k] = se[p]; d.row[k] = sr[p];
  }
}

}  // namespace pbgpu

extern "C" {

// receive arena: plain cudaMalloc memory (legacy CUDA IPC cannot export stream-ordered pool memory)
int pbgpu_peer_alloc(size_t bytes, void **d_ptr, unsigned char *handle_out /*[64]*/) {
  if (!d_ptr || !handle_out) return set_error(PBGPU_EINVAL, "NULL argument");
  *d_ptr = nullptr;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(e == cudaErrorMemoryAllocation ? PBGPU_ENOMEM : PBGPU_ECUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  }
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaFree(p);
    return set_error(PBGPU_ECUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  }
  memcpy(handle_out, &h, 64);
  *d_ptr = p;
  return PBGPU_OK;
}
int pbgpu_peer_free(void *d_ptr) {
  if (d_ptr) PB_CUDA(cudaFree(d_ptr));
  return PBGPU_OK;
}
// map another process's arena into this one (peer access is enabled on demand)
int pbgpu_peer_open(const unsigned char *handle /*[64]*/, void **d_ptr) {
  if (!d_ptr || !handle) return set_error(PBGPU_EINVAL, "NULL argument");
  *d_ptr = nullptr;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  cudaError_t e = cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *d_ptr = nullptr;
    return set_error(PBGPU_ECUDA, "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
  }
  return PBGPU_OK;
}
int pbgpu_peer_close(void *d_ptr) {
  if (d_ptr) PB_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return PBGPU_OK;
}

int pbgpu_peer_histogram(const int32_t *d_contig, int64_t n, int32_t n_contigs, int64_t *d_hist, void *stream) {
  if (n < 0 || n_contigs < 0) return set_error(PBGPU_EINVAL, "negative size");
  if (!d_hist) return set_error(PBGPU_EINVAL, "NULL argument");
  if (n > 0 && !d_contig) return set_error(PBGPU_EINVAL, "NULL column");
  int64_t grid = cdiv(n, 256 * 16);
  if (grid > kSMs * 8) grid = kSMs * 8;
  if (grid < 1) grid = 1;
  const size_t smem = n_contigs <= 4096 ? sizeof(unsigned int) * (size_t)n_contigs : 0;
  PB_LAUNCH(peer_hist_kernel, (unsigned)grid, 256, smem, (cudaStream_t)stream, d_contig, n, n_contigs, (unsigned long long *)d_hist);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}

int pbgpu_peer_plan(const int64_t *d_gathered, int32_t world, int32_t rank, int32_t n_tables, int32_t n_contigs,
                    const uint64_t *arena_base, const int64_t *cap_rows, int32_t *d_owner, void *d_dst, int64_t *d_result, void *stream) {
  if (world < 1 || world > kPeerMaxRanks) return set_error(PBGPU_EINVAL, "bad world (at most %d ranks)", kPeerMaxRanks);
  if (n_tables < 1 || n_tables > kPeerMaxTables) return set_error(PBGPU_EINVAL, "bad n_tables (at most %d)", kPeerMaxTables);
  if (rank < 0 || rank >= world || n_contigs < 0) return set_error(PBGPU_EINVAL, "bad rank / n_contigs");
  if (!d_gathered || !arena_base || !cap_rows || !d_owner || !d_dst || !d_result) return set_error(PBGPU_EINVAL, "NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  PeerPlanArgs a;
  memset(&a, 0, sizeof(a));
  for (int d = 0; d < world; ++d) a.arena[d] = arena_base[d];
  long long off = 0;
  for (int t = 0; t < n_tables; ++t) {
    if (cap_rows[t] < 0 || (cap_rows[t] & 63)) return set_error(PBGPU_EINVAL, "cap_rows must be non-negative multiples of 64");
    a.cap[t] = cap_rows[t];
    a.tab_off[t] = off;
    off += 16ll * cap_rows[t];
  }
  Scratch sc(s);
  unsigned long long *w = nullptr;
  int32_t *order = nullptr;
  PB_TRY(sc.get(&w, (size_t)n_contigs + 1));
  PB_TRY(sc.get(&order, (size_t)n_contigs + 1));
  PB_LAUNCH(peer_plan_kernel, 1, 1024, 0, s, (const long long *)d_gathered, world, rank, n_tables, n_contigs, a, w, order, d_owner,
            (PeerDst *)d_dst, (long long *)d_result);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}

// Every valid row of this rank's slice goes to the column arrays of its owner.  d_dst: n_ranks records of four device
// pointers (contig, start, end, row) = the start of THIS source's region at each destination (own arena or mapped peer
// memory), d_row_id_base: the global id of row 0, d_flag: non-zero = do nothing (arena overflow) -- all three are
// outputs of pbgpu_peer_plan and are read on the device, so the host enqueues plan and scatter back to back.
int pbgpu_peer_scatter(const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t n, const int32_t *d_owner,
                       int32_t n_contigs, int32_t n_ranks, const int64_t *d_row_id_base, const void *d_dst, const int64_t *d_flag,
                       void *stream) {
  if (n < 0 || n >= (1ll << 32) || n_ranks < 1 || n_ranks > kPeerMaxRanks) return set_error(PBGPU_EINVAL, "bad n / n_ranks (at most %d ranks)", kPeerMaxRanks);
  if (!d_owner || !d_dst || !d_row_id_base || !d_flag) return set_error(PBGPU_EINVAL, "NULL argument");
  if (n == 0) return PBGPU_OK;
  if (!d_contig || !d_start || !d_end) return set_error(PBGPU_EINVAL, "NULL column");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t nblk = cdiv(n, kPeerTile);
  Scratch sc(s);
  unsigned int *bc = nullptr;
  PB_TRY(sc.get(&bc, (size_t)nblk * n_ranks));
  PB_LAUNCH(peer_block_count_kernel, (unsigned)nblk, kPeerThreads, 0, s, d_contig, n, d_owner, n_contigs, n_ranks, (const long long *)d_flag, bc, nblk);
  PB_LAUNCH(peer_block_scan_kernel, (unsigned)n_ranks, 1024, 0, s, bc, nblk, (const long long *)d_flag);
  PB_LAUNCH(peer_scatter_kernel, (unsigned)nblk, kPeerThreads, 0, s, d_contig, d_start, d_end, n, d_owner, n_contigs, n_ranks,
            (const long long *)d_row_id_base, (const PeerDst *)d_dst, bc, nblk, (const long long *)d_flag);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}

}  // extern "C"
