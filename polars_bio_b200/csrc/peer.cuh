// peer.cuh -- the contig exchange as kernels over NVLink peer memory (one process per GPU, one node).
//
// rows arrive on arbitrary ranks and must reach the rank that owns their contig before the local join.
// The NCCL form of that step is pack (bucket by owner) -> all_to_all of 16-byte records -> unpack, with three host
// round trips (dist.py: shard_tables, kept as the fallback).  Here every rank maps every other rank's receive arena
// (CUDA IPC; NVSwitch gives full bandwidth to every peer) and ONE kernel per table reads the rank's slice once and
// stores each row straight into the column arrays of its owner: pack + transfer + unpack fused, no staging copy of
// the records, no collective on the data path.
//
//   peer_hist_publish_kernel  per-contig row counts of this rank's slices (+ the slice sizes), all tables in one launch
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
// overlaps the scatter.
//
// Synchronisation is also done through peer memory: every rank owns a small CONTROL block that all peers map --
// flag words written remotely with st.release.sys and polled locally with ld.acquire.sys (bounded spin: a dead peer
// turns into an error code after the timeout, never into a hung GPU) plus a slot per source for the histograms:
//   peer_hist_publish_kernel  the block that finishes last stores this rank's histograms into every peer's control
//                             block, then raises flag A[rank] there
//   peer_plan_kernel          first waits until flag A of every source shows this step
//   peer_signal_wait_kernel   (after the scatter of table t, same stream) raises flag B[t][rank] at every destination,
//                             then waits until flag B[t] of every source shows this step: table t is complete here
// so a step needs no NCCL call at all, and each table can run its scatter / signal / wait on its own stream (the index
// build over the small table overlaps the scatter of the big one).  Without control blocks (PBGPU_PEER_SYNC=nccl) the
// caller all-gathers the histograms and closes the step with a tiny all_reduce.
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

// control block: flag A [kPeerMaxRanks] u64 | flag B [kPeerMaxTables][kPeerMaxRanks] u64 | pad to 1024 B | histograms
// int64 [world][T][n_contigs+1] (slot s written by source s)
constexpr size_t kPeerCtlFlagB = sizeof(unsigned long long) * kPeerMaxRanks;
constexpr size_t kPeerCtlHist = 1024;
static_assert(kPeerCtlFlagB + sizeof(unsigned long long) * kPeerMaxTables * kPeerMaxRanks <= kPeerCtlHist, "control block header");

struct PeerCtlArgs {
  unsigned long long ctl[kPeerMaxRanks];  // base address of every rank's control block in THIS process
};

__device__ __forceinline__ unsigned long long peer_ld_flag(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_st_flag(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peer_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// wait until *p >= want; false after timeout_ns
__device__ __forceinline__ bool peer_spin(const unsigned long long *p, unsigned long long want, unsigned long long timeout_ns) {
  if (peer_ld_flag(p) >= want) return true;
  const unsigned long long t0 = peer_now_ns();
  while (peer_ld_flag(p) < want) {
    if (peer_now_ns() - t0 > timeout_ns) return false;
    __nanosleep(64);
  }
  return true;
}

struct PeerTablesArgs {
  const int32_t *c[kPeerMaxTables];
  long long n[kPeerMaxTables];
};

// The histograms of all tables in one launch (blockIdx.y = table) into hist[T][n_contigs+1] (+ one counter word behind
// it, zeroed with the rest); the block that finishes last publishes them (publish != 0): slot `rank` of every control
// block, then flag A[rank] = step there.
__global__ void __launch_bounds__(256) peer_hist_publish_kernel(PeerTablesArgs tb, int T, int32_t n_contigs, unsigned long long *__restrict__ hist_all,
                                                                PeerCtlArgs a, int world, int rank, unsigned long long step, int publish) {
  extern __shared__ unsigned int bins[];
  __shared__ bool last;
  const int t = blockIdx.y;
  const int32_t *__restrict__ c = tb.c[t];
  const int64_t n = tb.n[t];
  unsigned long long *hist = hist_all + (size_t)t * (n_contigs + 1);
  const bool use_smem = n_contigs <= 4096;
  if (blockIdx.x == 0 && threadIdx.x == 0) hist[n_contigs] = (unsigned long long)n;
  if (use_smem) { for (int i = threadIdx.x; i < n_contigs; i += blockDim.x) bins[i] = 0; __syncthreads(); }
  // a genome has a few dozen contigs: 32 lanes adding to two or three bins would serialise (r02f: 0.053 ms for 11 M
  // rows), so the lanes of a warp that hold the same contig add once, together (block-uniform loop: every lane takes
  // part in every match)
  const int lane = threadIdx.x & 31;
  for (int64_t b0 = (int64_t)blockIdx.x * blockDim.x; b0 < n; b0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = b0 + threadIdx.x;
    int32_t cc = i < n ? c[i] : -1;
    if (cc < 0 || cc >= n_contigs) cc = -1;
    const unsigned peers = __match_any_sync(0xffffffffu, cc);
    if (cc >= 0 && lane == __ffs(peers) - 1) {
      if (use_smem) atomicAdd(&bins[cc], (unsigned)__popc(peers)); else atomicAdd(hist + cc, (unsigned long long)__popc(peers));
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < n_contigs; i += blockDim.x) if (bins[i]) atomicAdd(hist + i, (unsigned long long)bins[i]);
  }
  if (!publish) return;
  __threadfence();
  __syncthreads();
  const int len = T * (n_contigs + 1);
  if (threadIdx.x == 0) last = atomicAdd((unsigned int *)(hist_all + len), 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int d = 0; d < world; ++d) {
    long long *slot = (long long *)((char *)a.ctl[d] + kPeerCtlHist) + (size_t)rank * len;
    for (int i = threadIdx.x; i < len; i += blockDim.x) slot[i] = (long long)__ldcg(hist_all + i);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) {
    __threadfence_system();
    peer_st_flag((unsigned long long *)a.ctl[threadIdx.x] + rank, step);
  }
}

// mode & 1: flag B[table][rank] = step at every destination (runs after the scatter of `table` on the same stream: the
// kernel boundary orders the scatter's remote stores before these); mode & 2: wait until every source's flag B[table]
// shows `step` here -- the table is complete on this rank; a timeout is reported through *status (= 2)
__global__ void __launch_bounds__(32) peer_signal_wait_kernel(PeerCtlArgs a, int world, int rank, int table, unsigned long long step,
                                                              unsigned long long timeout_ns, long long *__restrict__ status, int mode) {
  if (threadIdx.x < world) {
    const size_t off = (size_t)table * kPeerMaxRanks;
    if (mode & 1) {
      __threadfence_system();
      peer_st_flag((unsigned long long *)((char *)a.ctl[threadIdx.x] + kPeerCtlFlagB) + off + rank, step);
    }
    if (mode & 2) {
      const unsigned long long *f = (const unsigned long long *)((const char *)a.ctl[rank] + kPeerCtlFlagB) + off + threadIdx.x;
      if (!peer_spin(f, step, timeout_ns)) *status = 2;
    }
  }
}

// g: [world][T][n_contigs+1] all-gathered histograms.  result: int64 [3T+1] = received rows[T] | global row id base[T] |
// largest region any destination needs[T] | overflow flag.
// wait_flags != NULL: g is the histogram area of this rank's control block; wait until flag A of every source shows
// `step` before reading it (timeout -> result[3T] = 2 and nothing else is written).
__global__ void __launch_bounds__(1024) peer_plan_kernel(const long long *g /* peers write it: coherent loads only */, int world, int rank, int T, int nc,
                                                         PeerPlanArgs a, unsigned long long *__restrict__ w, int32_t *__restrict__ order,
                                                         int32_t *__restrict__ owner, PeerDst *__restrict__ dst, long long *__restrict__ result,
                                                         const unsigned long long *__restrict__ wait_flags, unsigned long long step,
                                                         unsigned long long timeout_ns, long long *host_result) {
  __shared__ unsigned long long cnt[kPeerMaxRanks * kPeerMaxTables * kPeerMaxRanks];
  __shared__ unsigned long long load[kPeerMaxRanks];
  __shared__ unsigned long long need[kPeerMaxTables];
  __shared__ int overflow;
  __shared__ int timed_out;
  const int tid = threadIdx.x, stride = nc + 1;
  if (wait_flags) {
    if (tid == 0) timed_out = 0;
    __syncthreads();
    if (tid < world && !peer_spin(wait_flags + tid, step, timeout_ns)) timed_out = 1;
    __syncthreads();
    if (timed_out) {
      if (tid == 0) {
        result[3 * T] = 2;
        if (host_result) {
          host_result[3 * T] = 2;
          __threadfence_system();
          *(volatile long long *)(host_result + 3 * T + 1) = (long long)step;
        }
      }
      return;
    }
  }
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
  // the same numbers straight into page-locked host memory, sequence word last: the host spins on it instead of
  // paying a copy-engine round trip (it needs the received rows to size the join)
  if (host_result) {
    __syncthreads();
    if (tid == 0) {
      for (int i = 0; i <= 3 * T; ++i) host_result[i] = result[i];
      __threadfence_system();
      *(volatile long long *)(host_result + 3 * T + 1) = (long long)step;
    }
  }
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

// One 1024-row tile per loop turn; the grid is either one block per tile or a few blocks per SM striding over the tiles
// (the kernel is bound by the NVLink stores, not by occupancy: a small resident grid leaves room for the kernels of the
// stream that overlaps it).  Tile t of the table always lands at bc[d][t], whichever block handles it.
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
  if (threadIdx.x < n_ranks) dst[threadIdx.x] = dst_table[threadIdx.x];
  const uint32_t id0 = (uint32_t)*row_id_base;
  for (int64_t tile = blockIdx.x; tile < nblk; tile += gridDim.x) {
    for (int i = threadIdx.x; i < kPeerItems * kPeerWarps * kPeerMaxRanks; i += kPeerThreads) (&wcnt[0][0])[i] = 0;
    if (threadIdx.x < n_ranks) gbase[threadIdx.x] = bc[(int64_t)threadIdx.x * nblk + tile];
    __syncthreads();
    const int64_t base = tile * kPeerTile;
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
    __syncthreads();  // the staging arrays and counters are rewritten by the next tile
  }
}


// ---- round 2: the same scatter with TMA bulk stores over NVLink --------------------------------------------------------
// peer_scatter_kernel leaves every (tile, destination) run as 4-byte-per-lane stores: a warp instruction covers 128 bytes
// that start anywhere, so most of them straddle two 128-byte lines of the remote arena and travel as two NVLink packets;
// measured ~400 GB/s of egress per GPU at N = 8 (770 GB/s peer copies).  Here a tile is 4096 rows (runs of ~512 rows per
// destination and column at N = 8), and the rows of every destination are staged in shared memory at an offset with the
// SAME residue mod 4 rows as their destination index (arenas are 2 MiB aligned, table offsets multiples of 1 KiB, column
// capacities multiples of 64 rows: the four columns of a run share that residue).  The 16-byte-aligned body of every
// column run then leaves as ONE cp.async.bulk shared -> global (SASS UBLKCP) straight into the owner's column on the
// peer GPU; the <= 3 head and <= 3 tail rows are stored by the issuing thread.  One thread per (destination, column).
constexpr int kPbThreads = 512, kPbItems = 8, kPbTile = kPbThreads * kPbItems, kPbWarps = kPbThreads / 32;
constexpr int kPbSub = kPbTile / kPeerTile;           // count-kernel blocks (1024 rows) per bulk tile
constexpr int kPbStageRows = kPbTile + 8 * kPeerMaxRanks;  // + alignment gaps between the destinations' segments
__global__ void __launch_bounds__(kPbThreads) peer_scatter_bulk_kernel(const int32_t *__restrict__ c, const int32_t *__restrict__ s,
                                                                       const int32_t *__restrict__ e, int64_t n, const int32_t *__restrict__ owner,
                                                                       int32_t n_contigs, int32_t n_ranks, const long long *__restrict__ row_id_base,
                                                                       const PeerDst *__restrict__ dst_table, const unsigned int *__restrict__ bc,
                                                                       int64_t nblk /*1024-row blocks*/, const long long *__restrict__ flag) {
  extern __shared__ __align__(16) unsigned char pb_smem[];
  int32_t *col[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) col[k] = reinterpret_cast<int32_t *>(pb_smem) + (size_t)k * kPbStageRows;
  __shared__ unsigned short wcnt[kPbItems * kPbWarps][kPeerMaxRanks];  // rows of (item, warp) per destination -> exclusive prefix
  __shared__ unsigned int seg[kPeerMaxRanks], cnt[kPeerMaxRanks];      // first staged row / rows of every destination
  __shared__ unsigned int gbase[kPeerMaxRanks];
  __shared__ PeerDst dst[kPeerMaxRanks];
  if (*flag) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = lanemask_lt();
  if (threadIdx.x < n_ranks) dst[threadIdx.x] = dst_table[threadIdx.x];
  const uint32_t id0 = (uint32_t)*row_id_base;
  const int64_t ntile = (nblk + kPbSub - 1) / kPbSub;
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    for (int i = threadIdx.x; i < kPbItems * kPbWarps * kPeerMaxRanks; i += kPbThreads) (&wcnt[0][0])[i] = 0;
    if (threadIdx.x < n_ranks) gbase[threadIdx.x] = bc[(int64_t)threadIdx.x * nblk + tile * kPbSub];
    __syncthreads();
    const int64_t base = tile * kPbTile;
    int32_t rc[kPbItems], rs[kPbItems], re[kPbItems];
    int rd[kPbItems];
    unsigned int slot[kPbItems];
#pragma unroll
    for (int j = 0; j < kPbItems; ++j) {
      const int64_t i = base + (int64_t)j * kPbThreads + threadIdx.x;
      rd[j] = -1;
      if (i < n) {
        rc[j] = c[i]; rs[j] = s[i]; re[j] = e[i];
        rd[j] = peer_dest(rc[j], owner, n_contigs, n_ranks);
      }
    }
    // rank of every row among the rows of its (item, warp) with the same destination; (item, warp, lane) order = row order
#pragma unroll
    for (int j = 0; j < kPbItems; ++j) {
      const unsigned peers = __match_any_sync(0xffffffffu, rd[j]);
      if (rd[j] >= 0 && lane == __ffs(peers) - 1) wcnt[j * kPbWarps + warp][rd[j]] = (unsigned short)__popc(peers);
      slot[j] = (unsigned)__popc(peers & lt);
    }
    __syncthreads();
    if (threadIdx.x < n_ranks) {  // exclusive prefix over the (item, warp) groups of this destination
      unsigned int run = 0;
      for (int k = 0; k < kPbItems * kPbWarps; ++k) { const unsigned int v = wcnt[k][threadIdx.x]; wcnt[k][threadIdx.x] = (unsigned short)run; run += v; }
      cnt[threadIdx.x] = run;
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // segments: destination r starts at the next multiple of 4 rows + the residue of its first remote row
      unsigned int end = 0;
      for (int r = 0; r < n_ranks; ++r) {
        const unsigned int a = (unsigned int)((reinterpret_cast<uintptr_t>(dst[r].contig + gbase[r]) >> 2) & 3u);
        seg[r] = ((end + 3u) & ~3u) + a;
        end = seg[r] + cnt[r];
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPbItems; ++j) {
      if (rd[j] < 0) continue;
      const unsigned int p = seg[rd[j]] + wcnt[j * kPbWarps + warp][rd[j]] + slot[j];
      const int64_t i = base + (int64_t)j * kPbThreads + threadIdx.x;
      col[0][p] = rc[j]; col[1][p] = rs[j]; col[2][p] = re[j]; col[3][p] = (int32_t)(id0 + (uint32_t)i);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staged rows are read by the bulk-copy engine
    __syncthreads();
    if (threadIdx.x < 4 * n_ranks) {
      const int r = threadIdx.x >> 2, k = threadIdx.x & 3;
      const unsigned int rows = cnt[r];
      if (rows) {
        int32_t *g = (k == 0 ? dst[r].contig : k == 1 ? dst[r].start : k == 2 ? dst[r].end : reinterpret_cast<int32_t *>(dst[r].row)) + gbase[r];
        const int32_t *src = col[k] + seg[r];
        const unsigned int head = min(rows, (4u - (seg[r] & 3u)) & 3u);  // rows before the first 16-byte boundary
        const unsigned int body = (rows - head) & ~3u;
        for (unsigned int q = 0; q < head; ++q) g[q] = src[q];
        if (body) {
          const uint64_t gdst = (uint64_t)reinterpret_cast<uintptr_t>(g + head);
          const uint32_t ssrc = (uint32_t)__cvta_generic_to_shared(src + head);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(body * 4u) : "memory");
        }
        for (unsigned int q = head + body; q < rows; ++q) g[q] = src[q];
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the staging arrays are rewritten by the next tile
    }
    __syncthreads();
  }
  // the block's shared memory goes away at exit and the signal kernel that follows promises delivery: wait for the writes
  if (threadIdx.x < 4 * n_ranks) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace pbgpu

extern "C" {

// receive arena: plain cudaMalloc memory (legacy CUDA IPC cannot export stream-ordered pool memory)
int pbgpu_peer_alloc(size_t bytes, void **d_ptr, unsigned char *handle_out /*[64]*/) {
  if (!d_ptr || !handle_out) return set_error(PBGPU_EINVAL, "NULL argument");
  *d_ptr = nullptr;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  // whole 2 MiB pages: a small cudaMalloc block shares its page (and its IPC mapping) with other allocations
  bytes = (bytes + ((size_t)2 << 20) - 1) / ((size_t)2 << 20) * ((size_t)2 << 20);
  if (!bytes) bytes = (size_t)2 << 20;
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
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

static unsigned long long peer_timeout_ns() {
  static unsigned long long v = [] {
    const char *e = getenv("PBGPU_PEER_TIMEOUT_MS");
    const long long ms = e ? atoll(e) : 60000;
    return (unsigned long long)(ms > 0 ? ms : 60000) * 1000000ull;
  }();
  return v;
}
static int peer_ctl_args(const uint64_t *ctl_base, int32_t world, PeerCtlArgs *a) {
  if (world < 1 || world > kPeerMaxRanks) return set_error(PBGPU_EINVAL, "bad world (at most %d ranks)", kPeerMaxRanks);
  if (!ctl_base) return set_error(PBGPU_EINVAL, "NULL argument");
  memset(a, 0, sizeof(*a));
  for (int d = 0; d < world; ++d) {
    if (!ctl_base[d]) return set_error(PBGPU_EINVAL, "NULL control block");
    a->ctl[d] = ctl_base[d];
  }
  return PBGPU_OK;
}

size_t pbgpu_peer_ctl_bytes(int32_t world, int32_t n_tables, int32_t n_contigs) {
  if (world < 1 || n_tables < 1 || n_contigs < 0) return 0;
  return kPeerCtlHist + sizeof(long long) * (size_t)world * (size_t)n_tables * ((size_t)n_contigs + 1);
}

static int peer_plan_impl(const int64_t *d_gathered, const void *d_own_ctl, uint64_t step, int32_t world, int32_t rank, int32_t n_tables,
                          int32_t n_contigs, const uint64_t *arena_base, const int64_t *cap_rows, int32_t *d_owner, void *d_dst,
                          int64_t *d_result, int64_t *h_result, void *stream, void *scratch_w, void *scratch_order) {
  if (world < 1 || world > kPeerMaxRanks) return set_error(PBGPU_EINVAL, "bad world (at most %d ranks)", kPeerMaxRanks);
  if (n_tables < 1 || n_tables > kPeerMaxTables) return set_error(PBGPU_EINVAL, "bad n_tables (at most %d)", kPeerMaxTables);
  if (rank < 0 || rank >= world || n_contigs < 0) return set_error(PBGPU_EINVAL, "bad rank / n_contigs");
  if ((!d_gathered && !d_own_ctl) || !arena_base || !cap_rows || !d_owner || !d_dst || !d_result) return set_error(PBGPU_EINVAL, "NULL argument");
  const unsigned long long *wait_flags = nullptr;
  if (d_own_ctl) {  // histograms were published into the control block: wait for flag A, read them from there
    wait_flags = (const unsigned long long *)d_own_ctl;
    d_gathered = (const int64_t *)((const char *)d_own_ctl + kPeerCtlHist);
  }
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
  unsigned long long *w = (unsigned long long *)scratch_w;
  int32_t *order = (int32_t *)scratch_order;
  if (!w || !order) {
    PB_TRY(sc.get(&w, (size_t)n_contigs + 1));
    PB_TRY(sc.get(&order, (size_t)n_contigs + 1));
  }
  PB_LAUNCH(peer_plan_kernel, 1, 1024, 0, s, (const long long *)d_gathered, world, rank, n_tables, n_contigs, a, w, order, d_owner,
            (PeerDst *)d_dst, (long long *)d_result, wait_flags, (unsigned long long)step, peer_timeout_ns(), (long long *)h_result);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}
int pbgpu_peer_plan(const int64_t *d_gathered, const void *d_own_ctl, uint64_t step, int32_t world, int32_t rank, int32_t n_tables,
                    int32_t n_contigs, const uint64_t *arena_base, const int64_t *cap_rows, int32_t *d_owner, void *d_dst, int64_t *d_result,
                    int64_t *h_result, void *stream) {
  return peer_plan_impl(d_gathered, d_own_ctl, step, world, rank, n_tables, n_contigs, arena_base, cap_rows, d_owner, d_dst, d_result, h_result,
                        stream, nullptr, nullptr);
}

// Every valid row of this rank's slice goes to the column arrays of its owner.  d_dst: n_ranks records of four device
// pointers (contig, start, end, row) = the start of THIS source's region at each destination (own arena or mapped peer
// memory), d_row_id_base: the global id of row 0, d_flag: non-zero = do nothing (arena overflow) -- all three are
// outputs of pbgpu_peer_plan and are read on the device, so the host enqueues plan and scatter back to back.
// $PBGPU_PEER_GRID = blocks per SM of the scatter kernel (persistent grid striding over the tiles), default 2: the
// kernel is bound by the NVLink stores, and a small resident grid lets the index build of the other table run beside it
// (r02d, N=2: 0.78 ms/step against 0.81 with one block per tile); 0 = one block per tile
// PBGPU_PEER_STORE=ldst: the first scatter kernel (4-byte stores through registers); default: TMA bulk stores
static bool peer_bulk_store() {
  static bool v = [] { const char *e = getenv("PBGPU_PEER_STORE"); return !(e && !strcmp(e, "ldst")); }();
  return v;
}
static int peer_grid_per_sm() {
  static int v = [] { const char *e = getenv("PBGPU_PEER_GRID"); const int x = e ? atoi(e) : 2; return x > 0 && x <= 8 ? x : 0; }();
  return v;
}

static int peer_scatter_impl(const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t n, const int32_t *d_owner,
                             int32_t n_contigs, int32_t n_ranks, const int64_t *d_row_id_base, const void *d_dst, const int64_t *d_flag,
                             void *stream, unsigned int *bc_scratch) {
  if (n < 0 || n >= (1ll << 32) || n_ranks < 1 || n_ranks > kPeerMaxRanks) return set_error(PBGPU_EINVAL, "bad n / n_ranks (at most %d ranks)", kPeerMaxRanks);
  if (!d_owner || !d_dst || !d_row_id_base || !d_flag) return set_error(PBGPU_EINVAL, "NULL argument");
  if (n == 0) return PBGPU_OK;
  if (!d_contig || !d_start || !d_end) return set_error(PBGPU_EINVAL, "NULL column");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t nblk = cdiv(n, kPeerTile);
  Scratch sc(s);
  unsigned int *bc = bc_scratch;
  if (!bc) PB_TRY(sc.get(&bc, (size_t)nblk * n_ranks));
  int64_t grid = nblk;
  if (peer_grid_per_sm() && grid > (int64_t)kSMs * peer_grid_per_sm()) grid = (int64_t)kSMs * peer_grid_per_sm();
  PB_LAUNCH(peer_block_count_kernel, (unsigned)nblk, kPeerThreads, 0, s, d_contig, n, d_owner, n_contigs, n_ranks, (const long long *)d_flag, bc, nblk);
  PB_LAUNCH(peer_block_scan_kernel, (unsigned)n_ranks, 1024, 0, s, bc, nblk, (const long long *)d_flag);
  if (peer_bulk_store()) {
    constexpr size_t smem = sizeof(int32_t) * 4 * (size_t)kPbStageRows;
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(peer_scatter_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
    int64_t bgrid = cdiv(nblk, kPbSub);
    if (peer_grid_per_sm() && bgrid > (int64_t)kSMs * peer_grid_per_sm()) bgrid = (int64_t)kSMs * peer_grid_per_sm();
    PB_LAUNCH(peer_scatter_bulk_kernel, (unsigned)bgrid, kPbThreads, smem, s, d_contig, d_start, d_end, n, d_owner, n_contigs, n_ranks,
              (const long long *)d_row_id_base, (const PeerDst *)d_dst, bc, nblk, (const long long *)d_flag);
  } else {
    PB_LAUNCH(peer_scatter_kernel, (unsigned)grid, kPeerThreads, 0, s, d_contig, d_start, d_end, n, d_owner, n_contigs, n_ranks,
              (const long long *)d_row_id_base, (const PeerDst *)d_dst, bc, nblk, (const long long *)d_flag);
  }
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}
int pbgpu_peer_scatter(const int32_t *d_contig, const int32_t *d_start, const int32_t *d_end, int64_t n, const int32_t *d_owner,
                       int32_t n_contigs, int32_t n_ranks, const int64_t *d_row_id_base, const void *d_dst, const int64_t *d_flag,
                       void *stream) {
  return peer_scatter_impl(d_contig, d_start, d_end, n, d_owner, n_contigs, n_ranks, d_row_id_base, d_dst, d_flag, stream, nullptr);
}

// layout of the caller-provided scratch of a step: w u64[nc+1] | order i32[nc+1] (padded) | per table bc u32[world * tiles]
struct PeerScratchLayout { size_t w, order, bc[kPeerMaxTables], total; };
static PeerScratchLayout peer_scratch_layout(const pbgpu_peer_step *d) {
  PeerScratchLayout l;
  size_t off = 0;
  auto take = [&off](size_t bytes) { const size_t at = off; off += (bytes + 255) / 256 * 256; return at; };
  l.w = take(sizeof(unsigned long long) * ((size_t)d->n_contigs + 1));
  l.order = take(sizeof(int32_t) * ((size_t)d->n_contigs + 1));
  for (int t = 0; t < kPeerMaxTables; ++t)
    l.bc[t] = t < d->n_tables ? take(sizeof(unsigned int) * (size_t)d->world * (size_t)cdiv(d->rows[t] > 0 ? d->rows[t] : 0, kPeerTile)) : 0;
  l.total = off;
  return l;
}
size_t pbgpu_peer_scratch_bytes(const pbgpu_peer_step *d) {
  if (!d || d->world < 1 || d->n_tables < 1 || d->n_tables > kPeerMaxTables || d->n_contigs < 0) return 0;
  return peer_scratch_layout(d).total;
}


static int peer_step_check(const pbgpu_peer_step *d) {
  if (!d) return set_error(PBGPU_EINVAL, "NULL descriptor");
  if (d->world < 1 || d->world > kPeerMaxRanks || d->rank < 0 || d->rank >= d->world) return set_error(PBGPU_EINVAL, "bad world / rank");
  if (d->n_tables < 1 || d->n_tables > kPeerMaxTables || d->n_contigs < 0) return set_error(PBGPU_EINVAL, "bad n_tables / n_contigs");
  if (!d->arena_base || !d->cap_rows || !d->d_hist || !d->d_owner || !d->d_dst || !d->d_result || !d->d_status)
    return set_error(PBGPU_EINVAL, "NULL argument");
  for (int t = 0; t < d->n_tables; ++t) {
    if (d->rows[t] < 0 || d->rows[t] >= (1ll << 32)) return set_error(PBGPU_EINVAL, "bad row count");
    if (d->rows[t] > 0 && (!d->contig[t] || !d->start[t] || !d->end[t])) return set_error(PBGPU_EINVAL, "NULL column");
  }
  return PBGPU_OK;
}

// First half of an exchange step on `stream`: histograms of all tables (one launch) and, with control blocks, their
// publication and the plan (which posts its result to h_result).  Without control blocks the caller all-gathers d_hist
// and calls pbgpu_peer_plan itself.
int pbgpu_peer_begin(const pbgpu_peer_step *d, void *stream) {
  PB_TRY(peer_step_check(d));
  cudaStream_t s = (cudaStream_t)stream;
  const int ph = d->phases ? d->phases : (PBGPU_PEER_HIST | PBGPU_PEER_PLAN);
  if (ph & PBGPU_PEER_HIST) {
    const int len = d->n_tables * (d->n_contigs + 1);
    PB_CUDA(cudaMemsetAsync(d->d_hist, 0, sizeof(int64_t) * ((size_t)len + 1), s));
    PeerTablesArgs tb;
    memset(&tb, 0, sizeof(tb));
    int64_t gx = 1;
    for (int t = 0; t < d->n_tables; ++t) {
      tb.c[t] = d->contig[t];
      tb.n[t] = d->rows[t];
      int64_t g = cdiv(d->rows[t], 256 * 16);
      if (g > gx) gx = g;
    }
    if (gx > kSMs * 8) gx = kSMs * 8;
    PeerCtlArgs ca;
    memset(&ca, 0, sizeof(ca));
    if (d->ctl_base) PB_TRY(peer_ctl_args(d->ctl_base, d->world, &ca));
    const size_t smem = d->n_contigs <= 4096 ? sizeof(unsigned int) * (size_t)d->n_contigs : 0;
    PB_LAUNCH(peer_hist_publish_kernel, dim3((unsigned)gx, (unsigned)d->n_tables), 256, smem, s, tb, d->n_tables, d->n_contigs,
              (unsigned long long *)d->d_hist, ca, d->world, d->rank, (unsigned long long)d->step, d->ctl_base ? 1 : 0);
    PB_CHECK_LAUNCH();
  }
  if (!d->ctl_base || !(ph & PBGPU_PEER_PLAN)) return PBGPU_OK;
  const PeerScratchLayout l = peer_scratch_layout(d);
  const bool own = d->d_scratch && d->scratch_bytes >= l.total;
  return peer_plan_impl(nullptr, (const void *)d->ctl_base[d->rank], d->step, d->world, d->rank, d->n_tables, d->n_contigs, d->arena_base,
                        d->cap_rows, d->d_owner, d->d_dst, d->d_result, d->h_result, stream, own ? (char *)d->d_scratch + l.w : nullptr,
                        own ? (char *)d->d_scratch + l.order : nullptr);
}

// Second half, once per table, on any stream ordered after pbgpu_peer_begin: the scatter of the table and, with control
// blocks, its signal + wait (after which the table is complete on this rank).
int pbgpu_peer_table(const pbgpu_peer_step *d, int32_t table, void *stream) {
  PB_TRY(peer_step_check(d));
  if (table < 0 || table >= d->n_tables) return set_error(PBGPU_EINVAL, "bad table");
  const int T = d->n_tables;
  const int ph = d->phases ? d->phases : (PBGPU_PEER_SCATTER | PBGPU_PEER_SIGNAL | PBGPU_PEER_WAIT);
  if (ph & PBGPU_PEER_SCATTER) {
    const PeerScratchLayout l = peer_scratch_layout(d);
    const bool own = d->d_scratch && d->scratch_bytes >= l.total;
    PB_TRY(peer_scatter_impl(d->contig[table], d->start[table], d->end[table], d->rows[table], d->d_owner, d->n_contigs, d->world,
                             d->d_result + T + table, (const char *)d->d_dst + sizeof(PeerDst) * (size_t)d->world * table,
                             d->d_result + 3 * T, stream, own ? (unsigned int *)((char *)d->d_scratch + l.bc[table]) : nullptr));
  }
  const int mode = ((ph & PBGPU_PEER_SIGNAL) ? 1 : 0) | ((ph & PBGPU_PEER_WAIT) ? 2 : 0);
  if (!d->ctl_base || !mode) return PBGPU_OK;
  PeerCtlArgs ca;
  PB_TRY(peer_ctl_args(d->ctl_base, d->world, &ca));
  PB_LAUNCH(peer_signal_wait_kernel, 1, 32, 0, (cudaStream_t)stream, ca, d->world, d->rank, table, (unsigned long long)d->step,
            peer_timeout_ns(), (long long *)d->d_status, mode);
  PB_CHECK_LAUNCH();
  return PBGPU_OK;
}

}  // extern "C"
