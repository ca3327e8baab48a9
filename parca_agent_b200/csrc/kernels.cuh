// kernels.cuh — sm_100a device code for the sample → Arrow aggregation path.
//
// Every kernel here is HBM-/L2-bound integer work (no tensor cores by design). Conventions:
//   * generic passes use a fixed grid (kGrid blocks x kThreads) and take their element count
//     from device memory, so the host never synchronises between dependent passes;
//   * "first-occurrence rank" (the invariant behind every dictionary index of the reference:
//     reporter/parca_reporter.go:425, reporter/arrow_v2.go:191,:302) is computed as
//     atomicMin(first position) -> flag -> exclusive scan, never by insertion order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pa {

constexpr uint32_t kNull = 0xFFFFFFFFu;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxCols = 48;  // label columns + 8 constant-ish REE columns

enum : uint32_t {
  ERR_TABLE_FULL = 1u, ERR_BAD_FRAME_ID = 2u, ERR_BAD_STRING_ID = 4u, ERR_BAD_LABELSET = 8u,
  ERR_BAD_KIND = 16u, ERR_INDEX_OVERFLOW = 32u, ERR_BAD_CPU = 64u, ERR_BAD_FRAME_RANGE = 128u,
};

struct __align__(16) Key128 { unsigned long long hi, lo; };

// open-address stack table entry: 32 B = one L2 sector
struct __align__(32) StackSlot {
  Key128 key;          // (0,0) = empty; claimed once by a 128-bit CAS, never changes afterwards
  uint32_t first_inv;  // 0xFFFFFFFF - first row, maintained with atomicMax (memset-0 = none yet)
  uint32_t count;      // occurrences (side table, not part of the reference's record)
  uint32_t offset;     // start of this stack's run in the location-index stream
  uint32_t size;       // nframes of the first occurrence (listEntryRef.listSize)
};

// device-resident counters, copied to the host once per flush
struct Counters {
  uint32_t err;
  uint32_t n_unique;
  unsigned long long n_indices64;
  uint32_t n_locations, n_lines, n_functions;
  uint32_t n_dict_type, n_dict_map, n_dict_bid, n_dict_file;
  uint32_t null_bid, null_file;
  uint32_t pad0;
  uint32_t n_runs[kMaxCols];
  uint32_t n_dict[kMaxCols];
  uint32_t n_null[kMaxCols];
  uint32_t last_nonnull_plus1[kMaxCols];
};

// ---------------------------------------------------------------------------------------------
// small helpers
__device__ __forceinline__ unsigned long long rotl64(unsigned long long x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ unsigned long long bswap64(unsigned long long x) {
  uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  return ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}
__device__ __forceinline__ unsigned long long ldg_stream64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ Key128 ld_key(const Key128* p) {
  Key128 k;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(k.hi), "=l"(k.lo) : "l"(p));
  return k;
}
__device__ __forceinline__ Key128 cas128(Key128* addr, Key128 cmp, Key128 val) {
  Key128 old;
  asm volatile(
      "{\n .reg .b128 c, v, o;\n mov.b128 c, {%2, %3};\n mov.b128 v, {%4, %5};\n"
      " atom.global.cas.b128 o, [%6], c, v;\n mov.b128 {%0, %1}, o;\n}"
      : "=l"(old.hi), "=l"(old.lo)
      : "l"(cmp.hi), "l"(cmp.lo), "l"(val.hi), "l"(val.lo), "l"(addr)
      : "memory");
  return old;
}
__device__ __forceinline__ bool key_eq(Key128 a, Key128 b) { return a.hi == b.hi && a.lo == b.lo; }
__device__ __forceinline__ bool key_zero(Key128 a) { return (a.hi | a.lo) == 0; }
__device__ __forceinline__ uint32_t mix_slot(Key128 k) {
  unsigned long long x = k.lo ^ (k.hi * 0x9E3779B97F4A7C15ull);
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
  return (uint32_t)x;
}
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}

// contiguous per-block range of [0,n), multiple of kThreads so tiles stay aligned
__device__ __forceinline__ void block_range(uint32_t n, uint32_t* begin, uint32_t* end) {
  uint32_t per = (n + gridDim.x - 1) / gridDim.x;
  per = (per + kThreads - 1) / kThreads * kThreads;
  unsigned long long b = (unsigned long long)blockIdx.x * per;
  *begin = b < n ? (uint32_t)b : n;
  unsigned long long e = b + per;
  *end = e < n ? (uint32_t)e : n;
}

// ---------------------------------------------------------------------------------------------
// stack table: find-or-claim on the 128-bit stack id (StacktraceDictBuilderV2.index, arrow_v2.go:230)
__device__ __forceinline__ uint32_t stack_find_or_insert(StackSlot* tab, uint32_t mask, Key128 k, Counters* ctr) {
  if (key_zero(k)) return mask + 1;  // the all-zero id lives in a dedicated slot past the table
  uint32_t idx = mix_slot(k) & mask;
  const Key128 zero{0ull, 0ull};
  for (uint32_t probe = 0; probe <= mask; probe++) {
    Key128 cur = ld_key(&tab[idx].key);
    if (cur.hi == 0 || cur.lo == 0) cur = cas128(&tab[idx].key, zero, k);  // empty / possibly torn: CAS is authoritative
    if (key_zero(cur) || key_eq(cur, k)) return idx;
    idx = (idx + 1) & mask;
  }
  atomicOr(&ctr->err, ERR_TABLE_FULL);
  return kNull;
}

// Warp-aggregated insert: lanes carrying the same id elect the lowest lane (== lowest row, rows
// ascend with the lane), which does one table walk, one atomicMax (first row) and one atomicAdd
// (count) for the whole group. All 32 lanes must call this converged.
__device__ __forceinline__ uint32_t warp_insert(StackSlot* tab, uint32_t mask, Key128 k, uint32_t row, bool valid, Counters* ctr) {
  const unsigned full = 0xFFFFFFFFu;
  int lane = threadIdx.x & 31;
  unsigned long long tag = valid ? (k.lo ^ rotl64(k.hi, 29)) : (0xDEAD00000000ull + lane);
  unsigned grp = __match_any_sync(full, tag);
  int leader = __ffs(grp) - 1;
  unsigned long long lhi = __shfl_sync(full, k.hi, leader), llo = __shfl_sync(full, k.lo, leader);
  int lvalid = __shfl_sync(full, (int)valid, leader);
  bool agree = valid && lvalid && lhi == k.hi && llo == k.lo;
  unsigned agree_mask = __ballot_sync(full, agree) & grp;
  bool own = valid && (lane == leader || !agree);  // tag collisions between different ids fall back to a private insert
  uint32_t idx = kNull;
  if (own) {
    idx = stack_find_or_insert(tab, mask, k, ctr);
    if (idx != kNull) {
      uint32_t inv = 0xFFFFFFFFu - row;
      if (*(volatile uint32_t*)&tab[idx].first_inv < inv) atomicMax(&tab[idx].first_inv, inv);
      atomicAdd(&tab[idx].count, (lane == leader) ? (uint32_t)__popc(agree_mask) : 1u);
    }
  }
  uint32_t lidx = __shfl_sync(full, idx, leader);
  return own ? idx : (agree ? lidx : kNull);
}

// ---------------------------------------------------------------------------------------------
// header pass: 64-byte AoS sample headers -> Arrow row columns + compact SoA side arrays.
// Replaces the per-sample appends of writeSampleV2 (reporter/parca_reporter.go:394-405).
struct HeaderArgs {
  const uint4* hdr;            // 4 x uint4 per row
  uint32_t row0, row1;         // this chunk
  long long* timestamp;        // Arrow column
  long long* value;            // Arrow column
  uint8_t* uuid;               // Arrow column (16 B/row), written here in provided-hash mode
  uint8_t* kind;
  uint16_t* nframes;
  unsigned long long* frame_off;
  uint32_t* ls;
  uint32_t* cpu;
  uint32_t* tid;
  uint32_t* comm;              // canonical string id of meta.Comm
  const uint32_t* sid2cid;
  uint32_t n_sids, n_labelsets;
  unsigned long long n_frame_ids;  // frames staged for this batch
  int provided;                // hash arrives in the header: insert here
  StackSlot* tab;
  uint32_t mask;
  uint32_t* slot_of_row;
  Counters* ctr;
};

__global__ void __launch_bounds__(kThreads) k_header(HeaderArgs a) {
  uint32_t stride = gridDim.x * kThreads;
  uint32_t span = a.row1 - a.row0;
  uint32_t iters = (span + stride - 1) / stride;  // uniform trip count: warp_insert needs converged warps
  for (uint32_t it = 0; it < iters; it++) {
    uint32_t r = a.row0 + it * stride + blockIdx.x * kThreads + threadIdx.x;
    bool valid = r < a.row1;
    Key128 k{0ull, 0ull};
    if (valid) {
      const uint4* h = a.hdr + 4ull * r;
      uint4 q0 = __ldg(h), q1 = __ldg(h + 1), q2 = __ldg(h + 2), q3 = __ldg(h + 3);
      k.hi = ((unsigned long long)q0.y << 32) | q0.x;
      k.lo = ((unsigned long long)q0.w << 32) | q0.z;
      long long ts = (long long)(((unsigned long long)q1.y << 32) | q1.x);
      long long val = (long long)(((unsigned long long)q1.w << 32) | q1.z);
      uint32_t tid = q2.y, comm_sid = q2.z, ls = q2.w;
      unsigned long long foff = ((unsigned long long)q3.y << 32) | q3.x;
      uint32_t cpu = q3.z;
      uint32_t nfr = q3.w & 0xFFFFu, knd = (q3.w >> 16) & 0xFFu;
      uint32_t err = 0;
      if (knd >= 7) { err |= ERR_BAD_KIND; knd = 0; }
      if (ls >= a.n_labelsets) { err |= ERR_BAD_LABELSET; ls = 0; }
      if (comm_sid >= a.n_sids) { err |= ERR_BAD_STRING_ID; comm_sid = 0; }
      if (cpu >= 65536u) { err |= ERR_BAD_CPU; cpu = 0; }
      if (foff + nfr > a.n_frame_ids) { err |= ERR_BAD_FRAME_RANGE; nfr = 0; foff = 0; }
      if (err) atomicOr(&a.ctr->err, err);
      a.timestamp[r] = ts;
      a.value[r] = (knd == 0) ? 1ll : val;  // TraceOriginSampling writes value 1 (:340)
      a.kind[r] = (uint8_t)knd;
      a.nframes[r] = (uint16_t)nfr;
      a.frame_off[r] = foff;
      a.ls[r] = ls;
      a.cpu[r] = cpu;
      a.tid[r] = tid;
      a.comm[r] = a.sid2cid[comm_sid];
      if (a.provided) {  // trace.Hash.Bytes(): big-endian hi||lo
        ulonglong2 id = make_ulonglong2(bswap64(k.hi), bswap64(k.lo));
        *reinterpret_cast<ulonglong2*>(a.uuid + 16ull * r) = id;
      }
    }
    if (a.provided) {
      uint32_t s = warp_insert(a.tab, a.mask, k, r, valid, a.ctr);
      if (valid) a.slot_of_row[r] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// XXH64 x 2 seeds over each sample's frame ids + fused table insert (PA_HASH_XXH64X2).
constexpr unsigned long long XP1 = 11400714785074694791ULL, XP2 = 14029467366897019727ULL, XP3 = 1609587929392839161ULL,
                             XP4 = 9650029242287828579ULL, XP5 = 2870177450012600261ULL;
constexpr unsigned long long kSeedLo = 0x9E3779B97F4A7C15ULL;

__device__ __forceinline__ unsigned long long xxh_round_pre(unsigned long long acc, unsigned long long in_p2) {
  return rotl64(acc + in_p2, 31) * XP1;  // in_p2 = input * PRIME64_2, shared by both seeds
}
__device__ __forceinline__ unsigned long long xxh_merge(unsigned long long h, unsigned long long v) {
  v = rotl64(v * XP2, 31) * XP1;
  return (h ^ v) * XP1 + XP4;
}
__device__ __forceinline__ unsigned long long xxh_avalanche(unsigned long long h) {
  h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
  return h;
}
__device__ __forceinline__ unsigned long long xxh_lane_init(unsigned long long seed, int j) {
  return j == 0 ? seed + XP1 + XP2 : (j == 1 ? seed + XP2 : (j == 2 ? seed : seed - XP1));
}
// finish one sample given the 4 lane accumulators of its group (all 4 lanes call this)
__device__ __forceinline__ unsigned long long xxh_finish(unsigned long long acc, unsigned long long seed, uint32_t n,
                                                         unsigned long long t0, unsigned long long t1, unsigned long long t2, int gbase) {
  const unsigned full = 0xFFFFFFFFu;
  unsigned long long v0 = __shfl_sync(full, acc, gbase), v1 = __shfl_sync(full, acc, gbase + 1),
                     v2 = __shfl_sync(full, acc, gbase + 2), v3 = __shfl_sync(full, acc, gbase + 3);
  unsigned long long h;
  if (n >= 4) {
    h = rotl64(v0, 1) + rotl64(v1, 7) + rotl64(v2, 12) + rotl64(v3, 18);
    h = xxh_merge(h, v0); h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3);
  } else {
    h = seed + XP5;
  }
  h += (unsigned long long)n * 8ull;
  uint32_t t = n & 3u;  // up to three trailing 8-byte words after the last full stripe
  if (t > 0) { h ^= rotl64(t0 * XP2, 31) * XP1; h = rotl64(h, 27) * XP1 + XP4; }
  if (t > 1) { h ^= rotl64(t1 * XP2, 31) * XP1; h = rotl64(h, 27) * XP1 + XP4; }
  if (t > 2) { h ^= rotl64(t2 * XP2, 31) * XP1; h = rotl64(h, 27) * XP1 + XP4; }
  return xxh_avalanche(h);
}

struct HashArgs {
  const unsigned long long* frames;
  const unsigned long long* frame_off;
  const uint16_t* nframes;
  uint32_t row0, row1;
  uint8_t* uuid;
  uint32_t* slot_of_row;
  StackSlot* tab;
  uint32_t mask;
  Counters* ctr;
};

// Variant A: 4 lanes per sample (one XXH64 accumulator lane each), 8 samples per warp, frame ids
// streamed straight from global memory (each lane reads one 8-byte word of every 32-byte stripe).
__global__ void __launch_bounds__(kThreads) k_hash_insert_direct(HashArgs a) {
  const unsigned full = 0xFFFFFFFFu;
  int lane = threadIdx.x & 31, j = lane & 3, g = lane >> 2, gbase = lane & ~3;
  uint32_t warp = (blockIdx.x * kThreads + threadIdx.x) >> 5, nwarps = (gridDim.x * kThreads) >> 5;
  uint32_t span = a.row1 - a.row0;
  uint32_t iters = (span + nwarps * 8 - 1) / (nwarps * 8);
  for (uint32_t it = 0; it < iters; it++) {
    uint32_t r = a.row0 + (it * nwarps + warp) * 8 + g;
    bool valid = r < a.row1;
    uint32_t n = valid ? a.nframes[r] : 0;
    const unsigned long long* p = a.frames + (valid ? a.frame_off[r] : 0ull);
    unsigned long long a0 = xxh_lane_init(0ull, j), a1 = xxh_lane_init(kSeedLo, j);
    uint32_t ns = n >> 2;
    const unsigned long long* q = p + j;
    uint32_t s = 0;
    for (; s + 4 <= ns; s += 4) {
      unsigned long long w0 = ldg_stream64(q + 4 * s), w1 = ldg_stream64(q + 4 * s + 4), w2 = ldg_stream64(q + 4 * s + 8),
                         w3 = ldg_stream64(q + 4 * s + 12);
      unsigned long long m0 = w0 * XP2, m1 = w1 * XP2, m2 = w2 * XP2, m3 = w3 * XP2;
      a0 = xxh_round_pre(a0, m0); a1 = xxh_round_pre(a1, m0);
      a0 = xxh_round_pre(a0, m1); a1 = xxh_round_pre(a1, m1);
      a0 = xxh_round_pre(a0, m2); a1 = xxh_round_pre(a1, m2);
      a0 = xxh_round_pre(a0, m3); a1 = xxh_round_pre(a1, m3);
    }
    for (; s < ns; s++) {
      unsigned long long m = ldg_stream64(q + 4 * s) * XP2;
      a0 = xxh_round_pre(a0, m); a1 = xxh_round_pre(a1, m);
    }
    const uint32_t nt = n & 3u;
    unsigned long long t0 = nt > 0 ? ldg_stream64(p + 4 * ns) : 0ull, t1 = nt > 1 ? ldg_stream64(p + 4 * ns + 1) : 0ull,
                       t2 = nt > 2 ? ldg_stream64(p + 4 * ns + 2) : 0ull;
    __syncwarp(full);
    Key128 k;
    k.hi = xxh_finish(a0, 0ull, n, t0, t1, t2, gbase);
    k.lo = xxh_finish(a1, kSeedLo, n, t0, t1, t2, gbase);
    bool mine = valid && j == 0;
    if (mine) *reinterpret_cast<ulonglong2*>(a.uuid + 16ull * r) = make_ulonglong2(bswap64(k.hi), bswap64(k.lo));
    uint32_t slot = warp_insert(a.tab, a.mask, k, r, mine, a.ctr);
    if (mine) a.slot_of_row[r] = slot;
  }
}

// Variant B (default): one warp-iteration covers 32 consecutive samples. The XXH64 rounds still run
// with 4 lanes per sample (8 samples at a time, 4 sub-iterations), but the per-sample epilogue —
// accumulator merge, avalanche, 16-byte id store, table insert — runs once with one *thread per
// sample* after a shuffle transpose, so its cost is amortised over 32 samples instead of 8 and
// the id store / slot store are fully coalesced (512 B / 128 B per warp).
__device__ __forceinline__ unsigned long long xxh_finish_own(const unsigned long long (&v)[4], unsigned long long seed, uint32_t n,
                                                             unsigned long long t0, unsigned long long t1, unsigned long long t2) {
  unsigned long long h;
  if (n >= 4) {
    h = rotl64(v[0], 1) + rotl64(v[1], 7) + rotl64(v[2], 12) + rotl64(v[3], 18);
    h = xxh_merge(h, v[0]); h = xxh_merge(h, v[1]); h = xxh_merge(h, v[2]); h = xxh_merge(h, v[3]);
  } else {
    h = seed + XP5;
  }
  h += (unsigned long long)n * 8ull;
  uint32_t t = n & 3u;
  if (t > 0) { h ^= rotl64(t0 * XP2, 31) * XP1; h = rotl64(h, 27) * XP1 + XP4; }
  if (t > 1) { h ^= rotl64(t1 * XP2, 31) * XP1; h = rotl64(h, 27) * XP1 + XP4; }
  if (t > 2) { h ^= rotl64(t2 * XP2, 31) * XP1; h = rotl64(h, 27) * XP1 + XP4; }
  return xxh_avalanche(h);
}

__global__ void __launch_bounds__(kThreads, 3) k_hash_insert(HashArgs a) {
  const unsigned full = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31, j = lane & 3, g = lane >> 2;
  const uint32_t warp = (blockIdx.x * kThreads + threadIdx.x) >> 5, nwarps = (gridDim.x * kThreads) >> 5;
  const uint32_t span = a.row1 - a.row0;
  const uint32_t iters = (span + nwarps * 32 - 1) / (nwarps * 32);
  for (uint32_t it = 0; it < iters; it++) {
    const uint32_t r = a.row0 + (it * nwarps + warp) * 32 + lane;  // the sample this lane finishes
    const bool valid = r < a.row1;
    const uint32_t n_me = valid ? a.nframes[r] : 0u;
    const unsigned long long off_me = valid ? a.frame_off[r] : 0ull;
    unsigned long long v0[4], v1[4];
#pragma unroll
    for (int sub = 0; sub < 4; sub++) {
      const int src = sub * 8 + g;  // lane that owns the sample my 4-lane group hashes now
      const uint32_t n = __shfl_sync(full, n_me, src);
      const unsigned long long off = __shfl_sync(full, off_me, src);
      const unsigned long long* q = a.frames + off + j;
      unsigned long long a0 = xxh_lane_init(0ull, j), a1 = xxh_lane_init(kSeedLo, j);
      const uint32_t ns = n >> 2;
      uint32_t s = 0;
      for (; s + 8 <= ns; s += 8) {  // 8 stripes = 8 independent 8-byte loads in flight per lane
        unsigned long long w[8];
#pragma unroll
        for (int u = 0; u < 8; u++) w[u] = ldg_stream64(q + 4 * (s + u));
#pragma unroll
        for (int u = 0; u < 8; u++) {
          unsigned long long m = w[u] * XP2;
          a0 = xxh_round_pre(a0, m);
          a1 = xxh_round_pre(a1, m);
        }
      }
      for (; s < ns; s++) {
        unsigned long long m = ldg_stream64(q + 4 * s) * XP2;
        a0 = xxh_round_pre(a0, m);
        a1 = xxh_round_pre(a1, m);
      }
      __syncwarp(full);
      // transpose: lane L (in octet `sub`) receives accumulator lane jj of sample L from lane 4*(L&7)+jj
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        unsigned long long x0 = __shfl_sync(full, a0, 4 * (lane & 7) + jj), x1 = __shfl_sync(full, a1, 4 * (lane & 7) + jj);
        if ((lane >> 3) == sub) { v0[jj] = x0; v1[jj] = x1; }
      }
    }
    const uint32_t nt = n_me & 3u;
    const unsigned long long* tp = a.frames + off_me + (n_me & ~3u);
    unsigned long long t0 = nt > 0 ? ldg_stream64(tp) : 0ull, t1 = nt > 1 ? ldg_stream64(tp + 1) : 0ull, t2 = nt > 2 ? ldg_stream64(tp + 2) : 0ull;
    Key128 k;
    k.hi = xxh_finish_own(v0, 0ull, n_me, t0, t1, t2);
    k.lo = xxh_finish_own(v1, kSeedLo, n_me, t0, t1, t2);
    if (valid) *reinterpret_cast<ulonglong2*>(a.uuid + 16ull * r) = make_ulonglong2(bswap64(k.hi), bswap64(k.lo));
    uint32_t slot = warp_insert(a.tab, a.mask, k, r, valid, a.ctr);
    if (valid) a.slot_of_row[r] = slot;
  }
}

// ---------------------------------------------------------------------------------------------
// block-wide exclusive scan (kThreads threads). T needs operator+ and a zero-initialised T().
struct Pair { uint32_t cnt; unsigned long long fr; };
__device__ __forceinline__ Pair operator+(Pair a, Pair b) { return Pair{a.cnt + b.cnt, a.fr + b.fr}; }
__device__ __forceinline__ uint32_t shfl_up_t(uint32_t v, int d) { return __shfl_up_sync(0xFFFFFFFFu, v, d); }
__device__ __forceinline__ Pair shfl_up_t(Pair v, int d) {
  return Pair{__shfl_up_sync(0xFFFFFFFFu, v.cnt, d), __shfl_up_sync(0xFFFFFFFFu, v.fr, d)};
}
__device__ __forceinline__ uint32_t shfl_idx_t(uint32_t v, int l) { return __shfl_sync(0xFFFFFFFFu, v, l); }
__device__ __forceinline__ Pair shfl_idx_t(Pair v, int l) { return Pair{__shfl_sync(0xFFFFFFFFu, v.cnt, l), __shfl_sync(0xFFFFFFFFu, v.fr, l)}; }
__device__ __forceinline__ uint32_t zero_of(uint32_t) { return 0u; }
__device__ __forceinline__ Pair zero_of(Pair) { return Pair{0u, 0ull}; }

template <class T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total) {
  __shared__ T s_w[kWarps];
  __shared__ T s_tot;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = shfl_up_t(inc, d);
    if (lane >= d) inc = o + inc;
  }
  if (lane == 31) s_w[w] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    T run = zero_of(v);
    for (int i = 0; i < kWarps; i++) { T t = s_w[i]; s_w[i] = run; run = run + t; }
    s_tot = run;
  }
  __syncthreads();
  T excl_in_warp = shfl_up_t(inc, 1);
  if (lane == 0) excl_in_warp = zero_of(v);
  T res = s_w[w] + excl_in_warp;
  *total = s_tot;
  __syncthreads();  // s_w / s_tot are reused by the next tile
  return res;
}

// Generic fixed-grid reduce-then-scan. F provides:
//   typedef T; uint32_t n(); T value(uint32_t i); void emit(uint32_t i, T exclusive, T v); void total(T t);
// blockIdx.y selects an independent job (F indexes its own job table with it).
template <class F>
__global__ void __launch_bounds__(kThreads) k_scan_reduce(F f, typename F::T* partial) {
  typedef typename F::T T;
  uint32_t n = f.n(), begin, end;
  block_range(n, &begin, &end);
  T acc = zero_of(T());
  for (uint32_t i = begin + threadIdx.x; i < end; i += kThreads) acc = acc + f.value(i);
  T tot;
  block_exclusive_scan(acc, &tot);
  if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = tot;
}
template <class F>
__global__ void k_scan_partials(F f, typename F::T* partial, int g) {  // grid = njobs blocks, 32 threads: warp scan over the block totals
  typedef typename F::T T;
  T* p = partial + (size_t)blockIdx.x * g;
  int lane = threadIdx.x;
  T run = zero_of(T());
  for (int base = 0; base < g; base += 32) {
    int i = base + lane;
    T v = i < g ? p[i] : zero_of(T());
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      T o = shfl_up_t(inc, d);
      if (lane >= d) inc = o + inc;
    }
    T ex = shfl_up_t(inc, 1);
    if (lane == 0) ex = zero_of(T());
    if (i < g) p[i] = run + ex;
    T tot = inc;  // lane 31 holds the chunk total
    tot = shfl_idx_t(tot, 31);
    run = run + tot;
  }
  if (lane == 0) f.total(blockIdx.x, run);
}
template <class F>
__global__ void __launch_bounds__(kThreads) k_scan_emit(F f, const typename F::T* partial) {
  typedef typename F::T T;
  uint32_t n = f.n(), begin, end;
  block_range(n, &begin, &end);
  T run = partial[blockIdx.y * gridDim.x + blockIdx.x];
  for (uint32_t tile = begin; tile < end; tile += kThreads) {
    uint32_t i = tile + threadIdx.x;
    bool in = i < end;
    T v = in ? f.value(i) : zero_of(T());
    T tot;
    T ex = block_exclusive_scan(v, &tot);
    if (in) f.emit(i, run + ex, v);
    run = run + tot;
  }
}

// ---------------------------------------------------------------------------------------------
// (1) unique stacks in first-occurrence order: ordinal + start offset of each first occurrence
struct RowFirstF {
  typedef Pair T;
  uint32_t n_rows;
  const uint32_t* slot_of_row;
  StackSlot* tab;
  const uint16_t* nframes;
  uint32_t* uniq_row;     // [ordinal] -> first row
  uint32_t* uniq_count;   // [ordinal] -> occurrences
  Counters* ctr;
  __device__ uint32_t n() const { return n_rows; }
  __device__ bool first(uint32_t r, uint32_t* s) const {
    *s = slot_of_row[r];
    return *s != kNull && (0xFFFFFFFFu - tab[*s].first_inv) == r;
  }
  __device__ Pair value(uint32_t r) const {
    uint32_t s;
    return first(r, &s) ? Pair{1u, (unsigned long long)nframes[r]} : Pair{0u, 0ull};
  }
  __device__ void emit(uint32_t r, Pair ex, Pair v) const {
    if (!v.cnt) return;
    uint32_t s = slot_of_row[r];
    uniq_row[ex.cnt] = r;
    uniq_count[ex.cnt] = tab[s].count;
    tab[s].offset = (uint32_t)ex.fr;  // startOffset := indices.Len() (arrow_v2.go:302)
    tab[s].size = nframes[r];
  }
  __device__ void total(int, Pair t) const {
    ctr->n_unique = t.cnt;
    ctr->n_indices64 = t.fr;
    if (t.fr > 0x7FFFFFFFull) atomicOr(&ctr->err, ERR_INDEX_OVERFLOW);  // ListView offsets are int32 (arrow_v2.go:233)
  }
};

// (2) per-row ListView offset/size: hit => reuse the first occurrence's (offset,size) (arrow_v2.go:293-299)
__global__ void __launch_bounds__(kThreads) k_rows_materialize(uint32_t n_rows, const uint32_t* slot_of_row, const StackSlot* tab,
                                                               int* st_offsets, int* st_sizes) {
  for (uint32_t r = blockIdx.x * kThreads + threadIdx.x; r < n_rows; r += gridDim.x * kThreads) {
    uint32_t s = slot_of_row[r];
    int o = 0, z = 0;
    if (s != kNull) { o = (int)tab[s].offset; z = (int)tab[s].size; }
    st_offsets[r] = o;
    st_sizes[r] = z;
  }
}

// (3) gather the frames of the unique stacks, in first-occurrence order, into the location-index
// stream and record each frame's first position (appendLocationV2 dedup key = the frame, :421)
__global__ void __launch_bounds__(kThreads) k_gather_unique(const Counters* ctr, const uint32_t* uniq_row, const uint32_t* slot_of_row,
                                                            const StackSlot* tab, const unsigned long long* frames,
                                                            const unsigned long long* frame_off, uint32_t n_frames_registered,
                                                            uint32_t* ustream, uint32_t* loc_first, Counters* ctr_w) {
  uint32_t nu = ctr->n_unique;
  int lane = threadIdx.x & 31;
  uint32_t warp = (blockIdx.x * kThreads + threadIdx.x) >> 5, nwarps = (gridDim.x * kThreads) >> 5;
  for (uint32_t u = warp; u < nu; u += nwarps) {
    uint32_t r = uniq_row[u];
    StackSlot e = tab[slot_of_row[r]];
    const unsigned long long* src = frames + frame_off[r];
    for (uint32_t jx = lane; jx < e.size; jx += 32) {
      unsigned long long fid = src[jx];
      uint32_t pos = e.offset + jx;
      if (fid >= n_frames_registered) { atomicOr(&ctr_w->err, ERR_BAD_FRAME_ID); fid = 0; }
      ustream[pos] = (uint32_t)fid;
      atomicMin(&loc_first[(uint32_t)fid], pos);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// first-occurrence ranking of 32-bit keys (dictionary index assignment). Batched: blockIdx.y = job.
struct FoJob {
  const uint32_t* keys;
  const uint32_t* n_ptr;   // element count lives on the device
  uint32_t* first;         // direct: [universe] first position (memset 0xFF). hashed: unused
  unsigned long long* hslots;  // hashed: (key<<32 | first position), memset 0xFF
  uint32_t hmask;
  uint32_t hashed;
  uint32_t nullable;       // keys == kNull are nulls (skipped, mapped to index 0 + validity 0)
  uint32_t skip_min;       // first[] already filled by a producer kernel
  uint32_t* rank;          // direct: [universe], hashed: [hmask+1]
  uint32_t* order;         // [n_unique] key by rank
  uint32_t* out;           // [n] rank per element (may alias keys); nullptr = skip
  uint32_t* validity;      // [ceil(n/32)] bitmap words; nullptr = skip
  uint32_t* n_unique;      // -> Counters
  uint32_t* n_null;        // -> Counters (may be nullptr)
};
__device__ __forceinline__ uint32_t fo_hfind(const FoJob& j, uint32_t key) {
  uint32_t idx = mix32(key) & j.hmask;
  while ((uint32_t)(j.hslots[idx] >> 32) != key) idx = (idx + 1) & j.hmask;
  return idx;
}
__global__ void __launch_bounds__(kThreads) k_fo_min(const FoJob* jobs) {
  const FoJob& j = jobs[blockIdx.y];
  if (j.skip_min) return;
  uint32_t n = *j.n_ptr;
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    uint32_t key = j.keys[i];
    if (j.nullable && key == kNull) continue;
    if (!j.hashed) {
      if (j.first[key] > i) atomicMin(&j.first[key], i);
    } else {
      unsigned long long packed = ((unsigned long long)key << 32) | i;
      uint32_t idx = mix32(key) & j.hmask;
      while (true) {
        unsigned long long cur = j.hslots[idx];
        if (cur == ~0ull) {
          cur = atomicCAS(&j.hslots[idx], ~0ull, packed);
          if (cur == ~0ull) break;
        }
        if ((uint32_t)(cur >> 32) == key) { if (cur > packed) atomicMin(&j.hslots[idx], packed); break; }
        idx = (idx + 1) & j.hmask;
      }
    }
  }
}
struct FoF {
  typedef uint32_t T;
  const FoJob* jobs;
  __device__ uint32_t n() const { return *jobs[blockIdx.y].n_ptr; }
  __device__ uint32_t value(uint32_t i) const {
    const FoJob& j = jobs[blockIdx.y];
    uint32_t key = j.keys[i];
    if (j.nullable && key == kNull) return 0;
    if (!j.hashed) return j.first[key] == i;
    return (uint32_t)(j.hslots[fo_hfind(j, key)] & 0xFFFFFFFFull) == i;
  }
  __device__ void emit(uint32_t i, uint32_t ex, uint32_t v) const {
    if (!v) return;
    const FoJob& j = jobs[blockIdx.y];
    uint32_t key = j.keys[i];
    j.rank[j.hashed ? fo_hfind(j, key) : key] = ex;
    j.order[ex] = key;
  }
  __device__ void total(int job, uint32_t t) const { *jobs[job].n_unique = t; }
};
// element -> dictionary index (+ validity bitmap, null count). Runs after k_scan_emit<FoF>.
__global__ void __launch_bounds__(kThreads) k_fo_map(const FoJob* jobs) {
  const FoJob& j = jobs[blockIdx.y];
  if (!j.out) return;
  uint32_t n = *j.n_ptr, begin, end;
  block_range(n, &begin, &end);
  uint32_t nulls = 0;
  for (uint32_t tile = begin; tile < end; tile += kThreads) {
    uint32_t i = tile + threadIdx.x;
    bool in = i < end;
    bool valid = false;
    if (in) {
      uint32_t key = j.keys[i];
      valid = !(j.nullable && key == kNull);
      j.out[i] = valid ? j.rank[j.hashed ? fo_hfind(j, key) : key] : 0u;
      nulls += valid ? 0u : 1u;
    }
    unsigned bits = __ballot_sync(0xFFFFFFFFu, valid);
    if (j.validity && (threadIdx.x & 31) == 0 && tile + (threadIdx.x & ~31) < end) j.validity[i >> 5] = bits;
  }
  if (j.n_null) {
    for (int d = 16; d > 0; d >>= 1) nulls += __shfl_down_sync(0xFFFFFFFFu, nulls, d);
    if ((threadIdx.x & 31) == 0 && nulls) atomicAdd(j.n_null, nulls);
  }
}

// ---------------------------------------------------------------------------------------------
// location dictionary: resolved per-frame attributes gathered in location order
struct FrameTable {  // device copy of the registered, pre-resolved frames (see agg.cu resolve_frame)
  const unsigned long long* addr;
  const uint32_t* type_cid;
  const uint32_t* map_cid;
  const uint32_t* bid_cid;   // kNull = null
  const unsigned long long* line;
  const uint32_t* func;      // kNull = no line (native / oomprof)
};
struct LocOut {
  unsigned long long* address;  // Arrow: location.address
  uint32_t* type_key;
  uint32_t* map_key;
  uint32_t* bid_key;
  int* line_off;                // Arrow: lines ListView offsets
  int* line_size;               // Arrow: lines ListView sizes
  uint32_t* line_valid;         // Arrow: lines validity words
  unsigned long long* line_no;  // Arrow: line.line
  uint32_t* func_key;
};
struct LocLinesF {  // scan of has_line over locations (lineListOffsets, parca_reporter.go:428)
  typedef uint32_t T;
  const Counters* ctr;
  Counters* ctr_w;
  const uint32_t* loc_order;
  FrameTable ft;
  LocOut o;
  __device__ uint32_t n() const { return ctr->n_locations; }
  __device__ uint32_t value(uint32_t i) const { return ft.func[loc_order[i]] != kNull; }
  __device__ void emit(uint32_t i, uint32_t ex, uint32_t v) const {
    uint32_t fid = loc_order[i];
    o.address[i] = ft.addr[fid];
    o.type_key[i] = ft.type_cid[fid];
    o.map_key[i] = ft.map_cid[fid];
    o.bid_key[i] = ft.bid_cid[fid];
    o.line_off[i] = (int)ex;
    o.line_size[i] = (int)v;
    if (v) { o.line_no[ex] = ft.line[fid]; o.func_key[ex] = ft.func[fid]; }
  }
  __device__ void total(int, uint32_t t) const { ctr_w->n_lines = t; }
};
// validity words of the lines ListView (null where the location has no line, arrow_v2.go:403-418)
__global__ void __launch_bounds__(kThreads) k_line_validity(const Counters* ctr, const int* line_size, uint32_t* words) {
  uint32_t n = ctr->n_locations;
  uint32_t nw = (n + 31) / 32;
  int lane = threadIdx.x & 31;
  uint32_t warp = (blockIdx.x * kThreads + threadIdx.x) >> 5, nwarps = (gridDim.x * kThreads) >> 5;
  for (uint32_t w = warp; w < nw; w += nwarps) {
    uint32_t i = w * 32 + lane;
    unsigned bits = __ballot_sync(0xFFFFFFFFu, i < n && line_size[i] > 0);
    if (lane == 0) words[w] = bits;
  }
}
// function table in function-dictionary order: filename keys for the nested dictionary
__global__ void __launch_bounds__(kThreads) k_func_keys(const Counters* ctr, const uint32_t* func_order, const uint32_t* fn_file_cid,
                                                        uint32_t* file_key) {
  uint32_t n = ctr->n_functions;
  for (uint32_t k = blockIdx.x * kThreads + threadIdx.x; k < n; k += gridDim.x * kThreads) file_key[k] = fn_file_cid[func_order[k]];
}

// ---------------------------------------------------------------------------------------------
// run-end encoding of every REE column in two fused passes over the rows
// (label columns: reporter/arrow.go:97-131; constant-ish columns: arrow.go:50-59,:166-207)
enum : uint32_t { COL_LS = 0, COL_CPU = 1, COL_TID = 2, COL_COMM = 3, COL_KIND = 4 };
struct ReeCol {
  uint32_t type, param;
  int* run_ends;       // Arrow: run_ends child
  uint32_t* run_keys;  // key of each run (kNull = null run)
};
struct ReeArgs {
  uint32_t n_rows, ncols;
  const ReeCol* cols;
  const uint32_t* ls; const uint32_t* cpu; const uint32_t* tid; const uint32_t* comm; const uint8_t* kind;
  const uint32_t* lsmat; uint32_t n_lscols;   // [labelset][label column] -> value local id or kNull
  const uint32_t* kindtab;                    // [column param][8] -> class id or kNull
  uint32_t* partial;                          // [ncols][grid]
  Counters* ctr;
};
__device__ __forceinline__ uint32_t ree_key(const ReeArgs& a, const ReeCol& c, uint32_t r, bool* null) {
  uint32_t v;
  switch (c.type) {
    case COL_LS: v = a.lsmat[(size_t)a.ls[r] * a.n_lscols + c.param]; *null = v == kNull; break;
    case COL_CPU: v = a.cpu[r]; *null = false; break;
    case COL_TID: v = a.tid[r]; *null = false; break;
    case COL_COMM: v = a.comm[r]; *null = v == 0; break;  // labels.Builder.Set(name, "") deletes the label
    default: v = a.kindtab[c.param * 8 + a.kind[r]]; *null = v == kNull; break;
  }
  return v;
}
__device__ __forceinline__ bool ree_boundary(const ReeArgs& a, const ReeCol& c, uint32_t r, uint32_t* key, bool* null) {
  *key = ree_key(a, c, r, null);
  if (r == 0 || *null) return true;
  bool pnull;
  uint32_t pk = ree_key(a, c, r - 1, &pnull);
  return pnull || pk != *key;
}
__global__ void __launch_bounds__(kThreads) k_ree_count(ReeArgs a) {
  __shared__ uint32_t s_cnt[kMaxCols], s_last[kMaxCols];
  if (threadIdx.x < kMaxCols) { s_cnt[threadIdx.x] = 0; s_last[threadIdx.x] = 0; }
  __syncthreads();
  uint32_t begin, end;
  block_range(a.n_rows, &begin, &end);
  int lane = threadIdx.x & 31;
  for (uint32_t tile = begin; tile < end; tile += kThreads) {
    uint32_t r = tile + threadIdx.x;
    bool in = r < end;
    for (uint32_t c = 0; c < a.ncols; c++) {
      uint32_t key; bool null = true; bool b = false;
      if (in) b = ree_boundary(a, a.cols[c], r, &key, &null);
      unsigned nn = __ballot_sync(0xFFFFFFFFu, in && !null);
      if (lane == 0 && nn) atomicMax(&s_last[c], tile + (threadIdx.x & ~31u) + (32 - __clz(nn)));
      int cnt = __syncthreads_count(b);
      if (threadIdx.x == 0) s_cnt[c] += (uint32_t)cnt;
    }
  }
  __syncthreads();
  if (threadIdx.x < a.ncols) {
    a.partial[threadIdx.x * gridDim.x + blockIdx.x] = s_cnt[threadIdx.x];
    if (s_last[threadIdx.x]) atomicMax(&a.ctr->last_nonnull_plus1[threadIdx.x], s_last[threadIdx.x]);
  }
}
__global__ void k_ree_scan_partials(ReeArgs a, int g) {  // grid = ncols, 32 threads
  uint32_t* p = a.partial + (size_t)blockIdx.x * g;
  int lane = threadIdx.x;
  uint32_t run = 0;
  for (int base = 0; base < g; base += 32) {
    int i = base + lane;
    uint32_t v = i < g ? p[i] : 0u, inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
    if (i < g) p[i] = run + inc - v;
    run += __shfl_sync(0xFFFFFFFFu, inc, 31);
  }
  if (lane == 0) {
    a.ctr->n_runs[blockIdx.x] = run;
    if (run) a.cols[blockIdx.x].run_ends[run - 1] = (int)a.n_rows;  // the last run ends at the row count
  }
}
__global__ void __launch_bounds__(kThreads) k_ree_emit(ReeArgs a) {
  __shared__ uint32_t s_base[kMaxCols];
  __shared__ uint32_t s_wt[kMaxCols][kWarps];
  if (threadIdx.x < a.ncols) s_base[threadIdx.x] = a.partial[threadIdx.x * gridDim.x + blockIdx.x];
  __syncthreads();
  uint32_t begin, end;
  block_range(a.n_rows, &begin, &end);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (uint32_t tile = begin; tile < end; tile += kThreads) {
    uint32_t r = tile + threadIdx.x;
    bool in = r < end;
    for (uint32_t c = 0; c < a.ncols; c++) {  // pass 1: per-warp boundary counts
      uint32_t key; bool null; bool b = in && ree_boundary(a, a.cols[c], r, &key, &null);
      unsigned m = __ballot_sync(0xFFFFFFFFu, b);
      if (lane == 0) s_wt[c][w] = (uint32_t)__popc(m);
    }
    __syncthreads();
    for (uint32_t c = 0; c < a.ncols; c++) {  // pass 2: rank within the column and write
      const ReeCol& col = a.cols[c];
      uint32_t key = 0; bool null = false; bool b = in && ree_boundary(a, col, r, &key, &null);
      unsigned m = __ballot_sync(0xFFFFFFFFu, b);
      if (b) {
        uint32_t k = s_base[c] + (uint32_t)__popc(m & ((1u << lane) - 1u));
        for (int i = 0; i < w; i++) k += s_wt[c][i];
        if (k > 0) col.run_ends[k - 1] = (int)r;  // run k starts at r => run k-1 ends at r
        col.run_keys[k] = null ? kNull : key;
      }
    }
    __syncthreads();
    if (threadIdx.x < a.ncols) {
      uint32_t t = 0;
      for (int i = 0; i < kWarps; i++) t += s_wt[threadIdx.x][i];
      s_base[threadIdx.x] += t;
    }
    __syncthreads();
  }
}

}  // namespace pa
