/*
 * oracle/xxh64.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Restatement of the published XXH64 algorithm (Yann Collet, xxHash specification,
 * doc/xxhash_spec.md, "XXH64 algorithm description"). The reference repository does not
 * contain a stack hash on this path: trace.Hash arrives precomputed
 * (reporter/parca_reporter.go:224, :394-395); the only in-tree hash call is
 * traceutil.HashTrace in parcagpu/parcagpu.go:48,:119, which lives in the un-vendored
 * github.com/parca-dev/opentelemetry-ebpf-profiler fork (go.mod:39,:186).
 * PA_HASH_XXH64X2 is therefore this project's own definition (SURVEY §8c "Hash mode decision"),
 * pinned by the known answers in tests/golden/xxh64_kat.json (generated with python-xxhash 3.7.0).
 */
#ifndef ORACLE_XXH64_H
#define ORACLE_XXH64_H
#include <stdint.h>
#include <string.h>

static const uint64_t ORC_P1 = 11400714785074694791ULL;
static const uint64_t ORC_P2 = 14029467366897019727ULL;
static const uint64_t ORC_P3 = 1609587929392839161ULL;
static const uint64_t ORC_P4 = 9650029242287828579ULL;
static const uint64_t ORC_P5 = 2870177450012600261ULL;

static inline uint64_t orc_rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t orc_rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t orc_rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t orc_xxh_round(uint64_t acc, uint64_t in) {
  acc += in * ORC_P2;
  acc = orc_rotl64(acc, 31);
  return acc * ORC_P1;
}
static inline uint64_t orc_xxh_merge(uint64_t h, uint64_t v) {
  v = orc_xxh_round(0, v);
  h ^= v;
  return h * ORC_P1 + ORC_P4;
}

static inline uint64_t orc_xxh64_impl(const void* data, uint64_t len, uint64_t seed) {
  const uint8_t* p = (const uint8_t*)data;
  const uint8_t* end = p + len;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = seed + ORC_P1 + ORC_P2, v2 = seed + ORC_P2, v3 = seed, v4 = seed - ORC_P1;
    const uint8_t* limit = end - 32;
    do {
      v1 = orc_xxh_round(v1, orc_rd64(p));
      v2 = orc_xxh_round(v2, orc_rd64(p + 8));
      v3 = orc_xxh_round(v3, orc_rd64(p + 16));
      v4 = orc_xxh_round(v4, orc_rd64(p + 24));
      p += 32;
    } while (p <= limit);
    h = orc_rotl64(v1, 1) + orc_rotl64(v2, 7) + orc_rotl64(v3, 12) + orc_rotl64(v4, 18);
    h = orc_xxh_merge(h, v1);
    h = orc_xxh_merge(h, v2);
    h = orc_xxh_merge(h, v3);
    h = orc_xxh_merge(h, v4);
  } else {
    h = seed + ORC_P5;
  }
  h += len;
  while (p + 8 <= end) {
    h ^= orc_xxh_round(0, orc_rd64(p));
    h = orc_rotl64(h, 27) * ORC_P1 + ORC_P4;
    p += 8;
  }
  if (p + 4 <= end) {
    h ^= (uint64_t)orc_rd32(p) * ORC_P1;
    h = orc_rotl64(h, 23) * ORC_P2 + ORC_P3;
    p += 4;
  }
  while (p < end) {
    h ^= (uint64_t)(*p) * ORC_P5;
    h = orc_rotl64(h, 11) * ORC_P1;
    p++;
  }
  h ^= h >> 33;
  h *= ORC_P2;
  h ^= h >> 29;
  h *= ORC_P3;
  h ^= h >> 32;
  return h;
}
#endif
