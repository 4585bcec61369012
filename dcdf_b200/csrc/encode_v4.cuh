// encode_v4.cuh -- Chunk::build for full 64x64 tiles whose fixed-point values fit 31 bits: one 64-thread CTA
// per (subchunk x time slice) unit, 64 cells (one level-3 quadtree node) per thread.
//
// Same reference functions as encode_tile.cuh (chunk.rs:42-96, snapshot.rs:108-156 + 439-500, log.rs:112-165 +
// 725-817, bitmap.rs:66-113, dac.rs:96-132 and the serializers), different mapping:
//   * thread t owns the level-3 node with Morton index t: 8x8 cells, kept as fixed-point quads in shared memory
//     (one image for the instant, one for the snapshot of the current block; a new snapshot just swaps the two)
//     and walked by short rolled loops -- the fully unrolled register version ran out of instruction cache.
//     Levels 6..3 of the min/max/equal pyramid never leave the thread, levels 2 and 1 take two shuffle rounds,
//     the root one 4-record exchange through shared memory.
//   * every DAC entry is classified by length ONCE into bit masks (entry longer than 1 / 2 / 3 bytes); the sizes
//     of both candidate encodings, the BFS layout and every output position follow from popcounts of those
//     masks, one packed warp scan and one cross-warp exchange.
//   * emission writes the exact byte stream into a shared-memory image: bitmaps (nodemap, equal, DAC
//     continuation bits) with atomicOr of whole runs, then (after a barrier, so no word-wide atomic races a
//     byte store) the DAC bytes, sibling groups of four as one 32-bit store; rank directories are summed from
//     the finished words; the image is copied to the arena with 16-byte stores.  Structures larger than the
//     image are emitted straight into the (pre-zeroed) arena piece by the same code.
// Bit convention for all masks: entry i of an N-entry group is bit N-1-i, so runs are MSB-first like the
// serialized bitmaps (bitmap.rs:44-64).
#pragma once
#include "encode_tile.cuh"

namespace dcdf {

constexpr int E4_THREADS = 64;
constexpr int E4_POOL = 16 * 1024;

// packed per-level counters: tree levels 2..5 (block totals: <=16, <=64, <=256, <=1024)
constexpr u32 E4_F2 = 1u, E4_F3 = 1u << 5, E4_F4 = 1u << 12, E4_F5 = 1u << 21;
DCDF_DEVINL u32 e4_f2(u32 x) { return x & 31u; }
DCDF_DEVINL u32 e4_f3(u32 x) { return (x >> 5) & 127u; }
DCDF_DEVINL u32 e4_f4(u32 x) { return (x >> 12) & 511u; }
DCDF_DEVINL u32 e4_f5(u32 x) { return x >> 21; }
DCDF_DEVINL u32 e4_fsum(u32 x) { return e4_f2(x) + e4_f3(x) + e4_f4(x) + e4_f5(x); }

struct E4Smem {
  __align__(16) u8 pool[E4_POOL];
  int4 cell[2][16][E4_THREADS];  // fixed-point cells of the instant and of the block's snapshot: [image][quad q][thread t]
  int2 l4[2][4][E4_THREADS];     // (max, min) of the four level-4 nodes of each image
  int4 rec1[4];       // level-1 nodes: tmax, tmin, first-cell diff, equal
  int2 ent1[4];       // level-1 log entries: tmax - smax, tmin - smin
  u32 wt[2][2][10];   // [warp][0 = snapshot, 1 = log][STRUCT, a1, b1, c1, a2, b2, c2, a3, b3, c3]
  u32 hi[2][2];       // [warp][candidate]: words 4..9 are valid (some entry below the level-1 nodes is longer than two bytes)
  u32 bm_off[12], bm_len[12];
  u32 n_bm;
  unsigned long long piece_off;
};

// entry (a signed difference) needs more than J bytes as a zigzag DAC code (dac.rs:109-121,134-137)
template <int J>
DCDF_DEVINL bool e4_longer(int e) {
  return J == 1 ? (u32)e + 128u > 255u : J == 2 ? (u32)e + 32768u > 65535u : (u32)e + 8388608u > 16777215u;
}
DCDF_DEVINL bool e4_longer_j(int e, int j) { return j == 1 ? e4_longer<1>(e) : j == 2 ? e4_longer<2>(e) : e4_longer<3>(e); }

DCDF_DEVINL u32 e4_expand4(u32 x) {  // bit i -> nibble i
  x &= 0xfu;
  const u32 t = (x | (x << 3) | (x << 6) | (x << 9)) & 0x1111u;
  return t * 0xfu;
}
DCDF_DEVINL u32 e4_expand8(u32 y) {
  u32 t = y & 0xffu;
  t = (t | (t << 12)) & 0x000f000fu;
  t = (t | (t << 6)) & 0x03030303u;
  t = (t | (t << 3)) & 0x11111111u;
  return t * 0xfu;
}
DCDF_DEVINL u64 e4_expand16(u32 x) { return ((u64)e4_expand8(x >> 8) << 32) | (u64)e4_expand8(x); }

DCDF_DEVINL u32 e4_bswap(u32 x) { return __byte_perm(x, 0u, 0x0123u); }

// OR without a return value; the staged image lives in shared memory, where the generic atomic is several times slower
// than the shared-space reduction.
DCDF_DEVINL void e4_red_or(u32* p, u32 v) {
  if (__isShared(p)) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(p)), "r"(v) : "memory");
  else atomicOr(p, v);
}

// OR a run of up to 64 bits (left-aligned in V: the first stream bit is bit 63) into the bit stream that starts
// at byte `bytes` (any alignment), at stream bit `bitpos`.  Works on the aligned 32-bit containers.
DCDF_DEVINL void e4_or_run(u8* bytes, u32 bitpos, u64 V) {
  if (V == 0) return;
  const uintptr_t A = (uintptr_t)bytes;
  u32* w = reinterpret_cast<u32*>(A & ~(uintptr_t)3);
  const u32 g = bitpos + 8u * (u32)(A & 3u);
  w += g >> 5;
  const u32 s = g & 31u;
  const u32 hi = (u32)(V >> 32), lo = (u32)V;
  const u32 w0 = hi >> s;
  const u32 w1 = __funnelshift_r(lo, hi, s);
  const u32 w2 = __funnelshift_r(0u, lo, s);
  if (w0) e4_red_or(w, e4_bswap(w0));
  if (w1) e4_red_or(w + 1, e4_bswap(w1));
  if (w2) e4_red_or(w + 2, e4_bswap(w2));
}
DCDF_DEVINL void e4_or_bits(u8* bytes, u32 bitpos, u32 bits, int n) {  // n <= 32 bits, right-aligned in `bits`
  if (n > 0 && bits) e4_or_run(bytes, bitpos, (u64)bits << (64 - n));
}
DCDF_DEVINL void e4_set_bit(u8* bytes, u32 bitpos) { e4_or_run(bytes, bitpos, 1ull << 63); }

DCDF_DEVINL void e4_store_be32(u8* p, u32 v) {
  if ((((uintptr_t)p) & 3u) == 0) *reinterpret_cast<u32*>(p) = e4_bswap(v);
  else store_be32(p, v);
}

DCDF_DEVINL u32 e4_warp_excl(u32 x, int lane) {
  u32 v = x;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u32 n = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += n;
  }
  return v - x;
}

// stream layout of one DAC (dac.rs:37-44): offsets of each level's bitmap header and bytes
struct E4Dac {
  u32 hdr[4];    // bitmap header (length, k) of level j
  u32 bytes[4];  // first byte of level j
  int levels;
};
DCDF_DEVINL u32 e4_lay_dac(E4Dac& D, u32 off, const u32* cnt) {
  D.levels = 0;
  off += 1;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    D.hdr[j] = off;
    D.bytes[j] = off;
    if (cnt[j] > 0) {
      D.levels = j + 1;
      D.bytes[j] = off + bitmap_size(cnt[j]);
      off = D.bytes[j] + cnt[j];
    }
  }
  return off;
}
DCDF_DEVINL u32 e4_dac_size(const u32* cnt) {
  u32 s = 1;
#pragma unroll
  for (int j = 0; j < 4; j++)
    if (cnt[j] > 0) s += bitmap_size(cnt[j]) + cnt[j];
  return s;
}
DCDF_DEVINL u8* e4_words_of(u8* out, u32 hdr, u32 len) { return out + hdr + 8u + 4u * (len >> 7); }

// Clipped tiles (FULL == false): cells outside the raster are None (snapshot.rs:448-456) -- kept as INT32_MIN in the cell
// images, nodes without any cell inside as (max, min) = (INT32_MIN, INT32_MAX).  max() / min() then skip them by
// themselves; every place that turns a max into an entry maps None to 0 (`None => 0`, snapshot.rs:122-130, log.rs:128-135).
constexpr int E4_NONE = INT32_MIN;
template <bool FULL>
DCDF_DEVINL int e4_o0(int v) { return FULL ? v : (v == E4_NONE ? 0 : v); }
DCDF_DEVINL int e4_sub(int a, int b) { return (int)((u32)a - (u32)b); }  // wraps (operands may be sentinels)
template <bool FULL>
DCDF_DEVINL bool e4_unif(int mx, int mn) { return FULL ? mx == mn : (mx == mn || mx == E4_NONE); }  // a None node is uniform
template <bool FULL>
DCDF_DEVINL void e4_qmm(const int4& t, int& qmax, int& qmin) {
  qmax = max(max(t.x, t.y), max(t.z, t.w));
  if (FULL) {
    qmin = min(min(t.x, t.y), min(t.z, t.w));
  } else {
    const int a = t.x == E4_NONE ? INT32_MAX : t.x, b = t.y == E4_NONE ? INT32_MAX : t.y;
    const int c = t.z == E4_NONE ? INT32_MAX : t.z, d = t.w == E4_NONE ? INT32_MAX : t.w;
    qmin = min(min(a, b), min(c, d));
  }
}

// The masks that describe one candidate encoding of the current instant, as seen by one thread.
struct E4Cand {
  u32 in5, in4;       // internal flags: 16 quads (quad q = bit 15-q), 4 level-4 nodes (node a = bit 3-a)
  bool in3, in2, in1; // own level-3 node, level-2 node of the 4-lane group, level-1 node of the half warp
  u64 ml[3];          // leaves longer than j+1 bytes (cell m = bit 63-m)
  u32 mq[3];          // quads: max entries bits 15..0, min entries bits 31..16
  u32 mu[3];          // level 4 max (bits 3..0), level 4 min (7..4), level 3 max (8) / min (9), level 2 max (10) / min (11)
};

// Snapshot entry masks for length class J (snapshot.rs:122-147: parent_max - child_max, child_min - parent_min).
template <int J, bool FULL>
DCDF_DEVINL void e4_snap_masks(const E4Smem& S, int cur, int tid, int t3max, int t3min, int t2max, int t2min, int t1max, int t1min,
                               u64& ml, u32& mq, u32& mu) {
  ml = 0; mq = 0; mu = 0;
#pragma unroll 1
  for (int a = 0; a < 4; a++) {
    const int2 n4 = S.l4[cur][a][tid];
    u32 leaf16 = 0, qx = 0, qn = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int4 t = S.cell[cur][4 * a + b][tid];
      int qmax, qmin;
      e4_qmm<FULL>(t, qmax, qmin);
      if (e4_longer<J>(e4_sub(qmax, e4_o0<FULL>(t.x)))) leaf16 |= 1u << (15 - 4 * b);  // a cell outside the raster counts as 0
      if (e4_longer<J>(e4_sub(qmax, e4_o0<FULL>(t.y)))) leaf16 |= 1u << (14 - 4 * b);
      if (e4_longer<J>(e4_sub(qmax, e4_o0<FULL>(t.z)))) leaf16 |= 1u << (13 - 4 * b);
      if (e4_longer<J>(e4_sub(qmax, e4_o0<FULL>(t.w)))) leaf16 |= 1u << (12 - 4 * b);
      if (e4_longer<J>(e4_sub(n4.x, e4_o0<FULL>(qmax)))) qx |= 1u << (3 - b);
      if (e4_longer<J>(e4_sub(qmin, n4.y))) qn |= 1u << (3 - b);
    }
    ml |= (u64)leaf16 << (48 - 16 * a);
    mq |= (qx << (12 - 4 * a)) | (qn << (28 - 4 * a));
    if (e4_longer<J>(e4_sub(t3max, e4_o0<FULL>(n4.x)))) mu |= 1u << (3 - a);
    if (e4_longer<J>(e4_sub(n4.y, t3min))) mu |= 1u << (7 - a);
  }
  if (e4_longer<J>(e4_sub(t2max, e4_o0<FULL>(t3max)))) mu |= 1u << 8;
  if (e4_longer<J>(e4_sub(t3min, t2min))) mu |= 1u << 9;
  if (e4_longer<J>(e4_sub(t1max, e4_o0<FULL>(t2max)))) mu |= 1u << 10;
  if (e4_longer<J>(e4_sub(t2min, t1min))) mu |= 1u << 11;
}

// Log entry masks for length class J (log.rs:128-158: max_t - max_s, min_t - min_s per node, t - s per leaf).
template <int J, bool FULL>
DCDF_DEVINL void e4_log_masks(const E4Smem& S, int cur, int ref, int tid, int t3max, int t3min, int t2max, int t2min, int s3max,
                              int s3min, int s2max, int s2min, u64& ml, u32& mq, u32& mu) {
  ml = 0; mq = 0; mu = 0;
#pragma unroll 1
  for (int a = 0; a < 4; a++) {
    const int2 n4 = S.l4[cur][a][tid], s4 = S.l4[ref][a][tid];
    u32 leaf16 = 0, qx = 0, qn = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int4 t = S.cell[cur][4 * a + b][tid], s = S.cell[ref][4 * a + b][tid];
      int qmax, qmin, sqmax, sqmin;
      e4_qmm<FULL>(t, qmax, qmin);
      e4_qmm<FULL>(s, sqmax, sqmin);
      if (e4_longer<J>(e4_sub(t.x, s.x))) leaf16 |= 1u << (15 - 4 * b);  // None - None = 0 (log.rs:751)
      if (e4_longer<J>(e4_sub(t.y, s.y))) leaf16 |= 1u << (14 - 4 * b);
      if (e4_longer<J>(e4_sub(t.z, s.z))) leaf16 |= 1u << (13 - 4 * b);
      if (e4_longer<J>(e4_sub(t.w, s.w))) leaf16 |= 1u << (12 - 4 * b);
      if (e4_longer<J>(e4_sub(e4_o0<FULL>(qmax), e4_o0<FULL>(sqmax)))) qx |= 1u << (3 - b);
      if (e4_longer<J>(e4_sub(qmin, sqmin))) qn |= 1u << (3 - b);
    }
    ml |= (u64)leaf16 << (48 - 16 * a);
    mq |= (qx << (12 - 4 * a)) | (qn << (28 - 4 * a));
    if (e4_longer<J>(e4_sub(e4_o0<FULL>(n4.x), e4_o0<FULL>(s4.x)))) mu |= 1u << (3 - a);
    if (e4_longer<J>(e4_sub(n4.y, s4.y))) mu |= 1u << (7 - a);
  }
  if (e4_longer<J>(e4_sub(e4_o0<FULL>(t3max), e4_o0<FULL>(s3max)))) mu |= 1u << 8;
  if (e4_longer<J>(e4_sub(t3min, s3min))) mu |= 1u << 9;
  if (e4_longer<J>(e4_sub(e4_o0<FULL>(t2max), e4_o0<FULL>(s2max)))) mu |= 1u << 10;
  if (e4_longer<J>(e4_sub(t2min, s2min))) mu |= 1u << 11;
}

// Per-thread packed counters of one candidate (relative to "the root is an internal node").
DCDF_DEVINL void e4_count(const E4Cand& C, bool owner2, u32* w /* [10] */) {
  const bool e2 = C.in1, e3 = e2 && C.in2, e4 = e3 && C.in3;
  const u32 ai4 = e4 ? C.in4 : 0u;
  const u32 X5 = e4_expand4(ai4);
  const u32 ai5 = C.in5 & X5;
  const u64 X6 = e4_expand16(ai5);
  w[0] = ((owner2 && e2 && C.in2) ? E4_F2 : 0u) + ((e3 && C.in3) ? E4_F3 : 0u) + (u32)__popc(ai4) * E4_F4 + (u32)__popc(ai5) * E4_F5;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const u32 mu = C.mu[j], mq = C.mq[j];
    w[1 + 3 * j] = ((owner2 && e2 && ((mu >> 10) & 1u)) ? E4_F2 : 0u) + ((e3 && ((mu >> 8) & 1u)) ? E4_F3 : 0u) +
                   (e4 ? (u32)__popc(mu & 0xfu) : 0u) * E4_F4 + (u32)__popc(mq & 0xffffu & X5) * E4_F5;
    w[2 + 3 * j] = (u32)__popcll(C.ml[j] & X6);
    w[3 + 3 * j] = ((owner2 && e2 && C.in2 && ((mu >> 11) & 1u)) ? E4_F2 : 0u) + ((e3 && C.in3 && ((mu >> 9) & 1u)) ? E4_F3 : 0u) +
                   (u32)__popc((mu >> 4) & ai4) * E4_F4 + (u32)__popc((mq >> 16) & ai5) * E4_F5;
  }
}

// High bytes (DAC levels 1..3) of a zigzag code longer than one byte; r1..r3 are the running positions of the
// current tree level in those DAC levels.
DCDF_DEVINL void e4_emit_hi(u8* b1, u8* b2, u8* b3, u32 z, u32& r1, u32& r2, u32& r3) {
  b1[r1++] = (u8)(z >> 8);
  if (z >> 16) {
    b2[r2++] = (u8)(z >> 16);
    if (z >> 24) b3[r3++] = (u8)(z >> 24);
  }
}

DCDF_DEVINL void e4_store4(u8* p, u32 z0, u32 z1, u32 z2, u32 z3) {
  if ((((uintptr_t)p) & 3u) == 0) {
    *reinterpret_cast<u32*>(p) = (z0 & 0xffu) | ((z1 & 0xffu) << 8) | ((z2 & 0xffu) << 16) | (z3 << 24);
  } else {
    p[0] = (u8)z0; p[1] = (u8)z1; p[2] = (u8)z2; p[3] = (u8)z3;
  }
}

// popcount of the `n` stream bytes starting at `p` (any alignment)
DCDF_DEVINL u32 e4_popc_bytes16(const u8* p) {
  const uintptr_t A = (uintptr_t)p;
  const u32* w = reinterpret_cast<const u32*>(A & ~(uintptr_t)3);
  const u32 a = (u32)(A & 3u);
  u32 c = __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
  if (a == 0) return c + __popc(w[0]);
  c += __popc(w[0] >> (8u * a));               // drop the `a` low-address bytes of the first container
  c += __popc(w[4] & ((1u << (8u * a)) - 1u)); // keep the `a` low-address bytes of the last one
  return c;
}


// Block totals of one candidate and the entry counts of its two DACs (cmax[j] / cmin[j] = entries longer than j
// bytes), from the per-warp packed counters and the root / level-1 entries (snapshot.rs:71-79, log.rs:77-86).
struct E4Tot {
  u32 tot[10];
  u32 cmax[4], cmin[4];
  u32 I0, I1, nm_len, n_int;
};
DCDF_DEVINL void e4_totals(E4Tot& T, const E4Smem& S, int cand, bool in0, u32 in1m, const int (&e1max)[4], const int (&e1min)[4],
                           int e0max, int e0min) {
  const bool hi = in0 && (S.hi[0][cand] | S.hi[1][cand]) != 0;
#pragma unroll
  for (int i = 0; i < 4; i++) T.tot[i] = in0 ? S.wt[0][cand][i] + S.wt[1][cand][i] : 0u;
#pragma unroll
  for (int i = 4; i < 10; i++) T.tot[i] = 0u;
  if (hi) {
#pragma unroll
    for (int i = 4; i < 10; i++) T.tot[i] = (S.hi[0][cand] ? S.wt[0][cand][i] : 0u) + (S.hi[1][cand] ? S.wt[1][cand][i] : 0u);
  }
  T.I0 = in0 ? 1u : 0u;
  T.I1 = in0 ? (u32)__popc(in1m) : 0u;
  const u32 upper = T.I0 + T.I1 + e4_f2(T.tot[0]) + e4_f3(T.tot[0]) + e4_f4(T.tot[0]);
  T.n_int = upper + e4_f5(T.tot[0]);
  T.nm_len = 1u + 4u * upper;
  T.cmax[0] = 1u + 4u * T.n_int;
  T.cmin[0] = T.n_int;
  // entries of the root and the level-1 nodes
  bool top_hi = e4_longer<2>(e0max) || e4_longer<2>(e0min);
#pragma unroll
  for (int k = 0; k < 4; k++) top_hi = top_hi || e4_longer<2>(e1max[k]) || e4_longer<2>(e1min[k]);
#pragma unroll
  for (int j = 1; j < 4; j++) {
    u32 tm = 0, tn = 0;
    if (j == 1 || hi || top_hi) {
      tm = e4_longer_j(e0max, j) ? 1u : 0u;
      if (in0) {
        tn = e4_longer_j(e0min, j) ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          tm += e4_longer_j(e1max[k], j) ? 1u : 0u;
          tn += (((in1m >> (3 - k)) & 1u) && e4_longer_j(e1min[k], j)) ? 1u : 0u;
        }
      }
    }
    T.cmax[j] = tm + e4_fsum(T.tot[3 * j - 2]) + T.tot[3 * j - 1];
    T.cmin[j] = tn + e4_fsum(T.tot[3 * j]);
  }
}

// Number of max-DAC entries longer than j bytes that precede this thread's entries of tree level `level` (2..6).
DCDF_DEVINL u32 e4_base_x(const E4Tot& T, const u32 (&pre)[10], int j, int level) {
  const u32 t = T.tot[3 * j - 2], p = pre[3 * j - 2];
  u32 b = T.cmax[j] - e4_fsum(t) - T.tot[3 * j - 1];  // root and level-1 entries come first
  if (level == 2) return b + e4_f2(p);
  b += e4_f2(t);
  if (level == 3) return b + e4_f3(p);
  b += e4_f3(t);
  if (level == 4) return b + e4_f4(p);
  b += e4_f4(t);
  if (level == 5) return b + e4_f5(p);
  return b + e4_f5(t) + pre[3 * j - 1];
}
// The same for the min DAC (tree levels 2..5).
DCDF_DEVINL u32 e4_base_n(const E4Tot& T, const u32 (&pre)[10], int j, int level) {
  const u32 t = T.tot[3 * j], p = pre[3 * j];
  u32 b = T.cmin[j] - e4_fsum(t);
  if (level == 2) return b + e4_f2(p);
  b += e4_f2(t);
  if (level == 3) return b + e4_f3(p);
  b += e4_f3(t);
  if (level == 4) return b + e4_f4(p);
  return b + e4_f4(t) + e4_f5(p);
}

// Continuation bits of one root / level-1 entry; p[j] = running position in DAC level j.
DCDF_DEVINL void e4_top_bits(u8* w0, u8* w1, u8* w2, u32 (&p)[4], int e) {
  const bool l1 = e4_longer<1>(e), l2 = e4_longer<2>(e), l3 = e4_longer<3>(e);
  if (l1) e4_set_bit(w0, p[0]);
  p[0]++;
  if (l1) {
    if (l2) e4_set_bit(w1, p[1]);
    p[1]++;
  }
  if (l2) {
    if (l3) e4_set_bit(w2, p[2]);
    p[2]++;
  }
  if (l3) p[3]++;
}

// to_fixed of one cell (a3).  `exact`: the stats pass proved that no value of the unit has more than `bits`
// fractional bits and none is infinite, so n * 2^(bits+1) is an integer and fixed.rs:47-57 never rounds or fails.
template <typename InT>
DCDF_DEVINL int e4_conv(InT x, int bits, bool do_round, bool exact, float scale2, u32& err) {
  return CellConv<InT, int32_t>::get(x, bits, do_round, err);
}
template <>
DCDF_DEVINL int e4_conv<float>(float x, int bits, bool do_round, bool exact, float scale2, u32& err) {
  if (exact) {
    const int f = __float2int_rz(x * scale2) + 1;
    return x != x ? 0 : f;  // NaN -> 0 (fixed.rs:35-37)
  }
  return CellConv<float, int32_t>::get(x, bits, do_round, err);
}

// Two rows of the thread's 8x8 block: x[0..7] = row `row`, x[8..15] = row `row + 1`.  Clipped tiles: `inb` gets one
// bit per cell that lies inside rows x cols (the others are not read).
template <typename InT, bool FULL>
DCDF_DEVINL void e4_fetch_pair(const InT* p, i64 sr, i64 sc, int row, int c0, bool vec, int rows, int cols, InT (&x)[16], u32& inb) {
  inb = 0xffffu;
  if (!FULL) {
    inb = 0;
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        if (sizeof(InT) == 4 && vec && row + rr < rows && c0 + 4 * h + 3 < cols) {  // whole group inside: one 128-bit load
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(p + (i64)(row + rr) * sr + c0 + 4 * h));
          const u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int cc = 0; cc < 4; cc++) x[8 * rr + 4 * h + cc] = *reinterpret_cast<const InT*>(&w[cc]);
          inb |= 0xfu << (8 * rr + 4 * h);
        } else {
#pragma unroll
          for (int cc = 0; cc < 4; cc++) {
            const bool in = row + rr < rows && c0 + 4 * h + cc < cols;
            x[8 * rr + 4 * h + cc] = in ? __ldg(p + (i64)(row + rr) * sr + (i64)(c0 + 4 * h + cc) * sc) : InT(0);
            inb |= in ? 1u << (8 * rr + 4 * h + cc) : 0u;
          }
        }
      }
    }
    return;
  }
#pragma unroll
  for (int rr = 0; rr < 2; rr++) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      if (sizeof(InT) == 4 && vec) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(p + (i64)(row + rr) * sr + c0 + 4 * h));
        const u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int cc = 0; cc < 4; cc++) x[8 * rr + 4 * h + cc] = *reinterpret_cast<const InT*>(&w[cc]);
      } else {
#pragma unroll
        for (int cc = 0; cc < 4; cc++) x[8 * rr + 4 * h + cc] = __ldg(p + (i64)(row + rr) * sr + (i64)(c0 + 4 * h + cc) * sc);
      }
    }
  }
}

template <typename InT, int MINB, bool FULL>
__global__ void __launch_bounds__(E4_THREADS, MINB) k_encode_v4(const EncParams P, const u32 stage_limit) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  E4Smem& S = *reinterpret_cast<E4Smem*>(smem_raw);

  if (blockIdx.x >= *P.order_count) return;
  const u32 unit_idx = P.order[blockIdx.x];
  const EncUnit unit = P.units[unit_idx];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool do_round = unit.flags & UF_ROUND;
  const bool exact = unit.flags & UF_EXACT;
  const float scale2 = (float)((i64)2 << unit.bits);
  const int r0 = 8 * (int)morton_row(tid), c0 = 8 * (int)morton_col(tid);
  const InT* base = static_cast<const InT*>(P.data) + unit.base;
  const bool vec = sizeof(InT) == 4 && P.stride_c == 1 && ((P.stride_r * 4) & 15) == 0 && ((P.stride_t * 4) & 15) == 0 &&
                   ((((uintptr_t)base)) & 15) == 0;
  const bool owner2 = (lane & 3) == 0;
  const int k1 = tid >> 4;  // own level-1 node

  {  // the emission image starts out all zero; every copy-out re-zeroes what it used
    uint4* z = reinterpret_cast<uint4*>(S.pool);
    for (int i = tid; i < E4_POOL / 16; i += E4_THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
#pragma unroll 1
  for (int q = 0; q < 16; q++) S.cell[1][q][tid] = make_int4(0, 0, 0, 0);
#pragma unroll 1
  for (int a = 0; a < 4; a++) S.l4[1][a][tid] = make_int2(0, 0);

  u32 err = 0;
  int cur = 0, ref = 1;  // images: cells of this instant / of the block's snapshot
  int s3max = 0, s3min = 0, s2max = 0, s2min = 0, s1max = 0, s1min = 0, s0max = 0, s0min = 0;
  u32 n_logs = 0, n_snap = 0, n_log_total = 0;
  u64 total_bytes = 0;

  InT nx[16];  // two rows of the block, loaded ahead of their use
  u32 nxb;
  e4_fetch_pair<InT, FULL>(base, P.stride_r, P.stride_c, r0, c0, vec, unit.rows, unit.cols, nx, nxb);

  for (int inst = 0; inst < unit.instants; inst++) {
    const bool first = inst == 0;
    // ---------------- load + convert (a3): two rows at a time, the next pair in flight while this one is converted
    {
      const InT* pi = base + (i64)inst * P.stride_t;
#pragma unroll 1
      for (int rp = 0; rp < 4; rp++) {
        InT cu[16];
        const u32 cub = nxb;
#pragma unroll
        for (int i = 0; i < 16; i++) cu[i] = nx[i];
        // next pair of this instant, or the first pair of the next instant (in flight during the rest of this one)
        if (rp < 3) e4_fetch_pair<InT, FULL>(pi, P.stride_r, P.stride_c, r0 + 2 * (rp + 1), c0, vec, unit.rows, unit.cols, nx, nxb);
        else if (inst + 1 < unit.instants) e4_fetch_pair<InT, FULL>(pi + P.stride_t, P.stride_r, P.stride_c, r0, c0, vec, unit.rows, unit.cols, nx, nxb);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          int4 v;
          v.x = e4_conv<InT>(cu[2 * j], unit.bits, do_round, exact, scale2, err);
          v.y = e4_conv<InT>(cu[2 * j + 1], unit.bits, do_round, exact, scale2, err);
          v.z = e4_conv<InT>(cu[8 + 2 * j], unit.bits, do_round, exact, scale2, err);
          v.w = e4_conv<InT>(cu[8 + 2 * j + 1], unit.bits, do_round, exact, scale2, err);
          if (!FULL) {
            if (!((cub >> (2 * j)) & 1u)) v.x = E4_NONE;
            if (!((cub >> (2 * j + 1)) & 1u)) v.y = E4_NONE;
            if (!((cub >> (8 + 2 * j)) & 1u)) v.z = E4_NONE;
            if (!((cub >> (8 + 2 * j + 1)) & 1u)) v.w = E4_NONE;
          }
          S.cell[cur][8 * (rp >> 1) + 4 * (j >> 1) + 2 * (rp & 1) + (j & 1)][tid] = v;
        }
      }
    }

    // ---------------- levels 6..3 inside the thread: min/max, uniform and equal flags, log length-1 masks
    u32 u5 = 0, eq5 = 0, u4 = 0, eq4 = 0;
    E4Cand L;  // log candidate
    L.ml[0] = 0; L.mq[0] = 0; L.mu[0] = 0;
    int t3max = INT32_MIN, t3min = INT32_MAX, diff3 = 0;
    bool eq3 = true;
#pragma unroll 1
    for (int a = 0; a < 4; a++) {
      const int2 s4 = S.l4[ref][a][tid];
      int amax = INT32_MIN, amin = INT32_MAX, dfirst = 0;
      bool aeq = true;
      u32 leaf16 = 0, u5n = 0, e5n = 0, qx = 0, qn = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int4 t = S.cell[cur][4 * a + b][tid], sq = S.cell[ref][4 * a + b][tid];
        int qmax, qmin, sqmax, sqmin;
        e4_qmm<FULL>(t, qmax, qmin);
        e4_qmm<FULL>(sq, sqmax, sqmin);
        const int d0 = e4_sub(t.x, sq.x), d1 = e4_sub(t.y, sq.y), d2 = e4_sub(t.z, sq.z), d3 = e4_sub(t.w, sq.w);  // None - None = 0
        const bool e5 = (((d1 ^ d0) | (d2 ^ d0) | (d3 ^ d0)) == 0);  // all four leaf diffs equal (log.rs:780-806)
        if (e4_unif<FULL>(qmax, qmin)) u5n |= 1u << (3 - b);
        if (e5) e5n |= 1u << (3 - b);
        if (e4_longer<1>(d0)) leaf16 |= 1u << (15 - 4 * b);
        if (e4_longer<1>(d1)) leaf16 |= 1u << (14 - 4 * b);
        if (e4_longer<1>(d2)) leaf16 |= 1u << (13 - 4 * b);
        if (e4_longer<1>(d3)) leaf16 |= 1u << (12 - 4 * b);
        if (e4_longer<1>(e4_sub(e4_o0<FULL>(qmax), e4_o0<FULL>(sqmax)))) qx |= 1u << (3 - b);
        if (e4_longer<1>(e4_sub(qmin, sqmin))) qn |= 1u << (3 - b);
        amax = max(amax, qmax); amin = min(amin, qmin);
        if (b == 0) dfirst = d0;
        aeq = aeq && e5 && d0 == dfirst;
      }
      S.l4[cur][a][tid] = make_int2(amax, amin);
      L.ml[0] |= (u64)leaf16 << (48 - 16 * a);
      L.mq[0] |= (qx << (12 - 4 * a)) | (qn << (28 - 4 * a));
      u5 |= u5n << (12 - 4 * a);
      eq5 |= e5n << (12 - 4 * a);
      if (e4_unif<FULL>(amax, amin)) u4 |= 1u << (3 - a);
      if (aeq) eq4 |= 1u << (3 - a);
      if (e4_longer<1>(e4_sub(e4_o0<FULL>(amax), e4_o0<FULL>(s4.x)))) L.mu[0] |= 1u << (3 - a);
      if (e4_longer<1>(e4_sub(amin, s4.y))) L.mu[0] |= 1u << (7 - a);
      t3max = max(t3max, amax); t3min = min(t3min, amin);
      if (a == 0) diff3 = dfirst;
      eq3 = eq3 && aeq && dfirst == diff3;
    }
    const bool u3 = e4_unif<FULL>(t3max, t3min);
    if (e4_longer<1>(e4_sub(e4_o0<FULL>(t3max), e4_o0<FULL>(s3max)))) L.mu[0] |= 1u << 8;
    if (e4_longer<1>(e4_sub(t3min, s3min))) L.mu[0] |= 1u << 9;

    // ---------------- levels 2 and 1 with shuffles (4 resp. 16 consecutive lanes)
    int t2max = max(t3max, shfl_xor(t3max, 1)); t2max = max(t2max, shfl_xor(t2max, 2));
    int t2min = min(t3min, shfl_xor(t3min, 1)); t2min = min(t2min, shfl_xor(t2min, 2));
    const int diff2 = shfl(diff3, lane & ~3);
    const u32 ok3 = __ballot_sync(0xffffffffu, eq3 && diff3 == diff2);
    const bool eq2 = ((ok3 >> (lane & ~3)) & 0xfu) == 0xfu;
    const bool u2 = e4_unif<FULL>(t2max, t2min);
    int t1max = max(t2max, shfl_xor(t2max, 4)); t1max = max(t1max, shfl_xor(t1max, 8));
    int t1min = min(t2min, shfl_xor(t2min, 4)); t1min = min(t1min, shfl_xor(t1min, 8));
    const int diff1 = shfl(diff2, lane & ~15);
    const u32 ok2 = __ballot_sync(0xffffffffu, eq2 && diff2 == diff1);
    const bool eq1 = ((ok2 >> (lane & ~15)) & 0xffffu) == 0xffffu;
    const bool u1 = e4_unif<FULL>(t1max, t1min);
    if (e4_longer<1>(e4_sub(e4_o0<FULL>(t2max), e4_o0<FULL>(s2max)))) L.mu[0] |= 1u << 10;
    if (e4_longer<1>(e4_sub(t2min, s2min))) L.mu[0] |= 1u << 11;

    // log entries longer than two bytes can only exist where the level-2 value ranges are that far apart
    L.ml[1] = L.ml[2] = 0; L.mq[1] = L.mq[2] = 0; L.mu[1] = L.mu[2] = 0;
    {
      const bool far = (((u32)e4_sub(t2min, s2max) + 32768u) | ((u32)e4_sub(t2max, s2min) + 32768u)) > 65535u;
      if (!first && __any_sync(0xffffffffu, far)) {
        e4_log_masks<2, FULL>(S, cur, ref, tid, t3max, t3min, t2max, t2min, s3max, s3min, s2max, s2min, L.ml[1], L.mq[1], L.mu[1]);
        e4_log_masks<3, FULL>(S, cur, ref, tid, t3max, t3min, t2max, t2min, s3max, s3min, s2max, s2min, L.ml[2], L.mq[2], L.mu[2]);
      }
    }
    // structure flags: snapshot internal = !uniform (snapshot.rs:133); log internal = !uniform && !equal (log.rs:137-152)
    E4Cand C;  // snapshot candidate (entry masks filled in on the slow path only)
    C.in5 = ~u5 & 0xffffu; C.in4 = ~u4 & 0xfu; C.in3 = !u3; C.in2 = !u2; C.in1 = !u1;
    L.in5 = ~u5 & ~eq5 & 0xffffu; L.in4 = ~u4 & ~eq4 & 0xfu; L.in3 = !u3 && !eq3; L.in2 = !u2 && !eq2; L.in1 = !u1 && !eq1;
#pragma unroll
    for (int j = 0; j < 3; j++) { C.ml[j] = 0; C.mq[j] = 0; C.mu[j] = 0; }

    // ---------------- per-warp totals and the four level-1 records go through shared memory
    u32 wl[10], ws[10];
    e4_count(L, owner2, wl);
    e4_count(C, owner2, ws);
    {
      const bool hi_l = __any_sync(0xffffffffu, (L.ml[1] | (u64)L.mq[1] | (u64)L.mu[1]) != 0);
      u32 r = __reduce_add_sync(0xffffffffu, ws[0]);
      if (lane == 0) S.wt[warp][0][0] = r;
#pragma unroll
      for (int i = 0; i < 10; i++) {
        if (i < 4 || hi_l) {
          r = __reduce_add_sync(0xffffffffu, wl[i]);
          if (lane == 0) S.wt[warp][1][i] = r;
        }
      }
      if (lane == 0) S.hi[warp][1] = hi_l ? 1u : 0u;
    }
    if ((lane & 15) == 0) {
      S.rec1[k1] = make_int4(t1max, t1min, diff1, eq1 ? 1 : 0);
      S.ent1[k1] = make_int2(e4_sub(e4_o0<FULL>(t1max), e4_o0<FULL>(s1max)), e4_sub(t1min, s1min));
    }
    __syncthreads();  // B1

    // ---------------- root (every thread, redundantly)
    int n1max[4], n1min[4], n1dif[4];
    bool n1eq[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int4 r = S.rec1[k];
      n1max[k] = r.x; n1min[k] = r.y; n1dif[k] = r.z; n1eq[k] = r.w != 0;
    }
    const int t0max = max(max(n1max[0], n1max[1]), max(n1max[2], n1max[3]));
    const int t0min = min(min(n1min[0], n1min[1]), min(n1min[2], n1min[3]));
    const bool u0 = t0max == t0min;
    const bool eq0 = n1eq[0] && n1eq[1] && n1eq[2] && n1eq[3] && n1dif[1] == n1dif[0] && n1dif[2] == n1dif[0] && n1dif[3] == n1dif[0];
    u32 l_in1 = 0, s_in1 = 0;  // level-1 internal flags, node k = bit 3-k
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const bool un = e4_unif<FULL>(n1max[k], n1min[k]);
      if (!un) s_in1 |= 1u << (3 - k);
      if (!un && !n1eq[k]) l_in1 |= 1u << (3 - k);
    }
    const bool l_in0 = !u0 && !eq0, s_in0 = !u0;
    int l_e1max[4], l_e1min[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { const int2 e = S.ent1[k]; l_e1max[k] = e.x; l_e1min[k] = e.y; }
    const int l_e0max = t0max - s0max, l_e0min = t0min - s0min;

    E4Tot T;
    bool as_snapshot = first || n_logs == 254u;  // chunk.rs:62 (Block caps logs at 254)
    u32 my_size = 0;
    bool slow = as_snapshot;
    if (!slow) {
      // exact size of the Log against a lower bound of the Snapshot (every entry takes at least one byte)
      e4_totals(T, S, 1, l_in0, l_in1, l_e1max, l_e1min, l_e0max, l_e0min);
      const u32 log_size = 13u + bitmap_size(T.nm_len) + bitmap_size(T.nm_len - T.n_int) + e4_dac_size(T.cmax) + e4_dac_size(T.cmin);
      const u32 st = s_in0 ? S.wt[0][0][0] + S.wt[1][0][0] : 0u;
      const u32 s_upper = (s_in0 ? 1u + (u32)__popc(s_in1) : 0u) + e4_f2(st) + e4_f3(st) + e4_f4(st);
      const u32 s_int = s_upper + e4_f5(st), s_nm = 1u + 4u * s_upper, s_nmax = 1u + 4u * s_int;
      const u32 snap_lb = 13u + bitmap_size(s_nm) + 1u + bitmap_size(s_nmax) + s_nmax + (s_int ? 1u + bitmap_size(s_int) + s_int : 1u);
      slow = snap_lb <= log_size;
      my_size = log_size;
    }
    int s_e1max[4], s_e1min[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { s_e1max[k] = e4_sub(t0max, e4_o0<FULL>(n1max[k])); s_e1min[k] = e4_sub(n1min[k], t0min); }
    if (slow) {
      // ---------------- exact Snapshot size
      const u32 log_size = my_size;
      e4_snap_masks<1, FULL>(S, cur, tid, t3max, t3min, t2max, t2min, t1max, t1min, C.ml[0], C.mq[0], C.mu[0]);
      // entries longer than two bytes need a value range that wide -- or, in a clipped tile, a cell outside the raster
      // next to large values (its entry is the quad's max itself)
      const bool far = !FULL || (u32)e4_sub(t1max, t1min) > 32767u;
      const bool hi_s = __any_sync(0xffffffffu, far);
      if (hi_s) {
        e4_snap_masks<2, FULL>(S, cur, tid, t3max, t3min, t2max, t2min, t1max, t1min, C.ml[1], C.mq[1], C.mu[1]);
        e4_snap_masks<3, FULL>(S, cur, tid, t3max, t3min, t2max, t2min, t1max, t1min, C.ml[2], C.mq[2], C.mu[2]);
      }
      e4_count(C, owner2, ws);
#pragma unroll
      for (int i = 0; i < 10; i++) {
        if (i < 4 || hi_s) {
          const u32 r = __reduce_add_sync(0xffffffffu, ws[i]);
          if (lane == 0) S.wt[warp][0][i] = r;
        }
      }
      if (lane == 0) S.hi[warp][0] = hi_s ? 1u : 0u;
      __syncthreads();  // B1'
      e4_totals(T, S, 0, s_in0, s_in1, s_e1max, s_e1min, t0max, t0min);
      const u32 snap_size = 13u + bitmap_size(T.nm_len) + e4_dac_size(T.cmax) + e4_dac_size(T.cmin);
      as_snapshot = as_snapshot || snap_size <= log_size;  // ties go to the snapshot (chunk.rs:62)
      if (as_snapshot) my_size = snap_size;
      else e4_totals(T, S, 1, l_in0, l_in1, l_e1max, l_e1min, l_e0max, l_e0min);
    }

    // ---------------- the winner, generic from here on
    E4Cand W;
    W.in5 = as_snapshot ? C.in5 : L.in5; W.in4 = as_snapshot ? C.in4 : L.in4;
    W.in3 = as_snapshot ? C.in3 : L.in3; W.in2 = as_snapshot ? C.in2 : L.in2; W.in1 = as_snapshot ? C.in1 : L.in1;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      W.ml[j] = as_snapshot ? C.ml[j] : L.ml[j]; W.mq[j] = as_snapshot ? C.mq[j] : L.mq[j]; W.mu[j] = as_snapshot ? C.mu[j] : L.mu[j];
    }
    int e1x[4], e1n[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { e1x[k] = as_snapshot ? s_e1max[k] : l_e1max[k]; e1n[k] = as_snapshot ? s_e1min[k] : l_e1min[k]; }
    const int e0x = as_snapshot ? t0max : l_e0max, e0n = as_snapshot ? t0min : l_e0min;
    const u32 nm_len = T.nm_len, n_int = T.n_int, I0 = T.I0, I1 = T.I1;
    const bool in0 = as_snapshot ? s_in0 : l_in0;
    const u32 in1m = as_snapshot ? s_in1 : l_in1;
    const int cand = as_snapshot ? 0 : 1;
    const bool staged = my_size <= stage_limit;
    if (tid == 0) {
      const u64 need = ((u64)my_size + 15ull) & ~15ull;
      const u64 off = atomicAdd(P.arena_head, (unsigned long long)need);
      S.piece_off = off;
      Piece pc;
      pc.off = off; pc.size = my_size; pc.kind = as_snapshot ? 1u : 0u;
      P.pieces[unit.piece_base + inst] = pc;
    }
    if (!staged) __syncthreads();
    u64 piece_off = 0;
    bool fits = true;
    u8* out = S.pool;
    if (!staged) {
      piece_off = S.piece_off;
      fits = piece_off + (((u64)my_size + 15ull) & ~15ull) <= P.arena_cap;
      if (!fits) err |= EF_ARENA_FULL;
      out = P.arena + piece_off;
      if (fits) {
        uint4* z = reinterpret_cast<uint4*>(out);
        for (u32 i = tid; i < (my_size + 15u) / 16u; i += E4_THREADS) z[i] = make_uint4(0, 0, 0, 0);
      }
      __syncthreads();
    }
    const bool emit = staged || fits;

    // stream layout (snapshot.rs:48-58 / log.rs:53-64)
    const u32 eq_len = nm_len - n_int;
    const u32 nm_hdr = 13u;
    const u32 eq_hdr = nm_hdr + bitmap_size(nm_len);
    E4Dac DX, DN;
    u32 end = e4_lay_dac(DX, as_snapshot ? eq_hdr : eq_hdr + bitmap_size(eq_len), T.cmax);
    end = e4_lay_dac(DN, end, T.cmin);
    if (end != my_size) err |= EF_BAD_FORMAT;
    u8* const nm_words = e4_words_of(out, nm_hdr, nm_len);
    u8* const eq_words = e4_words_of(out, eq_hdr, eq_len);
    u8* xw[3];  // continuation-bit words of max DAC levels 0..2
    u8* nw[3];
    u8* xb[4];
    u8* nb[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (j < 3) { xw[j] = e4_words_of(out, DX.hdr[j], T.cmax[j]); nw[j] = e4_words_of(out, DN.hdr[j], T.cmin[j]); }
      xb[j] = out + DX.bytes[j]; nb[j] = out + DN.bytes[j];
    }

    // block-wide positions
    const u32 T0 = T.tot[0];
    const u32 I2 = e4_f2(T0), I3 = e4_f3(T0), I4 = e4_f4(T0);
    const u32 Pn2 = 1u + 4u * I0, Pn3 = Pn2 + 4u * I1, Pn4 = Pn3 + 4u * I2, Pn5 = Pn4 + 4u * I3, Pn6 = Pn5 + 4u * I4;
    const u32 Mn2 = I0 + I1, Mn3 = Mn2 + I2, Mn4 = Mn3 + I3, Mn5 = Mn4 + I4;
    const u32 R1 = (u32)__popc(in1m >> (4 - k1));  // internal level-1 nodes before the own one
    const bool x2 = in0 && ((in1m >> (3 - k1)) & 1u);
    const bool x3 = x2 && W.in2, x4 = x3 && W.in3;
    const u32 ai4 = x4 ? W.in4 : 0u;
    const u32 X5 = e4_expand4(ai4);
    const u32 ai5 = W.in5 & X5;
    const u64 X6 = e4_expand16(ai5);
    // prefix over earlier threads (Morton order) of the packed counters
    u32 pre[10];
    const bool any_hi = (T.cmax[2] | T.cmin[2]) != 0;  // block-uniform: some entry takes three bytes or more
#pragma unroll
    for (int i = 0; i < 10; i++) {
      pre[i] = 0;
      if (i < 4 || any_hi) pre[i] = e4_warp_excl(as_snapshot ? ws[i] : wl[i], lane) + ((warp == 1 && in0 && (i < 4 || S.hi[0][cand])) ? S.wt[0][cand][i] : 0u);
    }
    const u32 R2own = e4_f2(pre[0]);                       // internal level-2 nodes before the own group (valid on owner lanes)
    const u32 R2p = e4_f2(shfl(pre[0], lane & ~3));        // the same, seen by every lane of the group
    const u32 R3 = e4_f3(pre[0]), R4 = e4_f4(pre[0]), R5 = e4_f5(pre[0]);
    // log "equal" flags of nodes that are neither uniform nor internal
    const u32 eqb5 = ~u5 & eq5 & 0xffffu, eqb4 = ~u4 & eq4 & 0xfu;
    const bool eqb3 = !u3 && eq3, eqb2 = !u2 && eq2;

    if (emit) {
      // ================= phase A: every bitmap bit (atomicOr of runs; no byte store is in flight) =================
      {  // quads and leaves
        u64 accL = 0; int nL = 0;
        u32 accM = 0; int nM = 0;
        u32 accE = 0; int nE = 0;
        if (ai5 == 0xffffu) {  // every quad is internal: the runs are the masks themselves
          accL = W.ml[0]; nL = 64;
          accM = W.mq[0] >> 16; nM = 16;
        } else
#pragma unroll 1
        for (int bit = 15; bit >= 0; bit--) {  // quad q = 15 - bit
          if ((ai5 >> bit) & 1u) {
            accL = (accL << 4) | ((W.ml[0] >> (4 * bit)) & 0xfull); nL += 4;
            accM = (accM << 1) | ((W.mq[0] >> (16 + bit)) & 1u); nM += 1;
          } else if ((X5 >> bit) & 1u) {
            accE = (accE << 1) | ((eqb5 >> bit) & 1u); nE += 1;
          }
        }
        if (nL) e4_or_run(xw[0], Pn6 + 4u * R5, accL << (64 - nL));
        e4_or_bits(nw[0], Mn5 + R5, accM, nM);
        if (!as_snapshot) e4_or_bits(eq_words, (Pn5 - Mn5) + 4u * R4 - R5, accE, nE);
        u32 accN = 0, accC = 0; int nN = 0;
#pragma unroll
        for (int a = 0; a < 4; a++) {
          if ((ai4 >> (3 - a)) & 1u) {
            accN = (accN << 4) | ((W.in5 >> (12 - 4 * a)) & 0xfu);
            accC = (accC << 4) | ((W.mq[0] >> (12 - 4 * a)) & 0xfu);
            nN += 4;
          }
        }
        e4_or_bits(nm_words, Pn5 + 4u * R4, accN, nN);
        e4_or_bits(xw[0], Pn5 + 4u * R4, accC, nN);
      }
      if (x4) {  // the four level-4 nodes
        e4_or_bits(nm_words, Pn4 + 4u * R3, W.in4, 4);
        e4_or_bits(xw[0], Pn4 + 4u * R3, W.mu[0] & 0xfu, 4);
        u32 accM = 0, accE = 0; int nM = 0, nE = 0;
#pragma unroll
        for (int a = 0; a < 4; a++) {
          if ((W.in4 >> (3 - a)) & 1u) { accM = (accM << 1) | ((W.mu[0] >> (7 - a)) & 1u); nM++; }
          else { accE = (accE << 1) | ((eqb4 >> (3 - a)) & 1u); nE++; }
        }
        e4_or_bits(nw[0], Mn4 + R4, accM, nM);
        if (!as_snapshot) e4_or_bits(eq_words, (Pn4 - Mn4) + 4u * R3 - R4, accE, nE);
      }
      if (x3) {  // own level-3 node
        const u32 pos = Pn3 + 4u * R2p + (u32)(tid & 3);
        if (W.in3) {
          e4_set_bit(nm_words, pos);
          if ((W.mu[0] >> 9) & 1u) e4_set_bit(nw[0], Mn3 + R3);
        } else if (!as_snapshot && eqb3) {
          e4_set_bit(eq_words, pos - (Mn3 + R3));
        }
        if ((W.mu[0] >> 8) & 1u) e4_set_bit(xw[0], pos);
      }
      if (x2 && owner2) {  // level-2 node of the group
        const u32 pos = Pn2 + 4u * R1 + (u32)((tid >> 2) & 3);
        if (W.in2) {
          e4_set_bit(nm_words, pos);
          if ((W.mu[0] >> 11) & 1u) e4_set_bit(nw[0], Mn2 + R2own);
        } else if (!as_snapshot && eqb2) {
          e4_set_bit(eq_words, pos - (Mn2 + R2own));
        }
        if ((W.mu[0] >> 10) & 1u) e4_set_bit(xw[0], pos);
      }
      if (any_hi) {
        // continuation bits of DAC levels 1 and 2 (entries of three and four bytes): rare, one bit at a time
#pragma unroll
        for (int j = 1; j < 3; j++) {
          // leaves
          const u64 lm = W.ml[j - 1] & X6;  // population of DAC level j on this thread's leaves
          u64 mk = W.ml[j] & X6;
          const u32 b6 = e4_base_x(T, pre, j, 6);
          while (mk) {
            const int b = 63 - __clzll((long long)mk);
            mk &= ~(1ull << b);
            e4_set_bit(xw[j], b6 + (u32)__popcll((lm >> b) >> 1));
          }
          {  // quads: max entries of existing quads, min entries of internal quads
            const u32 pm = W.mq[j - 1] & 0xffffu & X5, pn = (W.mq[j - 1] >> 16) & ai5;
            u32 mm = W.mq[j] & 0xffffu & X5, mn = (W.mq[j] >> 16) & ai5;
            const u32 bx = e4_base_x(T, pre, j, 5), bn = e4_base_n(T, pre, j, 5);
            while (mm) { const int b = 31 - __clz((int)mm); mm &= ~(1u << b); e4_set_bit(xw[j], bx + (u32)__popc((pm >> b) >> 1)); }
            while (mn) { const int b = 31 - __clz((int)mn); mn &= ~(1u << b); e4_set_bit(nw[j], bn + (u32)__popc((pn >> b) >> 1)); }
          }
          {  // level 4
            const u32 pm = x4 ? (W.mu[j - 1] & 0xfu) : 0u, pn = (W.mu[j - 1] >> 4) & ai4;
            u32 mm = x4 ? (W.mu[j] & 0xfu) : 0u, mn = (W.mu[j] >> 4) & ai4;
            const u32 bx = e4_base_x(T, pre, j, 4), bn = e4_base_n(T, pre, j, 4);
            while (mm) { const int b = 31 - __clz((int)mm); mm &= ~(1u << b); e4_set_bit(xw[j], bx + (u32)__popc((pm >> b) >> 1)); }
            while (mn) { const int b = 31 - __clz((int)mn); mn &= ~(1u << b); e4_set_bit(nw[j], bn + (u32)__popc((pn >> b) >> 1)); }
          }
          {
            const u32 bx3 = e4_base_x(T, pre, j, 3), bn3 = e4_base_n(T, pre, j, 3);
            if (x3 && ((W.mu[j] >> 8) & 1u)) e4_set_bit(xw[j], bx3);
            if (x3 && W.in3 && ((W.mu[j] >> 9) & 1u)) e4_set_bit(nw[j], bn3);
            const u32 bx2 = e4_base_x(T, pre, j, 2), bn2 = e4_base_n(T, pre, j, 2);
            if (x2 && owner2 && ((W.mu[j] >> 10) & 1u)) e4_set_bit(xw[j], bx2);
            if (x2 && owner2 && W.in2 && ((W.mu[j] >> 11) & 1u)) e4_set_bit(nw[j], bn2);
          }
        }
      }
      if (tid == 0) {
        // root and level-1 nodes: nodemap, equal, continuation bits
        if (in0) e4_set_bit(nm_words, 0);
        else if (!as_snapshot && !u0 && eq0) e4_set_bit(eq_words, 0);
        u32 px[4] = {0, 0, 0, 0}, pn[4] = {0, 0, 0, 0};  // running positions in DAC levels 0..3
        e4_top_bits(xw[0], xw[1], xw[2], px, e0x);
        if (in0) {
          e4_top_bits(nw[0], nw[1], nw[2], pn, e0n);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const bool ik = (in1m >> (3 - k)) & 1u;
            if (ik) e4_set_bit(nm_words, 1u + (u32)k);
            else if (!as_snapshot && !e4_unif<FULL>(n1max[k], n1min[k]) && n1eq[k]) e4_set_bit(eq_words, (u32)k - (u32)__popc(in1m >> (4 - k)));
            e4_top_bits(xw[0], xw[1], xw[2], px, e1x[k]);
            if (ik) e4_top_bits(nw[0], nw[1], nw[2], pn, e1n[k]);
          }
        }
      }
    }
    __syncthreads();  // B2a: all word-wide atomics done before the first byte store

    if (emit) {
      // ================= phase B: DAC bytes =================
      u32 rx1, rx2, rx3, rn1, rn2, rn3;
      // ---- leaves and quads: one pass over the alive level-4 nodes, their four quads side by side (independent
      // loads and arithmetic; only the rare multi-byte entries keep running positions)
      {
        u32 lx1 = e4_base_x(T, pre, 1, 6), lx2 = any_hi ? e4_base_x(T, pre, 2, 6) : 0u, lx3 = any_hi ? e4_base_x(T, pre, 3, 6) : 0u;
        rx1 = e4_base_x(T, pre, 1, 5); rx2 = any_hi ? e4_base_x(T, pre, 2, 5) : 0u; rx3 = any_hi ? e4_base_x(T, pre, 3, 5) : 0u;
        rn1 = e4_base_n(T, pre, 1, 5); rn2 = any_hi ? e4_base_n(T, pre, 2, 5) : 0u; rn3 = any_hi ? e4_base_n(T, pre, 3, 5) : 0u;
        u8* dst6 = xb[0] + Pn6 + 4u * R5;
        u8* dst5 = xb[0] + Pn5 + 4u * R4;
        u8* dmn = nb[0] + Mn5 + R5;
#pragma unroll 1
        for (int a = 0; a < 4; a++) {
          if (!((ai4 >> (3 - a)) & 1u)) continue;
          const int2 n4 = S.l4[cur][a][tid];
          const u32 in5a = (W.in5 >> (12 - 4 * a)) & 0xfu;
          const u32 ml16 = (u32)(W.ml[0] >> (48 - 16 * a)) & 0xffffu;
          u32 zx[4], zn[4], z[4][4];
#pragma unroll
          for (int b = 0; b < 4; b++) {
            const int4 t = S.cell[cur][4 * a + b][tid];
            int qmax, qmin;
            e4_qmm<FULL>(t, qmax, qmin);
            if (as_snapshot) {
              zx[b] = zigzag32(e4_sub(n4.x, e4_o0<FULL>(qmax)));
              zn[b] = zigzag32(e4_sub(qmin, n4.y));
              z[b][0] = zigzag32(e4_sub(qmax, e4_o0<FULL>(t.x))); z[b][1] = zigzag32(e4_sub(qmax, e4_o0<FULL>(t.y)));
              z[b][2] = zigzag32(e4_sub(qmax, e4_o0<FULL>(t.z))); z[b][3] = zigzag32(e4_sub(qmax, e4_o0<FULL>(t.w)));
            } else {
              const int4 sq = S.cell[ref][4 * a + b][tid];
              int sqmax, sqmin;
              e4_qmm<FULL>(sq, sqmax, sqmin);
              zx[b] = zigzag32(e4_sub(e4_o0<FULL>(qmax), e4_o0<FULL>(sqmax)));
              zn[b] = zigzag32(e4_sub(qmin, sqmin));
              z[b][0] = zigzag32(e4_sub(t.x, sq.x)); z[b][1] = zigzag32(e4_sub(t.y, sq.y));
              z[b][2] = zigzag32(e4_sub(t.z, sq.z)); z[b][3] = zigzag32(e4_sub(t.w, sq.w));
            }
          }
          e4_store4(dst5, zx[0], zx[1], zx[2], zx[3]);
          dst5 += 4;
#pragma unroll
          for (int b = 0; b < 4; b++)
            if ((in5a >> (3 - b)) & 1u) e4_store4(dst6 + 4u * (u32)__popc(in5a >> (4 - b)), z[b][0], z[b][1], z[b][2], z[b][3]);
          dst6 += 4u * (u32)__popc(in5a);
          if (ml16) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
              if (((in5a >> (3 - b)) & 1u) && ((ml16 >> (12 - 4 * b)) & 0xfu)) {
#pragma unroll
                for (int c = 0; c < 4; c++)
                  if (z[b][c] > 0xffu) e4_emit_hi(xb[1], xb[2], xb[3], z[b][c], lx1, lx2, lx3);
              }
            }
          }
#pragma unroll
          for (int b = 0; b < 4; b++)
            if (zx[b] > 0xffu) e4_emit_hi(xb[1], xb[2], xb[3], zx[b], rx1, rx2, rx3);
#pragma unroll
          for (int b = 0; b < 4; b++) {
            if ((in5a >> (3 - b)) & 1u) {
              *dmn++ = (u8)zn[b];
              if (zn[b] > 0xffu) e4_emit_hi(nb[1], nb[2], nb[3], zn[b], rn1, rn2, rn3);
            }
          }
        }
      }
      if (x4) {  // ---- level-4 nodes
        rx1 = e4_base_x(T, pre, 1, 4); rx2 = any_hi ? e4_base_x(T, pre, 2, 4) : 0u; rx3 = any_hi ? e4_base_x(T, pre, 3, 4) : 0u;
        rn1 = e4_base_n(T, pre, 1, 4); rn2 = any_hi ? e4_base_n(T, pre, 2, 4) : 0u; rn3 = any_hi ? e4_base_n(T, pre, 3, 4) : 0u;
        u32 zx[4], zn[4];
#pragma unroll
        for (int a = 0; a < 4; a++) {
          const int2 n4 = S.l4[cur][a][tid], s4 = S.l4[ref][a][tid];
          zx[a] = zigzag32(as_snapshot ? e4_sub(t3max, e4_o0<FULL>(n4.x)) : e4_sub(e4_o0<FULL>(n4.x), e4_o0<FULL>(s4.x)));
          zn[a] = zigzag32(as_snapshot ? e4_sub(n4.y, t3min) : e4_sub(n4.y, s4.y));
        }
        e4_store4(xb[0] + Pn4 + 4u * R3, zx[0], zx[1], zx[2], zx[3]);
        u8* dmn = nb[0] + Mn4 + R4;
#pragma unroll
        for (int a = 0; a < 4; a++)
          if (zx[a] > 0xffu) e4_emit_hi(xb[1], xb[2], xb[3], zx[a], rx1, rx2, rx3);
#pragma unroll
        for (int a = 0; a < 4; a++) {
          if ((W.in4 >> (3 - a)) & 1u) {
            *dmn++ = (u8)zn[a];
            if (zn[a] > 0xffu) e4_emit_hi(nb[1], nb[2], nb[3], zn[a], rn1, rn2, rn3);
          }
        }
      }
      {
        const u32 bx1 = e4_base_x(T, pre, 1, 3), bx2 = any_hi ? e4_base_x(T, pre, 2, 3) : 0u, bx3 = any_hi ? e4_base_x(T, pre, 3, 3) : 0u;
        const u32 bn1 = e4_base_n(T, pre, 1, 3), bn2 = any_hi ? e4_base_n(T, pre, 2, 3) : 0u, bn3 = any_hi ? e4_base_n(T, pre, 3, 3) : 0u;
        if (x3) {
          rx1 = bx1; rx2 = bx2; rx3 = bx3; rn1 = bn1; rn2 = bn2; rn3 = bn3;
          const u32 zx = zigzag32(as_snapshot ? e4_sub(t2max, e4_o0<FULL>(t3max)) : e4_sub(e4_o0<FULL>(t3max), e4_o0<FULL>(s3max)));
          xb[0][Pn3 + 4u * R2p + (u32)(tid & 3)] = (u8)zx;
          if (zx > 0xffu) e4_emit_hi(xb[1], xb[2], xb[3], zx, rx1, rx2, rx3);
          if (W.in3) {
            const u32 zn = zigzag32(as_snapshot ? e4_sub(t3min, t2min) : e4_sub(t3min, s3min));
            nb[0][Mn3 + R3] = (u8)zn;
            if (zn > 0xffu) e4_emit_hi(nb[1], nb[2], nb[3], zn, rn1, rn2, rn3);
          }
        }
      }
      {
        const u32 bx1 = e4_base_x(T, pre, 1, 2), bx2 = any_hi ? e4_base_x(T, pre, 2, 2) : 0u, bx3 = any_hi ? e4_base_x(T, pre, 3, 2) : 0u;
        const u32 bn1 = e4_base_n(T, pre, 1, 2), bn2 = any_hi ? e4_base_n(T, pre, 2, 2) : 0u, bn3 = any_hi ? e4_base_n(T, pre, 3, 2) : 0u;
        if (x2 && owner2) {
          rx1 = bx1; rx2 = bx2; rx3 = bx3; rn1 = bn1; rn2 = bn2; rn3 = bn3;
          const u32 zx = zigzag32(as_snapshot ? e4_sub(t1max, e4_o0<FULL>(t2max)) : e4_sub(e4_o0<FULL>(t2max), e4_o0<FULL>(s2max)));
          xb[0][Pn2 + 4u * R1 + (u32)((tid >> 2) & 3)] = (u8)zx;
          if (zx > 0xffu) e4_emit_hi(xb[1], xb[2], xb[3], zx, rx1, rx2, rx3);
          if (W.in2) {
            const u32 zn = zigzag32(as_snapshot ? e4_sub(t2min, t1min) : e4_sub(t2min, s2min));
            nb[0][Mn2 + R2own] = (u8)zn;
            if (zn > 0xffu) e4_emit_hi(nb[1], nb[2], nb[3], zn, rn1, rn2, rn3);
          }
        }
      }
      if (tid == 0) {
        // root, level-1 entries, every header field
        rx1 = rx2 = rx3 = rn1 = rn2 = rn3 = 0;
        {
          const u32 z = zigzag32(e0x);
          xb[0][0] = (u8)z;
          if (z > 0xffu) e4_emit_hi(xb[1], xb[2], xb[3], z, rx1, rx2, rx3);
        }
        if (in0) {
          const u32 zr = zigzag32(e0n);
          nb[0][0] = (u8)zr;
          if (zr > 0xffu) e4_emit_hi(nb[1], nb[2], nb[3], zr, rn1, rn2, rn3);
          u32 rm = 1;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const u32 z = zigzag32(e1x[k]);
            xb[0][1 + k] = (u8)z;
            if (z > 0xffu) e4_emit_hi(xb[1], xb[2], xb[3], z, rx1, rx2, rx3);
            if ((in1m >> (3 - k)) & 1u) {
              const u32 zn = zigzag32(e1n[k]);
              nb[0][rm++] = (u8)zn;
              if (zn > 0xffu) e4_emit_hi(nb[1], nb[2], nb[3], zn, rn1, rn2, rn3);
            }
          }
        }
        out[0] = 2;  // k
        store_be32(out + 1, (u32)unit.rows);
        store_be32(out + 5, (u32)unit.cols);
        store_be32(out + 9, 64u);  // sidelen
        u32 nb_ = 0;
        auto bitmap_hdr = [&](u32 hdr, u32 len) {
          store_be32(out + hdr, len);
          store_be32(out + hdr + 4, 4u);  // k = 4 (bitmap.rs:69)
          S.bm_off[nb_] = hdr; S.bm_len[nb_] = len; nb_++;
        };
        bitmap_hdr(nm_hdr, nm_len);
        if (!as_snapshot) bitmap_hdr(eq_hdr, eq_len);
        out[DX.hdr[0] - 1] = (u8)DX.levels;
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (j < DX.levels) bitmap_hdr(DX.hdr[j], T.cmax[j]);
        out[DN.hdr[0] - 1] = (u8)DN.levels;
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (j < DN.levels) bitmap_hdr(DN.hdr[j], T.cmin[j]);
        S.n_bm = nb_;
      }
    }
    __syncthreads();  // B2

    if (emit) {
      // ================= rank directories: index[b] = ones in bits [0, 128(b+1))  (bitmap.rs:97-104) =================
      const u32 n_bm = S.n_bm;
      for (u32 i = warp; i < n_bm; i += 2) {
        const u32 hdr = S.bm_off[i], len = S.bm_len[i];
        const u32 blocks = len >> 7;
        u8* const index = out + hdr + 8u;
        const u8* const words = index + 4u * blocks;
        u32 carry = 0;
        for (u32 b0 = 0; b0 < blocks; b0 += 32u) {
          const u32 b = b0 + lane;
          u32 c = b < blocks ? e4_popc_bytes16(words + 16u * b) : 0u;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const u32 n = __shfl_up_sync(0xffffffffu, c, d);
            if (lane >= d) c += n;
          }
          if (b < blocks) e4_store_be32(index + 4u * b, carry + c);
          carry += shfl(c, 31);
        }
      }
    }
    __syncthreads();  // B3

    // ================= copy-out (and re-zero the image) =================
    if (staged) {
      piece_off = S.piece_off;
      fits = piece_off + (((u64)my_size + 15ull) & ~15ull) <= P.arena_cap;
      if (!fits) err |= EF_ARENA_FULL;
      uint4* src = reinterpret_cast<uint4*>(S.pool);
      uint4* dst = reinterpret_cast<uint4*>(P.arena + piece_off);
      for (u32 i = tid; i < (my_size + 15u) / 16u; i += E4_THREADS) {
        const uint4 v = src[i];
        if (fits) dst[i] = v;
        src[i] = make_uint4(0, 0, 0, 0);
      }
    }

    // ---------------- bookkeeping: start a new block or extend the current one
    if (as_snapshot) {
      ref = cur;  // the instant's image becomes the block's snapshot; the old snapshot image is overwritten next
      cur ^= 1;
      s3max = t3max; s3min = t3min; s2max = t2max; s2min = t2min; s1max = t1max; s1min = t1min; s0max = t0max; s0min = t0min;
      n_snap++;
      n_logs = 0;
      total_bytes += 1;  // Block's n_instants byte (block.rs:88-95)
    } else {
      n_logs++;
      n_log_total++;
    }
    total_bytes += my_size;
  }

  if (tid == 0) {
    UnitResult r;
    r.bytes = total_bytes + 6;  // encoding + fractional_bits + n_blocks (chunk.rs:235-243)
    r.snapshots = n_snap;
    r.logs = n_log_total;
    P.results[unit_idx] = r;
  }
  err = __reduce_or_sync(0xffffffffu, err);
  if (lane == 0 && err) atomicOr(P.err, err);
}

}  // namespace dcdf
