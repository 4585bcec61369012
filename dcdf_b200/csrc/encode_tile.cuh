// encode_tile.cuh -- the Chunk::build kernel for tiles of side <= 64 (one CTA per Chunk).
//
// Replaces, per (subchunk x time slice) unit, the reference's
//   Chunk::build            chunk.rs:42-96      (heuristic: Snapshot vs Log per instant)
//   Snapshot::build         snapshot.rs:108-156 + K2TreeNode::_build  snapshot.rs:439-500
//   Log::build              log.rs:112-165      + K2PTreeNode::_build log.rs:725-817
//   BitMapBuilder::finish   bitmap.rs:66-113,   Dac::from dac.rs:96-132
//   and the Snapshot / Log / Dac / BitMap serializers (snapshot.rs:48-58, log.rs:53-64, dac.rs:37-44,
//   bitmap.rs:128-138).
//
// Layout of the work inside the CTA (256 threads, k = 2, tree of L <= 6 levels embedded in a 64x64
// frame; the tree root is the frame node at level lo = 6 - L, Morton index 0):
//   * thread t owns the level-4 frame node with Morton index t: a 4x4 block of cells kept in
//     registers in Morton order (cell m = quad (m>>2), child (m&3); row bit above column bit).
//   * levels 6 -> 5 -> 4 of the min/max/equal pyramid are reduced inside the thread, levels 3 and 2 with
//     warp shuffles, levels 1 and 0 through shared memory.
//   * node existence flows top-down (a node exists iff every ancestor is an internal node);
//     BFS output positions come from per-level ballot/popc ranks plus a cross-warp scan.
//   * DAC byte-length histograms of BOTH candidate encodings give their serialized sizes; only the
//     winner is emitted: values are staged in shared memory in BFS order and packed level by level
//     (byte, continuation bitmap, rank directory) with in-place compaction.
//   * each emitted Snapshot/Log is a "piece" in a bump-allocated arena; gather.cuh lays pieces out as
//     the final Chunk bytes.
#pragma once
#include "common.cuh"

namespace dcdf {

constexpr int ENC_THREADS = 256;
constexpr int ENC_WARPS = ENC_THREADS / 32;
constexpr int MAX_NODES = 5461;     // nodes of a 7-level (64x64) quadtree
constexpr int MAX_INTERNAL = 1365;  // nodes above the leaf level
constexpr int NM_WORDS = (MAX_INTERNAL + 31) / 32 + 1;
constexpr int STAGE_BYTES = 16 * 1024;  // structures up to this size are assembled in smem, then bulk-copied

struct EncUnit {
  i64 base;         // element offset of [t0, top, left] in the input array
  int rows, cols;   // in-bounds extent of the tile, 1..64
  int instants;     // 1..  (Block caps logs at 254, any length is fine)
  int bits;         // fractional bits written to the chunk header / used by to_fixed
  int flags;        // UF_*
  int lo;           // 6 - tree levels
  u32 piece_base;   // first entry of this unit in the piece table
  u32 slot;         // caller-defined (subchunk slot inside its slice)
  int row0, col0;   // origin of the tile inside its region (row-major order of the NaN quirk)
};
enum : int { UF_ROUND = 1, UF_NARROW = 2, UF_SKIP = 4 };

struct Piece {
  u64 off;   // byte offset in the arena (16-byte aligned)
  u32 size;  // serialized size of the Snapshot / Log
  u32 kind;  // 1 = Snapshot (starts a Block), 0 = Log
};

struct UnitResult {
  u64 bytes;       // Chunk::size()  (chunk.rs:269-278)
  u32 snapshots;   // == number of blocks
  u32 logs;
};

struct EncParams {
  const void* data;
  i64 stride_t, stride_r, stride_c;  // element strides
  const EncUnit* units;
  const u32* order;       // unit indices to process (narrow or wide list)
  const u32* order_count; // number of valid entries in `order` (device scalar)
  Piece* pieces;
  UnitResult* results;
  u8* arena;
  u64 arena_cap;
  unsigned long long* arena_head;
  u32* err;               // EF_* bits
};

template <typename V>
struct VT;
template <>
struct VT<int32_t> {
  typedef u32 U;
  static constexpr int32_t NONE_MAX = INT32_MIN, NONE_MIN = INT32_MAX;
  static constexpr int MAXLEN = 4;
  static DCDF_DEVINL u32 zz(int32_t v) { return zigzag32(v); }
};
template <>
struct VT<i64> {
  typedef u64 U;
  static constexpr i64 NONE_MAX = INT64_MIN, NONE_MIN = INT64_MAX;
  static constexpr int MAXLEN = 8;
  static DCDF_DEVINL u64 zz(i64 v) { return zigzag64(v); }
};

template <typename V>
DCDF_DEVINL V vmax(V a, V b) { return a > b ? a : b; }
template <typename V>
DCDF_DEVINL V vmin(V a, V b) { return a < b ? a : b; }
template <typename V>
DCDF_DEVINL V shfl(V v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <typename V>
DCDF_DEVINL V shfl_xor(V v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// Upper-level node record shared through smem (levels 2, 1, 0 of the 64x64 frame).
template <typename V>
struct UpNode {
  V tmax, tmin;  // sentinels when None
  V diff;        // log: t - s of the first cell (log.rs:751,781)
  int eq;        // log: K2PTreeNode.equal
};

// Shared-memory plan of the encode CTA.
template <typename V>
struct EncSmem {
  typename VT<V>::U A[MAX_NODES + 3];      // zigzag codes of `max` in BFS order
  typename VT<V>::U B[MAX_INTERNAL + 3];   // zigzag codes of `min`
  u32 nm[NM_WORDS];                        // nodemap bits, MSB first
  u32 eqw[NM_WORDS];                       // equal bits (logs)
  UpNode<V> l2[16];
  UpNode<V> l1[4];
  UpNode<V> l0;
  V s_l1max[4], s_l1min[4], s_l0max, s_l0min;  // cached snapshot upper levels (per-thread regs hold l3,l2)
  u32 cnt[24];                             // block-reduced counters
  u32 wtot[6][ENC_WARPS];                  // per-warp internal-node counts per level (scan input)
  u32 scan[ENC_WARPS];                     // scratch for the DAC compaction scan
  u32 scan2[ENC_WARPS];
  int decision;                            // 1 = snapshot
  u64 piece_off;
  __align__(16) u8 stage[STAGE_BYTES];
};

// ---------------------------------------------------------------------------------------------------
// DAC emission (dac.rs:96-132 + dac.rs:37-44 + bitmap.rs:66-113,128-138), block-wide.
// `a` holds n zigzag codes (shared memory), n_j = cntgt[j] = number of codes longer than j bytes.
// Level j is dense: byte = code & 0xff, continuation bit = (code >> 8) != 0; survivors are compacted in
// place (order preserved) to form level j+1.  Returns bytes written.  All threads must call.
template <typename U, int MAXLEN>
__device__ u32 block_dac_emit(U* a, u32 n, const u32* cntgt, u8* out, u32* scan_a, u32* scan_b) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int n_levels = 0;
  for (int j = 0; j < MAXLEN; j++)
    if (cntgt[j] > 0) n_levels = j + 1;
  if (tid == 0) out[0] = (u8)n_levels;
  u32 off = 1;
  for (int j = 0; j < n_levels; j++) {
    const u32 nj = cntgt[j];
    const u32 blocks = nj / 128u, words = (nj + 31u) / 32u;
    u8* p_len = out + off;
    u8* p_index = p_len + 8;
    u8* p_words = p_index + 4 * blocks;
    u8* p_bytes = p_words + 4 * words;
    if (tid == 0) {
      store_be32(p_len, nj);
      store_be32(p_len + 4, 4u);  // k = 4 (bitmap.rs:69)
    }
    // chunks of 1024 entries: warp w owns [128w, 128w+128) of the chunk, 4 ballots of 32
    u32 running = 0;  // survivors before this chunk
    for (u32 c0 = 0; c0 < nj; c0 += 1024u) {
      U v[4];
      u32 bal[4];
      u32 wcount = 0;
#pragma unroll
      for (int s = 0; s < 4; s++) {
        u32 i = c0 + warp * 128u + s * 32u + lane;
        v[s] = i < nj ? a[i] : (U)0;
        bool more = i < nj && (v[s] >> 8) != 0;
        bal[s] = __ballot_sync(0xffffffffu, more);
        wcount += __popc(bal[s]);
      }
      u32* sc = ((c0 >> 10) & 1) ? scan_b : scan_a;
      if (lane == 0) sc[warp] = wcount;
      __syncthreads();  // also: every read of this chunk happened before any compacting write
      u32 wbase = running, total = 0;
#pragma unroll
      for (int w = 0; w < ENC_WARPS; w++) {
        u32 x = sc[w];
        if (w < warp) wbase += x;
        total += x;
      }
      u32 pre = wbase;
#pragma unroll
      for (int s = 0; s < 4; s++) {
        u32 i0 = c0 + warp * 128u + s * 32u;  // first entry of this ballot
        u32 i = i0 + lane;
        if (i < nj) {
          p_bytes[i] = (u8)(v[s] & 0xff);
          if ((bal[s] >> lane) & 1u) a[pre + __popc(bal[s] & lanemask_lt())] = (U)(v[s] >> 8);
        }
        if (lane == 0 && i0 < nj) store_be32(p_words + 4 * (i0 >> 5), __brev(bal[s]));
        // rank directory: index[b] = ones in bits [0, 128(b+1))  (bitmap.rs:97-104)
        if (lane == 0 && s == 3 && (i0 + 32u) <= nj) store_be32(p_index + 4 * ((i0 + 32u) / 128u - 1u), pre + __popc(bal[s]));
        pre += __popc(bal[s]);
      }
      running += total;
    }
    __syncthreads();  // compacted level j+1 complete before it is read
    off += 8 + 4 * blocks + 4 * words + nj;
  }
  return off;
}

__host__ __device__ inline u32 dac_size_from_counts(const u32* cntgt, int maxlen) {
  u32 s = 1;
  for (int j = 0; j < maxlen; j++)
    if (cntgt[j] > 0) s += bitmap_size(cntgt[j]) + cntgt[j];
  return s;
}

// BitMap serialization from MSB-first words held in smem (bitmap.rs:128-138); block-wide.
__device__ inline u32 block_bitmap_emit(const u32* words_smem, u32 length, u8* out) {
  const int tid = threadIdx.x;
  const u32 blocks = length / 128u, words = (length + 31u) / 32u;
  if (tid == 0) {
    store_be32(out, length);
    store_be32(out + 4, 4u);
  }
  for (u32 b = tid; b < blocks; b += ENC_THREADS) {
    u32 c = 0;
    for (u32 w = 0; w < 4 * (b + 1); w++) c += __popc(words_smem[w]);
    store_be32(out + 8 + 4 * b, c);
  }
  for (u32 w = tid; w < words; w += ENC_THREADS) store_be32(out + 8 + 4 * blocks + 4 * w, words_smem[w]);
  return 8 + 4 * blocks + 4 * words;
}

// ---------------------------------------------------------------------------------------------------
// Tile loader: thread t reads its 4x4 block (Morton order) of instant `p`.
template <typename InT>
DCDF_DEVINL void load_block(const InT* p, i64 sr, i64 sc, int rows, int cols, int r0, int c0, InT out[16], u32& inb) {
  inb = 0;
  const bool full = (r0 + 4 <= rows) && (c0 + 4 <= cols);
  if (full && sc == 1 && sizeof(InT) == 4 && ((((uintptr_t)(p + (i64)r0 * sr + c0)) | ((uintptr_t)(sr * (i64)sizeof(InT)))) & 15) == 0) {
    inb = 0xffffu;
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(p + (i64)(r0 + rr) * sr + c0));
      const u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {
        const int m = ((rr >> 1) << 3) | ((cc >> 1) << 2) | ((rr & 1) << 1) | (cc & 1);
        out[m] = *reinterpret_cast<const InT*>(&w[cc]);
      }
    }
    return;
  }
#pragma unroll
  for (int rr = 0; rr < 4; rr++) {
#pragma unroll
    for (int cc = 0; cc < 4; cc++) {
      const int m = ((rr >> 1) << 3) | ((cc >> 1) << 2) | ((rr & 1) << 1) | (cc & 1);
      const bool in = (r0 + rr < rows) && (c0 + cc < cols);
      out[m] = in ? __ldg(p + (i64)(r0 + rr) * sr + (i64)(c0 + cc) * sc) : InT(0);
      inb |= in ? (1u << m) : 0u;
    }
  }
}

// Packed per-thread histogram: field j (8 bits) counts entries longer than j bytes.
template <typename V>
struct Hist;
template <>
struct Hist<int32_t> {
  u32 h = 0;
  DCDF_DEVINL void add(u32 zz) { h += 0x01010101u >> (32 - 8 * dac_len(zz)); }
  DCDF_DEVINL u32 field(int j) const { return (h >> (8 * j)) & 0xffu; }
};
template <>
struct Hist<i64> {
  u64 h = 0;
  DCDF_DEVINL void add(u64 zz) { h += 0x0101010101010101ull >> (64 - 8 * dac_len(zz)); }
  DCDF_DEVINL u32 field(int j) const { return (u32)(h >> (8 * j)) & 0xffu; }
};

// Block-reduce a histogram into smem counters cnt[base .. base+MAXLEN) (two 16-bit fields per redux).
template <typename V>
DCDF_DEVINL void reduce_hist(const Hist<V>& h, u32* cnt, int base) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < VT<V>::MAXLEN; j += 2) {
    u32 packed = h.field(j) | (h.field(j + 1) << 16);
    u32 r = __reduce_add_sync(0xffffffffu, packed);
    if (lane == 0 && r) {
      if (r & 0xffffu) atomicAdd(&cnt[base + j], r & 0xffffu);
      if (r >> 16) atomicAdd(&cnt[base + j + 1], r >> 16);
    }
  }
}

// Counter slots in EncSmem::cnt
enum { C_SMAX = 0, C_SMIN = 8, C_LMAX = 16 /* ..23 */ };
// log min histogram + lengths live in a second array to keep indices simple
enum { C2_LMIN = 0, C2_SNM = 8, C2_LNM = 9, C2_N = 10 };

// ---------------------------------------------------------------------------------------------------
template <typename InT, typename V>
__global__ void __launch_bounds__(ENC_THREADS) k_encode_tiles(const EncParams P) {
  typedef typename VT<V>::U U;
  constexpr int MAXLEN = VT<V>::MAXLEN;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EncSmem<V>& S = *reinterpret_cast<EncSmem<V>*>(smem_raw);
  __shared__ u32 cnt2[C2_N];

  if (blockIdx.x >= *P.order_count) return;
  const u32 unit_idx = P.order[blockIdx.x];
  const EncUnit unit = P.units[unit_idx];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lo = unit.lo;
  const bool do_round = unit.flags & UF_ROUND;
  const int r0 = 4 * (int)morton_row(tid), c0 = 4 * (int)morton_col(tid);
  const InT* base = static_cast<const InT*>(P.data) + unit.base;

  // tree membership of this thread's frame nodes (tree root = frame node (lo, 0))
  const bool intree4 = lo <= 4 ? ((tid >> (8 - 2 * lo)) == 0) : (tid == 0);
  const u32 quad_mask = lo <= 4 ? 0xfu : 0x1u;  // lo == 5: only quad 0 of thread 0 is in the tree

  u32 err = 0;
  // cached snapshot (reference instant of the current block): leaves + own ancestors
  V sv[16];
#pragma unroll
  for (int m = 0; m < 16; m++) sv[m] = 0;
  V s3max = 0, s3min = 0, s2max = 0, s2min = 0;
  u32 n_logs = 0, n_snap = 0, n_log_total = 0;
  u64 total_bytes = 0;

  InT raw[16];
  u32 inb;
  load_block<InT>(base, P.stride_r, P.stride_c, unit.rows, unit.cols, r0, c0, raw, inb);

  for (int inst = 0; inst < unit.instants; inst++) {
    // ---------------- convert (a3) and prefetch the next instant
    V tv[16];
#pragma unroll
    for (int m = 0; m < 16; m++) {
      i64 f = Conv<InT>::get(raw[m], unit.bits, do_round, err);
      tv[m] = (V)f;
      if (sizeof(V) == 4 && ((inb >> m) & 1u) && (f > (i64)0x3fffffff || f < -(i64)0x3fffffff)) err |= EF_BAD_FORMAT;  // narrow path mis-selected
    }
    const u32 cur_inb = inb;
    if (inst + 1 < unit.instants)
      load_block<InT>(base + (i64)(inst + 1) * P.stride_t, P.stride_r, P.stride_c, unit.rows, unit.cols, r0, c0, raw, inb);
    const bool first = inst == 0;

    // ---------------- bottom-up pyramid inside the thread: levels 6 -> 5 -> 4
    V t5max[4], t5min[4];
    V d0[4];        // diff of the first cell of each quad (log)
    u32 eq5 = 0;    // bit q: all four leaf diffs of quad q equal
#pragma unroll
    for (int q = 0; q < 4; q++) {
      V mx = VT<V>::NONE_MAX, mn = VT<V>::NONE_MIN;
      V dq[4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int m = 4 * q + c;
        const bool in = (cur_inb >> m) & 1u;
        if (in) { mx = vmax(mx, tv[m]); mn = vmin(mn, tv[m]); }
        dq[c] = (in && !first) ? (V)(tv[m] - sv[m]) : (V)0;  // OOB diff = 0 (log.rs:751)
      }
      t5max[q] = mx; t5min[q] = mn;
      d0[q] = dq[0];
      if (dq[1] == dq[0] && dq[2] == dq[0] && dq[3] == dq[0]) eq5 |= 1u << q;
    }
    V t4max = vmax(vmax(t5max[0], t5max[1]), vmax(t5max[2], t5max[3]));
    V t4min = vmin(vmin(t5min[0], t5min[1]), vmin(t5min[2], t5min[3]));
    const V diff4 = d0[0];
    const bool eq4 = eq5 == 0xfu && d0[1] == d0[0] && d0[2] == d0[0] && d0[3] == d0[0];

    // ---------------- levels 3 and 2 with shuffles (4 resp. 16 consecutive lanes)
    V t3max = vmax(t4max, shfl_xor(t4max, 1)); t3max = vmax(t3max, shfl_xor(t3max, 2));
    V t3min = vmin(t4min, shfl_xor(t4min, 1)); t3min = vmin(t3min, shfl_xor(t3min, 2));
    const V diff3 = shfl(diff4, lane & ~3);
    const u32 ok3 = __ballot_sync(0xffffffffu, eq4 && diff4 == diff3);
    const bool eq3 = ((ok3 >> (lane & ~3)) & 0xfu) == 0xfu;
    V t2max = vmax(t3max, shfl_xor(t3max, 4)); t2max = vmax(t2max, shfl_xor(t2max, 8));
    V t2min = vmin(t3min, shfl_xor(t3min, 4)); t2min = vmin(t2min, shfl_xor(t2min, 8));
    const V diff2 = shfl(diff3, lane & ~15);
    const u32 ok2 = __ballot_sync(0xffffffffu, eq3 && diff3 == diff2);
    const bool eq2 = ((ok2 >> (lane & ~15)) & 0xffffu) == 0xffffu;

    // ---------------- levels 1 and 0 through smem
    if ((lane & 15) == 0) {
      UpNode<V> n;
      n.tmax = t2max; n.tmin = t2min; n.diff = diff2; n.eq = eq2;
      S.l2[tid >> 4] = n;
    }
    if (tid < 24) S.cnt[tid] = 0;
    if (tid < C2_N) cnt2[tid] = 0;
    __syncthreads();
    if (tid < 4) {
      UpNode<V> a = S.l2[4 * tid], b = S.l2[4 * tid + 1], c = S.l2[4 * tid + 2], d = S.l2[4 * tid + 3];
      UpNode<V> n;
      n.tmax = vmax(vmax(a.tmax, b.tmax), vmax(c.tmax, d.tmax));
      n.tmin = vmin(vmin(a.tmin, b.tmin), vmin(c.tmin, d.tmin));
      n.diff = a.diff;
      n.eq = a.eq && b.eq && c.eq && d.eq && b.diff == a.diff && c.diff == a.diff && d.diff == a.diff;
      S.l1[tid] = n;
    }
    __syncthreads();
    if (tid == 0) {
      UpNode<V> a = S.l1[0], b = S.l1[1], c = S.l1[2], d = S.l1[3];
      UpNode<V> n;
      n.tmax = vmax(vmax(a.tmax, b.tmax), vmax(c.tmax, d.tmax));
      n.tmin = vmin(vmin(a.tmin, b.tmin), vmin(c.tmin, d.tmin));
      n.diff = a.diff;
      n.eq = a.eq && b.eq && c.eq && d.eq && b.diff == a.diff && c.diff == a.diff && d.diff == a.diff;
      S.l0 = n;
    }
    __syncthreads();
    const UpNode<V> N1 = S.l1[tid >> 6];
    const UpNode<V> N0 = S.l0;
    // snapshot upper levels (reference instant) for this thread's ancestors
    const V s1max = S.s_l1max[tid >> 6], s1min = S.s_l1min[tid >> 6], s0max = S.s_l0max, s0min = S.s_l0min;

    // ---------------- per-node classification for this thread's chain (levels 0..4) and its quads
    // U = "uniform in t" (min_t == max_t as Options; a None node is uniform)   snapshot.rs:133, log.rs:137
    auto uniform = [](V mx, V mn) { return mx == VT<V>::NONE_MAX || mx == mn; };
    const bool u0 = uniform(N0.tmax, N0.tmin), u1 = uniform(N1.tmax, N1.tmin), u2 = uniform(t2max, t2min),
               u3 = uniform(t3max, t3min), u4 = uniform(t4max, t4min);
    u32 u5 = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) u5 |= uniform(t5max[q], t5min[q]) ? (1u << q) : 0u;
    // internal flags: snapshot = !U ; log = !U && !equal
    const bool si[5] = {!u0, !u1, !u2, !u3, !u4};
    const bool li[5] = {!u0 && !N0.eq, !u1 && !N1.eq, !u2 && !eq2, !u3 && !eq3, !u4 && !eq4};
    const u32 si5 = ~u5 & 0xfu, li5 = ~u5 & ~eq5 & 0xfu;
    // values of the chain, "or 0" for None
    auto or0 = [](V v, V none) { return v == none ? (V)0 : v; };
    const V cmax[5] = {N0.tmax, N1.tmax, t2max, t3max, t4max};
    const V cmin[5] = {N0.tmin, N1.tmin, t2min, t3min, t4min};
    const V csmax[5] = {s0max, s1max, s2max, s3max, (V)0};  // level 4/5 snapshot values recomputed from sv below
    const V csmin[5] = {s0min, s1min, s2min, s3min, (V)0};
    const bool owner[5] = {tid == 0, (tid & 63) == 0, (tid & 15) == 0, (tid & 3) == 0, true};

    // snapshot pyramid levels 5/4 of the reference instant, recomputed from the cached leaves
    V s5max[4], s5min[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      V mx = VT<V>::NONE_MAX, mn = VT<V>::NONE_MIN;
#pragma unroll
      for (int c = 0; c < 4; c++)
        if ((cur_inb >> (4 * q + c)) & 1u) { mx = vmax(mx, sv[4 * q + c]); mn = vmin(mn, sv[4 * q + c]); }
      s5max[q] = mx; s5min[q] = mn;
    }
    const V s4max = vmax(vmax(s5max[0], s5max[1]), vmax(s5max[2], s5max[3]));
    const V s4min = vmin(vmin(s5min[0], s5min[1]), vmin(s5min[2], s5min[3]));

    // ---------------- histograms of both candidates (a5/a7 entry values, a9 byte lengths)
    Hist<V> hsmax, hsmin, hlmax, hlmin;
    u32 snm = 0, lnm = 0;  // nodemap lengths contributed by this thread
    {
      bool sal = intree4 || lo == 5, lal = sal;  // alive flags entering level lo (virtual above)
      // virtual levels above the root always pass; only the Morton-0 path is in the tree
#pragma unroll
      for (int l = 0; l < 5; l++) {
        if (l < lo) continue;
        const bool mine = owner[l] && intree4;
        const V tm = or0(cmax[l], VT<V>::NONE_MAX);
        const V pmax = cmax[l > 0 ? l - 1 : 0], pmin = cmin[l > 0 ? l - 1 : 0];  // parent (unused at the root)
        if (sal && mine) {
          const V e = l == lo ? tm : (V)(pmax - tm);
          hsmax.add(VT<V>::zz(e));
          snm++;
          if (si[l]) hsmin.add(VT<V>::zz(l == lo ? cmin[l] : (V)(cmin[l] - pmin)));
        }
        if (!first && lal && mine) {
          const V smx = l == 4 ? s4max : csmax[l], smn = l == 4 ? s4min : csmin[l];
          hlmax.add(VT<V>::zz((V)(tm - or0(smx, VT<V>::NONE_MAX))));
          lnm++;
          if (li[l]) hlmin.add(VT<V>::zz((V)(cmin[l] - smn)));
        }
        sal = sal && si[l];
        lal = lal && li[l];
      }
      // level 5 quads and level 6 leaves
      const bool s4alive = lo == 5 ? (tid == 0) : sal, l4alive = lo == 5 ? (tid == 0) : lal;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (!((quad_mask >> q) & 1u)) continue;
        const V tm = or0(t5max[q], VT<V>::NONE_MAX);
        if (s4alive) {
          hsmax.add(VT<V>::zz(lo == 5 ? tm : (V)(t4max - tm)));
          snm++;
          if ((si5 >> q) & 1u) {
            hsmin.add(VT<V>::zz(lo == 5 ? t5min[q] : (V)(t5min[q] - t4min)));
#pragma unroll
            for (int c = 0; c < 4; c++) {
              const int m = 4 * q + c;
              const V leaf = ((cur_inb >> m) & 1u) ? tv[m] : (V)0;
              hsmax.add(VT<V>::zz((V)(t5max[q] - leaf)));
            }
          }
        }
        if (!first && l4alive) {
          hlmax.add(VT<V>::zz((V)(tm - or0(s5max[q], VT<V>::NONE_MAX))));
          lnm++;
          if ((li5 >> q) & 1u) {
            hlmin.add(VT<V>::zz((V)(t5min[q] - s5min[q])));
#pragma unroll
            for (int c = 0; c < 4; c++) {
              const int m = 4 * q + c;
              const bool in = (cur_inb >> m) & 1u;
              hlmax.add(VT<V>::zz(in ? (V)(tv[m] - sv[m]) : (V)0));
            }
          }
        }
      }
    }
    reduce_hist<V>(hsmax, S.cnt, C_SMAX);
    reduce_hist<V>(hsmin, S.cnt, C_SMIN);
    if (!first) {
      reduce_hist<V>(hlmax, S.cnt, C_LMAX);
      reduce_hist<V>(hlmin, cnt2, C2_LMIN);
    }
    {
      u32 r = __reduce_add_sync(0xffffffffu, snm | (lnm << 16));
      if (lane == 0) {
        atomicAdd(&cnt2[C2_SNM], r & 0xffffu);
        atomicAdd(&cnt2[C2_LNM], r >> 16);
      }
    }
    __syncthreads();

    // ---------------- sizes and the heuristic (chunk.rs:62)
    u32 smaxc[8], sminc[8], lmaxc[8], lminc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      smaxc[j] = j < MAXLEN ? S.cnt[C_SMAX + j] : 0u;
      sminc[j] = j < MAXLEN ? S.cnt[C_SMIN + j] : 0u;
      lmaxc[j] = j < MAXLEN ? S.cnt[C_LMAX + j] : 0u;
      lminc[j] = j < MAXLEN ? cnt2[C2_LMIN + j] : 0u;
    }
    const u32 snm_len = cnt2[C2_SNM], lnm_len = cnt2[C2_LNM];
    const u32 leq_len = lnm_len - lminc[0];  // one equal bit per nodemap-0 node (log.rs:137-146)
    const u32 snap_size = 13u + bitmap_size(snm_len) + dac_size_from_counts(smaxc, 8) + dac_size_from_counts(sminc, 8);
    const u32 log_size = 13u + bitmap_size(lnm_len) + bitmap_size(leq_len) + dac_size_from_counts(lmaxc, 8) + dac_size_from_counts(lminc, 8);
    const bool as_snapshot = first || n_logs == 254u || snap_size <= log_size;
    const u32 my_size = as_snapshot ? snap_size : log_size;

    // ---------------- arena allocation for the winner
    if (tid == 0) {
      const u64 need = ((u64)my_size + 15ull) & ~15ull;
      const u64 off = atomicAdd(P.arena_head, (unsigned long long)need);
      S.piece_off = off;
      Piece pc;
      pc.off = off; pc.size = my_size; pc.kind = as_snapshot ? 1u : 0u;
      P.pieces[unit.piece_base + inst] = pc;
    }
    for (int w = tid; w < NM_WORDS; w += ENC_THREADS) { S.nm[w] = 0; S.eqw[w] = 0; }
    __syncthreads();
    const u64 piece_off = S.piece_off;
    const bool fits = piece_off + (((u64)my_size + 15ull) & ~15ull) <= P.arena_cap;
    if (!fits) err |= EF_ARENA_FULL;

    // ---------------- BFS positions of the winner: per-level ranks of internal nodes
    bool in_[5];
#pragma unroll
    for (int l = 0; l < 5; l++) in_[l] = as_snapshot ? si[l] : li[l];
    const u32 in5 = as_snapshot ? si5 : li5;
    bool al[6];  // al[l]: this thread's level-l frame node exists in the tree (l = 0..4), al[5] = quads exist
    {
      bool a = intree4 || lo == 5;
#pragma unroll
      for (int l = 0; l < 5; l++) {
        al[l] = l >= lo && a && intree4;
        if (l >= lo) a = a && in_[l];
      }
      al[5] = lo == 5 ? (tid == 0) : a;
    }
    // flags of internal nodes owned by this thread, per level
    u32 rank_in_warp[5];
    {
#pragma unroll
      for (int l = 1; l < 5; l++) {
        const bool f = owner[l] && al[l] && in_[l];
        const u32 b = __ballot_sync(0xffffffffu, f);
        rank_in_warp[l] = __popc(b & lanemask_lt());
        if (lane == 0) S.wtot[l][warp] = __popc(b);
      }
      rank_in_warp[0] = 0;
    }
    // level 5: quads
    u32 q_before = 0, q_count = 0;
    {
      const u32 mine = al[5] ? (in5 & quad_mask) : 0u;
      u32 before = 0, tot = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const u32 b = __ballot_sync(0xffffffffu, (mine >> q) & 1u);
        before += __popc(b & lanemask_lt());
        tot += __popc(b);
      }
      q_before = before;
      q_count = __popc(mine);
      if (lane == 0) S.wtot[5][warp] = tot;
    }
    __syncthreads();
    u32 I[6], Rbase[6];  // I[l]: internal nodes at level l; Rbase[l]: internal nodes at level l in earlier warps
    I[0] = (lo == 0 && in_[0]) ? 1u : 0u;  // root flag is uniform across the CTA
    Rbase[0] = 0;
#pragma unroll
    for (int l = 1; l < 6; l++) {
      u32 tot = 0, bs = 0;
#pragma unroll
      for (int w = 0; w < ENC_WARPS; w++) {
        const u32 x = S.wtot[l][w];
        if (w < warp) bs += x;
        tot += x;
      }
      I[l] = tot; Rbase[l] = bs;
    }
    // P[l]: BFS index of the first level-l node; M[l]: internal nodes on levels above l
    u32 Pn[8], Mn[8];
    {
      u32 p = 0, mm = 0, e = 1;  // e = nodes on the current level
#pragma unroll
      for (int l = 0; l <= 6; l++) {
        Pn[l] = p; Mn[l] = mm;
        if (l < lo) continue;
        p += e;
        const u32 il = l < 6 ? I[l] : 0u;
        mm += il;
        e = 4 * il;
      }
      Pn[7] = p; Mn[7] = mm;
    }
    const u32 n_max = Pn[7], n_min = Mn[6], nm_len = Pn[6];
    // rank (among internal nodes of its level) of each chain node, as seen by this thread
    u32 R[5];
    R[0] = 0;
    R[1] = 0;
    {
      // level 1: four nodes, flags known to everyone through smem-resident N1 of each... recompute from wtot:
      // owners of level-1 nodes are threads 0,64,128,192 = lane 0 of warps 0,2,4,6
      u32 r1 = 0;
      for (int w = 0; w < (tid >> 6) * 2; w++) r1 += S.wtot[1][w];
      R[1] = r1;
    }
    {
      const u32 r2own = Rbase[2] + rank_in_warp[2];
      R[2] = shfl(r2own, lane & ~15);
      const u32 r3own = Rbase[3] + rank_in_warp[3];
      R[3] = shfl(r3own, lane & ~3);
      R[4] = Rbase[4] + rank_in_warp[4];
    }
    const u32 R5 = Rbase[5] + q_before;  // rank of this thread's first internal quad

    // ---------------- stage values / bits in BFS order
    auto set_bit = [](u32* words, u32 pos) { atomicOr(&words[pos >> 5], 0x80000000u >> (pos & 31u)); };
#pragma unroll
    for (int l = 0; l < 5; l++) {
      if (l < lo || !(owner[l] && al[l])) continue;
      const u32 child = l == 0 ? 0u : ((u32)(tid >> (2 * (4 - l))) & 3u);
      const u32 pos = l == lo ? 0u : Pn[l] + 4u * R[l > 0 ? l - 1 : 0] + child;
      const V tm = or0(cmax[l], VT<V>::NONE_MAX);
      const V pmax = cmax[l > 0 ? l - 1 : 0], pmin = cmin[l > 0 ? l - 1 : 0];
      V e;
      if (as_snapshot) e = l == lo ? tm : (V)(pmax - tm);
      else e = (V)(tm - or0(l == 4 ? s4max : csmax[l], VT<V>::NONE_MAX));
      S.A[pos] = VT<V>::zz(e);
      const u32 ones_before = Mn[l] + R[l];
      if (in_[l]) {
        set_bit(S.nm, pos);
        V mv;
        if (as_snapshot) mv = l == lo ? cmin[l] : (V)(cmin[l] - pmin);
        else mv = (V)(cmin[l] - (l == 4 ? s4min : csmin[l]));
        S.B[ones_before] = VT<V>::zz(mv);
      } else if (!as_snapshot) {
        const bool eqf = l == 0 ? (bool)N0.eq : l == 1 ? (bool)N1.eq : l == 2 ? eq2 : l == 3 ? eq3 : eq4;
        const bool un = l == 0 ? u0 : l == 1 ? u1 : l == 2 ? u2 : l == 3 ? u3 : u4;
        if (!un && eqf) set_bit(S.eqw, pos - ones_before);
      }
    }
    if (al[5]) {
      u32 r5 = R5;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (!((quad_mask >> q) & 1u)) continue;
        const u32 pos = lo == 5 ? 0u : Pn[5] + 4u * R[4] + (u32)q;
        const V tm = or0(t5max[q], VT<V>::NONE_MAX);
        V e;
        if (as_snapshot) e = lo == 5 ? tm : (V)(t4max - tm);
        else e = (V)(tm - or0(s5max[q], VT<V>::NONE_MAX));
        S.A[pos] = VT<V>::zz(e);
        const u32 ones_before = Mn[5] + r5;
        if ((in5 >> q) & 1u) {
          set_bit(S.nm, pos);
          V mv;
          if (as_snapshot) mv = lo == 5 ? t5min[q] : (V)(t5min[q] - t4min);
          else mv = (V)(t5min[q] - s5min[q]);
          S.B[ones_before] = VT<V>::zz(mv);
          const u32 lpos = Pn[6] + 4u * r5;
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const int m = 4 * q + c;
            const bool in = (cur_inb >> m) & 1u;
            V le;
            if (as_snapshot) le = (V)(t5max[q] - (in ? tv[m] : (V)0));
            else le = in ? (V)(tv[m] - sv[m]) : (V)0;
            S.A[lpos + c] = VT<V>::zz(le);
          }
          r5++;
        } else if (!as_snapshot) {
          if (!((u5 >> q) & 1u) && ((eq5 >> q) & 1u)) set_bit(S.eqw, pos - ones_before);
        }
      }
    }
    (void)q_count;
    __syncthreads();

    // ---------------- serialize (snapshot.rs:48-58 / log.rs:53-64)
    {
      const bool staged = my_size <= (u32)STAGE_BYTES;
      u8* out = staged ? S.stage : (fits ? P.arena + piece_off : nullptr);
      if (out != nullptr) {
        const u32 sidelen = 64u >> lo;
        if (tid == 0) {
          out[0] = 2;  // k
          store_be32(out + 1, (u32)unit.rows);
          store_be32(out + 5, (u32)unit.cols);
          store_be32(out + 9, sidelen);
        }
        u32 off = 13;
        off += block_bitmap_emit(S.nm, nm_len, out + off);
        if (!as_snapshot) off += block_bitmap_emit(S.eqw, nm_len - n_min, out + off);
        const u32* cmaxc = as_snapshot ? smaxc : lmaxc;
        const u32* cminc = as_snapshot ? sminc : lminc;
        (void)n_max;
        off += block_dac_emit<U, MAXLEN>(S.A, cmaxc[0], cmaxc, out + off, S.scan, S.scan2);
        off += block_dac_emit<U, MAXLEN>(S.B, cminc[0], cminc, out + off, S.scan, S.scan2);
        if (off != my_size) err |= EF_BAD_FORMAT;  // internal consistency: emitted bytes == predicted size
        __syncthreads();
        if (staged && fits) {
          const uint4* src = reinterpret_cast<const uint4*>(S.stage);
          uint4* dst = reinterpret_cast<uint4*>(P.arena + piece_off);
          for (u32 i = tid; i < (my_size + 15u) / 16u; i += ENC_THREADS) dst[i] = src[i];
        }
      }
    }

    // ---------------- bookkeeping: start a new block or extend the current one
    if (as_snapshot) {
#pragma unroll
      for (int m = 0; m < 16; m++) sv[m] = tv[m];
      s3max = t3max; s3min = t3min; s2max = t2max; s2min = t2min;
      if (tid < 4) { S.s_l1max[tid] = S.l1[tid].tmax; S.s_l1min[tid] = S.l1[tid].tmin; }
      if (tid == 0) { S.s_l0max = N0.tmax; S.s_l0min = N0.tmin; }
      n_snap++;
      n_logs = 0;
      total_bytes += 1;  // Block's n_instants byte (block.rs:88-95)
    } else {
      n_logs++;
      n_log_total++;
    }
    total_bytes += my_size;
    __syncthreads();
  }

  if (tid == 0) {
    UnitResult r;
    r.bytes = total_bytes + 6;  // encoding + fractional_bits + n_blocks (chunk.rs:235-243)
    r.snapshots = n_snap;
    r.logs = n_log_total;
    P.results[unit_idx] = r;
  }
  // one atomic per warp for the error bits
  err = __reduce_or_sync(0xffffffffu, err);
  if (lane == 0 && err) atomicOr(P.err, err);
}

}  // namespace dcdf
