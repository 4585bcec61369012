// decode_tile_common.cuh -- pieces shared by the tile decoders (decode_tile4.cuh: windows, decode_search4.cuh:
// value-range search) for trees up to 64x64: constants, output conversion (from_fixed, fixed.rs:81-86), cp.async
// staging, four-entry DAC / bitmap fetches for the children of one node (their BFS indices are consecutive:
// 1 + 4 * rank(idx) .. + 3, snapshot.rs:177), the window-clipped quad writer.
#pragma once
#include "decode.cuh"

namespace dcdf {

constexpr int DT_THREADS = 256;
constexpr int DT_WARPS = DT_THREADS / 32;
constexpr int DT_NODES = 5461;
constexpr int DT_UPPER = 1365;

DCDF_DEVINL u32 lvl_off(int k) { return (0x55555555u >> (32 - 2 * k)) & (k ? 0xffffffffu : 0u); }  // (4^k - 1) / 3

// Output of one cell: raw fixed point, or the chunk's own encoding with from_fixed applied to floats (fixed.rs:81-86;
// the divide by 2^(bits+1) is an exact scaling, done as a multiply by the exact reciprocal).
struct CellOut {
  void* out;
  int kind;  // 0: i64 (raw / ENC_I64), 1: i32, 2: f32, 3: f64
  float inv32;
  double inv64;
  DCDF_DEVINL void init(const QuerySet& Q, void* out_, int raw, int bits) {
    out = out_;
    kind = (raw || Q.encoding == 8) ? 0 : Q.encoding == 4 ? 1 : Q.encoding == 32 ? 2 : 3;
    inv32 = __int_as_float((126 - bits) << 23);                          // 2^-(bits+1)
    inv64 = __longlong_as_double((long long)(1022 - bits) << 52);
  }
  // 32-bit values (narrow expansion): the same conversions without 64-bit arithmetic
  DCDF_DEVINL void put(u64 i, int32_t fixed) const {
    if (kind == 2) static_cast<float*>(out)[i] = fixed == 0 ? __int_as_float(0x7fc00000) : __int2float_rn(fixed - 1) * inv32;
    else if (kind == 0) static_cast<i64*>(out)[i] = (i64)fixed;
    else if (kind == 1) static_cast<int32_t*>(out)[i] = fixed;
    else static_cast<double*>(out)[i] = fixed == 0 ? __longlong_as_double(0x7ff8000000000000ll) : __int2double_rn(fixed - 1) * inv64;
  }
  DCDF_DEVINL void put(u64 i, i64 fixed) const {
    if (kind == 2) static_cast<float*>(out)[i] = fixed == 0 ? __int_as_float(0x7fc00000) : __ll2float_rn(fixed - 1) * inv32;
    else if (kind == 0) static_cast<i64*>(out)[i] = fixed;
    else if (kind == 1) static_cast<int32_t*>(out)[i] = (int32_t)fixed;
    else static_cast<double*>(out)[i] = fixed == 0 ? __longlong_as_double(0x7ff8000000000000ll) : __ll2double_rn(fixed - 1) * inv64;
  }
};

struct TileWindowParams {
  QuerySet Q;
  const CubeDev* cubes;   // validated, ordered
  const u64* out_off;     // [n] element offsets
  const u64* job_base;    // [n + 1] prefix of (slices x subchunks) per window
  u64 n_queries, n_jobs;
  void* out;
  int raw;
};

constexpr u32 W3_NONE = 0xffffffffu;
constexpr int W3_UPPER = DT_UPPER + 3;  // levels above the cells, level k >= 1 at (4^k - 1) / 3 + 3 (16-byte aligned groups)
constexpr int W3_DIRW = (int)(sizeof(InstDir) / 4);

DCDF_DEVINL u32 off3(int k) { return k ? lvl_off(k) + 3u : 0u; }

DCDF_DEVINL void cp_async16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((u32)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
DCDF_DEVINL void cp_async4(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((u32)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
DCDF_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Level 0 of a DAC (bytes + continuation bits) with the general accessor for longer codes.
struct Dac4 {
  const u8* bytes0;
  const u8* more0;
  const u8* chunk;
  const DacDir* d;
  u32 len0;
};
DCDF_DEVINL Dac4 dac4_of(const u8* chunk, const DacDir* d) {
  const u32 len = d->len[0], base = d->base[0];
  const u32 words = base + 8u + 4u * (len / 128u);
  return Dac4{chunk + words + 4u * ((len + 31u) / 32u), chunk + words, chunk, d, d->n_levels ? len : 0u};
}
// codes longer than one byte: out of line, they are rare and the rank loop would be inlined a dozen times
__device__ __noinline__ i64 dac_slow_get(const u8* chunk, const DacDir* d, u32 idx) { return DacRef{chunk, d}.get(idx); }
template <typename V>
DCDF_DEVINL V unzz8(u32 b) { return (V)(int)((b >> 1) ^ (0u - (b & 1u))); }
template <typename V>
DCDF_DEVINL V dac_get1(const Dac4& m, u32 idx) {  // dac.rs:80-93
  if (idx >= m.len0) return (V)0;
  if (!((m.more0[idx >> 3] >> (7u - (idx & 7u))) & 1u)) return unzz8<V>(m.bytes0[idx]);
  return (V)dac_slow_get(m.chunk, m.d, idx);
}
// entries idx .. idx + 3 (the children of one node); anything but four one-byte codes goes out of line
template <typename V>
struct __align__(16) Quad { V c[4]; };
struct __align__(16) Quad32 { u32 c[4]; };
template <typename V>
__device__ __noinline__ Quad<V> dac_get4_slow(const u8* chunk, const DacDir* d, u32 idx) {
  Quad<V> q;
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
    const V v = (V)DacRef{chunk, d}.get(idx + (u32)i);  // an empty DAC or an index past its end yields 0
    if (i == 0) q.c[0] = v; else if (i == 1) q.c[1] = v; else if (i == 2) q.c[2] = v; else q.c[3] = v;
  }
  return q;
}
// Four consecutive entries idx .. idx + 3 <= len0 of which some are longer than one byte (nib: continuation bits of
// level 0, bit 3 = entry idx).  The entries that continue sit next to each other on the next level as well, so every
// level costs ONE rank over the serialized directory (bitmap.rs:186-212) for the whole group, not one per entry
// (dac.rs:80-93 per entry would).
template <typename V>
__device__ __noinline__ Quad<V> dac_get4_multi(const u8* chunk, const DacDir* d, u32 idx, u32 nib) {
  u64 n0, n1, n2, n3;
  {
    const u32 len = d->len[0], words = d->base[0] + 8u + 4u * (len / 128u);
    const u8* b = chunk + words + 4u * ((len + 31u) / 32u) + idx;
    n0 = b[0]; n1 = b[1]; n2 = b[2]; n3 = b[3];
  }
  u32 cont = nib & 15u;  // bit 3 - i: entry i continues on the next level
  u32 p = idx;           // position of the group's first entry on the current level
  const u32 nl = d->n_levels;
#pragma unroll 1
  for (u32 j = 1; j < nl && cont; j++) {
    const BitMapRef up{chunk, d->len[j - 1], d->base[j - 1]};
    // the first continuing entry of the group sits at p + (entries before it that stopped); ones before it = ones before p
    // only if nothing between p and it is set, which holds: the entries in between did not continue
    const u32 r = up.rank(p);
    const u32 len = d->len[j], words = d->base[j] + 8u + 4u * (len / 128u);
    const u8* bytes = chunk + words + 4u * ((len + 31u) / 32u);
    const u8* more = chunk + words;
    const bool last = j + 1u >= nl;
    u32 next = 0, k = 0;
    const u32 sh = 8u * j;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if ((cont >> (3 - i)) & 1u) {
        const u32 pos = r + k;
        k++;
        if (pos < len) {  // malformed input guard
          const u64 byte = bytes[pos];
          if (i == 0) n0 |= byte << sh; else if (i == 1) n1 |= byte << sh; else if (i == 2) n2 |= byte << sh; else n3 |= byte << sh;
          if (!last && ((more[pos >> 3] >> (7u - (pos & 7u))) & 1u)) next |= 1u << (3 - i);
        }
      }
    }
    cont = next;
    p = r;
  }
  Quad<V> q;
  q.c[0] = (V)unzigzag64(n0); q.c[1] = (V)unzigzag64(n1); q.c[2] = (V)unzigzag64(n2); q.c[3] = (V)unzigzag64(n3);
  return q;
}
template <typename V>
DCDF_DEVINL void dac_get4(const Dac4& m, u32 idx, V (&d)[4]) {
  u32 nib = 1;
  if (idx + 4u <= m.len0) {
    const u32 by = idx >> 3;
    const u32 hw = ((u32)m.more0[by] << 8) | (u32)m.more0[by + 1];  // the byte after the bitmap is the DAC's first code
    nib = (hw >> (12u - (idx & 7u))) & 15u;
  }
  if (nib == 0) {
    const u8* b = m.bytes0 + idx;
#pragma unroll
    for (int i = 0; i < 4; i++) d[i] = unzz8<V>(b[i]);
  } else {
    const Quad<V> q = idx + 4u <= m.len0 ? dac_get4_multi<V>(m.chunk, m.d, idx, nib) : dac_get4_slow<V>(m.chunk, m.d, idx);
#pragma unroll
    for (int i = 0; i < 4; i++) d[i] = q.c[i];
  }
}
DCDF_DEVINL bool bit_at(const u8* bits, u32 i) { return (bits[i >> 3] >> (7u - (i & 7u))) & 1u; }
DCDF_DEVINL const u8* bitmap_bits(const u8* chunk, u32 len, u32 base) { return chunk + base + 8u + 4u * (len / 128u); }
// bits idx .. idx + 3 of an MSB-first bit stream as a mask (bit c = stream bit idx + c); positions >= len read as 0
DCDF_DEVINL u32 bits4(const u8* bits, u32 len, u32 idx) {
  const u32 by = idx >> 3;
  const u32 hw = ((u32)bits[by] << 8) | (u32)bits[by + 1];
  u32 nib = __brev((hw >> (12u - (idx & 7u))) & 15u) >> 28;
  if (idx + 4u > len) nib &= idx >= len ? 0u : (1u << (len - idx)) - 1u;
  return nib;
}

DCDF_DEVINL u32 below(u32 inb, int c) { return __popc(inb & ((1u << c) - 1u)); }

// Where the cells of the current (window, tile, instant) go.
struct QuadOut {
  static constexpr bool search_quirk = false;
  CellOut co;
  u64 base;       // element index of tile cell (0, 0) at this instant
  i64 pitch;      // window columns
  int top, bottom, left, right;  // window clipped to the tile, tile coordinates
  bool vec, vec4; // f32 output whose row pairs / row quadruples are 8 / 16-byte aligned
  template <typename V>
  DCDF_DEVINL float cvt(V v) const {  // from_fixed (fixed.rs:81-86): 0 -> NaN, else (v - 1) * 2^-(bits+1), exact scaling
    const float f = (sizeof(V) == 4 ? __int2float_rn((int)v - 1) : __ll2float_rn((i64)v - 1)) * co.inv32;
    return v == 0 ? __int_as_float(0x7fc00000) : f;
  }
  DCDF_DEVINL bool touches(int r0, int c0, int side) const { return r0 + side > top && r0 < bottom && c0 + side > left && c0 < right; }
  DCDF_DEVINL bool inside(int r0, int c0, int side) const { return r0 >= top && r0 + side <= bottom && c0 >= left && c0 + side <= right; }
  template <typename V>
  DCDF_DEVINL void put(int r0, int c0, const V (&v)[4]) const {  // one 2x2 quad
    if (!touches(r0, c0, 2)) return;
    const u64 i00 = base + (u64)((i64)r0 * pitch + c0);
    const bool in = inside(r0, c0, 2);
    if (in && vec) {
      float* o = static_cast<float*>(co.out) + i00;
      *reinterpret_cast<float2*>(o) = make_float2(cvt(v[0]), cvt(v[1]));
      *reinterpret_cast<float2*>(o + pitch) = make_float2(cvt(v[2]), cvt(v[3]));
      return;
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int r = r0 + (c >> 1), col = c0 + (c & 1);
      if (!in && (r < top || r >= bottom || col < left || col >= right)) continue;
      co.put(i00 + (u64)((c >> 1) ? pitch : 0) + (u64)(c & 1), v[c]);
    }
  }
  // two quads side by side (a 2x4 strip at (r0, c0)); fast = the whole 4x4 block is inside the window and vec4 holds
  template <typename V>
  DCDF_DEVINL void put_pair(bool fast, int r0, int c0, const V (&a)[4], const V (&b)[4]) const {
    if (fast) {
      float* o = static_cast<float*>(co.out) + (base + (u64)((i64)r0 * pitch + c0));
      *reinterpret_cast<float4*>(o) = make_float4(cvt(a[0]), cvt(a[1]), cvt(b[0]), cvt(b[1]));
      *reinterpret_cast<float4*>(o + pitch) = make_float4(cvt(a[2]), cvt(a[3]), cvt(b[2]), cvt(b[3]));
      return;
    }
    put(r0, c0, a);
    put(r0, c0 + 2, b);
  }
};

// Cell series through the tile decoder (k_cell_tiles4): the cells of a tile that some series of the batch asks for.
struct CellRef {
  u64 base;        // output element of the series' first instant
  i64 start, end;  // the series covers instants [start, end)
  u32 cell;        // 16 * block + 4 * quad + cell: block = Morton index of the 4x4 block, quad = 2 * (row / 2) + col / 2 and
  u32 pad_;        // cell = 2 * (row & 1) + (col & 1) inside the block -- the order the block decoder produces values in
};
// Same interface as QuadOut towards the block decoder: the values of the blocks that hold an asked-for cell go to an image
// of the tile in shared memory (one instant), from which the CTA then serves every series of the tile.
template <typename V>
struct SeriesOut {
  static constexpr bool search_quirk = false;
  V* img;     // the thread's 16 cells: img[4 * quad + cell]
  u32 mask;   // asked-for cells of the block (bit 4 * quad + cell); 0: nothing to decode
  static constexpr bool vec4 = false;
  DCDF_DEVINL bool touches(int, int, int) const { return mask != 0; }
  DCDF_DEVINL bool inside(int, int, int) const { return false; }
  DCDF_DEVINL void put(int r0, int c0, const V (&v)[4]) const {  // the 2x2 quad at (r0, c0): only its place inside the block matters
    const u32 quad = 2u * (((u32)r0 >> 1) & 1u) + (((u32)c0 >> 1) & 1u);
    Quad<V> q;
#pragma unroll
    for (int i = 0; i < 4; i++) q.c[i] = v[i];
    reinterpret_cast<Quad<V>*>(img)[quad] = q;
  }
  DCDF_DEVINL void put_pair(bool, int r0, int c0, const V (&a)[4], const V (&b)[4]) const {
    if (!(mask & (0xffu << (8u * (((u32)r0 >> 1) & 1u))))) return;  // nothing asked for in this half of the block
    put(r0, c0, a);
    put(r0, c0 + 2, b);
  }
};

// Value-range COUNTS of many windows over one tile (k_count_tiles4): the windows of a batch that touch the tile in this time
// slice share one decode of every instant.  An entry = one (window, tile): its rectangle inside the tile, its band, the
// instants of the slice it covers and where its per-instant counts go.
struct CountEntry {
  i64 cnt_base;          // counts[cnt_base + t] for the slice-local instant t
  i64 lower, upper;
  uint16_t t0, t1;       // slice-local instants [t0, t1)
  u8 top, bottom, left, right;  // tile coordinates, exclusive ends (<= 64)
};
constexpr int CT_MAX = 64;  // entries per job (a tile touched by more windows is split into several jobs)
template <typename V>
struct CountShared {
  CountEntry ent[CT_MAX];
  V lo[CT_MAX], hi[CT_MAX];  // the entry's band in the expansion's value type (clamped; lo > hi: no value can match)
  u32 cnt[2][CT_MAX];   // hits of the instant being decoded / of the previous one (flushed while the next one is decoded)
  u32 keep[CT_MAX];     // 0: pruned by the superchunk's min / max tables (superchunk.rs:480-493)
  unsigned long long act[2];  // entries that cover the instant being decoded (by instant parity)
  V smin0, smax0;       // root of the block's snapshot
  u32 n;
};
// Same interface as QuadOut towards the block decoder: the values of one half of the thread's block are tested against
// every entry whose rectangle touches the block.  The hit SET of the reference's traversal (snapshot.rs:310-421,
// log.rs:519-702) is "value inside the band" because every node test uses exact bounds -- with one exception that is
// reproduced here: a Log's ROOT is tested with min = snapshot.min.get(0) + log.min.get(0) and max = snapshot.max.get(0) +
// log.max.get(0) (log.rs:527-548), and an empty min Dac yields 0 (dac.rs:80-93).  So a Log over a single-node Snapshot, and
// a Log that is a single node itself, are tested against a wrong lower bound; a single-node Log is
// moreover read as "snapshot + root entry" whether or not its `equal` bit is set.  log4 builds those values
// (search_quirk); the root test is applied per entry below, to every Log instant (it changes nothing when it is exact).
template <typename V>
struct CountOut {
  static constexpr bool search_quirk = true;
  static constexpr bool vec4 = false;
  CountShared<V>* C;
  unsigned long long mine;  // entries whose rectangle touches the thread's block
  int R0, C0;               // the block's origin
  u32 t;                    // slice-local instant
  u32 buf;                  // counter set / activity mask of this instant
  bool is_log;              // the instant is a Log: the reference's root test applies
  V min_t, e_root;          // its root entries (min Dac: 0 when empty)
  DCDF_DEVINL bool touches(int, int, int) const { return mine != 0ull; }
  DCDF_DEVINL bool inside(int, int, int) const { return false; }
  DCDF_DEVINL void put_pair(bool, int r0, int c0, const V (&a)[4], const V (&b)[4]) const {
    unsigned long long m = mine & C->act[buf];
    if (!m) return;
    // the strip's own range first: most bands miss it or hold all of it
    const V vmin = min(min(min(a[0], a[1]), min(a[2], a[3])), min(min(b[0], b[1]), min(b[2], b[3])));
    const V vmax = max(max(max(a[0], a[1]), max(a[2], a[3])), max(max(b[0], b[1]), max(b[2], b[3])));
    while (m) {
      const int e = __ffsll((long long)m) - 1;
      m &= m - 1ull;
      const V lo = C->lo[e], hi = C->hi[e];
      bool all = false, none = false;
      if (is_log) {  // the traversal's root test (log.rs:527-548)
        const CountEntry& E = C->ent[e];
        const i64 mn = (i64)C->smin0 + (i64)min_t, mx = (i64)C->smax0 + (i64)e_root;
        all = mn >= E.lower && mx <= E.upper;
        none = !all && (mn > E.upper || mx < E.lower);
      }
      if (none || (!all && (vmax < lo || vmin > hi))) continue;
      // cells of this 2x4 strip inside the entry's rectangle: bits 0..3 = a, 4..7 = b (cell = 2 * (row & 1) + (col & 1))
      const uchar4 rc = *reinterpret_cast<const uchar4*>(&C->ent[e].top);  // top, bottom, left, right
      const u32 rowm = ((r0 >= (int)rc.x && r0 < (int)rc.y) ? 0x33u : 0u) | ((r0 + 1 >= (int)rc.x && r0 + 1 < (int)rc.y) ? 0xccu : 0u);
      const u32 colm = ((c0 >= (int)rc.z && c0 < (int)rc.w) ? 0x05u : 0u) | ((c0 + 1 >= (int)rc.z && c0 + 1 < (int)rc.w) ? 0x0au : 0u) |
                       ((c0 + 2 >= (int)rc.z && c0 + 2 < (int)rc.w) ? 0x50u : 0u) | ((c0 + 3 >= (int)rc.z && c0 + 3 < (int)rc.w) ? 0xa0u : 0u);
      const u32 in = rowm & colm;
      if (!in) continue;
      u32 hit = in;
      if (!all && !(vmin >= lo && vmax <= hi)) {
        hit = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          if (a[i] >= lo && a[i] <= hi) hit |= 1u << i;
          if (b[i] >= lo && b[i] <= hi) hit |= 16u << i;
        }
        hit &= in;
      }
      if (hit) atomicAdd(&C->cnt[buf][e], (u32)__popc(hit));
    }
  }
  DCDF_DEVINL void put(int r0, int c0, const V (&v)[4]) const {  // trees of side 2: one quad at (0, 0)
    // the second quad of the strip lies outside every rectangle of a 2x2 tile (right <= 2); it repeats the first one so
    // that the strip's range is the quad's
    CountOut o = *this;
    o.put_pair(false, r0, c0, v, v);
  }
};

// The `equal` bits of the children that stop here are consecutive: child c's bit is at e0 + (non-internal children
// before c), e0 = idx0 - rank1(idx0) (rank0(idx + 1) - 1, log.rs:265).  Returns a 16-bit window starting at e0's byte.
DCDF_DEVINL u32 eq_window(const u8* eq, u32 e0) {
  const u32 by = e0 >> 3;
  return ((u32)eq[by] << 8) | (u32)eq[by + 1];
}
DCDF_DEVINL bool eq_bit(u32 win, u32 e0, u32 k) { return (win >> (15u - (e0 & 7u) - k)) & 1u; }

}  // namespace dcdf
