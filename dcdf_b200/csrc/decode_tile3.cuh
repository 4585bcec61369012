// decode_tile3.cuh -- window decode of whole <=64x64 tiles, second generation of decode_tile.cuh.
//
// Same contract as k_window_tiles (Snapshot::fill_window snapshot.rs:204-301, Log::fill_window log.rs:311-508, routed
// as Chunk::fill_window chunk.rs:152-158 / Superchunk::fill_window superchunk.rs:402-457), fewer instructions:
//   * one thread per PARENT node: its four children have consecutive BFS indices (1 + 4 * rank .. + 3,
//     snapshot.rs:177), so their DAC entries are fetched together (one 4-bit test of the continuation bitmap, four
//     byte loads), their nodemap bits are one nibble, and their records are one 16-byte shared-memory store;
//   * every level of the snapshot pyramid is kept in Morton order: the four cells of a quad are one 16-byte word;
//   * the scan that turns "internal" flags into child BFS indices is fused into the level pass (flags stay in
//     registers between the ballot and the write), and the same scan yields rank1(idx), hence the position of a log
//     node's `equal` bit (rank0(idx + 1) - 1, log.rs:265) without touching the rank directory;
//   * the three top levels below the root are expanded by one warp without block barriers;
//   * the quad level is never written back: the thread that owns a 4x4 block of cells classifies its four quads and
//     writes the block's rows (16-byte stores for f32 output);
//   * a Snapshot instant inside the window is emitted by the same code as a Log ("equal" to its snapshot, offset 0);
//   * the next instant's bytes and directory entry are fetched with cp.async into the other half of a double
//     staging buffer while the current instant is expanded.
#pragma once
#include "decode_tile.cuh"

namespace dcdf {

constexpr u32 W3_NONE = 0xffffffffu;
constexpr int W3_UPPER = DT_UPPER + 3;  // levels above the cells, level k >= 1 at (4^k - 1) / 3 + 3 (16-byte aligned groups)
constexpr int W3_DIRW = (int)(sizeof(InstDir) / 4);

template <typename V>
struct Tile3Smem {
  static constexpr int BUF = sizeof(V) == 4 ? 10 * 1024 : 16 * 1024;
  __align__(16) V cells[4096];     // snapshot values of the cells, Morton order (quad q = cells[4q .. 4q+3])
  __align__(16) V sup[W3_UPPER];   // snapshot max values of the levels above the cells
  __align__(16) V pay[W3_UPPER];   // log expansion payload
  __align__(16) u32 meta[W3_UPPER];  // snapshot pass: BFS index of the first child or W3_NONE; log pass: (first child << 2) | mode
  u32 wtot[DT_WARPS];
  u32 nxt, isum, snap_single, pad_;
  __align__(16) InstDir dir[2];    // directory entries of the staged structures
  __align__(16) u8 stage[2][BUF + 32];
};

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
    const Quad<V> q = dac_get4_slow<V>(m.chunk, m.d, idx);
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

// Threads are parents in Morton order == BFS order among existing nodes.  inb = mask of this thread's internal
// children; ex = number of internal children of the threads before this one; returns the level total.
template <bool BLOCK>
DCDF_DEVINL u32 scan4(u32 inb, u32& ex, u32* wtot) {
  const u32 lt = lanemask_lt();
  u32 cnt = 0;
  ex = 0;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const u32 b = __ballot_sync(0xffffffffu, (inb >> c) & 1u);
    ex += __popc(b & lt);
    cnt += __popc(b);
  }
  if (!BLOCK) return cnt;
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) wtot[warp] = cnt;
  __syncthreads();
  u32 before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < DT_WARPS; w++) {
    const u32 x = wtot[w];
    if (w < warp) before += x;
    total += x;
  }
  ex += before;
  return total;
}
DCDF_DEVINL u32 below(u32 inb, int c) { return __popc(inb & ((1u << c) - 1u)); }

// Where the cells of the current (window, tile, instant) go.
struct QuadOut {
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

// ------------------------------------------------------------------ Snapshot
// Children of this thread's parent `par` (level lv-1): values and internal flags.
template <typename V>
DCDF_DEVINL void snap_children(const Tile3Smem<V>& S, const Dac4& mx, const u8* nm, u32 nm_len, u32 o0, u32 par, bool valid,
                               V (&val)[4], u32& inb) {
  inb = 0;
#pragma unroll
  for (int c = 0; c < 4; c++) val[c] = 0;
  if (valid) {
    const u32 pm = S.meta[o0 + par];
    const V pv = S.sup[o0 + par];
    V dq[4] = {0, 0, 0, 0};
    if (pm != W3_NONE) {
      dac_get4<V>(mx, pm, dq);  // snapshot.rs:179
      inb = bits4(nm, nm_len, pm);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) val[c] = pv - dq[c];
  }
}
// Produce level lv from level lv-1.  nxt = BFS index of the first node of level lv+1.
template <typename V, bool BLOCK>
DCDF_DEVINL u32 snap_level3(Tile3Smem<V>& S, const Dac4& mx, const u8* nm, u32 nm_len, int lv, u32 nxt) {
  const u32 par = BLOCK ? threadIdx.x : (threadIdx.x & 31u);
  const u32 n0 = 1u << (2 * (lv - 1)), o0 = off3(lv - 1), o1 = off3(lv);
  const bool valid = par < n0;
  V val[4];
  u32 inb, ex;
  snap_children<V>(S, mx, nm, nm_len, o0, par, valid, val, inb);
  const u32 total = scan4<BLOCK>(inb, ex, S.wtot);
  if (valid) {
    Quad<V> q;
    Quad32 m;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      q.c[c] = val[c];
      m.c[c] = (inb >> c) & 1u ? nxt + 4u * (ex + below(inb, c)) : W3_NONE;
    }
    *reinterpret_cast<Quad<V>*>(&S.sup[o1 + 4u * par]) = q;
    *reinterpret_cast<Quad32*>(&S.meta[o1 + 4u * par]) = m;
  }
  return total;
}

// Expand the Snapshot at `chunk + d.off` into S.sup / S.cells.  With `as_instant` the records of the quads' parents
// are left as "equal, offset 0", so that final3 emits the snapshot's own cells.  No trailing barrier: the caller's next
// barrier publishes the pyramid.
template <typename V>
DCDF_DEVINL void expand_snapshot3(const u8* chunk, const InstDir& d, int L, Tile3Smem<V>& S, bool as_instant) {
  const int tid = threadIdx.x;
  const u8* nm = bitmap_bits(chunk, d.nm_len, d.nm_base);
  const u32 nm_len = d.nm_len;
  const Dac4 mx = dac4_of(chunk, &d.max);
  if (tid < 32) {
    const bool internal0 = bit_at(nm, 0);
    if (tid == 0) {
      S.sup[0] = dac_get1<V>(mx, 0);
      S.meta[0] = internal0 ? 1u : W3_NONE;
      S.snap_single = internal0 ? 0u : 1u;
    }
    u32 nxt = 1u + (internal0 ? 4u : 0u);
    __syncwarp();
    for (int lv = 1; lv <= 3 && lv <= L - 2; lv++) {
      nxt += 4u * snap_level3<V, false>(S, mx, nm, nm_len, lv, nxt);
      __syncwarp();
    }
    if (tid == 0) S.nxt = nxt;
  }
  __syncthreads();
  u32 nxt = S.nxt;
  for (int lv = 4; lv <= L - 2; lv++) {
    nxt += 4u * snap_level3<V, true>(S, mx, nm, nm_len, lv, nxt);
    __syncthreads();
  }
  if (L == 1) {
    if (tid == 0) {
      const u32 pm = S.meta[0];
      const V pv = S.sup[0];
      V dd[4] = {0, 0, 0, 0};
      if (pm != W3_NONE) dac_get4<V>(mx, pm, dd);
      Quad<V> q;
#pragma unroll
      for (int i = 0; i < 4; i++) q.c[i] = pv - dd[i];
      reinterpret_cast<Quad<V>*>(S.cells)[0] = q;
      if (as_instant) { S.meta[0] = 2u; S.pay[0] = 0; }
    }
    return;
  }
  // the quads (level L-1) and their cells: one thread per parent of four quads
  const int lvp = L - 2;
  const u32 par = (u32)tid, n0 = 1u << (2 * lvp), oP = off3(lvp), oQ = off3(L - 1);
  const bool valid = par < n0;
  V val[4];
  u32 inb, ex;
  snap_children<V>(S, mx, nm, nm_len, oP, par, valid, val, inb);
  scan4<true>(inb, ex, S.wtot);
  if (valid) {
    Quad<V> qv;
#pragma unroll
    for (int c = 0; c < 4; c++) qv.c[c] = val[c];
    *reinterpret_cast<Quad<V>*>(&S.sup[oQ + 4u * par]) = qv;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      V dd[4] = {0, 0, 0, 0};
      if ((inb >> c) & 1u) dac_get4<V>(mx, nxt + 4u * (ex + below(inb, c)), dd);
      Quad<V> q;
#pragma unroll
      for (int i = 0; i < 4; i++) q.c[i] = val[c] - dd[i];
      reinterpret_cast<Quad<V>*>(S.cells)[4u * par + (u32)c] = q;
    }
    if (as_instant) { S.meta[oP + par] = 2u; S.pay[oP + par] = 0; }  // read back by this same thread in final3
  }
}

// ------------------------------------------------------------------ Log
// Per node: mode 0 internal (payload = max_t, replaced at every level, log.rs:233), 1 uniform (payload = value),
// 2 equal (value = payload + snapshot cell).  isum = internal nodes of the levels above, so that
// rank1(idx) = isum + (internal nodes of this level before idx).
template <typename V>
struct LogKids {
  u32 pmode, inb, idx0;
  V pp;
  V dq[4];
};
template <typename V>
DCDF_DEVINL void log_children(const Tile3Smem<V>& S, const Dac4& mx, const u8* nm, u32 nm_len, u32 o0, u32 par, bool valid, LogKids<V>& k) {
  k.pmode = 3; k.inb = 0; k.idx0 = 0; k.pp = 0;
#pragma unroll
  for (int c = 0; c < 4; c++) k.dq[c] = 0;
  if (valid) {
    const u32 pm = S.meta[o0 + par];
    k.pp = S.pay[o0 + par];
    k.pmode = pm & 3u;
    if (k.pmode == 0) {
      k.idx0 = pm >> 2;
      dac_get4<V>(mx, k.idx0, k.dq);
      k.inb = bits4(nm, nm_len, k.idx0);
    }
  }
}
// The `equal` bits of the children that stop here are consecutive: child c's bit is at e0 + (non-internal children
// before c), e0 = idx0 - rank1(idx0) (rank0(idx + 1) - 1, log.rs:265).  Returns a 16-bit window starting at e0's byte.
DCDF_DEVINL u32 eq_window(const u8* eq, u32 e0) {
  const u32 by = e0 >> 3;
  return ((u32)eq[by] << 8) | (u32)eq[by + 1];
}
DCDF_DEVINL bool eq_bit(u32 win, u32 e0, u32 k) { return (win >> (15u - (e0 & 7u) - k)) & 1u; }

template <typename V, bool BLOCK>
DCDF_DEVINL u32 log_level3(Tile3Smem<V>& S, const Dac4& mx, const u8* nm, u32 nm_len, const u8* eq, int lv, u32 nxt, u32 isum) {
  const u32 par = BLOCK ? threadIdx.x : (threadIdx.x & 31u);
  const u32 n0 = 1u << (2 * (lv - 1)), o0 = off3(lv - 1), o1 = off3(lv);
  const bool valid = par < n0;
  LogKids<V> k;
  u32 ex;
  log_children<V>(S, mx, nm, nm_len, o0, par, valid, k);
  const u32 total = scan4<BLOCK>(k.inb, ex, S.wtot);
  if (valid) {
    Quad<V> p;
    Quad32 m;
    if (k.pmode != 0) {
#pragma unroll
      for (int c = 0; c < 4; c++) { p.c[c] = k.pp; m.c[c] = k.pmode; }
    } else {
      const Quad<V> sn = *reinterpret_cast<const Quad<V>*>(&S.sup[o1 + 4u * par]);
      const u32 e0 = k.idx0 - (isum + ex);
      const u32 win = k.inb != 15u ? eq_window(eq, e0) : 0u;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const u32 bl = below(k.inb, c);
        if ((k.inb >> c) & 1u) {
          p.c[c] = k.dq[c];
          m.c[c] = (nxt + 4u * (ex + bl)) << 2;
        } else {
          const bool e = eq_bit(win, e0, (u32)c - bl);
          p.c[c] = e ? k.dq[c] : k.dq[c] + sn.c[c];  // uniform: max_t + max_s of this node (log.rs:266-268)
          m.c[c] = e ? 2u : 1u;
        }
      }
    }
    *reinterpret_cast<Quad<V>*>(&S.pay[o1 + 4u * par]) = p;
    *reinterpret_cast<Quad32*>(&S.meta[o1 + 4u * par]) = m;
  }
  return total;
}

// Root and the levels down to the quads' parents (level L-2) of the Log at `chunk + d.off`.  Ends with a barrier.
template <typename V>
DCDF_DEVINL void log_top3(const u8* chunk, const InstDir& d, int L, Tile3Smem<V>& S) {
  const int tid = threadIdx.x;
  const u8* nm = bitmap_bits(chunk, d.nm_len, d.nm_base);
  const u8* eq = bitmap_bits(chunk, d.eq_len, d.eq_base);
  const u32 nm_len = d.nm_len;
  const Dac4 mx = dac4_of(chunk, &d.max);
  if (tid < 32) {
    const bool internal0 = bit_at(nm, 0);
    if (tid == 0) {
      const V d0 = dac_get1<V>(mx, 0);
      if (internal0) {
        S.meta[0] = 1u << 2; S.pay[0] = d0;
      } else {
        // log.rs:180-186: a single-node log is uniform unless its equal bit says "snapshot + constant"
        const bool uniform = S.snap_single || !bit_at(eq, 0);
        S.meta[0] = uniform ? 1u : 2u;
        S.pay[0] = uniform ? d0 + S.sup[0] : d0;
      }
    }
    u32 isum = internal0 ? 1u : 0u, nxt = 1u + 4u * isum;
    __syncwarp();
    for (int lv = 1; lv <= 3 && lv <= L - 2; lv++) {
      const u32 tot = log_level3<V, false>(S, mx, nm, nm_len, eq, lv, nxt, isum);
      isum += tot; nxt += 4u * tot;
      __syncwarp();
    }
    if (tid == 0) { S.nxt = nxt; S.isum = isum; }
  }
  __syncthreads();
  u32 nxt = S.nxt, isum = S.isum;
  for (int lv = 4; lv <= L - 2; lv++) {
    const u32 tot = log_level3<V, true>(S, mx, nm, nm_len, eq, lv, nxt, isum);
    isum += tot; nxt += 4u * tot;
    __syncthreads();  // every thread has read S.nxt / S.isum (scan4's barrier) before they are replaced
    if (tid == 0) { S.nxt = nxt; S.isum = isum; }
  }
}

// The quads (level L-1) of the structure whose records stand at level L-2, and the window's cells of this instant.
// One thread per parent of four quads == one 4x4 block of cells.
template <typename V>
DCDF_DEVINL void final3(const u8* chunk, const InstDir& d, int L, Tile3Smem<V>& S, const QuadOut& O) {
  const int tid = threadIdx.x;
  const u8* nm = bitmap_bits(chunk, d.nm_len, d.nm_base);
  const u8* eq = bitmap_bits(chunk, d.eq_len, d.eq_base);
  const Dac4 mx = dac4_of(chunk, &d.max);
  if (L == 1) {
    // a 2x2 tile: the root is the quad
    if (tid == 0) {
      const u32 mode = S.meta[0] & 3u;
      const V pay = S.pay[0];
      const Quad<V> q = reinterpret_cast<const Quad<V>*>(S.cells)[0];
      V v[4], dd[4] = {0, 0, 0, 0};
      if (mode == 0) dac_get4<V>(mx, 1u, dd);
#pragma unroll
      for (int i = 0; i < 4; i++) v[i] = mode == 1 ? pay : mode == 2 ? pay + q.c[i] : dd[i] + q.c[i];
      O.put(0, 0, v);
    }
    return;
  }
  const int lvp = L - 2;
  const u32 par = (u32)tid, n0 = 1u << (2 * lvp), oP = off3(lvp), oQ = off3(L - 1);
  const bool valid = par < n0;
  LogKids<V> k;
  u32 ex;
  log_children<V>(S, mx, nm, d.nm_len, oP, par, valid, k);
  scan4<true>(k.inb, ex, S.wtot);
  const u32 nxt = S.nxt, isum = S.isum;  // published by log_top3 before its last barrier
  if (!valid) return;
  const int R0 = 4 * (int)morton_row(par), C0 = 4 * (int)morton_col(par);
  if (!O.touches(R0, C0, 4)) return;
  const bool fast = O.vec4 && O.inside(R0, C0, 4);
  Quad<V> sn;
#pragma unroll
  for (int c = 0; c < 4; c++) sn.c[c] = 0;
  u32 e0 = 0, win = 0;
  if (k.pmode == 0 && k.inb != 15u) {
    sn = *reinterpret_cast<const Quad<V>*>(&S.sup[oQ + 4u * par]);
    e0 = k.idx0 - (isum + ex);
    win = eq_window(eq, e0);
  }
#pragma unroll
  for (int h = 0; h < 2; h++) {
    V v[2][4];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int c = 2 * h + j;
      u32 mode = k.pmode;
      V pay = k.pp;
      const u32 bl = below(k.inb, c);
      if (k.pmode == 0 && !((k.inb >> c) & 1u)) {
        const bool e = eq_bit(win, e0, (u32)c - bl);
        mode = e ? 2u : 1u;
        pay = e ? k.dq[c] : k.dq[c] + sn.c[c];
      }
      if (mode == 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) v[j][i] = pay;
      } else {
        const Quad<V> q = reinterpret_cast<const Quad<V>*>(S.cells)[4u * par + (u32)c];
        if (mode == 2) {
#pragma unroll
          for (int i = 0; i < 4; i++) v[j][i] = pay + q.c[i];
        } else {
          V dd[4];
          dac_get4<V>(mx, nxt + 4u * (ex + bl), dd);  // leaves: max_t + max_s (log.rs:233,246)
#pragma unroll
          for (int i = 0; i < 4; i++) v[j][i] = dd[i] + q.c[i];
        }
      }
    }
    O.put_pair(fast, R0 + 2 * h, C0, v[0], v[1]);
  }
}

// One instant: `chunk` is the pointer that plays the role of the chunk start for the structure described by d
// (structure offsets are relative to it) -- the staging buffer, or global memory for structures that do not fit.
template <typename V>
DCDF_DEVINL void instant3(const u8* chunk, const InstDir& d, bool is_snap, int L, Tile3Smem<V>& S, const QuadOut& O) {
  if (is_snap) expand_snapshot3<V>(chunk, d, L, S, true);
  else log_top3<V>(chunk, d, L, S);
  final3<V>(chunk, d, L, S, O);
}
template <typename V>
__device__ __noinline__ void instant3_global(const u8* chunk, const InstDir* d, bool is_snap, int L, Tile3Smem<V>* S, const QuadOut* O) {
  instant3<V>(chunk, *d, is_snap, L, *S, *O);
}
template <typename V>
__device__ __noinline__ void snapshot3_global(const u8* chunk, const InstDir* d, int L, Tile3Smem<V>* S) {
  expand_snapshot3<V>(chunk, *d, L, *S, false);
}

// Start the copy of a structure and of its directory entry into staging half b.
template <typename V>
DCDF_DEVINL void prefetch3(const u8* chunk, const InstDir* dg, u32 off, u32 size, Tile3Smem<V>& S, int b) {
  const int tid = threadIdx.x;
  if (tid < W3_DIRW) cp_async4(reinterpret_cast<u32*>(&S.dir[b]) + tid, reinterpret_cast<const u32*>(dg) + tid);
  const u8* src = chunk + off;
  const u32 mis = (u32)((uintptr_t)src & 15u);
  if (size + mis + 4u > (u32)Tile3Smem<V>::BUF + 32u) return;
  const u8* g = src - mis;
  const u32 n16 = (size + mis + 4u + 15u) / 16u;  // +4: one word past the last byte may be read
  for (u32 i = tid; i < n16; i += DT_THREADS) cp_async16(S.stage[b] + 16u * i, g + 16u * i);
}
template <typename V>
DCDF_DEVINL bool staged3(const u8* chunk, const InstDir& d, u32& delta) {
  const u32 mis = (u32)((uintptr_t)(chunk + d.off) & 15u);
  delta = mis - d.off;
  return d.size + mis + 4u <= (u32)Tile3Smem<V>::BUF + 32u;
}

template <typename V>
__global__ void __launch_bounds__(DT_THREADS, sizeof(V) == 4 ? 4 : 2) k_window_tiles3(const TileWindowParams P) {
  extern __shared__ __align__(16) unsigned char dt3_smem_raw[];
  Tile3Smem<V>& S = *reinterpret_cast<Tile3Smem<V>*>(dt3_smem_raw);
  const QuerySet& Q = P.Q;
  const int tid = threadIdx.x;
  for (u64 ji = blockIdx.x; ji < P.n_jobs; ji += gridDim.x) {
    u64 lo_q = 0, hi_q = P.n_queries;
    while (hi_q - lo_q > 1) {
      const u64 mid = (lo_q + hi_q) >> 1;
      if (P.job_base[mid] <= ji) lo_q = mid; else hi_q = mid;
    }
    const u64 q = lo_q;
    const CubeDev c = P.cubes[q];
    const u64 local = ji - P.job_base[q];
    const i64 cs = Q.chunks_sidelen;
    const i64 cr0 = c.top / cs, cc0 = c.left / cs;
    const i64 ncr = (c.bottom - 1) / cs - cr0 + 1, ncc = (c.right - 1) / cs - cc0 + 1;
    const u64 nsub = (u64)(ncr * ncc);
    const u32 s = (u32)(c.start / Q.chunk_size) + (u32)(local / nsub);
    const u64 sub = local % nsub;
    const i64 cr = cr0 + (i64)(sub / (u64)ncc), cc = cc0 + (i64)(sub % (u64)ncc);
    const SliceMeta sm = Q.slices[s];
    const i64 t_lo = max(c.start, sm.t0), t_hi = min(c.end, sm.t0 + (i64)sm.instants);
    const i64 chunk_top = cr * cs, chunk_left = cc * cs;
    const i64 W_rows = c.bottom - c.top, W_cols = c.right - c.left;
    const u64 obase = P.out_off[q];
    const u32 slot = (u32)(cr * Q.subsidelen + cc);
    const int32_t u = Q.slot_unit[sm.slot_base + slot];
    const UnitMeta m = u >= 0 ? Q.units[u] : UnitMeta{};
    const bool stored = u >= 0 && m.stored;
    const i64 tile_org = (chunk_top - c.top) * W_cols + (chunk_left - c.left);  // element offset of tile cell (0, 0)
    QuadOut O;
    O.co.init(Q, P.out, P.raw, m.bits);
    O.pitch = W_cols;
    O.top = (int)(max(chunk_top, c.top) - chunk_top); O.bottom = (int)(min(chunk_top + cs, c.bottom) - chunk_top);
    O.left = (int)(max(chunk_left, c.left) - chunk_left); O.right = (int)(min(chunk_left + cs, c.right) - chunk_left);
    O.vec = O.co.kind == 2 && !(W_cols & 1) && !((obase + (u64)tile_org) & 1ull) && !((uintptr_t)P.out & 7u);
    O.vec4 = O.co.kind == 2 && !(W_cols & 3) && !((obase + (u64)tile_org) & 3ull) && !((uintptr_t)P.out & 15u);
    if (!stored) {
      // Elided: one value per instant from the max table, parent's fractional bits (superchunk.rs:426-433)
      const SlotDesc sdsc = Q.slot_desc[sm.slot_base + slot];
      const int wr = O.bottom - O.top, wc = O.right - O.left;
      for (i64 t = t_lo; t < t_hi; t++) {
        const i64 v = Q.tbl_max[sdsc.tbl0 + (u64)(t - sm.t0) * sdsc.stride];
        const u64 tb = obase + (u64)((t - c.start) * W_rows * W_cols + tile_org);
        for (int i = tid; i < wr * wc; i += DT_THREADS)
          emit(Q, P.out, tb + (u64)((i64)(O.top + i / wc) * W_cols + (O.left + i % wc)), v, sdsc.bits, P.raw);
      }
      continue;
    }
    const u8* chunk = Q.blob + m.blob_off;
    const InstDir* dir = Q.dir + m.dir_base;
    const int L = 31 - __clz(m.sidelen);
    if (t_hi <= t_lo) continue;
    const u32 ti0 = (u32)(t_lo - sm.t0), n_t = (u32)(t_hi - t_lo);
    __syncthreads();  // the previous job's readers are done with the staging buffers and the records
    const u32 snap0 = dir[ti0].snap;
    if (snap0 != ti0) {
      // the window starts inside a block: expand the block's snapshot first
      prefetch3<V>(chunk, dir + snap0, dir[snap0].off, dir[snap0].size, S, 1);
      cp_async_wait_all();
      __syncthreads();
      u32 delta;
      if (staged3<V>(chunk, S.dir[1], delta)) expand_snapshot3<V>(S.stage[1] + (int32_t)delta, S.dir[1], L, S, false);
      else snapshot3_global<V>(chunk, &S.dir[1], L, &S);
      __syncthreads();
    }
    prefetch3<V>(chunk, dir + ti0, dir[ti0].off, dir[ti0].size, S, 0);
    u32 noff = 0, nsize = 0;  // offset and size of the structure after the one being processed
    if (n_t > 1) { noff = dir[ti0 + 1].off; nsize = dir[ti0 + 1].size; }
    for (u32 i = 0; i < n_t; i++) {
      const int b = (int)(i & 1u);
      const u32 ti = ti0 + i;
      cp_async_wait_all();
      __syncthreads();  // structure i has landed; everyone is done with instant i-1 (the other half can be overwritten)
      if (i + 1 < n_t) {
        prefetch3<V>(chunk, dir + ti + 1, noff, nsize, S, b ^ 1);
        if (i + 2 < n_t) { noff = dir[ti + 2].off; nsize = dir[ti + 2].size; }
      }
      O.base = obase + (u64)((i64)(t_lo + i - c.start) * W_rows * W_cols + tile_org);
      const InstDir& D = S.dir[b];
      const bool is_snap = D.snap == ti;
      u32 delta;
      if (staged3<V>(chunk, D, delta)) {
        instant3<V>(S.stage[b] + (int32_t)delta, D, is_snap, L, S, O);
      } else {
        const QuadOut O2 = O;  // only the copy has its address taken
        instant3_global<V>(chunk, &D, is_snap, L, &S, &O2);
      }
    }
  }
}

}  // namespace dcdf
