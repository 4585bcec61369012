// decode_tile3.cuh -- window decode of whole <=64x64 tiles, second generation of decode_tile.cuh.
//
// Same contract as k_window_tiles (Snapshot::fill_window snapshot.rs:204-301, Log::fill_window log.rs:311-508, routed
// as Chunk::fill_window chunk.rs:152-158 / Superchunk::fill_window superchunk.rs:402-457), fewer instructions:
//   * every level of the snapshot pyramid is kept in Morton order, so the four cells of a quad are one 16-byte
//     shared-memory word and the four DAC entries of its children (BFS index 1 + 4 * rank .. + 3, snapshot.rs:177)
//     are fetched together: one 4-bit test of the continuation bitmap, four byte loads;
//   * the scan that turns "internal" flags into child BFS indices is fused into the level pass (flags stay in
//     registers between the ballot and the write), and the same scan yields rank1(idx), hence the position of a log
//     node's `equal` bit (rank0(idx + 1) - 1, log.rs:265) without touching the rank directory;
//   * the two top levels below the root are expanded by one warp without block barriers;
//   * the log's quad level is never written back: the thread that classified a quad emits its four cells
//     (two 8-byte stores for f32 output);
//   * snapshot and log share one staging buffer (the snapshot's bytes are dead once its pyramid is expanded).
#pragma once
#include "decode_tile.cuh"

namespace dcdf {

constexpr u32 W3_NONE = 0xffffffffu;

template <typename V>
struct Tile3Smem {
  static constexpr int STAGE = sizeof(V) == 4 ? 16 * 1024 : 12 * 1024;
  __align__(16) V cells[4096];   // snapshot values of the cells, Morton order (quad q = cells[4q .. 4q+3])
  V sup[DT_UPPER + 3];           // snapshot max values of the levels above the cells (level k at (4^k - 1) / 3)
  V pay[DT_UPPER + 3];           // log expansion payload
  u32 meta[DT_UPPER + 3];        // snapshot pass: BFS index of the first child or W3_NONE; log pass: (first child << 2) | mode
  u32 wtot[DT_WARPS];
  u32 nxt, isum, snap_single;
  InstDir dir;                   // directory entry of the structure being expanded
  __align__(16) u8 stage[STAGE + 32];
};

// Level 0 of a DAC (bytes + continuation bits) with the general accessor for longer codes.
struct Dac4 {
  const u8* bytes0;
  const u8* more0;
  DacRef slow;
  u32 len0;
};
DCDF_DEVINL Dac4 dac4_of(const u8* chunk, const DacDir* d) {
  const u32 len = d->len[0], base = d->base[0];
  const u32 words = base + 8u + 4u * (len / 128u);
  return Dac4{chunk + words + 4u * ((len + 31u) / 32u), chunk + words, DacRef{chunk, d}, d->n_levels ? len : 0u};
}
// codes longer than one byte: out of line, they are rare and the rank loop would be inlined a dozen times
__device__ __noinline__ i64 dac_slow_get(const u8* chunk, const DacDir* d, u32 idx) { return DacRef{chunk, d}.get(idx); }
template <typename V>
DCDF_DEVINL V unzz8(u32 b) { return (V)(int)((b >> 1) ^ (0u - (b & 1u))); }
template <typename V>
DCDF_DEVINL V dac_get1(const Dac4& m, u32 idx) {  // dac.rs:80-93
  if (idx >= m.len0) return (V)0;
  if (!((m.more0[idx >> 3] >> (7u - (idx & 7u))) & 1u)) return unzz8<V>(m.bytes0[idx]);
  return (V)dac_slow_get(m.slow.chunk, m.slow.d, idx);
}
// entries idx .. idx + 3 (the children of one node)
template <typename V>
DCDF_DEVINL void dac_get4(const Dac4& m, u32 idx, V (&d)[4]) {
  if (idx + 4u <= m.len0) {
    const u32 by = idx >> 3;
    const u32 hw = ((u32)m.more0[by] << 8) | (u32)m.more0[by + 1];  // the byte after the bitmap is the DAC's first code
    const u32 nib = (hw >> (12u - (idx & 7u))) & 15u;
    const u8* b = m.bytes0 + idx;
    if (nib == 0) {
#pragma unroll
      for (int i = 0; i < 4; i++) d[i] = unzz8<V>(b[i]);
      return;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) d[i] = (nib >> (3 - i)) & 1u ? (V)dac_slow_get(m.slow.chunk, m.slow.d, idx + (u32)i) : unzz8<V>(b[i]);
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) d[i] = dac_get1<V>(m, idx + (u32)i);
}
DCDF_DEVINL bool bit_at(const u8* bits, u32 i) { return (bits[i >> 3] >> (7u - (i & 7u))) & 1u; }
DCDF_DEVINL const u8* bitmap_bits(const u8* chunk, u32 len, u32 base) { return chunk + base + 8u + 4u * (len / 128u); }

// Positions of a level handled by this thread: BLOCK: warp w owns [w * seg, (w + 1) * seg), slot s covers 32 of them;
// otherwise (a level of at most 32 positions expanded by one warp) position = lane.
template <bool BLOCK>
struct LevelMap {
  u32 n1, seg, base;
  DCDF_DEVINL LevelMap(int lv) {
    n1 = 1u << (2 * lv);
    seg = BLOCK && n1 > 32u * DT_WARPS ? n1 / DT_WARPS : 32u;
    base = (BLOCK ? (threadIdx.x >> 5) * seg : 0u) + (threadIdx.x & 31u);
  }
  DCDF_DEVINL u32 pos(int s) const { return base + 32u * (u32)s; }
  DCDF_DEVINL bool valid(int s) const { return 32u * (u32)s < seg && pos(s) < n1; }
};

// Exclusive prefix of the flags over the level's positions (Morton order == BFS order among existing nodes).
// ex[s] = number of set flags before this thread's position of slot s; returns the level total.
template <bool BLOCK, int SL>
DCDF_DEVINL u32 level_scan(const u32 (&bal)[SL], u32 (&ex)[SL], u32* wtot) {
  const u32 lt = lanemask_lt();
  u32 cnt = 0;
#pragma unroll
  for (int s = 0; s < SL; s++) { ex[s] = cnt + __popc(bal[s] & lt); cnt += __popc(bal[s]); }
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
#pragma unroll
  for (int s = 0; s < SL; s++) ex[s] += before;
  return total;
}

// Snapshot: produce level lv (1 .. L-1) from level lv-1.  nxt = BFS index of the first node of level lv+1.
template <typename V, bool BLOCK>
DCDF_DEVINL u32 snap_level(Tile3Smem<V>& S, const Dac4& mx, const u8* nm, u32 nm_len, int lv, u32 nxt) {
  constexpr int SL = BLOCK ? 4 : 1;
  const LevelMap<BLOCK> M(lv);
  const u32 o0 = lvl_off(lv - 1), o1 = lvl_off(lv);
  u32 bal[SL], ex[SL];
  V val[SL];
#pragma unroll
  for (int s = 0; s < SL; s++) {
    bool in = false;
    V v = 0;
    if (M.valid(s)) {
      const u32 p1 = M.pos(s), par = p1 >> 2;
      const u32 pm = S.meta[o0 + par];
      v = S.sup[o0 + par];
      if (pm != W3_NONE) {
        const u32 idx = pm + (p1 & 3u);
        v -= dac_get1<V>(mx, idx);  // snapshot.rs:179
        in = idx < nm_len && bit_at(nm, idx);
      }
    }
    val[s] = v;
    bal[s] = __ballot_sync(0xffffffffu, in);
  }
  const u32 total = level_scan<BLOCK, SL>(bal, ex, S.wtot);
  const u32 lane = threadIdx.x & 31u;
#pragma unroll
  for (int s = 0; s < SL; s++) {
    if (M.valid(s)) {
      const u32 p1 = M.pos(s);
      S.sup[o1 + p1] = val[s];
      S.meta[o1 + p1] = (bal[s] >> lane) & 1u ? nxt + 4u * ex[s] : W3_NONE;
    }
  }
  return total;
}

template <typename V>
struct __align__(16) Quad { V c[4]; };

// Expand the Snapshot whose bytes start at `chunk + d.off` (staged) into S.sup / S.cells.  Ends with a barrier.
template <typename V>
DCDF_DEVINL void expand_snapshot3(const u8* chunk, int L, Tile3Smem<V>& S) {
  const int tid = threadIdx.x;
  const InstDir& d = S.dir;
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
    for (int lv = 1; lv <= 2 && lv <= L - 1; lv++) {
      nxt += 4u * snap_level<V, false>(S, mx, nm, nm_len, lv, nxt);
      __syncwarp();
    }
    if (tid == 0) S.nxt = nxt;
  }
  __syncthreads();
  u32 nxt = S.nxt;
  for (int lv = 3; lv <= L - 1; lv++) {
    nxt += 4u * snap_level<V, true>(S, mx, nm, nm_len, lv, nxt);
    __syncthreads();
  }
  // the cells, one quad (a node of level L-1) at a time
  const u32 n0 = 1u << (2 * (L - 1)), oP = lvl_off(L - 1);
  for (u32 p = tid; p < n0; p += DT_THREADS) {
    const u32 pm = S.meta[oP + p];
    const V pv = S.sup[oP + p];
    Quad<V> q;
    if (pm != W3_NONE) {
      V dd[4];
      dac_get4<V>(mx, pm, dd);
#pragma unroll
      for (int i = 0; i < 4; i++) q.c[i] = pv - dd[i];
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) q.c[i] = pv;
    }
    reinterpret_cast<Quad<V>*>(S.cells)[p] = q;
  }
  __syncthreads();
}

// Where the cells of the current (window, tile, instant) go.
struct QuadOut {
  CellOut co;
  u64 base;       // element index of tile cell (0, 0) at this instant
  i64 pitch;      // window columns
  int top, bottom, left, right;  // window clipped to the tile, tile coordinates
  bool vec;       // f32 output whose row pairs are 8-byte aligned
  template <typename V>
  DCDF_DEVINL float cvt(V v) const {  // from_fixed (fixed.rs:81-86): 0 -> NaN, else (v - 1) * 2^-(bits+1), exact scaling
    const float f = (sizeof(V) == 4 ? __int2float_rn((int)v - 1) : __ll2float_rn((i64)v - 1)) * co.inv32;
    return v == 0 ? __int_as_float(0x7fc00000) : f;
  }
  DCDF_DEVINL bool touches(int r0, int c0) const { return r0 + 1 >= top && r0 < bottom && c0 + 1 >= left && c0 < right; }
  template <typename V>
  DCDF_DEVINL void put(int r0, int c0, const V (&v)[4]) const {
    const u64 i00 = base + (u64)((i64)r0 * pitch + c0);
    const bool inside = r0 >= top && r0 + 1 < bottom && c0 >= left && c0 + 1 < right;
    if (inside && vec) {
      float* o = static_cast<float*>(co.out) + i00;
      *reinterpret_cast<float2*>(o) = make_float2(cvt(v[0]), cvt(v[1]));
      *reinterpret_cast<float2*>(o + pitch) = make_float2(cvt(v[2]), cvt(v[3]));
      return;
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int r = r0 + (c >> 1), col = c0 + (c & 1);
      if (!inside && (r < top || r >= bottom || col < left || col >= right)) continue;
      co.put(i00 + (u64)((c >> 1) ? pitch : 0) + (u64)(c & 1), v[c]);
    }
  }
};

// A Snapshot instant inside the window: the cells are the pyramid's last level.
template <typename V>
DCDF_DEVINL void emit_snapshot3(int L, const Tile3Smem<V>& S, const QuadOut& O) {
  const u32 n0 = 1u << (2 * (L - 1));
  for (u32 p = threadIdx.x; p < n0; p += DT_THREADS) {
    const int r0 = 2 * (int)morton_row(p), c0 = 2 * (int)morton_col(p);
    if (!O.touches(r0, c0)) continue;
    const Quad<V> q = reinterpret_cast<const Quad<V>*>(S.cells)[p];
    O.put(r0, c0, q.c);
  }
}

// Log: one level of the expansion (log.rs:207-293).  Per position: mode 0 internal (payload = max_t, replaced at every
// level, log.rs:233), 1 uniform (payload = value), 2 equal (value = payload + snapshot cell).  isum = internal nodes of
// the levels above, so that rank1(idx) = isum + (internal nodes of this level before idx).
template <typename V>
struct LogNode {
  u32 mode, idx;
  V pay;
};
template <typename V>
DCDF_DEVINL void log_derive(const Tile3Smem<V>& S, const Dac4& mx, const u8* nm, u32 nm_len, u32 o0, u32 p1, LogNode<V>& n, bool& in) {
  const u32 par = p1 >> 2;
  const u32 pm = S.meta[o0 + par];
  n.pay = S.pay[o0 + par];
  n.mode = pm & 3u;
  n.idx = 0;
  in = false;
  if (n.mode == 0) {
    n.idx = (pm >> 2) + (p1 & 3u);
    n.pay = dac_get1<V>(mx, n.idx);
    in = n.idx < nm_len && bit_at(nm, n.idx);
  }
}
// after the scan: a node that stops here is `equal` (snapshot + constant) or uniform (max_t + max_s of the node)
template <typename V>
DCDF_DEVINL void log_settle(const u8* eq, LogNode<V>& n, bool in, u32 rank1, V snap_node) {
  if (n.mode == 0 && !in) {
    const bool e = bit_at(eq, n.idx - rank1);  // rank0(idx + 1) - 1 (log.rs:265)
    n.mode = e ? 2u : 1u;
    if (!e) n.pay += snap_node;                // log.rs:266-268
  }
}

template <typename V, bool BLOCK>
DCDF_DEVINL u32 log_level(Tile3Smem<V>& S, const Dac4& mx, const u8* nm, u32 nm_len, const u8* eq, int lv, u32 nxt, u32 isum) {
  constexpr int SL = BLOCK ? 4 : 1;
  const LevelMap<BLOCK> M(lv);
  const u32 o0 = lvl_off(lv - 1), o1 = lvl_off(lv);
  u32 bal[SL], ex[SL];
  LogNode<V> nd[SL];
#pragma unroll
  for (int s = 0; s < SL; s++) {
    bool in = false;
    nd[s].mode = 3; nd[s].idx = 0; nd[s].pay = 0;
    if (M.valid(s)) log_derive<V>(S, mx, nm, nm_len, o0, M.pos(s), nd[s], in);
    bal[s] = __ballot_sync(0xffffffffu, in);
  }
  const u32 total = level_scan<BLOCK, SL>(bal, ex, S.wtot);
  const u32 lane = threadIdx.x & 31u;
#pragma unroll
  for (int s = 0; s < SL; s++) {
    if (M.valid(s)) {
      const u32 p1 = M.pos(s);
      const bool in = (bal[s] >> lane) & 1u;
      if (nd[s].mode == 0 && !in) log_settle<V>(eq, nd[s], in, isum + ex[s], S.sup[o1 + p1]);
      S.pay[o1 + p1] = nd[s].pay;
      S.meta[o1 + p1] = in ? (nxt + 4u * ex[s]) << 2 : nd[s].mode;
    }
  }
  return total;
}

// Expand the staged Log against the snapshot pyramid in S and write the window's cells of this instant.
template <typename V>
DCDF_DEVINL void expand_log3(const u8* chunk, int L, Tile3Smem<V>& S, const QuadOut& O) {
  const int tid = threadIdx.x;
  const InstDir& d = S.dir;
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
    for (int lv = 1; lv <= 2 && lv <= L - 2; lv++) {
      const u32 tot = log_level<V, false>(S, mx, nm, nm_len, eq, lv, nxt, isum);
      isum += tot; nxt += 4u * tot;
      __syncwarp();
    }
    if (tid == 0) { S.nxt = nxt; S.isum = isum; }
  }
  __syncthreads();
  u32 nxt = S.nxt, isum = S.isum;
  for (int lv = 3; lv <= L - 2; lv++) {
    const u32 tot = log_level<V, true>(S, mx, nm, nm_len, eq, lv, nxt, isum);
    isum += tot; nxt += 4u * tot;
    __syncthreads();
  }
  // the quads (level L-1): classified in registers, their cells written straight to the output
  constexpr int SL = 4;
  const int lvq = L - 1;
  const LevelMap<true> M(lvq);
  u32 bal[SL], ex[SL];
  LogNode<V> nd[SL];
  if (lvq == 0) {
    // a 2x2 tile: the root is the quad
#pragma unroll
    for (int s = 0; s < SL; s++) { nd[s].mode = 3; nd[s].idx = 0; nd[s].pay = 0; bal[s] = 0; ex[s] = 0; }
    if (tid == 0) {
      const u32 pm = S.meta[0];
      nd[0].mode = pm & 3u; nd[0].pay = S.pay[0];
      if (nd[0].mode == 0) bal[0] = 1u;  // children at BFS index nxt = 1
    }
  } else {
    const u32 o0 = lvl_off(lvq - 1);
#pragma unroll
    for (int s = 0; s < SL; s++) {
      bool in = false;
      nd[s].mode = 3; nd[s].idx = 0; nd[s].pay = 0;
      if (M.valid(s)) log_derive<V>(S, mx, nm, nm_len, o0, M.pos(s), nd[s], in);
      bal[s] = __ballot_sync(0xffffffffu, in);
    }
    level_scan<true, SL>(bal, ex, S.wtot);
  }
  const u32 lane = tid & 31u, oQ = lvl_off(lvq);
#pragma unroll
  for (int s = 0; s < SL; s++) {
    if (!M.valid(s)) continue;
    const u32 p1 = M.pos(s);
    const int r0 = 2 * (int)morton_row(p1), c0 = 2 * (int)morton_col(p1);
    if (!O.touches(r0, c0)) continue;
    const bool in = (bal[s] >> lane) & 1u;
    if (lvq > 0 && nd[s].mode == 0 && !in) log_settle<V>(eq, nd[s], in, isum + ex[s], S.sup[oQ + p1]);
    V v[4];
    if (nd[s].mode == 1) {
#pragma unroll
      for (int i = 0; i < 4; i++) v[i] = nd[s].pay;
    } else {
      const Quad<V> q = reinterpret_cast<const Quad<V>*>(S.cells)[p1];
      if (nd[s].mode == 2) {
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = nd[s].pay + q.c[i];
      } else {
        V dd[4];
        dac_get4<V>(mx, (lvq ? nxt : 1u) + 4u * ex[s], dd);  // leaves: max_t + max_s (log.rs:233,246)
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = dd[i] + q.c[i];
      }
    }
    O.put(r0, c0, v);
  }
}

// Copy a structure into the staging buffer (same address modulo 16); returns the pointer that plays the role of the
// chunk start (structure offsets are relative to it), or the global chunk pointer when the structure does not fit.
template <typename V>
DCDF_DEVINL bool stage3(const u8* chunk, const InstDir* dg, Tile3Smem<V>& S, u32& delta) {
  const int tid = threadIdx.x;
  if (tid < (int)(sizeof(InstDir) / 4)) reinterpret_cast<u32*>(&S.dir)[tid] = reinterpret_cast<const u32*>(dg)[tid];
  const u32 off = dg->off, size = dg->size;
  const u8* src = chunk + off;
  const u32 mis = (u32)((uintptr_t)src & 15u);
  delta = mis - off;  // staged "chunk start" = S.stage + delta (modular arithmetic)
  if (size + mis + 4u > (u32)Tile3Smem<V>::STAGE + 32u) return false;
  const uint4* g = reinterpret_cast<const uint4*>(src - mis);
  uint4* dst = reinterpret_cast<uint4*>(S.stage);
  const u32 n16 = (size + mis + 4u + 15u) / 16u;
  for (u32 i = tid; i < n16; i += DT_THREADS) dst[i] = g[i];
  return true;
}
// structures that do not fit the staging buffer are read from global memory by out-of-line copies of the same code
template <typename V>
__device__ __noinline__ void expand_snapshot3_global(const u8* chunk, int L, Tile3Smem<V>* S) { expand_snapshot3<V>(chunk, L, *S); }
template <typename V>
__device__ __noinline__ void expand_log3_global(const u8* chunk, int L, Tile3Smem<V>* S, const QuadOut* O) { expand_log3<V>(chunk, L, *S, *O); }

template <typename V>
__global__ void __launch_bounds__(DT_THREADS, sizeof(V) == 4 ? 4 : 3) k_window_tiles3(const TileWindowParams P) {
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
    u32 cur_snap = 0xffffffffu;
    for (i64 t = t_lo; t < t_hi; t++) {
      const u32 ti = (u32)(t - sm.t0);
      const u32 snap = dir[ti].snap;
      O.base = obase + (u64)((t - c.start) * W_rows * W_cols + tile_org);
      __syncthreads();  // the previous instant's readers are done with the staging buffer and the records
      if (snap != cur_snap) {
        u32 delta;
        const bool staged = stage3<V>(chunk, dir + snap, S, delta);
        __syncthreads();
        if (staged) expand_snapshot3<V>(S.stage + (int32_t)delta, L, S);
        else expand_snapshot3_global<V>(chunk, L, &S);
        cur_snap = snap;
        if (snap == ti) emit_snapshot3<V>(L, S, O);
      } else if (snap == ti) {
        emit_snapshot3<V>(L, S, O);
      }
      if (snap != ti) {
        u32 delta;
        const bool staged = stage3<V>(chunk, dir + ti, S, delta);
        __syncthreads();
        if (staged) expand_log3<V>(S.stage + (int32_t)delta, L, S, O);
        else {
          const QuadOut O2 = O;  // only the copy has its address taken
          expand_log3_global<V>(chunk, L, &S, &O2);
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace dcdf
