// decode_tile2.cuh -- window decode of <=64x64 tiles, one thread per 8x8 block of cells.
//
// Same job decomposition as decode_tile.cuh (one CTA per (window, time slice, subchunk), instants in order), but
// instead of expanding every level of the tree densely through shared memory with a barrier per level, each of
// the 64 threads walks the Snapshot and the Log of the instant in lock step down to its own 8x8 block, the way
// Snapshot::fill_window (snapshot.rs:204-301) and Log::fill_window (log.rs:311-508) recurse, but for a whole
// sub-tree at once:
//   * child positions are `1 + rank(index) * k^2` (snapshot.rs:177) with ONE rank-directory lookup per sibling
//     group (the ranks of the four siblings follow from their nodemap nibble);
//   * DAC entries of a sibling group are four consecutive bytes plus a continuation nibble (dac.rs:80-93), the
//     multi-byte path is taken only when a continuation bit is set;
//   * sub-trees that do not intersect the window are skipped, so small windows cost what they touch.
// The structure bytes of the instant are staged in shared memory by the whole CTA first (one barrier); nothing
// else is shared, so an instant needs two barriers instead of ~14 and no 100 KB pyramid.
#pragma once
#include "decode_tile.cuh"

namespace dcdf {

constexpr int DW_THREADS = 64;
constexpr int DW_STAGE = 14 * 1024;

struct Tile2Smem {
  InstDir dir_s, dir_l;
  __align__(16) u8 stage_s[DW_STAGE + 32];
  __align__(16) u8 stage_l[DW_STAGE + 32];
};

// Four consecutive bits of an MSB-first bit stream, first one in bit 3 (reads one byte past the last bit's byte).
DCDF_DEVINL u32 bits4_at(const u8* bits, u32 i) {
  const u32 w = ((u32)bits[i >> 3] << 8) | (u32)bits[(i >> 3) + 1];
  return (w >> (12u - (i & 7u))) & 0xfu;
}

// Out-of-line slow paths (kept out of the walker's loops: the kernel is instruction-cache sensitive).
__device__ __noinline__ i64 dw_dac_get(const u8* chunk, const DacDir* d, u32 idx) { return DacRef{chunk, d}.get(idx); }
__device__ __noinline__ u32 dw_rank(const u8* chunk, u32 len, u32 base, u32 i) { return BitMapRef{chunk, len, base}.rank(i); }

// One serialized Snapshot / Log as the walker needs it.
struct TreeRef {
  const u8* chunk;
  const InstDir* d;
  const u8* nm_bits;    // nodemap words as an MSB-first byte stream
  const u8* x_bytes0;   // level-0 bytes of the max DAC
  const u8* x_more0;    // their continuation bits
  u32 nm_len, n_max;
  DCDF_DEVINL void init(const u8* chunk_, const InstDir* d_) {
    chunk = chunk_; d = d_;
    nm_len = d_->nm_len;
    nm_bits = chunk_ + d_->nm_base + 8u + 4u * (nm_len / 128u);
    n_max = d_->max.n_levels ? d_->max.len[0] : 0u;
    const u32 words = d_->max.base[0] + 8u + 4u * (n_max / 128u);
    x_more0 = chunk_ + words;
    x_bytes0 = chunk_ + words + 4u * ((n_max + 31u) / 32u);
  }
  DCDF_DEVINL u32 rank(u32 i) const { return dw_rank(chunk, nm_len, d->nm_base, i); }
  DCDF_DEVINL bool internal(u32 idx) const { return idx < nm_len && ((nm_bits[idx >> 3] >> (7u - (idx & 7u))) & 1u); }
  DCDF_DEVINL i64 entry(u32 idx) const {  // max[idx]  (dac.rs:80-93)
    if (idx >= n_max) return 0;
    if (!((x_more0[idx >> 3] >> (7u - (idx & 7u))) & 1u)) return unzigzag64((u64)x_bytes0[idx]);
    return dw_dac_get(chunk, &d->max, idx);
  }
  // The level-0 bytes of the four sibling entries idx..idx+3 (little-endian in the result) and their continuation
  // bits (sibling c = bit 3-c).  False when the group is not entirely inside the DAC (malformed input guard).
  DCDF_DEVINL bool group(u32 idx, u32& bytes, u32& more) const {
    if (idx + 4u > n_max) return false;
    const uintptr_t a = (uintptr_t)(x_bytes0 + idx);
    const u32* w = reinterpret_cast<const u32*>(a & ~(uintptr_t)3);
    const u32 sh = (u32)(a & 3u) * 8u;
    const u32 lo = w[0];
    bytes = sh ? __funnelshift_r(lo, w[1], sh) : lo;
    more = bits4_at(x_more0, idx);
    return true;
  }
  // Entry idx + c of a fetched group.
  DCDF_DEVINL i64 group_entry(u32 idx, int c, u32 bytes, u32 more) const {
    if ((more >> (3 - c)) & 1u) return dw_dac_get(chunk, &d->max, idx + (u32)c);
    const int b = (int)((bytes >> (8 * c)) & 0xffu);
    return (i64)((b >> 1) ^ -(b & 1));  // zigzag decode of a one-byte code (dac.rs:134-137)
  }
};

// State of the lock-step walk at one tree node.
struct WalkNode {
  i64 sv;     // snapshot: value of the node (or of the ancestor it stopped at)
  u32 scb;    // snapshot: BFS index of the first child, 0 if the node is not an internal node
  i64 lp;     // log payload
  u32 lcb;    // log: BFS index of the first child when lmode == 0
  u32 lmode;  // log: 0 internal, 1 uniform (value = payload), 2 equal (value = payload + snapshot cell)
};

// The four children of a node: nodemap nibbles and the ranks that give their own child bases.
struct Group {
  u32 s_nib, s_rank;  // snapshot: nodemap bits of the children (child c = bit 3-c), rank(scb)
  u32 l_nib, l_rank;
  u32 s_bytes, s_more, l_bytes, l_more;  // DAC entries of the four children (TreeRef::group)
  bool s_ok, l_ok;
};

template <bool LOG>
DCDF_DEVINL Group group_of(const WalkNode& n, const TreeRef& S, const TreeRef& Lg, bool children_have_bits) {
  Group g;
  g.s_nib = g.s_rank = g.l_nib = g.l_rank = 0;
  g.s_bytes = g.s_more = g.l_bytes = g.l_more = 0;
  g.s_ok = n.scb && S.group(n.scb, g.s_bytes, g.s_more);
  g.l_ok = LOG && n.lmode == 0 && Lg.group(n.lcb, g.l_bytes, g.l_more);
  if (n.scb && children_have_bits && n.scb < S.nm_len) {
    g.s_nib = bits4_at(S.nm_bits, n.scb);
    if (n.scb + 4u > S.nm_len) g.s_nib &= 0xfu << (n.scb + 4u - S.nm_len);
    if (g.s_nib) g.s_rank = S.rank(n.scb);
  }
  if (LOG && n.lmode == 0 && children_have_bits && n.lcb < Lg.nm_len) {
    g.l_nib = bits4_at(Lg.nm_bits, n.lcb);
    if (n.lcb + 4u > Lg.nm_len) g.l_nib &= 0xfu << (n.lcb + 4u - Lg.nm_len);
    g.l_rank = Lg.rank(n.lcb);  // also needed for the equal bits of children that stop here
  }
  return g;
}

// Child c of node n (log.rs:207-293 / snapshot.rs:165-188).  `child_is_cell`: the child is on the last tree level.
template <bool LOG>
DCDF_DEVINL WalkNode child_of(const WalkNode& n, const Group& g, int c, const TreeRef& S, const TreeRef& Lg, const BitMapRef& eq,
                              bool child_is_cell) {
  WalkNode ch;
  ch.sv = n.sv; ch.scb = 0;
  if (n.scb) {
    const u32 idx = n.scb + (u32)c;
    ch.sv = n.sv - (g.s_ok ? S.group_entry(n.scb, c, g.s_bytes, g.s_more) : S.entry(idx));  // snapshot.rs:179
    if ((g.s_nib >> (3 - c)) & 1u) ch.scb = 1u + 4u * (g.s_rank + (u32)__popc(g.s_nib >> (4 - c)));
  }
  ch.lp = n.lp; ch.lcb = 0; ch.lmode = n.lmode;
  if (LOG && n.lmode == 0) {
    const u32 idx = n.lcb + (u32)c;
    const i64 d = g.l_ok ? Lg.group_entry(n.lcb, c, g.l_bytes, g.l_more) : Lg.entry(idx);  // max_t is replaced, not accumulated (log.rs:233)
    const u32 ones = g.l_rank + (u32)__popc(g.l_nib >> (4 - c));  // rank(idx)
    if ((g.l_nib >> (3 - c)) & 1u) {
      ch.lmode = 0; ch.lp = d; ch.lcb = 1u + 4u * ones;
    } else if (child_is_cell) {
      ch.lmode = 2; ch.lp = d;  // value = max_t + snapshot cell
    } else {
      const bool e = eq.get(idx - ones);  // rank0(idx + 1) - 1 (log.rs:265)
      ch.lmode = e ? 2u : 1u;
      ch.lp = e ? d : d + ch.sv;  // uniform: max_t + max_s of this node (log.rs:266-268)
    }
  }
  return ch;
}

template <bool LOG>
DCDF_DEVINL i64 cell_value(const WalkNode& n) {
  if (!LOG) return n.sv;
  return n.lmode == 1 ? n.lp : n.lp + n.sv;
}

// One instant of one tile: every thread walks down to its 8x8 block and writes the cells inside the window.
// L = tree levels (sidelen = 2^L), the tree root is the node (lo = 6 - L, 0) of the 64x64 frame.
template <bool LOG, typename Emit>
DCDF_DEVINL void walk_tile(const TreeRef& S, const TreeRef& Lg, const BitMapRef& eq, int L, int top, int bottom, int left, int right,
                           Emit emit_cell) {
  const int tid = threadIdx.x;
  const int lo = 6 - L;
  const int r0 = 8 * (int)morton_row((u32)tid), c0 = 8 * (int)morton_col((u32)tid);
  // in the tree, and touching the window?
  if (lo <= 3 ? (tid >> (2 * (3 - lo))) != 0 : tid != 0) return;
  if (r0 >= bottom || r0 + 8 <= top || c0 >= right || c0 + 8 <= left) return;

  // root
  WalkNode n;
  n.sv = S.entry(0);
  n.scb = (L > 0 && S.internal(0)) ? 1u : 0u;
  n.lp = 0; n.lcb = 0; n.lmode = 1;
  if (LOG) {
    const i64 d0 = Lg.entry(0);
    if (L > 0 && Lg.internal(0)) {
      n.lmode = 0; n.lp = d0; n.lcb = 1;
    } else {
      // log.rs:180-186: a single-node log is uniform unless its equal bit says "snapshot + constant"
      const bool uniform = !(L > 0 && S.internal(0)) || !eq.get(0);
      n.lmode = uniform ? 1u : 2u;
      n.lp = uniform ? d0 + n.sv : d0;
    }
  }
  // path from the root to the thread's level-3 node
  for (int k = lo; k < 3; k++) {
    const int c = (tid >> (2 * (2 - k))) & 3;
    const Group g = group_of<LOG>(n, S, Lg, k + 1 < 6);
    n = child_of<LOG>(n, g, c, S, Lg, eq, false);
  }
  // the 8x8 block: frame levels 4 (a), 5 (b), 6 (cells); levels above the root pass the state through child 0
  const bool kids3 = lo <= 3 && (n.scb || (LOG && n.lmode == 0));
  Group g3 = Group{0, 0, 0, 0, 0, 0, 0, 0, false, false};
  if (kids3) g3 = group_of<LOG>(n, S, Lg, true);
#pragma unroll 1
  for (int a = 0; a < 4; a++) {
    const int ra = r0 + 4 * (a >> 1), ca = c0 + 4 * (a & 1);
    WalkNode n4 = n;
    if (lo <= 3) {
      if (ra >= bottom || ra + 4 <= top || ca >= right || ca + 4 <= left) continue;
      if (kids3) n4 = child_of<LOG>(n, g3, a, S, Lg, eq, false);
    } else if (a != 0) {
      continue;
    }
    const bool kids4 = lo <= 4 && (n4.scb || (LOG && n4.lmode == 0));
    Group g4 = Group{0, 0, 0, 0, 0, 0, 0, 0, false, false};
    if (kids4) g4 = group_of<LOG>(n4, S, Lg, true);
#pragma unroll 1
    for (int b = 0; b < 4; b++) {
      const int rb = ra + 2 * (b >> 1), cb = ca + 2 * (b & 1);
      WalkNode n5 = n4;
      if (lo <= 4) {
        if (rb >= bottom || rb + 2 <= top || cb >= right || cb + 2 <= left) continue;
        if (kids4) n5 = child_of<LOG>(n4, g4, b, S, Lg, eq, false);
      } else if (b != 0) {
        continue;
      }
      const bool kids5 = lo <= 5 && (n5.scb || (LOG && n5.lmode == 0));
      Group g6;
      g6.s_nib = g6.s_rank = g6.l_nib = g6.l_rank = 0;
      g6.s_bytes = g6.s_more = g6.l_bytes = g6.l_more = 0;
      g6.s_ok = kids5 && n5.scb && S.group(n5.scb, g6.s_bytes, g6.s_more);
      g6.l_ok = LOG && kids5 && n5.lmode == 0 && Lg.group(n5.lcb, g6.l_bytes, g6.l_more);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int r = rb + (c >> 1), col = cb + (c & 1);
        WalkNode n6 = n5;
        if (lo <= 5) {
          if (r < top || r >= bottom || col < left || col >= right) continue;
          if (kids5) n6 = child_of<LOG>(n5, g6, c, S, Lg, eq, true);
        } else if (c != 0) {
          continue;
        }
        emit_cell(r, col, cell_value<LOG>(n6));
      }
    }
  }
}

__global__ void __launch_bounds__(DW_THREADS) k_window_tiles2(const TileWindowParams P) {
  __shared__ Tile2Smem S;
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
    const int top = (int)(max(chunk_top, c.top) - chunk_top), bottom = (int)(min(chunk_top + cs, c.bottom) - chunk_top);
    const int left = (int)(max(chunk_left, c.left) - chunk_left), right = (int)(min(chunk_left + cs, c.right) - chunk_left);
    const int wr = bottom - top, wc = right - left;
    const i64 W_rows = c.bottom - c.top, W_cols = c.right - c.left;
    const u64 obase = P.out_off[q];
    const u32 slot = (u32)(cr * Q.subsidelen + cc);
    const int32_t u = Q.slot_unit[sm.slot_base + slot];
    const UnitMeta m = u >= 0 ? Q.units[u] : UnitMeta{};
    const bool stored = u >= 0 && m.stored;
    // element offset of tile cell (0, 0) at instant t:  obase + (t - start) * W_rows * W_cols + tile_org
    const i64 tile_org = (chunk_top - c.top) * W_cols + (chunk_left - c.left);
    if (!stored) {
      // Elided: one value per instant from the max table, parent's fractional bits (superchunk.rs:426-433)
      const SlotDesc sdsc = Q.slot_desc[sm.slot_base + slot];
      for (i64 t = t_lo; t < t_hi; t++) {
        const i64 v = Q.tbl_max[sdsc.tbl0 + (u64)(t - sm.t0) * sdsc.stride];
        const u64 ob = obase + (u64)((t - c.start) * W_rows * W_cols + tile_org);
        for (int i = tid; i < wr * wc; i += DW_THREADS) emit(Q, P.out, ob + (u64)((top + i / wc) * W_cols + (left + i % wc)), v, sdsc.bits, P.raw);
      }
      continue;
    }
    const u8* chunk = Q.blob + m.blob_off;
    const InstDir* dir = Q.dir + m.dir_base;
    const int L = 31 - __clz(m.sidelen);
    u32 cur_snap = 0xffffffffu;
    const u8* chunk_s = chunk;  // chunk pointers rebased so that `ptr + structure offset` lands in the staged copy
    for (i64 t = t_lo; t < t_hi; t++) {
      const u32 ti = (u32)(t - sm.t0);
      const u32 snap = dir[ti].snap;
      const bool is_log = snap != ti;
      __syncthreads();  // previous instant done with the staged bytes
      if (snap != cur_snap) {
        if (tid == 0) S.dir_s = dir[snap];
        const u32 off = dir[snap].off;
        chunk_s = stage_bytes(chunk + off, dir[snap].size, S.stage_s, DW_STAGE + 32, DW_THREADS) - off;
        cur_snap = snap;
      }
      const u8* chunk_l = chunk;
      if (is_log) {
        if (tid == 0) S.dir_l = dir[ti];
        const u32 off = dir[ti].off;
        chunk_l = stage_bytes(chunk + off, dir[ti].size, S.stage_l, DW_STAGE + 32, DW_THREADS) - off;
      }
      __syncthreads();
      TreeRef TS, TL;
      TS.init(chunk_s, &S.dir_s);
      const u64 ob = obase + (u64)((t - c.start) * W_rows * W_cols + tile_org);
      CellOut co;
      co.init(Q, P.out, P.raw, m.bits);
      auto emit_cell = [&](int r, int col, i64 v) { co.put(ob + (u64)((i64)r * W_cols + col), v); };
      if (is_log) {
        TL.init(chunk_l, &S.dir_l);
        const BitMapRef eq{chunk_l, S.dir_l.eq_len, S.dir_l.eq_base};
        walk_tile<true>(TS, TL, eq, L, top, bottom, left, right, emit_cell);
      } else {
        const BitMapRef eq{chunk_s, 0, 0};
        walk_tile<false>(TS, TS, eq, L, top, bottom, left, right, emit_cell);
      }
    }
    __syncthreads();
  }
}

}  // namespace dcdf
