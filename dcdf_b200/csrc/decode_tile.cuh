// decode_tile.cuh -- window decode by level-synchronous expansion of whole <=64x64 tiles.
//
// Replaces the per-cell recursion of Snapshot::fill_window (snapshot.rs:204-301) and Log::fill_window
// (log.rs:311-508) for batched windows: one CTA per (window, time slice, subchunk) walks the requested
// instants of that chunk in order.  For a Snapshot the k^2 tree is expanded top-down into a dense
// max-value pyramid in shared memory (every level, Morton order): positions whose parent is an internal
// node read their own DAC entry (value = parent - max[idx], snapshot.rs:179), all others inherit the
// parent's value.  BFS indices of children come from a per-level ballot/popc scan of the internal flags,
// i.e. the same arithmetic as `1 + rank(index) * k^2` (snapshot.rs:177) without touching the rank
// directory.  A Log is expanded the same way against its block's snapshot pyramid (log.rs:207-293):
// a log node that stops with equal = 0 is uniform (max_t + max_s of that node), one that stops with
// equal = 1 adds its offset to the snapshot's cells below it.
#pragma once
#include "decode.cuh"

namespace dcdf {

constexpr int DT_THREADS = 256;
constexpr int DT_WARPS = DT_THREADS / 32;
constexpr int DT_NODES = 5461;
constexpr int DT_UPPER = 1365;
constexpr int DT_STAGE_S = 12 * 1024;   // shared-memory copy of the block's Snapshot bytes (when it fits)
constexpr int DT_STAGE_L = 8 * 1024;   // ... and of the current Log

// V = int32_t when every max DAC of the handle has at most three byte levels (entries below 2^23 in magnitude, so no
// sum along a root-to-cell path of <= 7 nodes leaves 32 bits), i64 otherwise; the narrow layout lets four CTAs share
// an SM instead of two.
template <typename V>
struct TileSmemT {
  V sval[DT_NODES];   // snapshot max pyramid: level k at offset (4^k - 1) / 3, Morton order inside a level
  V lpay[DT_UPPER + 3];   // log expansion payload (levels above the cells)
  unsigned short scb[DT_UPPER + 3];  // snapshot: BFS index of the first child (0xffff: not an internal node)
  unsigned short lcb[DT_UPPER + 3];  // log: same
  u8 lmode[DT_UPPER + 3];            // 0 internal, 1 uniform (value = payload), 2 equal (value = payload + snapshot cell)
  u32 wtot[DT_WARPS];
  u32 run;
  InstDir dir_s, dir_l;              // directory entries of the structures being expanded
  __align__(16) u8 stage_s[DT_STAGE_S + 32];
  __align__(16) u8 stage_l[DT_STAGE_L + 32];
};
typedef TileSmemT<i64> TileSmem;

DCDF_DEVINL u32 lvl_off(int k) { return (0x55555555u >> (32 - 2 * k)) & (k ? 0xffffffffu : 0u); }  // (4^k - 1) / 3

// Copy `size` bytes starting at global address `src` into `stage` keeping the address modulo 16, and return
// the pointer that corresponds to `src`.  Falls back to `src` itself when the structure does not fit.
DCDF_DEVINL const u8* stage_bytes(const u8* src, u32 size, u8* stage, u32 cap, u32 n_threads = DT_THREADS) {
  const u32 mis = (u32)((uintptr_t)src & 15u);
  if (size + mis + 4u > cap) return src;
  const uint4* g = reinterpret_cast<const uint4*>(src - mis);
  uint4* d = reinterpret_cast<uint4*>(stage);
  const u32 n16 = (size + mis + 4u + 15u) / 16u;  // +4: be32_at reads one aligned word past the last byte
  for (u32 i = threadIdx.x; i < n16; i += n_threads) d[i] = g[i];
  return stage + mis;
}

// Block-wide: turn the "internal" markers of level k (k <= 5, at most 1024 positions) into child BFS bases;
// returns the number of internal nodes.  Each warp owns a contiguous run of positions (ballots stay in
// registers), one barrier turns the per-warp counts into offsets.
template <typename V>
DCDF_DEVINL u32 assign_child_bases(unsigned short* cb, int k, u32 p_next, TileSmemT<V>& S) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 n = 1u << (2 * k), off = lvl_off(k);
  const u32 seg = n > 32u * DT_WARPS ? n / DT_WARPS : 32u;  // 32 or 128
  u32 bal[4], cnt = 0;
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const u32 p = warp * seg + s * 32u + lane;
    const bool f = s * 32u < seg && p < n && cb[off + p] != 0xffffu;
    bal[s] = __ballot_sync(0xffffffffu, f);
    cnt += __popc(bal[s]);
  }
  if (lane == 0) S.wtot[warp] = cnt;
  __syncthreads();
  u32 before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < DT_WARPS; w++) {
    const u32 x = S.wtot[w];
    if (w < warp) before += x;
    total += x;
  }
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const u32 p = warp * seg + s * 32u + lane;
    if ((bal[s] >> lane) & 1u) cb[off + p] = (unsigned short)(p_next + 4u * (before + __popc(bal[s] & lanemask_lt())));
    before += __popc(bal[s]);
  }
  __syncthreads();
  return total;
}

// Level L (the cells) is stored row-major with pitch 2^L so that the output loop reads consecutive addresses;
// the levels above stay in Morton order.
DCDF_DEVINL u32 cell_index(u32 p1, int L, bool last) {
  return last ? (morton_row(p1) << L) + morton_col(p1) : p1;
}

// Expand a Snapshot into S.sval / S.scb (L = tree levels, 1..6).
template <typename V>
DCDF_DEVINL void expand_snapshot(const ChunkView& cv, const InstDir& s, int L, TileSmemT<V>& S) {
  const int tid = threadIdx.x;
  const BitsFast nm = bits_fast(cv.chunk, s.nm_len, s.nm_base);
  const DacFast mx = dac_fast(cv.chunk, &s.max);
  if (tid == 0) {
    S.sval[0] = mx.getv<V>(0);
    S.scb[0] = nm.get(0) ? 1 : 0xffff;
  }
  __syncthreads();
  u32 p_next = 1;
  for (int k = 0; k < L; k++) {
    const u32 internal = assign_child_bases(S.scb, k, p_next, S);
    const u32 n1 = 1u << (2 * (k + 1)), o0 = lvl_off(k), o1 = lvl_off(k + 1);
    const bool has_bits = k + 1 < L;
    for (u32 p1 = tid; p1 < n1; p1 += DT_THREADS) {
      const u32 p = p1 >> 2;
      const u32 cb = S.scb[o0 + p];
      V v = S.sval[o0 + p];
      bool in = false;
      if (cb != 0xffffu) {
        const u32 idx = cb + (p1 & 3u);
        v -= mx.getv<V>(idx);
        in = has_bits && idx < s.nm_len && nm.get(idx);
      }
      S.sval[o1 + cell_index(p1, L, !has_bits)] = v;
      if (has_bits) S.scb[o1 + p1] = in ? 1 : 0xffff;
    }
    p_next += 4u * internal;
    __syncthreads();
  }
}

// Expand a Log against the snapshot pyramid already in S.sval down to level L - 1 (the 2x2 quads); the cells are
// produced by the caller's fused output pass.
template <typename V>
DCDF_DEVINL void expand_log(const ChunkView& cv, const ChunkView& cv_snap, const InstDir& l, const InstDir& s, int L, TileSmemT<V>& S) {
  const int tid = threadIdx.x;
  const BitsFast nm = bits_fast(cv.chunk, l.nm_len, l.nm_base);
  const BitMapRef nm_rank{cv.chunk, l.nm_len, l.nm_base}, eq{cv.chunk, l.eq_len, l.eq_base};
  const DacFast mx = dac_fast(cv.chunk, &l.max);
  if (tid == 0) {
    const V d0 = mx.getv<V>(0);
    const bool single_t = !nm.get(0);
    if (!single_t) {
      S.lmode[0] = 0; S.lpay[0] = d0; S.lcb[0] = 1;
    } else {
      // log.rs:180-186: a single-node log is uniform unless its equal bit says "snapshot + constant"
      BitMapRef nm_s{cv_snap.chunk, s.nm_len, s.nm_base};
      const bool snap_single = !nm_s.get(0);
      const bool uniform = snap_single || !eq.get(0);
      S.lmode[0] = uniform ? 1 : 2;
      S.lpay[0] = uniform ? d0 + S.sval[0] : d0;
      S.lcb[0] = 0xffff;
    }
  }
  __syncthreads();
  u32 p_next = 1;
  for (int k = 0; k < L; k++) {
    const u32 internal = assign_child_bases(S.lcb, k, p_next, S);
    if (k == L - 1) break;  // child bases of the quads are all the fused cell pass needs
    const u32 n1 = 1u << (2 * (k + 1)), o0 = lvl_off(k), o1 = lvl_off(k + 1);
    const bool has_bits = true;
    for (u32 p1 = tid; p1 < n1; p1 += DT_THREADS) {
      const u32 p = p1 >> 2;
      u8 mode = S.lmode[o0 + p];
      V pay = S.lpay[o0 + p];
      bool in = false;
      if (mode == 0) {
        const u32 idx = S.lcb[o0 + p] + (p1 & 3u);
        const V d = mx.getv<V>(idx);  // max_t is replaced, not accumulated (log.rs:233)
        in = has_bits && idx < l.nm_len && nm.get(idx);
        if (in) {
          mode = 0; pay = d;
        } else if (!has_bits) {
          mode = 2; pay = d;  // bottom level: value = max_t + snapshot cell
        } else {
          const bool e = eq.get(idx - nm_rank.rank(idx));  // rank0(idx + 1) - 1 (log.rs:265)
          mode = e ? 2 : 1;
          pay = e ? d : d + S.sval[o1 + p1];  // uniform: max_t + max_s of this node (log.rs:266-268)
        }
      }
      S.lmode[o1 + p1] = mode;
      S.lpay[o1 + p1] = pay;
      S.lcb[o1 + p1] = in ? 1 : 0xffff;
    }
    p_next += 4u * internal;
    __syncthreads();
  }
}

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

template <typename V>
__global__ void __launch_bounds__(DT_THREADS) k_window_tiles(const TileWindowParams P) {
  extern __shared__ __align__(16) unsigned char dt_smem_raw[];
  TileSmemT<V>& S = *reinterpret_cast<TileSmemT<V>*>(dt_smem_raw);
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
    const i64 tile_org = (chunk_top - c.top) * W_cols + (chunk_left - c.left);  // element offset of tile cell (0, 0)
    auto out_index = [&](i64 t, int r, int col) {
      return obase + (u64)((t - c.start) * W_rows * W_cols + tile_org) + (u64)((i64)r * W_cols + col);
    };
    CellOut co;
    co.init(Q, P.out, P.raw, m.bits);
    if (!stored) {
      // Elided: one value per instant from the max table, parent's fractional bits (superchunk.rs:426-433)
      const SlotDesc sdsc = Q.slot_desc[sm.slot_base + slot];
      for (i64 t = t_lo; t < t_hi; t++) {
        const i64 v = Q.tbl_max[sdsc.tbl0 + (u64)(t - sm.t0) * sdsc.stride];
        for (int i = tid; i < wr * wc; i += DT_THREADS) emit(Q, P.out, out_index(t, top + i / wc, left + i % wc), v, sdsc.bits, P.raw);
      }
      continue;
    }
    const u8* chunk = Q.blob + m.blob_off;
    const InstDir* dir = Q.dir + m.dir_base;
    const int L = 31 - __clz(m.sidelen);
    const u32 oL = lvl_off(L);
    u32 cur_snap = 0xffffffffu;
    ChunkView cv_s{chunk, dir, m.sidelen};
    const u8* cv_l_chunk = chunk;
    for (i64 t = t_lo; t < t_hi; t++) {
      const u32 ti = (u32)(t - sm.t0);
      const u32 snap = dir[ti].snap;
      __syncthreads();
      if (snap != cur_snap) {
        if (tid == 0) S.dir_s = dir[snap];
        const u32 off = dir[snap].off;
        cv_s.chunk = stage_bytes(chunk + off, dir[snap].size, S.stage_s, DT_STAGE_S + 32) - off;
        __syncthreads();
        expand_snapshot(cv_s, S.dir_s, L, S);
        cur_snap = snap;
      }
      const bool is_log = snap != ti;
      if (is_log) {
        if (tid == 0) S.dir_l = dir[ti];
        const u32 off = dir[ti].off;
        cv_l_chunk = stage_bytes(chunk + off, dir[ti].size, S.stage_l, DT_STAGE_L + 32) - off;
        ChunkView cv_l{cv_l_chunk, dir, m.sidelen};
        __syncthreads();
        expand_log(cv_l, cv_s, S.dir_l, S.dir_s, L, S);
      }
      if (!is_log) {
        // rows of the window inside this tile; consecutive threads write consecutive columns
        for (int r = top + (tid / 64); r < bottom; r += DT_THREADS / 64) {
          const int col = left + (tid & 63);
          if (col < right) co.put(out_index(t, r, col), S.sval[oL + ((u32)r << L) + (u32)col]);
        }
      } else {
        // fused cell pass: one thread per 2x2 quad (row-major over quads so that a warp writes whole row segments);
        // value = payload (uniform), payload + snapshot cell (equal), or the cell's own log entry + snapshot cell
        const DacFast mxl = dac_fast(cv_l_chunk, &S.dir_l.max);
        const int half = 1 << (L - 1);
        const u32 oQ = lvl_off(L - 1);
        for (int qi = tid; qi < half * half; qi += DT_THREADS) {
          const int qr = qi >> (L - 1), qc = qi & (half - 1);
          const int r0q = 2 * qr, c0q = 2 * qc;
          if (r0q + 1 < top || r0q >= bottom || c0q + 1 < left || c0q >= right) continue;
          const u32 q = L > 1 ? morton_encode((u32)qr, (u32)qc) : 0u;
          const u8 mode = S.lmode[oQ + q];
          const V pay = S.lpay[oQ + q];
          const u32 cb = S.lcb[oQ + q];
          const u64 i00 = out_index(t, r0q, c0q);  // the quad's cells are i00, +1, +W_cols, +W_cols+1
          const bool inside = r0q >= top && r0q + 1 < bottom && c0q >= left && c0q + 1 < right;
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const int r = r0q + (c >> 1), col = c0q + (c & 1);
            if (!inside && (r < top || r >= bottom || col < left || col >= right)) continue;
            const V sc = S.sval[oL + ((u32)r << L) + (u32)col];
            V v;
            if (mode == 1) v = pay;
            else if (mode == 2) v = pay + sc;
            else v = mxl.getv<V>(cb + (u32)c) + sc;
            co.put(i00 + (u64)((c >> 1) ? W_cols : 0) + (u64)(c & 1), v);
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace dcdf
