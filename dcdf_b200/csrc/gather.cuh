// gather.cuh -- layout of the variable-length outputs.
//   * k_table_dac     : Dac::from(max) / Dac::from(min) of a superchunk's per-(instant, subchunk) tables
//                       (superchunk.rs:190-198, 248-249; dac.rs:96-132) into arena pieces
//   * k_scan_units    : device-wide exclusive scan of Chunk sizes -> final byte offsets
//   * k_gather_chunks : Chunk::write_to / Block::write_to framing (chunk.rs:235-243, block.rs:88-95)
//                       around the Snapshot / Log pieces the encoder left in the arena
#pragma once
#include "common.cuh"
#include "encode_tile.cuh"

namespace dcdf {

// ------------------------------------------------------------------ superchunk min/max DACs
struct TableDacParams {
  const i64* tbl_min;
  const i64* tbl_max;
  const u64* table_base;  // [n_tables]
  const u64* table_len;   // [n_tables] instants * n_children
  const int* alive;       // [n_tables] stride 2 ints (NodeState): tables of nodes that are not built are skipped
  u64* scratch;           // zigzag codes + compaction partners, same indexing as the tables; [4][total]
  u64 total;              // sum of table_len
  Piece* pieces;          // [n_tables][2]  (0 = max, 1 = min)
  u8* arena;
  u64 arena_cap;
  unsigned long long* arena_head;
  u32* err;
};

// grid = (n_slices, 2); block = ENC_THREADS
__global__ void __launch_bounds__(ENC_THREADS) k_table_dac(const TableDacParams P) {
  const u32 s = blockIdx.x, which = blockIdx.y;  // which: 0 = max, 1 = min
  const int tid = threadIdx.x, lane = tid & 31;
  if (P.alive && !P.alive[2 * s]) {
    if (tid == 0) { Piece pc; pc.off = 0; pc.size = 0; pc.kind = 2u + which; P.pieces[2 * s + which] = pc; }
    return;
  }
  const i64* src = (which == 0 ? P.tbl_max : P.tbl_min) + P.table_base[s];
  u64* a = P.scratch + (u64)which * P.total + P.table_base[s];
  u64* a2 = P.scratch + (u64)(2 + which) * P.total + P.table_base[s];  // compaction partner
  const u32 n = (u32)P.table_len[s];
  __shared__ u32 cnt[8];
  __shared__ u32 scan_a[ENC_WARPS];
  __shared__ u64 s_off;
  if (tid < 8) cnt[tid] = 0;
  __syncthreads();
  // pass 1: zigzag + byte-length histogram
  u32 h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (u32 i = tid; i < n; i += ENC_THREADS) {
    const u64 zz = zigzag64(src[i]);
    a[i] = zz;
    const int len = dac_len(zz);
#pragma unroll
    for (int j = 0; j < 8; j++) h[j] += (len > j) ? 1u : 0u;
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    u32 r = __reduce_add_sync(0xffffffffu, h[j]);
    if (lane == 0 && r) atomicAdd(&cnt[j], r);
  }
  __syncthreads();
  u32 c[8];
#pragma unroll
  for (int j = 0; j < 8; j++) c[j] = cnt[j];
  const u32 size = dac_size_from_counts(c, 8);
  if (tid == 0) {
    const u64 need = ((u64)size + 15ull) & ~15ull;
    const u64 off = atomicAdd(P.arena_head, (unsigned long long)need);
    s_off = off;
    Piece pc;
    pc.off = off; pc.size = size; pc.kind = 2u + which;
    P.pieces[2 * s + which] = pc;
  }
  __syncthreads();
  const u64 off = s_off;
  if (off + (((u64)size + 15ull) & ~15ull) > P.arena_cap) {
    if (tid == 0) atomicOr(P.err, (u32)EF_ARENA_FULL);
    return;
  }
  __shared__ u32 c_s[8];
  if (tid < 8) c_s[tid] = c[tid];
  __syncthreads();
  const u32 wrote = block_dac_emit<u64, 8>(a, a2, c_s, P.arena + off, scan_a);
  if (tid == 0 && wrote != size) atomicOr(P.err, (u32)EF_BAD_FORMAT);
}

// ------------------------------------------------------------------ exclusive scan of chunk sizes
// Single CTA (1024 threads); n up to a few hundred thousand units.  out[n] = total.
__global__ void __launch_bounds__(1024) k_scan_units(const UnitResult* results, const u8* stored, u32 n, u64* out) {
  __shared__ u64 wsum[32];
  __shared__ u64 carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (u32 base = 0; base < n; base += 1024u) {
    const u32 i = base + tid;
    u64 v = (i < n && stored[i]) ? results[i].bytes : 0ull;
    u64 x = v;
    for (int o = 1; o < 32; o <<= 1) {
      u64 y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      u64 w = wsum[lane];
      u64 xs = w;
      for (int o = 1; o < 32; o <<= 1) {
        u64 y = __shfl_up_sync(0xffffffffu, xs, o);
        if (lane >= o) xs += y;
      }
      wsum[lane] = xs - w;  // exclusive
    }
    __syncthreads();
    const u64 carry = carry_s;
    if (i < n) out[i] = carry + wsum[warp] + x - v;
    __syncthreads();
    if (tid == 1023) carry_s = carry + wsum[warp] + x;
    __syncthreads();
  }
  if (tid == 0) out[n] = carry_s;
}

// ------------------------------------------------------------------ final Chunk bytes
struct GatherParams {
  const EncUnit* units;
  const UnitResult* results;
  const u8* stored;
  const Piece* pieces;
  const u64* chunk_off;  // from k_scan_units
  const u8* arena;
  u8* out;               // final blob
  u64 out_cap;
  int encoding;
  u32 n_units;
  u32* err;
};

constexpr int GATHER_MAX_T = 1024;

// Warp-cooperative copy of n bytes from a 4-byte aligned source to an arbitrarily aligned destination:
// destination-aligned 32-bit stores, each built from two aligned source words with a funnel shift.
DCDF_DEVINL void warp_copy_bytes(u8* dst, const u8* src, u32 n, int lane) {
  const u32 head = min(n, (u32)((4u - ((uintptr_t)dst & 3u)) & 3u));
  if ((u32)lane < head) dst[lane] = src[lane];
  const u32 body = (n - head) >> 2;
  u32* d4 = reinterpret_cast<u32*>(dst + head);
  const u8* s0 = src + head;
  const u32 sh = (u32)((uintptr_t)s0 & 3u) * 8u;
  const u32* s4 = reinterpret_cast<const u32*>((uintptr_t)s0 & ~(uintptr_t)3);
  u32 w = lane;
  for (; w + 96 < body; w += 128) {  // four independent word pairs in flight per lane
    u32 lo[4], hi[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { lo[j] = s4[w + 32 * j]; hi[j] = sh ? s4[w + 32 * j + 1] : 0u; }
#pragma unroll
    for (int j = 0; j < 4; j++) d4[w + 32 * j] = sh ? __funnelshift_r(lo[j], hi[j], sh) : lo[j];
  }
  for (; w < body; w += 32) {
    const u32 lo = s4[w];
    d4[w] = sh ? __funnelshift_r(lo, s4[w + 1], sh) : lo;
  }
  const u32 done = head + 4u * body;
  if (done + (u32)lane < n) dst[done + lane] = src[done + lane];
}

__global__ void __launch_bounds__(256) k_gather_chunks(const GatherParams P) {
  const u32 u = blockIdx.x;
  if (u >= P.n_units || !P.stored[u]) return;
  const EncUnit unit = P.units[u];
  const UnitResult res = P.results[u];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u64 base = P.chunk_off[u];
  if (base + res.bytes > P.out_cap) {
    if (tid == 0) atomicOr(P.err, (u32)EF_OUT_CAP);
    return;
  }
  u8* out = P.out + base;
  __shared__ u32 dst_off[GATHER_MAX_T];
  __shared__ Piece pcs[GATHER_MAX_T];  // piece table of the current batch, fetched by all threads at once
  __shared__ u32 carry_off, last_snap_pos, block_count;
  const Piece* pieces = P.pieces + unit.piece_base;
  if (tid == 0) {
    out[0] = (u8)P.encoding;                 // chunk.rs:236
    out[1] = (u8)unit.bits;                  // chunk.rs:237
    store_be32(out + 2, res.snapshots);      // chunk.rs:238 (number of blocks)
    carry_off = 6;
    last_snap_pos = 0;
    block_count = 0;
  }
  __syncthreads();
  for (int t0 = 0; t0 < unit.instants; t0 += GATHER_MAX_T) {
    const int nt = min(GATHER_MAX_T, unit.instants - t0);
    for (int i = tid; i < nt; i += 256) pcs[i] = pieces[t0 + i];
    __syncthreads();
    if (tid == 0) {
      u32 off = carry_off, lsp = last_snap_pos, bc = block_count;
      for (int i = 0; i < nt; i++) {
        const Piece pc = pcs[i];
        if (pc.kind == 1u) {
          if (lsp) out[lsp] = (u8)bc;        // close the previous block: n_instants (block.rs:89)
          lsp = off;
          bc = 0;
          off += 1;
        }
        bc++;
        dst_off[i] = off;
        off += pc.size;
      }
      carry_off = off; last_snap_pos = lsp; block_count = bc;
    }
    __syncthreads();
    for (int i = warp; i < nt; i += 8) {
      const Piece pc = pcs[i];
      const u8* src = P.arena + pc.off;
      u8* dst = out + dst_off[i];
      warp_copy_bytes(dst, src, pc.size, lane);
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (last_snap_pos) out[last_snap_pos] = (u8)block_count;
    if (carry_off != (u32)res.bytes) atomicOr(P.err, (u32)EF_BAD_FORMAT);
  }
}

// Copy the table DAC pieces into a dense blob: per slice [max DAC][min DAC]; offsets precomputed on host.
struct GatherDacParams {
  const Piece* pieces;   // [n_slices][2]
  const u64* dst_off;    // [n_slices][2]
  const u8* arena;
  u8* out;
};
__global__ void __launch_bounds__(256) k_gather_dacs(const GatherDacParams P) {
  const u32 id = blockIdx.x;
  const Piece pc = P.pieces[id];
  const u8* src = P.arena + pc.off;
  u8* dst = P.out + P.dst_off[id];
  for (u32 b = threadIdx.x; b < pc.size; b += blockDim.x) dst[b] = src[b];
}

}  // namespace dcdf
