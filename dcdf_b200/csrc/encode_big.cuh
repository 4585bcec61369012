// encode_big.cuh -- Chunk::build (chunk.rs:42-96) for rasters whose padded side exceeds 64 (up to 4096).
//
// The whole padded square of one instant is kept as a dense min/max pyramid in global memory, level-major
// (level l at offset (4^l - 1) / 3) and Morton-ordered inside a level (row bit above column bit), which is
// exactly the order of the reference's breadth-first traversal restricted to existing nodes
// (snapshot.rs:126-145, log.rs:131-154).  Every step is a flat data-parallel pass:
//   kb_leaves / kb_reduce   bottom-up pyramid of t and (t - s)         snapshot.rs:439-500, log.rs:725-817
//   kb_classify             node flags, existence (all ancestors internal), entry byte-length histograms
//   device-wide scan        BFS positions = exclusive scan of the existence flags in level-major order
//   kb_scatter              max / min codes, nodemap / equal flags at their BFS positions
//   kb_bitmap_*, kb_dac_*   BitMap words + rank directory, DAC levels with scan-computed offsets
// The host reads back a few counters per instant to take the Snapshot-vs-Log decision (chunk.rs:62).
#pragma once
#include "common.cuh"

namespace dcdf {

constexpr i64 BIG_NONE_MAX = INT64_MIN, BIG_NONE_MIN = INT64_MAX;
enum : u8 { BF_INT_S = 1, BF_INT_L = 2, BF_UNIFORM = 4, BF_EQ = 8, BF_ALIVE_S = 16, BF_ALIVE_L = 32 };

__host__ __device__ inline u64 big_lvl_off(int l) { return ((1ull << (2 * l)) - 1ull) / 3ull; }

struct BigPyr {
  i64* tmax;  // current instant
  i64* tmin;
  i64* smax;  // reference snapshot instant (same layout)
  i64* smin;
  i64* diff;  // log: t - s of the first cell below the node (log.rs:751,781)
  u8* fl;     // BF_* flags
  int L;      // tree levels: side = 2^L
  u64 N;      // nodes = (4^(L+1) - 1) / 3
};

template <typename InT>
__global__ void kb_leaves(const InT* data, i64 sr, i64 sc, int rows, int cols, int bits, int round, BigPyr P, u32* err) {
  const u64 n = 1ull << (2 * P.L), off = big_lvl_off(P.L);
  u32 e = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u32 r = morton_row((u32)i), c = morton_col((u32)i);
    i64 mx = BIG_NONE_MAX, mn = BIG_NONE_MIN, d = 0;
    if ((int)r < rows && (int)c < cols) {
      const i64 v = Conv<InT>::get(data[(i64)r * sr + (i64)c * sc], bits, round != 0, e);
      mx = mn = v;
      const i64 s = P.smax[off + i];
      d = v - (s == BIG_NONE_MAX ? 0 : s);
    }
    P.tmax[off + i] = mx; P.tmin[off + i] = mn; P.diff[off + i] = d;
    P.fl[off + i] = BF_EQ | BF_UNIFORM;  // a leaf is "equal" (log.rs:757) and never internal
  }
  if (e) atomicOr(err, e);
}

// level l from level l + 1
__global__ void kb_reduce(BigPyr P, int l) {
  const u64 n = 1ull << (2 * l), o = big_lvl_off(l), oc = big_lvl_off(l + 1);
  for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) {
    i64 mx = BIG_NONE_MAX, mn = BIG_NONE_MIN;
    const i64 d0 = P.diff[oc + 4 * p];
    bool eq = true;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const u64 ci = oc + 4 * p + c;
      const i64 a = P.tmax[ci], b = P.tmin[ci];
      mx = a > mx ? a : mx;
      mn = b < mn ? b : mn;
      eq = eq && (P.fl[ci] & BF_EQ) && P.diff[ci] == d0;
    }
    P.tmax[o + p] = mx; P.tmin[o + p] = mn; P.diff[o + p] = d0;
    const bool uniform = mx == BIG_NONE_MAX || mx == mn;  // Option equality: None == None (log.rs:137)
    u8 f = (eq ? BF_EQ : 0) | (uniform ? BF_UNIFORM : 0);
    if (!uniform) f |= BF_INT_S;                 // snapshot.rs:133
    if (!uniform && !eq) f |= BF_INT_L;          // log.rs:137-146
    P.fl[o + p] = f;
  }
}

DCDF_DEVINL i64 big_or0(i64 v) { return (v == BIG_NONE_MAX || v == BIG_NONE_MIN) ? 0 : v; }

// Counters: [cand][arr][j] (len > j) then structure counts.
struct BigCounts {
  unsigned long long h[2][2][8];
  unsigned long long n_alive[2], n_upper[2], n_int[2];
};

DCDF_DEVINL int big_level_of(u64 node, int L) {
  // level l holds nodes [off(l), off(l+1)); off(l) = (4^l - 1)/3  ->  3*node + 1 in [4^l, 4^(l+1))
  return (63 - __clzll((long long)(3ull * node + 1ull))) >> 1;
}

// Entry values of a node for one candidate.  Returns false if the node has no min entry.
DCDF_DEVINL void big_entries(const BigPyr& P, int cand, int l, u64 p, i64& emax, i64& emin) {
  const u64 o = big_lvl_off(l) + p;
  const i64 tm = big_or0(P.tmax[o]);
  if (cand == 0) {
    if (l == 0) { emax = tm; emin = P.tmin[o]; }
    else {
      const u64 po = big_lvl_off(l - 1) + (p >> 2);
      emax = P.tmax[po] - tm;           // snapshot.rs:139
      emin = P.tmin[o] - P.tmin[po];    // snapshot.rs:140
    }
  } else {
    emax = tm - big_or0(P.smax[o]);     // log.rs:133
    emin = P.tmin[o] - P.smin[o];       // log.rs:149
  }
}

// One thread per node: existence for both candidates, histograms (only when `count`).
__global__ void kb_classify(BigPyr P, int first, BigCounts* counts) {
  __shared__ unsigned long long sh[2][2][8];
  __shared__ unsigned long long ss[6];
  const int tid = threadIdx.x;
  if (tid < 32) (&sh[0][0][0])[tid] = 0;
  if (tid < 6) ss[tid] = 0;
  __syncthreads();
  for (u64 node = (u64)blockIdx.x * blockDim.x + tid; node < P.N; node += (u64)gridDim.x * blockDim.x) {
    const int l = big_level_of(node, P.L);
    const u64 p = node - big_lvl_off(l);
    bool aS = true, aL = !first;
    u64 a = p;
    for (int k = l - 1; k >= 0 && (aS || aL); k--) {
      a >>= 2;
      const u8 f = P.fl[big_lvl_off(k) + a];
      aS = aS && (f & BF_INT_S);
      aL = aL && (f & BF_INT_L);
    }
    u8 f = P.fl[node] & ~(BF_ALIVE_S | BF_ALIVE_L);
    if (aS) f |= BF_ALIVE_S;
    if (aL) f |= BF_ALIVE_L;
    P.fl[node] = f;
    for (int cand = 0; cand < 2; cand++) {
      if (!(cand ? aL : aS)) continue;
      i64 emax, emin;
      big_entries(P, cand, l, p, emax, emin);
      const int len = dac_len(zigzag64(emax));
      for (int j = 0; j < len; j++) atomicAdd(&sh[cand][0][j], 1ull);
      atomicAdd(&ss[cand], 1ull);
      if (l < P.L) atomicAdd(&ss[2 + cand], 1ull);
      if (l < P.L && (f & (cand ? BF_INT_L : BF_INT_S))) {
        const int len2 = dac_len(zigzag64(emin));
        for (int j = 0; j < len2; j++) atomicAdd(&sh[cand][1][j], 1ull);
        atomicAdd(&ss[4 + cand], 1ull);
      }
    }
  }
  __syncthreads();
  if (tid < 32 && (&sh[0][0][0])[tid]) atomicAdd(&counts->h[0][0][0] + tid, (&sh[0][0][0])[tid]);
  if (tid < 2) {
    if (ss[tid]) atomicAdd(&counts->n_alive[tid], ss[tid]);
    if (ss[2 + tid]) atomicAdd(&counts->n_upper[tid], ss[2 + tid]);
    if (ss[4 + tid]) atomicAdd(&counts->n_int[tid], ss[4 + tid]);
  }
}

// scan input: lo 32 bits = node exists in the winner, hi 32 bits = exists and internal
__global__ void kb_scan_input(BigPyr P, int cand, u64* out) {
  const u8 alive = cand ? BF_ALIVE_L : BF_ALIVE_S, in = cand ? BF_INT_L : BF_INT_S;
  for (u64 node = (u64)blockIdx.x * blockDim.x + threadIdx.x; node < P.N; node += (u64)gridDim.x * blockDim.x) {
    const u8 f = P.fl[node];
    const u64 a = (f & alive) ? 1ull : 0ull;
    out[node] = a | ((a && (f & in)) ? (1ull << 32) : 0ull);
  }
}

// ---- generic device-wide exclusive scan of u64 (packed counters add independently while sums < 2^32)
constexpr int SCAN_ITEMS = 4096;  // per CTA (1024 threads x 4)
__global__ void __launch_bounds__(1024) kb_scan_blocks(const u64* in, u64 n, u64* out, u64* block_sums) {
  __shared__ u64 wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u64 base = (u64)blockIdx.x * SCAN_ITEMS + (u64)tid * 4;
  u64 v[4], t = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { v[k] = base + k < n ? in[base + k] : 0ull; t += v[k]; }
  u64 x = t;
  for (int o = 1; o < 32; o <<= 1) {
    const u64 y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    const u64 w = wsum[lane];
    u64 xs = w;
    for (int o = 1; o < 32; o <<= 1) {
      const u64 y = __shfl_up_sync(0xffffffffu, xs, o);
      if (lane >= o) xs += y;
    }
    wsum[lane] = xs - w;
    if (lane == 31) block_sums[blockIdx.x] = xs;
  }
  __syncthreads();
  u64 run = wsum[warp] + x - t;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
}
// single-CTA exclusive scan of the block sums (in place); sums[n] is not written
__global__ void __launch_bounds__(1024) kb_scan_top(u64* sums, u64 n) {
  __shared__ u64 wsum[32];
  __shared__ u64 carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (u64 base = 0; base < n; base += 1024u) {
    const u64 i = base + tid;
    const u64 v = i < n ? sums[i] : 0ull;
    u64 x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const u64 y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      const u64 w = wsum[lane];
      u64 xs = w;
      for (int o = 1; o < 32; o <<= 1) {
        const u64 y = __shfl_up_sync(0xffffffffu, xs, o);
        if (lane >= o) xs += y;
      }
      wsum[lane] = xs - w;
    }
    __syncthreads();
    const u64 carry = carry_s;
    if (i < n) sums[i] = carry + wsum[warp] + x - v;
    __syncthreads();
    if (tid == 1023) carry_s = carry + wsum[warp] + x;
    __syncthreads();
  }
}
__global__ void kb_scan_add(u64* out, u64 n, const u64* block_offsets) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += block_offsets[i / SCAN_ITEMS];
}

// Place codes and flags at their BFS positions.
__global__ void kb_scatter(BigPyr P, int cand, const u64* scan, u64* A, u64* B, u8* nmf, u8* eqf) {
  const u8 alive = cand ? BF_ALIVE_L : BF_ALIVE_S, in = cand ? BF_INT_L : BF_INT_S;
  for (u64 node = (u64)blockIdx.x * blockDim.x + threadIdx.x; node < P.N; node += (u64)gridDim.x * blockDim.x) {
    const u8 f = P.fl[node];
    if (!(f & alive)) continue;
    const int l = big_level_of(node, P.L);
    const u64 p = node - big_lvl_off(l);
    const u64 pos = scan[node] & 0xffffffffull, pint = scan[node] >> 32;
    i64 emax, emin;
    big_entries(P, cand, l, p, emax, emin);
    A[pos] = zigzag64(emax);
    if (l < P.L) {
      const bool internal = f & in;
      nmf[pos] = internal ? 1 : 0;
      if (internal) B[pint] = zigzag64(emin);
      else if (cand) eqf[pos - pint] = (!(f & BF_UNIFORM) && (f & BF_EQ)) ? 1 : 0;  // log.rs:137-146
    }
  }
}

// ---- BitMap from flag bytes: words (big-endian) + per-word popcounts for the rank directory
__global__ void kb_bitmap_words(const u8* flags, u64 length, u8* out_words, u64* popc) {
  const u64 gw = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per word
  const int lane = threadIdx.x & 31;
  const u64 words = (length + 31) / 32;
  if (gw >= words) return;
  const u64 i = gw * 32 + lane;
  const u32 bal = __ballot_sync(0xffffffffu, i < length && flags[i] != 0);
  if (lane == 0) {
    store_be32(out_words + 4 * gw, __brev(bal));
    popc[gw] = __popc(bal);
  }
}
__global__ void kb_bitmap_index(const u64* popc_scan, const u64* popc, u64 length, u8* out) {
  const u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 blocks = length / 128;
  if (b == 0) {
    store_be32(out, (u32)length);
    store_be32(out + 4, 4u);
  }
  if (b < blocks) store_be32(out + 8 + 4 * b, (u32)(popc_scan[4 * b + 3] + popc[4 * b + 3]));  // ones in [0, 128(b+1))
}

// ---- one DAC level: flags of survivors, then bytes / words / directory / compaction
__global__ void kb_dac_more(const u64* x, u64 n, u64* more) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) more[i] = (x[i] >> 8) ? 1ull : 0ull;
}
__global__ void kb_dac_level(const u64* x, u64 n, const u64* more_scan, u8* out, u64* next) {
  // out -> BitMap header of this level; layout: [len][k][index][words][bytes]
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const u64 blocks = n / 128, words = (n + 31) / 32;
  u8* p_index = out + 8;
  u8* p_words = p_index + 4 * blocks;
  u8* p_bytes = p_words + 4 * words;
  if (i == 0) {
    store_be32(out, (u32)n);
    store_be32(out + 4, 4u);
  }
  const u64 v = i < n ? x[i] : 0ull;
  const bool more = (v >> 8) != 0;
  const u32 bal = __ballot_sync(0xffffffffu, more);
  if (i < n) {
    p_bytes[i] = (u8)(v & 0xff);
    if (more) next[more_scan[i]] = v >> 8;
    if (((i + 1) & 127ull) == 0) store_be32(p_index + 4 * ((i + 1) / 128 - 1), (u32)(more_scan[i] + (more ? 1 : 0)));
  }
  if (lane == 0 && (i - lane) < n) store_be32(p_words + 4 * (i >> 5), __brev(bal));
}

__global__ void kb_header(u8* out, int rows, int cols, int side) {
  out[0] = 2;
  store_be32(out + 1, (u32)rows);
  store_be32(out + 5, (u32)cols);
  store_be32(out + 9, (u32)side);
}
__global__ void kb_byte(u8* out, u8 v) { out[0] = v; }

// copy the t pyramid into the s pyramid (new block: this instant becomes the reference snapshot)
__global__ void kb_copy_ts(BigPyr P) {
  for (u64 node = (u64)blockIdx.x * blockDim.x + threadIdx.x; node < P.N; node += (u64)gridDim.x * blockDim.x) {
    P.smax[node] = P.tmax[node];
    P.smin[node] = P.tmin[node];
  }
}

}  // namespace dcdf
