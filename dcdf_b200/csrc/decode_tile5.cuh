// decode_tile5.cuh -- k_window_tiles4's per-thread walk behind a TMA / mbarrier pipeline.
//
// k_window_tiles4 needs one block barrier per instant only to publish the bytes its threads fetched with cp.async.
// Here one elected thread issues two bulk copies per structure (cp.async.bulk: its directory entry and its bytes,
// 16-byte aligned) into a ring of three slots; a slot's `full` mbarrier counts the bytes in, its `empty` mbarrier
// counts the eight warps out.  Warps never wait for each other, only for data (and the producer for the slowest warp,
// one ring turn behind), so a warp whose 32 blocks are cheap at one instant runs ahead instead of idling at a barrier.
// Nothing a warp reads is written by another warp: cells, quads and levels >= 2 of the snapshot pyramid are private
// to the thread's path, levels 0 and 1 are kept per warp (decode_tile4.cuh: sup_at), the rank table is per warp.
#pragma once
#include "decode_tile4.cuh"

namespace dcdf {

constexpr int W5_SLOTS = 3;

template <typename V>
struct Tile5Smem {
  static constexpr int BUF = sizeof(V) == 4 ? 8704 : 16 * 1024;
  static constexpr bool PRIVATE_TOP = true;
  struct Slot {
    __align__(16) InstDir dir;
    __align__(16) u8 stage[BUF + 32];
  };
  __align__(16) V cells[4096];
  __align__(16) V sup[W3_UPPER];
  RankTab tab[DT_WARPS];
  V suptop[DT_WARPS][8];
  u32 single[DT_WARPS];
  __align__(8) unsigned long long full[W5_SLOTS], empty[W5_SLOTS];
  Slot slot[W5_SLOTS];
};

DCDF_DEVINL u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
DCDF_DEVINL void mbar_init(unsigned long long* b, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
DCDF_DEVINL void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
DCDF_DEVINL void mbar_arrive_expect_tx(unsigned long long* b, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
DCDF_DEVINL void mbar_wait(unsigned long long* b, u32 parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "W5_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra W5_DONE;\n\t"
      "bra W5_WAIT;\n\t"
      "W5_DONE:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
DCDF_DEVINL void bulk_g2s(void* smem, const void* g, u32 bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)),
               "l"(g), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}

template <typename V>
__global__ void __launch_bounds__(DT_THREADS, sizeof(V) == 4 ? 4 : 2) k_window_tiles5(const TileWindowParams P) {
  extern __shared__ __align__(16) unsigned char dt5_smem_raw[];
  typedef Tile5Smem<V> SM;
  SM& S = *reinterpret_cast<SM*>(dt5_smem_raw);
  const QuerySet& Q = P.Q;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < W5_SLOTS; s++) { mbar_init(&S.full[s], 1u); mbar_init(&S.empty[s], (u32)DT_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  u32 g = 0;  // items (structures) this CTA has pushed through the ring; item g uses slot g % 3 for the (g / 3)-th time
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
    O.base = 0;
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
    const u32 snap0 = dir[ti0].snap;
    const u32 pre = snap0 != ti0 ? 1u : 0u;  // the window starts inside a block: its snapshot is expanded first, not emitted
    const u32 n_items = n_t + pre;
    // item j of this job -> directory index
    auto item_dir = [&](u32 j) { return (pre && j == 0) ? snap0 : ti0 + j - pre; };
    // producer (thread 0): bulk copies of item `gi` = job item j
    u32 noff = 0, nsize = 0;
    auto issue = [&](u32 gi, u32 j, u32 off, u32 size) {
      const u32 sl = gi % W5_SLOTS, use = gi / W5_SLOTS;
      if (use > 0) mbar_wait(&S.empty[sl], (use - 1u) & 1u);  // every warp is done with the slot's previous structure
      const u8* src = chunk + off;
      const u32 mis = (u32)((uintptr_t)src & 15u);
      const bool fits = size + mis + 4u <= (u32)SM::BUF + 32u;
      const u32 n = fits ? ((size + mis + 4u + 15u) & ~15u) : 0u;
      mbar_arrive_expect_tx(&S.full[sl], (u32)sizeof(InstDir) + n);
      bulk_g2s(&S.slot[sl].dir, dir + item_dir(j), (u32)sizeof(InstDir), &S.full[sl]);
      if (n) bulk_g2s(S.slot[sl].stage, src - mis, n, &S.full[sl]);
    };
    if (tid == 0) {
      const u32 d0 = item_dir(0);
      issue(g, 0, dir[d0].off, dir[d0].size);
      if (n_items > 1) { const u32 d1 = item_dir(1); noff = dir[d1].off; nsize = dir[d1].size; }
    }
    const u64 t_stride = (u64)(W_rows * W_cols);
    u64 tbase = obase + (u64)((t_lo - c.start) * W_rows * W_cols + tile_org);
    for (u32 j = 0; j < n_items; j++, g++) {
      if (tid == 0 && j + 1 < n_items) {
        issue(g + 1, j + 1, noff, nsize);
        if (j + 2 < n_items) { const u32 d2 = item_dir(j + 2); noff = dir[d2].off; nsize = dir[d2].size; }
      }
      const u32 sl = g % W5_SLOTS;
      mbar_wait(&S.full[sl], (g / W5_SLOTS) & 1u);
      const InstDir& D = S.slot[sl].dir;
      const bool is_pre = pre && j == 0;
      const bool is_snap = is_pre || D.snap == item_dir(j);
      O.base = tbase;
      if (!is_pre) tbase += t_stride;
      u32 delta;
      const u32 mis = (u32)((uintptr_t)(chunk + D.off) & 15u);
      delta = mis - D.off;
      if (D.size + mis + 4u <= (u32)SM::BUF + 32u) {
        instant4<V, SM>(S.slot[sl].stage + (int32_t)delta, D, is_snap, !is_pre, L, S, O);
      } else {
        const QuadOut O2 = O;  // only the copy has its address taken
        instant4_global<V, SM>(chunk, &D, is_snap, !is_pre, L, &S, &O2);
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&S.empty[sl]);
    }
  }
}

}  // namespace dcdf
