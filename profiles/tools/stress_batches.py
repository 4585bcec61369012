"""Randomised cross-check of the batch kernels: counts of value-range searches from the shared-decode pass
(k_count_tiles4) against the per-window kernel, cell series from the tile decoder (k_cell_tiles4) against the per-cell
walks and the input, on random rasters (dtype, NaNs, uniform tiles, constant offsets between instants -> single-node and
`equal` Logs, nested trees, short time slices).  Prints one line per case; exits non-zero on the first mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from dcdf_b200 import Context, Superchunk

def raster(rng, T, R, C, kind):
    base = rng.integers(-50, 50, (R, C))
    frames = []
    for t in range(T):
        mode = rng.integers(0, 6)
        if mode == 0: f = base + int(rng.integers(-5, 6))            # equal to the previous pattern plus a constant
        elif mode == 1: f = np.full((R, C), int(rng.integers(-9, 9)))  # uniform instant
        elif mode == 2: base = rng.integers(-50, 50, (R, C)); f = base
        else:
            f = base.copy(); k = int(rng.integers(1, 200))
            f[rng.integers(0, R, k), rng.integers(0, C, k)] += rng.integers(-3, 4, k)
        frames.append(f)
    a = np.stack(frames).astype(np.int64)
    for _ in range(int(rng.integers(0, 4))):                         # uniform / slowly changing tiles
        r0, c0 = int(rng.integers(0, R)), int(rng.integers(0, C))
        a[:, r0:r0 + 64, c0:c0 + 64] = (np.arange(T) // 3)[:, None, None]
    if kind == "f32":
        a = (a / 4).astype(np.float32)
        m = rng.random(a.shape) < 0.01
        a[m] = np.nan
        a[:, : R // 5, : C // 7] = np.nan
        return a
    if kind == "i32": return a.astype(np.int32)
    return a * (1 << 33)                                             # 64-bit expansion

def run(seed, n_cases):
    rng = np.random.default_rng(seed)
    ctxs = {}
    for name, opts in (("shared", {"search_share_min": 1, "cell_tile_min": 1}), ("plain", {"search_share_min": 0, "cell_tile_min": 0})):
        c = Context(0)
        for k, v in opts.items(): c.set_option(k, v)
        ctxs[name] = c
    for case in range(n_cases):
        kind = ("f32", "i32", "i64")[case % 3]
        levels = [[1, 6], [2, 6], [1, 1, 6], [2, 5], [1, 4]][int(rng.integers(0, 5))]
        side = 2 ** sum(levels)
        R, C = int(rng.integers(side // 2 + 1, side + 1)), int(rng.integers(side // 2 + 1, side + 1))
        T, cs = int(rng.integers(5, 40)), int(rng.integers(3, 17))
        data = raster(rng, T, R, C, kind)
        kw = dict(compute_bits=True) if kind == "f32" else {}
        try:
            scs = {n: Superchunk.build(c, data, levels, chunk_size=cs, **kw) for n, c in ctxs.items()}
        except Exception as ex:   # e.g. a clipped nested region that needs other levels (the reference raises the same error)
            print(f"case {case}: {kind} levels {levels} {T}x{R}x{C}: skipped ({str(ex)[:60]})")
            continue
        nw = 400
        t0 = rng.integers(0, T, nw); t1 = np.minimum(T, t0 + rng.integers(1, T + 1, nw))
        r0 = rng.integers(0, R, nw); r1 = np.minimum(R, r0 + rng.integers(1, 130, nw))
        c0 = rng.integers(0, C, nw); c1 = np.minimum(C, c0 + rng.integers(1, 130, nw))
        cubes = np.stack([t0, t1, r0, r1, c0, c1], axis=1)
        raw = scs["plain"].window(0, T, 0, R, 0, C, raw=True)
        vals = np.unique(raw)
        lo = rng.choice(vals, nw) - rng.integers(0, 3, nw)
        hi = lo + rng.integers(0, max(2, int((vals.max() - vals.min()) // 8) + 1), nw)
        ca, _ = scs["shared"].search_batch(cubes, lo, hi, want_cells=False)
        cb, _ = scs["plain"].search_batch(cubes, lo, hi, want_cells=False)
        ok_s = bool(np.array_equal(ca, cb))
        nq = 3000
        qs = rng.integers(0, T, nq); qe = np.minimum(T, qs + rng.integers(0, T + 1, nq))
        q = np.stack([qs, qe, rng.integers(0, R, nq), rng.integers(0, C, nq)], axis=1)
        fa, _ = scs["shared"].cell_batch(q, flat=True); fb, _ = scs["plain"].cell_batch(q, flat=True)
        ok_c = bool(np.array_equal(fa, fb, equal_nan=True))
        want = np.concatenate([data[a:b, r, c] for a, b, r, c in q]) if nq else fa
        ok_i = bool(np.array_equal(fa, want, equal_nan=True))
        # windows and single cells against the input
        wflat, woff = scs["plain"].window_batch(cubes[:60])
        ok_w = all(bool(np.array_equal(wflat[int(woff[i]):int(woff[i + 1])], data[a:b, r_a:r_b, c_a:c_b].ravel(), equal_nan=True))
                   for i, (a, b, r_a, r_b, c_a, c_b) in enumerate(cubes[:60]))
        irc = np.stack([rng.integers(0, T, 2000), rng.integers(0, R, 2000), rng.integers(0, C, 2000)], axis=1)
        ok_g = bool(np.array_equal(scs["plain"].get_batch(irc), data[irc[:, 0], irc[:, 1], irc[:, 2]], equal_nan=True))
        ok_i = ok_i and ok_w and ok_g
        # hits WITH cells (count + write passes of the per-window kernel) against the oracle, order included: windows
        # inside one time slice, two-level trees (for nested ones the reference's order is region-major)
        ok_o = True
        if len(levels) == 2:
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
            import oracle_lib as orc
            sl = int(rng.integers(0, (T + cs - 1) // cs))
            s0, s1 = sl * cs, min(T, (sl + 1) * cs)
            ref = orc.superchunk_build(data[s0:s1], levels, **({"compute_bits": True} if kind == "f32" else {}))
            k = 12
            a_ = rng.integers(0, s1 - s0, k); b_ = np.minimum(s1 - s0, a_ + rng.integers(1, s1 - s0 + 1, k))
            loc = np.stack([a_, b_, r0[:k], r1[:k], c0[:k], c1[:k]], axis=1)
            glob = loc.copy(); glob[:, 0] += s0; glob[:, 1] += s0
            gc, gcells = scs["plain"].search_batch(glob, lo[:k], hi[:k])
            rc, rcells, _ = ref.search_batch(loc, lo[:k], hi[:k])
            if rcells is not None and len(rcells):
                rcells = rcells.copy(); rcells[:, 0] += s0
            ok_o = gc.tolist() == rc.tolist() and (int(gc.sum()) == 0 or bool(np.array_equal(gcells, rcells)))
        print(f"case {case}: {kind} levels {levels} {T}x{R}x{C} chunk_size {cs}: search {ok_s} ({int(ca.sum())} hits)  cells {ok_c} input {ok_i} oracle order {ok_o}", flush=True)
        ok_s = ok_s and ok_o
        if not (ok_s and ok_c and ok_i):
            bad = np.nonzero(ca != cb)[0][:5]
            print("first differing windows", bad, cubes[bad].tolist(), lo[bad].tolist(), hi[bad].tolist(), "shared", ca[bad].tolist(), "per window", cb[bad].tolist())
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
            import oracle_lib as orc
            for w in bad[:2]:
                cu = cubes[w]
                s0 = int(cu[0]) // cs
                tot = 0
                for sl in range(int(cu[0]) // cs, (int(cu[1]) - 1) // cs + 1):
                    ref = orc.superchunk_build(data[sl * cs:(sl + 1) * cs], levels, **({"compute_bits": True} if kind == "f32" else {}))
                    a_, b_ = max(int(cu[0]), sl * cs) - sl * cs, min(int(cu[1]), (sl + 1) * cs) - sl * cs
                    rc, _, _ = ref.search_batch([[a_, b_, int(cu[2]), int(cu[3]), int(cu[4]), int(cu[5])]], int(lo[w]), int(hi[w]), want_cells=False)
                    tot += int(rc[0])
                    for name in ("shared", "plain"):
                        g, _ = scs[name].search_batch([[sl * cs + a_, sl * cs + b_, int(cu[2]), int(cu[3]), int(cu[4]), int(cu[5])]], int(lo[w]), int(hi[w]), want_cells=False)
                        print("  window", w, "slice", sl, name, int(g[0]), "oracle", int(rc[0]))
                    # per instant and tile
                    for t in range(a_, b_):
                        for (ra, rb) in ((int(cu[2]), min(int(cu[3]), 128)), (max(int(cu[2]), 128), int(cu[3]))):
                            for (ca_, cb_) in ((int(cu[4]), min(int(cu[5]), 128)), (max(int(cu[4]), 128), int(cu[5]))):
                                if rb <= ra or cb_ <= ca_: continue
                                one = [[t, t + 1, ra, rb, ca_, cb_]]
                                o, _, _ = ref.search_batch(one, int(lo[w]), int(hi[w]), want_cells=False)
                                one_g = [[sl * cs + t, sl * cs + t + 1, ra, rb, ca_, cb_]]
                                g1, _ = scs["shared"].search_batch(one_g, int(lo[w]), int(hi[w]), want_cells=False)
                                g2, _ = scs["plain"].search_batch(one_g, int(lo[w]), int(hi[w]), want_cells=False)
                                if not (int(o[0]) == int(g1[0]) == int(g2[0])):
                                    np.set_printoptions(linewidth=250)
                                    print("band", int(lo[w]), int(hi[w]))
                                    for tt in range(max(0, sl * cs + t - 3), sl * cs + t + 1):
                                        print("raw t", tt); print(raw[tt, ra:rb, ca_:cb_])
                                    oc, ocells, _ = ref.search_batch(one, int(lo[w]) - 1000, int(hi[w]) + 1000)
                                    print("oracle cells in a wide band:", int(oc[0]))
                                    rr, cc2 = np.nonzero((raw[sl * cs + t, ra:rb, ca_:cb_] >= int(lo[w])) & (raw[sl * cs + t, ra:rb, ca_:cb_] <= int(hi[w])))
                                    for r_, c_ in zip(rr, cc2):
                                        o1, _, _ = ref.search_batch([[t, t + 1, ra + int(r_), ra + int(r_) + 1, ca_ + int(c_), ca_ + int(c_) + 1]], int(lo[w]), int(hi[w]), want_cells=False)
                                        o2, _, _ = ref.search_batch([[t, t + 1, ra + int(r_), ra + int(r_) + 1, ca_ + int(c_), ca_ + int(c_) + 1]], -100000, 100000, want_cells=False)
                                        print("   cell", int(r_), int(c_), "value", int(raw[sl * cs + t, ra + r_, ca_ + c_]), "oracle single-cell query in band:", int(o1[0]), "any band:", int(o2[0]))
                                    kinds, child = ref.node_refs(0)
                                    info0 = ref.node_info(0)
                                    slot_i = (ra // info0.chunks_sidelen) * info0.subsidelen + ca_ // info0.chunks_sidelen
                                    print("   superchunk bits", info0.fractional_bits, "slot", slot_i, "kind", int(kinds[slot_i]), "child", int(child[slot_i]))
                                    v_, b_bits = ref.get_batch([[t, ra, ca_ + 5]])
                                    print("   get(t, 0, 5) =", int(v_[0]), "bits", int(b_bits[0]))
                                    if int(child[slot_i]) >= 0:
                                        chh = ref.chunk(int(child[slot_i]))
                                        print("   chunk-level search at t:", len(chh.search(t, t + 1, 0, rb - ra, 0, cb_ - ca_, int(lo[w]), int(hi[w]))),
                                              " blocks", chh.block_instants())
                                    for tt in range(max(0, t - 3), t + 1):
                                        o3, _, _ = ref.search_batch([[tt, tt + 1, ra, rb, ca_, cb_]], int(lo[w]), int(hi[w]), want_cells=False)
                                        o4, _, _ = ref.search_batch([[tt, tt + 1, ra, rb, ca_, cb_]], int(lo[w]), 100000, want_cells=False)
                                        o5, _, _ = ref.search_batch([[tt, tt + 1, ra, rb, ca_, cb_]], -100000, int(hi[w]), want_cells=False)
                                        print("   oracle at t", tt, "band:", int(o3[0]), " [lo, inf):", int(o4[0]), " (-inf, hi]:", int(o5[0]))
                                    print("    t", sl * cs + t, "rect", ra, rb, ca_, cb_, "oracle", int(o[0]), "shared", int(g1[0]), "per window", int(g2[0]),
                                          "blocks", scs["plain"].block_instants() if hasattr(scs["plain"], "block_instants") else "")
                print("  oracle total", tot)
            return False
        for s in scs.values(): s.close()
    for c in ctxs.values(): c.close()
    print("all cases agree")
    return True

if __name__ == "__main__":
    ok = run(int(sys.argv[1]) if len(sys.argv) > 1 else 1, int(sys.argv[2]) if len(sys.argv) > 2 else 12)
    sys.exit(0 if ok else 1)
