// tree_geo.hpp -- static geometry of Superchunk::build's recursion (superchunk.rs:88-181), shared by the encode entry
// points and by dcdf_superchunk_open (host code only).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "host.hpp"

namespace dcdf {

// sidelen exponent: ceil(log_k(longest)) computed in f64 exactly as the reference does
// (snapshot.rs:118-119, superchunk.rs:96-101).
inline uint32_t levels_for(int64_t longest, int k) {
  double l = std::ceil(std::log((double)longest) / std::log((double)k));
  if (!(l >= 0)) l = 0;
  return (uint32_t)l;
}

struct TreeGeo {
  std::vector<TreeNode> nodes;
  std::vector<TreeChild> children;
  std::vector<dcdf_superchunk::NodeGeom> geom;
  std::vector<int32_t> leaf_unit;
  int leaf_rows = 0, leaf_cols = 0, leaf_side = 1;
  int64_t leaf_grid = 1;
  uint64_t tbl_per_instant = 0;
};

// Static node tree of Superchunk::build's recursion (superchunk.rs:88-181) for a rows x cols raster.
inline void build_tree(TreeGeo& G, int64_t rows, int64_t cols, const uint32_t* levels, uint32_t n_levels) {
  const int D = (int)n_levels - 1;
  std::vector<uint32_t> suffix(n_levels + 1, 0);  // suffix[d] = sum of levels[d..]
  for (int d = (int)n_levels - 1; d >= 0; d--) suffix[d] = suffix[d + 1] + levels[d];
  const int ls = 1 << levels[n_levels - 1];
  G.leaf_side = ls;
  G.leaf_rows = (int)((rows + ls - 1) / ls);
  G.leaf_cols = (int)((cols + ls - 1) / ls);
  G.leaf_grid = ((int64_t)1 << suffix[0]) / ls;
  G.leaf_unit.assign((size_t)G.leaf_rows * G.leaf_cols, -1);
  for (int i = 0; i < G.leaf_rows * G.leaf_cols; i++) G.leaf_unit[i] = i;
  struct Pending { int depth; int64_t top, left; int parent; u32 child_slot; };
  std::vector<Pending> queue;
  queue.push_back({0, 0, 0, -1, 0});
  for (size_t qi = 0; qi < queue.size(); qi++) {
    const Pending pd = queue[qi];
    const int d = pd.depth;
    const int64_t side = (int64_t)1 << suffix[d];
    const int64_t nrows = std::min(side, rows - pd.top), ncols = std::min(side, cols - pd.left);
    TreeNode nd;
    nd.parent = pd.parent; nd.depth = d;
    nd.first_child = (u32)G.children.size();
    const int64_t sub = (int64_t)1 << levels[d];
    nd.n_children = (u32)(sub * sub);
    nd.tbl_off = (u32)G.tbl_per_instant;
    nd.levels_ok = levels_for(std::max(nrows, ncols), 2) == suffix[d] ? 1 : 0;
    G.tbl_per_instant += nd.n_children;
    const int id = (int)G.nodes.size();
    if (pd.parent >= 0) G.children[G.nodes[pd.parent].first_child + pd.child_slot].index = id;
    G.nodes.push_back(nd);
    const int64_t cs = side / sub;
    G.geom.push_back({pd.top, pd.left, nrows, ncols, side, cs, sub, levels[d]});
    G.children.resize(G.children.size() + nd.n_children);
    for (int64_t r = 0; r < sub; r++)
      for (int64_t c = 0; c < sub; c++) {
        TreeChild ch;
        memset(&ch, 0, sizeof ch);
        const int64_t ctop = pd.top + r * cs, cleft = pd.left + c * cs;
        const u32 slot = (u32)(r * sub + c);
        if (ctop >= rows || cleft >= cols) {
          ch.kind = 0;
        } else {
          const int64_t cr = std::min(cs, rows - ctop), cc = std::min(cs, cols - cleft);
          ch.gr0 = (int)(ctop / ls); ch.gc0 = (int)(cleft / ls);
          ch.gr1 = (int)((ctop + cr + ls - 1) / ls); ch.gc1 = (int)((cleft + cc + ls - 1) / ls);
          const bool at_bottom = d == D - 1;
          bool as_chunk = at_bottom;
          if (!at_bottom && levels_for(std::max(cr, cc), 2) <= levels[d + 1]) {  // superchunk.rs:153-163
            if (cr > ls || cc > ls)
              api_fail(DCDF_ERR_BAD_ARG, "a clipped region is demoted to a Chunk larger than the leaf subchunks; not built on the GPU yet");
            as_chunk = true;
          }
          if (as_chunk) {
            ch.kind = 1;
            ch.index = ch.gr0 * G.leaf_cols + ch.gc0;
          } else {
            ch.kind = 2;
            ch.index = -1;
            queue.push_back({d + 1, ctop, cleft, id, slot});
          }
        }
        G.children[nd.first_child + slot] = ch;
      }
  }
}


}  // namespace dcdf
