// tree_types.hpp -- plain structs of the static superchunk node tree (shared by host and device code).
#pragma once
#include <stdint.h>

namespace dcdf {
typedef uint32_t u32;

struct TreeNode {
  int parent;        // node index, -1 for the root
  int depth;
  u32 first_child;   // into TreeChild[]
  u32 n_children;    // subsidelen^2
  u32 tbl_off;       // per-instant offset of this node's table inside the slice's table block
  int levels_ok;     // 0: the clipped region does not need sum(sublevels) levels -> BAD_LEVELS when built
};
struct TreeChild {
  int kind;          // 0 out of bounds, 1 leaf unit, 2 node
  int index;         // unit index inside the slice / node index
  int gr0, gc0, gr1, gc1;  // leaf-tile rectangle covered by the child [gr0, gr1) x [gc0, gc1), clipped to the raster
};
struct NodeState {
  int alive;         // built (every ancestor alive and not elided)
  int bits;          // fractional bits of this node's buffer
};

}  // namespace dcdf
