// tree.cuh -- the reference-shaped KD tree on the device (topology only) and the pieces that
// need it inside other kernels.
//
// The reference's tree (src/kdtree.c:47-62) is an insertion-order BST over the log entries:
// entry s descends from the root, at depth d compares coordinate d % K, goes to the
// "strictly less" child or the "greater or equal" child, and becomes a leaf where the slot is
// empty.  We keep exactly that shape as two u32 child links per log entry (`child[2*s+side]`),
// the kd-points themselves stay where the scan reads them.  With it:
//   * tree_nearest_kernel replays kdtree_nearest_rec (src/kdtree.c:131-162) visit for visit;
//   * resolve_tie() picks, among entries at exactly equal distance, the one the reference's
//     near-side-first traversal reaches first.
#pragma once
#include "common.cuh"

namespace svdb {

constexpr uint32_t NODE_NONE = 0xffffffffu;

// Which of `nt` log entries, all at the SAME reference distance from q and all present in the
// tree, does kdtree_nearest_rec reach first?  (Strict '<' at kdtree.c:139 keeps the first.)
// The first minimum in near-first preorder is never pruned: it lies beyond a splitting plane
// only if its own distance is >= the plane distance, and the running best is still larger.
// Walk down from the root keeping the tied entries of the current subtree: a tied entry that
// IS the current node wins (visited before its descendants); otherwise continue into the near
// child if any tied entry lives there, else into the far child.  Called by one lane.
__device__ inline u64 resolve_tie(const double *__restrict__ pts, int stride, int K,
                                  const uint32_t *__restrict__ child, const double *__restrict__ q,
                                  u64 *tied, int nt) {
    uint32_t node = 0;
    int depth = 0;
    while (nt > 1) {
        for (int j = 0; j < nt; j++)
            if (tied[j] == (u64)node) return tied[j];
        const int cd = depth % K;
        const double pn = pts[(u64)node * stride + cd];
        const int near_side = (q[cd] < pn) ? 0 : 1;                 // kdtree.c:147-155
        int n_near = 0;
        for (int j = 0; j < nt; j++) n_near += ((pts[tied[j] * stride + cd] < pn) ? 0 : 1) == near_side;
        const int side = n_near > 0 ? near_side : 1 - near_side;
        int m = 0;
        for (int j = 0; j < nt; j++)
            if (((pts[tied[j] * stride + cd] < pn) ? 0 : 1) == side) tied[m++] = tied[j];
        nt = m;
        node = child[2 * (u64)node + side];
        depth++;
        if (node == NODE_NONE) break;                                // cannot happen for entries in the tree
    }
    return tied[0];
}

}  // namespace svdb
