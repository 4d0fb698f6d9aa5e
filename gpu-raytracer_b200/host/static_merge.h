// Host-side helpers of the static merge (see rebuild_static_merge in csrc/ptb_api.cu): pure functions over 80-byte CWBVH
// nodes (Src/BVH/BVH.h:61-80 layout: p[12] e[3] imask | base_child base_triangle meta[8] | quantised boxes).
// Shared by libptb.so (the product) and libptb_host.so (which exports C wrappers so the CPU test suite can exercise them).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace ptb_merge {

struct NodeView {
    const unsigned char* n;
    unsigned imask() const { return n[15]; }
    unsigned base_child() const { unsigned v; std::memcpy(&v, n + 16, 4); return v; }
    unsigned base_triangle() const { unsigned v; std::memcpy(&v, n + 20, 4); return v; }
    unsigned meta(int k) const { return n[24 + k]; }
};

// Every primitive index referenced by the leaves below `root` (a BLAS: triangles; a TLAS: instances), in traversal order.
inline void collect_leaf_primitives(const unsigned char* nodes, unsigned root, std::vector<int>& out) {
    std::vector<unsigned> stack{ root };
    while (!stack.empty()) {
        NodeView v{ nodes + (size_t)stack.back() * 80 }; stack.pop_back();
        unsigned internal = 0;
        for (int k = 0; k < 8; k++) {
            unsigned meta = v.meta(k);
            if (v.imask() & (1u << k)) { unsigned child = v.base_child() + internal++; if (meta) stack.push_back(child); continue; }
            if (!meta) continue;
            unsigned count = (unsigned)__builtin_popcount(meta >> 5), first = meta & 31u;
            for (unsigned t = 0; t < count; t++) out.push_back(int(v.base_triangle() + first + t));
        }
    }
}

// Longest root-to-leaf path, in nodes.  The traversal keeps at most two stack entries per level (the rest of a node group and a
// postponed triangle group), so 2 * depth + 2 entries bound its stack.
inline int max_depth(const unsigned char* nodes, unsigned root) {
    struct Item { unsigned node; int depth; };
    std::vector<Item> stack{ { root, 1 } };
    int deepest = 0;
    while (!stack.empty()) {
        Item it = stack.back(); stack.pop_back();
        if (it.depth > deepest) deepest = it.depth;
        NodeView v{ nodes + (size_t)it.node * 80 };
        unsigned internal = 0;
        for (int k = 0; k < 8; k++) if (v.imask() & (1u << k)) { unsigned child = v.base_child() + internal++; if (v.meta(k)) stack.push_back({ child, it.depth + 1 }); }
    }
    return deepest;
}

// Blank (meta = 0: "empty slot" to the node test) every leaf slot whose instances are all merged and every internal child whose
// whole subtree is; imask is left alone because it drives child indexing.  Returns true when nothing below `ni` is left.
inline bool prune_tlas(unsigned char* nodes, unsigned ni, const std::vector<char>& merged) {
    unsigned char* n = nodes + (size_t)ni * 80;
    NodeView v{ n };
    unsigned internal = 0; bool all = true;
    for (int k = 0; k < 8; k++) {
        unsigned meta = v.meta(k);
        if (v.imask() & (1u << k)) {
            unsigned child = v.base_child() + internal++;
            if (!meta) continue;
            if (prune_tlas(nodes, child, merged)) n[24 + k] = 0; else all = false;
            continue;
        }
        if (!meta) continue;
        unsigned count = (unsigned)__builtin_popcount(meta >> 5), first = meta & 31u;
        bool leaf_all = true;
        for (unsigned t = 0; t < count; t++) { unsigned inst = v.base_triangle() + first + t; if (inst >= merged.size() || !merged[inst]) leaf_all = false; }
        if (leaf_all) n[24 + k] = 0; else all = false;
    }
    return all;
}

// Breadth-first re-layout of a depth-first node array (root at 0, children of a node contiguous and in slot order): output node 0
// is the root and child indices become `base + new index`, so the first nodes of the array are the top levels of the tree.
inline void bfs_relayout(const unsigned char* dfs, int node_count, int base, unsigned char* bfs) {
    std::vector<int> queue{ 0 }; queue.reserve((size_t)node_count);
    int next = 1;
    for (size_t qi = 0; qi < queue.size(); qi++) {
        const unsigned char* src = dfs + (size_t)queue[qi] * 80;
        unsigned char* dst = bfs + qi * 80;
        std::memcpy(dst, src, 80);
        unsigned old_base; std::memcpy(&old_base, src + 16, 4);
        int kids = __builtin_popcount((unsigned)src[15]);
        unsigned new_base = (unsigned)(base + next);
        std::memcpy(dst + 16, &new_base, 4);
        for (int c = 0; c < kids; c++) queue.push_back((int)old_base + c);
        next += kids;
    }
}

} // namespace ptb_merge
