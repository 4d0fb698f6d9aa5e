// CPU BVH construction for the B200 wavefront path tracer (host side, no CUDA).
//
// The north star keeps BVH *build* on the CPU ("BVH build stays the reference's CPU path"),
// so this file restates -- in our own code -- the three host algorithms whose OUTPUT FORMAT
// the device traversal kernels consume, and which bench.py times as the CPU baseline:
//
//   * full-sweep SAH binary builder, one primitive per leaf
//         (reference: Src/BVH/Builders/SAHBuilder.cpp:12-104, BVHPartitions.cpp:7-51,
//          presort Src/Core/Sort.h:142-202)
//   * SAH leaf collapser for the plain BVH2 variant (config 1)
//         (reference: Src/BVH/BVHCollapser.cpp:14-114)
//   * BVH2 -> 80-byte CWBVH (BVH8) conversion: 7-entry DP cost table, greedy octant
//     slot assignment, quantised child boxes
//         (reference: Src/BVH/Converters/BVH8Converter.cpp:7-335, node Src/BVH/BVH.h:61-80)
//
// Exposed through a small C ABI (ptbh_*) that the Python scene front-end drives via ctypes;
// one call per mesh so the caller can fan the meshes out over a thread pool exactly like the
// reference's AssetManager does (Src/Assets/AssetManager.cpp:57-95).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <algorithm>
#include <thread>
#include <vector>
#include "static_merge.h"

namespace {

struct V3 { float x, y, z; };
inline V3 vmin(V3 a, V3 b) { return { std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z) }; }
inline V3 vmax(V3 a, V3 b) { return { std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z) }; }

struct Box {
    V3 lo, hi;
    static Box empty() {
        const float inf = std::numeric_limits<float>::infinity();
        return { { inf, inf, inf }, { -inf, -inf, -inf } };
    }
    void grow(const Box& b) { lo = vmin(lo, b.lo); hi = vmax(hi, b.hi); }
    void grow(V3 p) { lo = vmin(lo, p); hi = vmax(hi, p); }
    bool is_empty() const { return lo.x == std::numeric_limits<float>::infinity(); }
    // half surface area * 2, same operation order as the reference's AABB::surface_area
    float area() const {
        float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
        return 2.0f * (dx * dy + dy * dz + dz * dx);
    }
    V3 center() const { return { (lo.x + hi.x) * 0.5f, (lo.y + hi.y) * 0.5f, (lo.z + hi.z) * 0.5f }; }
    // widen degenerate (flat) boxes, Src/Math/AABB.h fix_if_needed
    void fatten(float epsilon = 0.001f) {
        if (is_empty()) return;
        float* mn = &lo.x; float* mx = &hi.x;
        for (int d = 0; d < 3; d++) {
            float eps = epsilon;
            while (mx[d] - mn[d] < eps) { mn[d] -= eps; mx[d] += eps; eps *= 2.0f; }
        }
    }
};

// 32-byte binary node, bit-compatible with the device's BVH2 node (Src/BVH/BVH.h:11-23)
struct Node2 {
    Box      box;
    int32_t  left_or_first;
    uint32_t count_axis; // count : 30 (low bits), axis : 2 (high bits)
    uint32_t count() const { return count_axis & 0x3fffffffu; }
    uint32_t axis()  const { return count_axis >> 30; }
    bool     leaf()  const { return count() > 0; }
    void set(uint32_t count, uint32_t axis) { count_axis = (count & 0x3fffffffu) | (axis << 30); }
};
static_assert(sizeof(Node2) == 32, "BVH2 node must be 32 bytes");

// 80-byte compressed wide node (Src/BVH/BVH.h:61-80)
struct Node8 {
    V3       p;
    uint8_t  e[3];
    uint8_t  imask;
    uint32_t base_child;
    uint32_t base_triangle;
    uint8_t  meta[8];
    uint8_t  qlo_x[8], qhi_x[8];
    uint8_t  qlo_y[8], qhi_y[8];
    uint8_t  qlo_z[8], qhi_z[8];
};
static_assert(sizeof(Node8) == 80, "CWBVH node must be 80 bytes");

struct Prims {
    std::vector<Box> box;
    std::vector<V3>  center;
};

inline uint32_t float_sort_key(float f) {
    uint32_t u; std::memcpy(&u, &f, 4);
    uint32_t mask = uint32_t(-int32_t(u >> 31)) | 0x80000000u;
    return u ^ mask;
}

// LSD radix sort of indices by float key, 4 passes of 8 bits (stable, like the reference's presort)
void radix_sort_indices(std::vector<int>& idx, const std::vector<uint32_t>& key) {
    size_t n = idx.size();
    if (n <= 1) return;
    std::vector<int> tmp(n);
    int* in = idx.data(); int* out = tmp.data();
    for (int pass = 0; pass < 4; pass++) {
        uint32_t hist[256] = {};
        int shift = pass * 8;
        for (size_t i = 0; i < n; i++) hist[(key[in[i]] >> shift) & 255u]++;
        uint32_t sum = 0;
        for (int b = 0; b < 256; b++) { uint32_t c = hist[b]; hist[b] = sum; sum += c; }
        for (size_t i = 0; i < n; i++) out[hist[(key[in[i]] >> shift) & 255u]++] = in[i];
        std::swap(in, out);
    }
    // 4 passes -> result is back in idx
}

struct SAHBuilder {
    const Prims& prims;
    std::vector<Node2>& nodes;
    std::vector<int> order[3];
    std::vector<float> sweep;
    std::vector<int>   scratch;
    std::vector<uint8_t> goes_left;

    SAHBuilder(const Prims& p, std::vector<Node2>& n) : prims(p), nodes(n) {}

    struct Split { int index; int dim; float cost; Box left, right; };

    Split find_split(int first, int count) {
        Split s; s.index = -1; s.dim = -1; s.cost = std::numeric_limits<float>::infinity();
        s.left = Box::empty(); s.right = Box::empty();
        for (int dim = 0; dim < 3; dim++) {
            const int* ord = order[dim].data();
            Box l = Box::empty(), r = Box::empty();
            for (int i = 1; i < count; i++) {
                l.grow(prims.box[ord[first + i - 1]]);
                sweep[i] = l.area() * float(i);
            }
            for (int i = count - 1; i > 0; i--) {
                r.grow(prims.box[ord[first + i]]);
                float cost = sweep[i] + r.area() * float(count - i);
                if (cost <= s.cost) { s.cost = cost; s.index = first + i; s.dim = dim; s.right = r; }
            }
        }
        const int* ord = order[s.dim].data();
        for (int i = first; i < s.index; i++) s.left.grow(prims.box[ord[i]]);
        return s;
    }

    void build_node(int node_index, int first, int count) {
        if (count == 1) { // always split down to one primitive per leaf
            nodes[node_index].left_or_first = first;
            nodes[node_index].set(1, 0);
            return;
        }
        Split s = find_split(first, count);
        const int* ord = order[s.dim].data();
        for (int i = first; i < s.index; i++)         goes_left[ord[i]] = 1;
        for (int i = s.index; i < first + count; i++) goes_left[ord[i]] = 0;
        for (int dim = 0; dim < 3; dim++) {
            if (dim == s.dim) continue;
            int l = 0, r = s.index - first;
            int* o = order[dim].data();
            for (int i = first; i < first + count; i++) {
                int id = o[i];
                if (goes_left[id]) scratch[l++] = id; else scratch[r++] = id;
            }
            std::memcpy(o + first, scratch.data(), size_t(count) * sizeof(int));
        }
        int left = int(nodes.size());
        nodes[node_index].left_or_first = left;
        nodes[node_index].set(0, uint32_t(s.dim));
        nodes.emplace_back(); nodes.emplace_back();
        nodes[left].box = s.left; nodes[left + 1].box = s.right;
        int nl = s.index - first;
        build_node(left, first, nl);
        build_node(left + 1, first + nl, count - nl);
    }

    void build(std::vector<int>& indices_out) {
        int n = int(prims.box.size());
        nodes.clear(); nodes.reserve(size_t(2) * n + 2);
        nodes.emplace_back(); nodes.emplace_back(); // root + dummy so siblings stay paired
        Box root = Box::empty();
        for (int i = 0; i < n; i++) root.grow(prims.box[i]);
        nodes[0].box = root;
        std::vector<uint32_t> key(n);
        for (int dim = 0; dim < 3; dim++) {
            order[dim].resize(n);
            for (int i = 0; i < n; i++) {
                order[dim][i] = i;
                key[i] = float_sort_key((&prims.center[i].x)[dim]);
            }
            radix_sort_indices(order[dim], key);
        }
        sweep.resize(n + 1); scratch.resize(n); goes_left.assign(n, 0);
        build_node(0, 0, n);
        indices_out = order[0];
    }
};

struct BVH2 { std::vector<Node2> nodes; std::vector<int> indices; };
struct BVH8 { std::vector<Node8> nodes; std::vector<int> indices; };

// 128-byte 4-wide node (Src/BVH/BVH.h:25-59): child boxes as structure of arrays, then (index, count) per child.
// count > 0: leaf (index = first primitive), count == 0: internal (index = node), count == -1: unused slot.
struct Node4 {
    float   lo_x[4], lo_y[4], lo_z[4], hi_x[4], hi_y[4], hi_z[4];
    struct { int32_t index, count; } child[4];
    int used() const { int n = 0; while (n < 4 && child[n].count != -1) n++; return n; }
};
static_assert(sizeof(Node4) == 128, "BVH4 node must be 128 bytes");
struct BVH4 { std::vector<Node4> nodes; std::vector<int> indices; };

// ---------------------------------------------------------------- BVH2 -> BVH4
// The reference's 4-wide tree (Src/BVH/Converters/BVH4Converter.cpp:3-148): node i keeps the index of BVH2 node i and first holds the
// boxes of that node's two children; then, top-down, a node repeatedly ADOPTS the children of its internal child with the largest
// surface area as long as the result still fits in four slots.  Node 1 (the BVH2 alignment dummy) becomes the entry: its slot 0 points
// at node 0, and traversal starts at (node 1, slot 0) -- for a BLAS at offset `root` in the shared array, at (root + 1, slot 0).
struct QuadConverter {
    const BVH2& in; BVH4& out;

    void set_slot(Node4& n, int slot, const Box& b, int index, int count) {
        n.lo_x[slot] = b.lo.x; n.lo_y[slot] = b.lo.y; n.lo_z[slot] = b.lo.z;
        n.hi_x[slot] = b.hi.x; n.hi_y[slot] = b.hi.y; n.hi_z[slot] = b.hi.z;
        n.child[slot].index = index; n.child[slot].count = count;
    }
    void copy_slot(Node4& dst, int d, const Node4& src, int s) {
        dst.lo_x[d] = src.lo_x[s]; dst.lo_y[d] = src.lo_y[s]; dst.lo_z[d] = src.lo_z[s];
        dst.hi_x[d] = src.hi_x[s]; dst.hi_y[d] = src.hi_y[s]; dst.hi_z[d] = src.hi_z[s];
        dst.child[d] = src.child[s];
    }
    void adopt(int ni) {
        Node4& n = out.nodes[size_t(ni)];
        while (true) {
            const int have = n.used();
            int best = -1; float best_area = -std::numeric_limits<float>::infinity();
            for (int i = 0; i < have; i++) {
                if (n.child[i].count != 0) continue;                                    // leaves cannot be opened
                const int theirs = out.nodes[size_t(n.child[i].index)].used();
                if (have + theirs - 1 > 4) continue;
                float dx = n.hi_x[i] - n.lo_x[i], dy = n.hi_y[i] - n.lo_y[i], dz = n.hi_z[i] - n.lo_z[i];
                float half_area = dx * dy + dy * dz + dz * dx;
                if (half_area > best_area) { best_area = half_area; best = i; }
            }
            if (best < 0) break;
            const Node4 victim = out.nodes[size_t(n.child[best].index)];              // copy: its slots move up into this node
            const int theirs = victim.used();
            copy_slot(n, best, victim, 0);
            for (int i = 1; i < theirs; i++) copy_slot(n, have + i - 1, victim, i);
        }
    }
    void run() {
        Node4 blank; std::memset(&blank, 0, sizeof(blank));
        for (int i = 0; i < 4; i++) blank.child[i].index = blank.child[i].count = -1;
        out.nodes.assign(in.nodes.size() < 2 ? 2 : in.nodes.size(), blank);
        out.indices = in.indices;
        for (size_t i = 0; i < in.nodes.size(); i++) {
            if (i == 1 || in.nodes[i].leaf()) continue;
            const int l = in.nodes[i].left_or_first;
            for (int c = 0; c < 2; c++) {
                const Node2& k = in.nodes[size_t(l + c)];
                if (k.leaf()) set_slot(out.nodes[i], c, k.box, k.left_or_first, int(k.count()));
                else          set_slot(out.nodes[i], c, k.box, l + c, 0);
            }
        }
        out.nodes[1].child[0].index = 0; out.nodes[1].child[0].count = 0;              // entry slot
        if (in.nodes[0].leaf()) { set_slot(out.nodes[0], 0, in.nodes[0].box, in.nodes[0].left_or_first, int(in.nodes[0].count())); return; }
        std::vector<int> todo{ 0 };
        while (!todo.empty()) {                                                         // top-down: a node is final before its children are opened
            int ni = todo.back(); todo.pop_back();
            adopt(ni);
            const Node4& n = out.nodes[size_t(ni)];
            for (int i = 0; i < n.used(); i++) if (n.child[i].count == 0) todo.push_back(n.child[i].index);
        }
    }
};

// ---------------------------------------------------------------- BVH2 leaf collapse
struct Collapser {
    const BVH2& in; BVH2& out; float c_node, c_leaf;
    std::vector<uint8_t> merge;
    struct Cost { int count; float sah; };

    Cost cost(int ni) {
        const Node2& n = in.nodes[ni];
        if (n.leaf()) return { int(n.count()), float(n.count()) * c_leaf };
        Cost l = cost(n.left_or_first), r = cost(n.left_or_first + 1);
        int total = l.count + r.count;
        float as_leaf = c_leaf * float(total);
        float as_node = c_node + (in.nodes[n.left_or_first].box.area() * l.sah +
                                  in.nodes[n.left_or_first + 1].box.area() * r.sah) / n.box.area();
        if (as_leaf < as_node) { merge[ni] = 1; return { total, as_leaf }; }
        return { total, as_node };
    }
    int gather(int ni) {
        const Node2& n = in.nodes[ni];
        if (n.leaf()) {
            for (uint32_t i = 0; i < n.count(); i++) out.indices.push_back(in.indices[n.left_or_first + i]);
            return int(n.count());
        }
        return gather(n.left_or_first) + gather(n.left_or_first + 1);
    }
    void emit(int dst, int src) {
        const Node2 n = in.nodes[src];
        out.nodes[dst].box = n.box;
        out.nodes[dst].count_axis = n.count_axis;
        if (n.leaf()) {
            out.nodes[dst].left_or_first = int(out.indices.size());
            for (uint32_t i = 0; i < n.count(); i++) out.indices.push_back(in.indices[n.left_or_first + i]);
        } else if (merge[src]) {
            int c = gather(src);
            out.nodes[dst].set(uint32_t(c), n.axis());
            out.nodes[dst].left_or_first = int(out.indices.size()) - c;
        } else {
            int l = int(out.nodes.size());
            out.nodes[dst].left_or_first = l;
            out.nodes.emplace_back(); out.nodes.emplace_back();
            emit(l, n.left_or_first);
            emit(l + 1, n.left_or_first + 1);
        }
    }
    void run() {
        merge.assign(in.nodes.size(), 0);
        cost(0);
        out.nodes.clear(); out.indices.clear();
        out.nodes.reserve(in.nodes.size()); out.indices.reserve(in.indices.size());
        out.nodes.emplace_back(); out.nodes.emplace_back();
        emit(0, 0);
    }
};

// ---------------------------------------------------------------- BVH2 -> CWBVH
struct WideConverter {
    const BVH2& in; BVH8& out;
    enum Kind : int8_t { LEAF, INTERNAL, DISTRIBUTE };
    struct Choice { Kind kind; int8_t dl, dr; float cost; };
    std::vector<Choice> table; // 7 per BVH2 node: cost of representing the subtree with <= i+1 slots

    float c_tri = 1.0f;          // relative cost of one triangle test (the reference's model: 1, BVH8Converter.cpp:24-117)
    WideConverter(const BVH2& i, BVH8& o, float tri_cost = 1.0f) : in(i), out(o), c_tri(tri_cost) {}
    Choice& at(int node, int i) { return table[size_t(node) * 7 + i]; }

    int fill_costs(int ni) {
        const Node2& n = in.nodes[ni];
        if (n.leaf()) {
            float c = n.box.area() * float(n.count()) * c_tri;
            for (int i = 0; i < 7; i++) { at(ni, i).kind = LEAF; at(ni, i).cost = c; }
            return int(n.count());
        }
        int L = n.left_or_first, R = L + 1;
        int prims = fill_costs(L) + fill_costs(R);
        const float inf = std::numeric_limits<float>::infinity();
        { // whole subtree as ONE slot: either a (<=3 triangle) leaf or an internal node with 8 slots
            float as_leaf = prims <= 3 ? float(prims) * n.box.area() * c_tri : inf;
            float best = inf; int8_t dl = -1, dr = -1;
            for (int k = 0; k < 7; k++) {
                float c = at(L, k).cost + at(R, 6 - k).cost;
                if (c < best) { best = c; dl = int8_t(k); dr = int8_t(6 - k); }
            }
            float as_internal = best + n.box.area();
            if (as_leaf < as_internal) { at(ni, 0).kind = LEAF; at(ni, 0).cost = as_leaf; }
            else                        { at(ni, 0).kind = INTERNAL; at(ni, 0).cost = as_internal; }
            at(ni, 0).dl = dl; at(ni, 0).dr = dr;
        }
        for (int i = 1; i < 7; i++) { // subtree spread over i+1 slots of the parent
            float best = at(ni, i - 1).cost; int8_t dl = -1, dr = -1;
            for (int k = 0; k < i; k++) {
                float c = at(L, k).cost + at(R, i - k - 1).cost;
                if (c < best) { best = c; dl = int8_t(k); dr = int8_t(i - k - 1); }
            }
            if (dl != -1) { at(ni, i).kind = DISTRIBUTE; at(ni, i).dl = dl; at(ni, i).dr = dr; at(ni, i).cost = best; }
            else          { at(ni, i) = at(ni, i - 1); }
        }
        return prims;
    }

    void collect(int ni, int i, int kids[8], int& nk) {
        const Node2& n = in.nodes[ni];
        if (n.leaf()) { kids[nk++] = ni; return; }
        int dl = at(ni, i).dl, dr = at(ni, i).dr;
        int L = n.left_or_first, R = L + 1;
        if (at(L, dl).kind == DISTRIBUTE) collect(L, dl, kids, nk); else kids[nk++] = L;
        if (at(R, dr).kind == DISTRIBUTE) collect(R, dr, kids, nk); else kids[nk++] = R;
    }

    // greedy slot assignment: slot s is "best" for the child furthest along octant direction s
    void assign_slots(int ni, int kids[8], int nk) {
        V3 pc = in.nodes[ni].box.center();
        float cost[8][8];
        for (int c = 0; c < nk; c++) {
            V3 cc = in.nodes[kids[c]].box.center();
            V3 d = { cc.x - pc.x, cc.y - pc.y, cc.z - pc.z };
            for (int s = 0; s < 8; s++) {
                float sx = (s & 4) ? -1.0f : 1.0f, sy = (s & 2) ? -1.0f : 1.0f, sz = (s & 1) ? -1.0f : 1.0f;
                cost[c][s] = d.x * sx + d.y * sy + d.z * sz;
            }
        }
        int slot_of[8]; bool used[8] = {};
        for (int c = 0; c < 8; c++) slot_of[c] = -1;
        for (;;) {
            float best = std::numeric_limits<float>::infinity(); int bs = -1, bc = -1;
            for (int c = 0; c < nk; c++) if (slot_of[c] < 0)
                for (int s = 0; s < 8; s++) if (!used[s] && cost[c][s] < best) { best = cost[c][s]; bs = s; bc = c; }
            if (bs < 0) break;
            used[bs] = true; slot_of[bc] = bs;
        }
        int copy[8];
        for (int i = 0; i < 8; i++) { copy[i] = kids[i]; kids[i] = -1; }
        for (int c = 0; c < nk; c++) kids[slot_of[c]] = copy[c];
    }

    int emit_leaf_prims(int ni) {
        const Node2& n = in.nodes[ni];
        if (n.leaf()) {
            for (uint32_t i = 0; i < n.count(); i++) out.indices.push_back(in.indices[n.left_or_first + i]);
            return int(n.count());
        }
        return emit_leaf_prims(n.left_or_first) + emit_leaf_prims(n.left_or_first + 1);
    }

    void emit(int dst, int src) {
        Node8 node; std::memset(&node, 0, sizeof(node));
        const Box& box = in.nodes[src].box;
        node.p = box.lo;
        const float denom = 1.0f / 255.0f;
        float ex = exp2f(ceilf(log2f((box.hi.x - box.lo.x) * denom)));
        float ey = exp2f(ceilf(log2f((box.hi.y - box.lo.y) * denom)));
        float ez = exp2f(ceilf(log2f((box.hi.z - box.lo.z) * denom)));
        float rx = 1.0f / ex, ry = 1.0f / ey, rz = 1.0f / ez;
        uint32_t u;
        std::memcpy(&u, &ex, 4); node.e[0] = uint8_t(u >> 23);
        std::memcpy(&u, &ey, 4); node.e[1] = uint8_t(u >> 23);
        std::memcpy(&u, &ez, 4); node.e[2] = uint8_t(u >> 23);

        int kids[8] = { -1, -1, -1, -1, -1, -1, -1, -1 }; int nk = 0;
        collect(src, 0, kids, nk);
        assign_slots(src, kids, nk);

        node.base_triangle = uint32_t(out.indices.size());
        node.base_child    = uint32_t(out.nodes.size());
        int n_internal = 0, n_tris = 0;
        for (int i = 0; i < 8; i++) {
            int k = kids[i];
            if (k < 0) continue;
            const Box& cb = in.nodes[k].box;
            node.qlo_x[i] = uint8_t(floorf((cb.lo.x - node.p.x) * rx));
            node.qlo_y[i] = uint8_t(floorf((cb.lo.y - node.p.y) * ry));
            node.qlo_z[i] = uint8_t(floorf((cb.lo.z - node.p.z) * rz));
            node.qhi_x[i] = uint8_t(ceilf((cb.hi.x - node.p.x) * rx));
            node.qhi_y[i] = uint8_t(ceilf((cb.hi.y - node.p.y) * ry));
            node.qhi_z[i] = uint8_t(ceilf((cb.hi.z - node.p.z) * rz));
            if (at(k, 0).kind == LEAF) {
                int t = emit_leaf_prims(k);
                for (int j = 0; j < t; j++) node.meta[i] |= uint8_t(1u << (j + 5)); // unary count in top 3 bits
                node.meta[i] |= uint8_t(n_tris);                                       // offset in low 5 bits
                n_tris += t;
            } else {
                node.meta[i] = uint8_t((i + 24) | 0x20);
                node.imask |= uint8_t(1u << i);
                n_internal++;
            }
        }
        for (int i = 0; i < n_internal; i++) out.nodes.emplace_back();
        out.nodes[dst] = node;
        int off = 0;
        for (int i = 0; i < 8; i++) {
            if (kids[i] < 0) continue;
            if (node.imask & (1u << i)) emit(int(node.base_child) + off++, kids[i]);
        }
    }

    void run() {
        out.nodes.clear(); out.indices.clear();
        out.indices.reserve(in.indices.size()); out.nodes.reserve(in.nodes.size());
        out.nodes.emplace_back();
        table.resize(in.nodes.size() * 7);
        fill_costs(0);
        emit(0, 0);
    }
};

// ---------------------------------------------------------------- split BVH (spatial splits)
// SBVH of Stich et al. 2009, the algorithm of the reference's second builder (Src/BVH/Builders/SBVHBuilder.cpp:12-366: full-sweep
// object split vs. binned spatial split with reference clipping, spatial splits only where the object split's children overlap by
// more than alpha x the root area, Config.h:58).  Own formulation: recursion over per-node reference vectors (the reference keeps three
// presorted index arrays), chopped binning over 128 bins, reference unsplitting.  Output: a BVH2 with ONE reference per leaf whose
// leaf boxes are the (clipped) reference boxes -- the input the CWBVH converter above expects.  A triangle may be referenced by
// several leaves.  Used for the merged static BVH of libptb (ptb_api.cu: rebuild_static_merge), where fewer node visits and
// triangle tests per ray are worth a longer one-off build.
struct Ref { Box box; int tri; };

struct SpatialBuilder {
    const float* pos;                 // 9 floats per triangle
    float alpha = 1e-5f;
    int   bins = 128;
    float inv_root_area = 0.0f;
    std::vector<Node2> nodes;
    std::vector<int>   indices;      // triangle id per leaf, in leaf order
    size_t max_refs = 0;             // stop splitting spatially once the reference count would exceed this

    static float axis(const V3& v, int d) { return (&v.x)[d]; }
    static float& axis(V3& v, int d) { return (&v.x)[d]; }

    // bounds of (triangle clipped to the slab lo <= x[d] <= hi), intersected with `within`
    Box clip_to_slab(int tri, int d, float lo, float hi, const Box& within) const {
        const float* t = pos + size_t(tri) * 9;
        V3 v[3] = { { t[0], t[1], t[2] }, { t[3], t[4], t[5] }, { t[6], t[7], t[8] } };
        Box b = Box::empty();
        for (int i = 0; i < 3; i++) {
            const V3& a = v[i]; const V3& c = v[(i + 1) % 3];
            float pa = axis(a, d), pc = axis(c, d);
            if (pa >= lo && pa <= hi) b.grow(a);
            // edge crossings with the two planes
            for (int k = 0; k < 2; k++) {
                float plane = k == 0 ? lo : hi;
                if ((pa < plane && pc > plane) || (pa > plane && pc < plane)) {
                    float f = (plane - pa) / (pc - pa);
                    V3 q = { a.x + f * (c.x - a.x), a.y + f * (c.y - a.y), a.z + f * (c.z - a.z) };
                    axis(q, d) = plane;
                    b.grow(q);
                }
            }
        }
        if (b.is_empty()) return b;
        b.lo = vmax(b.lo, within.lo); b.hi = vmin(b.hi, within.hi);
        if (b.lo.x > b.hi.x || b.lo.y > b.hi.y || b.lo.z > b.hi.z) return Box::empty();
        return b;
    }

    struct ObjSplit { int dim = -1, index = -1; float cost = std::numeric_limits<float>::infinity(); Box left, right; };
    struct SpaSplit { int dim = -1; float plane = 0.0f, cost = std::numeric_limits<float>::infinity(); int nl = 0, nr = 0; Box left, right; };

    // Object split.  Small nodes: full sweep over the references sorted by box centre (leaves `refs` sorted along the winning axis,
    // left = the first `index`).  Large nodes: 256 centroid bins per axis (the sweep's O(n log n) sorts would dominate the build);
    // `refs` is then partitioned so that the left side comes first.
    ObjSplit object_split(std::vector<Ref>& refs, std::vector<float>& sweep) {
        const int n = int(refs.size());
        return n > sweep_limit ? object_split_binned(refs) : object_split_sweep(refs, sweep);
    }
    int sweep_limit = 4096;

    ObjSplit object_split_binned(std::vector<Ref>& refs) {
        ObjSplit best;
        const int n = int(refs.size()), B = 256;
        Box cb = Box::empty();
        for (const Ref& r : refs) cb.grow(r.box.center());
        std::vector<Box> bin_box(B), right_acc(B); std::vector<int> cnt(B);
        int best_bin = -1; float best_lo = 0.0f, best_scale = 0.0f;
        for (int d = 0; d < 3; d++) {
            float lo = axis(cb.lo, d), extent = axis(cb.hi, d) - lo;
            if (!(extent > 0.0f)) continue;
            float scale = float(B) / extent;
            for (int b = 0; b < B; b++) { bin_box[b] = Box::empty(); cnt[b] = 0; }
            for (const Ref& r : refs) {
                int b = int((axis(r.box.center(), d) - lo) * scale); b = b < 0 ? 0 : (b >= B ? B - 1 : b);
                bin_box[b].grow(r.box); cnt[b]++;
            }
            Box acc = Box::empty();
            for (int b = B - 1; b > 0; b--) { acc.grow(bin_box[b]); right_acc[b] = acc; }
            Box l = Box::empty(); int nl = 0;
            for (int b = 1; b < B; b++) {
                l.grow(bin_box[b - 1]); nl += cnt[b - 1];
                if (nl == 0 || nl == n) continue;
                float c = l.area() * float(nl) + right_acc[b].area() * float(n - nl);
                if (c < best.cost) { best.cost = c; best.dim = d; best.index = nl; best.left = l; best.right = right_acc[b]; best_bin = b; best_lo = lo; best_scale = scale; }
            }
        }
        if (best.dim < 0) {      // all centres coincide: split the list in the middle
            best.dim = 0; best.index = n / 2; best.left = Box::empty(); best.right = Box::empty();
            for (int i = 0; i < n; i++) (i < best.index ? best.left : best.right).grow(refs[i].box);
            best.cost = best.left.area() * float(best.index) + best.right.area() * float(n - best.index);
            return best;
        }
        const int d = best.dim;
        std::stable_partition(refs.begin(), refs.end(), [&](const Ref& r) {
            int b = int((axis(r.box.center(), d) - best_lo) * best_scale); b = b < 0 ? 0 : (b >= B ? B - 1 : b);
            return b < best_bin; });
        return best;
    }

    ObjSplit object_split_sweep(std::vector<Ref>& refs, std::vector<float>& sweep) {
        ObjSplit best;
        const int n = int(refs.size());
        std::vector<Ref> sorted_best;
        for (int d = 0; d < 3; d++) {
            std::stable_sort(refs.begin(), refs.end(), [d](const Ref& a, const Ref& b) {
                return axis(a.box.lo, d) + axis(a.box.hi, d) < axis(b.box.lo, d) + axis(b.box.hi, d); });
            Box l = Box::empty(), r = Box::empty();
            for (int i = 1; i < n; i++) { l.grow(refs[i - 1].box); sweep[i] = l.area() * float(i); }
            bool won = false;
            for (int i = n - 1; i > 0; i--) {
                r.grow(refs[i].box);
                float c = sweep[i] + r.area() * float(n - i);
                if (c < best.cost) { best.cost = c; best.index = i; best.dim = d; best.right = r; won = true; }
            }
            if (won && d < 2) sorted_best = refs;
            if (won && d == 2) sorted_best.clear();
        }
        if (!sorted_best.empty()) refs.swap(sorted_best);
        best.left = Box::empty();
        for (int i = 0; i < best.index; i++) best.left.grow(refs[i].box);
        return best;
    }

    SpaSplit spatial_split(const std::vector<Ref>& refs, const Box& node_box) {
        SpaSplit best;
        const int B = bins;
        std::vector<Box> bin_box(B); std::vector<int> enter(B), leave(B);
        std::vector<Box> right_acc(B);
        for (int d = 0; d < 3; d++) {
            float lo = axis(node_box.lo, d), hi = axis(node_box.hi, d);
            float extent = hi - lo;
            if (!(extent > 0.0f)) continue;
            float scale = float(B) / extent, width = extent / float(B);
            for (int b = 0; b < B; b++) { bin_box[b] = Box::empty(); enter[b] = leave[b] = 0; }
            for (const Ref& r : refs) {
                int b0 = int((axis(r.box.lo, d) - lo) * scale), b1 = int((axis(r.box.hi, d) - lo) * scale);
                b0 = b0 < 0 ? 0 : (b0 >= B ? B - 1 : b0); b1 = b1 < b0 ? b0 : (b1 >= B ? B - 1 : b1);
                if (b0 == b1) bin_box[b0].grow(r.box);
                else for (int b = b0; b <= b1; b++) {
                    Box c = clip_to_slab(r.tri, d, lo + width * float(b), b == B - 1 ? hi : lo + width * float(b + 1), r.box);
                    if (!c.is_empty()) bin_box[b].grow(c);
                }
                enter[b0]++; leave[b1]++;
            }
            Box acc = Box::empty();
            for (int b = B - 1; b > 0; b--) { acc.grow(bin_box[b]); right_acc[b] = acc; }
            Box l = Box::empty(); int nl = 0, nr = int(refs.size());
            for (int b = 1; b < B; b++) {
                l.grow(bin_box[b - 1]); nl += enter[b - 1]; nr -= leave[b - 1];
                if (nl == 0 || nr == 0 || l.is_empty() || right_acc[b].is_empty()) continue;
                float c = l.area() * float(nl) + right_acc[b].area() * float(nr);
                if (c < best.cost) { best.cost = c; best.dim = d; best.plane = lo + width * float(b); best.nl = nl; best.nr = nr; best.left = l; best.right = right_acc[b]; }
            }
        }
        return best;
    }

    void make_leaf(int node_index, const Ref& r) {
        nodes[node_index].box = r.box; nodes[node_index].box.fatten();
        nodes[node_index].left_or_first = int(indices.size());
        nodes[node_index].set(1, 0);
        indices.push_back(r.tri);
    }

    void build_node(int node_index, std::vector<Ref>& refs, const Box& box, size_t& budget) {
        const int n = int(refs.size());
        nodes[node_index].box = box;
        if (n == 1) { make_leaf(node_index, refs[0]); return; }
        std::vector<float> sweep(size_t(n) + 1);
        ObjSplit os = object_split(refs, sweep);
        std::vector<Ref> left, right;
        Box lbox, rbox; int dim = os.dim;
        bool spatial = false;
        if (n > 2 && budget > 0) {
            Box ov; ov.lo = vmax(os.left.lo, os.right.lo); ov.hi = vmin(os.left.hi, os.right.hi);
            bool overlap = ov.lo.x < ov.hi.x && ov.lo.y < ov.hi.y && ov.lo.z < ov.hi.z;
            if (overlap && ov.area() * inv_root_area > alpha) {
                SpaSplit ss = spatial_split(refs, box);
                if (ss.dim >= 0 && ss.cost < os.cost) {
                    // distribute; straddling references are split unless moving them whole to one side is cheaper (unsplitting)
                    Box lb = Box::empty(), rb = Box::empty();
                    std::vector<Ref> straddle;
                    for (const Ref& r : refs) {
                        if (axis(r.box.hi, ss.dim) <= ss.plane) { left.push_back(r); lb.grow(r.box); }
                        else if (axis(r.box.lo, ss.dim) >= ss.plane) { right.push_back(r); rb.grow(r.box); }
                        else straddle.push_back(r);
                    }
                    int nl = int(left.size() + straddle.size()), nr = int(right.size() + straddle.size());
                    for (const Ref& r : straddle) {
                        Box cl = clip_to_slab(r.tri, ss.dim, axis(box.lo, ss.dim), ss.plane, r.box);
                        Box cr = clip_to_slab(r.tri, ss.dim, ss.plane, axis(box.hi, ss.dim), r.box);
                        if (cl.is_empty() && cr.is_empty()) { cl = r.box; }      // numerically degenerate: keep it whole on the left
                        if (cl.is_empty()) { right.push_back({ cr, r.tri }); rb.grow(cr); nl--; continue; }
                        if (cr.is_empty()) { left.push_back({ cl, r.tri }); lb.grow(cl); nr--; continue; }
                        Box l_split = lb; l_split.grow(cl); Box r_split = rb; r_split.grow(cr);
                        Box l_whole = lb; l_whole.grow(r.box); Box r_whole = rb; r_whole.grow(r.box);
                        float c_split = l_split.area() * float(nl) + r_split.area() * float(nr);
                        float c_left  = l_whole.area() * float(nl) + rb.area() * float(nr - 1);
                        float c_right = lb.area() * float(nl - 1) + r_whole.area() * float(nr);
                        if (rb.is_empty()) c_left = std::numeric_limits<float>::infinity();
                        if (lb.is_empty()) c_right = std::numeric_limits<float>::infinity();
                        if (c_split <= c_left && c_split <= c_right && budget > 0) {
                            left.push_back({ cl, r.tri }); right.push_back({ cr, r.tri }); lb = l_split; rb = r_split; budget--;
                        } else if (c_left <= c_right) { left.push_back(r); lb = l_whole; nr--; }
                        else { right.push_back(r); rb = r_whole; nl--; }
                    }
                    if (!left.empty() && !right.empty() && int(left.size()) < n && int(right.size()) < n) {
                        spatial = true; lbox = lb; rbox = rb; dim = ss.dim;
                    } else { left.clear(); right.clear(); }
                }
            }
        }
        if (!spatial) {
            left.assign(refs.begin(), refs.begin() + os.index);
            right.assign(refs.begin() + os.index, refs.end());
            lbox = os.left; rbox = os.right;
        }
        { std::vector<Ref>().swap(refs); }                       // the parent's references are no longer needed
        int l = int(nodes.size());
        nodes[node_index].left_or_first = l;
        nodes[node_index].set(0, uint32_t(dim));
        nodes.emplace_back(); nodes.emplace_back();
        if (n >= parallel_limit && left.size() > 1 && right.size() > 1) {
            // big node: the two subtrees are built concurrently, each into its own arrays, then spliced in
            SpatialBuilder a = child_builder(), b = child_builder();
            size_t bl = budget * left.size() / size_t(n), br = budget - bl;
            std::thread worker([&]() { a.nodes.emplace_back(); a.nodes.emplace_back(); a.build_node(0, left, lbox, bl); });
            b.nodes.emplace_back(); b.nodes.emplace_back(); b.build_node(0, right, rbox, br);
            worker.join();
            budget = bl + br;
            splice(l, a); splice(l + 1, b);
        } else {
            build_node(l, left, lbox, budget);
            build_node(l + 1, right, rbox, budget);
        }
    }

    int parallel_limit = 16384;
    SpatialBuilder child_builder() const {
        SpatialBuilder c; c.pos = pos; c.alpha = alpha; c.bins = bins; c.inv_root_area = inv_root_area; c.sweep_limit = sweep_limit; c.parallel_limit = parallel_limit;
        return c;
    }
    // sub-builder `c` holds a subtree with its root in slot 0 (slot 1 unused): root -> our slot `at`, the rest appended
    void splice(int at, const SpatialBuilder& c) {
        const int node_base = int(nodes.size()) - 2, index_base = int(indices.size());
        auto fix = [&](Node2 nd) { if (nd.leaf()) nd.left_or_first += index_base; else nd.left_or_first += node_base; return nd; };
        nodes[at] = fix(c.nodes[0]);
        for (size_t k = 2; k < c.nodes.size(); k++) nodes.push_back(fix(c.nodes[k]));
        indices.insert(indices.end(), c.indices.begin(), c.indices.end());
    }

    void build(int n) {
        std::vector<Ref> refs; refs.resize(size_t(n));
        Box root = Box::empty();
        for (int i = 0; i < n; i++) {
            const float* t = pos + size_t(i) * 9;
            Box b = Box::empty(); b.grow(V3{ t[0], t[1], t[2] }); b.grow(V3{ t[3], t[4], t[5] }); b.grow(V3{ t[6], t[7], t[8] });
            refs[i] = { b, i }; root.grow(b);
        }
        inv_root_area = 1.0f / root.area();
        nodes.clear(); indices.clear();
        nodes.reserve(size_t(4) * n + 2); indices.reserve(size_t(2) * n);
        nodes.emplace_back(); nodes.emplace_back();
        size_t budget = max_refs > size_t(n) ? max_refs - size_t(n) : 0;
        build_node(0, refs, root, budget);
        refit(0);
    }

    // leaf boxes were widened where flat (Box::fatten, like the plain builder's primitive boxes): make every inner box the union of
    // its children again, so that a child never sticks out of the box its quantisation grid is laid over
    Box refit(int ni) {
        if (nodes[ni].leaf()) return nodes[ni].box;
        int l = nodes[ni].left_or_first;
        Box b = refit(l); b.grow(refit(l + 1));
        nodes[ni].box = b;
        return b;
    }
};

// ---------------------------------------------------------------- insertion-based optimisation of a binary BVH
// Bittner, Hapala, Havran 2013 ("Fast insertion-based optimization of bounding volume hierarchies"; the idea behind the reference's
// Src/BVH/BVHOptimizer.cpp, own formulation): take a subtree out of the tree (its sibling moves up into the parent's slot), look for the
// position where putting it back costs the least surface area -- branch and bound over the tree, cost of a position = area of the new
// common parent + the growth of all its ancestors -- and link it in there with the freed pair of slots.  The original position is among
// the candidates, so the tree's SAH cost never rises.  Works on the one-reference-per-leaf BVH2 of both builders (children in adjacent
// slots (l, l + 1), l even, root 0, slot 1 unused); node boxes stay the exact union of their children.
struct ReinsertionOptimizer {
    std::vector<Node2>& nodes;
    std::vector<int> parent;
    explicit ReinsertionOptimizer(std::vector<Node2>& n) : nodes(n) {}

    void link_children(int i) { if (!nodes[i].leaf()) { parent[nodes[i].left_or_first] = i; parent[nodes[i].left_or_first + 1] = i; } }
    void refit_up(int i) {
        while (i >= 0) {
            Node2& n = nodes[i];
            Box b = nodes[n.left_or_first].box; b.grow(nodes[n.left_or_first + 1].box);
            n.box = b;
            i = parent[i];
        }
    }
    struct Item { float induced; int node; bool operator<(const Item& o) const { return induced > o.induced; } };   // min-heap on the induced cost

    // best position for a detached subtree with box `nb`
    int find_position(const Box& nb, std::vector<Item>& heap) {
        const float na = nb.area();
        float best = std::numeric_limits<float>::infinity(); int best_node = -1;
        heap.clear(); heap.push_back({ 0.0f, 0 });
        while (!heap.empty()) {
            std::pop_heap(heap.begin(), heap.end()); Item it = heap.back(); heap.pop_back();
            if (it.induced + na >= best) break;                       // nothing cheaper left
            const Node2& x = nodes[it.node];
            Box u = x.box; u.grow(nb);
            float direct = u.area(), total = it.induced + direct;
            if (total < best) { best = total; best_node = it.node; }
            float child_induced = total - x.box.area();
            if (!x.leaf() && child_induced + na < best) {
                heap.push_back({ child_induced, x.left_or_first }); std::push_heap(heap.begin(), heap.end());
                heap.push_back({ child_induced, x.left_or_first + 1 }); std::push_heap(heap.begin(), heap.end());
            }
        }
        return best_node;
    }

    // one reinsertion of the subtree in slot n; returns false when it cannot be moved (root, child of the root)
    bool reinsert(int n, std::vector<Item>& heap) {
        if (n < 2) return false;
        const int p = parent[n];
        if (p <= 0) return false;
        const int s = n ^ 1, g = parent[p];
        // detach: the sibling moves up into the parent's slot
        nodes[p] = nodes[s]; link_children(p);
        refit_up(g);
        const Box nb = nodes[n].box;
        int x = find_position(nb, heap);
        if (x < 0) x = p;
        // the freed pair (n, s) becomes the children of a new node in x's slot; x's old content moves to s
        nodes[s] = nodes[x]; link_children(s);
        Node2 np; np.box = nodes[s].box; np.box.grow(nb); np.left_or_first = n & ~1; np.set(0, 0);
        nodes[x] = np;
        parent[n] = x; parent[s] = x;
        refit_up(parent[x]);
        return true;
    }

    double sah() const {
        double sum = 0.0;
        for (size_t i = 0; i < nodes.size(); i++) if (i != 1) sum += double(nodes[i].box.area());
        return sum / double(nodes[0].box.area());
    }

    // `passes` sweeps; every sweep offers the `fraction` largest nodes (by surface area) for reinsertion, largest first
    void run(int passes, float fraction) {
        const int total = int(nodes.size());
        if (total < 8) return;
        parent.assign(size_t(total), -1);
        for (int i = 0; i < total; i++) if (i != 1) link_children(i);
        std::vector<Item> heap; heap.reserve(1024);
        std::vector<std::pair<float, int>> order; order.reserve(size_t(total));
        for (int pass = 0; pass < passes; pass++) {
            order.clear();
            for (int i = 2; i < total; i++) order.push_back({ nodes[i].box.area(), i });
            size_t take = size_t(double(order.size()) * double(fraction));
            if (take < 1) take = 1; if (take > order.size()) take = order.size();
            std::partial_sort(order.begin(), order.begin() + long(take), order.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first > b.first; });
            for (size_t k = 0; k < take; k++) reinsert(order[k].second, heap);
        }
    }
};

struct Built {
    BVH2 bvh2;
    BVH4 bvh4;
    BVH8 bvh8;
    int  kind; // 2, 4 or 8
};

void prims_from_triangles(const float* pos, int n, Prims& p) {
    p.box.resize(n); p.center.resize(n);
    for (int i = 0; i < n; i++) {
        const float* t = pos + size_t(i) * 9;
        V3 a = { t[0], t[1], t[2] }, b = { t[3], t[4], t[5] }, c = { t[6], t[7], t[8] };
        Box bx = Box::empty(); bx.grow(a); bx.grow(b); bx.grow(c); bx.fatten();
        p.box[i] = bx;
        p.center[i] = { (a.x + b.x + c.x) / 3.0f, (a.y + b.y + c.y) / 3.0f, (a.z + b.z + c.z) / 3.0f };
    }
}

Built* finish(Prims& prims, int kind, float sah_node, float sah_leaf) {
    Built* b = new Built(); b->kind = kind;
    BVH2 raw;
    SAHBuilder(prims, raw.nodes).build(raw.indices);
    if (kind == 8) {
        WideConverter(raw, b->bvh8).run();
    } else if (sah_leaf > 0.0f) {
        Collapser c{ raw, b->bvh2, sah_node, sah_leaf, {} };
        c.run();
    } else {
        b->bvh2 = std::move(raw);
    }
    if (kind == 4) QuadConverter{ b->bvh2, b->bvh4 }.run();          // 4-wide tree over the (leaf-collapsed) binary one
    return b;
}

} // namespace

extern "C" {

// Build over triangles: pos = n * 9 floats (v0,v1,v2). kind = 8 (CWBVH) or 2 (binary, SAH-collapsed leaves
// when sah_leaf > 0; raw one-primitive leaves when sah_leaf <= 0).
void* ptbh_build_triangles(const float* pos, int n, int kind, float sah_node, float sah_leaf) {
    if (n <= 0 || (kind != 2 && kind != 4 && kind != 8)) return nullptr;
    Prims p; prims_from_triangles(pos, n, p);
    return finish(p, kind, sah_node, sah_leaf);
}

// Build over already-boxed primitives (TLAS over instance boxes): aabb = n * 6 floats (min,max); centers = box centers.
// TLAS BVH2 is never leaf-collapsed (reference: BVH2Converter is a plain copy, BVHConverter.h:17-26).
void* ptbh_build_boxes(const float* aabb, int n, int kind) {
    if (n <= 0 || (kind != 2 && kind != 4 && kind != 8)) return nullptr;
    Prims p; p.box.resize(n); p.center.resize(n);
    for (int i = 0; i < n; i++) {
        const float* a = aabb + size_t(i) * 6;
        p.box[i] = { { a[0], a[1], a[2] }, { a[3], a[4], a[5] } };
        p.center[i] = p.box[i].center();
    }
    return finish(p, kind, 0.0f, 0.0f);
}


// Conversion of an already built binary BVH (the reference's `.bvh` cache, BVHLoader.cpp: raw one-primitive-per-leaf BVH2 + indices) to
// the kind the kernels walk, without rebuilding: kind 8 -> CWBVH, kind 2 -> SAH leaf collapse (sah_leaf > 0) or a plain copy.
// nodes: n_nodes x 32 bytes (BVH.h:11-23 layout, node 1 = dummy); returns null on a malformed tree.
void* ptbh_from_bvh2(const void* nodes, int n_nodes, const int* indices, int n_indices, int kind, float sah_node, float sah_leaf) {
    if (!nodes || !indices || n_nodes < 1 || n_indices < 1 || (kind != 2 && kind != 4 && kind != 8)) return nullptr;
    BVH2 raw;
    raw.nodes.resize(size_t(n_nodes)); std::memcpy(raw.nodes.data(), nodes, size_t(n_nodes) * sizeof(Node2));
    raw.indices.assign(indices, indices + n_indices);
    {   // every link in range and the graph a tree, so that a damaged file cannot walk out of the arrays or loop (an optimised BVH may
        // have children stored before their parents: only reachability is checked, not index order)
        std::vector<int> todo{ 0 };
        size_t visited = 0;
        while (!todo.empty()) {
            int i = todo.back(); todo.pop_back();
            if (++visited > size_t(n_nodes)) return nullptr;
            const Node2& nd = raw.nodes[size_t(i)];
            if (nd.leaf()) { if (nd.left_or_first < 0 || size_t(nd.left_or_first) + nd.count() > size_t(n_indices)) return nullptr; }
            else {
                if (nd.left_or_first < 1 || nd.left_or_first + 1 >= n_nodes) return nullptr;
                todo.push_back(nd.left_or_first); todo.push_back(nd.left_or_first + 1);
            }
        }
    }
    Built* b = new Built(); b->kind = kind;
    if (kind == 8) WideConverter(raw, b->bvh8).run();
    else if (sah_leaf > 0.0f) { Collapser c{ raw, b->bvh2, sah_node, sah_leaf, {} }; c.run(); }
    else b->bvh2 = std::move(raw);
    if (kind == 4) QuadConverter{ b->bvh2, b->bvh4 }.run();
    return b;
}

// Split-BVH build (spatial splits) -> CWBVH.  alpha: minimum overlap of the object split's children, relative to the root's area,
// for a spatial split to be tried (the reference's --sbvh-alpha, Config.h:58); max_dup: cap on references as a multiple of n
// (e.g. 1.5).  ptbh_index_count() is the number of REFERENCES (>= n): indices may repeat a triangle.
static thread_local float g_tri_cost = 1.0f;     // per calling thread: two contexts may build their merged trees concurrently
void ptbh_set_tri_cost(float c) { g_tri_cost = c; }     // tuning knob for tools/cpu_bvh_quality.py
static thread_local int g_opt_passes = 0; static thread_local float g_opt_fraction = 1.0f; static thread_local int g_opt_max_depth = 0;
static thread_local double g_opt_sah[2] = { 0.0, 0.0 };
// insertion-based optimisation of the binary tree before the wide collapse (ReinsertionOptimizer): `passes` sweeps over the `fraction`
// largest nodes; max_depth > 0 keeps the unoptimised tree when the optimised CWBVH would be deeper than that (traversal stack)
void ptbh_set_optimizer(int passes, float fraction, int max_depth) { g_opt_passes = passes; g_opt_fraction = fraction; g_opt_max_depth = max_depth; }
void ptbh_optimizer_sah(double* before_after) { before_after[0] = g_opt_sah[0]; before_after[1] = g_opt_sah[1]; }
void* ptbh_build_triangles_sbvh(const float* pos, int n, float alpha, int bins, float max_dup) {
    if (n <= 0) return nullptr;
    SpatialBuilder sb; sb.pos = pos; sb.alpha = alpha; sb.bins = bins < 8 ? 8 : bins;
    sb.max_refs = size_t(double(n) * double(max_dup < 1.0f ? 1.0f : max_dup));
    sb.build(n);
    Built* b = new Built(); b->kind = 8;
    BVH2 raw; raw.nodes.swap(sb.nodes); raw.indices.swap(sb.indices);
    if (g_opt_passes > 0) {
        BVH2 opt = raw;
        ReinsertionOptimizer ro(opt.nodes);
        g_opt_sah[0] = ro.sah();
        ro.run(g_opt_passes, g_opt_fraction);
        g_opt_sah[1] = ro.sah();
        WideConverter(opt, b->bvh8, g_tri_cost).run();
        if (g_opt_max_depth <= 0 || ptb_merge::max_depth(reinterpret_cast<const unsigned char*>(b->bvh8.nodes.data()), 0) <= g_opt_max_depth) return b;
        b->bvh8 = BVH8();                                   // too deep for the caller's stack: fall back to the tree as built
    }
    WideConverter(raw, b->bvh8, g_tri_cost).run();
    return b;
}

// CPU closest-hit walk of a built CWBVH (tuning / test aid: visit counts per ray predict the device kernel's work; BVH8.h:113-274
// semantics: octant-ordered child pops, triangles of a node tested before descending, culling against the current closest hit).
// rays: n x 6 floats (origin, direction).  Adds the visits to counts[0] (nodes) and counts[1] (triangle tests); hit_t / hit_tri may be null.
// optional sink for ptbh_trace_stats: (ray index, node index, closest hit so far) of every node visit, up to `capacity` records
static int* g_visit_ray = nullptr; static int* g_visit_node = nullptr; static float* g_visit_t = nullptr; static long long g_visit_cap = 0, g_visit_n = 0;
void ptbh_set_visit_sink(int* ray, int* node, float* t, long long capacity) { g_visit_ray = ray; g_visit_node = node; g_visit_t = t; g_visit_cap = capacity; g_visit_n = 0; }
long long ptbh_visit_count() { return g_visit_n; }
void ptbh_trace_stats(void* h, const float* pos, const float* rays, int n_rays, unsigned long long* counts, float* hit_t, int* hit_tri) {
    Built* b = static_cast<Built*>(h);
    const std::vector<Node8>& nodes = b->bvh8.nodes; const std::vector<int>& idx = b->bvh8.indices;
    unsigned long long n_nodes = 0, n_tris = 0;
    struct Entry { uint32_t base, mask; };
    for (int r = 0; r < n_rays; r++) {
        const float* ray = rays + size_t(r) * 6;
        float ox = ray[0], oy = ray[1], oz = ray[2], dx = ray[3], dy = ray[4], dz = ray[5];
        unsigned oct = (dx < 0.0f ? 0u : 4u) | (dy < 0.0f ? 0u : 2u) | (dz < 0.0f ? 0u : 1u);   // "inverse octant", BVH8.h:5-10
        float best = std::numeric_limits<float>::infinity(); int best_tri = -1;
        Entry stack[64]; int sp = 0;
        Entry cur = { 0u, 0x80000000u };      // root: pseudo group with one child bit
        for (;;) {
            uint32_t tri_base = 0, tri_mask = 0;
            if (cur.mask & 0xff000000u) {
                int bit = 31 - __builtin_clz(cur.mask);
                uint32_t imask_hits = cur.mask;
                cur.mask &= ~(1u << bit);
                if (cur.mask & 0xff000000u) stack[sp++] = cur;
                unsigned slot = unsigned(bit - 24) ^ oct;
                unsigned rel = unsigned(__builtin_popcount(imask_hits & ~(0xffffffffu << slot)));
                const Node8& nd = nodes[cur.base + rel];
                n_nodes++;
                if (g_visit_ray && g_visit_n < g_visit_cap) { g_visit_ray[g_visit_n] = r; g_visit_node[g_visit_n] = int(cur.base + rel); g_visit_t[g_visit_n] = best; g_visit_n++; }
                float ex, ey, ez; uint32_t u;
                u = uint32_t(nd.e[0]) << 23; std::memcpy(&ex, &u, 4); u = uint32_t(nd.e[1]) << 23; std::memcpy(&ey, &u, 4); u = uint32_t(nd.e[2]) << 23; std::memcpy(&ez, &u, 4);
                float aix = ex / dx, aiy = ey / dy, aiz = ez / dz;
                float aox = (nd.p.x - ox) / dx, aoy = (nd.p.y - oy) / dy, aoz = (nd.p.z - oz) / dz;
                uint32_t hm = 0;
                for (int j = 0; j < 8; j++) {
                    uint8_t meta = nd.meta[j];
                    if (!meta) continue;
                    float lx = float(dx < 0.0f ? nd.qhi_x[j] : nd.qlo_x[j]), hx = float(dx < 0.0f ? nd.qlo_x[j] : nd.qhi_x[j]);
                    float ly = float(dy < 0.0f ? nd.qhi_y[j] : nd.qlo_y[j]), hy = float(dy < 0.0f ? nd.qlo_y[j] : nd.qhi_y[j]);
                    float lz = float(dz < 0.0f ? nd.qhi_z[j] : nd.qlo_z[j]), hz = float(dz < 0.0f ? nd.qlo_z[j] : nd.qhi_z[j]);
                    float tmin = std::fmax(std::fmax(lx * aix + aox, ly * aiy + aoy), std::fmax(lz * aiz + aoz, 0.0f));
                    float tmax = std::fmin(std::fmin(hx * aix + aox, hy * aiy + aoy), std::fmin(hz * aiz + aoz, best));
                    if (tmin < tmax) {
                        bool inner = (meta & 0x18) == 0x18 && (meta & 0x20);     // 001xxxxx with index >= 24
                        if (nd.imask & (1u << j)) hm |= 1u << (24 + ((unsigned(meta) & 7u) ^ oct));
                        else hm |= (uint32_t(meta) >> 5) << (meta & 31u);
                        (void)inner;
                    }
                }
                cur.base = nd.base_child; cur.mask = (hm & 0xff000000u) | nd.imask;
                tri_base = nd.base_triangle; tri_mask = hm & 0x00ffffffu;
            } else {
                tri_base = cur.base; tri_mask = cur.mask; cur = { 0u, 0u };
            }
            while (tri_mask) {
                int tb = 31 - __builtin_clz(tri_mask); tri_mask &= ~(1u << tb);
                int ti = idx[tri_base + unsigned(tb)];
                n_tris++;
                const float* t = pos + size_t(ti) * 9;
                float e1x = t[3] - t[0], e1y = t[4] - t[1], e1z = t[5] - t[2], e2x = t[6] - t[0], e2y = t[7] - t[1], e2z = t[8] - t[2];
                float hx = dy * e2z - dz * e2y, hy = dz * e2x - dx * e2z, hz = dx * e2y - dy * e2x;
                float a = e1x * hx + e1y * hy + e1z * hz, f = 1.0f / a;
                float sx = ox - t[0], sy = oy - t[1], sz = oz - t[2];
                float uu = f * (sx * hx + sy * hy + sz * hz);
                if (uu >= 0.0f && uu <= 1.0f) {
                    float qx = sy * e1z - sz * e1y, qy = sz * e1x - sx * e1z, qz = sx * e1y - sy * e1x;
                    float vv = f * (dx * qx + dy * qy + dz * qz);
                    if (vv >= 0.0f && uu + vv <= 1.0f) {
                        float tt = f * (e2x * qx + e2y * qy + e2z * qz);
                        if (tt > 0.0f && tt < best) { best = tt; best_tri = ti; }
                    }
                }
            }
            if ((cur.mask & 0xff000000u) == 0) {
                if (sp == 0) break;
                cur = stack[--sp];
            }
        }
        if (hit_t) hit_t[r] = best;
        if (hit_tri) hit_tri[r] = best_tri;
    }
    counts[0] += n_nodes; counts[1] += n_tris;
}

int ptbh_kind(void* h)        { return static_cast<Built*>(h)->kind; }
int ptbh_node_count(void* h)  { Built* b = static_cast<Built*>(h); return int(b->kind == 8 ? b->bvh8.nodes.size() : b->kind == 4 ? b->bvh4.nodes.size() : b->bvh2.nodes.size()); }
int ptbh_index_count(void* h) { Built* b = static_cast<Built*>(h); return int(b->kind == 8 ? b->bvh8.indices.size() : b->bvh2.indices.size()); }
int ptbh_node_bytes(void* h)  { int k = static_cast<Built*>(h)->kind; return k == 8 ? 80 : k == 4 ? 128 : 32; }

// Copies nodes (80 or 32 bytes each) and primitive order out; adds the given offsets so the caller can
// concatenate many BLAS into one node/triangle array (reference: Integrator.cpp:252-277 / 184-207).
void ptbh_export(void* h, void* nodes_out, int* indices_out, int node_offset, int index_offset) {
    Built* b = static_cast<Built*>(h);
    if (b->kind == 8) {
        Node8* dst = static_cast<Node8*>(nodes_out);
        for (size_t i = 0; i < b->bvh8.nodes.size(); i++) {
            dst[i] = b->bvh8.nodes[i];
            dst[i].base_child    += uint32_t(node_offset);
            dst[i].base_triangle += uint32_t(index_offset);
        }
        std::memcpy(indices_out, b->bvh8.indices.data(), b->bvh8.indices.size() * sizeof(int));
    } else if (b->kind == 4) {
        // Integrator.cpp:216-246: leaf slots move by the primitive offset, internal slots (and the entry slot of node 1) by the node offset
        Node4* dst = static_cast<Node4*>(nodes_out);
        for (size_t i = 0; i < b->bvh4.nodes.size(); i++) {
            dst[i] = b->bvh4.nodes[i];
            for (int c = 0; c < 4; c++) {
                if (dst[i].child[c].count == -1) break;
                dst[i].child[c].index += dst[i].child[c].count > 0 ? index_offset : node_offset;
            }
        }
        std::memcpy(indices_out, b->bvh4.indices.data(), b->bvh4.indices.size() * sizeof(int));
    } else {
        Node2* dst = static_cast<Node2*>(nodes_out);
        for (size_t i = 0; i < b->bvh2.nodes.size(); i++) {
            dst[i] = b->bvh2.nodes[i];
            if (i == 1) continue; // dummy
            if (dst[i].leaf()) dst[i].left_or_first += index_offset; else dst[i].left_or_first += node_offset;
        }
        std::memcpy(indices_out, b->bvh2.indices.data(), b->bvh2.indices.size() * sizeof(int));
    }
}

void ptbh_free(void* h) { delete static_cast<Built*>(h); }

// C wrappers of host/static_merge.h for the CPU test suite (the product calls the inline functions directly)
int ptbh_collect_leaf_primitives(const void* nodes, unsigned root, int* out, int capacity) {
    std::vector<int> v;
    ptb_merge::collect_leaf_primitives(static_cast<const unsigned char*>(nodes), root, v);
    for (size_t i = 0; i < v.size() && (int)i < capacity; i++) out[i] = v[i];
    return (int)v.size();
}
int ptbh_prune_tlas(void* nodes, const char* merged, int instance_count) {
    std::vector<char> m(merged, merged + instance_count);
    return ptb_merge::prune_tlas(static_cast<unsigned char*>(nodes), 0, m) ? 1 : 0;
}
int ptbh_max_depth(const void* nodes, unsigned root) { return ptb_merge::max_depth(static_cast<const unsigned char*>(nodes), root); }
void ptbh_bfs_relayout(const void* dfs, int node_count, int base, void* bfs) {
    ptb_merge::bfs_relayout(static_cast<const unsigned char*>(dfs), node_count, base, static_cast<unsigned char*>(bfs));
}

} // extern "C"
