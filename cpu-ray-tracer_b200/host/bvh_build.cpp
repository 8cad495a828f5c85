// bvh_build.cpp — host-side builders that produce exactly what the reference's builders produce, for
// scenes that do not come from the reference's own loaders (synthetic meshes, instanced grids).
//
// SURVEY.md section 8f rank 1/2 ("next" rows either side of the hot path): the step immediately before
// the path.  Restated from the published algorithm (J. Bikker, "How to build a BVH", binned SAH) as the
// reference implements it; every routine cites the reference lines it must agree with, and
// tests/test_host_build.py checks that, fed the reference's triangles, it reproduces the reference's
// node array and triangle order bit for bit.  Compile with -ffp-contract=off (build.py does).
//
//   rtb_build_bvh     BVH::Build / BLASBVH::Build          (bvh.cpp:4-24,45-178; blas_bvh.cpp:82-257)
//   rtb_world_bounds  BLASBVH::SetTransform bounds          (blas_bvh.cpp:363-374)
//   rtb_invert_rigid  mat4::FastInvertedTransformNoScale    (tmplmath.h:745-768)
//   rtb_build_tlas    TLASBVH::Build / FindBestMatch        (tlas_bvh.cpp:17-70), without the 256-instance
//                     cap of `int nodeIdx[256]`; child indices stay 2 x 16 bit
//   rtb_kd_*          KDTree::Build / Subdivide             (kdtree.cpp:4-112), flattened into rt_kd_node[]
//   rtb_grid_*        Grid::Build                           (grid.cpp:4-60), flattened into rt_grid_desc arrays
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/rt_b200.h"

namespace {

constexpr int BINS = 8; // BVH_BINS, bvh.h:7

inline float minf(float a, float b) { return a < b ? a : b; } // tmplmath.h:122 fminf / min
inline float maxf(float a, float b) { return a > b ? a : b; }

// aabb as tmplmath.h:577-632 uses it: starts at (1e34, -1e34); Area = max(0, ex*ey + ex*ez + ey*ez)
struct Box {
    float mn[3] = { 1e34f, 1e34f, 1e34f }, mx[3] = { -1e34f, -1e34f, -1e34f };
    void grow(const float* p) { for (int a = 0; a < 3; a++) mn[a] = p[a] < mn[a] ? p[a] : mn[a], mx[a] = p[a] > mx[a] ? p[a] : mx[a]; } // _mm_min_ps(bmin, p)
    void grow(const Box& b) { for (int a = 0; a < 3; a++) mn[a] = b.mn[a] < mn[a] ? b.mn[a] : mn[a], mx[a] = b.mx[a] > mx[a] ? b.mx[a] : mx[a]; }
    float area() const
    {
        const float e0 = mx[0] - mn[0], e1 = mx[1] - mn[1], e2 = mx[2] - mn[2];
        return maxf(0.0f, e0 * e1 + e0 * e2 + e1 * e2);
    }
};

struct BvhBuilder {
    const rt_tri* tris;
    rt_bvh_node* nodes;
    uint32_t* idx;
    uint32_t nodesUsed = 1, maxDepth = 0; // bvh.h:41

    // UpdateNodeBounds bvh.cpp:45-61: min/max over the three vertices, starting from +-1e30
    void bounds(uint32_t n)
    {
        rt_bvh_node& node = nodes[n];
        for (int a = 0; a < 3; a++) node.aabb_min[a] = 1e30f, node.aabb_max[a] = -1e30f;
        for (uint32_t first = node.left_first, i = 0; i < node.tri_count; i++)
        {
            const rt_tri& t = tris[idx[first + i]];
            for (int a = 0; a < 3; a++)
            {
                node.aabb_min[a] = minf(minf(minf(node.aabb_min[a], t.v0[a]), t.v1[a]), t.v2[a]);
                node.aabb_max[a] = maxf(maxf(maxf(node.aabb_max[a], t.v0[a]), t.v1[a]), t.v2[a]);
            }
        }
    }

    // FindBestSplitPlane bvh.cpp:124-178: 8 centroid bins per axis, 7 candidate planes, cost N_l*A_l + N_r*A_r
    float best_split(const rt_bvh_node& node, int& axis, float& splitPos) const
    {
        float bestCost = 1e30f;
        for (int a = 0; a < 3; a++)
        {
            float bmin = 1e30f, bmax = -1e30f;
            for (uint32_t i = 0; i < node.tri_count; i++)
            {
                const float c = tris[idx[node.left_first + i]].centroid[a];
                bmin = minf(bmin, c), bmax = maxf(bmax, c);
            }
            if (bmin == bmax) continue;
            Box binBox[BINS];
            int binCount[BINS] = {};
            float scale = BINS / (bmax - bmin);
            for (uint32_t i = 0; i < node.tri_count; i++)
            {
                const rt_tri& t = tris[idx[node.left_first + i]];
                int b = (int)((t.centroid[a] - bmin) * scale);
                if (b > BINS - 1) b = BINS - 1;
                binCount[b]++;
                binBox[b].grow(t.v0), binBox[b].grow(t.v1), binBox[b].grow(t.v2);
            }
            float leftArea[BINS - 1], rightArea[BINS - 1];
            int leftCount[BINS - 1], rightCount[BINS - 1];
            Box leftBox, rightBox;
            int leftSum = 0, rightSum = 0;
            for (int i = 0; i < BINS - 1; i++)
            {
                leftSum += binCount[i], leftCount[i] = leftSum;
                leftBox.grow(binBox[i]), leftArea[i] = leftBox.area();
                rightSum += binCount[BINS - 1 - i], rightCount[BINS - 2 - i] = rightSum;
                rightBox.grow(binBox[BINS - 1 - i]), rightArea[BINS - 2 - i] = rightBox.area();
            }
            scale = (bmax - bmin) / BINS;
            for (int i = 0; i < BINS - 1; i++)
            {
                // an empty side has area +inf (the 1e34 sentinel box overflows) and count 0: 0 * inf = NaN,
                // which never compares less than bestCost — the reference skips such planes the same way
                const float cost = leftCount[i] * leftArea[i] + rightCount[i] * rightArea[i];
                if (cost < bestCost) axis = a, splitPos = bmin + scale * (i + 1), bestCost = cost;
            }
        }
        return bestCost;
    }

    // Subdivide bvh.cpp:63-115, with an explicit stack instead of recursion (same visiting order:
    // the left subtree is finished before the right one starts, so node numbering is identical)
    void build(uint32_t n)
    {
        for (uint32_t i = 0; i < n; i++) idx[i] = i;
        nodes[0].left_first = 0, nodes[0].tri_count = n;
        bounds(0);
        struct Item { uint32_t node, depth; };
        std::vector<Item> todo;
        todo.push_back({ 0u, 0u });
        while (!todo.empty())
        {
            const Item it = todo.back();
            todo.pop_back();
            rt_bvh_node& node = nodes[it.node];
            if (node.tri_count <= 2) continue;
            int axis = 0;
            float splitPos = 0;
            const float splitCost = best_split(node, axis, splitPos);
            const float ex = node.aabb_max[0] - node.aabb_min[0], ey = node.aabb_max[1] - node.aabb_min[1], ez = node.aabb_max[2] - node.aabb_min[2];
            const float noSplit = node.tri_count * (ex * ey + ey * ez + ez * ex); // CalculateNodeCost bvh.cpp:117-122
            if (splitCost >= noSplit) continue;
            int i = (int)node.left_first, j = i + (int)node.tri_count - 1;
            while (i <= j)
            {
                if (tris[idx[i]].centroid[axis] < splitPos) i++;
                else { const uint32_t t = idx[i]; idx[i] = idx[j]; idx[j] = t; j--; }
            }
            const int leftCount = i - (int)node.left_first;
            if (leftCount == 0 || leftCount == (int)node.tri_count) continue;
            const uint32_t l = nodesUsed++, r = nodesUsed++;
            nodes[l].left_first = node.left_first, nodes[l].tri_count = (uint32_t)leftCount;
            nodes[r].left_first = (uint32_t)i, nodes[r].tri_count = node.tri_count - (uint32_t)leftCount;
            node.left_first = l, node.tri_count = 0;
            bounds(l), bounds(r);
            if (it.depth > maxDepth) maxDepth = it.depth;
            todo.push_back({ r, it.depth + 1 }); // popped after the whole left subtree
            todo.push_back({ l, it.depth + 1 });
        }
    }
};

// KDTree::Build (kdtree.cpp:4-43) + Subdivide (:45-112): spatial median on the longest axis, depth <= 20, leaves
// of <= 2 triangles, a triangle goes left when its box ends before the plane, right when it starts after
// `splitPos - 0.001` (double arithmetic, :70), else into both.  Nodes are numbered the way the flattener of the
// reference build numbers them (children allocated when the parent is visited, depth first, left first) and leaf
// lists are concatenated in that visiting order.
struct KdBuilder {
    const rt_tri* tris;
    std::vector<Box> triBounds;
    std::vector<rt_kd_node> nodes;
    std::vector<uint32_t> leafIdx;
    uint32_t maxDepth = 0;

    void build(uint32_t n)
    {
        triBounds.resize(n);
        Box all;
        std::vector<uint32_t> idx(n);
        for (uint32_t i = 0; i < n; i++)
        {
            Box b;
            b.grow(tris[i].v0), b.grow(tris[i].v1), b.grow(tris[i].v2);
            all.grow(b);
            triBounds[i] = b, idx[i] = i;
        }
        nodes.emplace_back();
        memset(&nodes[0], 0, sizeof(rt_kd_node));
        memcpy(nodes[0].aabb_min, all.mn, 12), memcpy(nodes[0].aabb_max, all.mx, 12);
        subdivide(0, idx, 0);
    }

    void leaf(uint32_t n, const std::vector<uint32_t>& idx)
    {
        nodes[n].left = nodes[n].right = -1;
        nodes[n].tri_start = (uint32_t)leafIdx.size(), nodes[n].tri_count = (uint32_t)idx.size();
        leafIdx.insert(leafIdx.end(), idx.begin(), idx.end());
    }

    void subdivide(uint32_t n, std::vector<uint32_t>& idx, int depth)
    {
        if (depth >= 20 || idx.size() <= 2) { leaf(n, idx); return; } // m_maxBuildDepth kdtree.h:31, :48-50
        if ((uint32_t)depth > maxDepth) maxDepth = depth;
        float mn[3], mx[3], extent[3];
        memcpy(mn, nodes[n].aabb_min, 12), memcpy(mx, nodes[n].aabb_max, 12);
        for (int a = 0; a < 3; a++) extent[a] = mx[a] - mn[a];
        int axis = 0;
        if (extent[1] > extent[0]) axis = 1;
        if (extent[2] > extent[axis]) axis = 2;
        const float distance = extent[axis] * 0.5f;
        const float splitPos = mn[axis] + distance;
        std::vector<uint32_t> left, right;
        for (uint32_t t : idx)
        {
            if (triBounds[t].mx[axis] < splitPos) left.push_back(t);
            else if (triBounds[t].mn[axis] > splitPos - 0.001) right.push_back(t);
            else left.push_back(t), right.push_back(t);
        }
        idx.clear();
        idx.shrink_to_fit();
        const uint32_t l = (uint32_t)nodes.size(), r = l + 1;
        nodes.emplace_back(), nodes.emplace_back();
        memset(&nodes[l], 0, 2 * sizeof(rt_kd_node));
        nodes[n].left = (int32_t)l, nodes[n].right = (int32_t)r;
        nodes[n].split_axis = axis, nodes[n].split_distance = distance;
        memcpy(nodes[l].aabb_min, mn, 12), memcpy(nodes[l].aabb_max, mx, 12), nodes[l].aabb_max[axis] = splitPos;
        memcpy(nodes[r].aabb_min, mn, 12), memcpy(nodes[r].aabb_max, mx, 12), nodes[r].aabb_min[axis] = splitPos;
        subdivide(l, left, depth + 1);
        subdivide(r, right, depth + 1);
    }
};

// max(a, min(f, b)), tmplmath.h:436
inline int clampi(int f, int a, int b) { const int m = f < b ? f : b; return a > m ? a : m; }

// Grid::Build (grid.cpp:4-60)
struct GridBuilder {
    int32_t res[3];
    float cell[3], mn[3], mx[3];
    std::vector<uint32_t> cellStart, idx;

    void build(const rt_tri* tris, uint32_t n)
    {
        Box all;
        std::vector<Box> tb(n);
        for (uint32_t i = 0; i < n; i++)
        {
            tb[i].grow(tris[i].v0), tb[i].grow(tris[i].v1), tb[i].grow(tris[i].v2);
            all.grow(tb[i]);
        }
        memcpy(mn, all.mn, 12), memcpy(mx, all.mx, 12);
        float size[3];
        for (int a = 0; a < 3; a++) size[a] = mx[a] - mn[a];
        const float cubeRoot = powf(5 * (int)n / (size[0] * size[1] * size[2]), 1 / 3.f); // :18
        for (int a = 0; a < 3; a++)
        {
            res[a] = static_cast<int>(floorf(size[a] * cubeRoot));
            res[a] = res[a] < 128 ? res[a] : 128;  // max(1, min(resolution, 128)) :22
            res[a] = 1 > res[a] ? 1 : res[a];
        }
        for (int a = 0; a < 3; a++) cell[a] = size[a] / res[a];
        const size_t cells = (size_t)res[0] * res[1] * res[2];
        std::vector<uint32_t> count(cells + 1, 0);
        auto range = [&](uint32_t i, int lo[3], int hi[3]) {
            for (int a = 0; a < 3; a++)
            {
                lo[a] = clampi(static_cast<int>((tb[i].mn[a] - mn[a]) / cell[a]), 0, res[a] - 1);
                hi[a] = clampi(static_cast<int>((tb[i].mx[a] - mn[a]) / cell[a]), 0, res[a] - 1);
            }
        };
        for (int pass = 0; pass < 2; pass++)
        {
            // pass 0 counts, pass 1 fills: per cell the triangles end up in increasing index order, as push_back leaves them
            for (uint32_t i = 0; i < n; i++)
            {
                int lo[3], hi[3];
                range(i, lo, hi);
                for (int z = lo[2]; z <= hi[2]; ++z)
                    for (int y = lo[1]; y <= hi[1]; ++y)
                        for (int x = lo[0]; x <= hi[0]; ++x)
                        {
                            const size_t c = (size_t)x + (size_t)y * res[0] + (size_t)z * res[0] * res[1];
                            if (pass == 0) count[c + 1]++;
                            else idx[count[c]++] = i;
                        }
            }
            if (pass == 0)
            {
                for (size_t c = 0; c < cells; c++) count[c + 1] += count[c];
                cellStart = count;
                idx.resize(count[cells]);
            }
        }
    }
};

} // namespace

extern "C" {

// KD-tree / grid builders hand their variable-size output over through a handle: build, query the sizes, copy, free.
void* rtb_kd_build(const rt_tri* tris, uint32_t n)
{
    if (!tris || n == 0) return nullptr;
    KdBuilder* b = new KdBuilder();
    b->tris = tris;
    b->build(n);
    return b;
}
void rtb_kd_sizes(void* h, uint32_t* nodes, uint32_t* indices, uint32_t* max_depth)
{
    KdBuilder* b = (KdBuilder*)h;
    *nodes = (uint32_t)b->nodes.size(), *indices = (uint32_t)b->leafIdx.size(), *max_depth = b->maxDepth;
}
void rtb_kd_copy(void* h, rt_kd_node* nodes_out, uint32_t* indices_out)
{
    KdBuilder* b = (KdBuilder*)h;
    memcpy(nodes_out, b->nodes.data(), b->nodes.size() * sizeof(rt_kd_node));
    if (!b->leafIdx.empty()) memcpy(indices_out, b->leafIdx.data(), b->leafIdx.size() * 4);
}
void rtb_kd_free(void* h) { delete (KdBuilder*)h; }

void* rtb_grid_build(const rt_tri* tris, uint32_t n)
{
    if (!tris || n == 0) return nullptr;
    GridBuilder* g = new GridBuilder();
    g->build(tris, n);
    return g;
}
// header12: resolution (3 x int32), cell size, bounds min, bounds max (3 floats each) = the grid_header chunk
void rtb_grid_sizes(void* h, void* header12, uint32_t* cells, uint32_t* indices)
{
    GridBuilder* g = (GridBuilder*)h;
    char* o = (char*)header12;
    memcpy(o, g->res, 12), memcpy(o + 12, g->cell, 12), memcpy(o + 24, g->mn, 12), memcpy(o + 36, g->mx, 12);
    *cells = (uint32_t)g->cellStart.size() - 1, *indices = (uint32_t)g->idx.size();
}
void rtb_grid_copy(void* h, uint32_t* cell_start_out, uint32_t* indices_out)
{
    GridBuilder* g = (GridBuilder*)h;
    memcpy(cell_start_out, g->cellStart.data(), g->cellStart.size() * 4);
    if (!g->idx.empty()) memcpy(indices_out, g->idx.data(), g->idx.size() * 4);
}
void rtb_grid_free(void* h) { delete (GridBuilder*)h; }

// nodes_out: 2n - 1 entries, tri_indices_out: n entries.  Centroids are read from the triangles
// (the reference sets them to (v0 + v1 + v2) * 0.3333f when it loads a model, model.cpp:77).
int rtb_build_bvh(const rt_tri* tris, uint32_t n, rt_bvh_node* nodes_out, uint32_t* tri_indices_out, uint32_t* nodes_used, uint32_t* max_depth)
{
    if (!tris || !nodes_out || !tri_indices_out || n == 0) return RT_ERR_INVALID;
    memset(nodes_out, 0, sizeof(rt_bvh_node) * (2 * (size_t)n - 1));
    BvhBuilder b;
    b.tris = tris, b.nodes = nodes_out, b.idx = tri_indices_out;
    b.build(n);
    if (nodes_used) *nodes_used = b.nodesUsed;
    if (max_depth) *max_depth = b.maxDepth;
    return RT_OK;
}

// float4(a, 1) * M: tmplmath.cpp:155-165 (row . vector, left-to-right sum)
static inline void transform_position(const float* M, const float* a, float* o)
{
    for (int r = 0; r < 3; r++) o[r] = M[4 * r] * a[0] + M[4 * r + 1] * a[1] + M[4 * r + 2] * a[2] + M[4 * r + 3] * 1.0f;
}

// world bounds of a BLAS root box under T: the eight corners, blas_bvh.cpp:369-373. out = min.xyz, max.xyz
void rtb_world_bounds(const float* root_min, const float* root_max, const float* T, float* out6)
{
    Box w;
    for (int i = 0; i < 8; i++)
    {
        const float c[3] = { i & 1 ? root_max[0] : root_min[0], i & 2 ? root_max[1] : root_min[1], i & 4 ? root_max[2] : root_min[2] };
        float p[3];
        transform_position(T, c, p);
        w.grow(p);
    }
    memcpy(out6, w.mn, 12), memcpy(out6 + 3, w.mx, 12);
}

// tmplmath.h:745-768: transpose of the 3x3 part, translation = -(t . row)
void rtb_invert_rigid(const float* M, float* r)
{
    static const float I[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    memcpy(r, I, 64);
    r[0] = M[0], r[1] = M[4], r[2] = M[8];
    r[4] = M[1], r[5] = M[5], r[6] = M[9];
    r[8] = M[2], r[9] = M[6], r[10] = M[10];
    r[3] = -(M[3] * r[0] + M[7] * r[1] + M[11] * r[2]);
    r[7] = -(M[3] * r[4] + M[7] * r[5] + M[11] * r[6]);
    r[11] = -(M[3] * r[8] + M[7] * r[9] + M[11] * r[10]);
}

} // extern "C"

// world_bounds: n x 6 floats.  out: 2n entries (node 0 = copy of the root, leaves 1..n in BLAS order).
// One clustering loop for both node formats: the reference's TLASBVHNode (children 2 x 16 bit, tlas_bvh.h:10) and the
// ABI v5 node with 32-bit children (no 32 767-instance cap).
struct Tlas16 {
    static void leaf(rt_tlas_node& n, uint32_t blas) { n.blas = blas, n.left_right = 0; }
    static void join(rt_tlas_node& n, uint32_t a, uint32_t b) { n.left_right = a + (b << 16); }
};
struct Tlas32 {
    static void leaf(rt_tlas_node32& n, uint32_t blas) { n.left = 0, n.right = blas; }
    static void join(rt_tlas_node32& n, uint32_t a, uint32_t b) { n.left = a, n.right = b; }
};
template <class TNode, class F>
static int build_tlas_any(const float* world_bounds, uint32_t n, TNode* out, uint32_t* nodes_used)
{
    memset(out, 0, sizeof(TNode) * 2 * (size_t)n);
    std::vector<int> nodeIdx(n);
    int nodeIndices = (int)n;
    uint32_t used = 1;
    for (uint32_t i = 0; i < n; i++)
    {
        nodeIdx[i] = (int)used;
        memcpy(out[used].aabb_min, world_bounds + 6 * (size_t)i, 12);
        memcpy(out[used].aabb_max, world_bounds + 6 * (size_t)i + 3, 12);
        F::leaf(out[used], i);
        used++;
    }
    auto best_match = [&](int N, int A) { // FindBestMatch tlas_bvh.cpp:57-70
        float smallest = 1e30f;
        int bestB = -1;
        const TNode& a = out[nodeIdx[A]];
        for (int B = 0; B < N; B++)
            if (B != A)
            {
                const TNode& b = out[nodeIdx[B]];
                const float ex = maxf(a.aabb_max[0], b.aabb_max[0]) - minf(a.aabb_min[0], b.aabb_min[0]);
                const float ey = maxf(a.aabb_max[1], b.aabb_max[1]) - minf(a.aabb_min[1], b.aabb_min[1]);
                const float ez = maxf(a.aabb_max[2], b.aabb_max[2]) - minf(a.aabb_min[2], b.aabb_min[2]);
                const float area = ex * ey + ey * ez + ez * ex;
                if (area < smallest) smallest = area, bestB = B;
            }
        return bestB;
    };
    int A = 0, B = n > 1 ? best_match(nodeIndices, A) : -1;
    while (nodeIndices > 1)
    {
        const int Cc = best_match(nodeIndices, B);
        if (A == Cc)
        {
            const int ia = nodeIdx[A], ib = nodeIdx[B];
            TNode& nn = out[used];
            F::join(nn, (uint32_t)ia, (uint32_t)ib);
            for (int k = 0; k < 3; k++)
                nn.aabb_min[k] = minf(out[ia].aabb_min[k], out[ib].aabb_min[k]), nn.aabb_max[k] = maxf(out[ia].aabb_max[k], out[ib].aabb_max[k]);
            nodeIdx[A] = (int)used++;
            nodeIdx[B] = nodeIdx[nodeIndices - 1];
            B = best_match(--nodeIndices, A);
        }
        else A = B, B = Cc;
    }
    out[0] = out[nodeIdx[A]];
    if (nodes_used) *nodes_used = used;
    return RT_OK;
}

extern "C" {

int rtb_build_tlas(const float* world_bounds, uint32_t n, rt_tlas_node* out, uint32_t* nodes_used)
{
    if (!world_bounds || !out || n == 0) return RT_ERR_INVALID;
    if (2 * (uint64_t)n > 65535) return RT_ERR_UNSUPPORTED; // children are packed 2 x 16 bit (tlas_bvh.h:10): use rtb_build_tlas32
    return build_tlas_any<rt_tlas_node, Tlas16>(world_bounds, n, out, nodes_used);
}

int rtb_build_tlas32(const float* world_bounds, uint32_t n, rt_tlas_node32* out, uint32_t* nodes_used)
{
    if (!world_bounds || !out || n == 0) return RT_ERR_INVALID;
    return build_tlas_any<rt_tlas_node32, Tlas32>(world_bounds, n, out, nodes_used);
}

} // extern "C"
