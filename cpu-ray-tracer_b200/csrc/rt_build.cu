// rt_build.cu — the reference's binned-SAH BVH builder on the GPU (the step
// immediately before the hot path; the host builder takes 13 s for 10 M triangles).
//
// Replaces BVH::Build / BLASBVH::Build (bvh.cpp:4-24, 45-178 = blas_bvh.cpp:82-257) and produces the SAME
// arrays bit for bit: node boxes, node numbering, and the order of triangleIndices.
//   * The reference recurses depth-first; here every level of the tree is processed at once (all triangles
//     in parallel).  What a node decides depends only on the SET of its triangles: centroid bounds, the
//     8 bins per axis and the node boxes are min / max / count reductions, exact in any order (float
//     atomics through the usual integer-ordering trick).
//   * The reference partitions a node's index range with a two-cursor swap loop (bvh.cpp:88-95), whose
//     result depends on the order of the elements.  That loop has a closed form (derived and checked against
//     the loop on 200 000 random cases, tests/test_host_build.py): with G = number of "left" elements,
//       left-region left elements stay; the m-th misplaced left-region element (ascending) goes to the end of
//       the range for m = 1, else right below where the (m-1)-th right-region left element (descending) was;
//       that element fills the hole of the m-th misplaced one; the element AT position G, if it belongs
//       right, behaves like one more misplaced element; every other right element moves one slot left.
//     Ranks come from segmented prefix sums (cub::DeviceScan::ExclusiveSumByKey, key = node of the position).
//   * Node numbers: the reference hands out ids as nodes split, depth first (children of the k-th split get
//     1 + 2k and 2 + 2k).  The level-order build numbers nodes breadth first, then one bottom-up pass counts
//     the interior nodes below every node and one top-down pass derives each node's pre-order rank.
// Arithmetic: this file is compiled with -fmad=false like the rest of the library; every cost expression
// keeps the reference's operation order (including 0 * inf = NaN for empty sides, which never wins).
#include <cub/cub.cuh>

#include <vector>

#include "rt_internal.h"

namespace rtb {

constexpr int BINS = 8; // BVH_BINS, bvh.h:7

struct BNode {
    float bmin[3], bmax[3];
    float cmin[3], cmax[3]; // centroid bounds (FindBestSplitPlane, bvh.cpp:129-136)
    uint32_t start, count;
    int left;               // breadth-first id of the left child (right = left + 1); -1 = leaf
    int splitting;          // this level: 1 while the node is going to split
    int axis;
    float splitPos;
    uint32_t leftCount, misplaced;
    int rank;               // exclusive rank among the splitting nodes of the level
    int interior, pre, finalId;
};

__device__ __forceinline__ void atomic_min_float(float* addr, float v)
{
    if (__float_as_int(v) >= 0) atomicMin((int*)addr, __float_as_int(v));
    else atomicMax((unsigned int*)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v)
{
    if (__float_as_int(v) >= 0) atomicMax((int*)addr, __float_as_int(v));
    else atomicMin((unsigned int*)addr, __float_as_uint(v));
}

// per triangle: centroid (as the loader stored it), vertex bounds; identity permutation; everything in node 0
__global__ void k_build_init(const rt_tri* __restrict__ tris, int n, float* __restrict__ cen, float* __restrict__ tmin, float* __restrict__ tmax,
    uint32_t* __restrict__ idx, int* __restrict__ posNode, BNode* __restrict__ nodes)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const rt_tri& t = tris[i];
        for (int a = 0; a < 3; a++)
        {
            cen[3 * (size_t)i + a] = t.centroid[a];
            const float lo = tfminf(tfminf(t.v0[a], t.v1[a]), t.v2[a]), hi = tfmaxf(tfmaxf(t.v0[a], t.v1[a]), t.v2[a]);
            tmin[3 * (size_t)i + a] = lo, tmax[3 * (size_t)i + a] = hi;
            atomic_min_float(&nodes[0].bmin[a], lo), atomic_max_float(&nodes[0].bmax[a], hi); // UpdateNodeBounds(root)
        }
        idx[i] = (uint32_t)i, posNode[i] = 0;
    }
}

__global__ void k_build_node_init(BNode* nodes, int first, int count, uint32_t start0, uint32_t count0)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        BNode& n = nodes[first + i];
        for (int a = 0; a < 3; a++) n.bmin[a] = 1e30f, n.bmax[a] = -1e30f; // bvh.cpp:48-49
        n.left = -1, n.splitting = 0, n.interior = 0, n.pre = 0, n.finalId = 0;
        if (count == 1) n.start = start0, n.count = count0;
    }
}

// level set-up: centroid bounds and bins of every node that may split (triCount > 2, bvh.cpp:67)
__global__ void k_build_level_reset(BNode* nodes, int first, int count, float* __restrict__ binB, int* __restrict__ binC)
{
    const int total = count * 3 * BINS;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    {
        const int k = i / (3 * BINS);
        BNode& n = nodes[first + k];
        if (i % (3 * BINS) == 0)
        {
            for (int a = 0; a < 3; a++) n.cmin[a] = 1e30f, n.cmax[a] = -1e30f;
            n.splitting = n.count > 2 ? 1 : 0;
            n.rank = 0;
        }
        float* b = binB + 6 * (size_t)i;
        b[0] = b[1] = b[2] = 1e34f, b[3] = b[4] = b[5] = -1e34f; // aabb(), tmplmath.h:626
        binC[i] = 0;
    }
}

// The three reductions into per-node records (centroid bounds, bins, child boxes) are where the build spends
// its time: at the top levels millions of triangles target the same few addresses (95 % of the first version's
// time, profiles/r1_gpu_bvh_build.jsonl).  Positions of a node are contiguous, so every CTA walks one
// contiguous chunk of positions and keeps the record of the node it is currently inside in shared memory
// (fast atomics, no global contention); it is flushed with one global atomic per word when the CTA crosses
// into another node and at the end.  Positions of other nodes inside a step go straight to global memory,
// which is where nodes are small and contention is low anyway.
constexpr int BUILD_BLOCK = 256;

__global__ void __launch_bounds__(BUILD_BLOCK) k_build_centroid_bounds(int n, int chunk, int first, const int* __restrict__ posNode,
    const uint32_t* __restrict__ idx, const float* __restrict__ cen, BNode* nodes)
{
    __shared__ float sbox[6];
    __shared__ int skey;
    const int begin = blockIdx.x * chunk, end = min(n, begin + chunk);
    int kcur = -1;
    for (int base = begin; base < end; base += BUILD_BLOCK)
    {
        const int pos = base + threadIdx.x;
        int k = pos < end ? posNode[pos] : -1;
        if (k >= 0 && (k < first || nodes[k].count <= 2)) k = -1; // not a node that may split at this level
        if (threadIdx.x == 0) skey = k;
        __syncthreads();
        const int kstep = skey;
        if (kstep != kcur)
        {
            if (kcur >= 0 && threadIdx.x < 6)
            {
                if (threadIdx.x < 3) atomic_min_float(&nodes[kcur].cmin[threadIdx.x], sbox[threadIdx.x]);
                else atomic_max_float(&nodes[kcur].cmax[threadIdx.x - 3], sbox[threadIdx.x]);
            }
            __syncthreads();
            if (threadIdx.x < 6) sbox[threadIdx.x] = threadIdx.x < 3 ? 1e30f : -1e30f;
            kcur = kstep;
            __syncthreads();
        }
        if (k >= 0)
        {
            const float* c = cen + 3 * (size_t)idx[pos];
            if (k == kcur)
                for (int a = 0; a < 3; a++) atomic_min_float(&sbox[a], c[a]), atomic_max_float(&sbox[3 + a], c[a]);
            else
                for (int a = 0; a < 3; a++) atomic_min_float(&nodes[k].cmin[a], c[a]), atomic_max_float(&nodes[k].cmax[a], c[a]);
        }
        __syncthreads();
    }
    if (kcur >= 0 && threadIdx.x < 6)
    {
        if (threadIdx.x < 3) atomic_min_float(&nodes[kcur].cmin[threadIdx.x], sbox[threadIdx.x]);
        else atomic_max_float(&nodes[kcur].cmax[threadIdx.x - 3], sbox[threadIdx.x]);
    }
}

// bvh.cpp:138-149: bin index per axis, bin count and bin bounds
__global__ void __launch_bounds__(BUILD_BLOCK) k_build_bin(int n, int chunk, int first, const int* __restrict__ posNode, const uint32_t* __restrict__ idx,
    const float* __restrict__ cen, const float* __restrict__ tmin, const float* __restrict__ tmax, const BNode* __restrict__ nodes,
    float* binB, int* binC)
{
    __shared__ float sB[3 * BINS * 6];
    __shared__ int sC[3 * BINS];
    __shared__ int skey;
    const int begin = blockIdx.x * chunk, end = min(n, begin + chunk);
    int kcur = -1;
    auto flush = [&]() {
        if (kcur < 0) return;
        const size_t slot0 = (size_t)(kcur - first) * 3 * BINS;
        for (int i = threadIdx.x; i < 3 * BINS; i += BUILD_BLOCK)
            if (sC[i])
            {
                atomicAdd(&binC[slot0 + i], sC[i]);
                float* bb = binB + 6 * (slot0 + i);
                for (int c = 0; c < 3; c++) atomic_min_float(&bb[c], sB[6 * i + c]), atomic_max_float(&bb[3 + c], sB[6 * i + 3 + c]);
            }
    };
    for (int base = begin; base < end; base += BUILD_BLOCK)
    {
        const int pos = base + threadIdx.x;
        int k = pos < end ? posNode[pos] : -1;
        if (k >= 0 && (k < first || nodes[k].count <= 2)) k = -1;
        if (threadIdx.x == 0) skey = k;
        __syncthreads();
        const int kstep = skey;
        if (kstep != kcur)
        {
            flush();
            __syncthreads();
            for (int i = threadIdx.x; i < 3 * BINS; i += BUILD_BLOCK)
            {
                sC[i] = 0;
                for (int c = 0; c < 3; c++) sB[6 * i + c] = 1e34f, sB[6 * i + 3 + c] = -1e34f;
            }
            kcur = kstep;
            __syncthreads();
        }
        if (k >= 0)
        {
            const uint32_t t = idx[pos];
            for (int a = 0; a < 3; a++)
            {
                const float lo = nodes[k].cmin[a], hi = nodes[k].cmax[a];
                if (lo == hi) continue;
                const float scale = BINS / (hi - lo);
                int b = (int)((cen[3 * (size_t)t + a] - lo) * scale);
                if (b > BINS - 1) b = BINS - 1;
                if (k == kcur)
                {
                    const int i = a * BINS + b;
                    atomicAdd(&sC[i], 1);
                    for (int c = 0; c < 3; c++) atomic_min_float(&sB[6 * i + c], tmin[3 * (size_t)t + c]), atomic_max_float(&sB[6 * i + 3 + c], tmax[3 * (size_t)t + c]);
                }
                else
                {
                    const size_t slot = ((size_t)(k - first) * 3 + a) * BINS + b;
                    atomicAdd(&binC[slot], 1);
                    float* bb = binB + 6 * slot;
                    for (int c = 0; c < 3; c++) atomic_min_float(&bb[c], tmin[3 * (size_t)t + c]), atomic_max_float(&bb[3 + c], tmax[3 * (size_t)t + c]);
                }
            }
        }
        __syncthreads();
    }
    flush();
}

__device__ __forceinline__ float box_area(const float* mn, const float* mx) // aabb::Area, tmplmath.h:593-598
{
    const float e0 = mx[0] - mn[0], e1 = mx[1] - mn[1], e2 = mx[2] - mn[2];
    return tfmaxf(0.0f, e0 * e1 + e0 * e2 + e1 * e2);
}

// FindBestSplitPlane (bvh.cpp:124-178) + the split / no-split decision (bvh.cpp:69-76), one thread per node
__global__ void k_build_split(BNode* nodes, int first, int count, const float* __restrict__ binB, const int* __restrict__ binC)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        BNode& n = nodes[first + i];
        if (!n.splitting) continue;
        float bestCost = 1e30f, splitPos = 0;
        int axis = 0;
        for (int a = 0; a < 3; a++)
        {
            const float lo = n.cmin[a], hi = n.cmax[a];
            if (lo == hi) continue;
            const float* bb = binB + 6 * (((size_t)i * 3 + a) * BINS);
            const int* bc = binC + ((size_t)i * 3 + a) * BINS;
            float leftArea[BINS - 1], rightArea[BINS - 1];
            int leftCount[BINS - 1], rightCount[BINS - 1];
            float lmn[3] = { 1e34f, 1e34f, 1e34f }, lmx[3] = { -1e34f, -1e34f, -1e34f };
            float rmn[3] = { 1e34f, 1e34f, 1e34f }, rmx[3] = { -1e34f, -1e34f, -1e34f };
            int leftSum = 0, rightSum = 0;
            for (int b = 0; b < BINS - 1; b++)
            {
                leftSum += bc[b], leftCount[b] = leftSum;
                const float* L = bb + 6 * b;
                for (int c = 0; c < 3; c++) lmn[c] = L[c] < lmn[c] ? L[c] : lmn[c], lmx[c] = L[3 + c] > lmx[c] ? L[3 + c] : lmx[c];
                leftArea[b] = box_area(lmn, lmx);
                rightSum += bc[BINS - 1 - b], rightCount[BINS - 2 - b] = rightSum;
                const float* R = bb + 6 * (BINS - 1 - b);
                for (int c = 0; c < 3; c++) rmn[c] = R[c] < rmn[c] ? R[c] : rmn[c], rmx[c] = R[3 + c] > rmx[c] ? R[3 + c] : rmx[c];
                rightArea[BINS - 2 - b] = box_area(rmn, rmx);
            }
            const float scale = (hi - lo) / BINS;
            for (int b = 0; b < BINS - 1; b++)
            {
                const float cost = leftCount[b] * leftArea[b] + rightCount[b] * rightArea[b];
                if (cost < bestCost) axis = a, splitPos = lo + scale * (b + 1), bestCost = cost;
            }
        }
        const float ex = n.bmax[0] - n.bmin[0], ey = n.bmax[1] - n.bmin[1], ez = n.bmax[2] - n.bmin[2];
        const float noSplit = n.count * (ex * ey + ey * ez + ez * ex); // CalculateNodeCost bvh.cpp:117-122
        if (bestCost >= noSplit) n.splitting = 0;
        n.axis = axis, n.splitPos = splitPos;
    }
}

// which side every triangle of a splitting node goes to (bvh.cpp:91)
__global__ void k_build_flags(int n, int first, const int* __restrict__ posNode, const uint32_t* __restrict__ idx,
    const float* __restrict__ cen, const BNode* __restrict__ nodes, int* __restrict__ good)
{
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n; pos += gridDim.x * blockDim.x)
    {
        const int k = posNode[pos];
        int g = 0;
        if (k >= first && nodes[k].splitting) g = cen[3 * (size_t)idx[pos] + nodes[k].axis] < nodes[k].splitPos ? 1 : 0;
        good[pos] = g;
    }
}

// leftCount per node (bvh.cpp:98-99: no split when one side would be empty)
__global__ void k_build_left_count(int n, int first, const int* __restrict__ posNode, const int* __restrict__ good,
    const int* __restrict__ goodBefore, BNode* nodes)
{
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n; pos += gridDim.x * blockDim.x)
    {
        const int k = posNode[pos];
        if (k < first || !nodes[k].splitting) continue;
        BNode& nd = nodes[k];
        if ((uint32_t)pos != nd.start + nd.count - 1) continue;
        const uint32_t left = (uint32_t)(goodBefore[pos] + good[pos]);
        nd.leftCount = left;
        if (left == 0 || left == nd.count) nd.splitting = 0;
    }
}

__global__ void k_build_flags2(int n, int first, const int* __restrict__ posNode, const int* __restrict__ good, const BNode* __restrict__ nodes,
    int* __restrict__ leftBad, int* __restrict__ rightGood)
{
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n; pos += gridDim.x * blockDim.x)
    {
        const int k = posNode[pos];
        int lb = 0, rg = 0;
        if (k >= first && nodes[k].splitting)
        {
            const bool left = (uint32_t)pos - nodes[k].start < nodes[k].leftCount;
            lb = (!good[pos] && left) ? 1 : 0, rg = (good[pos] && !left) ? 1 : 0;
        }
        leftBad[pos] = lb, rightGood[pos] = rg;
    }
}

__global__ void k_build_misplaced(int n, int first, const int* __restrict__ posNode, const int* __restrict__ rightGood,
    const int* __restrict__ rgBefore, BNode* nodes)
{
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n; pos += gridDim.x * blockDim.x)
    {
        const int k = posNode[pos];
        if (k < first || !nodes[k].splitting) continue;
        if ((uint32_t)pos == nodes[k].start + nodes[k].count - 1) nodes[k].misplaced = (uint32_t)(rgBefore[pos] + rightGood[pos]);
    }
}

__global__ void k_build_scatter(int n, int first, const int* __restrict__ posNode, const BNode* __restrict__ nodes,
    const int* __restrict__ leftBad, const int* __restrict__ lbRank, const int* __restrict__ rightGood, const int* __restrict__ rgBefore,
    int* __restrict__ leftBadPos, int* __restrict__ goodPosDesc)
{
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n; pos += gridDim.x * blockDim.x)
    {
        const int k = posNode[pos];
        if (k < first || !nodes[k].splitting) continue;
        const BNode& nd = nodes[k];
        if (leftBad[pos]) leftBadPos[nd.start + lbRank[pos]] = pos;
        if (rightGood[pos]) goodPosDesc[nd.start + (nd.misplaced - 1 - rgBefore[pos])] = pos;
    }
}

// the closed form of the swap loop bvh.cpp:88-95 (see the file header)
__global__ void k_build_permute(int n, int first, const int* __restrict__ posNode, const BNode* __restrict__ nodes, const int* __restrict__ good,
    const int* __restrict__ leftBad, const int* __restrict__ lbRank, const int* __restrict__ rightGood, const int* __restrict__ rgBefore,
    const int* __restrict__ leftBadPos, const int* __restrict__ goodPosDesc, const uint32_t* __restrict__ idxIn, uint32_t* __restrict__ idxOut)
{
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n; pos += gridDim.x * blockDim.x)
    {
        const int k = posNode[pos];
        int dst = pos;
        if (k >= first && nodes[k].splitting)
        {
            const BNode& nd = nodes[k];
            const int end = (int)(nd.start + nd.count - 1), M = (int)nd.misplaced;
            const bool boundary = (uint32_t)pos == nd.start + nd.leftCount; // first slot of the right region
            if (leftBad[pos] || (boundary && !good[pos]))
            {
                const int m = leftBad[pos] ? lbRank[pos] + 1 : M + 1;
                dst = m == 1 ? end : goodPosDesc[nd.start + (m - 2)] - 1;
            }
            else if (rightGood[pos]) dst = leftBadPos[nd.start + (M - 1 - rgBefore[pos])];
            else if (!good[pos]) dst = pos - 1; // every other right element moves one slot left
        }
        idxOut[dst] = idxIn[pos];
    }
}

__global__ void k_build_split_flags(const BNode* __restrict__ nodes, int first, int count, int* __restrict__ flags)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) flags[i] = nodes[first + i].splitting;
}

// bvh.cpp:101-109: children of every splitting node, numbered breadth first in node order
__global__ void k_build_children(BNode* nodes, int first, int count, const int* __restrict__ rank, int nextFirst)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        BNode& nd = nodes[first + i];
        if (!nd.splitting) continue;
        const int l = nextFirst + 2 * rank[i];
        nd.left = l;
        nodes[l].start = nd.start, nodes[l].count = nd.leftCount;
        nodes[l + 1].start = nd.start + nd.leftCount, nodes[l + 1].count = nd.count - nd.leftCount;
    }
}

// UpdateNodeBounds of both children (bvh.cpp:110-111) + the node of every position for the next level;
// the two child boxes of the node a CTA is inside live in shared memory (see k_build_centroid_bounds)
__global__ void __launch_bounds__(BUILD_BLOCK) k_build_descend(int n, int chunk, int first, int* __restrict__ posNode, const uint32_t* __restrict__ idx,
    const float* __restrict__ tmin, const float* __restrict__ tmax, BNode* nodes)
{
    __shared__ float sbox[12]; // left min, left max, right min, right max
    __shared__ int skey;
    const int begin = blockIdx.x * chunk, end = min(n, begin + chunk);
    int kcur = -1;
    auto flush = [&]() {
        if (kcur < 0 || threadIdx.x >= 12) return;
        BNode& c = nodes[nodes[kcur].left + threadIdx.x / 6];
        const int w = threadIdx.x % 6;
        if (w < 3) atomic_min_float(&c.bmin[w], sbox[threadIdx.x]);
        else atomic_max_float(&c.bmax[w - 3], sbox[threadIdx.x]);
    };
    for (int base = begin; base < end; base += BUILD_BLOCK)
    {
        const int pos = base + threadIdx.x;
        int k = pos < end ? posNode[pos] : -1;
        if (k >= 0 && (k < first || !nodes[k].splitting)) k = -1;
        if (threadIdx.x == 0) skey = k;
        __syncthreads();
        const int kstep = skey;
        if (kstep != kcur)
        {
            flush();
            __syncthreads();
            if (threadIdx.x < 12) sbox[threadIdx.x] = (threadIdx.x % 6) < 3 ? 1e30f : -1e30f;
            kcur = kstep;
            __syncthreads();
        }
        if (k >= 0)
        {
            const int side = ((uint32_t)pos - nodes[k].start < nodes[k].leftCount) ? 0 : 1;
            const int child = nodes[k].left + side;
            posNode[pos] = child;
            const uint32_t t = idx[pos];
            if (k == kcur)
                for (int a = 0; a < 3; a++) atomic_min_float(&sbox[6 * side + a], tmin[3 * (size_t)t + a]), atomic_max_float(&sbox[6 * side + 3 + a], tmax[3 * (size_t)t + a]);
            else
                for (int a = 0; a < 3; a++)
                    atomic_min_float(&nodes[child].bmin[a], tmin[3 * (size_t)t + a]), atomic_max_float(&nodes[child].bmax[a], tmax[3 * (size_t)t + a]);
        }
        __syncthreads();
    }
    flush();
}

__global__ void k_build_count_interior(BNode* nodes, int first, int count)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        BNode& nd = nodes[first + i];
        nd.interior = nd.left >= 0 ? 1 + nodes[nd.left].interior + nodes[nd.left + 1].interior : 0;
    }
}

// depth-first numbering of the reference: children of the k-th split in pre-order get 1 + 2k and 2 + 2k
__global__ void k_build_number(BNode* nodes, int first, int count)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        const BNode& nd = nodes[first + i];
        if (nd.left < 0) continue;
        BNode& l = nodes[nd.left];
        BNode& r = nodes[nd.left + 1];
        l.finalId = 1 + 2 * nd.pre, r.finalId = 2 + 2 * nd.pre;
        l.pre = nd.pre + 1, r.pre = nd.pre + 1 + l.interior;
    }
}

__global__ void k_build_emit(const BNode* __restrict__ nodes, int count, rt_bvh_node* __restrict__ out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        const BNode& nd = nodes[i];
        rt_bvh_node o;
        for (int a = 0; a < 3; a++) o.aabb_min[a] = nd.bmin[a], o.aabb_max[a] = nd.bmax[a];
        o.left_first = nd.left >= 0 ? (uint32_t)nodes[nd.left].finalId : nd.start;
        o.tri_count = nd.left >= 0 ? 0u : nd.count;
        out[nd.finalId] = o;
    }
}

struct DevBuf {
    std::vector<void*> ptrs;
    ~DevBuf() { for (void* p : ptrs) cudaFree(p); }
    template <class T> bool alloc(T** p, size_t count)
    {
        *p = nullptr;
        if (cudaMalloc((void**)p, (count ? count : 1) * sizeof(T)) != cudaSuccess) return false;
        ptrs.push_back(*p);
        return true;
    }
    void keep(void* p) // hand an allocation over to the caller
    {
        for (void*& q : ptrs) if (q == p) q = nullptr;
    }
};

} // namespace rtb

using namespace rtb;

namespace rtb {

void DeviceBvh::release()
{
    cudaFree(tris), cudaFree(nodes), cudaFree(idx);
    tris = nullptr, nodes = nullptr, idx = nullptr;
}

// The build proper: host triangles in, the reference-layout arrays left in device memory (rt_internal.h DeviceBvh).
// rt_build_bvh downloads them; rt_scene_create lays them out for traversal without leaving the device (rt_construct.cu).
rt_status build_bvh_on_device(int device, const rt_tri* tris, uint32_t n, DeviceBvh& result)
{
    if (!tris || n == 0 || n > (1u << 30)) { set_error("rt_build_bvh: bad argument"); return RT_ERR_INVALID; }
    int devices = 0;
    if (cudaGetDeviceCount(&devices) != cudaSuccess || device < 0 || device >= devices)
    {
        cudaGetLastError();
        set_error("rt_build_bvh: no such CUDA device (there is no CPU fallback)");
        return RT_ERR_NO_DEVICE;
    }
    RT_CUDA(cudaSetDevice(device));
    const int N = (int)n;
    const size_t maxNodes = 2 * (size_t)n;
    DevBuf B;
    rt_tri* dTris;
    float *cen, *tmin, *tmax, *binB;
    uint32_t *idxA, *idxB;
    int *posNode, *good, *goodBefore, *leftBad, *lbRank, *rightGood, *rgBefore, *leftBadPos, *goodPosDesc, *binC, *flags, *rank;
    BNode* nodes;
    rt_bvh_node* dOut;
    // a level has at most n / 3 + 1 nodes with more than two triangles, but bins are indexed by level position
    const size_t maxLevelNodes = (size_t)n + 1;
    bool ok = B.alloc(&dTris, n) && B.alloc(&cen, 3 * (size_t)n) && B.alloc(&tmin, 3 * (size_t)n) && B.alloc(&tmax, 3 * (size_t)n) &&
              B.alloc(&idxA, n) && B.alloc(&idxB, n) && B.alloc(&posNode, n) && B.alloc(&good, n) && B.alloc(&goodBefore, n) &&
              B.alloc(&leftBad, n) && B.alloc(&lbRank, n) && B.alloc(&rightGood, n) && B.alloc(&rgBefore, n) && B.alloc(&leftBadPos, n) &&
              B.alloc(&goodPosDesc, n) && B.alloc(&nodes, maxNodes) && B.alloc(&dOut, maxNodes) && B.alloc(&flags, maxLevelNodes) &&
              B.alloc(&rank, maxLevelNodes);
    if (!ok) { cudaGetLastError(); set_error("rt_build_bvh: out of device memory"); return RT_ERR_CUDA; }
    size_t binCapacity = 0; // in nodes; grown per level
    binB = nullptr, binC = nullptr;
    // scan scratch
    size_t tempBytes = 0, t2 = 0;
    cub::DeviceScan::ExclusiveSumByKey(nullptr, tempBytes, posNode, good, goodBefore, N);
    cub::DeviceScan::ExclusiveSum(nullptr, t2, flags, rank, N);
    if (t2 > tempBytes) tempBytes = t2;
    void* temp;
    if (!B.alloc((char**)&temp, tempBytes)) { cudaGetLastError(); set_error("rt_build_bvh: out of device memory"); return RT_ERR_CUDA; }

    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int grid = sms * 8, block = 256;
    // contiguous chunk of positions per CTA for the privatised reductions (a multiple of the CTA size)
    const int chunk = (int)((((size_t)N + grid - 1) / grid + BUILD_BLOCK - 1) / BUILD_BLOCK * BUILD_BLOCK);
    auto gridFor = [&](size_t items) { const size_t g = (items + block - 1) / block; return (int)(g < (size_t)grid ? (g ? g : 1) : grid); };

    RT_CUDA(cudaMemcpy(dTris, tris, (size_t)n * sizeof(rt_tri), cudaMemcpyHostToDevice));
    struct Events {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        ~Events() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
    } ev;
    struct Bins { // grown per level, released on every exit path
        float* b = nullptr;
        int* c = nullptr;
        ~Bins() { cudaFree(b), cudaFree(c); }
    } bins;
    cudaEventCreate(&ev.e0), cudaEventCreate(&ev.e1);
    cudaEvent_t e0 = ev.e0, e1 = ev.e1;
    cudaEventRecord(e0);
    k_build_node_init<<<1, 32>>>(nodes, 0, 1, 0u, n);
    k_build_init<<<grid, block>>>(dTris, N, cen, tmin, tmax, idxA, posNode, nodes);
    std::vector<std::pair<int, int>> levels; // (first node, node count)
    int first = 0, count = 1, total = 1;
    uint32_t *idx = idxA, *idxNext = idxB;
    while (count > 0)
    {
        levels.push_back({ first, count });
        if ((size_t)count > binCapacity)
        {
            // (re)allocate the bins for this level; levels first grow, then shrink
            cudaFree(bins.b), cudaFree(bins.c);
            bins.b = nullptr, bins.c = nullptr;
            binCapacity = (size_t)count * 2;
            if (cudaMalloc((void**)&bins.b, binCapacity * 3 * BINS * 6 * sizeof(float)) != cudaSuccess ||
                cudaMalloc((void**)&bins.c, binCapacity * 3 * BINS * sizeof(int)) != cudaSuccess)
            {
                cudaGetLastError();
                set_error("rt_build_bvh: out of device memory (bins)");
                return RT_ERR_CUDA;
            }
            binB = bins.b, binC = bins.c;
        }
        k_build_level_reset<<<gridFor((size_t)count * 3 * BINS), block>>>(nodes, first, count, binB, binC);
        k_build_centroid_bounds<<<grid, BUILD_BLOCK>>>(N, chunk, first, posNode, idx, cen, nodes);
        k_build_bin<<<grid, BUILD_BLOCK>>>(N, chunk, first, posNode, idx, cen, tmin, tmax, nodes, binB, binC);
        k_build_split<<<gridFor(count), block>>>(nodes, first, count, binB, binC);
        k_build_flags<<<grid, block>>>(N, first, posNode, idx, cen, nodes, good);
        cub::DeviceScan::ExclusiveSumByKey(temp, tempBytes, posNode, good, goodBefore, N);
        k_build_left_count<<<grid, block>>>(N, first, posNode, good, goodBefore, nodes);
        k_build_flags2<<<grid, block>>>(N, first, posNode, good, nodes, leftBad, rightGood);
        cub::DeviceScan::ExclusiveSumByKey(temp, tempBytes, posNode, leftBad, lbRank, N);
        cub::DeviceScan::ExclusiveSumByKey(temp, tempBytes, posNode, rightGood, rgBefore, N);
        k_build_misplaced<<<grid, block>>>(N, first, posNode, rightGood, rgBefore, nodes);
        k_build_scatter<<<grid, block>>>(N, first, posNode, nodes, leftBad, lbRank, rightGood, rgBefore, leftBadPos, goodPosDesc);
        k_build_permute<<<grid, block>>>(N, first, posNode, nodes, good, leftBad, lbRank, rightGood, rgBefore, leftBadPos, goodPosDesc, idx, idxNext);
        { uint32_t* t = idx; idx = idxNext; idxNext = t; }
        k_build_split_flags<<<gridFor(count), block>>>(nodes, first, count, flags);
        cub::DeviceScan::ExclusiveSum(temp, tempBytes, flags, rank, count);
        int lastRank = 0, lastFlag = 0;
        RT_CUDA(cudaMemcpy(&lastRank, rank + count - 1, 4, cudaMemcpyDeviceToHost));
        RT_CUDA(cudaMemcpy(&lastFlag, flags + count - 1, 4, cudaMemcpyDeviceToHost));
        const int splits = lastRank + lastFlag;
        const int nextFirst = first + count, nextCount = 2 * splits;
        if (nextCount > 0)
        {
            k_build_node_init<<<gridFor(nextCount), block>>>(nodes, nextFirst, nextCount, 0u, 0u);
            k_build_children<<<gridFor(count), block>>>(nodes, first, count, rank, nextFirst);
            k_build_descend<<<grid, BUILD_BLOCK>>>(N, chunk, first, posNode, idx, tmin, tmax, nodes);
        }
        total += nextCount;
        first = nextFirst, count = nextCount;
    }
    for (int L = (int)levels.size() - 1; L >= 0; L--) k_build_count_interior<<<gridFor(levels[L].second), block>>>(nodes, levels[L].first, levels[L].second);
    for (size_t L = 0; L < levels.size(); L++) k_build_number<<<gridFor(levels[L].second), block>>>(nodes, levels[L].first, levels[L].second);
    k_build_emit<<<gridFor(total), block>>>(nodes, total, dOut);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (!cuda_ok(err, "rt_build_bvh kernels") || !cuda_ok(cudaGetLastError(), "rt_build_bvh kernels")) return RT_ERR_CUDA;
    // the three result arrays outlive the scratch buffers
    B.keep(dTris), B.keep(dOut), B.keep(idx);
    result.tris = dTris, result.nodes = dOut, result.idx = idx;
    result.n = n, result.total = (uint32_t)total, result.depth = (int)levels.size(), result.ms = ms;
    return RT_OK;
}

} // namespace rtb

extern "C" rt_status rt_build_bvh(int device, const rt_tri* tris, uint32_t n, rt_bvh_node* nodes_out, uint32_t* tri_indices_out,
    uint32_t* nodes_used, double* device_ms)
{
    if (!tris || !nodes_out || !tri_indices_out || n == 0 || n > (1u << 30)) { set_error("rt_build_bvh: bad argument"); return RT_ERR_INVALID; }
    struct Holder { DeviceBvh b; ~Holder() { b.release(); } } h;
    const rt_status st = build_bvh_on_device(device, tris, n, h.b);
    if (st != RT_OK) return st;
    RT_CUDA(cudaMemcpy(nodes_out, h.b.nodes, (size_t)h.b.total * sizeof(rt_bvh_node), cudaMemcpyDeviceToHost));
    RT_CUDA(cudaMemcpy(tri_indices_out, h.b.idx, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (nodes_used) *nodes_used = h.b.total;
    if (device_ms) *device_ms = h.b.ms;
    return RT_OK;
}
