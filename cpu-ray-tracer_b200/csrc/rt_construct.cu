// rt_construct.cu — the scene-construction steps that run on the device:
//   * layout_built_bvh      reference-layout BVH (rt_build.cu, still in device memory) -> fat nodes + leaf-ordered triangle
//                           records + shading records, no host round trip (rt_scene_create with blas.nodes == NULL);
//   * build_tlas_on_device  TLASBVH::Build / FindBestMatch (tlas_bvh.cpp:17-70), the reference's agglomerative clustering with
//                           32-bit children: no `int nodeIdx[256]` (tlas_bvh.cpp:21), no 2 x 16-bit leftRight (tlas_bvh.h:10);
//   * rt_scene_refit        BVH::Refit / BLASBVH::Refit (bvh.cpp:26-43 = blas_bvh.cpp:104-121) on the traversal layout;
//   * rt_scene_download_bvh the traversal layout read back as reference arrays (parity tests).
// Arithmetic: min / max by comparison-select exactly as the reference's float3 fminf / fmaxf (tmplmath.h:122-123, 256-257),
// surface areas with the reference's operation order; compiled with -fmad=false like the rest of the library.
#include <cooperative_groups.h>

#include <cstring>
#include <string>
#include <vector>

#include "rt_internal.h"

namespace rtb {

// ---------------------------------------------------------------------------------------------------------------------
// device-built BVH -> traversal layout
// ---------------------------------------------------------------------------------------------------------------------
// The reference numbers nodes as they split, depth first: the k-th split (pre-order) creates nodes 1 + 2k and 2 + 2k
// (bvh.cpp:100-101 with nodesUsed = 1).  So interior node i is the ((left_first - 1) / 2)-th interior node in pre-order, which
// is the index its fat node gets: a node's left subtree follows it directly.
__device__ __forceinline__ int fat_of(const rt_bvh_node& n) { return (int)((n.left_first - 1u) >> 1); }

__global__ void __launch_bounds__(256) k_layout_fat_nodes(const rt_bvh_node* __restrict__ nodes, const uint32_t total, const int fatBase,
    const int triBase, float4* __restrict__ out, uint8_t* __restrict__ leafEnd)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    {
        const rt_bvh_node p = nodes[i];
        if (p.tri_count > 0)
        {
            leafEnd[p.left_first + p.tri_count - 1] = 1; // the last triangle of a leaf carries LAST_BIT
            continue;
        }
        const rt_bvh_node L = nodes[p.left_first], R = nodes[p.left_first + 1];
        const int lref = L.tri_count > 0 ? ~(triBase + (int)L.left_first) : fatBase + fat_of(L);
        const int rref = R.tri_count > 0 ? ~(triBase + (int)R.left_first) : fatBase + fat_of(R);
        float4* f = out + 4 * (size_t)(fatBase + fat_of(p));
        f[0] = make_float4(L.aabb_min[0], L.aabb_min[1], L.aabb_max[0], L.aabb_max[1]);
        f[1] = make_float4(R.aabb_min[0], R.aabb_min[1], R.aabb_max[0], R.aabb_max[1]);
        f[2] = make_float4(L.aabb_min[2], L.aabb_max[2], R.aabb_min[2], R.aabb_max[2]);
        f[3] = make_float4(__int_as_float(lref), __int_as_float(rref), 0, 0);
    }
}

// triangle record of slot j <- triangle triangleIndices[j]; shading record k <- triangle k (rt_device.cuh layout)
__device__ __forceinline__ void write_tri_record(float4* __restrict__ o, const rt_tri& t, const int tag)
{
    // edge1 / edge2 exactly as bvh.cpp:205-206 computes them per test (fp32 subtraction)
    o[0] = make_float4(t.v0[0], t.v0[1], t.v0[2], __int_as_float(tag));
    o[1] = make_float4(t.v1[0] - t.v0[0], t.v1[1] - t.v0[1], t.v1[2] - t.v0[2], __int_as_float(t.obj_idx));
    o[2] = make_float4(t.v2[0] - t.v0[0], t.v2[1] - t.v0[1], t.v2[2] - t.v0[2], 0);
}
__device__ __forceinline__ void write_shade_record(float4* __restrict__ o, const rt_tri& t)
{
    o[0] = make_float4(t.n0[0], t.n0[1], t.n0[2], t.n1[0]);
    o[1] = make_float4(t.n1[1], t.n1[2], t.n2[0], t.n2[1]);
    o[2] = make_float4(t.n2[2], t.uv0[0], t.uv0[1], t.uv1[0]);
    o[3] = make_float4(t.uv1[1], t.uv2[0], t.uv2[1], __int_as_float(t.obj_idx));
}

__global__ void __launch_bounds__(256) k_layout_tris(const rt_tri* __restrict__ tris, const uint32_t* __restrict__ idx, const uint32_t n,
    const uint8_t* __restrict__ leafEnd, float4* __restrict__ outTris, float4* __restrict__ outShade)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    {
        const uint32_t triIdx = idx[j];
        write_tri_record(outTris + 3 * (size_t)j, tris[triIdx], (int)triIdx | (leafEnd[j] ? LAST_BIT : 0));
        write_shade_record(outShade + 4 * (size_t)j, tris[j]);
    }
}

static int grid_of(size_t items, int block = 256)
{
    const size_t g = (items + block - 1) / block;
    return (int)(g < 1 ? 1 : g > 148 * 16 ? 148 * 16 : g);
}

rt_status layout_built_bvh(const DeviceBvh& b, int fatBase, int triBase, float4* nodes, float4* tris, float4* shade, cudaStream_t stream)
{
    uint8_t* leafEnd = nullptr;
    RT_CUDA(cudaMalloc((void**)&leafEnd, b.n));
    cudaMemsetAsync(leafEnd, 0, b.n, stream);
    k_layout_fat_nodes<<<grid_of(b.total), 256, 0, stream>>>(b.nodes, b.total, fatBase, triBase, nodes, leafEnd);
    k_layout_tris<<<grid_of(b.n), 256, 0, stream>>>(b.tris, b.idx, b.n, leafEnd, tris + 3 * (size_t)triBase, shade + 4 * (size_t)triBase);
    const cudaError_t e = cudaStreamSynchronize(stream);
    cudaFree(leafEnd);
    if (!cuda_ok(e, "layout of the device-built BVH") || !cuda_ok(cudaGetLastError(), "layout of the device-built BVH")) return RT_ERR_CUDA;
    return RT_OK;
}

// BLASBVH::SetTransform (blas_bvh.cpp:369-373): the eight corners of the root box through T, float4(a, 1) * M with the
// left-to-right sum of tmplmath.cpp:155-165.  Host code (compiled with -ffp-contract=off).
void world_bounds_of(const float* root_min, const float* root_max, const float* T, float* out6)
{
    float mn[3] = { 1e30f, 1e30f, 1e30f }, mx[3] = { -1e30f, -1e30f, -1e30f };
    for (int i = 0; i < 8; i++)
    {
        const float c[3] = { i & 1 ? root_max[0] : root_min[0], i & 2 ? root_max[1] : root_min[1], i & 4 ? root_max[2] : root_min[2] };
        for (int r = 0; r < 3; r++)
        {
            const float p = T[4 * r] * c[0] + T[4 * r + 1] * c[1] + T[4 * r + 2] * c[2] + T[4 * r + 3] * 1.0f;
            mn[r] = mn[r] < p ? mn[r] : p, mx[r] = mx[r] > p ? mx[r] : p; // aabb::Grow: fminf / fmaxf of tmplmath.h
        }
    }
    for (int r = 0; r < 3; r++) out6[r] = mn[r], out6[3 + r] = mx[r];
}

// ---------------------------------------------------------------------------------------------------------------------
// TLASBVH::Build on the device
// ---------------------------------------------------------------------------------------------------------------------
// The clustering is a serial chain of FindBestMatch calls (each an argmin over the live clusters: O(n) work, first candidate
// wins ties), so the parallelism is INSIDE a call.  Two kernels build the same tree (build_tlas_on_device picks by size):
// k_build_tlas, one CTA of 1024 threads: live cluster boxes compact in a (min, max) float4 pair array in global memory, indexed like the
// reference's nodeIdx[] (.w of a min = the slot's node index); every thread scans a strided part, the block reduces (area, index) with
// warp shuffles + one shared-memory stage.  k_build_tlas_cluster (below): a thread-block cluster with the boxes in distributed shared memory.
constexpr int TLAS_THREADS = 1024;

__device__ __forceinline__ void best_of(float& area, int& idx, const float oa, const int oi)
{
    // FindBestMatch keeps the FIRST candidate with the smallest area (strict <, ascending B): smaller index wins ties
    if (oa < area || (oa == area && oi < idx)) area = oa, idx = oi;
}

__device__ int tlas_find_best_match(const float4* __restrict__ boxMin, const float4* __restrict__ boxMax, const int N, const int A, float* sArea, int* sIdx)
{
    const float4 amin = boxMin[A], amax = boxMax[A];
    float best = 1e30f;
    int bestB = 0x7fffffff;
    for (int B = threadIdx.x; B < N; B += TLAS_THREADS)
        if (B != A)
        {
            const float4 bmin = boxMin[B], bmax = boxMax[B];
            // tlas_bvh.cpp:62-66: e = fmaxf(a.max, b.max) - fminf(a.min, b.min); area = e.x * e.y + e.y * e.z + e.z * e.x
            const float ex = tfmaxf(amax.x, bmax.x) - tfminf(amin.x, bmin.x);
            const float ey = tfmaxf(amax.y, bmax.y) - tfminf(amin.y, bmin.y);
            const float ez = tfmaxf(amax.z, bmax.z) - tfminf(amin.z, bmin.z);
            const float area = ex * ey + ey * ez + ez * ex;
            if (area < best) best = area, bestB = B; // ascending B per thread: strict < keeps the first
        }
    for (int off = 16; off; off >>= 1)
    {
        const float oa = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bestB, off);
        best_of(best, bestB, oa, oi);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sArea[warp] = best, sIdx[warp] = bestB;
    __syncthreads();
    if (warp == 0)
    {
        best = sArea[lane], bestB = sIdx[lane];
        for (int off = 16; off; off >>= 1)
        {
            const float oa = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bestB, off);
            best_of(best, bestB, oa, oi);
        }
        if (lane == 0) sIdx[32] = (best < 1e30f) ? bestB : -1; // nothing below 1e30f: the reference returns -1
    }
    __syncthreads();
    const int r = sIdx[32];
    __syncthreads(); // sIdx / sArea are reused by the next call
    return r;
}

__global__ void __launch_bounds__(TLAS_THREADS, 1) k_build_tlas(const float* __restrict__ worldBounds, const int n, rt_tlas_node32* __restrict__ out,
    float4* __restrict__ boxMin, float4* __restrict__ boxMax, int* __restrict__ depthOf, int* __restrict__ depthOut)
{
    __shared__ float sArea[32];
    __shared__ int sIdx[33];
    // tlas_bvh.cpp:22-30: a leaf per BLAS, nodes 1..n
    for (int i = threadIdx.x; i < n; i += TLAS_THREADS)
    {
        const float* w = worldBounds + 6 * (size_t)i;
        rt_tlas_node32 leaf;
        leaf.aabb_min[0] = w[0], leaf.aabb_min[1] = w[1], leaf.aabb_min[2] = w[2];
        leaf.aabb_max[0] = w[3], leaf.aabb_max[1] = w[4], leaf.aabb_max[2] = w[5];
        leaf.left = 0, leaf.right = (uint32_t)i;
        out[1 + i] = leaf;
        boxMin[i] = make_float4(w[0], w[1], w[2], __int_as_float(1 + i));
        boxMax[i] = make_float4(w[3], w[4], w[5], 0);
        depthOf[1 + i] = 1;
    }
    __syncthreads();
    int nodeIndices = n, used = 1 + n;
    int A = 0, B = n > 1 ? tlas_find_best_match(boxMin, boxMax, nodeIndices, A, sArea, sIdx) : -1;
    while (nodeIndices > 1)
    {
        const int C = tlas_find_best_match(boxMin, boxMax, nodeIndices, B, sArea, sIdx);
        if (A == C)
        {
            if (threadIdx.x == 0)
            {
                const float4 amin = boxMin[A], amax = boxMax[A], bmin = boxMin[B], bmax = boxMax[B];
                const int ia = __float_as_int(amin.w), ib = __float_as_int(bmin.w);
                rt_tlas_node32 nn;
                nn.left = (uint32_t)ia, nn.right = (uint32_t)ib;
                nn.aabb_min[0] = tfminf(amin.x, bmin.x), nn.aabb_min[1] = tfminf(amin.y, bmin.y), nn.aabb_min[2] = tfminf(amin.z, bmin.z);
                nn.aabb_max[0] = tfmaxf(amax.x, bmax.x), nn.aabb_max[1] = tfmaxf(amax.y, bmax.y), nn.aabb_max[2] = tfmaxf(amax.z, bmax.z);
                out[used] = nn;
                const int da = depthOf[ia], db = depthOf[ib];
                depthOf[used] = 1 + (da > db ? da : db);
                boxMin[A] = make_float4(nn.aabb_min[0], nn.aabb_min[1], nn.aabb_min[2], __int_as_float(used));
                boxMax[A] = make_float4(nn.aabb_max[0], nn.aabb_max[1], nn.aabb_max[2], 0);
                // nodeIdx[B] = nodeIdx[nodeIndices - 1]
                boxMin[B] = boxMin[nodeIndices - 1], boxMax[B] = boxMax[nodeIndices - 1];
            }
            used++, nodeIndices--;
            __syncthreads();
            B = tlas_find_best_match(boxMin, boxMax, nodeIndices, A, sArea, sIdx);
        }
        else A = B, B = C;
    }
    if (threadIdx.x == 0)
    {
        const int root = __float_as_int(boxMin[A].w);
        out[0] = out[root]; // tlas_bvh.cpp:52
        *depthOut = depthOf[root];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same build on a THREAD-BLOCK CLUSTER with the live cluster boxes in DISTRIBUTED SHARED MEMORY.
// The single-CTA kernel above streams all live boxes (32 B each) through one SM for every FindBestMatch: 20 129 instances =
// 640 KB per call, more than that SM's L1 - 5 us per call, bound by one SM's L2 port.  Here TLAS_CLUSTER CTAs (one per SM) each
// keep an interleaved slice of the slot array (slot s lives in CTA s % TLAS_CLUSTER at index s / TLAS_CLUSTER) in their own
// shared memory, scan it in ~3 boxes per thread, and exchange the per-CTA (area, slot) minima through DSMEM: one
// cluster barrier per FindBestMatch, no global-memory traffic inside the loop.  Every CTA computes the same argmin from the same
// partials, so all stay in lock step on A / B / C without a broadcast.  Slot updates of a merge are made by the CTAs that own the
// slots.  Same tree, node for node (the scan order and the tie rule are those of the reference: smaller slot index first).
// Capacity: TLAS_CLUSTER x (shared memory / 32 B) slots = ~56 000 instances; larger scenes use the single-CTA kernel.
// ---------------------------------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int TLAS_CLUSTER = 8; // portable cluster size

struct TlasPartial { float area; int slot; };

__device__ __forceinline__ int tlas_cluster_best_match(cg::cluster_group& cluster, const float4* __restrict__ bmin, const float4* __restrict__ bmax,
    const int N, const int A, const int rank, float* sArea, int* sIdx, TlasPartial* part, int& parity)
{
    // box A from the CTA that owns slot A
    const float4* rmin = cluster.map_shared_rank(bmin, A % TLAS_CLUSTER);
    const float4* rmax = cluster.map_shared_rank(bmax, A % TLAS_CLUSTER);
    const float4 amin = rmin[A / TLAS_CLUSTER], amax = rmax[A / TLAS_CLUSTER];
    float best = 1e30f;
    int bestB = 0x7fffffff;
    for (int l = threadIdx.x; l * TLAS_CLUSTER + rank < N; l += (int)blockDim.x)
    {
        const int B = l * TLAS_CLUSTER + rank;
        if (B == A) continue;
        const float4 bmn = bmin[l], bmx = bmax[l];
        // tlas_bvh.cpp:62-66
        const float ex = tfmaxf(amax.x, bmx.x) - tfminf(amin.x, bmn.x);
        const float ey = tfmaxf(amax.y, bmx.y) - tfminf(amin.y, bmn.y);
        const float ez = tfmaxf(amax.z, bmx.z) - tfminf(amin.z, bmn.z);
        const float area = ex * ey + ey * ez + ez * ex;
        if (area < best) best = area, bestB = B; // ascending B per thread: strict < keeps the first
    }
    for (int off = 16; off; off >>= 1)
    {
        const float oa = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bestB, off);
        best_of(best, bestB, oa, oi);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sArea[warp] = best, sIdx[warp] = bestB;
    __syncthreads();
    if (warp == 0)
    {
        const int warps = (int)blockDim.x >> 5;
        best = lane < warps ? sArea[lane] : 1e30f, bestB = lane < warps ? sIdx[lane] : 0x7fffffff;
        for (int off = 16; off; off >>= 1)
        {
            const float oa = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bestB, off);
            best_of(best, bestB, oa, oi);
        }
        if (lane == 0) part[parity].area = best, part[parity].slot = bestB;
    }
    cluster.sync(); // every CTA's partial of this call is in its shared memory (and the previous call's reads are over)
    best = 1e30f, bestB = 0x7fffffff;
#pragma unroll
    for (int r = 0; r < TLAS_CLUSTER; r++)
    {
        const TlasPartial* q = cluster.map_shared_rank(part, r);
        best_of(best, bestB, q[parity].area, q[parity].slot);
    }
    parity ^= 1; // the next call writes the other buffer: a fast CTA cannot overwrite what a slow one still reads
    return best < 1e30f ? bestB : -1;
}

__global__ void __launch_bounds__(1024, 1) k_build_tlas_cluster(const float* __restrict__ worldBounds, const int n, rt_tlas_node32* __restrict__ out,
    int* __restrict__ depthOf, int* __restrict__ depthOut)
{
    extern __shared__ float4 sBoxes[]; // [slice] mins, then [slice] maxs; .w of a min = node index of the slot (the reference's nodeIdx[])
    __shared__ float sArea[32];
    __shared__ int sIdx[32];
    __shared__ TlasPartial part[2];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int slice = (n + TLAS_CLUSTER - 1) / TLAS_CLUSTER;
    float4* bmin = sBoxes;
    float4* bmax = sBoxes + slice;
    // tlas_bvh.cpp:22-30: a leaf per BLAS, nodes 1..n; this CTA's slots
    for (int l = threadIdx.x; l * TLAS_CLUSTER + rank < n; l += (int)blockDim.x)
    {
        const int i = l * TLAS_CLUSTER + rank;
        const float* w = worldBounds + 6 * (size_t)i;
        rt_tlas_node32 leaf;
        leaf.aabb_min[0] = w[0], leaf.aabb_min[1] = w[1], leaf.aabb_min[2] = w[2];
        leaf.aabb_max[0] = w[3], leaf.aabb_max[1] = w[4], leaf.aabb_max[2] = w[5];
        leaf.left = 0, leaf.right = (uint32_t)i;
        out[1 + i] = leaf;
        bmin[l] = make_float4(w[0], w[1], w[2], __int_as_float(1 + i));
        bmax[l] = make_float4(w[3], w[4], w[5], 0);
        depthOf[1 + i] = 1;
    }
    __threadfence();
    cluster.sync();
    int parity = 0;
    int nodeIndices = n, used = 1 + n;
    int A = 0, B = n > 1 ? tlas_cluster_best_match(cluster, bmin, bmax, nodeIndices, A, rank, sArea, sIdx, part, parity) : -1;
    while (nodeIndices > 1)
    {
        const int C = tlas_cluster_best_match(cluster, bmin, bmax, nodeIndices, B, rank, sArea, sIdx, part, parity);
        if (A == C)
        {
            // everyone reads the two boxes (and the last slot) before their owners overwrite them
            const int last = nodeIndices - 1;
            const float4 amin = cluster.map_shared_rank(bmin, A % TLAS_CLUSTER)[A / TLAS_CLUSTER], amax = cluster.map_shared_rank(bmax, A % TLAS_CLUSTER)[A / TLAS_CLUSTER];
            const float4 bmn = cluster.map_shared_rank(bmin, B % TLAS_CLUSTER)[B / TLAS_CLUSTER], bmx = cluster.map_shared_rank(bmax, B % TLAS_CLUSTER)[B / TLAS_CLUSTER];
            const float4 lmin = cluster.map_shared_rank(bmin, last % TLAS_CLUSTER)[last / TLAS_CLUSTER], lmax = cluster.map_shared_rank(bmax, last % TLAS_CLUSTER)[last / TLAS_CLUSTER];
            const int ia = __float_as_int(amin.w), ib = __float_as_int(bmn.w);
            const float4 nmin = make_float4(tfminf(amin.x, bmn.x), tfminf(amin.y, bmn.y), tfminf(amin.z, bmn.z), __int_as_float(used));
            const float4 nmax = make_float4(tfmaxf(amax.x, bmx.x), tfmaxf(amax.y, bmx.y), tfmaxf(amax.z, bmx.z), 0);
            cluster.sync();
            if (threadIdx.x == 0)
            {
                if (rank == 0)
                {
                    rt_tlas_node32 nn;
                    nn.left = (uint32_t)ia, nn.right = (uint32_t)ib;
                    nn.aabb_min[0] = nmin.x, nn.aabb_min[1] = nmin.y, nn.aabb_min[2] = nmin.z;
                    nn.aabb_max[0] = nmax.x, nn.aabb_max[1] = nmax.y, nn.aabb_max[2] = nmax.z;
                    out[used] = nn;
                    const int da = depthOf[ia], db = depthOf[ib];
                    depthOf[used] = 1 + (da > db ? da : db);
                }
                // nodeIdx[A] = nodesUsed++;  nodeIdx[B] = nodeIdx[nodeIndices - 1]  (tlas_bvh.cpp:46-47; when the last slot IS A, B receives the new node)
                if (A % TLAS_CLUSTER == rank) bmin[A / TLAS_CLUSTER] = nmin, bmax[A / TLAS_CLUSTER] = nmax;
                if (B % TLAS_CLUSTER == rank) bmin[B / TLAS_CLUSTER] = last == A ? nmin : lmin, bmax[B / TLAS_CLUSTER] = last == A ? nmax : lmax;
            }
            used++, nodeIndices--;
            // (the next call's first cluster barrier comes too late for these writes: it would let a CTA scan before the owners wrote)
            cluster.sync();
            B = tlas_cluster_best_match(cluster, bmin, bmax, nodeIndices, A, rank, sArea, sIdx, part, parity);
        }
        else A = B, B = C;
    }
    if (rank == 0 && threadIdx.x == 0)
    {
        __threadfence();
        const int root = __float_as_int(cluster.map_shared_rank(bmin, A % TLAS_CLUSTER)[A / TLAS_CLUSTER].w);
        out[0] = out[root]; // tlas_bvh.cpp:52 (merged nodes were written by this thread; a leaf root is slot 0's, written by this CTA)
        *depthOut = depthOf[root];
    }
    cluster.sync(); // no CTA may exit while another still reads its shared memory
}

// d_out: 2n entries; d_depth: one int (levels of the tree, a single leaf = 1)
rt_status build_tlas_on_device(int device, const float* d_world_bounds, uint32_t n, rt_tlas_node32* d_out, int* d_depth, cudaStream_t stream)
{
    float4 *boxMin = nullptr, *boxMax = nullptr;
    int* depthOf = nullptr;
    struct Free { float4*& a; float4*& b; int*& c; ~Free() { cudaFree(a), cudaFree(b), cudaFree(c); } } guard{ boxMin, boxMax, depthOf };
    if (cudaMalloc((void**)&depthOf, 2 * (size_t)n * 4) != cudaSuccess) { cudaGetLastError(); set_error("rt_build_tlas: out of device memory"); return RT_ERR_CUDA; }
    // cluster + distributed shared memory when the live boxes fit the cluster's shared memory (RT_B200_TLAS_BUILD=single forces the other)
    const size_t slice = ((size_t)n + TLAS_CLUSTER - 1) / TLAS_CLUSTER, smem = slice * 32;
    int maxSmem = 0;
    cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    // measured (profiles/r2_tlas_build.txt): below ~8 000 instances the single CTA wins (a call costs ~2 us of cluster barrier + DSMEM
    // latency however few boxes there are); RT_B200_TLAS_BUILD=cluster / single forces one of them (cluster needs n >= 64)
    const char* mode = getenv("RT_B200_TLAS_BUILD");
    const bool force = mode && strcmp(mode, "cluster") == 0;
    bool useCluster = n >= (force ? 64u : 8192u) && smem + 1024 <= (size_t)maxSmem && !(mode && strcmp(mode, "single") == 0);
    if (useCluster)
    {
        cudaLaunchConfig_t cfg = {};
        // the loop is a chain of cluster barriers: a barrier of 8 x 256 threads completes sooner than one of 8 x 1024, and 256 threads
        // still scan a 20 000-instance scene in ~10 boxes each (RT_B200_TLAS_THREADS for experiments)
        int threads = 256;
        if (const char* e = getenv("RT_B200_TLAS_THREADS")) { const int v = atoi(e); if (v == 128 || v == 256 || v == 512 || v == 1024) threads = v; }
        cfg.gridDim = dim3(TLAS_CLUSTER), cfg.blockDim = dim3(threads), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = TLAS_CLUSTER, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
        cfg.attrs = &attr, cfg.numAttrs = 1;
        const int ni = (int)n;
        if (cudaFuncSetAttribute(k_build_tlas_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaLaunchKernelEx(&cfg, k_build_tlas_cluster, d_world_bounds, ni, d_out, depthOf, d_depth) != cudaSuccess)
        {
            cudaGetLastError(); // no cluster of this shape on this device / partition: the single-CTA kernel builds the same tree
            useCluster = false;
        }
    }
    if (!useCluster)
    {
        if (cudaMalloc((void**)&boxMin, (size_t)n * 16) != cudaSuccess || cudaMalloc((void**)&boxMax, (size_t)n * 16) != cudaSuccess)
        {
            cudaGetLastError();
            set_error("rt_build_tlas: out of device memory");
            return RT_ERR_CUDA;
        }
        k_build_tlas<<<1, TLAS_THREADS, 0, stream>>>(d_world_bounds, (int)n, d_out, boxMin, boxMax, depthOf, d_depth);
    }
    const cudaError_t e = cudaStreamSynchronize(stream);
    if (!cuda_ok(e, "rt_build_tlas kernel") || !cuda_ok(cudaGetLastError(), "rt_build_tlas kernel")) return RT_ERR_CUDA;
    return RT_OK;
}

// reference-order TLAS (leaves 1..n, merged nodes n+1..2n-1 in creation order, root = the last one) -> fat nodes:
// merged node j gets fat index fatBase + (2n - 1 - j), so the root is fatBase and parents precede their children
__global__ void __launch_bounds__(256) k_layout_tlas(const rt_tlas_node32* __restrict__ t, const uint32_t n, const int fatBase, float4* __restrict__ out)
{
    const uint32_t last = 2 * n - 1;
    for (uint32_t j = n + 1 + blockIdx.x * blockDim.x + threadIdx.x; j <= last; j += gridDim.x * blockDim.x)
    {
        const rt_tlas_node32 p = t[j];
        const rt_tlas_node32 L = t[p.left], R = t[p.right];
        const int lref = L.left == 0 ? ~(INSTANCE_BIT | (int)L.right) : fatBase + (int)(last - p.left);
        const int rref = R.left == 0 ? ~(INSTANCE_BIT | (int)R.right) : fatBase + (int)(last - p.right);
        float4* f = out + 4 * (size_t)(fatBase + (int)(last - j));
        f[0] = make_float4(L.aabb_min[0], L.aabb_min[1], L.aabb_max[0], L.aabb_max[1]);
        f[1] = make_float4(R.aabb_min[0], R.aabb_min[1], R.aabb_max[0], R.aabb_max[1]);
        f[2] = make_float4(L.aabb_min[2], L.aabb_max[2], R.aabb_min[2], R.aabb_max[2]);
        f[3] = make_float4(__int_as_float(lref), __int_as_float(rref), 0, 0);
    }
}

rt_status layout_built_tlas(const rt_tlas_node32* d_tlas, uint32_t n, int fatBase, float4* nodes, cudaStream_t stream)
{
    if (n < 2) return RT_OK; // a single instance: the root is the leaf itself, no fat node
    k_layout_tlas<<<grid_of(n), 256, 0, stream>>>(d_tlas, n, fatBase, nodes);
    RT_CUDA(cudaStreamSynchronize(stream));
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Refit on the traversal layout
// ---------------------------------------------------------------------------------------------------------------------
// new positions / normals / uvs into the mesh's triangle and shading records (tags, objIdx and leaf order stay)
__global__ void __launch_bounds__(256) k_refit_tris(const rt_tri* __restrict__ tris, const uint32_t n, float4* __restrict__ recs, float4* __restrict__ shade)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    {
        float4* o = recs + 3 * (size_t)j;
        const int tag = __float_as_int(o[0].w);
        const float keepObj = o[1].w;
        const rt_tri& t = tris[tag & ~LAST_BIT];
        o[0] = make_float4(t.v0[0], t.v0[1], t.v0[2], __int_as_float(tag));
        o[1] = make_float4(t.v1[0] - t.v0[0], t.v1[1] - t.v0[1], t.v1[2] - t.v0[2], keepObj);
        o[2] = make_float4(t.v2[0] - t.v0[0], t.v2[1] - t.v0[1], t.v2[2] - t.v0[2], 0);
        float4* s = shade + 4 * (size_t)j;
        const float keepShadeObj = s[3].w;
        const rt_tri& u = tris[j];
        s[0] = make_float4(u.n0[0], u.n0[1], u.n0[2], u.n1[0]);
        s[1] = make_float4(u.n1[1], u.n1[2], u.n2[0], u.n2[1]);
        s[2] = make_float4(u.n2[2], u.uv0[0], u.uv0[1], u.uv1[0]);
        s[3] = make_float4(u.uv1[1], u.uv2[0], u.uv2[1], keepShadeObj);
    }
}

// parent links of the mesh's fat nodes: parent[f] = (parent fat node - fatBase) * 2 + which child; -1 for the root
__global__ void __launch_bounds__(256) k_refit_parents(const float4* __restrict__ nodes, const int fatBase, const int fatCount, int* __restrict__ parent, int* __restrict__ arrived)
{
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < fatCount; f += gridDim.x * blockDim.x)
    {
        const float4 c = nodes[4 * (size_t)(fatBase + f) + 3];
        const int refs[2] = { __float_as_int(c.x), __float_as_int(c.y) };
        for (int k = 0; k < 2; k++)
            if (refs[k] >= 0) parent[refs[k] - fatBase] = f * 2 + k;
        arrived[f] = 0;
        if (f == 0) parent[0] = -1; // the root is the first fat node of a mesh in both layouts
    }
}

// UpdateNodeBounds (bvh.cpp:45-61) of the leaf that starts at `slot`, from the (already updated) triangle records:
// v1 = v0 + edge1 would not reproduce the vertex, so the vertices come from the new rt_tri array
__device__ __forceinline__ void leaf_bounds(const float4* __restrict__ recs, const rt_tri* __restrict__ tris, int slot, float* mn, float* mx)
{
    for (int a = 0; a < 3; a++) mn[a] = 1e30f, mx[a] = -1e30f;
    while (true)
    {
        const int tag = __float_as_int(recs[3 * (size_t)slot].w);
        const rt_tri& t = tris[tag & ~LAST_BIT];
        for (int a = 0; a < 3; a++)
        {
            mn[a] = tfminf(mn[a], t.v0[a]), mn[a] = tfminf(mn[a], t.v1[a]), mn[a] = tfminf(mn[a], t.v2[a]);
            mx[a] = tfmaxf(mx[a], t.v0[a]), mx[a] = tfmaxf(mx[a], t.v1[a]), mx[a] = tfmaxf(mx[a], t.v2[a]);
        }
        if (tag & LAST_BIT) break;
        slot++;
    }
}

__device__ __forceinline__ void store_child_box(float4* __restrict__ f, const int k, const float* mn, const float* mx)
{
    // n0 = L (min.x, min.y, max.x, max.y), n1 = R likewise, n2 = (L.min.z, L.max.z, R.min.z, R.max.z)
    f[k] = make_float4(mn[0], mn[1], mx[0], mx[1]);
    float* z = (float*)(f + 2) + 2 * k;
    z[0] = mn[2], z[1] = mx[2];
}

// Bottom-up: every fat node first refreshes the boxes of its LEAF children, then the thread that completes a node (both
// children final) writes the node's own box - fminf / fmaxf of the child boxes, bvh.cpp:39-40 - into its parent's record
// and continues there if it was the second to arrive.  skipNode1: the reference never refits node 1 = the root's left child.
__global__ void __launch_bounds__(256) k_refit_up(float4* __restrict__ nodes, const float4* __restrict__ recs, const rt_tri* __restrict__ tris,
    const int fatBase, const int fatCount, const int triBase, const int* __restrict__ parent, int* __restrict__ arrived, const int skipNode1)
{
    for (int f0 = blockIdx.x * blockDim.x + threadIdx.x; f0 < fatCount; f0 += gridDim.x * blockDim.x)
    {
        float4* f = nodes + 4 * (size_t)(fatBase + f0);
        const float4 c = f[3];
        const int refs[2] = { __float_as_int(c.x), __float_as_int(c.y) };
        int interior = 0;
        for (int k = 0; k < 2; k++)
        {
            if (refs[k] >= 0) { interior++; continue; }
            if (skipNode1 && f0 == 0 && k == 0) continue;
            float mn[3], mx[3];
            leaf_bounds(recs, tris, ~refs[k], mn, mx);
            store_child_box(f, k, mn, mx);
        }
        if (interior > 0)
        {
            // wait for the interior children: the last of them to arrive carries on
            __threadfence();
            if (atomicAdd(&arrived[f0], 2 - interior) + (2 - interior) < 2) continue;
            __threadfence();
        }
        // this node is final: propagate
        int cur = f0;
        while (true)
        {
            const int p = parent[cur];
            if (p < 0) break;
            const int pf = p >> 1, k = p & 1;
            const float4* me = nodes + 4 * (size_t)(fatBase + cur);
            const float4 a = __ldcg(me), b = __ldcg(me + 1), z = __ldcg(me + 2);
            float mn[3] = { tfminf(a.x, b.x), tfminf(a.y, b.y), tfminf(z.x, z.z) };
            float mx[3] = { tfmaxf(a.z, b.z), tfmaxf(a.w, b.w), tfmaxf(z.y, z.w) };
            if (!(skipNode1 && pf == 0 && k == 0)) store_child_box(nodes + 4 * (size_t)(fatBase + pf), k, mn, mx);
            __threadfence();
            if (atomicAdd(&arrived[pf], 1) + 1 < 2) break;
            __threadfence(); // the sibling's stores to pf's record happened before its arrival: order this thread's reads after ours
            cur = pf;
        }
    }
}

} // namespace rtb

using namespace rtb;

extern "C" {

rt_status rt_build_tlas(int device, const float* world_bounds, uint32_t n, rt_tlas_node32* out, uint32_t* nodes_used, double* device_ms)
{
    if (!world_bounds || !out || n == 0 || n > (1u << 29)) { set_error("rt_build_tlas: bad argument"); return RT_ERR_INVALID; }
    if (device < 0 || device >= rt_device_count()) { set_error("rt_build_tlas: no such CUDA device (there is no CPU fallback)"); return RT_ERR_NO_DEVICE; }
    RT_CUDA(cudaSetDevice(device));
    float* dw = nullptr;
    rt_tlas_node32* dout = nullptr;
    int* ddepth = nullptr;
    struct Free { float*& a; rt_tlas_node32*& b; int*& c; ~Free() { cudaFree(a), cudaFree(b), cudaFree(c); } } guard{ dw, dout, ddepth };
    RT_CUDA(cudaMalloc((void**)&dw, (size_t)n * 24));
    RT_CUDA(cudaMalloc((void**)&dout, 2 * (size_t)n * sizeof(rt_tlas_node32)));
    RT_CUDA(cudaMalloc((void**)&ddepth, 4));
    RT_CUDA(cudaMemcpy(dw, world_bounds, (size_t)n * 24, cudaMemcpyHostToDevice));
    RT_CUDA(cudaMemset(dout, 0, 2 * (size_t)n * sizeof(rt_tlas_node32)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0, nullptr);
    const rt_status st = build_tlas_on_device(device, dw, n, dout, ddepth, nullptr);
    cudaEventRecord(e1, nullptr);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    if (st != RT_OK) return st;
    RT_CUDA(cudaMemcpy(out, dout, 2 * (size_t)n * sizeof(rt_tlas_node32), cudaMemcpyDeviceToHost));
    if (nodes_used) *nodes_used = 2 * n;
    if (device_ms) *device_ms = ms;
    return RT_OK;
}

rt_status rt_scene_refit(rt_scene* s, uint32_t blas_index, const rt_tri* tris, uint32_t tri_count, uint32_t flags)
{
    if (!s || !tris) { set_error("rt_scene_refit: null argument"); return RT_ERR_INVALID; }
    if (s->d.kind != RT_SCENE_FLAT && s->d.kind != RT_SCENE_TLAS) { set_error("rt_scene_refit: BVH scenes only (the reference has no Refit for its KD-tree / grid)"); return RT_ERR_UNSUPPORTED; }
    if (blas_index >= s->blas_geometry.size()) { set_error("rt_scene_refit: BLAS index out of range"); return RT_ERR_INVALID; }
    const Geometry& g = s->geometries[s->blas_geometry[blas_index]];
    if (tri_count != g.triCount) { set_error("rt_scene_refit: the triangle count of the mesh cannot change (Refit keeps the topology)"); return RT_ERR_INVALID; }
    if ((flags & RT_REFIT_REBUILD_TLAS) && s->d.kind != RT_SCENE_TLAS) { set_error("rt_scene_refit: RT_REFIT_REBUILD_TLAS on a scene without TLAS"); return RT_ERR_INVALID; }
    RT_CUDA(cudaSetDevice(s->device));
    std::lock_guard<std::mutex> lock(s->scratch_mutex);
    rt_tri* dTris = nullptr;
    int *parent = nullptr, *arrived = nullptr;
    struct Free { rt_tri*& a; int*& b; int*& c; ~Free() { cudaFree(a), cudaFree(b), cudaFree(c); } } guard{ dTris, parent, arrived };
    RT_CUDA(cudaMalloc((void**)&dTris, (size_t)tri_count * sizeof(rt_tri)));
    RT_CUDA(cudaMemcpyAsync(dTris, tris, (size_t)tri_count * sizeof(rt_tri), cudaMemcpyHostToDevice, s->stream));
    k_refit_tris<<<grid_of(tri_count), 256, 0, s->stream>>>(dTris, tri_count, s->tris + 3 * (size_t)g.triBase, s->shade + 4 * (size_t)g.triBase);
    if (g.fatCount > 0)
    {
        RT_CUDA(cudaMalloc((void**)&parent, (size_t)g.fatCount * 4));
        RT_CUDA(cudaMalloc((void**)&arrived, (size_t)g.fatCount * 4));
        k_refit_parents<<<grid_of(g.fatCount), 256, 0, s->stream>>>(s->nodes, g.fatBase, g.fatCount, parent, arrived);
        k_refit_up<<<grid_of(g.fatCount), 256, 0, s->stream>>>(s->nodes, s->tris, dTris, g.fatBase, g.fatCount, g.triBase, parent, arrived,
                                                              (flags & RT_REFIT_ALL_NODES) ? 0 : 1);
    }
    RT_CUDA(cudaStreamSynchronize(s->stream));
    RT_CUDA(cudaGetLastError());
    if (flags & RT_REFIT_REBUILD_TLAS)
    {
        // SetTransform for every instance (root box of its mesh through T), then TLASBVH::Build; same fat-node range
        const uint32_t n = (uint32_t)s->blas_geometry.size();
        std::vector<float> roots(6 * s->geometries.size()), wb(6 * (size_t)n);
        for (size_t gi = 0; gi < s->geometries.size(); gi++)
        {
            const Geometry& gg = s->geometries[gi];
            float* r = &roots[6 * gi];
            if (gg.fatCount == 0)
            {
                // a mesh of <= 2 triangles is a single leaf with no fat node: its root box is kept on the host (rt_scene_create) and
                // follows a refit here (UpdateNodeBounds of the root leaf, bvh.cpp:45-61)
                if (&gg == &g)
                {
                    float* rb = &s->root_boxes[6 * gi];
                    for (int a = 0; a < 3; a++) rb[a] = 1e30f, rb[3 + a] = -1e30f;
                    for (uint32_t j = 0; j < gg.triCount; j++)
                        for (int a = 0; a < 3; a++)
                        {
                            const float v[3] = { tris[j].v0[a], tris[j].v1[a], tris[j].v2[a] };
                            for (int q = 0; q < 3; q++) rb[a] = rb[a] < v[q] ? rb[a] : v[q], rb[3 + a] = rb[3 + a] > v[q] ? rb[3 + a] : v[q];
                        }
                }
                memcpy(r, &s->root_boxes[6 * gi], 24);
                continue;
            }
            float4 f[3];
            RT_CUDA(cudaMemcpy(f, s->nodes + 4 * (size_t)gg.fatBase, 48, cudaMemcpyDeviceToHost));
            // BVH root box as Refit leaves it: fminf / fmaxf of its children (bvh.cpp:39-40)
            r[0] = f[0].x < f[1].x ? f[0].x : f[1].x, r[1] = f[0].y < f[1].y ? f[0].y : f[1].y, r[2] = f[2].x < f[2].z ? f[2].x : f[2].z;
            r[3] = f[0].z > f[1].z ? f[0].z : f[1].z, r[4] = f[0].w > f[1].w ? f[0].w : f[1].w, r[5] = f[2].y > f[2].w ? f[2].y : f[2].w;
        }
        for (uint32_t i = 0; i < n; i++)
        {
            const float* r = &roots[6 * (size_t)s->blas_geometry[i]];
            world_bounds_of(r, r + 3, &s->blas_T[16 * (size_t)i], &wb[6 * (size_t)i]);
        }
        float* dw = nullptr;
        rt_tlas_node32* dt = nullptr;
        int* dd = nullptr;
        struct Free2 { float*& a; rt_tlas_node32*& b; int*& c; ~Free2() { cudaFree(a), cudaFree(b), cudaFree(c); } } guard2{ dw, dt, dd };
        RT_CUDA(cudaMalloc((void**)&dw, (size_t)n * 24));
        RT_CUDA(cudaMalloc((void**)&dt, 2 * (size_t)n * sizeof(rt_tlas_node32)));
        RT_CUDA(cudaMalloc((void**)&dd, 4));
        RT_CUDA(cudaMemcpy(dw, wb.data(), (size_t)n * 24, cudaMemcpyHostToDevice));
        rt_status st = build_tlas_on_device(s->device, dw, n, dt, dd, s->stream);
        if (st != RT_OK) return st;
        int depth = 0;
        RT_CUDA(cudaMemcpy(&depth, dd, 4, cudaMemcpyDeviceToHost));
        const int entries = (depth - 1) + 1 + (s->max_blas_depth > 0 ? s->max_blas_depth - 1 : 0);
        if (entries > STACK_SIZE) { set_error("rt_scene_refit: the rebuilt TLAS is too deep for the traversal stack; the scene keeps its previous TLAS"); return RT_ERR_UNSUPPORTED; }
        if ((st = layout_built_tlas(dt, n, s->tlas_fat_base, s->nodes, s->stream)) != RT_OK) return st;
        s->stack_entries = entries;
    }
    return RT_OK;
}

rt_status rt_scene_validate(rt_scene* s)
{
    if (!s) { set_error("rt_scene_validate: null argument"); return RT_ERR_INVALID; }
    if (s->d.kind != RT_SCENE_FLAT && s->d.kind != RT_SCENE_TLAS) { set_error("rt_scene_validate: BVH scenes only"); return RT_ERR_UNSUPPORTED; }
    RT_CUDA(cudaSetDevice(s->device));
    const size_t nFat = s->node_count, nSlots = s->tri_count, nInst = s->inst_count;
    std::vector<float4> fat(4 * nFat), recs(3 * nSlots), inst(4 * nInst);
    if (nFat) RT_CUDA(cudaMemcpy(fat.data(), s->nodes, fat.size() * 16, cudaMemcpyDeviceToHost));
    if (nSlots) RT_CUDA(cudaMemcpy(recs.data(), s->tris, recs.size() * 16, cudaMemcpyDeviceToHost));
    if (nInst) RT_CUDA(cudaMemcpy(inst.data(), s->inst, inst.size() * 16, cudaMemcpyDeviceToHost));
    auto asi = [](float f) { int i; memcpy(&i, &f, 4); return i; };
    auto bad = [&](const std::string& what) { set_error("rt_scene_validate: " + what); return RT_ERR_INVALID; };
    // mesh of every triangle slot (tags index the mesh's own shading records)
    std::vector<int> meshOfSlot(nSlots, -1);
    for (size_t gi = 0; gi < s->geometries.size(); gi++)
    {
        const Geometry& g = s->geometries[gi];
        if ((size_t)g.triBase + g.triCount > nSlots || (size_t)g.fatBase + g.fatCount > nFat) return bad("mesh " + std::to_string(gi) + " lies outside the device arrays");
        for (uint32_t j = 0; j < g.triCount; j++)
        {
            if (meshOfSlot[g.triBase + j] != -1) return bad("meshes overlap in the triangle array");
            meshOfSlot[g.triBase + j] = (int)gi;
            const uint32_t triIdx = (uint32_t)(asi(recs[3 * (size_t)(g.triBase + j)].w) & ~LAST_BIT);
            if (triIdx >= g.triCount) return bad("triangle tag " + std::to_string(triIdx) + " outside its mesh (" + std::to_string(g.triCount) + " triangles)");
        }
        if (g.triCount && !(asi(recs[3 * (size_t)(g.triBase + g.triCount - 1)].w) & LAST_BIT)) return bad("the last triangle of mesh " + std::to_string(gi) + " does not end a leaf");
    }
    std::vector<uint8_t> seen(nFat, 0);
    // depth-first walk of one tree: returns its depth in levels (a single leaf = 1), -1 on a violation (error set)
    auto walk = [&](int rootRef, bool tlasLevel, int lo, int hi, const char* what) -> int {
        struct Item { int ref; int depth; };
        std::vector<Item> todo;
        todo.push_back({ rootRef, 1 });
        int depth = 0;
        while (!todo.empty())
        {
            const Item it = todo.back();
            todo.pop_back();
            if (it.depth > depth) depth = it.depth;
            if (it.ref >= 0)
            {
                if (it.ref < lo || it.ref >= hi) { bad(std::string(what) + ": node reference " + std::to_string(it.ref) + " outside [" + std::to_string(lo) + ", " + std::to_string(hi) + ")"); return -1; }
                if (seen[it.ref]) { bad(std::string(what) + ": node " + std::to_string(it.ref) + " is reached twice (not a tree)"); return -1; }
                seen[it.ref] = 1;
                todo.push_back({ asi(fat[4 * (size_t)it.ref + 3].x), it.depth + 1 });
                todo.push_back({ asi(fat[4 * (size_t)it.ref + 3].y), it.depth + 1 });
                continue;
            }
            const int payload = ~it.ref;
            if (payload == SENTINEL_PAYLOAD) { bad(std::string(what) + ": the stack marker appears as a child"); return -1; }
            if (tlasLevel)
            {
                if (!(payload & INSTANCE_BIT) || (size_t)(payload & ~INSTANCE_BIT) >= nInst) { bad(std::string(what) + ": TLAS leaf without a valid instance"); return -1; }
            }
            else
            {
                if ((payload & INSTANCE_BIT) || (size_t)payload >= nSlots) { bad(std::string(what) + ": leaf slot " + std::to_string(payload) + " outside the triangle array"); return -1; }
                size_t j = (size_t)payload;
                const int mesh = meshOfSlot[j];
                while (!(asi(recs[3 * j].w) & LAST_BIT))
                    if (++j >= nSlots || meshOfSlot[j] != mesh) { bad(std::string(what) + ": a leaf's triangle run does not end inside its mesh"); return -1; }
            }
        }
        return depth;
    };
    int maxBlas = 0;
    for (size_t gi = 0; gi < s->geometries.size(); gi++)
    {
        const Geometry& g = s->geometries[gi];
        const int d = walk(g.rootRef, false, g.fatBase, g.fatBase + g.fatCount, "mesh BVH");
        if (d < 0) return RT_ERR_INVALID;
        if (d > maxBlas) maxBlas = d;
        for (int f = g.fatBase; f < g.fatBase + g.fatCount; f++)
            if (!seen[f]) return bad("mesh " + std::to_string(gi) + ": fat node " + std::to_string(f) + " is unreachable");
    }
    int entries = maxBlas > 0 ? maxBlas - 1 : 0;
    if (s->d.kind == RT_SCENE_TLAS)
    {
        for (size_t i = 0; i < nInst; i++)
        {
            const int root = asi(inst[4 * i + 3].x);
            const Geometry& g = s->geometries[s->blas_geometry[i]];
            if (root != g.rootRef) return bad("instance " + std::to_string(i) + " does not point at the root of its mesh");
        }
        const int d = walk(s->d.root_ref, true, s->tlas_fat_base, s->tlas_fat_base + s->tlas_fat_count, "TLAS");
        if (d < 0) return RT_ERR_INVALID;
        std::vector<uint8_t> used(nInst, 0);
        for (int f = s->tlas_fat_base; f < s->tlas_fat_base + s->tlas_fat_count; f++)
        {
            if (!seen[f]) return bad("TLAS fat node " + std::to_string(f) + " is unreachable");
            for (int k = 0; k < 2; k++)
            {
                const int ref = asi(k ? fat[4 * (size_t)f + 3].y : fat[4 * (size_t)f + 3].x);
                if (ref < 0)
                {
                    uint8_t& u = used[~ref & ~INSTANCE_BIT];
                    if (u) return bad("an instance hangs under two TLAS leaves");
                    u = 1;
                }
            }
        }
        entries = (d - 1) + 1 + (maxBlas > 0 ? maxBlas - 1 : 0);
    }
    else if (s->d.root_ref != s->geometries[s->blas_geometry[0]].rootRef) return bad("the scene root is not the root of its BVH");
    if (entries > STACK_SIZE) return bad("a ray can have " + std::to_string(entries) + " entries pending: more than the traversal stack holds");
    if (entries > s->stack_entries) return bad("the trees are deeper (" + std::to_string(entries) + " pending entries) than the scene recorded (" + std::to_string(s->stack_entries) + "): the stream kernel sizes its stack from that");
    return RT_OK;
}

rt_status rt_scene_download_bvh(rt_scene* s, uint32_t blas_index, rt_bvh_node* nodes_out, uint32_t* tri_indices_out, uint32_t* nodes_used)
{
    if (!s || !nodes_out || !tri_indices_out) { set_error("rt_scene_download_bvh: null argument"); return RT_ERR_INVALID; }
    if (s->d.kind != RT_SCENE_FLAT && s->d.kind != RT_SCENE_TLAS) { set_error("rt_scene_download_bvh: BVH scenes only"); return RT_ERR_UNSUPPORTED; }
    if (blas_index >= s->blas_geometry.size()) { set_error("rt_scene_download_bvh: BLAS index out of range"); return RT_ERR_INVALID; }
    const Geometry& g = s->geometries[s->blas_geometry[blas_index]];
    RT_CUDA(cudaSetDevice(s->device));
    std::vector<float4> fat(4 * (size_t)g.fatCount), recs(3 * (size_t)g.triCount);
    if (g.fatCount) RT_CUDA(cudaMemcpy(fat.data(), s->nodes + 4 * (size_t)g.fatBase, fat.size() * 16, cudaMemcpyDeviceToHost));
    RT_CUDA(cudaMemcpy(recs.data(), s->tris + 3 * (size_t)g.triBase, recs.size() * 16, cudaMemcpyDeviceToHost));
    auto asi = [](float f) { int i; memcpy(&i, &f, 4); return i; };
    for (uint32_t j = 0; j < g.triCount; j++) tri_indices_out[j] = (uint32_t)(asi(recs[3 * (size_t)j].w) & ~LAST_BIT);
    auto leaf = [&](rt_bvh_node& o, int ref) {
        const int slot = ~ref - g.triBase;
        int count = 1;
        while (!(asi(recs[3 * (size_t)(slot + count - 1)].w) & LAST_BIT)) count++;
        o.left_first = (uint32_t)slot, o.tri_count = (uint32_t)count;
    };
    memset(nodes_out, 0, sizeof(rt_bvh_node) * (2 * (size_t)g.triCount - 1));
    if (g.fatCount == 0)
    {
        leaf(nodes_out[0], g.rootRef);
        for (int a = 0; a < 3; a++) nodes_out[0].aabb_min[a] = 1e30f, nodes_out[0].aabb_max[a] = -1e30f;
        if (nodes_used) *nodes_used = 1;
        return RT_OK;
    }
    // pre-order walk: the k-th interior node visited owns the reference nodes 1 + 2k and 2 + 2k (bvh.cpp:100-101)
    struct Item { int fatIdx; uint32_t refNode; };
    std::vector<Item> todo;
    todo.push_back({ g.rootRef - g.fatBase, 0u });
    uint32_t splits = 0;
    while (!todo.empty())
    {
        const Item it = todo.back();
        todo.pop_back();
        const float4* f = &fat[4 * (size_t)it.fatIdx];
        const uint32_t l = 1 + 2 * splits, r = l + 1;
        splits++;
        nodes_out[it.refNode].left_first = l, nodes_out[it.refNode].tri_count = 0;
        rt_bvh_node& L = nodes_out[l];
        rt_bvh_node& R = nodes_out[r];
        L.aabb_min[0] = f[0].x, L.aabb_min[1] = f[0].y, L.aabb_max[0] = f[0].z, L.aabb_max[1] = f[0].w;
        R.aabb_min[0] = f[1].x, R.aabb_min[1] = f[1].y, R.aabb_max[0] = f[1].z, R.aabb_max[1] = f[1].w;
        L.aabb_min[2] = f[2].x, L.aabb_max[2] = f[2].y, R.aabb_min[2] = f[2].z, R.aabb_max[2] = f[2].w;
        const int refs[2] = { asi(f[3].x), asi(f[3].y) };
        // the left subtree is numbered before the right one: push right first
        if (refs[1] >= 0) todo.push_back({ refs[1] - g.fatBase, r }); else leaf(R, refs[1]);
        if (refs[0] >= 0) todo.push_back({ refs[0] - g.fatBase, l }); else leaf(L, refs[0]);
    }
    {
        rt_bvh_node& root = nodes_out[0];
        const rt_bvh_node &L = nodes_out[1], &R = nodes_out[2];
        for (int a = 0; a < 3; a++)
            root.aabb_min[a] = L.aabb_min[a] < R.aabb_min[a] ? L.aabb_min[a] : R.aabb_min[a], root.aabb_max[a] = L.aabb_max[a] > R.aabb_max[a] ? L.aabb_max[a] : R.aabb_max[a];
    }
    if (nodes_used) *nodes_used = 1 + 2 * splits;
    return RT_OK;
}

} // extern "C"
