// rt_device.cuh — device-side core of the B200 ray tracer: scene layout, slab / triangle tests,
// the two-level ordered stack traversal, analytic primitives, textures and hit shading queries.
//
// Numerics contract (DESIGN.md "FP discipline"): this translation unit is compiled with -fmad=false,
// IEEE division and square root, no fast-math.  Every expression keeps the reference's operation
// order so that hits (t, u, v, objIdx, triIdx) are bit-identical to the reference's CPU code; the
// reference lines each routine follows are cited next to it (paths relative to /root/reference).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
// RT_B200_GLIBC_MATH (default 1): expf / atan2f / acosf of the shading code are glibc 2.39's routines restated for the
// device (rt_glibc_math.cuh: same bits as the host libm the reference's CPU build calls) instead of CUDA's (<= 2 ulp away).
#ifndef RT_B200_GLIBC_MATH
#define RT_B200_GLIBC_MATH 1
#endif
// the two halves separately, for A/B builds only (tools/ncu_ab_libm.sh): Beer's law expf, sky lookup atan2f / acosf
#ifndef RT_B200_GLIBC_EXPF
#define RT_B200_GLIBC_EXPF RT_B200_GLIBC_MATH
#endif
#ifndef RT_B200_GLIBC_SKY
#define RT_B200_GLIBC_SKY RT_B200_GLIBC_MATH
#endif
#include "rt_glibc_math.cuh"
#if RT_B200_GLIBC_EXPF
#define rt_expf rt_glibc_expf
#else
#define rt_expf expf
#endif
#if RT_B200_GLIBC_SKY
#define rt_atan2f rt_glibc_atan2f
#define rt_acosf rt_glibc_acosf
#else
#define rt_atan2f atan2f
#define rt_acosf acosf
#endif

namespace rtb {

// ------------------------------------------------------------------------------------------------
// Device scene layout (built by rt_scene.cu from the reference-layout arrays of rt_scene_desc)
//
//   nodes   "fat" BVH2 nodes, 64 B = 4 x float4, one per INTERIOR node of a reference BVH / TLAS:
//             n0 = (L.min.x, L.min.y | L.max.x, L.max.y)
//             n1 = (R.min.x, R.min.y | R.max.x, R.max.y)
//             n2 = (L.min.z, L.max.z | R.min.z, R.max.z)
//             n3 = (int left_ref, int right_ref, -, -)
//           Both child boxes travel in one 64-byte, 64-byte-aligned record (two sectors), so an
//           interior visit costs one node fetch instead of the reference's parent + 2 children.
//           The boxes are bit copies of the reference's, the test order is the reference's.
//           The component order pairs the values that meet the same ray constants - (x, y) planes with
//           (O.x, O.y) / (rD.x, rD.y), z planes with (O.z, O.z) / (rD.z, rD.z) - so that the 24 subtractions and
//           multiplications of the two slab tests are 12 packed FADD2 / FMUL2 instructions (slab_both below).
//   ref     >= 0: index of a fat node.   < 0: leaf, payload = ~ref:
//             payload == SENTINEL            stack marker: leave the current instance
//             payload & INSTANCE_BIT         TLAS leaf: instance id = payload & ~INSTANCE_BIT
//             otherwise                      first triangle slot of a triangle leaf
//   tris    leaf-ordered triangles (the order of the reference's triangleIndices), 48 B = 3 x float4:
//             t0 = (v0.xyz, int triIdx | LAST_BIT)      LAST_BIT marks the last triangle of a leaf
//             t1 = (v1-v0, int objIdx)                  objIdx: Tri::objIdx (flat BVH only)
//             t2 = (v2-v0, -)
//           edge vectors are precomputed with the same fp32 subtraction the reference does per test.
//   inst    per BLAS instance, 64 B: rows 0..2 of invT (3 x float4) + (int root_ref, int objIdx, -, -)
//   shade   per triangle (global id = instance tri_base + triIdx), 64 B:
//             s0 = (n0.xyz, n1.x) s1 = (n1.yz, n2.xy) s2 = (n2.z, uv0.xy, uv1.x) s3 = (uv1.y, uv2.xy, int objIdx)
//   inst_shade per instance, 64 B: rows 0..2 of T + (int tri_base, -, -, -)
// ------------------------------------------------------------------------------------------------
constexpr int SENTINEL_PAYLOAD = 0x7fffffff;
constexpr int INSTANCE_BIT = 0x40000000;
constexpr int LAST_BIT = (int)0x80000000u;
constexpr int STACK_SIZE = 64; // the reference's own limit: BVHNode* stack[64] (bvh.cpp:227)

struct __align__(16) DMaterial { // rt_material padded to 48 B: one material = three 128-bit loads
    float reflectivity, refractivity;
    float absorption[3];
    float albedo[3];
    int is_light;
    int texture;
    int pad[2];
};

struct DTexture {
    const uint32_t* pixels;
    int width, height;
};

struct DScene {
    const float4* nodes;
    const float4* tris;
    const float4* inst;
    const float4* shade;
    const float4* inst_shade;
    const int* obj_material;
    const DMaterial* materials;
    const DTexture* textures;
    int root_ref;
    int kind;          // RT_SCENE_FLAT / RT_SCENE_TLAS / RT_SCENE_FLAT_KDTREE / RT_SCENE_FLAT_GRID
    int flat_obj_idx;  // objIdx override for the flat BVH (-1: take the triangle's)
    int skydome_texture, floor_texture;
    float floor_n[3], floor_d, floor_invto;
    float light_T[16], light_inv_T[16], light_size;
    float light_color[3], light_pos[3];
    // the other two FileScene accelerators (layouts documented next to their traversals below)
    const float4* kd_nodes;    // KD-tree kinds: 32-byte nodes of every tree (flat scene: node 0 = root)
    const int2* grid_cells;    // grid kinds: (first triangle slot, count) per cell, all grids concatenated
    const float4* grid_params; // grid kinds: 64 B per grid = (int res.xyz, int first cell) (cellSize.xyz, -) (boundsMin.xyz, -) (boundsMax.xyz, -)
};

// which accelerator a kernel is compiled for (template parameter, so the BVH kernels carry no extra code):
// the values are the scene kinds of include/rt_b200.h
enum { ACCEL_BVH = 0 /* flat BVH and TLAS over BVHs */, ACCEL_KD = 2, ACCEL_GRID = 3, ACCEL_TLAS_KD = 4, ACCEL_TLAS_GRID = 5 };
__host__ __device__ __forceinline__ bool kind_is_tlas(int kind) { return kind == 1 || kind == ACCEL_TLAS_KD || kind == ACCEL_TLAS_GRID; }

// ------------------------------------------------------------------------------------------------
// float3 helpers, written so that each reference expression maps 1:1 (template/tmplmath.h)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   // tmplmath.h:458
__device__ __forceinline__ float3 cross(float3 a, float3 b)                                               // tmplmath.h:512
{
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float3 normalize(float3 v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; } // :124,:480
__device__ __forceinline__ float3 reflect(float3 i, float3 n) { return i - 2.0f * n * dot(n, i); }        // tmplmath.h:506
__device__ __forceinline__ float3 recip(float3 d) { return f3(1 / d.x, 1 / d.y, 1 / d.z); }              // ray.h:19

// std::min / std::max exactly as the host compiler evaluates them in the slab test: a NaN in the
// FIRST operand propagates, a NaN in the second is dropped (CUDA's fminf/fmaxf drop either).
__device__ __forceinline__ float smin(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float smax(float a, float b) { return (a < b) ? b : a; }
// template/tmplmath.h:122-123,435
__device__ __forceinline__ float tfminf(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float tfmaxf(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float clampf(float f, float a, float b) { return tfmaxf(a, tfminf(f, b)); }
__device__ __forceinline__ int clampi(int f, int a, int b) { int m = f < b ? f : b; return a > m ? a : m; }

#define RT_PI 3.14159265358979323846264f
#define RT_INVPI 0.31830988618379067153777f
#define RT_INV2PI 0.15915494309189533576888f

// RNG: template/tmplmath.cpp:5-34
__device__ __forceinline__ uint32_t wang_hash(uint32_t s)
{
    s = (s ^ 61) ^ (s >> 16);
    s *= 9, s = s ^ (s >> 4);
    s *= 0x27d4eb2d;
    s = s ^ (s >> 15);
    return s;
}
__device__ __forceinline__ uint32_t init_seed(uint32_t base) { return wang_hash((base + 1) * 17); }
__device__ __forceinline__ float random_float(uint32_t& s)
{
    s ^= s << 13, s ^= s >> 17, s ^= s << 5;
    return s * 2.3283064365387e-10f;
}

// ------------------------------------------------------------------------------------------------
// hit record carried through traversal (Ray::t, barycentric, objIdx, triIdx, traversed, tested)
// ------------------------------------------------------------------------------------------------
struct HitRec {
    float t, u, v;
    int obj, tri;
    int traversed, tested;
};

// A ray needs the NaN-exact slab path iff a direction component is zero (rD = inf, so
// (b - O) * rD can be 0 * inf) or something is non-finite.  Everything else cannot produce NaN
// and fminf/fmaxf give the same values as std::min/max there.
__device__ __forceinline__ bool needs_exact_slab(float3 O, float3 D)
{
    const bool dirOk = fabsf(D.x) > 0 && fabsf(D.y) > 0 && fabsf(D.z) > 0 && fabsf(D.x) < 3e38f && fabsf(D.y) < 3e38f && fabsf(D.z) < 3e38f;
    const bool orgOk = fabsf(O.x) < 3e38f && fabsf(O.y) < 3e38f && fabsf(O.z) < 3e38f;
    return !(dirOk && orgOk);
}

// slab test: bvh.cpp:181-190 = blas_bvh.cpp:259-268 = tlas_bvh.cpp:72-81.  Returns tmin or 1e30f.
__device__ __forceinline__ float slab(float3 O, float3 rD, float rayT, bool exact,
    float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz)
{
    const float tx1 = (bminx - O.x) * rD.x, tx2 = (bmaxx - O.x) * rD.x;
    const float ty1 = (bminy - O.y) * rD.y, ty2 = (bmaxy - O.y) * rD.y;
    const float tz1 = (bminz - O.z) * rD.z, tz2 = (bmaxz - O.z) * rD.z;
    float tmin, tmax;
    if (!exact)
    {
        tmin = fminf(tx1, tx2), tmax = fmaxf(tx1, tx2);
        tmin = fmaxf(tmin, fminf(ty1, ty2)), tmax = fminf(tmax, fmaxf(ty1, ty2));
        tmin = fmaxf(tmin, fminf(tz1, tz2)), tmax = fminf(tmax, fmaxf(tz1, tz2));
    }
    else
    {
        tmin = smin(tx1, tx2), tmax = smax(tx1, tx2);
        tmin = smax(tmin, smin(ty1, ty2)), tmax = smin(tmax, smax(ty1, ty2));
        tmin = smax(tmin, smin(tz1, tz2)), tmax = smin(tmax, smax(tz1, tz2));
    }
    return (tmax >= tmin && tmin < rayT && tmax > 0) ? tmin : 1e30f;
}

// ------------------------------------------------------------------------------------------------
// Packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2, PTX add/sub/mul.rn.f32x2).  Each half is an independent IEEE-754
// operation, round-to-nearest, denormals kept: the same bits as the scalar instruction, at half the issue slots.
// ------------------------------------------------------------------------------------------------
typedef unsigned long long u64;
__device__ __forceinline__ u64 f2pack(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void f2unpack(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 f2sub(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 f2mul(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// the ray constants of the slab test, paired the way the node layout pairs the box planes
struct RaySlab {
    u64 Oxy, Ozz, Rxy, Rzz;
};
__device__ __forceinline__ RaySlab make_ray_slab(float3 O, float3 rD)
{
    RaySlab r;
    r.Oxy = f2pack(O.x, O.y), r.Ozz = f2pack(O.z, O.z), r.Rxy = f2pack(rD.x, rD.y), r.Rzz = f2pack(rD.z, rD.z);
    return r;
}

struct FatNode {
    ulonglong2 a, b, c; // n0, n1, n2 as pairs of packed halves
    int left, right;
};
__device__ __forceinline__ FatNode load_node(const float4* __restrict__ nodes, int cur)
{
    const ulonglong2* nd = (const ulonglong2*)(nodes + 4 * (size_t)cur);
    FatNode n;
    n.a = __ldg(nd), n.b = __ldg(nd + 1), n.c = __ldg(nd + 2);
    const int2 ch = __ldg((const int2*)(nd + 3));
    n.left = ch.x, n.right = ch.y;
    return n;
}

// min / max and the hit decision of one box from its six plane distances (second half of slab() above)
template <bool EXACT>
__device__ __forceinline__ float slab_finish(float tx1, float tx2, float ty1, float ty2, float tz1, float tz2, float rayT)
{
    float tmin, tmax;
    if (!EXACT)
    {
        tmin = fminf(tx1, tx2), tmax = fmaxf(tx1, tx2);
        tmin = fmaxf(tmin, fminf(ty1, ty2)), tmax = fminf(tmax, fmaxf(ty1, ty2));
        tmin = fmaxf(tmin, fminf(tz1, tz2)), tmax = fminf(tmax, fmaxf(tz1, tz2));
    }
    else
    {
        tmin = smin(tx1, tx2), tmax = smax(tx1, tx2);
        tmin = smax(tmin, smin(ty1, ty2)), tmax = smin(tmax, smax(ty1, ty2));
        tmin = smax(tmin, smin(tz1, tz2)), tmax = smin(tmax, smax(tz1, tz2));
    }
    // (tmax >= tmin && tmin < rayT && tmax > 0) ? tmin : 1e30f as one predicate chain + one select (the compiler's own
    // translation is a chain of three selects per box); comparisons with a NaN are false, as in C
    float d;
    asm("{ .reg .pred p;\n\t"
        "setp.ge.f32 p, %2, %1;\n\t"
        "setp.lt.and.f32 p, %1, %3, p;\n\t"
        "setp.gt.and.f32 p, %2, 0f00000000, p;\n\t"
        "selp.f32 %0, %1, 0f7149F2CA, p; }" // 0f7149F2CA = 1e30f
        : "=f"(d) : "f"(tmin), "f"(tmax), "f"(rayT));
    return d;
}

// both slab tests of a fat node (bvh.cpp:244-247 calls IntersectAABB on child1, then child2): d1 = left, d2 = right
template <bool EXACT>
__device__ __forceinline__ void slab_both(const RaySlab& r, const float rayT, const FatNode& n, float& d1, float& d2)
{
    float ltx1, lty1, ltx2, lty2, rtx1, rty1, rtx2, rty2, ltz1, ltz2, rtz1, rtz2;
    f2unpack(f2mul(f2sub(n.a.x, r.Oxy), r.Rxy), ltx1, lty1);
    f2unpack(f2mul(f2sub(n.a.y, r.Oxy), r.Rxy), ltx2, lty2);
    f2unpack(f2mul(f2sub(n.b.x, r.Oxy), r.Rxy), rtx1, rty1);
    f2unpack(f2mul(f2sub(n.b.y, r.Oxy), r.Rxy), rtx2, rty2);
    f2unpack(f2mul(f2sub(n.c.x, r.Ozz), r.Rzz), ltz1, ltz2);
    f2unpack(f2mul(f2sub(n.c.y, r.Ozz), r.Rzz), rtz1, rtz2);
    d1 = slab_finish<EXACT>(ltx1, ltx2, lty1, lty2, ltz1, ltz2, rayT);
    d2 = slab_finish<EXACT>(rtx1, rtx2, rty1, rty2, rtz1, rtz2, rayT);
}
__device__ __forceinline__ void slab_both(const RaySlab& r, const float rayT, const bool exact, const FatNode& n, float& d1, float& d2)
{
    if (exact) slab_both<true>(r, rayT, n, d1, d2);
    else slab_both<false>(r, rayT, n, d1, d2);
}

// Moeller-Trumbore: bvh.cpp:203-222 / blas_bvh.cpp:281-300.  Returns true when the hit was accepted.
__device__ __forceinline__ bool intersect_tri(float3 O, float3 D, float3 v0, float3 edge1, float3 edge2,
    float& rayT, float& outU, float& outV)
{
    const float3 h = cross(D, edge2);
    const float a = dot(edge1, h);
    if (a > -0.0001f && a < 0.0001f) return false;
    const float f = 1 / a;
    const float3 s = O - v0;
    const float u = f * dot(s, h);
    if (u < 0 || u > 1) return false;
    const float3 q = cross(s, edge1);
    const float v = f * dot(D, q);
    if (v < 0 || u + v > 1) return false;
    const float t = f * dot(edge2, q);
    if (t > 0.0001f && t < rayT)
    {
        rayT = t, outU = u, outV = v;
        return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// Two-level ordered traversal.
//   closest hit: bvh.cpp:224-258 (flat), tlas_bvh.cpp:83-111 + blas_bvh.cpp:376-389 + :302-336 (TLAS).
//   Same visiting order as the reference: near child first by slab tmin, left first on ties, far
//   child pushed only when hit, miss sentinel 1e30f compared with ==; the root box is never tested.
//   One stack serves both levels: entering an instance pushes SENTINEL, so that the instance's whole
//   subtree is finished (as the reference's nested call does) before the TLAS stack continues.
//   ANYHIT: return at the first accepted triangle (IsOccluded, file_scene.cpp:177-187: the reference
//   runs the closest-hit traversal with t = 1e34 and only asks whether anything was hit).
// ------------------------------------------------------------------------------------------------
template <bool ANYHIT, bool COUNTERS>
__device__ __forceinline__ void traverse(const DScene& s, const float3 wO, const float3 wD, HitRec& hit)
{
    float3 O = wO, D = wD;
    bool exact = needs_exact_slab(O, D);
    RaySlab rs = make_ray_slab(O, recip(D));
    int instObj = s.flat_obj_idx;
    int stack[STACK_SIZE];
    int sp = 0;
    int cur = s.root_ref;
    const float4* __restrict__ nodes = s.nodes;
    const float4* __restrict__ tris = s.tris;
    while (true)
    {
        if (cur >= 0)
        {
            if (COUNTERS) hit.traversed++;
            const FatNode n = load_node(nodes, cur);
            float d1, d2;
            slab_both(rs, hit.t, exact, n, d1, d2);
            int c1 = n.left, c2 = n.right;
            if (d1 > d2) { const float tf = d1; d1 = d2; d2 = tf; const int tc = c1; c1 = c2; c2 = tc; }
            if (d1 == 1e30f)
            {
                if (sp == 0) break;
                cur = stack[--sp];
            }
            else
            {
                cur = c1;
                if (d2 != 1e30f) stack[sp++] = c2;
            }
            continue;
        }
        const int payload = ~cur;
        if (payload == SENTINEL_PAYLOAD)
        {
            // leave the instance: blas_bvh.cpp:385-388 restores O, D, rD
            O = wO, D = wD;
            exact = needs_exact_slab(O, D);
            rs = make_ray_slab(O, recip(D));
        }
        else if (payload & INSTANCE_BIT)
        {
            // TLAS leaf -> BLASBVH::Intersect (blas_bvh.cpp:376-389), SSE lane-sum order of
            // TransformPosition_SSE / TransformVector_SSE (tmplmath.cpp:170-191)
            if (COUNTERS) hit.traversed++, hit.tested = 0;
            const float4* I = s.inst + 4 * (size_t)(payload & ~INSTANCE_BIT);
            const float4 r0 = __ldg(I), r1 = __ldg(I + 1), r2 = __ldg(I + 2);
            const int4 meta = __ldg((const int4*)(I + 3));
            O = f3((wO.x * r0.x + wO.y * r0.y) + (wO.z * r0.z + r0.w),
                   (wO.x * r1.x + wO.y * r1.y) + (wO.z * r1.z + r1.w),
                   (wO.x * r2.x + wO.y * r2.y) + (wO.z * r2.z + r2.w));
            D = f3((wD.x * r0.x + wD.y * r0.y) + wD.z * r0.z,
                   (wD.x * r1.x + wD.y * r1.y) + wD.z * r1.z,
                   (wD.x * r2.x + wD.y * r2.y) + wD.z * r2.z);
            exact = needs_exact_slab(O, D);
            rs = make_ray_slab(O, recip(D));
            instObj = meta.y;
            stack[sp++] = ~SENTINEL_PAYLOAD;
            cur = meta.x;
            continue;
        }
        else
        {
            // triangle leaf: bvh.cpp:232-241
            if (COUNTERS) hit.traversed++;
            int slot = payload;
            while (true)
            {
                const float4* T = tris + 3 * (size_t)slot;
                const float4 t0 = __ldg(T), t1 = __ldg(T + 1), t2 = __ldg(T + 2);
                const int tag = __float_as_int(t0.w);
                if (COUNTERS) hit.tested++;
                if (intersect_tri(O, D, f3(t0.x, t0.y, t0.z), f3(t1.x, t1.y, t1.z), f3(t2.x, t2.y, t2.z), hit.t, hit.u, hit.v))
                {
                    hit.tri = tag & ~LAST_BIT;
                    hit.obj = instObj >= 0 ? instObj : __float_as_int(t1.w);
                    if (ANYHIT) return;
                }
                if (tag & LAST_BIT) break;
                slot++;
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
}

// ------------------------------------------------------------------------------------------------
// analytic primitives tested for every ray before the BVH (file_scene.cpp:172-173)
// ------------------------------------------------------------------------------------------------
// Quad::Intersect / IsOccluded share this test: primitives.h:331-362
__device__ __forceinline__ bool quad_test(const DScene& s, float3 O, float3 D, float rayT, float& tOut)
{
    const float* m = s.light_inv_T;
    const float Oy = m[4] * O.x + m[5] * O.y + m[6] * O.z + m[7];
    const float Dy = m[4] * D.x + m[5] * D.y + m[6] * D.z;
    const float t = Oy / -Dy;
    if (t < rayT && t > 0)
    {
        const float Ox = m[0] * O.x + m[1] * O.y + m[2] * O.z + m[3];
        const float Oz = m[8] * O.x + m[9] * O.y + m[10] * O.z + m[11];
        const float Dx = m[0] * D.x + m[1] * D.y + m[2] * D.z;
        const float Dz = m[8] * D.x + m[9] * D.y + m[10] * D.z;
        const float Ix = Ox + t * Dx, Iz = Oz + t * Dz;
        const float size = s.light_size;
        if (Ix > -size && Ix < size && Iz > -size && Iz < size) { tOut = t; return true; }
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// Persistent-warp traversal with dynamic ray replacement.
//
// The plain per-thread loop above leaves a warp running until its slowest ray is done: on the path
// tracer's mixed queues ncu measured 7.3 of 32 lanes active per issued instruction (profiles/r1_v1_*).
// Here every lane runs a small state machine (one node visit, leaf, instance entry or instance exit
// per step) and a warp that has REFILL_LANES or more idle lanes pulls that many new rays from the
// queue with ONE atomicAdd (ballot / popc / shfl), so lanes stay occupied while the queue lasts.
// The per-ray visiting order is exactly the one of traverse<> (and the reference): only which ray
// occupies which lane changes.
//
// Src supplies the rays and takes the results:
//     bool load(int i, float3& O, float3& D, float& tmax)      ray i of the queue
//     void world(int i, float3& O, float3& D)                  reload of the world-space ray (instance exit)
//     void store(int i, const HitRec& h)
// ------------------------------------------------------------------------------------------------
constexpr int REFILL_LANES = 8;

template <bool ANYHIT, bool COUNTERS, class Src>
__device__ __forceinline__ void trace_queue(const DScene& s, Src& src, const int n, int* __restrict__ fetchCounter)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const float4* __restrict__ nodes = s.nodes;
    const float4* __restrict__ tris = s.tris;
    int stack[STACK_SIZE];
    int sp = 0, cur = 0, rayIdx = -1, instObj = s.flat_obj_idx;
    float3 O = f3(0, 0, 0), D = f3(0, 0, 0);
    RaySlab rs = make_ray_slab(O, O);
    bool exact = false, has = false, queueEmpty = false;
    HitRec hit;
    hit.t = 0, hit.u = 0, hit.v = 0, hit.obj = -1, hit.tri = -1, hit.traversed = 0, hit.tested = 0;
    while (true)
    {
        const unsigned idle = __ballot_sync(FULL, !has);
        if (idle == FULL && queueEmpty) break;
        if (!queueEmpty && (idle == FULL || __popc(idle) >= REFILL_LANES))
        {
            const int nIdle = __popc(idle);
            const int leader = __ffs(idle) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(fetchCounter, nIdle);
            base = __shfl_sync(FULL, base, leader);
            if (base + nIdle >= n) queueEmpty = true;
            if (!has)
            {
                const int i = base + __popc(idle & ((1u << lane) - 1));
                float tmax;
                if (i < n && src.load(i, O, D, tmax))
                {
                    rayIdx = i, has = true;
                    // FindNearest prologue: light quad, floor plane (file_scene.cpp:172-173); IsOccluded: quad only
                    hit.t = tmax, hit.u = 0, hit.v = 0, hit.obj = -1, hit.tri = -1, hit.traversed = 0, hit.tested = 0;
                    float tq;
                    bool done = false;
                    if (ANYHIT)
                    {
                        if (quad_test(s, O, D, tmax, tq)) hit.obj = 0, done = true;
                        hit.t = 1e34f;
                    }
                    else
                    {
                        if (quad_test(s, O, D, hit.t, tq)) hit.t = tq, hit.obj = 0;
                        const float3 N = f3(s.floor_n[0], s.floor_n[1], s.floor_n[2]);
                        const float tp = -(dot(O, N) + s.floor_d) / (dot(D, N));
                        if (tp < hit.t && tp > 0) hit.t = tp, hit.obj = 1;
                    }
                    if (done) { src.store(rayIdx, hit); has = false; }
                    else
                    {
                        rs = make_ray_slab(O, recip(D)), exact = needs_exact_slab(O, D);
                        sp = 0, cur = s.root_ref, instObj = s.flat_obj_idx;
                    }
                }
            }
            continue;
        }
        if (!has) continue;
        // ---- one traversal step ----
        bool pop = false;
        if (cur >= 0)
        {
            if (COUNTERS) hit.traversed++;
            const FatNode nd = load_node(nodes, cur);
            float d1, d2;
            slab_both(rs, hit.t, exact, nd, d1, d2);
            int c1 = nd.left, c2 = nd.right;
            if (d1 > d2) { const float tf = d1; d1 = d2; d2 = tf; const int tc = c1; c1 = c2; c2 = tc; }
            if (d1 == 1e30f) pop = true;
            else
            {
                cur = c1;
                if (d2 != 1e30f) stack[sp++] = c2;
            }
        }
        else
        {
            const int payload = ~cur;
            pop = true;
            if (payload == SENTINEL_PAYLOAD)
            {
                src.world(rayIdx, O, D); // blas_bvh.cpp:385-388
                rs = make_ray_slab(O, recip(D)), exact = needs_exact_slab(O, D);
            }
            else if (payload & INSTANCE_BIT)
            {
                if (COUNTERS) hit.traversed++, hit.tested = 0;
                const float4* I = s.inst + 4 * (size_t)(payload & ~INSTANCE_BIT);
                const float4 r0 = __ldg(I), r1 = __ldg(I + 1), r2 = __ldg(I + 2);
                const int4 meta = __ldg((const int4*)(I + 3));
                const float3 wO = O, wD = D; // in world space here: instances do not nest
                O = f3((wO.x * r0.x + wO.y * r0.y) + (wO.z * r0.z + r0.w),
                       (wO.x * r1.x + wO.y * r1.y) + (wO.z * r1.z + r1.w),
                       (wO.x * r2.x + wO.y * r2.y) + (wO.z * r2.z + r2.w));
                D = f3((wD.x * r0.x + wD.y * r0.y) + wD.z * r0.z,
                       (wD.x * r1.x + wD.y * r1.y) + wD.z * r1.z,
                       (wD.x * r2.x + wD.y * r2.y) + wD.z * r2.z);
                rs = make_ray_slab(O, recip(D)), exact = needs_exact_slab(O, D);
                instObj = meta.y;
                stack[sp++] = ~SENTINEL_PAYLOAD;
                cur = meta.x;
                pop = false;
            }
            else
            {
                if (COUNTERS) hit.traversed++;
                int slot = payload;
                while (true)
                {
                    const float4* T = tris + 3 * (size_t)slot;
                    const float4 t0 = __ldg(T), t1 = __ldg(T + 1), t2 = __ldg(T + 2);
                    const int tag = __float_as_int(t0.w);
                    if (COUNTERS) hit.tested++;
                    if (intersect_tri(O, D, f3(t0.x, t0.y, t0.z), f3(t1.x, t1.y, t1.z), f3(t2.x, t2.y, t2.z), hit.t, hit.u, hit.v))
                    {
                        hit.tri = tag & ~LAST_BIT;
                        hit.obj = instObj >= 0 ? instObj : __float_as_int(t1.w);
                        if (ANYHIT) { sp = 0; break; }
                    }
                    if (tag & LAST_BIT) break;
                    slot++;
                }
            }
        }
        if (pop)
        {
            if (sp == 0)
            {
                src.store(rayIdx, hit);
                has = false;
            }
            else cur = stack[--sp];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// KD-tree (FileScene as the reference ships it, file_scene.h:10-12).
//   Device node, 32 B = 2 x float4 (built by rt_scene.cu from rt_kd_node, children made adjacent):
//     k0 = (min.x, min.y, min.z, max.x)
//     k1 = (max.y, max.z, int a, b)    interior: a = left << 2 | axis (right = left + 1), b = float split plane
//                                      leaf:     a = ~first triangle slot,               b = int triangle count
//   Leaf triangles are copies in leaf order in `tris` (same 48-byte records as the BVH; a triangle that
//   straddles split planes is stored once per leaf that lists it), so a leaf is one contiguous run.
//   Traversal: KDTree::IntersectKDTree (kdtree.cpp:148-209) with the recursion unrolled onto a stack of
//   (far child, split t): the reference returns from a node without visiting the far child when
//   `ray.t < t` after the near child (:184, :203) - that test is made when the entry is popped, with the
//   ray.t of that moment, exactly as the recursion does.  Every node's own box is slab-tested on entry
//   (:151-152, tmin / tmax feed the child choice), the split distance is a true division (:172), and the
//   `tmin + 0.001` / `tmax - 0.001` comparisons are carried out in double like the reference's literals.
// ------------------------------------------------------------------------------------------------
constexpr int KD_STACK_SIZE = 32; // rt_scene_create rejects trees deeper than this (the reference stops at depth 20)

__device__ __forceinline__ bool slab_range(float3 O, float3 rD, float rayT, bool exact,
    float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz, float& tminOut, float& tmaxOut)
{
    const float tx1 = (bminx - O.x) * rD.x, tx2 = (bmaxx - O.x) * rD.x;
    const float ty1 = (bminy - O.y) * rD.y, ty2 = (bmaxy - O.y) * rD.y;
    const float tz1 = (bminz - O.z) * rD.z, tz2 = (bmaxz - O.z) * rD.z;
    float tmin, tmax;
    if (!exact)
    {
        tmin = fminf(tx1, tx2), tmax = fmaxf(tx1, tx2);
        tmin = fmaxf(tmin, fminf(ty1, ty2)), tmax = fminf(tmax, fmaxf(ty1, ty2));
        tmin = fmaxf(tmin, fminf(tz1, tz2)), tmax = fminf(tmax, fmaxf(tz1, tz2));
    }
    else
    {
        tmin = smin(tx1, tx2), tmax = smax(tx1, tx2);
        tmin = smax(tmin, smin(ty1, ty2)), tmax = smin(tmax, smax(ty1, ty2));
        tmin = smax(tmin, smin(tz1, tz2)), tmax = smin(tmax, smax(tz1, tz2));
    }
    tminOut = tmin, tmaxOut = tmax;
    return tmax >= tmin && tmin < rayT && tmax > 0;
}

__device__ __forceinline__ float axis_of(float3 a, int axis) { return axis == 0 ? a.x : (axis == 1 ? a.y : a.z); }

// Cursor protocol shared by the KD-tree and the grid.  A traversal alternates between ADVANCING through the structure
// (one node / one cell per step) and testing a RUN of consecutive triangle records (a leaf / a cell); the two are
// separate steps so that a warp can vote them as separate actions: ncu on a first version that tested the run inside
// the node step showed the triangle code at 2 of 32 lanes and 70 % of all instructions
// (profiles/r1_k_pt_streams_alt_*_source_hotspots.txt).
//   start()      position on the root / the entry cell; true = the ray is finished (misses the structure)
//   step()       CUR_CONTINUE another step() is due, CUR_RUN a run (runSlot, runCount > 0) is pending, CUR_DONE finished
//   tri_step()   test ONE triangle of the pending run (kdtree.cpp:155-161, grid.cpp:124-137); same return codes, CUR_RUN
//                while triangles remain; after the last one the cursor continues the way the reference does after its
//                leaf loop (KD-tree: return to the parent's pending far child; grid: advance to the next cell)
enum { CUR_CONTINUE = 0, CUR_DONE = 1, CUR_RUN = 2 };

template <bool COUNTERS>
__device__ __forceinline__ bool test_one(const float4* __restrict__ tris, int slot, int ownObj, float3 O, float3 D, HitRec& hit)
{
    const float4* T = tris + 3 * (size_t)slot;
    const float4 t0 = __ldg(T), t1 = __ldg(T + 1), t2 = __ldg(T + 2);
    if (COUNTERS) hit.tested++;
    if (intersect_tri(O, D, f3(t0.x, t0.y, t0.z), f3(t1.x, t1.y, t1.z), f3(t2.x, t2.y, t2.z), hit.t, hit.u, hit.v))
    {
        hit.tri = __float_as_int(t0.w) & ~LAST_BIT;
        hit.obj = ownObj >= 0 ? ownObj : __float_as_int(t1.w); // BLAS' objIdx (blas_kdtree.cpp:333, blas_grid.cpp:185) or Tri::objIdx
        return true;
    }
    return false;
}

struct KdCursor {
    int cur, sp;
    int runSlot, runCount;
    int ownObj; // -1: KDTree of a flat scene; >= 0: BLASKDTree with this objIdx
    float3 rD;
    bool exact;
    int stackNode[KD_STACK_SIZE];
    float stackT[KD_STACK_SIZE];

    // ref = index of the tree's root in kd_nodes
    __device__ __forceinline__ bool start(const DScene&, const float3 O, const float3 D, const HitRec&, int ref, int own)
    {
        rD = recip(D), exact = needs_exact_slab(O, D), cur = ref, sp = 0, runCount = 0, ownObj = own;
        return false;
    }

    // the recursion's unwinding: skip the pending far children the reference returns past (`if (ray.t < t) return;`, :189 / :208).
    // (Tried: (node, t) packed into 8-byte entries plus a running minimum of the pending t so that "the hit lies before
    // every pending plane" ends the ray without walking the stack - no faster, profiles/r1_kdtree_grid_accelerators.txt.)
    __device__ __forceinline__ int pop(const HitRec& hit)
    {
        while (true)
        {
            if (sp == 0) return CUR_DONE;
            sp--;
            // BLASKDTree additionally requires that the current hit belongs to this BLAS (blas_kdtree.cpp:373,392)
            if (!((ownObj < 0 || hit.obj == ownObj) && hit.t < stackT[sp])) { cur = stackNode[sp]; return CUR_CONTINUE; }
        }
    }

    template <bool COUNTERS>
    __device__ __forceinline__ int step(const DScene& s, const float3 O, const float3 D, HitRec& hit)
    {
        if (COUNTERS) hit.traversed++;
        const float4* __restrict__ nodes = s.kd_nodes;
        const float4 k0 = __ldg(nodes + 2 * (size_t)cur), k1 = __ldg(nodes + 2 * (size_t)cur + 1);
        float tmin, tmax;
        if (slab_range(O, rD, hit.t, exact, k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, tmin, tmax))
        {
            const int a = __float_as_int(k1.z);
            if (a < 0)
            {
                runSlot = ~a, runCount = __float_as_int(k1.w);
                if (runCount > 0) return CUR_RUN;
            }
            else
            {
                const int axis = a & 3, left = a >> 2;
                const float Da = axis_of(D, axis);
                const float t = (k1.w - axis_of(O, axis)) / Da;        // kdtree.cpp:171-172
                const bool pos = Da > 0;
                const int nearC = pos ? left : left + 1, farC = pos ? left + 1 : left;
                if ((double)t < (double)tmin + 0.001) cur = farC;       // :177 / :196: only the far side is crossed
                else if ((double)t > (double)tmax - 0.001) cur = nearC; // :182 / :201
                else stackNode[sp] = farC, stackT[sp] = t, sp++, cur = nearC;
                return CUR_CONTINUE;
            }
        }
        return pop(hit);
    }

    template <bool ANYHIT, bool COUNTERS>
    __device__ __forceinline__ int tri_step(const DScene& s, const float3 O, const float3 D, HitRec& hit)
    {
        const bool accepted = test_one<COUNTERS>(s.tris, runSlot, ownObj, O, D, hit);
        if (ANYHIT && accepted) return CUR_DONE;
        runSlot++;
        if (--runCount > 0) return CUR_RUN;
        return pop(hit);
    }
};

// one thread per ray: the cursor driven to completion
template <class Cursor, bool ANYHIT, bool COUNTERS>
__device__ __forceinline__ void run_cursor(const DScene& s, const float3 O, const float3 D, HitRec& hit)
{
    Cursor c;
    if (c.start(s, O, D, hit, 0, -1)) return;
    while (true)
    {
        int r = c.template step<COUNTERS>(s, O, D, hit);
        while (r == CUR_RUN) r = c.template tri_step<ANYHIT, COUNTERS>(s, O, D, hit);
        if (r == CUR_DONE) return;
    }
}


// ------------------------------------------------------------------------------------------------
// Uniform grid: Grid::IntersectGrid (grid.cpp:94-153), 3D-DDA without mailboxing (grid.h:7).
//   Device layout: grid_cells[cell] = (first triangle slot, count), triangles copied in cell order into
//   `tris` (a triangle overlapping several cells is stored once per cell).
//   cvtt_x86: static_cast<int>(float) on the reference's x86 build (cvttss2si) yields INT_MIN for NaN and for
//   values outside int range; CUDA's conversion saturates, so the out-of-range case is restated.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cvtt_x86(float x)
{
    return (x >= 2147483648.0f || x < -2147483648.0f || x != x) ? (int)0x80000000u : (int)x;
}

// Cursor (protocol above): start() runs the bounds test and the DDA set-up (grid.cpp:96-120), step() fetches the
// current cell's run, tri_step() tests one triangle and, after the last one, advances the DDA (grid.cpp:139-152).
struct GridCursor {
    int cell[3], stp[3], exitc[3];
    int runSlot, runCount;
    int ownObj;                  // -1: Grid of a flat scene; >= 0: BLASGrid with this objIdx
    int resX, resXY, cellBase;   // cell index = cellBase + x + y * resX + z * resXY
    float deltaT[3], nextT[3];

    // ref = index of the grid's 64-byte parameter record
    __device__ __forceinline__ bool start(const DScene& s, const float3 O, const float3 D, const HitRec& hit, int ref, int own)
    {
        const float4* P = s.grid_params + 4 * (size_t)ref;
        const int4 g0 = __ldg((const int4*)P);
        const float4 g1 = __ldg(P + 1), g2 = __ldg(P + 2), g3 = __ldg(P + 3);
        const int res[3] = { g0.x, g0.y, g0.z };
        const float cs[3] = { g1.x, g1.y, g1.z }, bmin[3] = { g2.x, g2.y, g2.z };
        const float3 rD = recip(D);
        float tminU, tmaxU;
        runCount = 0, ownObj = own, resX = g0.x, resXY = g0.x * g0.y, cellBase = g0.w;
        if (!slab_range(O, rD, hit.t, needs_exact_slab(O, D), g2.x, g2.y, g2.z, g3.x, g3.y, g3.z, tminU, tmaxU)) return true;
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            const float rayOrigCell = axis_of(O, i) - bmin[i];
            cell[i] = clampi(cvtt_x86(floorf(rayOrigCell / cs[i])), 0, res[i] - 1);
            if (axis_of(D, i) < 0)
            {
                deltaT[i] = -cs[i] * axis_of(rD, i);
                nextT[i] = (cell[i] * cs[i] - rayOrigCell) * axis_of(rD, i);
                exitc[i] = -1, stp[i] = -1;
            }
            else
            {
                deltaT[i] = cs[i] * axis_of(rD, i);
                nextT[i] = ((cell[i] + 1) * cs[i] - rayOrigCell) * axis_of(rD, i);
                exitc[i] = res[i], stp[i] = 1;
            }
        }
        return false;
    }

    // grid.cpp:139-152: pick the axis of the nearest cell boundary, stop when the hit lies before it or the grid ends
    __device__ __forceinline__ int advance(const HitRec& hit)
    {
        // k = (x<y)<<2 | (x<z)<<1 | (y<z), map = {2,1,2,1,2,2,0,0}
        const bool xy = nextT[0] < nextT[1], xz = nextT[0] < nextT[2], yz = nextT[1] < nextT[2];
        const int axis = xy ? (xz ? 0 : 2) : (yz ? 1 : 2);
        // unrolled selects keep cell / nextT in registers (no dynamically indexed local arrays)
        const float nt = axis == 0 ? nextT[0] : (axis == 1 ? nextT[1] : nextT[2]);
        if (hit.t < nt) return CUR_DONE;
#pragma unroll
        for (int i = 0; i < 3; i++)
            if (axis == i) cell[i] += stp[i];
        const int ca = axis == 0 ? cell[0] : (axis == 1 ? cell[1] : cell[2]);
        const int ea = axis == 0 ? exitc[0] : (axis == 1 ? exitc[1] : exitc[2]);
        if (ca == ea) return CUR_DONE;
#pragma unroll
        for (int i = 0; i < 3; i++)
            if (axis == i) nextT[i] += deltaT[i];
        return CUR_CONTINUE;
    }

    template <bool COUNTERS>
    __device__ __forceinline__ int step(const DScene& s, const float3 O, const float3 D, HitRec& hit)
    {
        if (COUNTERS) hit.traversed++;
        const int2 c = __ldg(s.grid_cells + (unsigned)(cellBase + cell[0] + cell[1] * resX + cell[2] * resXY));
        runSlot = c.x, runCount = c.y;
        if (runCount > 0) return CUR_RUN;
        return advance(hit);
    }

    template <bool ANYHIT, bool COUNTERS>
    __device__ __forceinline__ int tri_step(const DScene& s, const float3 O, const float3 D, HitRec& hit)
    {
        const bool accepted = test_one<COUNTERS>(s.tris, runSlot, ownObj, O, D, hit);
        if (ANYHIT && accepted) return CUR_DONE;
        runSlot++;
        if (--runCount > 0) return CUR_RUN;
        return advance(hit);
    }
};


// ------------------------------------------------------------------------------------------------
// TLASFileScene over per-object KD-trees / grids (TLAS_USE_KDTree / TLAS_USE_Grid, tlas_file_scene.h:12-14).
// The top level is the same agglomerative BVH as for TLAS_USE_BVH (tlas_kdtree.cpp:83-110 = tlas_grid.cpp:83-110 =
// tlas_bvh.cpp:83-111), stored as the same fat nodes; a TLAS leaf enters BLASKDTree::Intersect (blas_kdtree.cpp:420-431) /
// BLASGrid::Intersect (blas_grid.cpp:232-246): world -> object transform in the SSE lane-sum order, the BLAS cursor run
// to completion, then the TLAS stack continues.  The cursor protocol is the one above; world-space (O, D) come from the
// caller, the object-space ray lives in the cursor.
// ------------------------------------------------------------------------------------------------
template <class Blas>
struct TlasCursor {
    Blas blas;
    float3 Ol, Dl;
    RaySlab rsW;
    bool exactW, inBlas;
    int cur, sp;
    int stack[STACK_SIZE];

    __device__ __forceinline__ bool start(const DScene& s, const float3 O, const float3 D, const HitRec&, int, int)
    {
        rsW = make_ray_slab(O, recip(D)), exactW = needs_exact_slab(O, D), cur = s.root_ref, sp = 0, inBlas = false;
        blas.runCount = 0;
        return false;
    }

    __device__ __forceinline__ int pop_tlas()
    {
        if (sp == 0) return CUR_DONE;
        cur = stack[--sp];
        return CUR_CONTINUE;
    }

    template <bool COUNTERS>
    __device__ __forceinline__ int step(const DScene& s, const float3 O, const float3 D, HitRec& hit)
    {
        if (inBlas)
        {
            const int r = blas.template step<COUNTERS>(s, Ol, Dl, hit);
            if (r != CUR_DONE) return r;
            inBlas = false; // blas_kdtree.cpp:427-430: O, D, rD restored, the hit record carried over
            return pop_tlas();
        }
        if (cur >= 0)
        {
            if (COUNTERS) hit.traversed++;
            const FatNode n = load_node(s.nodes, cur);
            float d1, d2;
            slab_both(rsW, hit.t, exactW, n, d1, d2);
            int c1 = n.left, c2 = n.right;
            if (d1 > d2) { const float tf = d1; d1 = d2; d2 = tf; const int tc = c1; c1 = c2; c2 = tc; }
            if (d1 == 1e30f) return pop_tlas();
            cur = c1;
            if (d2 != 1e30f) stack[sp++] = c2;
            return CUR_CONTINUE;
        }
        // TLAS leaf: enter the instance (Ray(const Ray&) drops `tested`, ray.h:10-14)
        if (COUNTERS) hit.traversed++, hit.tested = 0;
        const float4* I = s.inst + 4 * (size_t)(~cur & ~INSTANCE_BIT);
        const float4 r0 = __ldg(I), r1 = __ldg(I + 1), r2 = __ldg(I + 2);
        const int4 meta = __ldg((const int4*)(I + 3));
        Ol = f3((O.x * r0.x + O.y * r0.y) + (O.z * r0.z + r0.w),
                (O.x * r1.x + O.y * r1.y) + (O.z * r1.z + r1.w),
                (O.x * r2.x + O.y * r2.y) + (O.z * r2.z + r2.w));
        Dl = f3((D.x * r0.x + D.y * r0.y) + D.z * r0.z,
                (D.x * r1.x + D.y * r1.y) + D.z * r1.z,
                (D.x * r2.x + D.y * r2.y) + D.z * r2.z);
        if (blas.start(s, Ol, Dl, hit, meta.x, meta.y)) return pop_tlas();
        inBlas = true;
        return CUR_CONTINUE;
    }

    template <bool ANYHIT, bool COUNTERS>
    __device__ __forceinline__ int tri_step(const DScene& s, const float3 O, const float3 D, HitRec& hit)
    {
        const int r = blas.template tri_step<ANYHIT, COUNTERS>(s, Ol, Dl, hit);
        if (r != CUR_DONE) return r;
        if (ANYHIT && hit.obj > -1) return CUR_DONE; // occluded: the whole query ends, not just this instance
        inBlas = false;
        return pop_tlas();
    }
};

template <int ACCEL> struct CursorOf { typedef KdCursor type; };
template <> struct CursorOf<ACCEL_GRID> { typedef GridCursor type; };
template <> struct CursorOf<ACCEL_TLAS_KD> { typedef TlasCursor<KdCursor> type; };
template <> struct CursorOf<ACCEL_TLAS_GRID> { typedef TlasCursor<GridCursor> type; };

template <int ACCEL, bool ANYHIT, bool COUNTERS>
__device__ __forceinline__ void accel_traverse(const DScene& s, const float3 O, const float3 D, HitRec& hit)
{
    if (ACCEL == ACCEL_BVH) traverse<ANYHIT, COUNTERS>(s, O, D, hit);
    else run_cursor<typename CursorOf<ACCEL>::type, ANYHIT, COUNTERS>(s, O, D, hit);
}

// One interior-node visit of the ordered traversal (bvh.cpp:242-257) on register state: both child boxes from one
// 64-byte record, near child first, left on ties, far child pushed only when hit; selects instead of branches.
template <bool EXACT>
__device__ __forceinline__ void node_step(const float4* __restrict__ nodes, const RaySlab& rs, const float ht,
    int* stack, int& sp, int& cur, bool& end)
{
    const FatNode n = load_node(nodes, cur);
    float a1, a2;
    slab_both<EXACT>(rs, ht, n, a1, a2);
    const bool swp = a1 > a2;
    const float d1 = swp ? a2 : a1, d2 = swp ? a1 : a2;
    const int c1 = swp ? n.right : n.left, c2 = swp ? n.left : n.right;
    const bool miss = d1 == 1e30f, both = !miss && d2 != 1e30f;
    const int top = stack[sp > 0 ? sp - 1 : 0];
    if (both) stack[sp] = c2;
    end = miss && sp == 0;
    cur = miss ? top : c1;
    sp += both ? 1 : (miss && sp > 0 ? -1 : 0);
}

// trace_queue with a warp vote, the schedule of the path tracer's stream kernel applied to ray queues: every iteration the
// warp runs ONE action - interior-node visits (repeated while >= 3/4 of the lanes that entered are still on interior nodes)
// or leaf / instance steps - whichever more lanes wait for, instead of serialising both inside every iteration.
// Same per-ray visiting order, same refill rule, same Src interface as trace_queue<>.
template <bool ANYHIT, bool COUNTERS, class Src>
__device__ __forceinline__ void trace_queue_voted(const DScene& s, Src& src, const int n, int* __restrict__ fetchCounter)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const float4* __restrict__ nodes = s.nodes;
    const float4* __restrict__ tris = s.tris;
    int stack[STACK_SIZE];
    int sp = 0, cur = 0, rayIdx = -1, instObj = s.flat_obj_idx;
    float3 O = f3(0, 0, 0), D = f3(0, 0, 0);
    RaySlab rs = make_ray_slab(O, O);
    bool exact = false, has = false, queueEmpty = false;
    HitRec hit;
    hit.t = 0, hit.u = 0, hit.v = 0, hit.obj = -1, hit.tri = -1, hit.traversed = 0, hit.tested = 0;
    while (true)
    {
        const unsigned idle = __ballot_sync(FULL, !has);
        if (idle == FULL && queueEmpty) break;
        if (!queueEmpty && (idle == FULL || __popc(idle) >= REFILL_LANES))
        {
            const int nIdle = __popc(idle);
            const int leader = __ffs(idle) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(fetchCounter, nIdle);
            base = __shfl_sync(FULL, base, leader);
            if (base + nIdle >= n) queueEmpty = true;
            if (!has)
            {
                const int i = base + __popc(idle & ((1u << lane) - 1));
                float tmax;
                if (i < n && src.load(i, O, D, tmax))
                {
                    rayIdx = i, has = true;
                    // FindNearest prologue: light quad, floor plane (file_scene.cpp:172-173); IsOccluded: quad only
                    hit.t = tmax, hit.u = 0, hit.v = 0, hit.obj = -1, hit.tri = -1, hit.traversed = 0, hit.tested = 0;
                    float tq;
                    bool done = false;
                    if (ANYHIT)
                    {
                        if (quad_test(s, O, D, tmax, tq)) hit.obj = 0, done = true;
                        hit.t = 1e34f;
                    }
                    else
                    {
                        if (quad_test(s, O, D, hit.t, tq)) hit.t = tq, hit.obj = 0;
                        const float3 N = f3(s.floor_n[0], s.floor_n[1], s.floor_n[2]);
                        const float tp = -(dot(O, N) + s.floor_d) / (dot(D, N));
                        if (tp < hit.t && tp > 0) hit.t = tp, hit.obj = 1;
                    }
                    if (done) { src.store(rayIdx, hit); has = false; }
                    else
                    {
                        rs = make_ray_slab(O, recip(D)), exact = needs_exact_slab(O, D);
                        sp = 0, cur = s.root_ref, instObj = s.flat_obj_idx;
                    }
                }
            }
            continue;
        }
        const int nN = __popc(__ballot_sync(FULL, has && cur >= 0)), nL = __popc(__ballot_sync(FULL, has && cur < 0));
        if (nN >= nL)
        {
            const int keep = nN - (nN >> 2);
            bool inNode = has && cur >= 0;
            const bool anyExact = __any_sync(FULL, inNode && exact);
            do
            {
                if (inNode)
                {
                    if (COUNTERS) hit.traversed++;
                    bool end;
                    if (anyExact) node_step<true>(nodes, rs, hit.t, stack, sp, cur, end);
                    else node_step<false>(nodes, rs, hit.t, stack, sp, cur, end);
                    if (end) src.store(rayIdx, hit), has = false;
                    inNode = !end && cur >= 0;
                }
            } while (__popc(__ballot_sync(FULL, inNode)) >= keep);
        }
        else if (has && cur < 0)
        {
            const int payload = ~cur;
            bool pop = true;
            if (payload == SENTINEL_PAYLOAD)
            {
                src.world(rayIdx, O, D); // blas_bvh.cpp:385-388
                rs = make_ray_slab(O, recip(D)), exact = needs_exact_slab(O, D);
            }
            else if (payload & INSTANCE_BIT)
            {
                if (COUNTERS) hit.traversed++, hit.tested = 0;
                const float4* I = s.inst + 4 * (size_t)(payload & ~INSTANCE_BIT);
                const float4 r0 = __ldg(I), r1 = __ldg(I + 1), r2 = __ldg(I + 2);
                const int4 meta = __ldg((const int4*)(I + 3));
                const float3 wO = O, wD = D; // in world space here: instances do not nest
                O = f3((wO.x * r0.x + wO.y * r0.y) + (wO.z * r0.z + r0.w),
                       (wO.x * r1.x + wO.y * r1.y) + (wO.z * r1.z + r1.w),
                       (wO.x * r2.x + wO.y * r2.y) + (wO.z * r2.z + r2.w));
                D = f3((wD.x * r0.x + wD.y * r0.y) + wD.z * r0.z,
                       (wD.x * r1.x + wD.y * r1.y) + wD.z * r1.z,
                       (wD.x * r2.x + wD.y * r2.y) + wD.z * r2.z);
                rs = make_ray_slab(O, recip(D)), exact = needs_exact_slab(O, D);
                instObj = meta.y;
                stack[sp++] = ~SENTINEL_PAYLOAD;
                cur = meta.x;
                pop = false;
            }
            else
            {
                if (COUNTERS) hit.traversed++;
                int slot = payload;
                while (true)
                {
                    const float4* T = tris + 3 * (size_t)slot;
                    const float4 t0 = __ldg(T), t1 = __ldg(T + 1), t2 = __ldg(T + 2);
                    const int tag = __float_as_int(t0.w);
                    if (COUNTERS) hit.tested++;
                    if (intersect_tri(O, D, f3(t0.x, t0.y, t0.z), f3(t1.x, t1.y, t1.z), f3(t2.x, t2.y, t2.z), hit.t, hit.u, hit.v))
                    {
                        hit.tri = tag & ~LAST_BIT;
                        hit.obj = instObj >= 0 ? instObj : __float_as_int(t1.w);
                        if (ANYHIT) { sp = 0; break; }
                    }
                    if (tag & LAST_BIT) break;
                    slot++;
                }
            }
            if (pop)
            {
                if (sp == 0) src.store(rayIdx, hit), has = false;
                else cur = stack[--sp];
            }
        }
    }
}

// Persistent-warp traversal with ray replacement for the cursor accelerators: the KD-tree / grid counterpart of
// trace_queue<> above (same Src interface, same refill rule).  Lanes are IDLE, advancing (TRAV) or inside a triangle
// run (TRI); the warp executes the action most lanes wait for and repeats it while >= 3/4 of them stay in that state.
template <class Cursor, bool ANYHIT, bool COUNTERS, class Src>
__device__ __forceinline__ void trace_queue_cursor(const DScene& s, Src& src, const int n, int* __restrict__ fetchCounter)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    enum { Q_IDLE = 0, Q_TRAV = 1, Q_TRI = 2 };
    int state = Q_IDLE, rayIdx = -1;
    bool queueEmpty = false;
    float3 O = f3(0, 0, 0), D = f3(0, 0, 0);
    Cursor cursor;
    HitRec hit;
    hit.t = 0, hit.u = 0, hit.v = 0, hit.obj = -1, hit.tri = -1, hit.traversed = 0, hit.tested = 0;
    while (true)
    {
        const unsigned idle = __ballot_sync(FULL, state == Q_IDLE);
        if (idle == FULL && queueEmpty) break;
        if (!queueEmpty && (idle == FULL || __popc(idle) >= REFILL_LANES))
        {
            const int nIdle = __popc(idle);
            const int leader = __ffs(idle) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(fetchCounter, nIdle);
            base = __shfl_sync(FULL, base, leader);
            if (base + nIdle >= n) queueEmpty = true;
            if (state == Q_IDLE)
            {
                const int i = base + __popc(idle & ((1u << lane) - 1));
                float tmax;
                if (i < n && src.load(i, O, D, tmax))
                {
                    rayIdx = i;
                    // FindNearest prologue: light quad, floor plane (file_scene.cpp:172-173); IsOccluded: quad only
                    hit.t = tmax, hit.u = 0, hit.v = 0, hit.obj = -1, hit.tri = -1, hit.traversed = 0, hit.tested = 0;
                    float tq;
                    bool done = false;
                    if (ANYHIT)
                    {
                        if (quad_test(s, O, D, tmax, tq)) hit.obj = 0, done = true;
                        hit.t = 1e34f;
                    }
                    else
                    {
                        if (quad_test(s, O, D, hit.t, tq)) hit.t = tq, hit.obj = 0;
                        const float3 N = f3(s.floor_n[0], s.floor_n[1], s.floor_n[2]);
                        const float tp = -(dot(O, N) + s.floor_d) / (dot(D, N));
                        if (tp < hit.t && tp > 0) hit.t = tp, hit.obj = 1;
                    }
                    if (!done) done = cursor.start(s, O, D, hit, 0, -1);
                    if (done) src.store(rayIdx, hit);
                    else state = Q_TRAV;
                }
            }
            continue;
        }
        const int nT = __popc(__ballot_sync(FULL, state == Q_TRAV)), nR = __popc(__ballot_sync(FULL, state == Q_TRI));
        if (nT >= nR)
        {
            const int keep = (nT * 3 + 3) >> 2;
            do
            {
                if (state == Q_TRAV)
                {
                    const int r = cursor.template step<COUNTERS>(s, O, D, hit);
                    state = r == CUR_DONE ? Q_IDLE : (r == CUR_RUN ? Q_TRI : Q_TRAV);
                    if (r == CUR_DONE) src.store(rayIdx, hit);
                }
            } while (keep > 0 && __popc(__ballot_sync(FULL, state == Q_TRAV)) >= keep);
        }
        else
        {
            const int keep = (nR * 3 + 3) >> 2;
            do
            {
                if (state == Q_TRI)
                {
                    const int r = cursor.template tri_step<ANYHIT, COUNTERS>(s, O, D, hit);
                    state = r == CUR_DONE ? Q_IDLE : (r == CUR_RUN ? Q_TRI : Q_TRAV);
                    if (r == CUR_DONE) src.store(rayIdx, hit);
                }
            } while (__popc(__ballot_sync(FULL, state == Q_TRI)) >= keep);
        }
    }
}

// the persistent-warp queue traversal of the accelerator a kernel is compiled for (VOTED: trace_queue_voted for the BVH;
// measured +3..13 % on incoherent closest-hit queues and -5..17 % on coherent and any-hit ones, so it is opt-in:
// profiles/r1_ray_queue_voted_vs_plain.txt)
template <int ACCEL, bool ANYHIT, bool COUNTERS, class Src, bool VOTED = false>
__device__ __forceinline__ void accel_trace_queue(const DScene& s, Src& src, const int n, int* __restrict__ fetchCounter)
{
    if (ACCEL == ACCEL_BVH && VOTED) trace_queue_voted<ANYHIT, COUNTERS>(s, src, n, fetchCounter);
    else if (ACCEL == ACCEL_BVH) trace_queue<ANYHIT, COUNTERS>(s, src, n, fetchCounter);
    else trace_queue_cursor<typename CursorOf<ACCEL>::type, ANYHIT, COUNTERS>(s, src, n, fetchCounter);
}

// BaseScene::FindNearest: file_scene.cpp:170-175 = tlas_file_scene.cpp:201-206
template <bool COUNTERS, int ACCEL = ACCEL_BVH>
__device__ __forceinline__ void find_nearest(const DScene& s, float3 O, float3 D, float tmax, HitRec& hit)
{
    hit.t = tmax, hit.u = 0, hit.v = 0, hit.obj = -1, hit.tri = -1, hit.traversed = 0, hit.tested = 0;
    float tq;
    if (quad_test(s, O, D, hit.t, tq)) hit.t = tq, hit.obj = 0;
    {
        // Plane::Intersect primitives.h:107-111
        const float3 N = f3(s.floor_n[0], s.floor_n[1], s.floor_n[2]);
        const float t = -(dot(O, N) + s.floor_d) / (dot(D, N));
        if (t < hit.t && t > 0) hit.t = t, hit.obj = 1;
    }
    accel_traverse<ACCEL, false, COUNTERS>(s, O, D, hit);
}

// BaseScene::IsOccluded: file_scene.cpp:177-187 = tlas_file_scene.cpp:208-218 (SURVEY quirk Q2:
// geometry is tested with t = 1e34, the floor never occludes, the light quad uses the real t).
// Any-hit is exact for all three accelerators: the reference runs its closest-hit traversal and only asks
// whether objIdx was set, and until the first accepted triangle the visiting order is the same.
template <int ACCEL = ACCEL_BVH>
__device__ __forceinline__ bool is_occluded(const DScene& s, float3 O, float3 D, float tmax)
{
    float tq;
    if (quad_test(s, O, D, tmax, tq)) return true;
    HitRec hit;
    hit.t = 1e34f, hit.u = 0, hit.v = 0, hit.obj = -1, hit.tri = -1, hit.traversed = 0, hit.tested = 0;
    accel_traverse<ACCEL, true, false>(s, O, D, hit);
    return hit.obj > -1;
}

// ------------------------------------------------------------------------------------------------
// textures, sky, hit shading queries
// ------------------------------------------------------------------------------------------------
// Texture::Sample: texture.h:61-96 (nearest, clamp, v flip, 0x00RRGGBB), split into texel choice and fetch
__device__ __forceinline__ void texture_texel(const DTexture& T, float u, float v, int& x, int& y, float& fx, float& fy)
{
    u = clampf(u, 0.0f, 1.0f);
    v = 1 - clampf(v, 0.0f, 1.0f);
    fx = u * T.width, fy = v * T.height;
    x = clampi((int)fx, 0, T.width - 1);
    y = clampi((int)fy, 0, T.height - 1);
}

__device__ __forceinline__ float3 texture_fetch(const DTexture& T, int x, int y)
{
    const uint32_t pixel = __ldg(T.pixels + (x + y * T.width));
    const float rgbScale = 1 / 255.0f;
    return f3(((pixel >> 16) & 0xFF) * rgbScale, ((pixel >> 8) & 0xFF) * rgbScale, (pixel & 0xFF) * rgbScale);
}

__device__ __forceinline__ float3 texture_sample(const DScene& s, int tex, float u, float v)
{
    if (tex < 0) return f3(0, 0, 0);
    const DTexture T = s.textures[tex];
    if (T.width * T.height == 0) return f3(0, 0, 0);
    int x, y;
    float fx, fy;
    texture_texel(T, u, v, x, y, fx, fy);
    return texture_fetch(T, x, y);
}

// GetSkyColor: file_scene.cpp:142-154.  The texel must be the one the reference picks, i.e. the one glibc's atan2f / acosf lead to.
// Half of the bench scene's primary rays end here; running the restated glibc routines (two IEEE divisions and a square root more
// than CUDA's, no FMA contraction) for every lookup measured 5-20 % on the whole path tracer (profiles/r1_glibc_math.txt).
// Instead the texel is first computed with CUDA's atan2f / acosf and ACCEPTED only when u * width and v * height are further from
// a texel border than the two libraries can disagree; otherwise (about 4e-6 * (width + height) of the lookups, 2.5 % for a 4096 x 2048
// sky) the restated routines decide.  Bound: each library is within 4 ulp of the true angle (CUDA documents 2 ulp for both functions,
// fdlibm's routines stay below 2 including the y / x rounding), so phi and theta differ by at most 8 ulp(pi) = 1.9e-6, u = phi / 2pi
// by 3.0e-7 and v = theta / pi by 6.1e-7, plus 1.2e-7 for the two roundings of each product: 4.2e-7 * width and 7.3e-7 * height in texel
// units, plus one ulp of the product itself (1.2e-7 * size).  the v flip 1 - v adds another 1.2e-7 * height.
// RT_SKY_TEXEL_MARGIN = 2e-6 per unit of size is twice the larger of the two sums.
// NaN directions and the clamped ends (u * width = 0 or width) fail the test and take the exact path too.
#define RT_SKY_TEXEL_MARGIN 2e-6f
__device__ __forceinline__ void sky_texel_exact(const DTexture& T, float3 D, int& x, int& y)
{
    const float phi = rt_atan2f(-D.z, D.x) + RT_PI;
    const float theta = rt_acosf(-D.y);
    float fx, fy;
    texture_texel(T, phi * RT_INV2PI, theta * RT_INVPI, x, y, fx, fy);
}
// The same, out of line: ~2 % of the sky lookups get here, and inlined the two restated glibc routines (IEEE divisions, a square root,
// fdlibm's branch trees) cost the hot kernels 7 KB of SASS and registers on every path (round-1 A/B: +5 ms of 68 on the bench job,
// profiles/r2_libm_ab_*).  Returns x | y << 32.
static __device__ __noinline__ unsigned long long sky_texel_exact_cold(int width, int height, float dx, float dy, float dz)
{
    DTexture T;
    T.pixels = nullptr, T.width = width, T.height = height;
    int x, y;
    sky_texel_exact(T, f3(dx, dy, dz), x, y);
    return (unsigned long long)(unsigned)x | ((unsigned long long)(unsigned)y << 32);
}

__device__ __forceinline__ bool sky_texel_filtered(const DTexture& T, float3 D, int& x, int& y)
{
    const float phi = atan2f(-D.z, D.x) + RT_PI;
    const float theta = acosf(-D.y);
    float fx, fy;
    texture_texel(T, phi * RT_INV2PI, theta * RT_INVPI, x, y, fx, fy);
    return fabsf(fx - rintf(fx)) > RT_SKY_TEXEL_MARGIN * T.width && fabsf(fy - rintf(fy)) > RT_SKY_TEXEL_MARGIN * T.height;
}

__device__ __forceinline__ float3 sky_color(const DScene& s, float3 D)
{
    const int tex = s.skydome_texture;
    if (tex < 0) return f3(0, 0, 0);
    const DTexture T = s.textures[tex];
    if (T.width * T.height == 0) return f3(0, 0, 0);
    int x, y;
#if RT_B200_GLIBC_SKY
    if (!sky_texel_filtered(T, D, x, y))
    {
        const unsigned long long xy = sky_texel_exact_cold(T.width, T.height, D.x, D.y, D.z);
        x = (int)(unsigned)xy, y = (int)(xy >> 32);
    }
#else
    sky_texel_filtered(T, D, x, y); // CUDA's atan2f / acosf only (<= 2 ulp away: a lookup on a texel border can flip)
#endif
    return texture_fetch(T, x, y);
}

// Beer's law factors exp(-absorption * t) (3. PathTracer/renderer.cpp:76-80, 2. WhittedStyle/renderer.cpp:81-88).  The restated glibc
// expf works in double precision; only rays that travelled inside glass get here, so it is kept out of line and the shading code
// of the hot kernels carries no FP64 instructions.
#if RT_B200_GLIBC_EXPF
static __device__ __noinline__ float3 beer_scale(float ax, float ay, float az) { return f3(rt_expf(ax), rt_expf(ay), rt_expf(az)); }
#else
__device__ __forceinline__ float3 beer_scale(float ax, float ay, float az) { return f3(rt_expf(ax), rt_expf(ay), rt_expf(az)); }
#endif

struct ShadeHit {
    float3 N, albedo, absorption;
    float reflectivity, refractivity;
    bool isLight;
};

// GetHitInfo (file_scene.cpp:189-214 / tlas_file_scene.cpp:220-260) + Material::GetAlbedo (material.h:28-35)
__device__ __forceinline__ void hit_info(const DScene& s, float3 D, float3 I, int obj, int tri, float bu, float bv, ShadeHit& h, float& outU, float& outV)
{
    h.isLight = false, h.reflectivity = 0, h.refractivity = 0, h.absorption = f3(0, 0, 0);
    float u = 0, v = 0;
    int tex = -1;
    float3 matAlbedo = f3(1, 1, 1);
    if (obj == 0)
    {
        h.N = f3(-s.light_T[1], -s.light_T[5], -s.light_T[9]); // Quad::GetNormal primitives.h:363-367
        h.isLight = true;
    }
    else if (obj == 1)
    {
        h.N = f3(s.floor_n[0], s.floor_n[1], s.floor_n[2]);
        if (s.floor_n[1] == 1) // Plane::GetUV primitives.h:116-133
        {
            u = I.x, v = I.z;
            u *= s.floor_invto, v *= s.floor_invto;
            u = u - floorf(u), v = v - floorf(v);
        }
        tex = s.floor_texture;
    }
    else
    {
        int triBase = 0;
        float4 r0, r1, r2;
        const bool tlas = kind_is_tlas(s.kind);
        if (tlas)
        {
            const float4* IS = s.inst_shade + 4 * (size_t)(obj - 2);
            r0 = __ldg(IS), r1 = __ldg(IS + 1), r2 = __ldg(IS + 2);
            triBase = __float_as_int(__ldg(IS + 3).x);
        }
        const float4* S = s.shade + 4 * (size_t)(triBase + tri);
        const float4 s0 = __ldg(S), s1 = __ldg(S + 1), s2 = __ldg(S + 2), s3 = __ldg(S + 3);
        const float3 n0 = f3(s0.x, s0.y, s0.z), n1 = f3(s0.w, s1.x, s1.y), n2 = f3(s1.z, s1.w, s2.x);
        // GetNormal bvh.cpp:290-297 / blas_bvh.cpp:391-398
        float3 N = (1 - bu - bv) * n0 + bu * n1 + bv * n2;
        if (tlas) // TransformVector(N, T): float4(N, 0) * M, tmplmath.cpp:155-169
            N = f3(r0.x * N.x + r0.y * N.y + r0.z * N.z + r0.w * 0.0f,
                   r1.x * N.x + r1.y * N.y + r1.z * N.z + r1.w * 0.0f,
                   r2.x * N.x + r2.y * N.y + r2.z * N.z + r2.w * 0.0f);
        h.N = normalize(N);
        // GetUV bvh.cpp:299-305
        const float w = 1 - bu - bv;
        u = w * s2.y + bu * s2.w + bv * s3.y;
        v = w * s2.z + bu * s3.x + bv * s3.z;
        const int objIdx = tlas ? obj : __float_as_int(s3.w);
        const DMaterial m = s.materials[s.obj_material[objIdx - 2]];
        h.reflectivity = m.reflectivity, h.refractivity = m.refractivity;
        h.absorption = f3(m.absorption[0], m.absorption[1], m.absorption[2]);
        h.isLight = m.is_light != 0;
        matAlbedo = f3(m.albedo[0], m.albedo[1], m.albedo[2]);
        tex = m.texture;
    }
    if (dot(h.N, D) > 0) h.N = -h.N;
    h.albedo = tex < 0 ? matAlbedo : texture_sample(s, tex, u, v);
    outU = u, outV = v;
}

// Camera::GetPrimaryRay: camera.h:23-30
struct DCamera {
    float3 pos, topLeft, topRight, bottomLeft;
    float invW, invH; // 1.0f / SCRWIDTH, 1.0f / SCRHEIGHT
};

__device__ __forceinline__ float3 primary_dir(const DCamera& c, float x, float y)
{
    const float u = x * c.invW;
    const float v = y * c.invH;
    const float3 P = c.topLeft + u * (c.topRight - c.topLeft) + v * (c.bottomLeft - c.topLeft);
    return normalize(P - c.pos);
}

} // namespace rtb
